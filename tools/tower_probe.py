"""Times the non-DCN pieces of the KGDet head step on one B200 (CUDA events, 20 reps): towers in NCHW vs
channels_last, the stage-1 plain block, the 1x1 output convolutions.  Decides which of them are worth
replacing."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from kgdet_b200.head import KGDetHead  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return round(a.elapsed_time(b) / reps * 1e3, 1)


def main():
    torch.manual_seed(0)
    head = KGDetHead().cuda().eval()
    x = torch.randn(16, 256, 25, 42, device='cuda')
    out = {}
    with torch.no_grad():
        def towers(xx):
            c = p = xx
            for m in head.cls_convs:
                c = m(c)
            for m in head.reg_convs:
                p = m(p)
            return c, p
        out['towers_nchw_us'] = timed(lambda: towers(x))
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            towers(x)
        out['towers_nchw_graph_us'] = timed(g.replay)
        head_cl = KGDetHead().cuda().eval().to(memory_format=torch.channels_last)
        xcl = x.contiguous(memory_format=torch.channels_last)
        def towers_cl(xx):
            c = p = xx
            for m in head_cl.cls_convs:
                c = m(c)
            for m in head_cl.reg_convs:
                p = m(p)
            return c, p
        out['towers_channels_last_us'] = timed(lambda: towers_cl(xcl))
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2):
            r = towers_cl(xcl)
        out['towers_channels_last_graph_us'] = timed(g2.replay)
        out['towers_cl_out_is_channels_last'] = bool(r[0].is_contiguous(memory_format=torch.channels_last))
        c, p = towers(x)
        out['plain_block_us'] = timed(lambda: head.kp_rep_block_1(c, p))
        cat = torch.randn(16, 768, 25, 42, device='cuda')
        b2 = head.kp_rep_block_2
        out['cls_out_1x1_us'] = timed(lambda: b2.cls_out(cat))
        out['keypts_out_1x1_us'] = timed(lambda: b2.keypts_out(cat))
        k = b2.keypts_out(cat)
        out['reppts_out_1x1_us'] = timed(lambda: b2.reppts_out(k))
        a16 = cat.permute(0, 2, 3, 1).reshape(-1, 768).to(torch.bfloat16).contiguous()
        w16 = torch.randn(768, 768, device='cuda', dtype=torch.bfloat16)
        out['cublas_bf16_16800x768x768_us'] = timed(lambda: a16 @ w16.t())
        conv = head.cls_convs[0].conv
        out['conv3x3_nchw_us'] = timed(lambda: conv(x))
        gn = head.cls_convs[0].gn
        y = conv(x)
        out['gn_relu_nchw_us'] = timed(lambda: F.relu(gn(y)))
        conv_cl, gn_cl = head_cl.cls_convs[0].conv, head_cl.cls_convs[0].gn
        ycl = conv_cl(xcl)
        out['conv3x3_channels_last_us'] = timed(lambda: conv_cl(xcl))
        out['gn_relu_channels_last_us'] = timed(lambda: F.relu(gn_cl(ycl)))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
