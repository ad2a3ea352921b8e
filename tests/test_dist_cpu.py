"""world_size-2 gloo tests of the multi-rank host logic (image sharding, max-over-ranks timing,
gradient averaging with the reference's semantics, overlapped bucketed all-reduce)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    from kgdet_b200 import dist as kd
    r, w, _ = kd.init_from_env('gloo')
    assert (r, w) == (rank, world)
    out = {}
    # image sharding: contiguous, disjoint, complete
    out['shard'] = kd.shard_range(33, rank, world)
    # max over ranks
    out['max'] = kd.max_over_ranks(1.5 + rank, device='cpu')
    # gradient averaging == reference semantics (sum, then divide by world size)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    x = torch.randn(5, 8, generator=torch.Generator().manual_seed(100 + rank))
    net(x).square().sum().backward()
    local = [p.grad.clone() for p in net.parameters()]
    kd.allreduce_grads(net.parameters(), coalesce=True, bucket_size_mb=-1)
    out['avg'] = [p.grad.clone() for p in net.parameters()]
    out['local'] = local
    # overlapped bucketed variant gives the same averages
    net2 = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    net2.load_state_dict(net.state_dict())
    gb = kd.GradBucketer(net2.parameters(), bucket_size_mb=0.0002)     # several tiny buckets
    net2(x).square().sum().backward()
    gb.finish()
    out['avg2'] = [p.grad.clone() for p in net2.parameters()]
    # second step reuses the bucketer
    for p in net2.parameters():
        p.grad = None
    net2(x * 2).square().sum().backward()
    gb.finish()
    out['avg3_ok'] = all(torch.isfinite(p.grad).all().item() for p in net2.parameters())
    gb.remove()
    # flat gradient buffer: backward accumulates into views of one tensor, one collective, no copies
    net3 = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    net3.load_state_dict(net.state_dict())
    fg = kd.FlatGrads(net3.parameters())
    for scale in (3.0, 1.0):                          # second pass: zero() really clears the first one
        fg.zero()
        net3(x * scale).square().sum().backward()
    ptrs = [p.grad.data_ptr() for p in net3.parameters()]
    fg.allreduce()
    out['flat_views_kept'] = ptrs == [p.grad.data_ptr() for p in net3.parameters()] and \
        ptrs[0] == fg.flat.data_ptr()
    out['avg_flat'] = [p.grad.clone().tolist() for p in net3.parameters()]
    # bucketed all-reduce on slices of the flat buffer, launched from autograd hooks (the form bench.py captures
    # into the training graph); two steps through the same object
    net4 = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    net4.load_state_dict(net.state_dict())
    fg4 = kd.FlatGrads(net4.parameters())
    ov = kd.FlatBucketAllReduce(fg4, bucket_size_mb=0.0002)
    out['n_buckets'] = len(ov.buckets)
    for scale in (3.0, 1.0):
        ov.start()
        net4(x * scale).square().sum().backward()
        ov.finish()
    out['flat_bucket_views'] = fg4.attached()
    out['avg_flat_buckets'] = [p.grad.clone().tolist() for p in net4.parameters()]
    ov.remove()
    # a detached gradient view is detected (zero_grad(set_to_none=True) drops the aliases)
    net4.zero_grad(set_to_none=True)
    net4(x).square().sum().backward()
    out['detached_detected'] = not fg4.attached()
    try:
        fg4.allreduce()
        out['detached_raises'] = False
    except RuntimeError:
        out['detached_raises'] = True
    fg4.reattach()
    out['reattached'] = fg4.attached()
    # non-coalesced path
    for p, l in zip(net.parameters(), local):
        p.grad.copy_(l)
    kd.allreduce_grads(net.parameters(), coalesce=False)
    out['avg_nc'] = [p.grad.clone() for p in net.parameters()]
    for k in ('avg', 'local', 'avg2', 'avg_nc'):
        out[k] = [t.tolist() for t in out[k]]       # plain lists: no shared-memory handles in the queue
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_host_logic():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0]['shard'] == (0, 17) and res[1]['shard'] == (17, 33)
    assert res[0]['max'] == res[1]['max'] == 2.5
    for i in range(len(res[0]['avg'])):
        want = (torch.tensor(res[0]['local'][i]) + torch.tensor(res[1]['local'][i])) / 2
        for key in ('avg', 'avg2', 'avg_nc', 'avg_flat', 'avg_flat_buckets'):
            assert torch.allclose(torch.tensor(res[0][key][i]), want, atol=1e-6), key
            assert torch.allclose(torch.tensor(res[1][key][i]), want, atol=1e-6), key
    assert res[0]['avg3_ok'] and res[1]['avg3_ok']
    assert res[0]['flat_views_kept'] and res[1]['flat_views_kept']
    assert res[0]['n_buckets'] > 1 and res[0]['flat_bucket_views'] and res[1]['flat_bucket_views']
    for r in (0, 1):
        assert res[r]['detached_detected'] and res[r]['detached_raises'] and res[r]['reattached']


def test_single_process_is_a_no_op():
    from kgdet_b200 import dist as kd
    assert kd.world_size() == 1
    assert kd.shard_range(16, 0, 1) == (0, 16)
    assert kd.max_over_ranks(3.25) == 3.25
    net = torch.nn.Linear(4, 4)
    net(torch.randn(2, 4)).sum().backward()
    g = net.weight.grad.clone()
    kd.allreduce_grads(net.parameters())
    assert torch.equal(net.weight.grad, g)
