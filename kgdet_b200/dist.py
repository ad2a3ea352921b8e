"""Image-sharded data parallelism for the KGDet head path (one process per GPU).

Inference shards the image batch and needs no data-path collective.  The only
exchange step of the reference is the training gradient all-reduce,
``mmdet/core/utils/dist_utils.py:9-41`` (one flat bucket per dtype after the whole backward, then
``div_(world_size)``); ``allreduce_grads`` keeps that contract and adds the overlapped form
(``GradBucketer``): gradients are packed into ~25 MB buckets as autograd produces them and each
bucket's all-reduce is launched on a side stream so that NCCL traffic over NVLink/NVSwitch overlaps
the remaining backward kernels.  Plumbing only: ``torch.distributed`` (NCCL on GPUs, gloo in the CPU
tests).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).  Returns (rank, world, local)."""
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kw = {}
        if backend == 'nccl':
            torch.cuda.set_device(local)
            kw['device_id'] = torch.device('cuda', local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def shard_range(total, rank, world):
    """Contiguous image range [lo, hi) of `rank` when `total` images are split over `world` ranks."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value, device=None):
    """MAX-reduce a python float over ranks (device timings are reported as the slowest rank's)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' else 'cpu'
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _buckets(tensors, bucket_bytes):
    if bucket_bytes <= 0:                       # reference default: one bucket per dtype (dist_utils.py:14-20)
        by_type = {}
        for t in tensors:
            by_type.setdefault(t.dtype, []).append(t)
        return list(by_type.values())
    out, cur, size = [], [], 0
    for t in tensors:
        nb = t.numel() * t.element_size()
        if cur and (size + nb > bucket_bytes or t.dtype != cur[0].dtype):
            out.append(cur)
            cur, size = [], 0
        cur.append(t)
        size += nb
    if cur:
        out.append(cur)
    return out


def allreduce_grads(params, coalesce=True, bucket_size_mb=-1):
    """Average gradients over ranks; same semantics as mmdet/core/utils/dist_utils.py:31-41."""
    grads = [p.grad.data for p in params if p.requires_grad and p.grad is not None]
    ws = world_size()
    if ws == 1 or not grads:
        return
    if not coalesce:
        for g in grads:
            dist.all_reduce(g.div_(ws))
        return
    for bucket in _buckets(grads, int(bucket_size_mb * 1024 * 1024) if bucket_size_mb > 0 else -1):
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat)
        flat.div_(ws)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()


class FlatGrads(object):
    """All gradients as views of ONE flat fp32 buffer: the all-reduce of the reference
    (mmdet/core/utils/dist_utils.py:14-25 flattens, all-reduces, divides and copies back every step) becomes a single
    collective on memory the gradients already live in -- no flatten / unflatten copies -- and the buffer is a fixed
    address, so a CUDA-graph-captured backward accumulates straight into it.  `zero()` once per step (inside the
    captured region), `allreduce()` after backward; same result as `allreduce_grads` (sum, then / world size)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, 'no trainable parameters'
        p0 = self.params[0]
        assert all(p.dtype == p0.dtype and p.device == p0.device for p in self.params), 'one dtype / device'
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=p0.dtype, device=p0.device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def attached(self):
        """True while every parameter's .grad still aliases its slice of the flat buffer.  The default
        `optimizer.zero_grad()` / `module.zero_grad()` (set_to_none=True) drops the views: use
        `zero_grad(set_to_none=False)` or `FlatGrads.zero()` with this class."""
        off = 0
        base, esz = self.flat.data_ptr(), self.flat.element_size()
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != base + off * esz:
                return False
            off += p.numel()
        return True

    def reattach(self):
        """Point every .grad back at its slice (copying a detached gradient's values in first)."""
        off = 0
        for p in self.params:
            view = self.flat[off:off + p.numel()].view_as(p)
            if p.grad is not None and p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
            p.grad = view
            off += p.numel()

    def allreduce(self, check=True):
        if check and not self.attached():
            raise RuntimeError('FlatGrads: a parameter gradient no longer aliases the flat buffer (zero_grad('
                               'set_to_none=True)?): the all-reduce would average stale data. Use FlatGrads.zero() '
                               'or zero_grad(set_to_none=False), or call reattach().')
        ws = world_size()
        if ws > 1:
            dist.all_reduce(self.flat)
            self.flat.div_(ws)


class FlatBucketAllReduce(object):
    """Overlapped gradient averaging into a `FlatGrads` buffer.  The flat buffer is cut into contiguous buckets of
    whole parameters (~`bucket_size_mb` each).  Backward runs with `p.grad = None` (autograd then simply hands its
    fresh gradient tensors over: no zero-fill of the flat buffer, no read-modify-write accumulation, no extra add
    kernel per parameter); a post-accumulate-grad hook counts the gradients of a bucket as autograd finishes them
    and, when the last one arrives, ON A SIDE STREAM packs them into the bucket's slice of the flat buffer (one
    `torch.cat(out=slice)`) and all-reduces that slice in place (NCCL: ReduceOp.AVG, so there is no separate
    divide pass).  `finish()` joins the side stream and re-points every `p.grad` at its view of the flat buffer:
    the same values as the reference's one flat all-reduce after backward (mmdet/core/utils/dist_utils.py:14-25),
    just earlier in time.  Every call is capture-safe (stream waits, copies and NCCL only), so when
    `start() ... backward ... finish()` runs under `torch.cuda.graph` the collectives become a parallel branch of
    the captured backward."""

    def __init__(self, flat, bucket_size_mb=10):
        self.fg = flat
        self.ws = world_size()
        limit = int(bucket_size_mb * 1024 * 1024)
        self.buckets = []                # (lo, hi, [params]) element ranges of the flat buffer
        self.where = {}
        lo = off = 0
        cur = []
        esz = flat.flat.element_size()
        for p in flat.params:
            if cur and (off + p.numel() - lo) * esz > limit:
                self.buckets.append((lo, off, cur))
                lo, cur = off, []
            self.where[id(p)] = len(self.buckets)
            cur.append(p)
            off += p.numel()
        self.buckets.append((lo, off, cur))
        self.pending = None
        self.stream = None
        self.avg = dist.is_initialized() and dist.get_backend() == 'nccl'       # gloo has no ReduceOp.AVG
        self.handles = [p.register_post_accumulate_grad_hook(self._hook) for p in flat.params] if self.ws > 1 else []

    def start(self):
        for p in self.fg.params:
            p.grad = None
        self.pending = [len(b[2]) for b in self.buckets]
        self.launched = [False] * len(self.buckets)

    def _hook(self, p):
        if self.pending is None:
            return
        bi = self.where[id(p)]
        self.pending[bi] -= 1
        if self.pending[bi] == 0:
            self._launch(bi)

    def _launch(self, bi):
        if self.launched[bi]:
            return
        self.launched[bi] = True
        lo, hi, params = self.buckets[bi]
        flat = self.fg.flat

        def pack_and_reduce():
            off = lo
            pieces = []
            for p in params:                       # a parameter that received no gradient contributes zeros
                pieces.append(p.grad.reshape(-1) if p.grad is not None else flat.new_zeros(p.numel()))
                off += p.numel()
            torch.cat(pieces, out=flat[lo:hi])
            if self.avg:
                dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.AVG)
            else:
                dist.all_reduce(flat[lo:hi])
                flat[lo:hi].div_(self.ws)
        if flat.is_cuda:
            if self.stream is None:
                self.stream = torch.cuda.Stream(device=flat.device)
            main = torch.cuda.current_stream(flat.device)
            self.stream.wait_event(main.record_event())
            with torch.cuda.stream(self.stream):
                pack_and_reduce()
        else:
            pack_and_reduce()

    def finish(self):
        if self.ws == 1 or self.pending is None:
            return
        for bi in range(len(self.buckets)):          # buckets with parameters that received no gradient this step
            self._launch(bi)
        flat = self.fg.flat
        if flat.is_cuda and self.stream is not None:
            torch.cuda.current_stream(flat.device).wait_event(self.stream.record_event())
        off = 0
        for p in self.fg.params:                     # the averaged gradients live in the flat buffer
            p.grad = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.pending = None

    def remove(self):
        for h in self.handles:
            h.remove()
        self.handles = []


class GradBucketer(object):
    """Overlap the gradient all-reduce with backward.

    Registers post-accumulate-grad hooks; parameters are bucketed in reverse registration order (the
    order autograd finishes them).  When the last gradient of a bucket arrives, the bucket is flattened
    and all-reduced asynchronously (on a side CUDA stream for NCCL).  ``finish()`` waits for all buckets,
    divides by the world size and scatters the averages back -- numerically the same result as
    ``allreduce_grads`` (sum then divide), just earlier in time.
    """

    def __init__(self, params, bucket_size_mb=25):
        self.params = [p for p in params if p.requires_grad]
        self.ws = world_size()
        self.buckets = _buckets(list(reversed(self.params)), int(bucket_size_mb * 1024 * 1024))
        self.where = {}
        for bi, b in enumerate(self.buckets):
            for p in b:
                self.where[id(p)] = bi
        self.pending = None
        self.inflight = []
        self.stream = None
        self.handles = []
        if self.ws > 1:
            for p in self.params:
                self.handles.append(p.register_post_accumulate_grad_hook(self._hook))
        self.start()

    def start(self):
        self.pending = [len(b) for b in self.buckets]
        self.inflight = []

    def _hook(self, p):
        bi = self.where[id(p)]
        self.pending[bi] -= 1
        if self.pending[bi] == 0:
            self._launch(bi)

    def _launch(self, bi):
        bucket = [p for p in self.buckets[bi] if p.grad is not None]
        if not bucket:
            return
        if bucket[0].grad.is_cuda:
            if self.stream is None:
                self.stream = torch.cuda.Stream()
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                flat = torch.cat([p.grad.reshape(-1) for p in bucket])
                work = dist.all_reduce(flat, async_op=True)
        else:
            flat = torch.cat([p.grad.reshape(-1) for p in bucket])
            work = dist.all_reduce(flat, async_op=True)
        self.inflight.append((bucket, flat, work))

    def finish(self):
        """Call after backward(): completes outstanding buckets and writes averaged grads back."""
        if self.ws == 1:
            return
        for bi, n in enumerate(self.pending):       # parameters that received no grad this step
            if n > 0:
                self.pending[bi] = 0
                self._launch(bi)
        for bucket, flat, work in self.inflight:
            work.wait()
            if flat.is_cuda:
                torch.cuda.current_stream().wait_stream(self.stream)
            flat.div_(self.ws)
            off = 0
            for p in bucket:
                p.grad.copy_(flat[off:off + p.grad.numel()].view_as(p.grad))
                off += p.grad.numel()
        self.start()

    def remove(self):
        for h in self.handles:
            h.remove()
        self.handles = []
