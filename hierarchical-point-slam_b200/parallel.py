"""Multi-GPU mapping: one process per GPU, the ray batch is sharded contiguously across ranks,
cloud / feature tables / decoders are replicated, and one gradient all-reduce per iteration (NCCL over
NVLink on the GPU box, gloo in the CPU tests) makes every rank apply the identical optimiser step.

The reference has no distributed code at all; this is the data-parallel form of Mapper.py:783-860.
"""
import torch


def shard_bounds(n, rank, world):
    """Contiguous shard [lo, hi) of n rays for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(rank, world, *tensors):
    """Slice every per-ray tensor (first dim = rays) to this rank's shard."""
    n = tensors[0].shape[0]
    lo, hi = shard_bounds(n, rank, world)
    return tuple(t[lo:hi] for t in tensors)


class GradExchange:
    """Gradient all-reduce plan for a fixed set of tensors whose `.grad` buffers keep their addresses (CUDA-graph
    replays; the fused backward hands out all its gradients as views of ONE zero-initialised allocation).

    Gradients that live in the same storage are reduced with a single collective over the span that covers them
    (gaps - alignment padding, scratch - ride along); the stragglers (a few small tensors) share one flat bucket.
    Typical mapper iteration: one all-reduce over the fused backward's buffer + one small bucket, no per-tensor
    host work."""

    MAX_WASTE = 4.0     # coalesce a storage only if span <= MAX_WASTE x the gradient bytes inside it

    def __init__(self, tensors, skip_missing=False):
        """skip_missing: tensors whose .grad is None when the plan is built are left out (parameters the current stage
        does not touch - the set is the same on every rank because every rank runs the same code path); otherwise
        they get a zero gradient and take part."""
        self.all_tensors = list(tensors)
        self.tensors = self.all_tensors
        self.skip_missing = skip_missing
        self.key = None

    def _plan(self):
        if self.skip_missing:
            self.tensors = [t for t in self.all_tensors if t.grad is not None]
        for t in self.tensors:
            if t.grad is None:
                t.grad = torch.zeros_like(t)
        grads = [t.grad for t in self.tensors]
        self.key = tuple((g.data_ptr(), g.numel()) for g in grads)
        by_storage = {}
        for g in grads:
            ok = g.is_contiguous() and g.dtype == torch.float32
            by_storage.setdefault(g.untyped_storage().data_ptr() if ok else ("solo", id(g)), []).append(g)
        self.spans, self.bucket = [], []
        for k, gs in by_storage.items():
            lo = min(g.storage_offset() for g in gs)
            hi = max(g.storage_offset() + g.numel() for g in gs)
            used = sum(g.numel() for g in gs)
            if isinstance(k, tuple):                       # other dtype / layout: reduced on its own
                self.spans.extend(gs)
            elif (len(gs) == 1 and used < (1 << 16)) or (hi - lo) > self.MAX_WASTE * used:
                for g in gs:
                    (self.bucket if g.numel() < (1 << 16) else self.spans).append(g)
            else:
                self.spans.append(gs[0].as_strided((hi - lo,), (1,), lo))
        self.flat = None
        if self.bucket:
            self.flat = torch.empty(sum(g.numel() for g in self.bucket), dtype=self.bucket[0].dtype,
                                    device=self.bucket[0].device)
            self.views, off = [], 0
            for g in self.bucket:
                self.views.append(self.flat[off:off + g.numel()].view(g.shape))
                off += g.numel()

    def __call__(self, dist, group=None):
        if self.key is None or len(self.key) != len(self.tensors) or \
                any(t.grad is None or t.grad.data_ptr() != k[0] for t, k in zip(self.tensors, self.key)):
            self._plan()
        items = list(self.spans)
        if self.bucket:
            torch._foreach_copy_(self.views, self.bucket)
            items.append(self.flat)
        # NCCL: all buffers go out as ONE grouped launch (ncclGroupStart/End through the coalescing manager)
        cm = getattr(dist, "_coalescing_manager", None)
        if len(items) > 1 and cm is not None and items[0].is_cuda and dist.get_backend(group) == "nccl":
            with cm(group=group, device=items[0].device, async_ops=False):
                for t in items:
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        else:
            for t in items:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        if self.bucket:
            torch._foreach_copy_(self.bucket, self.views)


def allreduce_grads(tensors, dist, small_numel=1 << 16, group=None):
    """Sum `.grad` of every tensor across ranks (missing grads count as zero): GradExchange built on the fly.  Keep a
    GradExchange object instead when the same tensors are exchanged every iteration."""
    GradExchange(tensors)(dist, group)


def allreduce_scalars(values, dist, group=None):
    """Sum a few python/0-d values across ranks (tracker outlier threshold: sum and count of tmp,
    Tracker.py:217-218)."""
    t = torch.stack([torch.as_tensor(v, dtype=torch.float64).reshape(()) for v in values])
    dev = values[0].device if torch.is_tensor(values[0]) else "cpu"
    t = t.to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def render_img_sharded(renderer, dist, *args, group=None, gather=True, **kwargs):
    """Renderer.render_img (Renderer.py:256-337) with the image rows dealt out to the ranks ("shard the 103
    chunks / image rows; outputs are disjoint slices"): rank r renders rows shard_bounds(H, r, world) through
    render_img(row_range=...) against its replica of the cloud - no data-path collective.  gather=True: the row blocks
    are all-gathered (6 MB per 640x480 frame) and every rank returns the full (H,W) depth, uncertainty and (H,W,3)
    colour, equal to the single-process render; gather=False: the rank-local rows and their (r0, r1)."""
    rank = dist.get_rank(group) if dist is not None else 0
    world = dist.get_world_size(group) if dist is not None else 1
    H, W = renderer.H, renderer.W
    r0, r1 = shard_bounds(H, rank, world)
    depth, unc, color = renderer.render_img(*args, row_range=(r0, r1), **kwargs)
    if not gather or world == 1:
        return (depth, unc, color) if gather else (depth, unc, color, (r0, r1))
    rows = max(h - l for l, h in (shard_bounds(H, r, world) for r in range(world)))
    # one packed block per rank: [depth | unc | r | g | b] per pixel as float64 (depth / unc are float64 in the reference)
    block = torch.zeros(rows, W, 5, dtype=torch.float64, device=depth.device)
    block[:r1 - r0, :, 0], block[:r1 - r0, :, 1], block[:r1 - r0, :, 2:] = depth, unc, color.double()
    full = torch.empty(world * rows, W, 5, dtype=torch.float64, device=depth.device)
    dist.all_gather_into_tensor(full, block, group=group)
    parts = []
    for r in range(world):
        l, h = shard_bounds(H, r, world)
        parts.append(full[r * rows:r * rows + (h - l)])
    full = torch.cat(parts, 0)
    return full[..., 0].contiguous(), full[..., 1].contiguous(), full[..., 2:].to(color.dtype).contiguous()
