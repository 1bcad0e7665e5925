"""Multi-GPU mapping: one process per GPU, the ray batch is sharded contiguously across ranks,
cloud / feature tables / decoders are replicated, and one gradient all-reduce per iteration (NCCL over
NVLink on the GPU box, gloo in the CPU tests) makes every rank apply the identical optimiser step.

The reference has no distributed code at all; this is the data-parallel form of Mapper.py:783-860.
"""
import ctypes

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


class PeerAllReduce:
    """In-place sum of a symmetric fp32 buffer over the ranks of one box by ONE kernel of this package
    (csrc/peer.cu: barrier, reduce by slice owner over NVLink peer loads or one NVLS multimem.ld_reduce per vector,
    broadcast, barrier) instead of a library collective.  PyTorch is plumbing here: torch.distributed._symmetric_memory
    allocates the buffer and the signal pad, exchanges the handles and maps the peers' copies (and the multicast
    address when the fabric has NVLS); no kernel of it runs.

        ex = PeerAllReduce(n_floats, device, dist)     # collective: every rank constructs it
        ex.buffer                                      # this rank's copy: write gradients here (RenderSpec.grad_buffer)
        ex()                                           # every rank: sum in place, on the current stream, graph-capturable

    Raises if symmetric memory cannot be set up (no peer access): callers fall back to GradExchange (NCCL)."""

    def __init__(self, n_floats, device, dist, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self._lib = _lib
        self.dist = dist
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise _lib.HpsError("PeerAllReduce: at most 8 ranks (one NVSwitch box)")
        self.n = (int(n_floats) + 3) // 4 * 4
        lib = _lib.load()
        pad_words = int(lib.hps_peer_allreduce_pad_words())
        self.buffer = symm_mem.empty(self.n, dtype=torch.float32, device=device)
        self._pad = symm_mem.empty(pad_words, dtype=torch.int32, device=device)
        self._pad.zero_()
        self.buffer.zero_()
        torch.cuda.synchronize(device)
        hb = symm_mem.rendezvous(self.buffer, group.group_name)
        hp = symm_mem.rendezvous(self._pad, group.group_name)
        self._handles = (hb, hp)
        self._bufs = (ctypes.c_void_p * self.world)(*[int(p) for p in hb.buffer_ptrs])
        self._pads = (ctypes.c_void_p * self.world)(*[int(p) for p in hp.buffer_ptrs])
        self.multicast_ptr = int(getattr(hb, "multicast_ptr", 0) or 0)
        self._shift_cache = {}
        dist.barrier(group)                 # every pad is zeroed before anybody raises a flag

    def variants(self):
        return ["p2p"] + (["multimem"] if self.multicast_ptr else [])

    def default_variant(self):
        """Measured on the B200 box (3.4 MB): with two ranks the P2P form is faster (33 vs 38 us: one remote load per
        vector either way, and no switch reduction to pay for); from four ranks on one multimem.ld_reduce replaces W - 1
        remote loads."""
        return "multimem" if (self.multicast_ptr and self.world > 2) else "p2p"

    def check(self):
        """Host-side check (synchronises): raises if any exchange so far gave up waiting for a peer."""
        lib = self._lib.load()
        st = int(self._pad[int(lib.hps_peer_allreduce_ctas()) * 8].item())
        if st:
            raise self._lib.HpsError("PeerAllReduce: a peer did not arrive within 2 s (status %d); the sums are invalid" % st)

    def __call__(self, variant=None, n_floats=None, offset_floats=0):
        """Sum floats [offset_floats, offset_floats + n_floats) (default: all) of every rank's buffer in place.  variant:
        "multimem" (NVLS; the default when the buffer has a multicast address) or "p2p".  offset_floats: a multiple of 4
        (the kernel is handed the shifted peer addresses)."""
        from .ops import _call, _stream
        if variant is None:
            variant = self.default_variant()
        off = int(offset_floats)
        n = self.n - off if n_floats is None else (int(n_floats) + 3) // 4 * 4
        if off < 0 or off % 4 or n < 0 or off + n > self.n:
            raise self._lib.HpsError("PeerAllReduce: floats [%d, %d) asked (offset must be a multiple of 4), the buffer "
                                     "holds %d" % (off, off + n, self.n))
        if variant == "multimem" and not self.multicast_ptr:
            raise self._lib.HpsError("PeerAllReduce: no multicast address (NVLS) on this fabric")
        mc = ctypes.c_void_p(self.multicast_ptr + 4 * off) if variant == "multimem" else None
        bufs = self._bufs if off == 0 else self._shifted(off)
        with torch.cuda.device(self.buffer.device):
            _call("hps_peer_allreduce", bufs, self._pads, mc, self.rank, self.world, n, _stream())

    def _shifted(self, off):
        tab = self._shift_cache.get(off)
        if tab is None:
            tab = (ctypes.c_void_p * self.world)(*[int(p) + 4 * off for p in self._handles[0].buffer_ptrs])
            self._shift_cache[off] = tab
        return tab


class PeerGradExchange:
    """The mapper's gradient exchange with nothing but this package's kernel on the wire: the fused backward lays every
    gradient of the iteration out in `pool` (an ops.GradPool over a symmetric buffer - pass it as grad_buffer to
    Renderer.render_batch_ray_loss) and `exchange()` sums the used prefix over the ranks in place (PeerAllReduce).  A
    plain kernel launch: it can sit inside the iteration's CUDA graph.  capacity_floats: an upper bound of what one
    backward needs (live rows x 32 x tables + decoder parameters + padding); every rank must run the same code path."""

    def __init__(self, capacity_floats, device, dist, group=None, zero_in_forward=False):
        from .ops import GradPool
        self.peer = PeerAllReduce(capacity_floats, device, dist, group)
        self.pool = GradPool(self.peer.buffer, zero_in_forward)     # (see GradPool for what that asks of the caller)
        # The exchange kernel spends its time waiting for peers, and its 192 small CTAs must not queue behind the grid of
        # whatever runs beside it (the next batch's neighbour search: 3,125 CTAs at the same priority pushed the
        # exchange to the end of the step, round 2 measurement): its own stream, highest priority.
        self.stream = torch.cuda.Stream(device=device, priority=-5)
        self.side = torch.cuda.Stream(device=device)
        self._features_sent = 0

    def features(self, variant=None):
        """The early half (RenderSpec.after_feature_grads): sums pool.tensor[:pool.features] - the feature-table gradients,
        complete behind the scatter - on the exchange's stream while the backward goes on; the next __call__ then only
        moves the rest and joins both.  Every rank must take the same path."""
        n = self.pool.features
        if n:
            self.stream.wait_stream(torch.cuda.current_stream(self.peer.buffer.device))
            with torch.cuda.stream(self.stream):
                self.peer(variant, n_floats=n)
        self._features_sent = n

    def __call__(self, variant=None, beside=None):
        """Issued from the current stream: the exchange waits for it, runs on the high-priority stream, and the current
        stream waits for the exchange.  beside: a callable issued on a second stream at the same time - work that needs
        neither the gradients nor the parameters, i.e. the next batch's Renderer.prepare_batch (sampling + neighbour
        search): the exchange is latency-bound (~35 us of cross-GPU barriers for 3.4 MB), the search issue-bound, so the
        two share the GPU well; the exchange is launched first so that its 192 CTAs are placed before the search's grid."""
        cur = torch.cuda.current_stream(self.peer.buffer.device)
        sent, self._features_sent = self._features_sent, 0
        if self.pool.used:
            self.stream.wait_stream(cur)
            with torch.cuda.stream(self.stream):
                if self.pool.used > sent:
                    self.peer(variant, n_floats=self.pool.used - sent, offset_floats=sent)
        if beside is not None:
            self.side.wait_stream(cur)
            with torch.cuda.stream(self.side):
                beside()
            cur.wait_stream(self.side)
        if self.pool.used:
            cur.wait_stream(self.stream)

    def check(self):
        self.peer.check()
