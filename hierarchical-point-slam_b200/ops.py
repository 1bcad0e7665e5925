"""Operator layer: torch tensors in, C-ABI kernel calls out, autograd wiring around them.

PyTorch is plumbing here (device memory, streams, autograd graph); every floating-point operation of
the hot path runs in libhps_b200.so.  No op has a fallback: non-CUDA tensors raise.
"""
import ctypes
import os

import torch

from . import _lib

K_NN = 8
C_DIM = 32
MLP_GEO, MLP_COL = 0, 1
N_PARAM = 23


def _stream(stream=None):
    return ctypes.c_void_p((stream if stream is not None else torch.cuda.current_stream()).cuda_stream)


# The two decoders of a colour stage are independent between the gather and the compositing kernel: the fused
# render runs the geometry decoder (and all weight re-packing) on a side stream so that the hardware block
# scheduler fills the tail of one persistent kernel with CTAs of the other.  Everything is allocated on the main
# stream before the fork and the side stream joins before anything is released, so the caching allocator needs no
# record_stream calls; the fork/join pattern is capturable in a CUDA graph.  HPS_STREAMS=0 serialises.
MULTI_STREAM = os.environ.get("HPS_STREAMS", "1") != "0"
_side_streams = {}


def _side_stream(dev, which=0, priority=0):
    key = (torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device(), which, priority)
    s = _side_streams.get(key)
    if s is None:
        s = torch.cuda.Stream(device=dev, priority=priority)
        _side_streams[key] = s
    return s


# ---- instrumentation: kernel-launch counter and optional CUDA-event timers per C-ABI call -----------
LAUNCHES = {"n": 0}
_KERNELS_PER_CALL = {"hps_grid_build": 6, "hps_grid_append": 3, "hps_knn_query": 1, "hps_sample_knn": 1, "hps_interp_fwd": 1, "hps_interp_bwd": 1, "hps_rows_store": 1,
                     "hps_mlp_pack": 1, "hps_mlp_fwd": 1, "hps_mlp_bwd": 1, "hps_mlp_tc_pack": 1, "hps_mlp_tc_fwd": 1,
                     "hps_mlp_tc_bwd_pack": 2, "hps_mlp_tc_bwd": 3, "hps_tc_selftest": 1, "hps_composite_fwd": 1,
                     "hps_composite_bwd": 1, "hps_composite_mapper_loss": 1, "hps_sample_rays": 1, "hps_sample_rays_bwd": 1, "hps_mapper_loss": 1,
                     "hps_tracker_loss_stats": 1, "hps_tracker_loss": 1, "hps_exposure_fwd": 1, "hps_exposure_bwd": 1, "hps_camera_from_tensor": 1, "hps_adam_rows": 1, "hps_adam_multi": 1,
                     "hps_camera_from_tensor_bwd": 1, "hps_rays_from_uv": 1, "hps_rays_from_uv_bwd": 1, "hps_peer_allreduce": 1}
TIMERS = None   # set to {} to record (start, end) CUDA events per call name on the current stream


# Scheduling of RenderSpec.tail_work (see _RenderRays._backward): stream priority and launch position.  Defaults from
# scripts/tail_experiment.py on the B200 (round 2, mean of 60 flushed replays of the bench iteration): tail work in front
# of the feature scatter as ONE launch, weight-gradient streams at high priority: 0.2670 ms against 0.2751 ms without any
# overlap; the other fifteen combinations lie between 0.2697 and 0.2806 ms.  The tail work's priority is one step ABOVE
# the scatter's: both become ready at the same instant (the dgrad kernels' end) and inside a replayed graph their launch
# order is the driver's choice, not the program's - at equal priority a change elsewhere in the graph flipped it, the
# scatter's 3,125 CTAs were placed first and the search started 40 us late (0.265 vs 0.256 ms).
TAIL_PRIORITY = int(os.environ.get("HPS_TAIL_PRIORITY", "-1"))
TAIL_AFTER_SCATTER = os.environ.get("HPS_TAIL_AFTER_SCATTER", "0") != "0"
WGRAD_PRIORITY = int(os.environ.get("HPS_WGRAD_PRIORITY", "-2"))
SCATTER_PRIORITY = int(os.environ.get("HPS_SCATTER_PRIORITY", "-2"))      # (above the tail work where an exchange waits for it)
SCATTER_STREAM_ALWAYS = os.environ.get("HPS_SCATTER_STREAM", "0") != "0"
COLOR_DGRAD_PRIORITY = int(os.environ.get("HPS_COLOR_DGRAD_PRIORITY", "-1"))

# NVTX ranges: HPS_NVTX=1 wraps every C-ABI call in a range named after the entry point (decoder
# calls carry the decoder kind) and the fused render's forward / backward in "hps.render_rays.fwd/bwd"; nsys / ncu
# --nvtx then attribute kernels to calls.  Off by default: two extra Python calls per launch.
NVTX = os.environ.get("HPS_NVTX", "0") != "0"


def _call(name, *args):
    """Invoke one C-ABI entry point on the current stream (raises on a non-zero return code)."""
    fn = getattr(_lib.load(), name)
    LAUNCHES["n"] += _KERNELS_PER_CALL.get(name, 0)
    if NVTX:
        torch.cuda.nvtx.range_push("%s[%d]" % (name, args[0]) if name.startswith("hps_mlp") and isinstance(args[0], int) else name)
        try:
            _lib.check(fn(*args))
        finally:
            torch.cuda.nvtx.range_pop()
        return
    if TIMERS is None:
        _lib.check(fn(*args))
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _lib.check(fn(*args))
    b.record()
    key = "%s[%d]" % (name, args[0]) if name in ("hps_mlp_fwd", "hps_mlp_bwd", "hps_mlp_tc_fwd", "hps_mlp_tc_bwd") else name
    TIMERS.setdefault(key, []).append((a, b))


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.HpsError("hps_b200 ops need CUDA tensors (got %s); there is no CPU fallback" % t.device)


def _f32c(t):
    t = t.detach()
    if t.dtype is torch.float32 and t.is_contiguous():
        return t
    return t.to(torch.float32).contiguous()


class _ZeroPool:
    """Zero-initialised scratch tensors carved out of ONE allocation (one memset kernel for all of them); every
    tensor starts on a 16-byte boundary.  want() before alloc(), get() after."""

    def __init__(self, device):
        self.device, self.total, self.reqs, self.flat = device, 0, [], None
        self.features = 0       # floats of the leading feature-table gradients (see _RenderRays._plan_pool)

    def want(self, shape):
        n = 1
        for d in shape:
            n *= int(d)
        self.reqs.append((self.total, n, tuple(shape)))
        self.total += (n + 3) // 4 * 4
        return len(self.reqs) - 1

    def want_like(self, tensors, skip_first):
        return [None if (skip_first and i == 0) else self.want(t.shape) for i, t in enumerate(tensors)]

    def alloc(self, stream=None, into=None, clear=True):
        """stream: issue the memset there (the caller orders it: that stream must already wait for the work that last
        touched the allocator's block, and the consumers must wait for that stream).  into: a GradPool whose tensor
        provides the memory (e.g. a symmetric buffer the gradient exchange reduces in place) instead of the allocator.
        clear=False: the caller clears `flat` itself (total is a multiple of 4 floats, flat is 16-byte aligned)."""
        n = max(self.total, 4)
        if into is not None:
            if into.tensor.numel() < n or into.tensor.dtype != torch.float32 or not into.tensor.is_contiguous():
                raise _lib.HpsError("grad_buffer: need %d contiguous fp32 elements, got %d" % (n, into.tensor.numel()))
            self.flat = into.tensor[:n]
            into.used, into.features = self.total, self.features
        else:
            self.flat = torch.empty(n, dtype=torch.float32, device=self.device)
        if clear:
            if stream is None:
                self.flat.zero_()
            else:
                with torch.cuda.stream(stream):
                    self.flat.zero_()
        return self

    def get(self, h):
        if h is None:
            return None
        if isinstance(h, list):
            return [self.get(x) for x in h]
        off, n, shape = self.reqs[h]
        return self.flat[off:off + n].view(shape)


class GradPool:
    """Caller-owned memory for the gradients of one fused backward (RenderSpec.grad_buffer): `tensor` is a contiguous
    fp32 buffer, `used` the number of leading floats the last backward laid its gradients out in (every .grad of that
    backward is a view of tensor[:used]).  With a symmetric-memory tensor (parallel.PeerAllReduce.buffer) the multi-GPU
    exchange sums the gradients where backward wrote them.
    zero_in_forward=True lets the FORWARD of a training iteration clear the buffer (its compositing kernel does it on the
    side, which takes the memset out of the backward's critical path): only for callers that are done with the previous
    iteration's gradients before they render again with parameters that require grad (forward -> backward -> exchange ->
    optimiser step -> forward ...); the default leaves them alone until the next backward."""

    def __init__(self, tensor, zero_in_forward=False):
        self.tensor, self.used, self.zero_in_forward = tensor, 0, bool(zero_in_forward)
        self.features = 0       # tensor[:features]: the feature-table gradients (complete first), tensor[features:used]: the rest


def _zero_grads_like(tensors, skip_first):
    """Zero-initialised gradient buffers for a parameter list as views of ONE flat allocation (one memset instead
    of one per parameter); every view starts on a 16-byte boundary (the wgrad kernels use red.global.add.v4)."""
    sizes = [0 if (skip_first and i == 0) else (t.numel() + 3) // 4 * 4 for i, t in enumerate(tensors)]
    flat = torch.zeros(sum(sizes), dtype=torch.float32, device=tensors[-1].device)
    out, off = [], 0
    for t, n in zip(tensors, sizes):
        out.append(None if n == 0 else flat[off:off + t.numel()].view(t.shape))
        off += n
    return out


# --------------------------------------------------------------------------------------------------
# radius thresholds: fp32 equivalents of the reference's float64 comparisons
# --------------------------------------------------------------------------------------------------
def radius_thresholds(radius, n, device):
    """radius: python float or tensor ((n,), (n,1); float64 in the reference) -> (t_up, t_dn) fp32 (n,).

    D < r64^2  <=>  D < t_up ;   D > r64^2  <=>  D > t_dn   for every fp32 D."""
    if torch.is_tensor(radius):
        r2 = radius.detach().to(device=device, dtype=torch.float64).reshape(-1) ** 2
        if r2.numel() != n:
            raise ValueError("shape mis-match for input points and dynamic radius")
    else:
        r2 = torch.full((n,), float(radius) ** 2, dtype=torch.float64, device=device)
    t = r2.to(torch.float32)
    inf = torch.tensor(float("inf"), device=device)
    t_dn = torch.where(t.double() > r2, torch.nextafter(t, -inf), t)
    t_up = torch.where(t.double() < r2, torch.nextafter(t, inf), t)
    return t_up.contiguous(), t_dn.contiguous()


# --------------------------------------------------------------------------------------------------
# neighbour index
# --------------------------------------------------------------------------------------------------
class VoxelHashGrid:
    """Owner of one hps_grid_t handle (one per hierarchy level)."""

    def __init__(self):
        lib = _lib.load()
        h = ctypes.c_void_p()
        _lib.check(lib.hps_grid_create(ctypes.byref(h)))
        self._h = h
        self.n = 0
        self.cell = None
        self.generation = 0

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.load().hps_grid_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def build(self, pos, cell, check=True):
        """(Re)build over pos (N,3).  check: read the status word back (one host synchronisation per build - builds are
        rare: insertions go through append()) and raise if a point fell outside the 21-bit cell range, where the exact-
        search guarantee would silently break (cell size far too small for the extent)."""
        _need_cuda(pos)
        pos = _f32c(pos).reshape(-1, 3)
        with torch.cuda.device(pos.device):
            _call("hps_grid_build", self._h, _lib.ptr(pos), pos.shape[0], float(cell), _stream())
            self._keep = pos            # keep the source alive until the stream has consumed it
            self.n = pos.shape[0]
            self.cell = float(cell)
            self.generation += 1        # device addresses / hash mask may have changed: captured CUDA graphs are stale
            if check and self.info()["status"] & 1:
                raise _lib.HpsError("voxel grid: points outside the 21-bit cell range (cell %.3g m is too small for the "
                                    "cloud's extent)" % self.cell)

    def append(self, pos, n_old):
        """pos (N,3): the first n_old rows are the points the grid holds, rows [n_old, N) are hashed into the existing
        structure (O(new points); device addresses and the hash mask stay, so captured CUDA graphs remain valid).
        -> status (host int, one synchronisation): 0 = done, otherwise the caller must build() again (a cell or the
        arrays ran out of the room they were built with, or a point lies outside the origin's coordinate range)."""
        _need_cuda(pos)
        pos = _f32c(pos).reshape(-1, 3)
        st = ctypes.c_int(0)
        with torch.cuda.device(pos.device):
            _call("hps_grid_append", self._h, _lib.ptr(pos), int(n_old), pos.shape[0], ctypes.byref(st), _stream())
        if st.value == 0:
            self._keep = pos
            self.n = pos.shape[0]
        return st.value

    def info(self):
        n, T, nc, cell, st = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_float(), ctypes.c_int()
        _lib.check(_lib.load().hps_grid_info(self._h, ctypes.byref(n), ctypes.byref(T), ctypes.byref(nc),
                                             ctypes.byref(cell), ctypes.byref(st), _stream()))
        return dict(n_points=n.value, table_size=T.value, n_cells=nc.value, cell_size=cell.value, status=st.value)

    def query(self, q, t_up, t_dn, want_cand=False, group=None):
        """q (Q,3) -> I (Q,8) int32, D (Q,8) f32, count (Q,) int32 [, cand (Q,) int32].  group=5: consecutive
        queries in fives (the samples of a ray) share one sweep over the bounding box of their five balls - only for
        queries that are that close; the result is the same."""
        group = group or 1
        _need_cuda(q, t_up, t_dn)
        q = _f32c(q).reshape(-1, 3)
        Q = q.shape[0]
        I = torch.empty(Q, K_NN, dtype=torch.int32, device=q.device)
        D = torch.empty(Q, K_NN, dtype=torch.float32, device=q.device)
        cnt = torch.empty(Q, dtype=torch.int32, device=q.device)
        cand = torch.empty(Q, dtype=torch.int32, device=q.device) if want_cand else None
        with torch.cuda.device(q.device):
            _call("hps_knn_query", self._h, _lib.ptr(q), _lib.ptr(t_up), _lib.ptr(t_dn), Q,
                  _lib.ptr(I), _lib.ptr(D), _lib.ptr(cnt), _lib.ptr(cand), int(group), _stream())
        return (I, D, cnt, cand) if want_cand else (I, D, cnt)


# --------------------------------------------------------------------------------------------------
# inverse-distance interpolation
# --------------------------------------------------------------------------------------------------
class _Interp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, feats_a, feats_b, cloud_pos, I, D, t_dn, count, fill_a, fill_b, min_nn, tracker):
        _need_cuda(p, feats_a, feats_b, I, D, t_dn, count, fill_a)
        P = I.shape[0]
        dev = I.device
        p_c = _f32c(p).reshape(-1, 3)
        fa = _f32c(feats_a)
        fb = _f32c(feats_b) if feats_b is not None else None
        pos = _f32c(cloud_pos).reshape(-1, 3) if tracker else None
        w = torch.empty(P, K_NN, dtype=torch.float32, device=dev)
        c_a = torch.empty(P, C_DIM, dtype=torch.float32, device=dev)
        c_b = torch.empty(P, C_DIM, dtype=torch.float32, device=dev) if fb is not None else None
        has = torch.empty(P, dtype=torch.uint8, device=dev)
        fill_a = _f32c(fill_a)
        fill_b = _f32c(fill_b) if fill_b is not None else None
        with torch.cuda.device(dev):
            _call("hps_interp_fwd", 
                _lib.ptr(p_c), _lib.ptr(I), _lib.ptr(D), _lib.ptr(t_dn), _lib.ptr(count), _lib.ptr(pos),
                _lib.ptr(fa), _lib.ptr(fb), _lib.ptr(fill_a), _lib.ptr(fill_b), P, int(min_nn),
                _lib.ptr(w), _lib.ptr(c_a), _lib.ptr(c_b), _lib.ptr(has), _stream())
        ctx.save_for_backward(p_c, fa, fb, pos, I, w, has, c_a, c_b)
        ctx.tracker = tracker
        ctx.mark_non_differentiable(has)
        if c_b is None:
            c_b = torch.empty(0, device=dev)
            ctx.mark_non_differentiable(c_b)
        return c_a, c_b, has

    @staticmethod
    def backward(ctx, dc_a, dc_b, _dhas):
        p_c, fa, fb, pos, I, w, has, c_a, c_b = ctx.saved_tensors
        P = I.shape[0]
        need_p, need_a, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        dc_a = _f32c(dc_a) if dc_a is not None else None
        dc_b = _f32c(dc_b) if (dc_b is not None and fb is not None) else None
        dfa = torch.zeros_like(fa) if (need_a and dc_a is not None) else None
        dfb = torch.zeros_like(fb) if (need_b and dc_b is not None) else None
        dp = torch.empty(P, 3, dtype=torch.float32, device=I.device) if (need_p and ctx.tracker) else None
        if dfa is not None or dfb is not None or dp is not None:
            with torch.cuda.device(I.device):
                _call("hps_interp_bwd", 
                    _lib.ptr(p_c), _lib.ptr(I), _lib.ptr(w), _lib.ptr(has), _lib.ptr(pos), _lib.ptr(fa),
                    _lib.ptr(fb), _lib.ptr(c_a), _lib.ptr(c_b), _lib.ptr(dc_a), _lib.ptr(dc_b), P,
                    _lib.ptr(dfa), _lib.ptr(dfb), _lib.ptr(dp), None, _stream())
        return dp, dfa, dfb, None, None, None, None, None, None, None, None, None


def interpolate(p, feats_a, feats_b, cloud_pos, I, D, t_dn, count, fill_a, fill_b, min_nn=2, tracker=False):
    """IDW features for up to two tables sharing one neighbour list.  Returns c_a, c_b|None, has (bool)."""
    c_a, c_b, has = _Interp.apply(p, feats_a, feats_b, cloud_pos, I, D, t_dn, count, fill_a, fill_b, min_nn,
                                  bool(tracker))
    return c_a, (c_b if feats_b is not None else None), has.bool()


# --------------------------------------------------------------------------------------------------
# decoder MLPs
# --------------------------------------------------------------------------------------------------
def _ptr_array(ts):
    arr = (ctypes.c_void_p * N_PARAM)()
    for i, t in enumerate(ts):
        arr[i] = None if t is None else t.data_ptr()
    return arr


_packed_cache = {}

# Which edition of the decoder forward runs: True = tcgen05 tensor-core kernel (mlp_tc.cu), False = fp32 SIMT
# kernel (mlp.cu, the parity anchor).  Both fill the same activation workspace for the backward.
USE_TC = os.environ.get("HPS_USE_TC", "1") != "0"


def _packed_buffer(kind, device, tc=False):
    key = (kind, device, tc)
    buf = _packed_cache.get(key)
    if buf is None:
        lib = _lib.load()
        if tc:
            buf = torch.empty(int(lib.hps_mlp_tc_packed_bytes(kind)), dtype=torch.uint8, device=device)
        else:
            buf = torch.empty(int(lib.hps_mlp_packed_floats(kind)), dtype=torch.float32, device=device)
        _packed_cache[key] = buf
    return buf


# Compositing backward inside the fused compositing + mapper-loss kernel of the forward pass (HPS_FUSE_COMPOSITE_BWD=0:
# the separate hps_composite_bwd launch at the head of the backward).
ZERO_IN_COMPOSITE = os.environ.get("HPS_ZERO_IN_COMPOSITE", "1") != "0"   # gradient buffers cleared by the compositing kernel
FUSE_COMPOSITE_BWD = os.environ.get("HPS_FUSE_COMPOSITE_BWD", "1") != "0"

# True: the backward also runs on tensor cores (dgrad chain + streaming wgrad kernels); False: fp32 SIMT backward.
USE_TC_BWD = os.environ.get("HPS_USE_TC_BWD", "1") != "0"


def _mlp_workspace(kind, P, need_grad, dev):
    """Saved-for-backward workspace of one decoder forward: ("tc"|"simt", tensor) or None."""
    lib = _lib.load()
    if not need_grad:
        return None
    if USE_TC and USE_TC_BWD:
        return ("tc", torch.empty(int(lib.hps_mlp_tc_ws_bytes(kind, P)), dtype=torch.uint8, device=dev))
    return ("simt", torch.empty((P + 63) // 64 * 64 * int(lib.hps_mlp_act_floats(kind)), dtype=torch.float32,
                                device=dev))


def _mlp_pack(kind, arr, dev, stream=None, bwd=None):
    """Re-pack the (possibly just updated) decoder weights into the tensor-core operand images.  bwd: a buffer
    (_new_bwd_images) owned by the caller that receives the backward images too, so that a later backward of
    this very forward finds the weights it was run with."""
    _call("hps_mlp_tc_pack", kind, arr, _lib.ptr(_packed_buffer(kind, dev, tc=True)), _stream(stream))
    if bwd is not None:
        _call("hps_mlp_tc_bwd_pack", kind, arr, _lib.ptr(bwd), _stream(stream))


def _new_bwd_images(kind, dev):
    return torch.empty(int(_lib.load().hps_mlp_tc_bwd_packed_bytes(kind)), dtype=torch.uint8, device=dev)


N_SM = 148
# Ray-grouped neighbour search (one warp answers the 5 samples of a ray) trades parallelism for shared probes and
# loads: measured on the bench cloud (scripts/knn_group_bench.py) 55 vs 47 us at 25,000 queries (too few warps to
# hide latency), 143 vs 152 us at 100,000, 269 vs 292 us at 200,000.
KNN_GROUP_MIN_QUERIES = 100000
# per-tile cost of the colour decoder relative to the geometry decoder (measured: 27/16 us forward, 40/21 us dgrad)
_COST_RATIO = {"fwd": 1.7, "bwd": 1.9}


def _sm_split(P, which):
    """(colour CTAs, geometry CTAs) for two decoder kernels that run concurrently on two streams.  Each kernel owns
    its SMs (shared memory does not fit twice), so the split that minimises max(ceil(T/nc)*cost_c, ceil(T/ng)*cost_g)
    beats letting one kernel queue behind the other's first wave."""
    tiles = (P + 127) // 128
    if tiles * 2 <= N_SM:
        return 0, 0                      # everything fits in one wave: no cap
    best, r = None, _COST_RATIO[which]
    for nc in range(1, N_SM):
        ng = N_SM - nc
        t = max(-(-tiles // nc) * r, -(-tiles // ng) * 1.0)
        if best is None or t < best[0] - 1e-9:
            best = (t, nc, ng)
    return best[1], best[2]


# Debugging switch: verify every decoder output for inf / NaN (one host synchronisation per decoder call).  The
# tensor-core decoders multiply fp16 hi/lo halves: operands beyond fp16's range (|activation| or |weight| > 65,504) overflow
# to inf, which always reaches the output - never a plausible finite number ("dynamic range").
CHECK_FINITE = os.environ.get("HPS_CHECK_FINITE", "0") != "0"


def _mlp_forward(kind, arr, p_c, c_c, aff, P, out, need_grad, dev, stream=None, saved=None, packed=False, max_ctas=0):
    """pack + forward of one decoder on `stream` (default: current).  Returns the saved workspace ("tc"|"simt",
    tensor) that _mlp_backward consumes (None when no backward will follow)."""
    lib = _lib.load()
    if USE_TC:
        if saved is None:
            saved = _mlp_workspace(kind, P, need_grad, dev)
        pk = _packed_buffer(kind, dev, tc=True)
        act = saved[1] if (saved and saved[0] == "simt") else None
        ws = saved[1] if (saved and saved[0] == "tc") else None
        if not packed:
            _call("hps_mlp_tc_pack", kind, arr, _lib.ptr(pk), _stream(stream))
        _call("hps_mlp_tc_fwd", kind, arr, _lib.ptr(pk), _lib.ptr(p_c), _lib.ptr(c_c), _lib.ptr(aff), P, _lib.ptr(out),
              _lib.ptr(act), _lib.ptr(ws), int(max_ctas), _stream(stream))
    else:
        pk = _packed_buffer(kind, dev)
        act = None
        if need_grad:
            act = torch.empty((P + 63) // 64 * 64 * int(lib.hps_mlp_act_floats(kind)), dtype=torch.float32, device=dev)
            saved = ("simt", act)
        _call("hps_mlp_pack", kind, arr, _lib.ptr(pk), _stream())
        _call("hps_mlp_fwd", kind, arr, _lib.ptr(pk), _lib.ptr(p_c), _lib.ptr(c_c), _lib.ptr(aff), P, _lib.ptr(out),
              _lib.ptr(act), _stream())
    if CHECK_FINITE:
        (stream if stream is not None else torch.cuda.current_stream()).synchronize()
        if not bool(torch.isfinite(out).all()):
            raise _lib.HpsError("decoder %s: non-finite output - an operand left the range of the fp16-split tensor-core "
                                "path (|x| <= 65,504); set HPS_USE_TC=0 for the fp32 SIMT decoders" % ("geo", "col")[kind])
    return saved


_bwd_packed_cache = {}


def _bwd_packed_buffer(kind, dev):
    key = (kind, dev)
    pk = _bwd_packed_cache.get(key)
    if pk is None:
        pk = torch.empty(int(_lib.load().hps_mlp_tc_bwd_packed_bytes(kind)), dtype=torch.uint8, device=dev)
        _bwd_packed_cache[key] = pk
    return pk


def _mlp_dp_buffer(saved_kind, P, need_dp, dev):
    if not need_dp:
        return None
    if saved_kind == "tc":
        return torch.zeros(P, 3, dtype=torch.float32, device=dev)      # accumulated by the kernel
    return torch.empty(P, 3, dtype=torch.float32, device=dev)


def _absmax_word(kind, packed):
    """Address of the max|dout| word behind a backward weight image (see hps_mlp_tc_bwd in the header)."""
    return ctypes.c_void_p(packed.data_ptr() + int(_lib.load().hps_mlp_tc_bwd_absmax_offset(kind)))


def _mlp_backward(kind, prm, saved_kind, saved, p_c, c_c, aff, dout, P, dprm, dc, need_dp, daff, dev, stream=None,
                  dp=None, packed=None, absmax_done=False, max_ctas=0, wgrad_stream=None, gscale=None):
    """Backward of one decoder on `stream`; returns dp (P,3) or None.  dprm: list of gradient buffers (or None).
    dp: preallocated _mlp_dp_buffer (required when `stream` is not the current stream); packed: backward weight
    images built at forward time by _mlp_pack (None: build them now)."""
    darr = _ptr_array(dprm) if dprm is not None else None
    if dp is None:
        dp = _mlp_dp_buffer(saved_kind, P, need_dp, dev)
    if saved_kind == "tc":
        arr = _ptr_array(prm)
        pk = packed
        if pk is None:
            pk = _bwd_packed_buffer(kind, dev)
            _call("hps_mlp_tc_bwd_pack", kind, arr, _lib.ptr(pk), _stream(stream))
        _call("hps_mlp_tc_bwd", kind, arr, _lib.ptr(pk), _lib.ptr(saved), _lib.ptr(p_c), _lib.ptr(aff), _lib.ptr(dout), P,
              darr, _lib.ptr(dc), _lib.ptr(dp), _lib.ptr(daff), 1 if (absmax_done and packed is not None) else 0,
              int(max_ctas), _stream(wgrad_stream) if wgrad_stream is not None else None, _lib.ptr(gscale),
              _stream(stream))
        return dp
    _call("hps_mlp_bwd", kind, _ptr_array(prm), _lib.ptr(p_c), _lib.ptr(c_c), _lib.ptr(aff), _lib.ptr(saved),
          _lib.ptr(dout), P, darr, _lib.ptr(dc), _lib.ptr(dp), _lib.ptr(daff), _stream(stream))
    return dp


class _Mlp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kind, p, c, affine, *params):
        _need_cuda(p, c, affine, *params)
        lib = _lib.load()
        dev = c.device
        P = c.shape[0]
        p_c, c_c = _f32c(p).reshape(-1, 3), _f32c(c)
        prm = [_f32c(t) for t in params]
        aff = _f32c(affine) if affine is not None else None
        nout = 1 if kind == MLP_GEO else 3
        out = torch.empty((P,) if nout == 1 else (P, 3), dtype=torch.float32, device=dev)
        need_grad = any(ctx.needs_input_grad)
        arr = _ptr_array(prm)
        with torch.cuda.device(dev):
            saved = _mlp_forward(kind, arr, p_c, c_c, aff, P, out, need_grad, dev)
        ctx.kind = kind
        ctx.has_affine = aff is not None
        ctx.saved_kind = saved[0] if saved else None
        ctx.save_for_backward(p_c, c_c, aff, saved[1] if saved else None, *prm)
        return out

    @staticmethod
    def backward(ctx, dout):
        p_c, c_c, aff, act = ctx.saved_tensors[:4]
        prm = list(ctx.saved_tensors[4:])
        kind, dev, P = ctx.kind, c_c.device, c_c.shape[0]
        need = ctx.needs_input_grad          # (kind, p, c, affine, *params)
        want_w = any(need[4 + i] for i in range(1, N_PARAM))
        want_b = need[4] and kind == MLP_GEO
        dprm = [None] * N_PARAM
        if want_w:
            dprm = _zero_grads_like(prm, skip_first=not want_b)
        elif want_b:
            raise _lib.HpsError("gradient w.r.t. embedder._B alone is not supported")
        dc = torch.empty(P, C_DIM, dtype=torch.float32, device=dev)
        daff = torch.zeros(12, dtype=torch.float32, device=dev) if (ctx.has_affine and need[3]) else None
        with torch.cuda.device(dev):
            dp = _mlp_backward(kind, prm, ctx.saved_kind, act, p_c, c_c, aff, _f32c(dout), P, dprm if want_w else None,
                               dc, need[1], daff, dev)
        grads = [g if (g is not None and need[4 + i]) else None for i, g in enumerate(dprm)]
        return (None, dp, dc if need[2] else None, daff, *grads)


def mlp(kind, p, c, params, affine=None):
    """params: the 23 tensors [B, (W_i, b_i, Wc_i, bc_i) x5, Wo, bo] in the reference layouts."""
    assert len(params) == N_PARAM
    return _Mlp.apply(kind, p, c, affine, *params)


# --------------------------------------------------------------------------------------------------
# compositing
# --------------------------------------------------------------------------------------------------
class _CompositeRaw(torch.autograd.Function):
    """raw (R,S,4) packed [rgb, occ]; point_mask (R*S,) uint8/bool or None; z (R,S)."""

    @staticmethod
    def forward(ctx, raw, z, point_mask, coef, want_weights):
        _need_cuda(raw, z, point_mask)
        R, S = z.shape
        dev = raw.device
        raw_c, z_c = _f32c(raw).reshape(R, S, 4), _f32c(z)
        pm = point_mask.detach().to(torch.uint8).contiguous() if point_mask is not None else None
        depth = torch.empty(R, dtype=torch.float32, device=dev)
        var = torch.empty(R, dtype=torch.float32, device=dev)
        color = torch.empty(R, 3, dtype=torch.float32, device=dev)
        weights = torch.empty(R, S, dtype=torch.float32, device=dev) if want_weights else None
        base = raw_c.data_ptr()
        with torch.cuda.device(dev):
            _call("hps_composite_fwd", 
                ctypes.c_void_p(base), 4, ctypes.c_void_p(base + 12), 4, _lib.ptr(pm), _lib.ptr(z_c), R, S,
                float(coef), _lib.ptr(depth), _lib.ptr(var), _lib.ptr(color), _lib.ptr(weights), 0, None, None, None, 0,
                _stream())
        ctx.save_for_backward(raw_c, z_c, pm)
        ctx.coef = float(coef)
        ctx.mark_non_differentiable(var)
        if weights is None:
            weights = torch.empty(0, device=dev)
        ctx.mark_non_differentiable(weights)
        return depth, var, color, weights

    @staticmethod
    def backward(ctx, g_depth, _g_var, g_color, _g_w):
        raw_c, z_c, pm = ctx.saved_tensors
        R, S = z_c.shape
        d_raw = torch.empty_like(raw_c)
        gd = _f32c(g_depth) if g_depth is not None else None
        gc = _f32c(g_color) if g_color is not None else None
        base, dbase = raw_c.data_ptr(), d_raw.data_ptr()
        with torch.cuda.device(raw_c.device):
            _call("hps_composite_bwd", 
                ctypes.c_void_p(base), 4, ctypes.c_void_p(base + 12), 4, _lib.ptr(pm), _lib.ptr(z_c), R, S, ctx.coef,
                _lib.ptr(gd), _lib.ptr(gc), None, ctypes.c_void_p(dbase), 4, ctypes.c_void_p(dbase + 12), 4, None, None,
                _stream())
        return d_raw, None, None, None, None


def composite(raw, z, point_mask=None, coef=0.1, want_weights=False):
    depth, var, color, w = _CompositeRaw.apply(raw, z, point_mask, coef, want_weights)
    return depth, var, color, (w if want_weights else None)


# --------------------------------------------------------------------------------------------------
# live (frustum-selected) feature rows
# --------------------------------------------------------------------------------------------------
class LiveRows:
    """The rows of a feature table that are being optimised (reference Mapper.py:512-536: frustum-selected rows are
    cloned into leaf tensors and written back into the full table every iteration, :615-630, :862-874).

    `param` (M,32) is the autograd leaf the optimiser owns; `table` (N,32) is the full, replicated table the
    kernels gather from.  Passing a LiveRows as npc_geo_feats[level] / npc_col_feats[level] makes the fused
    render path (a) refresh the live rows of `table` from `param` with one index copy and (b) scatter feature
    gradients straight into a compact (M,32) buffer through `row_map` - no (N,32) gradient is materialised, and
    the multi-GPU exchange shrinks from the whole table to the live rows."""

    def __init__(self, table, indices, row_map=None):
        self.table = table.detach()
        assert self.table.is_contiguous() and self.table.dtype is torch.float32
        self.indices = indices.to(device=table.device, dtype=torch.long)
        if row_map is None:
            row_map = torch.full((table.shape[0],), -1, dtype=torch.int32, device=table.device)
            row_map[self.indices] = torch.arange(self.indices.numel(), dtype=torch.int32, device=table.device)
        self.row_map = row_map
        self.param = self.table[self.indices].clone().requires_grad_(True)
        self.fresh = False       # True: table[indices] == param is guaranteed (RowAdam.step writes both at once)

    @classmethod
    def pair(cls, geo_table, col_table, indices):
        a = cls(geo_table, indices)
        return a, cls(col_table, indices, a.row_map)

    def refresh(self):
        """table[indices] = param (no autograd).  Skipped while RowAdam owns the update (it stores the new rows into the
        table itself); with any other optimiser on `param` the copy runs before every render."""
        if self.fresh:
            return
        with torch.no_grad():
            self.table.index_copy_(0, self.indices, self.param)

    @staticmethod
    def refresh_pair(a, b, stream=None):
        """refresh() of up to two LiveRows (either may be None) as ONE launch (hps_rows_store) instead of one torch
        index_copy_ each: the write-back sits at the head of every iteration's critical path."""
        todo = [r for r in (a, b) if r is not None and not r.fresh]
        if not todo:
            return
        for r in todo:
            _need_cuda(r.table, r.param)
            if r.table.shape[1] != C_DIM or not r.param.is_contiguous():
                raise _lib.HpsError("LiveRows: (M,%d) contiguous fp32 leaves expected" % C_DIM)
        x, y = todo[0], (todo[1] if len(todo) > 1 else None)
        with torch.cuda.device(x.table.device):
            _call("hps_rows_store", _lib.ptr(x.table), _lib.ptr(x.param), _lib.ptr(x.indices), x.indices.numel(),
                  _lib.ptr(y.table) if y else None, _lib.ptr(y.param) if y else None,
                  _lib.ptr(y.indices) if y else None, y.indices.numel() if y else 0, _stream(stream))

    def dense(self):
        """Differentiable (N,32) view for code paths that want a plain tensor (reference semantics)."""
        return self.table.index_put((self.indices,), self.param)


# --------------------------------------------------------------------------------------------------
# fused ray rendering: sampling -> neighbour search -> interpolation -> decoders -> compositing
# --------------------------------------------------------------------------------------------------
class RayPrep:
    """The parameter-free head of a render call, computed ahead of time: depth-guided samples, their radius
    thresholds and the exact neighbour lists (Renderer.py:163-229 + neural_point.py:366-417).  Nothing in it depends
    on features or decoder weights, so a multi-GPU mapper can run it for iteration i+1 while the gradients of
    iteration i are still being exchanged (bench.py does).  Build with prepare_rays(); pass as RenderSpec.prep."""
    __slots__ = ("z", "pts", "t_up", "t_dn", "I", "D", "cnt", "R", "S")


def _z_zero_stride(zz, R, S):
    """z for rays without a depth reading: one shared row (S,) -> stride 0, or one row per ray (R,S) -> stride S."""
    if zz is None or zz.dim() == 1:
        if zz is not None and zz.numel() != S:
            raise ValueError("z_zero must hold S values (shared) or R x S (per ray)")
        return 0
    if tuple(zz.shape) != (R, S):
        raise ValueError("z_zero must hold S values (shared) or R x S (per ray)")
    return S


def prepare_rays(grid, rays_o, rays_d, gt_depth, radius, S, near, far, z_zero=None, out=None):
    """Sampling + neighbour search for R rays (two kernel launches).  `out`: a RayPrep to refill in place."""
    _need_cuda(rays_o, rays_d, gt_depth)
    dev = rays_o.device
    R = rays_o.shape[0]
    P = R * S
    f32 = dict(dtype=torch.float32, device=dev)
    ro, rd, gd = _f32c(rays_o).reshape(-1, 3), _f32c(rays_d).reshape(-1, 3), _f32c(gt_depth).reshape(-1)
    rad_t = radius.detach().to(device=dev, dtype=torch.float64).reshape(-1).contiguous() if torch.is_tensor(radius) else None
    if rad_t is not None and rad_t.numel() != R:
        raise ValueError("shape mis-match for rays and dynamic radius")
    zz = _f32c(z_zero) if z_zero is not None else None
    pr = out
    if pr is None or pr.R != R or pr.S != S:
        pr = RayPrep()
        pr.R, pr.S = R, S
        pr.z, pr.pts = torch.empty(R, S, **f32), torch.empty(P, 3, **f32)
        pr.t_up, pr.t_dn = torch.empty(P, **f32), torch.empty(P, **f32)
        pr.I = torch.empty(P, K_NN, dtype=torch.int32, device=dev)
        pr.D = torch.empty(P, K_NN, **f32)
        pr.cnt = torch.empty(P, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        # one launch (the search computes its query points itself) below the ray-grouped size: a prepared head is
        # typically issued beside other work, where a second launch waits for thread slots a second time
        _sample_and_search(grid, ro, rd, gd, rad_t, radius, zz, R, S, near, far, pr.z, pr.pts, pr.t_up, pr.t_dn, pr.I,
                           pr.D, pr.cnt, _stream(), fused=PREPARE_FUSED)
    return pr


# hps_sample_knn (one launch) instead of hps_sample_rays + hps_knn_query: bit-identical, measured 1 % slower on the
# bench workload (the merged kernel repeats the sampling arithmetic in every lane), so off unless asked for
SAMPLE_IN_SEARCH = os.environ.get("HPS_SAMPLE_IN_SEARCH", "0") != "0"


PREPARE_FUSED = os.environ.get("HPS_PREPARE_FUSED", "1") != "0"


def _sample_and_search(grid, ro, rd, gd, rad_t, rad, zz, R, S, near, far, z, pts, t_up, t_dn, I, D, cnt, stream, fused=None):
    """Sample points + thresholds + exact neighbour search: the sampling kernel, then the search (ray-grouped from
    KNN_GROUP_MIN_QUERIES queries on); with SAMPLE_IN_SEARCH one launch (the search computes its query points itself)."""
    P = R * S
    scalar = 0.0 if rad_t is not None else float(rad)
    if fused is None:
        fused = SAMPLE_IN_SEARCH
    if (S == 5 and P >= KNN_GROUP_MIN_QUERIES) or not fused:
        _call("hps_sample_rays", _lib.ptr(ro), _lib.ptr(rd), _lib.ptr(gd), _lib.ptr(rad_t), scalar, _lib.ptr(zz),
              _z_zero_stride(zz, R, S), R, S, float(near), float(far), _lib.ptr(z), _lib.ptr(pts), _lib.ptr(t_up),
              _lib.ptr(t_dn), stream)
        _call("hps_knn_query", grid._h, _lib.ptr(pts), _lib.ptr(t_up), _lib.ptr(t_dn), P, _lib.ptr(I), _lib.ptr(D),
              _lib.ptr(cnt), None, 5 if (S == 5 and P >= KNN_GROUP_MIN_QUERIES) else 1, stream)
    else:
        _call("hps_sample_knn", grid._h, _lib.ptr(ro), _lib.ptr(rd), _lib.ptr(gd), _lib.ptr(rad_t), scalar, _lib.ptr(zz),
              _z_zero_stride(zz, R, S), R, S, float(near), float(far), _lib.ptr(z), _lib.ptr(pts), _lib.ptr(t_up),
              _lib.ptr(t_dn), _lib.ptr(I), _lib.ptr(D), _lib.ptr(cnt), None, stream)


class RenderSpec:
    """Non-tensor arguments of render_rays (one per call)."""
    __slots__ = ("grid", "radius", "z_zero", "S", "near", "far", "coef", "min_nn", "min_valid", "color_stage",
                 "tracker", "cloud_pos", "fill_geo", "fill_col", "live_geo", "live_col", "prep", "loss", "aff_mlp",
                 "tail_work", "grad_buffer", "after_grads", "after_feature_grads")

    def __init__(self, **kw):
        for k in self.__slots__:
            setattr(self, k, kw.get(k))


def _exposure_grad_floats(eprm):
    hidden, n_in = eprm[1].shape
    return sum((n + 3) // 4 * 4 for n in (hidden * n_in, hidden, 12 * hidden, 12, n_in))


def _exposure_backward(eprm, a1, dout, gscale, e_need, x_shape, stream, buf=None):
    """d/d(latent, W1, b1, W2, b2) of the exposure MLP from d/d affine (12 floats, times the device scalar `gscale`
    when given): one kernel on `stream`, one allocation (or `buf`: _exposure_grad_floats(eprm) floats, fully
    overwritten).  e_need: which of the five gradients are wanted."""
    if not any(e_need):
        return (None,) * 5
    hidden, n_in = eprm[1].shape
    sizes = [hidden * n_in, hidden, 12 * hidden, 12, n_in]
    if buf is None:
        buf = torch.empty(sum((n + 3) // 4 * 4 for n in sizes), dtype=torch.float32, device=a1.device)
    parts, off = [], 0
    for n in sizes:
        parts.append(buf[off:off + n])
        off += (n + 3) // 4 * 4
    _call("hps_exposure_bwd", _lib.ptr(eprm[0]), _lib.ptr(eprm[1]), _lib.ptr(eprm[3]), _lib.ptr(a1), _lib.ptr(dout),
          _lib.ptr(gscale), n_in, hidden, 12, _lib.ptr(parts[0]), _lib.ptr(parts[1]), _lib.ptr(parts[2]),
          _lib.ptr(parts[3]), _lib.ptr(parts[4]) if e_need[0] else None, _stream(stream))
    shaped = (parts[4].view(x_shape), parts[0].view(hidden, n_in), parts[1], parts[2].view(12, hidden), parts[3])
    return tuple(g if e_need[i] else None for i, g in enumerate(shaped))


def _exposure_params(tensors):
    """(latent, W1, b1, W2, b2) as contiguous fp32 with the sizes checked -> (list, hidden, n_in)."""
    ep = [_f32c(t) for t in tensors]
    if len(ep) != 5:
        raise _lib.HpsError("exposure MLP: expected latent, W1, b1, W2, b2")
    ep[0] = ep[0].reshape(-1)
    hidden, n_in = ep[1].shape
    if ep[0].numel() != n_in or ep[3].shape[0] != 12:
        raise _lib.HpsError("exposure MLP: latent / output size mismatch")
    return ep, hidden, n_in


class MapperLossSpec:
    """Fuses the mapper's loss (Mapper.py:791-856) into render_rays: `gt_color` (R,3) or None (geometry stages),
    `w_color`, and what follows the decoder parameters in the argument list: `exposure` = "mlp": the five tensors of
    the exposure MLP (latent, W1, b1, W2, b2) - the colour term then is |gt - sigmoid(colour x affine)| with affine =
    MLP(latent), evaluated inside the call; "affine": the affine itself, (12,) or (F,12) with `frame_index` (R,) for a
    batch that mixes rays of F keyframes (Mapper.py:805-822); None: the colour as rendered."""
    __slots__ = ("gt_color", "w_color", "exposure", "frame_index")

    def __init__(self, gt_color, w_color, exposure=None, frame_index=None):
        if exposure not in (None, "mlp", "affine"):
            raise ValueError("exposure: None, 'mlp' or 'affine'")
        self.gt_color, self.w_color, self.exposure, self.frame_index = gt_color, float(w_color), exposure, frame_index


class TrackerLossSpec:
    """Fuses the tracker's loss (Tracker.py:213-231) into render_rays: the batch statistic sum |gt - depth| / sqrt(var)
    is accumulated by the compositing kernel, the loss kernel applies the 10 x mean outlier mask.  `gt_color` (R,3) or
    None (depth only), `w_color`, `reduce_fn` as in tracker_loss (rays sharded over ranks), `zero_missing`: the
    renderer reports depth 0 for rays without a depth reading (Renderer.py:250-251)."""
    __slots__ = ("gt_color", "w_color", "reduce_fn", "zero_missing")

    def __init__(self, gt_color, w_color, reduce_fn=None, zero_missing=False):
        self.gt_color, self.w_color, self.reduce_fn, self.zero_missing = gt_color, float(w_color), reduce_fn, zero_missing


class _RenderRays(torch.autograd.Function):
    """depth (R,), var (R,), color (R,3), valid (R,) uint8  =  f(rays_o, rays_d, gt_depth, feats_geo, feats_col,
    affine, 23 geometry-decoder tensors, 23 colour-decoder tensors).  Ten kernel launches forward, five or
    six backward; nothing but buffer allocation happens in PyTorch."""

    @staticmethod
    def forward(ctx, spec, rays_o, rays_d, gt_depth, feats_geo, feats_col, affine, *params):
        if NVTX:
            torch.cuda.nvtx.range_push("hps.render_rays.fwd")
            try:
                return _RenderRays._forward(ctx, spec, rays_o, rays_d, gt_depth, feats_geo, feats_col, affine, *params)
            finally:
                torch.cuda.nvtx.range_pop()
        return _RenderRays._forward(ctx, spec, rays_o, rays_d, gt_depth, feats_geo, feats_col, affine, *params)

    @staticmethod
    def _forward(ctx, spec, rays_o, rays_d, gt_depth, feats_geo, feats_col, affine, *params):
        _need_cuda(rays_o, rays_d, gt_depth, feats_geo)
        dev = rays_o.device
        R, S = rays_o.shape[0], spec.S
        P = R * S
        f32 = dict(dtype=torch.float32, device=dev)
        ro, rd, gd = _f32c(rays_o).reshape(-1, 3), _f32c(rays_d).reshape(-1, 3), _f32c(gt_depth).reshape(-1)
        # live-row mode: feats_* are the compact (M,32) leaves; the kernels gather from the refreshed full tables
        row_map = None
        fork = MULTI_STREAM and TIMERS is None and USE_TC and spec.color_stage
        if spec.live_geo is not None:
            if not fork:
                spec.live_geo.refresh()
            fg = spec.live_geo.table
            row_map = spec.live_geo.row_map
        else:
            fg = _f32c(feats_geo)
        fc = None
        if spec.color_stage:
            if spec.live_col is not None:
                if not fork:
                    spec.live_col.refresh()
                fc = spec.live_col.table
                row_map = spec.live_col.row_map
            else:
                fc = _f32c(feats_col)
        if (spec.live_geo is None) != (spec.live_col is None) and spec.color_stage:
            raise _lib.HpsError("pass LiveRows for both feature tables of a level or for neither")
        aff = _f32c(affine) if (affine is not None and spec.color_stage) else None
        gprm = [_f32c(t) for t in params[:N_PARAM]]
        cprm = [_f32c(t) for t in params[N_PARAM:2 * N_PARAM]] if spec.color_stage else []
        # loss mode: [loss, pad x3 | d_affine (12) | g_depth (R) | g_color (3R)] in one buffer, head zeroed (accumulators)
        lm = spec.loss
        eprm, flat, a1_l, aff_l, frm_l, head, n_aff, aff_in_shape = [], None, None, None, None, 16, 12, None
        trk_loss = isinstance(lm, TrackerLossSpec)
        if lm is not None:
            color_loss = spec.color_stage and lm.gt_color is not None
            gtc = _f32c(lm.gt_color).reshape(-1, 3) if color_loss else None
            if not trk_loss and lm.exposure == "affine" and color_loss:
                aff_in = params[len(gprm) + len(cprm)]
                aff_in_shape = aff_in.shape
                aff_l = _f32c(aff_in)
                frm_l, n_aff = _loss_frames(aff_l, lm.frame_index, R)
                head = _loss_head_floats(n_aff)
            flat = torch.empty(head + R * (4 if color_loss else 1), **f32)   # tracker: [4] = the batch statistic
            if not trk_loss and lm.exposure == "mlp" and color_loss:
                ctx.exp_x_shape = params[len(gprm) + len(cprm)].shape
                eprm, e_hidden, e_in = _exposure_params(params[len(gprm) + len(cprm):])
                a1_l, aff_l = torch.empty(e_hidden, **f32), torch.empty(12, **f32)
        # decoder-side exposure affine (tracker, decoder.py:606-612) from the MLP's own tensors: evaluated beside the
        # neighbour search instead of in front of the call, its backward beside the feature scatter
        mprm, a1_d = [], None
        if spec.aff_mlp and spec.color_stage:
            if eprm or aff_l is not None or aff is not None:
                raise _lib.HpsError("exposure: either the decoder-side MLP or a loss-side correction, not both")
            ctx.aff_x_shape = params[len(gprm) + len(cprm)].shape
            mprm, m_hidden, m_in = _exposure_params(params[len(gprm) + len(cprm):])
            a1_d, aff = torch.empty(m_hidden, **f32), torch.empty(12, **f32)

        def aff_head(s):
            _call("hps_exposure_fwd", _lib.ptr(mprm[0]), _lib.ptr(mprm[1]), _lib.ptr(mprm[2]), _lib.ptr(mprm[3]),
                  _lib.ptr(mprm[4]), m_in, m_hidden, 12, _lib.ptr(a1_d), _lib.ptr(aff), _stream(s))

        def loss_head(s):
            """zero the accumulators and evaluate the exposure MLP (independent of the rays): off the critical path"""
            with torch.cuda.stream(s):
                flat[:head].zero_()
            if eprm:
                _call("hps_exposure_fwd", _lib.ptr(eprm[0]), _lib.ptr(eprm[1]), _lib.ptr(eprm[2]), _lib.ptr(eprm[3]),
                      _lib.ptr(eprm[4]), e_in, e_hidden, 12, _lib.ptr(a1_l), _lib.ptr(aff_l), _stream(s))
        rad = spec.radius
        rad_t = rad.detach().to(device=dev, dtype=torch.float64).reshape(-1).contiguous() if torch.is_tensor(rad) else None
        if rad_t is not None and rad_t.numel() != R:
            raise ValueError("shape mis-match for rays and dynamic radius")
        prep = spec.prep
        if prep is not None:
            if prep.R != R or prep.S != S:
                raise _lib.HpsError("RayPrep was built for %d rays x %d samples, the call has %d x %d" % (prep.R, prep.S, R, S))
            z, pts, t_up, t_dn, I, D, cnt = prep.z, prep.pts, prep.t_up, prep.t_dn, prep.I, prep.D, prep.cnt
        else:
            z = torch.empty(R, S, **f32)
            pts = torch.empty(P, 3, **f32)
            t_up, t_dn = torch.empty(P, **f32), torch.empty(P, **f32)
            I = torch.empty(P, K_NN, dtype=torch.int32, device=dev)
            D = torch.empty(P, K_NN, **f32)
            cnt = torch.empty(P, dtype=torch.int32, device=dev)
        w = torch.empty(P, K_NN, **f32)
        c_geo = torch.empty(P, C_DIM, **f32)
        c_col = torch.empty(P, C_DIM, **f32) if spec.color_stage else None
        has = torch.empty(P, dtype=torch.uint8, device=dev)
        occ = torch.empty(P, **f32)
        rgb = torch.empty(P, 3, **f32) if spec.color_stage else None
        depth, var, color = torch.empty(R, **f32), torch.empty(R, **f32), torch.empty(R, 3, **f32)
        valid = torch.empty(R, dtype=torch.uint8, device=dev)
        need_grad = any(ctx.needs_input_grad)
        pos = _f32c(spec.cloud_pos).reshape(-1, 3) if spec.tracker else None
        fill_g, fill_c = _f32c(spec.fill_geo), (_f32c(spec.fill_col) if spec.color_stage else None)
        zz = _f32c(spec.z_zero) if spec.z_zero is not None else None
        garr, carr = _ptr_array(gprm), (_ptr_array(cprm) if spec.color_stage else None)
        st = _stream
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            bpk_g = bpk_c = None
            if fork:
                # side stream: weight images for forward (and backward) while the main stream samples and searches
                side = _side_stream(dev)
                ws_g = _mlp_workspace(MLP_GEO, P, need_grad, dev)
                ws_c = _mlp_workspace(MLP_COL, P, need_grad, dev)
                if need_grad and USE_TC_BWD:
                    bpk_g, bpk_c = _new_bwd_images(MLP_GEO, dev), _new_bwd_images(MLP_COL, dev)
                side.wait_stream(main)
                refreshed = None
                if spec.live_geo is not None or spec.live_col is not None:
                    LiveRows.refresh_pair(spec.live_geo, spec.live_col, side)     # table[live rows] = leaves, one launch
                    refreshed = side.record_event()
                # forward images first (the decoders wait for them), the backward images afterwards
                _mlp_pack(MLP_COL, carr, dev, side)
                if mprm:
                    aff_head(side)
                packed_col = side.record_event()
                _mlp_pack(MLP_GEO, garr, dev, side)
                # the loss's accumulators and exposure affine (two tiny kernels) on a stream of their own: in front of or
                # behind the geometry decoder on the side stream they sat on the critical path (that kernel finishes last)
                lh_s = None
                if lm is not None:
                    lh_s = _side_stream(dev, 8)
                    lh_s.wait_stream(main)
                    loss_head(lh_s)
                aux = None
                if bpk_c is not None:      # backward images on a third stream: nothing in the forward waits for them but
                    aux = _side_stream(dev, 1)      # the compositing (it folds max |d raw| into the images' scale words)
                    aux.wait_stream(main)
                    _call("hps_mlp_tc_bwd_pack", MLP_COL, carr, _lib.ptr(bpk_c), _stream(aux))
                    _call("hps_mlp_tc_bwd_pack", MLP_GEO, garr, _lib.ptr(bpk_g), _stream(aux))
            if lm is not None and not fork:
                loss_head(main)
            if mprm and not fork:
                aff_head(main)
            if prep is None:
                _sample_and_search(spec.grid, ro, rd, gd, rad_t, rad, zz, R, S, spec.near, spec.far, z, pts, t_up, t_dn,
                                   I, D, cnt, st())
            if fork and refreshed is not None:
                main.wait_event(refreshed)
            _call("hps_interp_fwd", _lib.ptr(pts), _lib.ptr(I), _lib.ptr(D), _lib.ptr(t_dn), _lib.ptr(cnt),
                  _lib.ptr(pos), _lib.ptr(fg), _lib.ptr(fc), _lib.ptr(fill_g), _lib.ptr(fill_c), P, int(spec.min_nn),
                  _lib.ptr(w), _lib.ptr(c_geo), _lib.ptr(c_col), _lib.ptr(has), st())
            if fork:
                side.wait_stream(main)
                main.wait_event(packed_col)
                n_c, n_g = _sm_split(P, "fwd")
                sv_c = _mlp_forward(MLP_COL, carr, pts, c_col, aff, P, rgb, need_grad, dev, main, ws_c, packed=True,
                                    max_ctas=n_c)
                sv_g = _mlp_forward(MLP_GEO, garr, pts, c_geo, None, P, occ, need_grad, dev, side, ws_g, packed=True,
                                    max_ctas=n_g)
                main.wait_stream(side)
                if lh_s is not None:
                    main.wait_stream(lh_s)
                if aux is not None:
                    main.wait_stream(aux)
            else:
                sv_g = _mlp_forward(MLP_GEO, garr, pts, c_geo, None, P, occ, need_grad, dev)
                sv_c = _mlp_forward(MLP_COL, carr, pts, c_col, aff, P, rgb, need_grad, dev) if spec.color_stage else None
            ctx.bwd_packed = (bpk_g, bpk_c)
            act_g, act_c = (sv_g[1] if sv_g else None), (sv_c[1] if sv_c else None)
            ctx.saved_kinds = (sv_g[0] if sv_g else None, sv_c[0] if sv_c else None)
            if lm is None:
                _call("hps_composite_fwd", _lib.ptr(rgb), 3, _lib.ptr(occ), 1, _lib.ptr(has), _lib.ptr(z), R, S,
                      float(spec.coef), _lib.ptr(depth), _lib.ptr(var), _lib.ptr(color), None, int(spec.min_valid),
                      _lib.ptr(valid), None, None, 0, st())
            elif trk_loss:
                _call("hps_composite_fwd", _lib.ptr(rgb), 3, _lib.ptr(occ), 1, _lib.ptr(has), _lib.ptr(z), R, S,
                      float(spec.coef), _lib.ptr(depth), _lib.ptr(var), _lib.ptr(color), None, int(spec.min_valid),
                      _lib.ptr(valid), _lib.ptr(gd), _lib.ptr(flat[4:]), 1 if lm.zero_missing else 0, st())
                count = float(R)
                if lm.reduce_fn is not None:       # rays sharded over ranks: global sum and count
                    flat[5:6].fill_(float(R))
                    count = float(lm.reduce_fn(flat[4:6]))
                # (rays without a depth reading are masked out of the loss by gt_depth > 0, whatever their depth)
                _call("hps_tracker_loss", _lib.ptr(depth), _lib.ptr(var), _lib.ptr(color) if color_loss else None,
                      _lib.ptr(gd), _lib.ptr(gtc), _lib.ptr(flat[4:]), count, lm.w_color, R, _lib.ptr(flat),
                      _lib.ptr(flat[16:]), _lib.ptr(flat[16 + R:]) if color_loss else None, st())
            else:
                # with forward-time backward images (the fork path of a training iteration) the compositing backward is
                # evaluated right here for an upstream gradient of 1 and max |d raw| lands in the images' scale words: the
                # backward then starts with the decoders' dgrad kernels (gscale = the real upstream gradient)
                fuse_bwd = (FUSE_COMPOSITE_BWD and fork and need_grad and bpk_g is not None and bpk_c is not None
                            and aff is None and not spec.tracker)
                zpool = None
                if fuse_bwd:
                    d_occ_f = torch.empty(P, **f32)
                    d_rgb_f = torch.empty(P, 3, **f32) if spec.color_stage else None
                    if ZERO_IN_COMPOSITE and (spec.grad_buffer is None or spec.grad_buffer.zero_in_forward):
                        # the backward's gradient buffers, laid out now and cleared by the compositing kernel on the side:
                        # no memset (and no stream hop) between the loss and the first dgrad kernel
                        zpool = _RenderRays._plan_pool(
                            spec, ctx.needs_input_grad, gprm, cprm, aff,
                            feats_geo.shape[0] if row_map is not None else fg.shape[0], P, dev, eprm, mprm, sv_g[0])
                        zpool[0].alloc(into=spec.grad_buffer, clear=False)
                        ctx.zeroed_pool = zpool
                _call("hps_composite_mapper_loss", _lib.ptr(rgb), 3, _lib.ptr(occ), 1, _lib.ptr(has), _lib.ptr(z), R, S,
                      float(spec.coef), int(spec.min_valid), _lib.ptr(gd), _lib.ptr(gtc), _lib.ptr(aff_l),
                      _lib.ptr(frm_l), lm.w_color, _lib.ptr(depth), _lib.ptr(var), _lib.ptr(color), _lib.ptr(valid),
                      _lib.ptr(flat), _lib.ptr(flat[head:]), _lib.ptr(flat[head + R:]) if color_loss else None,
                      _lib.ptr(flat[4:]) if aff_l is not None else None,
                      _lib.ptr(d_rgb_f) if fuse_bwd else None, _lib.ptr(d_occ_f) if fuse_bwd else None,
                      _absmax_word(MLP_COL, bpk_c) if (fuse_bwd and d_rgb_f is not None) else None,
                      _absmax_word(MLP_GEO, bpk_g) if fuse_bwd else None,
                      _lib.ptr(zpool[0].flat) if zpool is not None else None,
                      (zpool[0].total + 3) // 4 * 4 if zpool is not None else 0, st())
                if fuse_bwd:
                    ctx.fused_d_raw = (d_occ_f, d_rgb_f)
        if not hasattr(ctx, "fused_d_raw"):
            ctx.fused_d_raw = None
        if not hasattr(ctx, "zeroed_pool"):
            ctx.zeroed_pool = None
        ctx.loss_state = None if lm is None else (flat, a1_l, eprm, color_loss, head, n_aff, aff_in_shape)
        ctx.aff_mlp_state = (mprm, a1_d) if mprm else None
        ctx.spec = spec
        ctx.n_g = len(gprm)
        ctx.row_map = row_map
        ctx.n_live = feats_geo.shape[0] if row_map is not None else None
        ctx.save_for_backward(z, pts, I, w, has, c_geo, c_col, occ, rgb, act_g, act_c, fg, fc, aff, pos, *gprm, *cprm)
        ctx.set_materialize_grads(False)        # no zero tensors for the outputs nobody differentiates
        if lm is not None:
            ctx.mark_non_differentiable(depth, var, color, valid)
            return flat[0], depth, var, color, valid
        ctx.mark_non_differentiable(var, valid)
        return depth, var, color, valid

    @staticmethod
    def _plan_pool(spec, need, gprm, cprm, aff, n_rows, P, dev, l_eprm, mprm, geo_kind):
        """Layout of every zero-initialised buffer of the backward inside ONE allocation (_ZeroPool, not yet allocated)
        and the handles: (dg, dcw, daff, dfg, dfc, exposure, dp_geo, dp_col).  Depends only on shapes and on which
        inputs want a gradient (l_eprm / mprm: the loss-side / decoder-side exposure MLP's tensors or nothing; n_rows:
        rows of the feature gradients), so the forward of a training iteration can lay it out ahead of the backward."""
        zp = _ZeroPool(dev)
        want_gw = any(need[7 + i] for i in range(1, N_PARAM))
        want_cw = spec.color_stage and any(need[7 + N_PARAM + i] for i in range(1, N_PARAM))
        need_p = spec.tracker and (need[1] or need[2])
        # the feature-table gradients lead: they are complete with the scatter, long before the decoders' weight gradients,
        # and a multi-GPU exchange can take that prefix early (GradPool.features, RenderSpec.after_feature_grads)
        h_dfg = zp.want((n_rows, C_DIM)) if need[4] else None
        h_dfc = zp.want((n_rows, C_DIM)) if (spec.color_stage and need[5]) else None
        zp.features = zp.total
        h_dg = zp.want_like(gprm, skip_first=not need[7]) if want_gw else None
        h_dcw = zp.want_like(cprm, skip_first=True) if want_cw else None   # colour _B is not a parameter
        m_need = need[7 + len(gprm) + len(cprm):] if mprm else ()
        h_daff = zp.want((12,)) if (aff is not None and (need[6] or any(m_need))) else None
        tc_dp = need_p and geo_kind == "tc"
        # the exposure MLP's gradients (loss-side or decoder-side) come out of the same allocation, so that a caller-owned
        # grad_buffer holds EVERY gradient of the iteration (one exchange, nothing left outside)
        h_exp = None
        if l_eprm and any(need[7 + len(gprm) + len(cprm):]):
            h_exp = zp.want((_exposure_grad_floats(l_eprm),))
        elif mprm and any(m_need):
            h_exp = zp.want((_exposure_grad_floats(mprm),))
        h_dpg = zp.want((P, 3)) if tc_dp else None                          # accumulated by the dgrad kernels
        h_dpc = zp.want((P, 3)) if (tc_dp and spec.color_stage) else None
        return zp, (h_dg, h_dcw, h_daff, h_dfg, h_dfc, h_exp, h_dpg, h_dpc)

    @staticmethod
    def backward(ctx, *gouts):
        if NVTX:
            torch.cuda.nvtx.range_push("hps.render_rays.bwd")
            try:
                return _RenderRays._backward(ctx, *gouts)
            finally:
                torch.cuda.nvtx.range_pop()
        return _RenderRays._backward(ctx, *gouts)

    @staticmethod
    def _backward(ctx, *gouts):
        spec = ctx.spec
        gscale = None
        if ctx.loss_state is not None:          # gradients of the fused loss, scaled by the upstream gradient in-kernel
            l_flat, l_a1, l_eprm, l_color, l_head, l_naff, l_aff_shape = ctx.loss_state
            l_R = ctx.saved_tensors[0].shape[0]
            gscale = _f32c(gouts[0]).reshape(1)
            g_depth = l_flat[l_head:l_head + l_R]
            g_color = l_flat[l_head + l_R:].view(l_R, 3) if l_color else None
        else:
            g_depth, _g_var, g_color, _g_valid = gouts
        (z, pts, I, w, has, c_geo, c_col, occ, rgb, act_g, act_c, fg, fc, aff, pos) = ctx.saved_tensors[:15]
        gprm = list(ctx.saved_tensors[15:15 + N_PARAM])
        cprm = list(ctx.saved_tensors[15 + N_PARAM:])
        need = ctx.needs_input_grad   # (spec, rays_o, rays_d, gt_depth, feats_geo, feats_col, affine, *params)
        dev = z.device
        R, S = z.shape
        P = R * S
        f32 = dict(dtype=torch.float32, device=dev)
        need_p = spec.tracker and (need[1] or need[2])
        gd = _f32c(g_depth) if g_depth is not None else None
        gc = _f32c(g_color) if (g_color is not None and spec.color_stage) else None
        fused_raw = ctx.fused_d_raw
        if fused_raw is not None:
            d_occ, d_rgb = fused_raw
        else:
            d_occ = torch.empty(P, **f32)
            d_rgb = torch.empty(P, 3, **f32) if spec.color_stage else None
        dc_g = torch.empty(P, C_DIM, **f32)
        dc_c = torch.empty(P, C_DIM, **f32) if spec.color_stage else None
        fork = MULTI_STREAM and TIMERS is None and spec.color_stage and ctx.saved_kinds == ("tc", "tc")
        if fused_raw is not None and not fork:
            raise _lib.HpsError("the forward fused the compositing backward but the backward is not on the multi-stream "
                                "tensor-core path (HPS_STREAMS / TIMERS changed between forward and backward?)")
        m_need = need[7 + len(gprm) + len(cprm):] if ctx.aff_mlp_state is not None else ()
        # every zero-initialised buffer of the backward comes out of one allocation, cleared by one memset kernel - or,
        # in a training iteration on the fused path, by the forward's compositing kernel already (ctx.zeroed_pool).
        # (Clearing it at the START of the forward, beside the decoders, was tried in round 2: the iteration gets 4-10 us
        # SLOWER - by the time the scatter and the weight-gradient kernels add into the buffer the decoders' 100 MB of
        # workspace writes have pushed the zero lines out of L2, and every red.add fetches from DRAM.)
        ready, ctx.zeroed_pool = ctx.zeroed_pool, None
        if ready is not None:
            zp, (h_dg, h_dcw, h_daff, h_dfg, h_dfc, h_exp, h_dpg, h_dpc) = ready
            if fork:
                main, side = torch.cuda.current_stream(dev), _side_stream(dev)
        else:
            zp, (h_dg, h_dcw, h_daff, h_dfg, h_dfc, h_exp, h_dpg, h_dpc) = _RenderRays._plan_pool(
                spec, need, gprm, cprm, aff, ctx.n_live if ctx.row_map is not None else fg.shape[0], P, dev,
                ctx.loss_state[2] if ctx.loss_state is not None else None,
                ctx.aff_mlp_state[0] if ctx.aff_mlp_state is not None else None, ctx.saved_kinds[0])
            with torch.cuda.device(dev):
                if fork:        # the memset runs beside the compositing backward instead of in front of it
                    main, side = torch.cuda.current_stream(dev), _side_stream(dev)
                    side.wait_stream(main)
                    zp.alloc(side, into=spec.grad_buffer)
                else:
                    zp.alloc(into=spec.grad_buffer)
        dg, dcw, daff = zp.get(h_dg), zp.get(h_dcw), zp.get(h_daff)
        dfg, dfc = zp.get(h_dfg), zp.get(h_dfc)
        dp_i = torch.empty(P, 3, **f32) if need_p else None
        st = _stream
        with torch.cuda.device(dev):
            # forward-time backward images exist => their max|dout| words are zeroed: reduce into them right here
            bpk_g, bpk_c = ctx.bwd_packed
            folded = fork and bpk_g is not None and bpk_c is not None
            if fused_raw is None:
                _call("hps_composite_bwd", _lib.ptr(rgb), 3, _lib.ptr(occ), 1, _lib.ptr(has), _lib.ptr(z), R, S,
                      float(spec.coef), _lib.ptr(gd), _lib.ptr(gc), _lib.ptr(gscale), _lib.ptr(d_rgb), 3, _lib.ptr(d_occ), 1,
                      _absmax_word(MLP_COL, bpk_c) if folded else None, _absmax_word(MLP_GEO, bpk_g) if folded else None,
                      st())
            up = gscale if fused_raw is not None else None      # d raw was written for an upstream gradient of 1
            dp_c = None
            wgrad_streams = ()
            x_s = g_exp_early = None
            if fork:
                dp_g, dp_c = zp.get(h_dpg), zp.get(h_dpc)
                main.wait_stream(side)          # the zeroed buffers
                side.wait_stream(main)
                if ctx.loss_state is not None and l_eprm:
                    # exposure MLP backward from d loss / d affine (accumulated by the forward's fused loss kernel): it
                    # depends on nothing in this backward, so it goes first on a stream of its own - at the end of the
                    # main stream (round 2 timeline) its single CTA was the last 5 us of the iteration
                    x_s = _side_stream(dev, 3)
                    x_s.wait_stream(main)
                    g_exp_early = _exposure_backward(l_eprm, l_a1, l_flat[4:], gscale, need[7 + len(gprm) + len(cprm):],
                                                     ctx.exp_x_shape, x_s, zp.get(h_exp))
                n_c, n_g = _sm_split(P, "bwd")
                # the HBM-bound weight-gradient kernels go to streams of their own: the feature scatter below needs
                # only dc of both decoders and overlaps with them
                # (highest priority: the two weight-gradient kernels are the critical path of the tail - the second can only
                # be placed once the first has left the SMs (shared memory) - and must not queue behind the scatter's or the
                # tail work's CTAs for shared memory and thread slots)
                w_c = _side_stream(dev, 5, priority=WGRAD_PRIORITY) if dcw is not None else None     # (no weight gradients: nothing to fork)
                w_g = _side_stream(dev, 2, priority=WGRAD_PRIORITY) if dg is not None else None
                # The colour dgrad kernel goes to a stream of higher priority than the geometry one's: its 98 CTAs own an
                # SM each and must all be placed before the geometry kernel's 100 half-size CTAs (two per SM) take the rest
                # - placed in the other order, the geometry CTAs spread one per SM and the colour kernel waits for them.
                c_s = _side_stream(dev, 7, priority=COLOR_DGRAD_PRIORITY)
                c_s.wait_stream(main)
                _mlp_backward(MLP_COL, cprm, "tc", act_c, pts, c_col, aff, d_rgb, P, dcw, dc_c, need_p, daff, dev, c_s,
                              dp_c, packed=ctx.bwd_packed[1], absmax_done=folded, max_ctas=n_c, wgrad_stream=w_c, gscale=up)
                e_s = None
                if any(m_need):            # d affine is complete with the colour dgrad: the MLP's backward runs beside
                    e_s = _side_stream(dev, 3)                                      # the feature scatter
                    e_s.wait_stream(c_s)
                    g_aff_mlp = _exposure_backward(ctx.aff_mlp_state[0], ctx.aff_mlp_state[1], daff, None, m_need,
                                                   ctx.aff_x_shape, e_s, zp.get(h_exp))
                _mlp_backward(MLP_GEO, gprm, "tc", act_g, pts, c_geo, None, d_occ, P, dg, dc_g, need_p, None, dev, side,
                              dp_g, packed=ctx.bwd_packed[0], absmax_done=folded, max_ctas=n_g, wgrad_stream=w_g, gscale=up)
                main.wait_stream(side)
                main.wait_stream(c_s)
                wgrad_streams = tuple(w for w in (w_c, w_g) if w is not None)
            else:
                dp_g = _mlp_backward(MLP_GEO, gprm, ctx.saved_kinds[0], act_g, pts, c_geo, None, d_occ, P, dg, dc_g,
                                     need_p, None, dev, dp=zp.get(h_dpg))
                if spec.color_stage:
                    dp_c = _mlp_backward(MLP_COL, cprm, ctx.saved_kinds[1], act_c, pts, c_col, aff, d_rgb, P, dcw, dc_c,
                                         need_p, daff, dev, dp=zp.get(h_dpc))
                e_s = None
                if any(m_need):
                    g_aff_mlp = _exposure_backward(ctx.aff_mlp_state[0], ctx.aff_mlp_state[1], daff, None, m_need,
                                                   ctx.aff_x_shape, None, zp.get(h_exp))
            t_s = None
            if spec.tail_work is not None and not TAIL_AFTER_SCATTER:
                # Work the caller wants issued beside the backward's HBM-bound tail (weight-gradient streams + feature
                # scatter), typically the parameter-free head of the NEXT iteration (Renderer.prepare_batch: sampling +
                # neighbour search, issue-bound and L2-resident; its CTAs fit beside the weight-gradient CTAs).  It starts
                # once both decoders' data gradients are done and joins the main stream before backward returns.
                if fork:
                    t_s = _side_stream(dev, 4, priority=TAIL_PRIORITY)
                    t_s.wait_stream(main)
                    with torch.cuda.stream(t_s):
                        spec.tail_work()
                else:
                    spec.tail_work()
            if dfg is not None or dfc is not None or need_p:
                # Beside tail work the scatter goes to a stream of higher priority than the tail's: it PRODUCES gradients
                # (the multi-GPU exchange waits for it), the tail work is only needed by the next iteration - at equal
                # priority the tail's grid, launched first, is placed first and the scatter finished last (round 2).
                # (only with an exchange hook: on one GPU nothing waits for the scatter and the plain order is 4 % faster)
                sc_s = _side_stream(dev, 6, priority=SCATTER_PRIORITY) \
                    if (fork and t_s is not None and (spec.after_grads is not None or SCATTER_STREAM_ALWAYS)) else None
                if sc_s is not None:
                    sc_s.wait_stream(main)
                _call("hps_interp_bwd", _lib.ptr(pts), _lib.ptr(I), _lib.ptr(w), _lib.ptr(has), _lib.ptr(pos),
                      _lib.ptr(fg), _lib.ptr(fc), _lib.ptr(c_geo), _lib.ptr(c_col), _lib.ptr(dc_g), _lib.ptr(dc_c), P,
                      _lib.ptr(dfg), _lib.ptr(dfc), _lib.ptr(dp_i), _lib.ptr(ctx.row_map), st(sc_s))
                if spec.after_feature_grads is not None:
                    # the feature-table gradients (the bulk of what a multi-GPU exchange moves) are complete HERE, while
                    # the weight-gradient kernels still run: the caller's hook can put them on the wire now
                    if sc_s is not None:
                        with torch.cuda.stream(sc_s):
                            spec.after_feature_grads()
                    else:
                        spec.after_feature_grads()
                if sc_s is not None:
                    main.wait_stream(sc_s)
            if spec.tail_work is not None and TAIL_AFTER_SCATTER:
                if fork:
                    t_s = _side_stream(dev, 4, priority=TAIL_PRIORITY)
                    t_s.wait_stream(w_c if w_c is not None else main)      # (experiment: behind the colour weight gradients)
                    with torch.cuda.stream(t_s):
                        spec.tail_work()
                else:
                    spec.tail_work()
            g_ro = g_rd = None
            if need_p:
                g_ro, g_rd = torch.empty(R, 3, **f32), torch.empty(R, 3, **f32)
                _call("hps_sample_rays_bwd", _lib.ptr(dp_g), _lib.ptr(dp_i), _lib.ptr(dp_c), _lib.ptr(z), R, S,
                      _lib.ptr(g_ro), _lib.ptr(g_rd), st())
            g_exp = ()
            if ctx.loss_state is not None and l_aff_shape is not None:
                # the loss-side affine was an input: its gradient, scaled by the upstream gradient (main stream has slack)
                g_exp = ((l_flat[4:4 + l_naff] * gscale).view(l_aff_shape) if need[7 + len(gprm) + len(cprm)] else None,)
            if ctx.loss_state is not None and l_eprm:
                if x_s is not None:
                    g_exp = g_exp_early
                    main.wait_stream(x_s)
                else:
                    g_exp = _exposure_backward(l_eprm, l_a1, l_flat[4:], gscale, need[7 + len(gprm) + len(cprm):],
                                               ctx.exp_x_shape, None, zp.get(h_exp))
            if ctx.aff_mlp_state is not None:
                g_exp = g_aff_mlp if any(m_need) else (None,) * 5
                if e_s is not None:
                    main.wait_stream(e_s)
            for w_s in wgrad_streams:
                main.wait_stream(w_s)
            if spec.after_grads is not None:
                # Every gradient of this backward is complete on the main stream here, while tail_work may still be
                # running on its own: the caller's hook - the multi-GPU exchange over grad_buffer - is issued in front of
                # the join, so that the exchange's latency hides behind the next batch's neighbour search.
                spec.after_grads()
            if t_s is not None:
                main.wait_stream(t_s)
        gg = [(g if (g is not None and need[7 + i]) else None) for i, g in enumerate(dg)] if dg else [None] * N_PARAM
        if spec.color_stage:
            gcw = [(g if (g is not None and need[7 + N_PARAM + i]) else None) for i, g in enumerate(dcw)] \
                if dcw else [None] * N_PARAM
        else:
            gcw = []
        return (None, g_ro if need[1] else None, g_rd if need[2] else None, None, dfg, dfc,
                daff if need[6] else None, *gg, *gcw, *g_exp)


def render_rays(spec, rays_o, rays_d, gt_depth, feats_geo, feats_col, affine, geo_params, col_params,
                exposure_params=None):
    """exposure_params (with spec.loss.exposure): (latent, W1, b1, W2, b2) of the exposure MLP; the call then returns
    (loss, depth, var, color, valid) with only the loss differentiable."""
    spec.live_geo = feats_geo if isinstance(feats_geo, LiveRows) else None
    spec.live_col = feats_col if (isinstance(feats_col, LiveRows) and spec.color_stage) else None
    if spec.live_geo is not None:
        feats_geo = spec.live_geo.param
    if isinstance(feats_col, LiveRows):
        feats_col = feats_col.param
    params = list(geo_params) + (list(col_params) if spec.color_stage else [])
    if spec.aff_mlp and spec.color_stage:
        params += list(exposure_params)      # decoder-side exposure MLP: (latent, W1, b1, W2, b2)
    elif isinstance(spec.loss, MapperLossSpec) and spec.loss.exposure and spec.color_stage \
            and spec.loss.gt_color is not None:
        params += list(exposure_params)      # "mlp": (latent, W1, b1, W2, b2); "affine": (affine,)
    return _RenderRays.apply(spec, rays_o, rays_d, gt_depth, feats_geo, feats_col if spec.color_stage else None,
                             affine if spec.color_stage else None, *params)


# --------------------------------------------------------------------------------------------------
# optimiser over live rows (Mapper.py:592-603, 859)
# --------------------------------------------------------------------------------------------------
class RowAdam:
    """torch.optim.Adam semantics (no amsgrad, no weight decay) for LiveRows leaves, one kernel per table and step,
    fused with the write-back of the new rows into the full tables - the mapper's `optimizer.step()` plus its
    per-iteration `table[indices] = rows` (Mapper.py:615-630, 859, 862-874).  `groups`: list of (LiveRows, lr).
    The step counter is a device tensor, so step() can be captured in a CUDA graph with the iteration."""

    class _Dense:
        """A whole (N,C) table as the leaf (first mapped frame / frustum_feature_selection off, Mapper.py:538-556):
        no separate table to write back into."""
        table = indices = None

        def __init__(self, param):
            _need_cuda(param)
            if param.dim() != 2 or param.shape[1] % 4 or param.dtype != torch.float32 or not param.is_contiguous():
                raise _lib.HpsError("RowAdam takes LiveRows or a contiguous fp32 (N,C) leaf with C % 4 == 0")
            self.param, self.fresh = param, False

    def __init__(self, groups, betas=(0.9, 0.999), eps=1e-8):
        self.groups = [[self._Dense(r) if torch.is_tensor(r) else r, float(lr)] for r, lr in groups]
        self.betas, self.eps = betas, float(eps)
        self.state = []
        for rows, _ in self.groups:
            p = rows.param
            self.state.append(dict(m=torch.zeros_like(p), v=torch.zeros_like(p),
                                   step=torch.zeros(2, dtype=torch.int32, device=p.device)))

    def set_lr(self, index, lr):
        self.groups[index][1] = float(lr)

    def zero_grad(self):
        for rows, _ in self.groups:
            rows.param.grad = None

    @torch.no_grad()
    def step(self):
        for (rows, lr), st in zip(self.groups, self.state):
            p = rows.param
            if p.grad is None:
                continue
            _need_cuda(p, p.grad)
            g = _f32c(p.grad)
            with torch.cuda.device(p.device):
                _call("hps_adam_rows", _lib.ptr(p), _lib.ptr(g), _lib.ptr(st["m"]), _lib.ptr(st["v"]), p.shape[0], p.shape[1],
                      lr, float(self.betas[0]), float(self.betas[1]), self.eps, _lib.ptr(st["step"]), _lib.ptr(rows.table),
                      _lib.ptr(rows.indices), _stream())
            rows.fresh = rows.table is not None


class FusedAdam:
    """torch.optim.Adam semantics (no amsgrad, no weight decay) for a few small dense tensors - the tracker's pose,
    exposure latent and exposure MLPs (Tracker.py:383-416) - as ONE kernel per step instead of the ~20 launches of
    torch's capturable foreach path.  `groups`: list of (tensor or list of tensors, lr) = parameter groups.
    own_grads=True: every gradient lives in one flat buffer (p.grad are views of it, created here and kept: backward
    accumulates into them) and is cleared by the step kernel itself, so zero_grad() costs nothing.
    own_grads=False: step() reads whatever p.grad holds (parameters without a gradient are skipped) and
    zero_grad() sets the gradients to None, so that backward hands its buffers over instead of adding into ours -
    one small kernel less per parameter; inside a CUDA graph the buffers keep their addresses.
    Device step counter (one per launch of up to 24 tensors, so a parameter that sits out an iteration shares the
    bias correction of the others): capturable.
    per_tensor_steps=True: one device counter per tensor instead, exactly torch.optim.Adam's `state[p]['step']` - what
    the mapper's optimiser needs, whose parameter list holds all four decoders although a stage gives gradients to
    one or two of them (Mapper.py:555-569, 592-603); use it with own_grads=False (tensors without .grad sit out)."""

    MAX = 24

    def __init__(self, groups, betas=(0.9, 0.999), eps=1e-8, own_grads=True, per_tensor_steps=False):
        self.own_grads = bool(own_grads)
        self.per_tensor_steps = bool(per_tensor_steps)
        self.params, self.lrs = [], []
        for tensors, lr in groups:
            for t in ([tensors] if torch.is_tensor(tensors) else list(tensors)):
                _need_cuda(t)
                if t.dtype != torch.float32 or not t.is_contiguous():
                    raise _lib.HpsError("FusedAdam takes contiguous fp32 tensors")
                self.params.append(t)
                self.lrs.append(float(lr))
        if not self.params:
            raise ValueError("no parameters")
        dev = self.params[0].device
        self.betas, self.eps = betas, float(eps)
        sizes = [(p.numel() + 3) // 4 * 4 for p in self.params]
        self._flat = torch.zeros(3, sum(sizes), dtype=torch.float32, device=dev)        # grad | exp_avg | exp_avg_sq
        self._views, off = [], 0
        for p, n in zip(self.params, sizes):
            g, m, v = (self._flat[k, off:off + p.numel()].view(p.shape) for k in range(3))
            if self.own_grads:
                p.grad = g
            self._views.append((g, m, v))
            off += n
        n_chunks = (len(self.params) + self.MAX - 1) // self.MAX
        self.step_count = torch.zeros(n_chunks, 2, dtype=torch.int32, device=dev)     # one counter per launch
        # per_tensor_steps: one counter per tensor; column 1 of step_count row c is launch c's scratch word
        self.tensor_steps = torch.zeros(len(self.params), dtype=torch.int32, device=dev) if self.per_tensor_steps else None

    def set_lr(self, index, lr):
        self.lrs[index] = float(lr)

    def zero_grad(self, set_to_none=None):
        """own_grads: nothing to do after a step() (the kernel cleared the gradients).  Otherwise: gradients -> None."""
        if self.own_grads:
            if set_to_none:
                raise _lib.HpsError("FusedAdam(own_grads=True) keeps its gradient buffers (views of one allocation)")
            return
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self):
        ps, gs, ms, vs, lr_of, which = [], [], [], [], [], []
        for k, (p, (g, m, v), lr) in enumerate(zip(self.params, self._views, self.lrs)):
            if self.own_grads:
                if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                    raise _lib.HpsError("FusedAdam: a .grad was replaced; keep the buffers (or use own_grads=False)")
            else:
                if p.grad is None:
                    continue
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = p.grad = g.float().contiguous()
            ps.append(p); gs.append(g); ms.append(m); vs.append(v); lr_of.append(lr); which.append(k)
        if not ps:
            return
        dev = self.params[0].device
        with torch.cuda.device(dev):
            for s in range(0, len(ps), self.MAX):
                e = min(s + self.MAX, len(ps))
                n = e - s
                numel = (ctypes.c_int64 * n)(*[p.numel() for p in ps[s:e]])
                lrs = (ctypes.c_float * n)(*lr_of[s:e])
                arrs = [(ctypes.c_void_p * n)(*[t.data_ptr() for t in ts[s:e]]) for ts in (ps, gs, ms, vs)]
                if self.per_tensor_steps:
                    base = self.tensor_steps.data_ptr()
                    steps = (ctypes.c_void_p * n)(*[base + 4 * k for k in which[s:e]])
                    _call("hps_adam_multi_own_steps", n, arrs[0], arrs[1], arrs[2], arrs[3], numel, lrs, steps,
                          float(self.betas[0]), float(self.betas[1]), self.eps,
                          ctypes.c_void_p(self.step_count.data_ptr() + 4 * (2 * (s // self.MAX) + 1)),
                          1 if self.own_grads else 0, _stream())
                    continue
                _call("hps_adam_multi", n, arrs[0], arrs[1], arrs[2], arrs[3], numel, lrs, float(self.betas[0]),
                      float(self.betas[1]), self.eps, _lib.ptr(self.step_count[s // self.MAX]),
                      1 if self.own_grads else 0, _stream())


# --------------------------------------------------------------------------------------------------
# tracker pose helpers (common.py:304-346, 103-119)
# --------------------------------------------------------------------------------------------------
class _CameraFromTensor(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cam):
        _need_cuda(cam)
        c = _f32c(cam).reshape(7)
        RT = torch.empty(3, 4, dtype=torch.float32, device=c.device)
        with torch.cuda.device(c.device):
            _call("hps_camera_from_tensor", _lib.ptr(c), _lib.ptr(RT), _stream())
        ctx.save_for_backward(c)
        return RT

    @staticmethod
    def backward(ctx, g):
        (c,) = ctx.saved_tensors
        out = torch.empty(7, dtype=torch.float32, device=c.device)
        with torch.cuda.device(c.device):
            _call("hps_camera_from_tensor_bwd", _lib.ptr(c), _lib.ptr(_f32c(g)), _lib.ptr(out), _stream())
        return out


def camera_from_tensor(cam):
    """(7,) [quaternion, translation] CUDA tensor -> (3,4); one kernel each way."""
    return _CameraFromTensor.apply(cam)


class _RaysFromUv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, i, j, c2w, fx, fy, cx, cy):
        _need_cuda(i, j, c2w)
        dev = c2w.device
        pi, pj, m = _f32c(i).reshape(-1), _f32c(j).reshape(-1), _f32c(c2w)
        if m.dim() != 2 or m.shape[1] != 4 or m.shape[0] < 3:
            raise ValueError("c2w must be (3,4) or (4,4)")
        R = pi.numel()
        ro = torch.empty(R, 3, dtype=torch.float32, device=dev)
        rd = torch.empty(R, 3, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _call("hps_rays_from_uv", _lib.ptr(pi), _lib.ptr(pj), _lib.ptr(m), 4, R, float(fx), float(fy), float(cx),
                  float(cy), _lib.ptr(ro), _lib.ptr(rd), _stream())
        ctx.save_for_backward(pi, pj)
        ctx.k = (float(fx), float(fy), float(cx), float(cy), tuple(c2w.shape))
        return ro, rd

    @staticmethod
    def backward(ctx, g_o, g_d):
        pi, pj = ctx.saved_tensors
        fx, fy, cx, cy, shape = ctx.k
        g = torch.zeros(shape, dtype=torch.float32, device=pi.device)
        with torch.cuda.device(pi.device):
            _call("hps_rays_from_uv_bwd", _lib.ptr(pi), _lib.ptr(pj), _lib.ptr(_f32c(g_o) if g_o is not None else None),
                  _lib.ptr(_f32c(g_d) if g_d is not None else None), pi.numel(), fx, fy, cx, cy, _lib.ptr(g), _stream())
        return None, None, g, None, None, None, None


def rays_from_uv(i, j, c2w, fx, fy, cx, cy):
    """Rays through pixels (i, j) of camera c2w ((3,4) or (4,4) CUDA tensor): rays_o (R,3), rays_d (R,3)."""
    return _RaysFromUv.apply(i, j, c2w, fx, fy, cx, cy)


# --------------------------------------------------------------------------------------------------
# exposure MLP (decoder.py:380-396)
# --------------------------------------------------------------------------------------------------
class _ExposureMlp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W1, b1, W2, b2):
        _need_cuda(x, W1, b1, W2, b2)
        dev = x.device
        xs, w1, bb1, w2, bb2 = _f32c(x).reshape(-1), _f32c(W1), _f32c(b1), _f32c(W2), _f32c(b2)
        hidden, n_in = w1.shape
        n_out = w2.shape[0]
        if xs.numel() != n_in:
            raise ValueError("exposure feature has %d entries, the MLP takes %d" % (xs.numel(), n_in))
        a1 = torch.empty(hidden, dtype=torch.float32, device=dev)
        out = torch.empty(n_out, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _call("hps_exposure_fwd", _lib.ptr(xs), _lib.ptr(w1), _lib.ptr(bb1), _lib.ptr(w2), _lib.ptr(bb2), n_in, hidden,
                  n_out, _lib.ptr(a1), _lib.ptr(out), _stream())
        ctx.save_for_backward(xs, w1, w2, a1)
        ctx.x_shape = x.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        xs, w1, w2, a1 = ctx.saved_tensors
        hidden, n_in = w1.shape
        n_out = w2.shape[0]
        dev = xs.device
        # one allocation: dW1 | db1 | dW2 | db2 | dx (all fully written by the kernel)
        sizes = [hidden * n_in, hidden, n_out * hidden, n_out, n_in]
        flat = torch.empty(sum((n + 3) // 4 * 4 for n in sizes), dtype=torch.float32, device=dev)
        parts, off = [], 0
        for n in sizes:
            parts.append(flat[off:off + n])
            off += (n + 3) // 4 * 4
        dW1, db1, dW2, db2, dx = parts
        want_x = ctx.needs_input_grad[0]
        with torch.cuda.device(dev):
            _call("hps_exposure_bwd", _lib.ptr(xs), _lib.ptr(w1), _lib.ptr(w2), _lib.ptr(a1), _lib.ptr(_f32c(dout)), None,
                  n_in, hidden, n_out, _lib.ptr(dW1), _lib.ptr(db1), _lib.ptr(dW2), _lib.ptr(db2),
                  _lib.ptr(dx) if want_x else None, _stream())
        return (dx.view(ctx.x_shape) if want_x else None, dW1.view(hidden, n_in), db1, dW2.view(n_out, hidden), db2)


def exposure_mlp(x, W1, b1, W2, b2):
    """MLP_exposure.forward on one latent vector: W2 softplus_100(W1 x + b1) + b2, one kernel each way."""
    return _ExposureMlp.apply(x, W1, b1, W2, b2)


# --------------------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------------------
class _MapperLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, color, valid, gt_depth, gt_color, affine, w_color, color_stage, frame):
        _need_cuda(depth, valid, gt_depth)
        dev, R = depth.device, depth.shape[0]
        f32 = dict(dtype=torch.float32, device=dev)
        d, gdp = _f32c(depth), _f32c(gt_depth).reshape(-1)
        col = _f32c(color) if color_stage else None
        gcl = _f32c(gt_color) if color_stage else None
        aff = _f32c(affine) if (affine is not None and color_stage) else None
        frm, n_aff = _loss_frames(aff, frame, R)
        v = valid.detach().contiguous()
        v = v.view(torch.uint8) if v.dtype == torch.bool else v.to(torch.uint8)
        # one buffer: [loss, pad x3 | d_affine (12 per frame) | g_depth (R) | g_color (3R)]; the head is zeroed
        # (accumulators)
        head = _loss_head_floats(n_aff)
        flat = torch.empty(head + R * (4 if color_stage else 1), **f32)
        flat[:head].zero_()
        loss = flat[0]
        d_aff = flat[4:4 + n_aff] if aff is not None else None
        g_depth = flat[head:head + R]
        g_color = flat[head + R:].view(R, 3) if color_stage else None
        with torch.cuda.device(dev):
            _call("hps_mapper_loss", _lib.ptr(d), _lib.ptr(col), _lib.ptr(v), _lib.ptr(gdp), _lib.ptr(gcl), _lib.ptr(aff),
                  _lib.ptr(frm), float(w_color), R, _lib.ptr(loss), _lib.ptr(g_depth), _lib.ptr(g_color), _lib.ptr(d_aff),
                  _stream())
        ctx.save_for_backward(flat)
        ctx.dims = (R, color_stage, affine.shape if aff is not None else None, head, n_aff)
        return loss

    @staticmethod
    def backward(ctx, gout):
        (flat,) = ctx.saved_tensors
        R, color_stage, aff_shape, head, n_aff = ctx.dims
        g = flat[4:] * gout                      # one kernel scales d_affine, g_depth and g_color
        return (g[head - 4:head - 4 + R], g[head - 4 + R:].view(R, 3) if color_stage else None, None, None, None,
                g[:n_aff].view(aff_shape) if aff_shape is not None else None, None, None, None)


def _loss_head_floats(n_aff):
    """[loss, pad x3 | d_affine (n_aff)] rounded up to 16 floats."""
    return max(16, (4 + n_aff + 15) // 16 * 16)


def _loss_frames(aff, frame, R):
    """Checks the (affine, frame index) pair of a multi-keyframe batch -> (int32 frame tensor or None, #affine floats)."""
    if aff is None:
        return None, 0
    if aff.numel() % 12 != 0:
        raise ValueError("exposure affine: 12 values per frame")
    n_frames = aff.numel() // 12
    if frame is None:
        if n_frames != 1:
            raise ValueError("%d exposure affines but no frame index per ray" % n_frames)
        return None, 12
    frm = frame.detach().reshape(-1).to(device=aff.device, dtype=torch.int32).contiguous()
    if frm.numel() != R:
        raise ValueError("frame index: one entry per ray")
    return frm, 12 * n_frames


def mapper_loss(depth, color, valid_ray_mask, gt_depth, gt_color, stage, w_color_loss=0.1, affine=None,
                frame_index=None):
    """Mapper.py:791-856 in one kernel (loss + gradients).  affine: (12,) = one frame's exposure correction, or (F,12)
    with frame_index (R,) in [0,F) for a batch that mixes rays of F keyframes (Mapper.py:805-822; the kernel does not
    range-check the indices)."""
    cs = stage.startswith("color")
    return _MapperLoss.apply(depth, color if cs else None, valid_ray_mask, gt_depth, gt_color if cs else None,
                             affine if cs else None, w_color_loss, cs, frame_index if cs else None)


class _TrackerLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, uncertainty, color, gt_depth, gt_color, w_color, reduce_fn):
        _need_cuda(depth, uncertainty, gt_depth)
        dev, R = depth.device, depth.shape[0]
        f32 = dict(dtype=torch.float32, device=dev)
        d, u, gdp = _f32c(depth), _f32c(uncertainty), _f32c(gt_depth).reshape(-1)
        col = _f32c(color) if color is not None else None
        gcl = _f32c(gt_color) if color is not None else None
        stats = torch.full((2,), float(R), **f32)      # [sum of tmp (accumulated by the kernel), number of rays]
        stats[:1].zero_()                              # (no scalar assignment: that is a host-to-device copy, which a
                                                       #  CUDA-graph capture of the tracker iteration would refuse)
        loss = torch.zeros((), **f32)
        g_depth = torch.empty(R, **f32)
        g_color = torch.empty(R, 3, **f32) if col is not None else None
        with torch.cuda.device(dev):
            _call("hps_tracker_loss_stats", _lib.ptr(d), _lib.ptr(u), _lib.ptr(gdp), R, _lib.ptr(stats), _stream())
            count = float(R)
            if reduce_fn is not None:      # rays sharded over ranks: global sum and count
                count = float(reduce_fn(stats))
            _call("hps_tracker_loss", _lib.ptr(d), _lib.ptr(u), _lib.ptr(col), _lib.ptr(gdp), _lib.ptr(gcl),
                  _lib.ptr(stats), count, float(w_color), R, _lib.ptr(loss), _lib.ptr(g_depth), _lib.ptr(g_color),
                  _stream())
        ctx.save_for_backward(g_depth, g_color)
        return loss

    @staticmethod
    def backward(ctx, gout):
        g_depth, g_color = ctx.saved_tensors
        return g_depth * gout, None, g_color * gout if g_color is not None else None, None, None, None, None


def tracker_loss(depth, uncertainty, color, gt_depth, gt_color, w_color_loss=0.5, reduce_fn=None):
    """Tracker.py:213-231 in two kernels.  reduce_fn(stats[2] = [sum_tmp, n_rays]) all-reduces in place across
    ranks and returns the global ray count when the batch is sharded."""
    return _TrackerLoss.apply(depth, uncertainty, color, gt_depth, gt_color, w_color_loss, reduce_fn)
