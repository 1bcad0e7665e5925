"""CUDA-graph capture of a whole render + loss + backward iteration.

One mapper / tracker iteration is ~20 kernel launches of 10-900 us each plus a handful of tiny PyTorch ops; issued
eagerly it is host-bound (~2 ms of Python per iteration for ~1.5 ms of GPU work).  Because nothing on the path
synchronises with the host (no boolean indexing, no .item(), neighbour index prebuilt), the whole iteration -
including autograd's backward and an NCCL all-reduce - can be captured once and replayed.
"""
import torch


class GraphedStep:
    """fn() must read its per-iteration inputs only from `static_inputs` (a dict of preallocated CUDA tensors) and
    leave its results in tensors it returns / in `.grad` fields; call the object with new input tensors.

    A captured graph holds DEVICE ADDRESSES: of the feature tables, the position buffer and - inside the neighbour
    search's launch parameters - the voxel grid's arrays and hash mask.  Point insertion keeps all of them (tables and
    positions grow inside capacity buffers, the grid is appended to in place), but a capacity doubling, a grid rebuild,
    set_cloud() or a whole-table update_*_feats() moves them, and replaying an older graph would then read freed or stale
    memory.  Pass `guard` = a callable returning a value that changes when that happens - e.g.
    `lambda: npc.storage_generation(level)` - and replay raises instead; re-capture (build a new GraphedStep) then."""

    def __init__(self, fn, static_inputs, warmup=3, before_capture=None, guard=None):
        self.fn = fn
        self.static = static_inputs
        self.guard = guard
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                out = fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if before_capture is not None:
            before_capture()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = fn()
        self._guard_value = guard() if guard is not None else None
        del out

    def load(self, **inputs):
        for k, v in inputs.items():
            self.static[k].copy_(v, non_blocking=True)

    def check(self):
        if self.guard is not None and self.guard() != self._guard_value:
            raise RuntimeError("GraphedStep: device storage captured by this graph has moved since capture (%r -> %r); "
                               "capture the iteration again" % (self._guard_value, self.guard()))

    def __call__(self, **inputs):
        self.check()
        self.load(**inputs)
        self.graph.replay()
        return self.outputs
