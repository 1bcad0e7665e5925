"""The mapper's optimiser and its stage schedule, the step right after the render path.

Reference: `Mapper.optimize_map` builds ONE `torch.optim.Adam` over seven parameter groups - decoders, the four
feature tables (or their frustum-selected rows), the keyframe poses (BA) and the exposure latent - all created with
lr 0 (src/Mapper.py:590-603), then rewrites the groups' learning rates every iteration from the stage the iteration
falls into (:637-667) before `loss.backward(); optimizer.step()` (:858-859).

Here the same thing is two kinds of kernels: `RowAdam` (one launch per feature table, fused with the write-back of
the rows into the table the gather kernels read) and `FusedAdam(per_tensor_steps=True)` (one launch per 24 small
tensors: decoder weights, poses, exposure latent).  Both keep torch.optim.Adam's arithmetic and its per-parameter
step counters, and both skip parameters the stage gave no gradient - so a run follows the reference's optimiser
state step for step.  The schedule functions are pure host arithmetic, restated from the cited lines.
"""
from . import _lib
from .ops import FusedAdam, RowAdam

STAGES = ("geometry_mid", "color_mid", "geometry_fine", "color_fine")


def joint_iters_when_adding(num_joint_iters, frame_pts_add, min_iter_ratio):
    """more_iters_when_adding (Mapper.py:605-608): np.clip(int(n * added / 300), int(ratio * n), 2 n)."""
    lo, hi = int(min_iter_ratio * num_joint_iters), 2 * num_joint_iters
    return max(lo, min(int(num_joint_iters * frame_pts_add / 300), hi))


def split_iters(num_joint_iters, mid_iter_ratio):
    """(num_mid_iters, num_fine_iters), Mapper.py:611-612."""
    return int(num_joint_iters * mid_iter_ratio), int(num_joint_iters * (1 - mid_iter_ratio))


def mapper_stage(joint_iter, num_joint_iters, mid_iter_ratio, geo_iter_ratio, init=False, geo_iter_first=0):
    """Stage of iteration `joint_iter` (Mapper.py:637-644; the comparisons are `<=`, so every boundary iteration
    belongs to the earlier stage)."""
    num_mid, num_fine = split_iters(num_joint_iters, mid_iter_ratio)
    if joint_iter <= (geo_iter_first if init else int(num_mid * geo_iter_ratio)):
        return "geometry_mid"
    if joint_iter <= int(num_mid):
        return "color_mid"
    if joint_iter <= int(num_mid + num_fine * geo_iter_ratio):
        return "geometry_fine"
    return "color_fine"


def stage_learning_rates(mapping_cfg, stage, init=False, color_refine_last=False):
    """Learning rates of the five always-present groups, in the reference's group order (Mapper.py:648-660).
    color_refine_last: the last frame's colour refinement (`idx == n_img-1 and color_refine`)."""
    table = mapping_cfg["init" if init else "stage"]
    if color_refine_last:
        c = table["color_fine"]
        return {"decoders": c["decoders_lr"], "geometry_mid": 0.0, "geometry_fine": 0.0,
                "color_mid": c["color_lr"] / 10.0, "color_fine": c["color_lr"] / 10.0}
    c = table[stage]
    return {"decoders": c["decoders_lr"], "geometry_mid": c["geometry_mid_lr"], "geometry_fine": c["geometry_fine_lr"],
            "color_mid": c["color_lr"], "color_fine": c["color_lr"]}


def ba_camera_lr(joint_iter, num_joint_iters, mid_iter_ratio, geo_iter_ratio, ba_cam_lr):
    """Pose learning rate under BA: on in two windows, 20-30 % into each level's colour phase (Mapper.py:662-667)."""
    num_mid, num_fine = split_iters(num_joint_iters, mid_iter_ratio)
    if num_mid * (geo_iter_ratio + 0.2) <= joint_iter <= num_mid * (geo_iter_ratio + 0.3):
        return ba_cam_lr
    if joint_iter <= num_mid:
        return 0.0
    if num_mid + num_fine * (geo_iter_ratio + 0.2) <= joint_iter <= num_mid + num_fine * (geo_iter_ratio + 0.3):
        return ba_cam_lr
    return 0.0


class MapperOptimizer:
    """`torch.optim.Adam(optim_para_list)` of Mapper.py:592-603 with the per-iteration learning-rate rewrite of
    :646-667, on this package's kernels.

      decoder_params : list of decoder Parameters (the reference's `decoders_para_list`, Mapper.py:555-569)
      rows           : dict geometry_mid / geometry_fine / color_mid / color_fine -> `LiveRows` (frustum-selected
                       rows) or a full (N,32) leaf tensor; missing keys are simply not optimised
      cameras        : list of 7-vectors under BA (group 5), exposure_feat: the latent (last group, lr 0.001)

    set_iteration(joint_iter) returns the stage and installs its learning rates; step() runs the kernels (no host
    synchronisation, capturable in a CUDA graph - learning rates are launch arguments, so capture one graph per stage);
    zero_grad() drops the gradients like `optimizer.zero_grad(set_to_none=True)`."""

    ROW_KEYS = ("geometry_mid", "geometry_fine", "color_mid", "color_fine")

    def __init__(self, mapping_cfg, decoder_params=(), rows=None, cameras=(), exposure_feat=None, mid_iter_ratio=None,
                 exposure_lr=0.001, betas=(0.9, 0.999), eps=1e-8):
        self.cfg = mapping_cfg
        self.mid_iter_ratio = mapping_cfg["mid_iter_ratio"] if mid_iter_ratio is None else mid_iter_ratio
        self.geo_iter_ratio = mapping_cfg["geo_iter_ratio"]
        rows = dict(rows or {})
        unknown = set(rows) - set(self.ROW_KEYS)
        if unknown:
            raise ValueError("unknown feature tables: %s" % sorted(unknown))
        self.row_keys = [k for k in self.ROW_KEYS if k in rows]
        self.row_opt = RowAdam([(rows[k], 0.0) for k in self.row_keys], betas, eps) if self.row_keys else None
        groups, self._slot = [], {}

        def add(name, tensors, lr):
            tensors = [t for t in tensors if t is not None]
            if tensors:
                first = sum(len(g[0]) for g in groups)
                self._slot[name] = range(first, first + len(tensors))
                groups.append((tensors, lr))

        add("decoders", list(decoder_params), 0.0)
        add("cameras", list(cameras), 0.0)
        add("exposure", [exposure_feat], exposure_lr)
        self.small_opt = FusedAdam(groups, betas, eps, own_grads=False, per_tensor_steps=True) if groups else None
        if self.row_opt is None and self.small_opt is None:
            raise ValueError("nothing to optimise")
        self.stage = None

    def _set_small(self, name, lr):
        for k in self._slot.get(name, ()):
            self.small_opt.set_lr(k, lr)

    def set_learning_rates(self, lrs, camera_lr=None):
        """lrs: dict as returned by stage_learning_rates."""
        for i, k in enumerate(self.row_keys):
            self.row_opt.set_lr(i, lrs[k])
        self._set_small("decoders", lrs["decoders"])
        if camera_lr is not None:
            self._set_small("cameras", camera_lr)

    def set_iteration(self, joint_iter, num_joint_iters, init=False, color_refine_last=False, BA=False):
        """Mapper.py:637-667 for one iteration: pick the stage, install its learning rates, return the stage name."""
        self.stage = mapper_stage(joint_iter, num_joint_iters, self.mid_iter_ratio, self.geo_iter_ratio, init,
                                  self.cfg.get("geo_iter_first", 0))
        cam_lr = None
        if BA:
            cam_lr = ba_camera_lr(joint_iter, num_joint_iters, self.mid_iter_ratio, self.geo_iter_ratio,
                                  self.cfg["BA_cam_lr"])
        self.set_learning_rates(stage_learning_rates(self.cfg, self.stage, init, color_refine_last), cam_lr)
        return self.stage

    def zero_grad(self, set_to_none=True):
        if not set_to_none:
            raise _lib.HpsError("MapperOptimizer.zero_grad drops the gradients (set_to_none=True): backward hands its "
                                "buffers over instead of accumulating")
        if self.row_opt is not None:
            self.row_opt.zero_grad()
        if self.small_opt is not None:
            self.small_opt.zero_grad()

    def step(self):
        if self.row_opt is not None:
            self.row_opt.step()
        if self.small_opt is not None:
            self.small_opt.step()
