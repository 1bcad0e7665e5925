"""The tracker's optimiser and per-frame bookkeeping, the steps around the tracker's render call.

Reference: for every tracked frame `Tracker.run` builds a fresh `torch.optim.Adam` over the pose - one 7-vector at
`tracking.lr`, or translation at `lr` and quaternion at `lr / 5` with `separate_LR` - plus, with `encode_exposure`, the
exposure latent and both exposure MLPs at 0.001 (src/Tracker.py:383-415).  It then runs `tracking.iters` iterations, the
first `ratio_iter_mid` = half of them against the mid level, the rest against the fine level (:190-212), reads the
loss back to the host every iteration and keeps the pose of the iteration with the smallest loss (:426-440).

Here the optimiser is one `FusedAdam` launch per step (ten small tensors in the ScanNet configuration) and the
keep-the-best-pose bookkeeping stays on the device (two `torch.where`), so a tracked frame needs no host
synchronisation until its result is read.  Schedule and bookkeeping are plain host / tensor code, tested on the CPU
against the reference's statements; the Adam kernel is `hps_adam_multi` (tests/test_gpu_kernels.py).
"""
import torch

from .ops import FusedAdam

RATIO_ITER_MID = 0.5            # Tracker.py:40


def tracker_stage(cam_iter, num_cam_iters, ratio_iter_mid=RATIO_ITER_MID):
    """Tracker.py:427, 190-212: `progress = cam_iter / num_cam_iters`; mid level while progress < ratio_iter_mid."""
    return "color_mid" if cam_iter / num_cam_iters < ratio_iter_mid else "color_fine"


def tracker_param_groups(cam_lr, camera_tensor, separate_LR=False, exposure_feat=None, decoders=None,
                         encode_exposure=True, exposure_lr=0.001):
    """Parameter groups of Tracker.py:383-414 as (tensors, lr) pairs, in the reference's order, and the pose leaves.
    camera_tensor: (7,) quaternion (w,x,y,z) + translation.  separate_LR: translation and quaternion become two leaves
    (T at cam_lr, quad at cam_lr * 0.2) and the pose used by an iteration is cat([quad, T])."""
    cam = camera_tensor.detach().clone()
    if separate_LR:
        T = cam[-3:].clone().requires_grad_(True)
        quad = cam[:4].clone().requires_grad_(True)
        groups, leaves = [([T], cam_lr), ([quad], cam_lr * 0.2)], (quad, T)
    else:
        cam.requires_grad_(True)
        groups, leaves = [([cam], cam_lr)], (cam,)
    if encode_exposure:
        if exposure_feat is None or decoders is None:
            raise ValueError("encode_exposure: pass the exposure latent and the decoders")
        groups.append(([exposure_feat], exposure_lr))
        groups.append((list(decoders.color_decoder_mid.mlp_exposure.parameters()), exposure_lr))
        groups.append((list(decoders.color_decoder_fine.mlp_exposure.parameters()), exposure_lr))
    return groups, leaves


class TrackerOptimizer:
    """One tracked frame's optimiser (Tracker.py:383-415) and best-pose bookkeeping (:426-440).

        opt = TrackerOptimizer(cfg['tracking'], camera_tensor, exposure_feat, decoders)
        for cam_iter in range(opt.num_cam_iters):
            stage = opt.stage(cam_iter)
            cam = opt.camera_tensor()                     # (7,), differentiable
            loss = ...render + tracker loss...; loss.backward()
            opt.step(loss)                                # Adam step + keep the pose with the smallest loss
        best = opt.best_camera_tensor                     # device tensor; no host synchronisation so far
    """

    def __init__(self, tracking_cfg, camera_tensor, exposure_feat=None, decoders=None, encode_exposure=True,
                 ratio_iter_mid=RATIO_ITER_MID):
        self.num_cam_iters = int(tracking_cfg["iters"])
        self.ratio_iter_mid = ratio_iter_mid
        self.separate_LR = bool(tracking_cfg.get("separate_LR", False))
        groups, self.leaves = tracker_param_groups(float(tracking_cfg["lr"]), camera_tensor, self.separate_LR,
                                                   exposure_feat, decoders, encode_exposure)
        # own_grads=False: backward hands its gradient buffers over, zero_grad() drops them (Tracker.py:233-235).
        # per_tensor_steps: torch.optim.Adam keeps state[p]['step'] per parameter and skips parameters without a
        # gradient - the fine level's exposure MLP sees its first gradient at cam_iter = iters / 2 (the mid level renders
        # first, :190-212) and must then start its bias correction at t = 1, not at t = iters / 2 + 1.
        self.opt = FusedAdam(groups, own_grads=False, per_tensor_steps=True)
        dev = camera_tensor.device
        self.best_loss = torch.full((), float(1e20), device=dev)          # current_min_loss, Tracker.py:419
        self.best_camera_tensor = camera_tensor.detach().clone()          # candidate_cam_tensor
        self._cam_of_iteration = None

    def stage(self, cam_iter):
        return tracker_stage(cam_iter, self.num_cam_iters, self.ratio_iter_mid)

    def camera_tensor(self):
        """The pose this iteration renders with: the leaf, or cat([quad, T]) under separate_LR (Tracker.py:428-430)."""
        cam = torch.cat(self.leaves, 0) if self.separate_LR else self.leaves[0]
        self._cam_of_iteration = cam
        return cam

    @torch.no_grad()
    def step(self, loss):
        """optimizer.step(); optimizer.zero_grad() (Tracker.py:234-235), then the bookkeeping of :435-438: the pose of
        the iteration with the smallest loss so far.  As in the reference, that is the pose the loss was computed with
        under separate_LR (the concatenation made before the step) and the already-updated leaf otherwise."""
        self.opt.step()
        self.opt.zero_grad()
        cam = self._cam_of_iteration if self.separate_LR else self.leaves[0]
        if cam is None:
            raise RuntimeError("call camera_tensor() for every iteration")
        better = loss.detach() < self.best_loss
        # in place: inside a CUDA graph the two buffers keep their addresses from replay to replay
        self.best_camera_tensor.copy_(torch.where(better, cam.detach(), self.best_camera_tensor))
        self.best_loss.copy_(torch.where(better, loss.detach().to(self.best_loss.dtype), self.best_loss))
