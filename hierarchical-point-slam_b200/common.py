"""Drop-ins for the hot-path helpers of the reference's src/common.py.

raw2outputs_nerf_color (common.py:385-428) runs in the compositing kernel; the ray / pose helpers
(common.py:103-119, 304-346, 431-448) stay small torch expressions so that autograd chains the kernels'
d/d rays_o, d/d rays_d into the quaternion - they are written device-agnostically
(the reference's `.to(quad.get_device())` is CUDA-only).
"""
import numpy as np
import torch

from . import ops


def raw2outputs_nerf_color(raw, z_vals, rays_d=None, occupancy=False, device="cuda:0", coef=0.1):
    """raw (R,S,4) [rgb, occupancy logit], z_vals (R,S) -> depth (R,), depth_var (R,), rgb (R,3), weights (R,S).

    Only occupancy=True exists in the reference configuration (cfg['occupancy'] True, point_slam.yaml:4).
    Like the reference (common.py:414) the occupancy channel of `raw` is overwritten with alpha."""
    if not occupancy:
        raise NotImplementedError("volume-density compositing (occupancy=False) is dead code in the reference cfg")
    depth, var, rgb, weights = ops.composite(raw, z_vals.float(), None, coef, want_weights=True)
    with torch.no_grad():
        raw[..., -1] = torch.sigmoid(coef * raw[..., -1])
    return depth, var, rgb, weights


def quad2rotation(quad):
    """(B,4) quaternion (w,x,y,z) -> (B,3,3); common.py:304-328."""
    qr, qi, qj, qk = quad[:, 0], quad[:, 1], quad[:, 2], quad[:, 3]
    two_s = 2.0 / (quad * quad).sum(-1)
    rows = [
        torch.stack([1 - two_s * (qj ** 2 + qk ** 2), two_s * (qi * qj - qk * qr), two_s * (qi * qk + qj * qr)], -1),
        torch.stack([two_s * (qi * qj + qk * qr), 1 - two_s * (qi ** 2 + qk ** 2), two_s * (qj * qk - qi * qr)], -1),
        torch.stack([two_s * (qi * qk - qj * qr), two_s * (qj * qk + qi * qr), 1 - two_s * (qi ** 2 + qj ** 2)], -1),
    ]
    return torch.stack(rows, 1)


def get_camera_from_tensor(inputs):
    """(7,) or (B,7) [quat, T] -> (3,4) or (B,3,4); common.py:330-346."""
    single = inputs.dim() == 1
    if single and inputs.is_cuda and inputs.numel() == 7 and inputs.dtype == torch.float32:
        return ops.camera_from_tensor(inputs)              # one kernel each way (the tracker's per-iteration call)
    if single:
        inputs = inputs.unsqueeze(0)
    RT = torch.cat([quad2rotation(inputs[:, :4]), inputs[:, 4:, None]], 2)
    return RT[0] if single else RT


def get_rays_from_uv(i, j, c2w, H, W, fx, fy, cx, cy, device):
    """common.py:103-119."""
    if isinstance(c2w, np.ndarray):
        c2w = torch.from_numpy(c2w).to(device)
    if (torch.is_tensor(i) and i.is_cuda and i.dim() == 1 and i.is_floating_point() and c2w.is_cuda
            and c2w.dtype == torch.float32 and c2w.dim() == 2 and c2w.shape[1] == 4):
        return ops.rays_from_uv(i, j, c2w, fx, fy, cx, cy)  # one kernel each way, bit-identical to the expressions below
    dirs = torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1).to(device).reshape(-1, 1, 3)
    rays_d = torch.sum(dirs * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def get_rays(H, W, fx, fy, cx, cy, c2w, device, crop_edge=0):
    """common.py:431-448 (pixel grid built directly on `device`)."""
    if isinstance(c2w, np.ndarray):
        c2w = torch.from_numpy(c2w)
    c2w = c2w.to(device)
    i, j = torch.meshgrid(torch.linspace(crop_edge, W - 1 - crop_edge, W - 2 * crop_edge, device=device),
                          torch.linspace(crop_edge, H - 1 - crop_edge, H - 2 * crop_edge, device=device),
                          indexing="ij")
    i, j = i.t(), j.t()
    dirs = torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1)
    dirs = dirs.reshape(H - 2 * crop_edge, W - 2 * crop_edge, 1, 3)
    rays_d = torch.sum(dirs * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d
