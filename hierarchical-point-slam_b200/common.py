"""Drop-ins for the hot-path helpers of the reference's src/common.py.

raw2outputs_nerf_color (common.py:385-428) runs in the compositing kernel; the ray / pose helpers
(common.py:103-119, 304-346, 431-448) stay small torch expressions so that autograd chains the kernels'
d/d rays_o, d/d rays_d into the quaternion - they are written device-agnostically
(the reference's `.to(quad.get_device())` is CUDA-only).
"""
import numpy as np
import torch

from . import frame, ops


def setup_seed(seed):
    """common.py:31-37: the reference seeds torch, numpy and `random` (1219 everywhere) before building the decoders
    and the cloud; NICER's seeded init equals the reference's only after the same call."""
    import random
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    random.seed(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False


def depth_inside_mask(batch_gt_depth):
    """The batch filter both loops apply right before render_batch_ray (Tracker.py:179-181, Mapper.py:766-768, `nice`
    branch without a bound): keep rays whose depth is <= min(10 median(d), 1.2 max(d)).  A batch-global statistic: under
    ray sharding compute it BEFORE the split.  Device ops only; torch.median = the lower median, as there."""
    with torch.no_grad():
        return batch_gt_depth <= torch.minimum(10 * batch_gt_depth.median(), 1.2 * torch.max(batch_gt_depth))


def clone_kf_dict(keyframes_dict):
    """common.py:523-531: a list of keyframe dicts with every tensor cloned (what npc.get_keyframe_dict hands out)."""
    return [{key: value.clone() for key, value in keyframe.items()} for keyframe in keyframes_dict]


def as_intrinsics_matrix(intrinsics):
    """common.py:40-50: (fx, fy, cx, cy) -> 3x3 numpy matrix."""
    K = np.eye(3)
    K[0, 0], K[1, 1], K[0, 2], K[1, 2] = intrinsics[0], intrinsics[1], intrinsics[2], intrinsics[3]
    return K


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


# ---- pixel / ray sampling that feeds the path.  Everything stays on the frame's device; the reference
# ---- builds a dense (u,v) meshgrid per call and, for the gradient-guided variants, goes through skimage on the host.

def random_select(l, k):
    """k distinct values from range(l), numpy RNG like the reference (common.py:95-100)."""
    return list(np.random.permutation(l)[:min(l, k)])


def select_uv(i, j, n, depth, color, device="cuda:0"):
    """n pixels drawn with replacement from dense (u,v) maps; common.py:122-137 (same torch.randint call, so the
    same pixels for the same seed)."""
    i, j = i.reshape(-1), j.reshape(-1)
    pick = torch.randint(i.shape[0], (n,), device=device)
    return i[pick], j[pick], depth.reshape(-1)[pick], color.reshape(-1, 3)[pick]


def get_sample_uv(H0, H1, W0, W1, n, depth, color, device="cuda:0"):
    """n (u,v) from the window rows H0..H1-1, columns W0..W1-1; common.py:159-171.  The reference materialises the
    window's meshgrid (two (H1-H0, W1-W0) float maps per call) and gathers from it; the maps hold the integers
    u = W0 + column, v = H0 + row exactly, so the draw is decoded arithmetically instead."""
    rows, cols = H1 - H0, W1 - W0
    pick = torch.randint(rows * cols, (n,), device=device)
    r = torch.div(pick, cols, rounding_mode="floor")
    c = pick - r * cols
    return (c + W0).float(), (r + H0).float(), depth[r + H0, c + W0], color[r + H0, c + W0]


def _depth_filter(sample_depth, depth_limit):
    mask = sample_depth > 0
    if depth_limit is not None:
        mask = mask & (sample_depth < depth_limit)
    return mask


def get_samples(H0, H1, W0, W1, n, H, W, fx, fy, cx, cy, c2w, depth, color, device,
                depth_filter=False, return_index=False, depth_limit=None):
    """n rays from the window with their depth / colour (and pixel indices); common.py:236-258."""
    i, j, sample_depth, sample_color = get_sample_uv(H0, H1, W0, W1, n, depth, color, device=device)
    if depth_filter:
        keep = _depth_filter(sample_depth, depth_limit)        # filtering before the ray kernel: same rays, less work
        i, j, sample_depth, sample_color = i[keep], j[keep], sample_depth[keep], sample_color[keep]
    rays_o, rays_d = get_rays_from_uv(i, j, c2w, H, W, fx, fy, cx, cy, device)
    if return_index:
        return rays_o, rays_d, sample_depth, sample_color, i.to(torch.int64), j.to(torch.int64)
    return rays_o, rays_d, sample_depth, sample_color


def _top_gradient_pixels(image, count, H0, H1, W0, W1):
    """flat indices (int64, on image.device) of the `count` pixels with the largest Sobel colour-gradient magnitude
    that also lie in the window, and the (H,W) float64 magnitude map; the host part of common.py:174-233."""
    mag = frame.color_gradient_magnitude(image, float("inf"))
    Himg, Wimg = mag.shape
    top = torch.topk(mag.reshape(-1), min(count, mag.numel()), sorted=False).indices     # np.argpartition's set
    h = torch.div(top, Wimg, rounding_mode="floor")
    w = top - h * Wimg
    inside = (h >= H0) & (h < H1) & (w >= W0) & (w < W1)
    return top, h, w, inside, mag


def get_sample_uv_with_grad(H0, H1, W0, W1, n, image):
    """n distinct flat pixel indices among the top 5n colour-gradient pixels inside the window; common.py:174-196
    (np.random.choice without replacement -> a device permutation)."""
    top, _, _, inside, _ = _top_gradient_pixels(image, 5 * n, H0, H1, W0, W1)
    cand = top[inside]
    if cand.numel() < n:
        raise ValueError("Cannot take a larger sample than population when 'replace=False'")
    return cand[torch.randperm(cand.numel(), device=cand.device)[:n]]


def get_selected_index_with_grad(H0, H1, W0, W1, n, image, ratio=15, gt_depth=None, depth_limit=False,
                                 as_numpy=True):
    """Flat indices of the top ratio*n colour-gradient pixels inside the window (with a depth reading, <= 5 m when
    depth_limit) and the gradient-magnitude map; common.py:199-233.  as_numpy=True returns ndarrays like the
    reference (one device->host read per frame); False keeps both on the device."""
    top, h, w, keep, mag = _top_gradient_pixels(image, ratio * n, H0, H1, W0, W1)
    if gt_depth is not None:
        d = gt_depth[h, w]
        keep = keep & (d > 0.0)
        if depth_limit:
            keep = keep & (d <= 5.0)
    sel = top[keep]
    if as_numpy:
        return sel.cpu().numpy(), mag.cpu().numpy()
    return sel, mag


def get_samples_with_pixel_grad(H0, H1, W0, W1, num_tuple, H, W, fx, fy, cx, cy, c2w, depth, color, device, est_normal,
                                depth_filter=True, return_index=True, depth_limit=None):
    """Rays through the union of colour-gradient and normal-gradient guided pixels; common.py:261-301."""
    n_color, n_normal = num_tuple
    assert n_color > 0 or n_normal > 0, "invalid number of rays to sample."
    picks = []
    if n_color > 0:
        picks.append(get_sample_uv_with_grad(H0, H1, W0, W1, n_color, color))
    if n_normal > 0:
        picks.append(get_sample_uv_with_grad(H0 + 10, H1 - 10, W0 + 10, W1 - 10, n_normal, est_normal))
    merged = torch.unique(torch.cat(picks))                      # np.union1d: sorted, duplicates dropped
    j = torch.div(merged, W, rounding_mode="floor")
    i = merged - j * W
    sample_depth, sample_color = depth[j, i], color[j, i]
    if depth_filter:
        keep = _depth_filter(sample_depth, depth_limit)
        i, j, sample_depth, sample_color = i[keep], j[keep], sample_depth[keep], sample_color[keep]
    rays_o, rays_d = get_rays_from_uv(i.float(), j.float(), c2w, H, W, fx, fy, cx, cy, device)
    if return_index:
        return rays_o, rays_d, sample_depth, sample_color, i, j
    return rays_o, rays_d, sample_depth, sample_color


def get_rotation_from_tensor(quad):
    """(4,) or (B,4) quaternion -> (3,3) or (B,3,3); common.py:349-354."""
    if quad.dim() == 1:
        quad = quad.unsqueeze(0)
    return quad2rotation(quad).squeeze(0)


def get_tensor_from_camera(RT, Tquad=False):
    """(3|4,4) camera matrix -> float32 [quat(w,x,y,z), T] (or [T, quat]); common.py:357-382.  The reference goes
    through scipy on the host (Rotation.from_matrix(R).as_quat(), float64: nearest rotation by SVD, then the
    largest-of-(diagonal, trace) branch, normalised); the same arithmetic runs here on RT's device and the result
    stays there (the reference always hands back a CPU tensor, which its callers move to the device again)."""
    if isinstance(RT, np.ndarray):
        RT = torch.from_numpy(RT)
    RT = RT.detach()
    m = RT[:3, :3].to(torch.float64)
    if torch.linalg.det(m) <= 0:
        raise ValueError("Non-positive determinant (left-handed or null coordinate frame) in rotation matrix")
    if not torch.allclose(m @ m.T, torch.eye(3, dtype=m.dtype, device=m.device), rtol=1e-5, atol=1e-12):
        U, _, Vt = torch.linalg.svd(m)
        m = U @ Vt
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    choice = int(torch.argmax(torch.stack([m[0, 0], m[1, 1], m[2, 2], tr])))
    q = torch.empty(4, dtype=m.dtype, device=m.device)             # (x, y, z, w)
    if choice < 3:
        a, b, c = choice, (choice + 1) % 3, (choice + 2) % 3
        q[a] = 1 - tr + 2 * m[a, a]
        q[b] = m[b, a] + m[a, b]
        q[c] = m[c, a] + m[a, c]
        q[3] = m[c, b] - m[b, c]
    else:
        q[0] = m[2, 1] - m[1, 2]
        q[1] = m[0, 2] - m[2, 0]
        q[2] = m[1, 0] - m[0, 1]
        q[3] = 1 + tr
    q = q / torch.linalg.norm(q)
    quad = torch.roll(q, 1)                                          # (w, x, y, z)
    T = RT[:3, 3].to(torch.float64)
    return (torch.cat([T, quad]) if Tquad else torch.cat([quad, T])).float()
