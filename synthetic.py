"""Synthetic ScanNet-shaped inputs: intrinsics, a smooth room-like depth image, a point cloud unprojected
from it and rays through random pixels.  Pure data generation - no algorithm of the path lives here - shared by
bench.py (both arms), the GPU-box scripts, the tests and the oracle."""
import torch

# ----------------------------------------------------------------------------------------------
# synthetic ScanNet-shaped inputs shared by tests / bench
# ----------------------------------------------------------------------------------------------
SCANNET_INTRINSICS = dict(H=480, W=640, fx=577.590698, fy=578.729797, cx=318.905426, cy=242.683609)


def synthetic_depth(H=480, W=640):
    v, u = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32),
                          indexing="ij")
    return 2.0 + 0.3 * torch.sin(u / 50.0) + 0.2 * torch.cos(v / 40.0)


def synthetic_scene(n_points, n_rays, level_radius=(0.04, 0.16), seed=1219, extent_scale=1.0, c_dim=32):
    """Cloud = unprojected depth-surface samples jittered by the N_add pattern, rays through random
    pixels, per-ray f64 radii in [rmin, rmax].  Identity pose (camera looks down -z)."""
    g = torch.Generator().manual_seed(seed)
    K = SCANNET_INTRINSICS
    depth = synthetic_depth()

    def unproject(u, v, d):
        return torch.stack([(u - K["cx"]) / K["fx"] * d, -(v - K["cy"]) / K["fy"] * d, -d], -1)

    # cloud
    u = torch.rand(n_points, generator=g) * (K["W"] - 1) * extent_scale - (extent_scale - 1) * K["W"] / 2
    v = torch.rand(n_points, generator=g) * (K["H"] - 1) * extent_scale - (extent_scale - 1) * K["H"] / 2
    ui = u.clamp(0, K["W"] - 1).long()
    vi = v.clamp(0, K["H"] - 1).long()
    d = depth[vi, ui] * (0.96 + 0.04 * torch.randint(0, 3, (n_points,), generator=g).float())
    pts = unproject(u, v, d).float().contiguous()
    geo = torch.randn(n_points, c_dim, generator=g) * 0.1
    col = torch.randn(n_points, c_dim, generator=g) * 0.1
    # rays
    ru = torch.randint(10, K["W"] - 10, (n_rays,), generator=g).float()
    rv = torch.randint(10, K["H"] - 10, (n_rays,), generator=g).float()
    rays_d = torch.stack([(ru - K["cx"]) / K["fx"], -(rv - K["cy"]) / K["fy"], -torch.ones(n_rays)], -1)
    rays_o = torch.zeros(n_rays, 3)
    gt_depth = depth[rv.long(), ru.long()].clone()
    gt_color = torch.rand(n_rays, 3, generator=g)
    rmin, rmax = level_radius
    radius = (rmin + (rmax - rmin) * torch.rand(n_rays, generator=g, dtype=torch.float64))
    return dict(points=pts, geo_feats=geo, col_feats=col, rays_o=rays_o, rays_d=rays_d.float(),
                gt_depth=gt_depth, gt_color=gt_color, radius=radius)
