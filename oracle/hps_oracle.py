"""CPU oracle for the Hierarchical Point-SLAM render-and-optimise hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this module; it is used by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the checker
(and as the reported CPU baseline), never as the thing shipped.

It is a restatement, in plain torch-CPU fp32 ops, of the reference algorithm.  Every function cites the
reference file:line it follows (paths relative to the reference root).  Parity pinning: the reference
has no tests or golden vectors of its own, so this restatement is pinned against
outputs of the reference's *own imported modules* run in the build container by oracle/make_golden.py
(neighbour index forced to exhaustive search, see oracle/ref_import.py); the resulting fixtures are
committed under tests/golden/ and tests/test_oracle_golden.py replays them.

Decoder parameters are passed as flat dicts keyed by the reference's state_dict names
("pts_linears.0.weight", "fc_c.3.bias", "output_linear.weight", "embedder._B", ...).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

K_NN = 8


# ----------------------------------------------------------------------------------------------
# neighbour search  (neural_point.py:366-417 with the faiss index forced to exhaustive search)
# ----------------------------------------------------------------------------------------------
def sqdist_direct(q, p):
    """fl(fl(fl(dx*dx)+fl(dy*dy))+fl(dz*dz)), dx = fl(q-p); fp32, no FMA."""
    dx = q[:, None, 0] - p[None, :, 0]
    dy = q[:, None, 1] - p[None, :, 1]
    dz = q[:, None, 2] - p[None, :, 2]
    return (dx * dx + dy * dy) + dz * dz


def exhaustive_knn(points, queries, k=K_NN, chunk_elems=1 << 26):
    """k smallest squared distances per query in ascending (D, index) order.

    Restates index.search(x, k) (neural_point.py:393) for an IVF index with nprobe == nlist.
    Requires points.shape[0] >= k (true for every reference cloud)."""
    points = points.detach().float().reshape(-1, 3)
    queries = queries.detach().float().reshape(-1, 3)
    n = points.shape[0]
    assert n >= k
    Ds, Is = [], []
    step = max(1, chunk_elems // n)
    for s in range(0, queries.shape[0], step):
        D = sqdist_direct(queries[s:s + step], points)
        Dk, Ik = torch.sort(D, dim=1, stable=True)
        Ds.append(Dk[:, :k].contiguous())
        Is.append(Ik[:, :k].contiguous())
    if not Ds:
        return torch.zeros(0, k), torch.zeros(0, k, dtype=torch.long)
    return torch.cat(Ds), torch.cat(Is)


def neighbor_count(D, radius):
    """(D < r**2).sum(-1).int(); r is float64 when dynamic (neural_point.py:410-415)."""
    if torch.is_tensor(radius):
        return (D < radius.reshape(-1, 1) ** 2).sum(-1).int()
    return (D < radius ** 2).sum(-1).int()


def find_neighbors(points, queries, radius, k=K_NN):
    """find_neighbors_faiss (neural_point.py:366-417): returns D (Q,k) f32, I (Q,k) i64, n (Q,) i32."""
    D, I = exhaustive_knn(points, queries, k)
    return D, I, neighbor_count(D, radius)


def radius_thresholds(radius, n=None):
    """fp32 thresholds equivalent to the reference's float64 comparisons.

    D < r64**2  <=>  D < t_up   (t_up = smallest f32 >= r64**2)   [count, neural_point.py:412]
    D > r64**2  <=>  D > t_dn   (t_dn = largest  f32 <= r64**2)   [weight zeroing, decoder.py:236]
    """
    if not torch.is_tensor(radius):
        radius = torch.full((n,), float(radius), dtype=torch.float64)
    r2 = radius.double().reshape(-1) ** 2
    t = r2.float()
    inf = torch.tensor(float("inf"))
    t_dn = torch.where(t.double() > r2, torch.nextafter(t, -inf), t)
    t_up = torch.where(t.double() < r2, torch.nextafter(t, inf), t)
    return t_up, t_dn


class KDTreeNeighbors:
    """Fast exact CPU neighbour search for the *timed* CPU baseline (scipy cKDTree, all cores).  Stands in
    for faiss-cpu, which is not installable here.  Same return contract as find_neighbors;
    distances are re-evaluated with the fp32 direct formula so downstream weights match the oracle."""

    def __init__(self):
        self._trees = {}

    def __call__(self, points, queries, radius, k=K_NN):
        from scipy.spatial import cKDTree
        key = (points.data_ptr(), points.shape[0])
        if key not in self._trees:
            self._trees[key] = cKDTree(points.detach().numpy())
        q = queries.detach().float()
        _, I = self._trees[key].query(q.numpy(), k=k, workers=-1)
        I = torch.from_numpy(np.ascontiguousarray(I)).long()
        d = q[:, None, :] - points[I]
        D = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        D, order = torch.sort(D, dim=1, stable=True)
        I = torch.gather(I, 1, order)
        return D, I, neighbor_count(D, radius)


# ----------------------------------------------------------------------------------------------
# inverse-distance interpolation  (decoder.py:195-272 == :491-559)
# ----------------------------------------------------------------------------------------------
def idw_features(p, feats, D, I, nn_count, radius, c_fill, min_nn=2, is_tracker=False, cloud_pos=None):
    """get_feature_at_pos with nn_weighting='distance', encode_rel_pos off, normals off.

    p (P,3); feats (N,C); D,I,nn_count from find_neighbors; radius: f64 (P,) / (P,1) tensor or float.
    c_fill (C,): the vector the reference draws with normal_(0, 0.01) for no-neighbour points
    (decoder.py:262-263) - an explicit input here."""
    p = p.reshape(-1, 3)
    if torch.is_tensor(radius):
        bound = radius.clone().reshape(-1, 1) ** 2                         # decoder.py:207-208 (f64)
    else:
        bound = radius ** 2
    if is_tracker:                                                         # decoder.py:209-218
        D = torch.sum(torch.square(cloud_pos[I] - p.reshape(-1, 1, 3)), dim=-1)
        D = D.reshape(-1, I.shape[1])
        D = torch.where(D > bound, torch.full_like(D, 1e4), D)
    has_neighbors = nn_count > (min_nn - 1)                                # decoder.py:222
    w = 1.0 / (D + 1e-10)                                                  # decoder.py:229
    w = torch.where((D > bound).detach(), torch.zeros_like(w), w)         # decoder.py:235-236
    w = F.normalize(w, p=1, dim=1).unsqueeze(-1)                           # decoder.py:239
    c = (w * feats[I]).sum(dim=1).reshape(-1, feats.shape[1])              # decoder.py:245,260-261
    c = torch.where(has_neighbors[:, None], c, c_fill[None, :].to(c.dtype))  # decoder.py:262-263
    return c, has_neighbors


# ----------------------------------------------------------------------------------------------
# decoders  (decoder.py:36-45, 274-342, 561-618)
# ----------------------------------------------------------------------------------------------
def fourier_embed(p, B, concat):
    """GaussianFourierFeatureTransform.forward (decoder.py:36-45)."""
    x = (2 * math.pi * p) @ B
    if concat:
        return torch.cat((torch.sin(x), torch.cos(x)), dim=-1)
    return torch.sin(x)


def _blocks(e, c, prm, act, n_blocks=5, skips=(2,)):
    h = e
    for i in range(n_blocks):
        h = F.linear(h, prm["pts_linears.%d.weight" % i], prm["pts_linears.%d.bias" % i])
        h = act(h)
        h = h + F.linear(c, prm["fc_c.%d.weight" % i], prm["fc_c.%d.bias" % i])
        if i in skips:
            h = torch.cat([e, h], -1)
    return F.linear(h, prm["output_linear.weight"], prm["output_linear.bias"])


def geo_preactivations(p, c, prm, n_blocks=5, skips=(2,)):
    """Pre-activations a_i = W_i x_i + b_i of the ReLU decoder, (P, n_blocks, hidden).  Used by the tests to
    find sample points that sit on a ReLU kink within fp32 noise (|a| ~ 1e-4 from the rounding of the
    Fourier arguments 2*pi*p*B ~ 1e3 rad), where the reference's own gradient is discontinuous."""
    e = fourier_embed(p.reshape(-1, 3), prm["embedder._B"], concat=False)
    h, out = e, []
    for i in range(n_blocks):
        a = F.linear(h, prm["pts_linears.%d.weight" % i], prm["pts_linears.%d.bias" % i])
        out.append(a)
        h = F.relu(a) + F.linear(c, prm["fc_c.%d.weight" % i], prm["fc_c.%d.bias" % i])
        if i in skips:
            h = torch.cat([e, h], -1)
    return torch.stack(out, 1)


def geo_mlp(p, c, prm):
    """MLP_geometry.forward body (decoder.py:316-341): sin embedding (B learnable, 3x93), ReLU."""
    e = fourier_embed(p.reshape(-1, 3), prm["embedder._B"], concat=False)
    return _blocks(e, c, prm, F.relu).squeeze(-1)


def exposure_affine(prm, exposure_feat):
    """MLP_exposure.forward (decoder.py:392-396) -> (12,) affine tensor."""
    h = F.linear(exposure_feat, prm["mlp_exposure.linear1.weight"], prm["mlp_exposure.linear1.bias"])
    h = F.softplus(h, beta=100)
    return F.linear(h, prm["mlp_exposure.linear2.weight"], prm["mlp_exposure.linear2.bias"])


def col_mlp(p, c, prm, exposure_feat=None, encode_exposure=True):
    """MLP_color.forward body (decoder.py:581-618): sin|cos embedding (B fixed 3x20), Softplus(100)."""
    e = fourier_embed(p.reshape(-1, 3), prm["embedder._B"], concat=True)
    out = _blocks(e, c, prm, lambda x: F.softplus(x, beta=100))
    if encode_exposure:
        if exposure_feat is None:
            return out                                                     # decoder.py:613-614
        a = exposure_affine(prm, exposure_feat)
        return torch.sigmoid(out @ a[:9].reshape(3, 3) + a[-3:])           # decoder.py:608-612
    return torch.sigmoid(out)


def nicer_forward(p, points, stage, geo_feats, col_feats, radius, dec, c_fill_geo, c_fill_col,
                  n_surface=5, min_nn=2, is_tracker=False, exposure_feat=None, neighbors=None, knn_fn=None):
    """NICER.forward for the four live stages (decoder.py:660-723).

    points / geo_feats / col_feats / radius: dicts level -> tensor (radius: per-point f64).
    dec: dict decoder name -> parameter dict.  Returns raw (P,4), ray_mask (R,), point_mask (P,)."""
    level = "mid" if stage.endswith("mid") else "fine"
    p = p.reshape(-1, 3)
    if neighbors is None:
        neighbors = (knn_fn or find_neighbors)(points[level], p, radius[level])
    D, I, n = neighbors
    cg, has = idw_features(p, geo_feats[level], D, I, n, radius[level], c_fill_geo, min_nn,
                           is_tracker, points[level])
    ray_mask = ~(has.view(-1, n_surface).sum(1) < int(n_surface / 2 + 1))  # decoder.py:310-311
    occ = geo_mlp(p, cg, dec["geo_decoder_" + level])
    if stage.startswith("geometry"):                                       # decoder.py:682-699
        raw = torch.cat([torch.zeros(occ.shape[0], 3, dtype=occ.dtype), occ[:, None]], -1)
        return raw, ray_mask, has
    cc, _ = idw_features(p, col_feats[level], D, I, n, radius[level], c_fill_col, min_nn,
                         is_tracker, points[level])
    rgb = col_mlp(p, cc, dec["color_decoder_" + level], exposure_feat)
    return torch.cat([rgb, occ[:, None]], -1), ray_mask, has               # decoder.py:710,722


# ----------------------------------------------------------------------------------------------
# compositing  (common.py:385-428 with occupancy=True)
# ----------------------------------------------------------------------------------------------
def composite(raw, z_vals, coef=0.1):
    """raw (R,S,4) with the occupancy logit in channel 3.  Returns depth, var, rgb, weights."""
    alpha = torch.sigmoid(coef * raw[..., 3])
    ones = torch.ones(alpha.shape[0], 1, dtype=alpha.dtype)
    trans = torch.cumprod(torch.cat([ones, 1.0 - alpha + 1e-10], -1), -1)[:, :-1]
    w = alpha * trans
    wsum = w.sum(-1, keepdim=True) + 1e-10
    rgb = (w[..., None] * raw[..., :3]).sum(-2) / wsum
    depth = (w * z_vals).sum(-1) / wsum.squeeze(-1)
    d = z_vals - depth.unsqueeze(-1)
    var = (w * d * d).sum(1)
    return depth, var, rgb, w


# ----------------------------------------------------------------------------------------------
# render_batch_ray  (Renderer.py:84-254; N_samples=0, use_bound False, sample_near_pcl False)
# ----------------------------------------------------------------------------------------------
def surface_z_vals(gt_depth, near_end_surface=0.96, far_end_surface=1.04, n_surface=5,
                   near_end=0.3):
    """Depth-guided samples (Renderer.py:163-208): z = near*d*(1-t) + far*d*t, t = linspace(0,1,S);
    zero-depth rays get linspace(near_end, max far) with far = clamp(far_bb, 0, max(1.2 d))."""
    gt = gt_depth.reshape(-1, 1)
    t = torch.linspace(0.0, 1.0, steps=n_surface, dtype=gt.dtype)
    nz = gt.squeeze(-1) > 0
    d = gt[nz].repeat(1, n_surface)
    z = torch.zeros(gt.shape[0], n_surface, dtype=gt.dtype)
    z[nz] = near_end_surface * d * (1.0 - t) + far_end_surface * d * t
    if nz.sum() < gt.shape[0]:
        far_bb = torch.minimum(5 * gt.mean(), torch.max(gt * 1.2))          # Renderer.py:133-134
        if torch.max(gt) > 0:
            far = torch.clamp(far_bb, 0, torch.max(gt * 1.2).item())        # Renderer.py:150
        else:
            far = far_bb
        z[~nz] = torch.linspace(near_end, float(far), steps=n_surface).repeat(int((~nz).sum()), 1)
    return z, nz


def render_batch_ray(points, dec, rays_d, rays_o, stage, gt_depth, geo_feats, col_feats, radius_ray,
                     c_fill_geo, c_fill_col, is_tracker=False, exposure_feat=None, coef=0.1,
                     near_end_surface=0.96, far_end_surface=1.04, n_surface=5, min_nn=2, knn_fn=None):
    """Returns depth (R,), uncertainty (R,), color (R,3), valid_ray_mask (R,).

    radius_ray: dict level -> per-RAY float64 radii (repeated per sample, Renderer.py:226-229)."""
    z, nz = surface_z_vals(gt_depth, near_end_surface, far_end_surface, n_surface)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]     # Renderer.py:218-219
    radius_pt = {k: v.reshape(-1, 1).repeat_interleave(n_surface, dim=0) for k, v in radius_ray.items()}
    raw, ray_mask, point_mask = nicer_forward(pts.reshape(-1, 3), points, stage, geo_feats, col_feats,
                                              radius_pt, dec, c_fill_geo, c_fill_col, n_surface, min_nn,
                                              is_tracker, exposure_feat, knn_fn=knn_fn)
    # Renderer.py:237-239: overwrite under no_grad, straight-through for the backward
    occ = raw[:, 3]
    occ = occ + (torch.where(point_mask, occ, torch.full_like(occ, -100.0)) - occ).detach()
    raw = torch.cat([raw[:, :3], occ[:, None]], -1).reshape(-1, n_surface, 4)
    depth, var, rgb, _ = composite(raw, z, coef)
    depth = torch.where(nz, depth, torch.zeros_like(depth))                 # Renderer.py:250-251
    return depth, var, rgb, ray_mask


# ----------------------------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------------------------
def mapper_loss(depth, color, valid, gt_depth, gt_color, stage, w_color=0.1, affine=None):
    """Mapper.py:791-856.  affine: (12,) exposure tensor applied to the whole batch (single frame) or
    None; colour stages apply  sigmoid(color @ A[:9].view(3,3) + A[-3:])  before the L1."""
    mask = (gt_depth > 0) & valid & (~torch.isnan(depth))
    loss = torch.abs(gt_depth[mask] - depth[mask]).sum()
    if stage.startswith("color"):
        if affine is not None:
            color = color @ affine[:9].reshape(3, 3) + affine[-3:]
        color = torch.sigmoid(color)
        loss = loss + w_color * torch.abs(gt_color[mask] - color[mask]).sum()
    return loss


def tracker_loss(depth, uncertainty, color, gt_depth, gt_color, w_color=0.5):
    """Tracker.py:213-231 (handle_dynamic True, use_color_in_tracking True)."""
    unc = uncertainty.detach()
    nan_mask = (~torch.isnan(depth)) & (~torch.isnan(unc))
    tmp = torch.abs(gt_depth - depth) / torch.sqrt(unc + 1e-10)
    mask = (tmp < 10 * tmp.mean()) & (gt_depth > 0) & nan_mask
    geo = torch.clamp(tmp, min=0.0, max=1e3)[mask].sum()
    col = torch.abs(gt_color - color)[mask].sum()
    return geo + w_color * col


# ----------------------------------------------------------------------------------------------
# pose / ray helpers  (common.py:103-119, 304-346)
# ----------------------------------------------------------------------------------------------
def quat_to_rot(q):
    """quad2rotation for one quaternion (w,x,y,z) (common.py:304-328)."""
    qr, qi, qj, qk = q[0], q[1], q[2], q[3]
    s = 2.0 / (q * q).sum()
    return torch.stack([
        torch.stack([1 - s * (qj ** 2 + qk ** 2), s * (qi * qj - qk * qr), s * (qi * qk + qj * qr)]),
        torch.stack([s * (qi * qj + qk * qr), 1 - s * (qi ** 2 + qk ** 2), s * (qj * qk - qi * qr)]),
        torch.stack([s * (qi * qk - qj * qr), s * (qj * qk + qi * qr), 1 - s * (qi ** 2 + qj ** 2)])])


def camera_from_tensor(cam):
    """get_camera_from_tensor (common.py:330-346): (7,) quat+T -> (3,4)."""
    return torch.cat([quat_to_rot(cam[:4]), cam[4:, None]], 1)


def rays_from_uv(i, j, c2w, fx, fy, cx, cy):
    """get_rays_from_uv (common.py:103-119)."""
    dirs = torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1).reshape(-1, 1, 3)
    rays_d = torch.sum(dirs * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


# ----------------------------------------------------------------------------------------------
# dynamic radius maps.  PARITY UNPINNED: skimage (rgb2gray, filters.sobel_h / sobel_v) is not
# importable here; this restates scikit-image 0.19.3 (the reference's env.yaml pin) with scipy.ndimage.
# ----------------------------------------------------------------------------------------------
def sobel_magnitude(gt_color):
    """skimage rgb2gray -> filters.sobel_h / sobel_v -> magnitude (common.py:174-233, Tracker.py:297-325) restated with
    scipy.ndimage; float64 (H,W).  PARITY UNPINNED (skimage is not installed here)."""
    import numpy as np
    from scipy import ndimage as ndi
    img = np.asarray(gt_color, dtype=np.float64)          # frames are float64 here (datasets.py: uint8 / 255.)
    gray = img @ np.array([0.2125, 0.7154, 0.0721])

    def sobel(axis):          # skimage.filters.sobel(image, axis=axis): edge [1,0,-1] along axis, smooth [1,2,1]/4 across
        out = ndi.correlate1d(gray, [1, 0, -1], axis=axis, mode="reflect")
        return ndi.correlate1d(out, np.array([1, 2, 1]) / 4.0, axis=1 - axis, mode="reflect")
    gy, gx = sobel(0), sobel(1)
    return np.sqrt(gx ** 2 + gy ** 2)


def dynamic_radius_maps(gt_color, radius_hierarchy, ratio, threshold):
    """Tracker.py:297-325 / Mapper.py:1026-1050 in numpy + scipy: level -> (r_add, r_query) float64 (H,W) arrays."""
    import numpy as np
    from scipy.interpolate import interp1d
    mag = np.clip(sobel_magnitude(gt_color), 0.0, threshold)
    out = {}
    for level, d in radius_hierarchy.items():
        vals = list(d.values())
        f_add = interp1d([0, 0.01, threshold], [vals[0], vals[0], vals[1]])
        f_q = interp1d([0, 0.01, threshold], [ratio * vals[0], ratio * vals[0], ratio * vals[1]])
        out[level] = (f_add(mag), f_q(mag))
    return out


# ----------------------------------------------------------------------------------------------
# frustum feature selection.  PINNED: oracle/make_golden_frustum.py executes the reference's
# Mapper.get_mask_from_c2w with the real cv2 (4.13, installed in the build image) and commits the selected indices;
# tests/test_oracle_golden.py requires this restatement to return the same sets and its sampler to equal cv2.remap.
# ----------------------------------------------------------------------------------------------
def remap_bilinear_cv2(depth, x, y):
    """cv2.remap(depth, x, y, interpolation=cv2.INTER_LINEAR) for a float32 image and float32 maps, default constant-0
    border, as called at Mapper.py:231-236.  OpenCV (imgproc/remap: INTER_BITS = 5) rounds the sampling position to
    1/32 pixel - sx = cvRound(x * 32), pixel sx >> 5, fraction (sx & 31) / 32 - takes the four weights from its float
    table ((1 - fy)(1 - fx), (1 - fy) fx, fy (1 - fx), fy fx; exact products of multiples of 1/32) and sums
    S00 w0 + S01 w1 + S10 w2 + S11 w3 left to right in float32; taps outside the image contribute 0; the integer pixel
    coordinate saturates to int16.  Pinned bit-exact against cv2 4.13 by tests/golden/frustum_small.pt."""
    import numpy as np
    d = np.asarray(depth, dtype=np.float32)
    H, W = d.shape
    x = np.asarray(x, dtype=np.float32)
    y = np.asarray(y, dtype=np.float32)

    def fixed(t):                                          # cvRound = round half to even; NaN / overflow -> INT_MIN
        t32 = t * np.float32(32.0)
        bad = ~np.isfinite(t32) | (np.abs(t32) >= 2.0 ** 31)
        r = np.rint(np.where(bad, 0.0, t32).astype(np.float64)).astype(np.int64)
        r = np.where(bad, -2 ** 31, r)
        return np.clip(r >> 5, -32768, 32767), (r & 31)
    ix, fx5 = fixed(x)
    iy, fy5 = fixed(y)
    ax = fx5.astype(np.float32) * np.float32(1.0 / 32.0)
    ay = fy5.astype(np.float32) * np.float32(1.0 / 32.0)
    w = [(np.float32(1) - ay) * (np.float32(1) - ax), (np.float32(1) - ay) * ax, ay * (np.float32(1) - ax), ay * ax]

    def at(yy, xx):
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        out = np.zeros(xx.shape, dtype=np.float32)
        out[ok] = d[yy[ok], xx[ok]]
        return out
    acc = at(iy, ix) * w[0]
    acc = acc + at(iy, ix + 1) * w[1]
    acc = acc + at(iy + 1, ix) * w[2]
    acc = acc + at(iy + 1, ix + 1) * w[3]
    return acc.astype(np.float32)


def frustum_mask(points, c2w, depth, fx, fy, cx, cy, edge):
    """Mapper.get_mask_from_c2w (Mapper.py:194-250): indices of the cloud points that project inside the frame
    (cropped by `edge` pixels) and lie in front of the camera, not further than depth + 0.5 m.  numpy; the types follow
    the reference: w2c is the float32 inverse of the float32 pose (:208-209), the projection runs in float64 on the
    float64 point list (:213-227), uv is cast to float32 (:228), the depth is sampled by cv2.remap's fixed-point
    bilinear (:231-236), `depths + 0.5` is a float32 sum (:246).  Pinned against the reference's own method executed
    with the real cv2 (oracle/make_golden_frustum.py).  Also returns each point's distance to the nearest threshold."""
    import numpy as np
    H, W = depth.shape
    pts = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    w2c = np.linalg.inv(np.asarray(c2w, dtype=np.float32)).astype(np.float64)
    cam = (w2c @ np.concatenate([pts, np.ones_like(pts[:, :1])], 1).T).T[:, :3]
    cam[:, 0] *= -1                                         # Mapper.py:225 axis flip
    Kmat = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    uvz = (Kmat @ cam.T).T
    z = uvz[:, 2:] + 1e-5
    with np.errstate(divide="ignore", invalid="ignore"):
        uv = (uvz[:, :2] / z).astype(np.float32)
    dep = remap_bilinear_cv2(depth, uv[:, 0], uv[:, 1])
    mask = (uv[:, 0] < W - edge) & (uv[:, 0] > edge) & (uv[:, 1] < H - edge) & (uv[:, 1] > edge)
    dep[dep == 0] = dep.max()
    mz = -z[:, 0]
    far = (dep + np.float32(0.5)).astype(np.float64)        # float32 sum, compared against the float64 -z
    mask = mask & (0 <= mz) & (mz <= far)
    margin = np.minimum.reduce([np.abs(uv[:, 0] - (W - edge)), np.abs(uv[:, 0] - edge), np.abs(uv[:, 1] - (H - edge)),
                                np.abs(uv[:, 1] - edge), np.abs(mz), np.abs(mz - far)])
    return np.where(mask)[0], margin


# ----------------------------------------------------------------------------------------------
# synthetic inputs: generated by the neutral module synthetic.py (repo root); re-exported for the tests
# ----------------------------------------------------------------------------------------------
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _ROOT not in _sys.path:
    _sys.path.insert(0, _ROOT)
from synthetic import SCANNET_INTRINSICS, synthetic_depth, synthetic_scene  # noqa: E402,F401
