"""-m gpu: every kernel, called through the C ABI, against the CPU oracle on identical seeded inputs."""
import pytest
import torch

from oracle import hps_oracle as O
from tests import parity

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = parity.TOL   # 1e-4 relative, fp32 (BASELINE.json north_star)


def _check_knn(hps, pts, q, radius, cell=None):
    """Product query vs exhaustive oracle: counts equal; slots < #(D <= t_dn) bit-equal; the rest empty."""
    idx = hps.GridIndex(DEV)
    idx.reset(pts.to(DEV))
    idx.cell_hint = cell
    res = idx.radius_search(q.to(DEV), radius.to(DEV) if torch.is_tensor(radius) else radius)
    I, D, cnt = res.I.cpu(), res.D.cpu(), res.count.cpu()
    Q = q.shape[0]
    k = min(8, pts.shape[0])
    if pts.shape[0] >= 8:
        Do, Io = O.exhaustive_knn(pts, q, 8)
    else:   # oracle needs N >= k: pad with far-away points that can never be in radius
        pad = torch.full((8 - pts.shape[0], 3), 1e6)
        Do, Io = O.exhaustive_knn(torch.cat([pts, pad]), q, 8)
    t_up, t_dn = O.radius_thresholds(radius, Q)
    n_ref = (Do < t_up[:, None]).sum(1).int()
    assert torch.equal(cnt, n_ref), "neighbor_num differs on %d queries" % int((cnt != n_ref).sum())
    keep = Do <= t_dn[:, None]          # slots that can carry weight (decoder.py:236)
    assert torch.equal(I[keep].long(), Io[keep]), "indices differ"
    assert torch.equal(D[keep], Do[keep]), "distances differ"
    assert bool((I[~keep] == -1).all()) and bool((D[~keep] == torch.finfo(torch.float32).max).all())
    return cnt


def test_knn_golden_fixture(hps, golden):
    scene, knn = golden("scene_small.pt"), golden("knn_small.pt")
    for level in ("mid", "fine"):
        k = knn[level]
        npc = hps.NeuralPointCloud(hps.scannet_cfg(DEV))
        npc.set_cloud(level, scene["clouds"][level]["points"].to(DEV))
        D, I, n = npc.find_neighbors_faiss(k["q"].to(DEV), step="query", dynamic_radius=k["r"].to(DEV), level=level)
        assert D.dtype == torch.float32 and I.dtype == torch.int64 and n.dtype == torch.int32
        assert torch.equal(n.cpu(), k["n"])
        _, t_dn = O.radius_thresholds(k["r"])
        keep = k["D"] <= t_dn[:, None]
        assert torch.equal(I.cpu()[keep], k["I"].long()[keep]) and torch.equal(D.cpu()[keep], k["D"][keep])
        _, _, ns = npc.find_neighbors_faiss(k["q"].to(DEV), step="query", level=level)   # scalar radius_query
        assert torch.equal(ns.cpu(), k["n_scalar"])


@pytest.mark.parametrize("n,q,rmin,rmax,cell", [
    (20000, 3000, 0.04, 0.16, 0.08),
    (20000, 2000, 0.24, 0.96, 0.48),
    (5000, 1000, 0.05, 0.3, None),       # cell derived from the batch
    (5000, 1000, 0.3, 0.3, 0.02),        # cells much smaller than the radius (multi-pass staging)
    (3, 50, 0.5, 2.0, 0.5),              # fewer points than k
    (1, 10, 0.5, 2.0, 0.5),
])
def test_knn_random_vs_exhaustive(hps, n, q, rmin, rmax, cell):
    g = torch.Generator().manual_seed(n + q)
    sc = O.synthetic_scene(max(n, 8), 8, seed=n)
    pts = sc["points"][:n].contiguous()
    base = pts[torch.randint(0, n, (q,), generator=g)]
    qs = base + torch.randn(q, 3, generator=g) * rmax * 0.4
    r = rmin + (rmax - rmin) * torch.rand(q, generator=g, dtype=torch.float64)
    cnt = _check_knn(hps, pts, qs, r, cell)
    if n >= 1000:
        assert 0 < cnt.float().mean() < 8.01


def test_knn_edge_cases(hps):
    g = torch.Generator().manual_seed(5)
    pts = torch.rand(4000, 3, generator=g)
    pts[100:140] = pts[100]                       # 40 coincident points: ties broken by index
    pts[200:216, 0] = 0.25                        # exactly representable coordinates
    q = torch.cat([pts[100:101], pts[:200] + 0.01, torch.full((3, 3), 50.0), pts[200:203] + torch.tensor([0.25, 0, 0])])
    r = torch.full((q.shape[0],), 0.25, dtype=torch.float64)      # r^2 = 0.0625 exactly representable
    _check_knn(hps, pts, q, r, 0.125)
    _check_knn(hps, pts, q, 0.08, 0.04)                           # scalar python radius
    _check_knn(hps, pts, q[:0], r[:0], 0.125)                     # empty query batch
    idx = hps.GridIndex(DEV)                                      # empty cloud
    idx.reset(torch.zeros(0, 3))
    idx.cell_hint = 0.1
    res = idx.radius_search(q.to(DEV), 0.3)
    assert bool((res.I == -1).all()) and bool((res.count == 0).all())


def test_knn_index_search_is_exact_knn(hps):
    g = torch.Generator().manual_seed(9)
    pts = torch.rand(6000, 3, generator=g) * 2
    q = torch.rand(300, 3, generator=g) * 2.4 - 0.2
    idx = hps.GridIndex(DEV)
    idx.train(pts)
    idx.add(pts[:3000].to(DEV))
    idx.add(pts[3000:].to(DEV))
    assert idx.ntotal == 6000 and idx.is_trained
    D, I = idx.search(q.to(DEV), 8)
    Do, Io = O.exhaustive_knn(pts, q, 8)
    assert torch.equal(I.cpu(), Io) and torch.equal(D.cpu(), Do)


def test_knn_full_size_properties(hps):
    """BASELINE configs[1] size (1M fine points, 25k queries): size-independent properties +
    exhaustive spot check of 64 queries."""
    sc = O.synthetic_scene(1_000_000, 5000, (0.04, 0.16), seed=77, extent_scale=3.0)
    z, _ = O.surface_z_vals(sc["gt_depth"])
    q = (sc["rays_o"][:, None, :] + sc["rays_d"][:, None, :] * z[..., None]).reshape(-1, 3)
    r = sc["radius"].repeat_interleave(5)
    idx = hps.GridIndex(DEV)
    idx.reset(sc["points"].to(DEV))
    idx.cell_hint = 0.08
    res, cand = idx.radius_search(q.to(DEV), r.to(DEV), want_cand=True)
    res2 = idx.radius_search(q.to(DEV), r.to(DEV))
    assert torch.equal(res.I, res2.I) and torch.equal(res.D, res2.D)                  # idempotent
    I, D, cnt = res.I.cpu(), res.D.cpu(), res.count.cpu()
    used = I >= 0
    assert bool((D[:, 1:] >= D[:, :-1]).all())                                        # ascending
    assert bool((used[:, 1:] <= used[:, :-1]).all())                                  # used slots are a prefix
    t_up, t_dn = O.radius_thresholds(r)
    assert bool((D[used] <= t_dn[:, None].expand_as(D)[used]).all())
    assert torch.equal(cnt, (D < t_up[:, None]).sum(1).int())
    pi = I.clamp(min=0).long()
    d = q[:, None, :] - sc["points"][pi]
    Dre = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    assert torch.equal(Dre[used], D[used])                                            # D is the stated formula
    assert bool((cand.cpu() >= used.sum(1)).all())
    sel = torch.randperm(q.shape[0], generator=torch.Generator().manual_seed(1))[:64]
    Do, Io = O.exhaustive_knn(sc["points"], q[sel], 8)
    keep = Do <= t_dn[sel][:, None]
    assert torch.equal(I[sel][keep].long(), Io[keep]) and torch.equal(D[sel][keep], Do[keep])
    info = idx._grid.info()
    assert info["status"] == 0 and info["n_points"] == 1_000_000


@pytest.mark.parametrize("tracker", [False, True])
def test_interp_fwd_bwd(hps, tracker):
    sc = O.synthetic_scene(8000, 200, (0.04, 0.16), seed=3)
    z, _ = O.surface_z_vals(sc["gt_depth"])
    p = (sc["rays_o"][:, None, :] + sc["rays_d"][:, None, :] * z[..., None]).reshape(-1, 3)
    r = sc["radius"].repeat_interleave(5).reshape(-1, 1)
    fill_a, fill_b = torch.randn(32) * 0.01, torch.randn(32) * 0.01
    # oracle
    po = p.clone().requires_grad_(tracker)
    fa, fb = sc["geo_feats"].clone().requires_grad_(True), sc["col_feats"].clone().requires_grad_(True)
    D, I, n = O.find_neighbors(sc["points"], p, r)
    ca, has = O.idw_features(po, fa, D, I, n, r, fill_a, 2, tracker, sc["points"])
    cb, _ = O.idw_features(po, fb, D, I, n, r, fill_b, 2, tracker, sc["points"])
    ga, gb = torch.randn_like(ca), torch.randn_like(cb)
    ((ca * ga).sum() + (cb * gb).sum()).backward()
    # product
    idx = hps.GridIndex(DEV)
    idx.reset(sc["points"].to(DEV))
    idx.cell_hint = 0.08
    qr = idx.radius_search(p.to(DEV), r.to(DEV))
    pd = p.to(DEV).clone().requires_grad_(tracker)
    fad, fbd = sc["geo_feats"].to(DEV).requires_grad_(True), sc["col_feats"].to(DEV).requires_grad_(True)
    cad, cbd, hasd = hps.ops.interpolate(pd, fad, fbd, idx.positions if tracker else None, qr.I, qr.D, qr.t_dn,
                                         qr.count, fill_a.to(DEV), fill_b.to(DEV), 2, tracker)
    ((cad * ga.to(DEV)).sum() + (cbd * gb.to(DEV)).sum()).backward()
    assert torch.equal(hasd.cpu(), has)
    assert 0.2 < has.float().mean() < 1.0
    assert parity.rel(cad, ca) < 1e-5 and parity.rel(cbd, cb) < 1e-5
    assert parity.rel(fad.grad, fa.grad) < 1e-5 and parity.rel(fbd.grad, fb.grad) < 1e-5
    if tracker:
        assert parity.rel(pd.grad, po.grad) < TOL
    # single-table variant (geometry stages)
    c1, c2, _ = hps.ops.interpolate(p.to(DEV), sc["geo_feats"].to(DEV), None, None, qr.I, qr.D, qr.t_dn, qr.count,
                                    fill_a.to(DEV), None, 2, False)
    assert c2 is None and parity.rel(c1, O.idw_features(p, sc["geo_feats"], D, I, n, r, fill_a)[0]) < 1e-5


@pytest.mark.parametrize("kind,P,affine", [("geo", 200, False), ("col", 200, False), ("col", 333, True),
                                           ("geo", 64, False), ("col", 1, True)])
def test_mlp_fwd_bwd(hps, golden, kind, P, affine):
    scene = golden("scene_small.pt")
    name = "geo_decoder_fine" if kind == "geo" else "color_decoder_fine"
    prm = scene["decoders"][name]
    g = torch.Generator().manual_seed(P)
    p = torch.randn(P, 3, generator=g) * 1.5 + torch.tensor([0.0, 0.0, -2.0])
    c = torch.randn(P, 32, generator=g) * 0.1
    ef = torch.randn(8, generator=g) * 0.5
    # oracle
    po, co = p.clone().requires_grad_(True), c.clone().requires_grad_(True)
    prmo = {k: v.clone().requires_grad_(True) for k, v in prm.items()}
    efo = ef.clone().requires_grad_(True)
    out_o = O.geo_mlp(po, co, prmo) if kind == "geo" else O.col_mlp(po, co, prmo, efo if affine else None)
    gout = torch.randn(out_o.shape, generator=g)
    (out_o * gout).sum().backward()
    # product
    _, _, dec, _ = parity.product(scene, DEV)
    m = getattr(dec, name)
    pd, cd = p.to(DEV).requires_grad_(True), c.to(DEV).requires_grad_(True)
    efd = ef.to(DEV).requires_grad_(True)
    out_d = m(pd, cd, efd if affine else None)
    assert out_d.shape == out_o.shape
    (out_d * gout.to(DEV)).sum().backward()
    assert parity.rel(out_d, out_o) < TOL
    assert parity.rel(cd.grad, co.grad) < TOL and parity.rel(pd.grad, po.grad) < TOL
    named = dict(m.named_parameters())
    for k, v in prmo.items():
        if v.grad is None:
            continue
        if k == "embedder._B" and kind == "col":
            continue
        assert named[k].grad is not None, k
        assert parity.rel(named[k].grad, v.grad) < TOL, k
    if affine:
        assert parity.rel(efd.grad, efo.grad) < TOL
    # inference path (no activation workspace) gives the same output
    with torch.no_grad():
        out_i = m(p.to(DEV), c.to(DEV), ef.to(DEV) if affine else None)
    assert torch.equal(out_i, out_d.detach())


@pytest.mark.parametrize("N,K,a_major,b_major", [(128, 128, 0, 0), (32, 64, 0, 0), (128, 128, 1, 1), (192, 128, 1, 1),
                                                 (32, 128, 1, 1), (128, 64, 0, 1), (64, 128, 1, 0), (176, 128, 1, 1)])
def test_tcgen05_descriptor_selftest(hps, N, K, a_major, b_major):
    """K-major and MN-major SWIZZLE_128B operand descriptors: exact integer GEMM through tcgen05.mma."""
    g = torch.Generator().manual_seed(N * 1000 + K + a_major * 2 + b_major)
    A = torch.randint(-4, 5, (128, K), generator=g).float()
    B = torch.randint(-4, 5, (N, K), generator=g).float()
    want = A @ B.T
    Ad = (A if a_major == 0 else A.T).contiguous().to(DEV)
    Bd = (B if b_major == 0 else B.T).contiguous().to(DEV)
    D = torch.full((128, N), float("nan"), device=DEV)
    L = hps._lib
    hps.ops._call("hps_tc_selftest", L.ptr(Ad), L.ptr(Bd), L.ptr(D), N, K, a_major, b_major, hps.ops._stream())
    torch.cuda.synchronize()
    assert torch.equal(D.cpu(), want), "max err %g" % float((D.cpu() - want).abs().max())


@pytest.mark.parametrize("kind,P,affine", [("col", 1000, False), ("col", 130, True), ("geo", 1000, False), ("geo", 77, False),
                                           ("col", 40000, False), ("geo", 40000, False)])
def test_mlp_tensor_core_edition_matches_simt(hps, golden, kind, P, affine):
    """tcgen05 edition (forward + dgrad / wgrad / wfinal backward) vs the fp32 SIMT edition of the decoders, up to the
    bench's point count: outputs agree to 1e-5 (split-precision MMA, measured ~1e-6), input, feature and every
    parameter gradient to 1e-4 - this is where the backward's dy-free algebra (dc through Q = W Wc, dWc / dbc through the
    (H x 32) sums) is checked against a plain chain-rule implementation at scale."""
    scene = golden("scene_small.pt")
    name = "geo_decoder_fine" if kind == "geo" else "color_decoder_fine"
    g = torch.Generator().manual_seed(P)
    p = (torch.randn(P, 3, generator=g) * 1.5 + torch.tensor([0.0, 0.0, -2.0])).to(DEV)
    c = (torch.randn(P, 32, generator=g) * 0.1).to(DEV)
    ef = (torch.randn(8, generator=g) * 0.5).to(DEV)
    gout = torch.randn((P,) if kind == "geo" else (P, 3), generator=g).to(DEV)
    _, _, dec, _ = parity.product(scene, DEV)
    m = getattr(dec, name)
    res = {}
    for tc in (False, True):
        hps.ops.USE_TC = tc
        try:
            m.zero_grad()
            pd, cd = p.clone().requires_grad_(True), c.clone().requires_grad_(True)
            out = m(pd, cd, ef if affine else None)
            (out * gout).sum().backward()
            res[tc] = (out.detach(), pd.grad, cd.grad, {k: v.grad.clone() for k, v in m.named_parameters() if v.grad is not None})
        finally:
            hps.ops.USE_TC = True
    assert parity.rel(res[True][0], res[False][0]) < 1e-5

    def rows_off(a, b):
        """fraction of rows whose error exceeds TOL x max|b|"""
        e = (a - b).abs().reshape(a.shape[0], -1).amax(1)
        return float((e > TOL * float(b.abs().max())).float().mean())
    if kind == "col":       # Softplus decoder: smooth, gradients agree everywhere
        assert parity.rel(res[True][1], res[False][1]) < TOL and parity.rel(res[True][2], res[False][2]) < TOL
    else:
        # ReLU decoder: a pre-activation within ~1e-6 of zero may take the other branch of ReLU' in the two
        # editions (the forward values differ by ~1e-6); that flips one point's input gradient.  Allow 0.1% of points.
        assert rows_off(res[True][1], res[False][1]) < 1e-3 and rows_off(res[True][2], res[False][2]) < 1e-3
    for k, gq in res[False][3].items():
        assert parity.rel(res[True][3][k], gq) < (TOL if kind == "col" else 5e-3), k


@pytest.mark.parametrize("R,S", [(257, 5), (64, 8), (33, 12), (1, 5)])
def test_composite_fwd_bwd(hps, R, S):
    g = torch.Generator().manual_seed(R * S)
    raw = torch.randn(R, S, 4, generator=g)
    raw[..., 3] *= 30.0
    z = torch.sort(torch.rand(R, S, generator=g) * 3 + 0.5, dim=1).values
    mask = torch.rand(R * S, generator=g) > 0.2
    ro = raw.clone().requires_grad_(True)
    occ = ro[..., 3].reshape(-1)
    occ = occ + (torch.where(mask, occ, torch.full_like(occ, -100.0)) - occ).detach()
    d, v, c, w = O.composite(torch.cat([ro[..., :3], occ.reshape(R, S, 1)], -1), z, 0.1)
    gd, gc = torch.randn(R, generator=g), torch.randn(R, 3, generator=g)
    ((d * gd).sum() + (c * gc).sum()).backward()
    rd = raw.to(DEV).requires_grad_(True)
    dd, vd, cd, wd = hps.ops.composite(rd, z.to(DEV), mask.to(DEV), 0.1, want_weights=True)
    ((dd * gd.to(DEV)).sum() + (cd * gc.to(DEV)).sum()).backward()
    assert parity.rel(dd, d) < 1e-6 and parity.rel(vd, v) < 1e-5 and parity.rel(cd, c) < 1e-6
    assert parity.rel(wd, w) < 1e-6
    assert parity.rel(rd.grad, ro.grad) < TOL


def test_raw2outputs_dropin(hps):
    g = torch.Generator().manual_seed(2)
    raw = torch.randn(100, 5, 4, generator=g)
    z = torch.sort(torch.rand(100, 5, generator=g) + 1, dim=1).values
    d, v, c, w = O.composite(raw, z, 0.1)
    rawd = raw.to(DEV)
    dd, vd, cd, wd = hps.raw2outputs_nerf_color(rawd, z.to(DEV), None, occupancy=True, device=DEV, coef=0.1)
    assert parity.rel(dd, d) < 1e-6 and parity.rel(cd, c) < 1e-6 and parity.rel(wd, w) < 1e-6
    assert parity.rel(rawd[..., 3], torch.sigmoid(0.1 * raw[..., 3])) < 1e-6      # in-place alpha, common.py:414
    with pytest.raises(NotImplementedError):
        hps.raw2outputs_nerf_color(rawd, z.to(DEV), None, occupancy=False)


def test_exposure_mlp_matches_oracle(hps, golden):
    """decoder.py:380-396 MLP_exposure: fused kernel pair vs the oracle restatement (fp64), forward and all gradients."""
    torch.manual_seed(7)
    prm = {"mlp_exposure.linear1.weight": torch.randn(128, 8) * 0.3, "mlp_exposure.linear1.bias": torch.randn(128) * 0.05,
           "mlp_exposure.linear2.weight": torch.randn(12, 128) * 0.1, "mlp_exposure.linear2.bias": torch.randn(12) * 0.1}
    x = torch.randn(8) * 0.5
    ref_in = {k: v.double().requires_grad_(True) for k, v in prm.items()}
    xr = x.double().requires_grad_(True)
    ref = O.exposure_affine(ref_in, xr)
    g = torch.randn(12, dtype=torch.float64)
    (ref * g).sum().backward()
    dev_in = [x.to(DEV).requires_grad_(True)] + [prm[k].to(DEV).requires_grad_(True) for k in (
        "mlp_exposure.linear1.weight", "mlp_exposure.linear1.bias", "mlp_exposure.linear2.weight", "mlp_exposure.linear2.bias")]
    out = hps.ops.exposure_mlp(*dev_in)
    (out * g.float().to(DEV)).sum().backward()
    # tolerance: 1e-4 relative to the largest reference entry (fp32 kernel vs fp64 oracle)
    def close(a, b):
        return float((a.detach().cpu().double() - b).abs().max()) <= TOL * max(float(b.abs().max()), 1e-12)
    assert close(out, ref.detach())
    assert close(dev_in[0].grad, xr.grad)
    for t, k in zip(dev_in[1:], ("mlp_exposure.linear1.weight", "mlp_exposure.linear1.bias", "mlp_exposure.linear2.weight",
                                 "mlp_exposure.linear2.bias")):
        assert close(t.grad, ref_in[k].grad), k


def test_composite_bwd_reports_gradient_maxima(hps):
    """The optional abs-max words of hps_composite_bwd equal max|d_rgb| / max|d_occ| (consumed by hps_mlp_tc_bwd)."""
    import ctypes
    from hps_b200 import _lib
    torch.manual_seed(3)
    R, S = 301, 5
    rgb, occ = torch.rand(R * S, 3, device=DEV), torch.randn(R * S, device=DEV) * 5
    z = torch.sort(torch.rand(R, S, device=DEV) + 1, dim=1).values
    gd, gc = torch.randn(R, device=DEV) * 1e-3, torch.randn(R, 3, device=DEV) * 1e-3
    d_rgb, d_occ = torch.empty_like(rgb), torch.empty_like(occ)
    words = torch.zeros(2, dtype=torch.int32, device=DEV)
    lib = _lib.load()
    rc = lib.hps_composite_bwd(_lib.ptr(rgb), 3, _lib.ptr(occ), 1, None, _lib.ptr(z), R, S, 0.1, _lib.ptr(gd), _lib.ptr(gc),
                               None, _lib.ptr(d_rgb), 3, _lib.ptr(d_occ), 1, ctypes.c_void_p(words.data_ptr()),
                               ctypes.c_void_p(words.data_ptr() + 4), None)
    _lib.check(rc)
    torch.cuda.synchronize()
    got = words.view(torch.float32).cpu()
    assert float(got[0]) == float(d_rgb.abs().max()) and float(got[1]) == float(d_occ.abs().max())


@pytest.mark.parametrize("Q", [5, 23, 5000])
def test_knn_ray_grouped_sweep_equals_per_query(hps, Q):
    """hps_knn_query(group=5) (the samples of a ray share voxel probes and candidate loads) returns bit-identical
    neighbour lists, distances and counts to group=1, including a ragged last group, dead queries (t_dn < 0 cannot
    occur through the host API, so: far-away queries) and mixed radii inside a group."""
    g = torch.Generator().manual_seed(11 + Q)
    pts = torch.rand(40000, 3, generator=g) * 2 - 1
    # rays: 5 samples close to each other, like Renderer.py:163-229
    o = torch.rand((Q + 4) // 5, 1, 3, generator=g) * 2 - 1
    d = torch.nn.functional.normalize(torch.randn((Q + 4) // 5, 1, 3, generator=g), dim=-1)
    z = torch.linspace(0.0, 0.12, 5).view(1, 5, 1)
    q = (o + d * z).reshape(-1, 3)[:Q].contiguous()
    q[::37] += 50.0                                            # queries with no voxel at all
    radius = 0.04 + 0.12 * torch.rand(Q, generator=g, dtype=torch.float64)
    idx = hps.GridIndex(DEV)
    idx.reset(pts.to(DEV))
    t_up, t_dn = hps.ops.radius_thresholds(radius.to(DEV), Q, DEV)
    grid = idx._ensure_grid(t_dn)
    a = grid.query(q.to(DEV), t_up, t_dn, group=1)
    b = grid.query(q.to(DEV), t_up, t_dn, group=5)
    for x, y, name in zip(a, b, ("I", "D", "count")):
        assert torch.equal(x, y), name
    assert int(a[2].max()) > 0


def test_pose_helpers_match_torch_expressions(hps):
    """get_camera_from_tensor / get_rays_from_uv (common.py:304-346, 103-119) as single kernels: forward bit-identical
    to the reference's torch expressions evaluated on the same device, backward (analytic) within 1e-6 of autograd."""
    K = O.SCANNET_INTRINSICS
    g = torch.Generator().manual_seed(21)
    for trial in range(4):
        cam = torch.randn(7, generator=g)
        cam[:4] = torch.nn.functional.normalize(cam[:4], dim=0) * (1.0 + 0.05 * trial)     # also off the unit sphere
        R = 777
        pi = torch.randint(0, K["W"], (R,), generator=g).float().to(DEV)
        pj = torch.randint(0, K["H"], (R,), generator=g).float().to(DEV)
        w_o, w_d = torch.randn(R, 3, generator=g).to(DEV), torch.randn(R, 3, generator=g).to(DEV)
        # reference expressions (oracle restatement, run on the GPU so that rounding is torch-CUDA's)
        cr = cam.to(DEV).clone().requires_grad_(True)
        c2w_r = O.camera_from_tensor(cr)
        ro_r, rd_r = O.rays_from_uv(pi, pj, c2w_r, K["fx"], K["fy"], K["cx"], K["cy"])
        ((ro_r * w_o).sum() + (rd_r * w_d).sum()).backward()
        # product
        cp = cam.to(DEV).clone().requires_grad_(True)
        c2w_p = hps.get_camera_from_tensor(cp)
        ro_p, rd_p = hps.get_rays_from_uv(pi, pj, c2w_p, K["H"], K["W"], K["fx"], K["fy"], K["cx"], K["cy"], DEV)
        ((ro_p * w_o).sum() + (rd_p * w_d).sum()).backward()
        assert torch.equal(c2w_p, c2w_r.detach()), "camera matrix differs"
        assert torch.equal(rd_p, rd_r.detach()) and torch.equal(ro_p, ro_r.detach()), "rays differ"
        err = float((cp.grad - cr.grad).abs().max()) / float(cr.grad.abs().max())
        assert err < 1e-5, err            # sums over 777 rays in a different order


def test_row_adam_matches_torch_adam_and_writes_table(hps):
    """RowAdam (Mapper.py:592-603, 859 + the row write-back :862-874) vs torch.optim.Adam on the same compact rows over
    several steps with changing gradients and a learning-rate change; the full table must carry the new rows."""
    torch.manual_seed(5)
    N, M, C = 5000, 777, 32
    table = torch.randn(N, C, device=DEV) * 0.1
    idx = torch.randperm(N, device=DEV)[:M]
    rows = hps.LiveRows(table.clone(), idx)
    ref_p = rows.param.detach().clone().requires_grad_(True)
    ref = torch.optim.Adam([ref_p], lr=5e-3, betas=(0.9, 0.999), eps=1e-8)
    opt = hps.RowAdam([(rows, 5e-3)])
    for it in range(7):
        g = torch.randn(M, C, device=DEV) * (10.0 ** (-it % 4))
        if it == 4:
            for grp in ref.param_groups:
                grp["lr"] = 1e-3
            opt.set_lr(0, 1e-3)
        ref_p.grad = g.clone()
        rows.param.grad = g.clone()
        ref.step()
        opt.step()
        # fp32, same formula: differences are a few ulp of the update (tolerance 1e-6 of the parameter scale)
        assert float((rows.param.detach() - ref_p.detach()).abs().max()) <= 1e-6 * float(ref_p.detach().abs().max())
        assert torch.equal(rows.table[idx], rows.param.detach())
        assert rows.fresh
    untouched = torch.ones(N, dtype=torch.bool, device=DEV)
    untouched[idx] = False
    assert torch.equal(rows.table[untouched], table[untouched])
    assert int(opt.state[0]["step"][0]) == 7 and int(opt.state[0]["step"][1]) == 0


def test_fused_adam_matches_torch_adam(hps):
    """FusedAdam (the tracker's optimiser, Tracker.py:383-416: pose / exposure latent / exposure MLP groups with their
    own learning rates) vs torch.optim.Adam over several steps; gradients accumulate into the optimiser's buffers and
    are cleared by the step kernel; the whole loss.backward(); step() pair replays from a CUDA graph."""
    torch.manual_seed(11)
    shapes = [(7,), (8,), (128, 8), (128,), (12, 128), (12,), (3,), (4,)]
    lrs = [5e-4, 1e-3, 1e-3, 1e-3, 1e-3, 1e-3, 2e-3, 4e-4]
    mine = [torch.randn(s, device=DEV).requires_grad_(True) for s in shapes]
    ref = [p.detach().clone().requires_grad_(True) for p in mine]
    groups = [([mine[0]], lrs[0]), (mine[1:6], lrs[1]), (mine[6], lrs[6]), ([mine[7]], lrs[7])]
    opt = hps.FusedAdam(groups)
    topt = torch.optim.Adam([{"params": [ref[0]], "lr": lrs[0]}, {"params": ref[1:6], "lr": lrs[1]},
                             {"params": [ref[6]], "lr": lrs[6]}, {"params": [ref[7]], "lr": lrs[7]}])
    coef = [torch.randn(s, device=DEV) for s in shapes]

    def loss_of(ps, k):
        return sum(((p * c).sum() * (k + 1.0)) ** 2 + (p ** 3).sum() * 0.1 for p, c in zip(ps, coef))

    for k in range(6):
        if k == 3:
            opt.set_lr(0, 1e-4)
            topt.param_groups[0]["lr"] = 1e-4
        opt.zero_grad()
        topt.zero_grad()
        loss_of(mine, k).backward()
        loss_of(ref, k).backward()
        opt.step()
        topt.step()
        for a, b in zip(mine, ref):
            assert float((a.detach() - b.detach()).abs().max()) <= 2e-6 * max(1.0, float(b.detach().abs().max()))
            assert float(a.grad.abs().max()) == 0.0             # cleared by the step kernel
    assert int(opt.step_count[0, 0]) == 6 and int(opt.step_count[0, 1]) == 0

    # graph replay: the device step counter advances, results keep following torch.optim.Adam
    def it():
        loss_of(mine, 1).backward()
        opt.step()
    g = hps.graphs.GraphedStep(it, {}, warmup=1)                  # one executed warm-up step, then the recording
    topt.zero_grad()
    loss_of(ref, 1).backward()
    topt.step()
    for _ in range(3):
        g()
        topt.zero_grad()
        loss_of(ref, 1).backward()
        topt.step()
    torch.cuda.synchronize()
    assert int(opt.step_count[0, 0]) == 10
    for a, b in zip(mine, ref):
        assert float((a.detach() - b.detach()).abs().max()) <= 5e-6 * max(1.0, float(b.detach().abs().max()))
    with pytest.raises(hps._lib.HpsError):
        mine[0].grad = torch.zeros_like(mine[0])
        opt.step()


def test_sample_rays_bwd_sums_its_three_terms(hps):
    R, S = 333, 5
    g = torch.Generator(device=DEV).manual_seed(2)
    dps = [torch.randn(R * S, 3, device=DEV, generator=g) for _ in range(3)]
    z = torch.rand(R, S, device=DEV, generator=g) + 1
    L = hps._lib
    for n in (1, 2, 3):
        go, gd = torch.empty(R, 3, device=DEV), torch.empty(R, 3, device=DEV)
        hps.ops._call("hps_sample_rays_bwd", L.ptr(dps[0]), L.ptr(dps[1]) if n > 1 else None, L.ptr(dps[2]) if n > 2 else None,
                      L.ptr(z), R, S, L.ptr(go), L.ptr(gd), hps.ops._stream())
        dp = sum(dps[:n]).view(R, S, 3)
        assert torch.allclose(go, dp.sum(1), rtol=1e-5, atol=1e-5)
        assert torch.allclose(gd, (dp * z[..., None]).sum(1), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("layout", ["segments", "interleaved"])
def test_mapper_loss_with_several_keyframes(hps, layout):
    """A mapper batch mixes rays of several keyframes, each with its own exposure affine (Mapper.py:805-822): the loss
    kernel against the reference's expressions written out in torch (slice-wise colour @ rot + trans, sigmoid, masked
    L1 sums) under autograd.  'segments' = contiguous runs per frame (block sums), 'interleaved' = frames change from
    ray to ray (per-ray atomics)."""
    torch.manual_seed(8)
    R, F, w = 1500, 4, 0.1
    depth = (1 + torch.rand(R, device=DEV)).requires_grad_(True)
    color = torch.rand(R, 3, device=DEV).requires_grad_(True)
    gt_d = 1 + torch.rand(R, device=DEV)
    gt_d[::9] = 0
    gt_c = torch.rand(R, 3, device=DEV)
    valid = torch.rand(R, device=DEV) > 0.1
    affine = (torch.eye(3, device=DEV).reshape(1, 9).repeat(F, 1) + 0.2 * torch.randn(F, 9, device=DEV))
    affine = torch.cat([affine, 0.1 * torch.randn(F, 3, device=DEV)], 1).requires_grad_(True)
    if layout == "segments":
        bounds = [0, 300, 777, 778, R]
        frame = torch.cat([torch.full((bounds[i + 1] - bounds[i],), i, device=DEV) for i in range(F)])
    else:
        frame = torch.arange(R, device=DEV) % F
    loss = hps.mapper_loss(depth, color, valid, gt_d, gt_c, "color_fine", w, affine, frame_index=frame)
    (loss * 3.0).backward()
    got = [loss.detach(), depth.grad.clone(), color.grad.clone(), affine.grad.clone()]
    for t in (depth, color, affine):
        t.grad = None
    mask = (gt_d > 0) & valid & ~torch.isnan(depth)
    rot, trans = affine[:, :9].reshape(F, 3, 3), affine[:, 9:]
    c2 = torch.sigmoid(torch.einsum("rk,rkc->rc", color, rot[frame]) + trans[frame])
    ref = (gt_d[mask] - depth[mask]).abs().sum() + w * (gt_c[mask] - c2[mask]).abs().sum()
    (ref * 3.0).backward()
    want = [ref.detach(), depth.grad, color.grad, affine.grad]
    for name, a, b in zip(("loss", "g_depth", "g_color", "d_affine"), got, want):
        err = float((a - b).abs().max() / (b.abs().max() + 1e-30))
        assert err < 1e-5, (name, err)
    with pytest.raises(ValueError):
        hps.mapper_loss(depth, color, valid, gt_d, gt_c, "color_fine", w, affine)          # (F,12) needs frame_index


def test_fused_adam_on_handed_over_gradients(hps):
    """FusedAdam(own_grads=False): backward hands its gradient buffers over (zero_grad sets them to None), the step
    kernel reads them where they are; parameters without a gradient are skipped."""
    torch.manual_seed(12)
    mine = [torch.randn(s, device=DEV).requires_grad_(True) for s in ((7,), (8,), (5, 3))]
    ref = [p.detach().clone().requires_grad_(True) for p in mine]
    opt = hps.FusedAdam([(mine[:2], 5e-4), ([mine[2]], 1e-3)], own_grads=False)
    topt = torch.optim.Adam([{"params": ref[:2], "lr": 5e-4}, {"params": [ref[2]], "lr": 1e-3}])
    assert all(p.grad is None for p in mine)
    for k in range(5):
        opt.zero_grad()
        topt.zero_grad()
        for ps in (mine, ref):
            (sum((p * (k + 1.0)).sin().sum() for p in ps)).backward()
        held = [p.grad.data_ptr() for p in mine]
        opt.step()
        topt.step()
        assert [p.grad.data_ptr() for p in mine] == held and float(mine[0].grad.abs().max()) > 0    # left as they were
        for a, b in zip(mine, ref):
            assert float((a.detach() - b.detach()).abs().max()) <= 2e-6 * max(1.0, float(b.detach().abs().max()))
    opt.zero_grad()
    assert all(p.grad is None for p in mine)
    opt.step()                                                       # nothing to do, no launch
    assert int(opt.step_count[0, 0]) == 5


def test_mapper_optimizer_follows_torch_adam_through_the_stages(hps):
    """MapperOptimizer (RowAdam per feature table + FusedAdam with per-tensor step counters for decoders / poses /
    exposure latent) vs the reference's construction: ONE torch.optim.Adam over the seven groups of Mapper.py:592-603
    whose learning rates are rewritten per iteration (:637-667).  A stage gives gradients to some tensors only; torch
    skips the others and keeps their own step counts - so must we (51 small tensors; a colour stage touches 27 of them = two launches, whose
    make-up changes with the stage)."""
    torch.manual_seed(21)
    mo = hps.mapper_optim
    cfg = hps.scannet_cfg()["mapping"]
    N, M, C = 900, 130, 32
    shapes = [(32, 96), (32,), (32, 32), (32,), (4, 32), (4,), (128, 40), (128,), (128, 168), (128,), (3, 128), (3,),
              (40, 3), (5, 7), (9,), (32, 32), (32,), (128, 128), (128,), (128, 128), (128,), (1, 300), (2,), (257,)]
    dec_mid = [torch.randn(s, device=DEV).requires_grad_(True) for s in shapes]
    dec_fine = [torch.randn(s, device=DEV).requires_grad_(True) for s in shapes]
    cams = [torch.randn(7, device=DEV).requires_grad_(True) for _ in range(2)]
    expo = torch.randn(8, device=DEV).requires_grad_(True)
    tab = {k: torch.randn(N, C, device=DEV) * 0.1 for k in mo.MapperOptimizer.ROW_KEYS}
    idx = torch.randperm(N, device=DEV)[:M]
    g_mid, c_mid = hps.LiveRows.pair(tab["geometry_mid"].clone(), tab["color_mid"].clone(), idx)   # selected rows
    g_fine = tab["geometry_fine"].clone().requires_grad_(True)                                    # whole tables
    c_fine = tab["color_fine"].clone().requires_grad_(True)
    rows = {"geometry_mid": g_mid, "color_mid": c_mid, "geometry_fine": g_fine, "color_fine": c_fine}
    opt = mo.MapperOptimizer(cfg, decoder_params=dec_mid + dec_fine, rows=rows, cameras=cams, exposure_feat=expo)
    assert len(opt.small_opt.params) == 51 and opt.small_opt.step_count.shape[0] == 3

    leaf = {"geometry_mid": g_mid.param, "color_mid": c_mid.param, "geometry_fine": g_fine, "color_fine": c_fine}
    mine = dec_mid + dec_fine + [leaf[k] for k in mo.MapperOptimizer.ROW_KEYS] + cams + [expo]
    ref = [p.detach().clone().requires_grad_(True) for p in mine]
    nd = 2 * len(shapes)
    topt = torch.optim.Adam([{"params": ref[:nd], "lr": 0}, {"params": [ref[nd]], "lr": 0},
                             {"params": [ref[nd + 1]], "lr": 0}, {"params": [ref[nd + 2]], "lr": 0},
                             {"params": [ref[nd + 3]], "lr": 0}, {"params": ref[nd + 4:nd + 6], "lr": 0},
                             {"params": [ref[nd + 6]], "lr": 0.001}])
    # which tensors a stage gives gradients to (positions in `mine`)
    ns = len(shapes)
    touched = {"geometry_mid": list(range(0, 6)) + [nd],
               "color_mid": list(range(0, ns)) + [nd, nd + 2, nd + 4, nd + 5, nd + 6],
               "geometry_fine": list(range(ns, ns + 6)) + [nd + 1],
               "color_fine": list(range(ns, 2 * ns)) + [nd + 1, nd + 3, nd + 4, nd + 5, nd + 6]}
    n_iters, seen = 24, set()
    expect_steps = [0] * len(mine)
    for j in range(n_iters):
        stage = opt.set_iteration(j, n_iters, BA=True)
        seen.add(stage)
        lr = mo.stage_learning_rates(cfg, stage)
        cam_lr = mo.ba_camera_lr(j, n_iters, cfg["mid_iter_ratio"], cfg["geo_iter_ratio"], cfg["BA_cam_lr"])
        for grp, v in zip(topt.param_groups[:6], [lr["decoders"], lr["geometry_mid"], lr["geometry_fine"],
                                                  lr["color_mid"], lr["color_fine"], cam_lr]):
            grp["lr"] = v
        opt.zero_grad()
        topt.zero_grad()
        assert all(p.grad is None for p in mine)
        for k in touched[stage]:
            g = torch.randn(mine[k].shape, device=DEV) * (10.0 ** (-(j + k) % 3))
            mine[k].grad = g.clone()
            ref[k].grad = g.clone()
            expect_steps[k] += 1
        opt.step()
        topt.step()
        for k, (a, b) in enumerate(zip(mine, ref)):
            err = float((a.detach() - b.detach()).abs().max())
            assert err <= 1e-5 * max(1.0, float(b.detach().abs().max())), (j, stage, k, err)
    assert seen == set(mo.STAGES)
    torch.cuda.synchronize()
    # per-tensor step counters = torch's state[p]['step'];  scratch words back to zero
    small = opt.small_opt
    pos = {id(p): k for k, p in enumerate(mine)}
    for p, st in zip(small.params, small.tensor_steps.tolist()):
        assert st == expect_steps[pos[id(p)]] == int(topt.state[ref[pos[id(p)]]]["step"]), (pos[id(p)], st)
    assert small.step_count.abs().sum().item() == 0
    # live rows were written back into their tables, the other rows left alone
    for r, key in ((g_mid, "geometry_mid"), (c_mid, "color_mid")):
        assert torch.equal(r.table[idx], r.param.detach())
        other = torch.ones(N, dtype=torch.bool, device=DEV)
        other[idx] = False
        assert torch.equal(r.table[other], tab[key][other])
    with pytest.raises(ValueError):
        mo.MapperOptimizer(cfg, rows={"geometry_coarse": g_fine})


def test_knn_largest_cloud_properties(hps):
    """BASELINE configs[4] size (8M fine points, 8x the area of configs[1] at the same ScanNet-like density): the
    size-independent properties of test_knn_full_size_properties + an exhaustive check of 16 queries against the
    oracle's distance formula over all 8M points (in-radius slots ordered by (D, index))."""
    g = torch.Generator().manual_seed(79)
    K, depth = O.SCANNET_INTRINSICS, O.synthetic_depth()
    n, ext, Q = 8_000_000, 28.0, 20_000
    u = torch.rand(n, generator=g) * (K["W"] - 1) * ext - (ext - 1) * K["W"] / 2
    v = torch.rand(n, generator=g) * (K["H"] - 1) * ext - (ext - 1) * K["H"] / 2
    d = depth[v.clamp(0, K["H"] - 1).long(), u.clamp(0, K["W"] - 1).long()]
    d = d * (0.96 + 0.04 * torch.randint(0, 3, (n,), generator=g).float())
    pts = torch.stack([(u - K["cx"]) / K["fx"] * d, -(v - K["cy"]) / K["fy"] * d, -d], -1).float().contiguous()
    q = pts[torch.randint(0, n, (Q,), generator=g)] + (torch.rand(Q, 3, generator=g) - 0.5) * 0.05
    r = 0.04 + 0.12 * torch.rand(Q, generator=g, dtype=torch.float64)
    idx = hps.GridIndex(DEV)
    idx.reset(pts.to(DEV))
    idx.cell_hint = 0.08
    res = idx.radius_search(q.to(DEV), r.to(DEV))
    I, D, cnt = res.I.cpu(), res.D.cpu(), res.count.cpu()
    info = idx._grid.info()
    assert info["status"] == 0 and info["n_points"] == n
    used = I >= 0
    assert int(used.sum()) > 4 * Q                                                    # the queries do find neighbours
    assert bool((D[:, 1:] >= D[:, :-1]).all()) and bool((used[:, 1:] <= used[:, :-1]).all())
    t_up, t_dn = O.radius_thresholds(r)
    assert bool((D[used] <= t_dn[:, None].expand_as(D)[used]).all())
    assert torch.equal(cnt, (D < t_up[:, None]).sum(1).int())
    dd = q[:, None, :] - pts[I.clamp(min=0).long()]
    assert torch.equal(((dd[..., 0] * dd[..., 0] + dd[..., 1] * dd[..., 1]) + dd[..., 2] * dd[..., 2])[used], D[used])
    assert bool((I[used].long() < n).all())
    for j in range(16):
        Dall = O.sqdist_direct(q[j:j + 1], pts)[0]
        inr = torch.nonzero(Dall <= t_dn[j]).flatten()                                # ascending index order
        order = torch.sort(Dall[inr], stable=True).indices[:8]                        # (D, index) order
        want_i, want_d = inr[order], Dall[inr][order]
        m = want_i.numel()
        assert int(used[j].sum()) == m, (j, m)
        assert torch.equal(I[j, :m].long(), want_i) and torch.equal(D[j, :m], want_d)


def test_grid_append_equals_a_fresh_build(hps):
    """hps_grid_append (index.add on a built index, neural_point.py:327): points hashed into the existing structure in
    batches - new surface patches next to the old cloud (new cells), patches far outside the first bounding box (the
    origin's margin), and a batch dropped INTO occupied cells until one outgrows its range (status -> rebuild) - must
    answer every query exactly like an index built from scratch over the same points: indices, distances, counts."""
    g = torch.Generator().manual_seed(5)

    def patch(n, cx, cy, spread=1.0):
        uv = torch.rand(n, 2, generator=g) * spread
        return torch.stack([cx + uv[:, 0], cy + uv[:, 1], -2.0 + 0.05 * torch.sin(3 * uv[:, 0]) + 0.01 * torch.randn(n, generator=g)], 1)

    batches = [patch(20000, 0.0, 0.0), patch(3000, 1.0, 0.0), patch(3000, 0.5, 1.0), patch(2000, -30.0, 12.0),
               patch(500, 0.2, 0.2, 0.05),                       # 500 points into a 5 cm square of the first patch
               patch(3000, 2.0, 0.0)]
    idx = hps.GridIndex(DEV)
    idx.cell_hint = 0.08
    q_all, total = [], 0
    for k, b in enumerate(batches):
        idx.add(b.to(DEV)) if k else idx.reset(b.to(DEV))
        total += b.shape[0]
        q = torch.cat([b[:400] + 0.01 * torch.randn(400, 3, generator=g), batches[0][:200]])
        q_all.append(q)
        qs = torch.cat(q_all).to(DEV)
        rad = (0.04 + 0.12 * torch.rand(qs.shape[0], generator=g, dtype=torch.float64)).to(DEV)
        got = idx.radius_search(qs, rad)
        fresh = hps.GridIndex(DEV)
        fresh.cell_hint = 0.08
        fresh.reset(torch.cat(batches[:k + 1]).to(DEV))
        want = fresh.radius_search(qs, rad)
        assert idx.ntotal == total
        assert torch.equal(got.I, want.I) and torch.equal(got.D, want.D) and torch.equal(got.count, want.count), k
        assert int(got.count.max()) > 0
    # every batch was appended in place: new cells behind the old ones, the far patch inside the origin's margin, and the
    # cells the dense batch overflowed moved to larger ranges
    print("appends", idx.appends, "rebuilds", idx.rebuilds, "failed append status words", idx.append_failures)
    # (status 8 = the build-time capacity, 1.5x the points it was built over, is spent: the one legitimate rebuild here)
    assert idx.appends >= 4 and idx.rebuilds <= 2 and set(idx.append_failures) <= {8}, (idx.appends, idx.rebuilds, idx.append_failures)
    # exhaustive oracle on the final cloud
    pts = torch.cat(batches)
    sub = torch.cat(q_all)[::37]
    r = 0.04 + 0.12 * torch.rand(sub.shape[0], generator=g, dtype=torch.float64)
    De, Ie, ne = O.find_neighbors(pts, sub, r.reshape(-1, 1))
    got = idx.radius_search(sub.to(DEV), r.to(DEV))
    sl = torch.arange(8)[None, :] < ne[:, None].long()
    assert torch.equal(got.count.cpu(), ne)
    assert not ((got.I.cpu().long() != Ie) & sl).any() and not ((got.D.cpu() != De) & sl).any()


def test_point_insertion_appends_in_place(hps):
    """NeuralPointCloud.add_neural_points over several frames: positions and both feature tables grow inside their
    capacity buffers (same storage, old rows untouched), the neighbour grid is appended to, not rebuilt, and
    storage_generation() only moves when a buffer is re-allocated."""
    cfg = hps.scannet_cfg(DEV)
    npc = hps.NeuralPointCloud(cfg)
    sc = O.synthetic_scene(8, 6000, (0.04, 0.16), seed=3)
    ro, rd, gd, gc = (sc[k].to(DEV) for k in ("rays_o", "rays_d", "gt_depth", "gt_color"))
    n0 = int(npc.add_neural_points(ro[:2000], rd[:2000], gd[:2000], gc[:2000], level="fine", idx=0))
    assert n0 > 0
    npc.find_neighbors_faiss(npc.cloud_pos_tensor("fine")[:10], step="query", level="fine")       # builds the grid
    gen, ptr_g, ptr_p = npc.storage_generation("fine"), npc.get_geo_feats("fine").data_ptr(), npc.cloud_pos_tensor("fine").data_ptr()
    row0 = npc.get_geo_feats("fine")[:5].clone()
    index = npc.get_index("fine")
    rebuilds = index.rebuilds
    grew = 0
    for k in range(1, 3):
        sl = slice(2000 * k, 2000 * (k + 1))
        shift = torch.tensor([0.4 * k, 0.0, 0.0], device=DEV)
        n = int(npc.add_neural_points(ro[sl] + shift, rd[sl], gd[sl], gc[sl], level="fine", idx=k))
        grew += n
        npc.find_neighbors_faiss(npc.cloud_pos_tensor("fine")[:10], step="query", level="fine")
    assert grew > 0 and npc.index_ntotal("fine") == 3 * (n0 + grew) == npc.get_geo_feats("fine").shape[0]
    # appended in place; a rebuild only when the grid's build-time capacity (1.5x its points) is spent (status 8)
    assert index.appends >= 1 and set(index.append_failures) <= {8} and index.rebuilds - rebuilds == len(index.append_failures)
    if npc.get_geo_feats("fine").data_ptr() == ptr_g and npc.cloud_pos_tensor("fine").data_ptr() == ptr_p:
        assert npc.storage_generation("fine") == gen                 # nothing moved: captured graphs stay valid
    else:
        assert npc.storage_generation("fine") != gen
    assert torch.equal(npc.get_geo_feats("fine")[:5], row0)


def test_live_rows_write_back_in_one_launch(hps):
    """LiveRows.refresh_pair (hps_rows_store: table[idx] = leaves for both feature tables in one launch, the per-iteration
    write-back of Mapper.py:615-630) against torch's index_copy_, bit for bit: two tables with different row sets, one
    table alone, rows already fresh (RowAdam owns the write-back) skipped, an empty row set."""
    torch.manual_seed(11)
    N, C = 40_000, 32
    ta, tb = torch.randn(N, C, device=DEV), torch.randn(N, C, device=DEV)
    ia, ib = torch.randperm(N, device=DEV)[:11_419], torch.randperm(N, device=DEV)[:777]
    a, b = hps.LiveRows(ta.clone(), ia), hps.LiveRows(tb.clone(), ib)
    with torch.no_grad():
        a.param.copy_(torch.randn_like(a.param))
        b.param.copy_(torch.randn_like(b.param))
    want_a = ta.clone().index_copy_(0, ia, a.param.detach())
    want_b = tb.clone().index_copy_(0, ib, b.param.detach())
    hps.LiveRows.refresh_pair(a, b)
    assert torch.equal(a.table, want_a) and torch.equal(b.table, want_b)
    # one table alone / the other one fresh: untouched
    with torch.no_grad():
        a.param.mul_(2.0)
        b.param.mul_(2.0)
    b.fresh = True
    hps.LiveRows.refresh_pair(a, b)
    assert torch.equal(a.table, ta.clone().index_copy_(0, ia, a.param.detach())) and torch.equal(b.table, want_b)
    b.fresh = False
    hps.LiveRows.refresh_pair(None, b)
    assert torch.equal(b.table, tb.clone().index_copy_(0, ib, b.param.detach()))
    e = hps.LiveRows(tb.clone(), torch.empty(0, dtype=torch.long, device=DEV))
    hps.LiveRows.refresh_pair(e, None)
    assert torch.equal(e.table, tb)
