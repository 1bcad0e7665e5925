"""-m gpu: the drop-in render_batch_ray + losses + backward against fixtures produced by the imported
reference (tests/golden/, oracle/make_golden.py)."""
import pytest
import torch

from oracle import hps_oracle as O
from tests import parity

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def built(hps):
    scene = parity.load("scene_small.pt")
    return (scene,) + parity.product(scene, DEV)


@pytest.mark.parametrize("path", ["fused", "modular", "fused_loss"])
@pytest.mark.parametrize("stage", ["geometry_mid", "color_mid", "geometry_fine", "color_fine"])
def test_mapper_iteration_matches_reference(built, stage, path):
    scene, hps, npc, dec, renderer = built
    ref = parity.load("render_mapper_small.pt")[stage]
    out = parity.mapper_step(hps, npc, dec, renderer, scene, stage, DEV, modular=(path == "modular"),
                             fused_loss=(path == "fused_loss"))
    errs = parity.check_mapper(stage, out, ref, dec)
    print(stage, errs)


@pytest.mark.parametrize("streams", [True, False])
def test_render_batch_ray_loss_equals_render_plus_loss(built, streams):
    """renderer.render_batch_ray_loss (compositing + loss in one kernel, exposure MLP beside the search, upstream
    gradient as a device scalar) against render_batch_ray + mapper_loss: every gradient, including the exposure MLP's
    and the exposure latent's, with an upstream gradient different from 1."""
    scene, hps, npc, dec, renderer = built
    saved = hps.ops.MULTI_STREAM
    hps.ops.MULTI_STREAM = streams
    try:
        for stage in ("color_fine", "color_mid", "geometry_fine"):
            level = "mid" if stage.endswith("mid") else "fine"
            mods = [getattr(dec, "color_decoder_" + level), getattr(dec, "geo_decoder_" + level)]
            grads = []
            for fused in (False, True):
                out = parity.mapper_step(hps, npc, dec, renderer, scene, stage, DEV, fused_loss=fused, upstream=-2.5)
                g = {n + "/" + k: v.grad.clone() for n, m in zip("cg", mods) for k, v in m.named_parameters()
                     if v.grad is not None}
                grads.append((out, g))
            (a, ga), (b, gb) = grads
            assert torch.equal(a["valid"], b["valid"]) and torch.equal(a["depth"], b["depth"])
            assert torch.equal(a["color"], b["color"]) and torch.equal(a["var"], b["var"])
            assert parity.rel(b["loss"], a["loss"]) <= 1e-6
            assert set(ga) == set(gb), set(ga) ^ set(gb)
            if stage.startswith("color"):
                assert any("mlp_exposure" in k for k in gb), "the exposure MLP must receive gradients"
                assert parity.rel(b["g_ef"], a["g_ef"]) <= 1e-5
                assert parity.rel(b["g_col"], a["g_col"]) <= 1e-5
            else:
                assert b["g_ef"] is None
            assert parity.rel(b["g_geo"], a["g_geo"]) <= 1e-5
            worst = max(parity.rel(gb[k], ga[k]) for k in ga)
            assert worst <= 2e-5, (stage, worst)
    finally:
        hps.ops.MULTI_STREAM = saved


@pytest.mark.parametrize("switch", ["simt", "simt_bwd", "single_stream"])
def test_mapper_iteration_other_kernel_editions(built, switch):
    """The fused mapper iteration through the code paths the defaults do not take: fp32 SIMT decoders
    (HPS_USE_TC=0), tensor-core forward with SIMT backward (HPS_USE_TC_BWD=0), and everything on one stream
    (HPS_STREAMS=0) - same fixtures, same bar."""
    scene, hps, npc, dec, renderer = built
    ops = hps.ops
    saved = (ops.USE_TC, ops.USE_TC_BWD, ops.MULTI_STREAM)
    try:
        if switch == "simt":
            ops.USE_TC = False
        elif switch == "simt_bwd":
            ops.USE_TC_BWD = False
        else:
            ops.MULTI_STREAM = False
        for stage in ("color_fine", "geometry_mid"):
            ref = parity.load("render_mapper_small.pt")[stage]
            out = parity.mapper_step(hps, npc, dec, renderer, scene, stage, DEV)
            parity.check_mapper(stage, out, ref, dec)
    finally:
        ops.USE_TC, ops.USE_TC_BWD, ops.MULTI_STREAM = saved


@pytest.mark.parametrize("path", ["fused", "modular", "fused_loss"])
@pytest.mark.parametrize("stage", ["color_mid", "color_fine"])
def test_tracker_iteration_matches_reference(built, stage, path):
    scene, hps, npc, dec, renderer = built
    ref = parity.load("render_tracker_small.pt")
    out = parity.tracker_step(hps, npc, dec, renderer, scene, ref["inputs"], stage, DEV, modular=(path == "modular"),
                              fused_loss=(path == "fused_loss"))
    errs = parity.check_tracker(stage, out, ref[stage], scene, ref["inputs"])
    print(stage, errs)


@pytest.mark.parametrize("stage", ["geometry_fine", "color_fine"])
def test_live_rows_compact_gradients(built, stage):
    """LiveRows (frustum-selected rows as compact leaves, Mapper.py:512-536) gives the same forward and, on the
    live rows, the same gradients as optimising the whole table; frozen rows receive none."""
    scene, hps, npc, dec, renderer = built
    cl, rays, cfg = scene["clouds"], scene["rays"], scene["cfg"]
    level = "fine"
    ref = parity.mapper_step(hps, npc, dec, renderer, scene, stage, DEV)
    g = torch.Generator().manual_seed(3)
    N = cl[level]["points"].shape[0]
    live_idx = torch.randperm(N, generator=g)[: N * 2 // 3].sort().values.to(DEV)
    geo_t, col_t = cl[level]["geo_feats"].to(DEV).clone(), cl[level]["col_feats"].to(DEV).clone()
    lg, lc = hps.LiveRows.pair(geo_t, col_t, live_idx)
    with torch.no_grad():      # the optimiser moved the live rows since the table was last refreshed
        geo_t[live_idx] += 1.0
        col_t[live_idx] -= 1.0
    gf = {l: (lg if l == level else cl[l]["geo_feats"].to(DEV)) for l in cl}
    cf = {l: (lc if l == level else cl[l]["col_feats"].to(DEV)) for l in cl}
    dec.zero_grad()
    rq = {l: cl[l]["radius"].to(DEV).clone() for l in cl}
    gt_d, gt_c = rays["gt_depth"].to(DEV), rays["gt_color"].to(DEV)
    depth, var, color, valid = renderer.render_batch_ray(
        npc, dec, rays["rays_d"].to(DEV), rays["rays_o"].to(DEV), DEV, stage, gt_depth=gt_d, npc_geo_feats=gf,
        npc_col_feats=cf, dynamic_r_query=rq)
    affine = dec.color_decoder_fine.mlp_exposure(scene["exposure_feat"].to(DEV)) if stage.startswith("color") else None
    hps.mapper_loss(depth, color, valid, gt_d, gt_c, stage, cfg["w_color_mapper"], affine).backward()
    assert torch.equal(depth, ref["depth"]) and torch.equal(color, ref["color"])
    assert lg.param.grad.shape == (live_idx.numel(), 32)
    assert parity.rel(lg.param.grad, ref["g_geo"][live_idx]) < 1e-5
    if stage.startswith("color"):
        assert parity.rel(lc.param.grad, ref["g_col"][live_idx]) < 1e-5
    assert torch.equal(geo_t[live_idx], lg.param.detach())            # table refreshed from the leaves


def test_render_batch_does_not_need_dict_copy(built):
    """Renderer.py:226-229 mutates dynamic_r_query in place; the drop-in keeps that visible behaviour."""
    scene, hps, npc, dec, renderer = built
    cl, rays = scene["clouds"], scene["rays"]
    rq = {l: cl[l]["radius"].to(DEV).clone() for l in cl}
    with torch.no_grad():
        renderer.render_batch_ray(npc, dec, rays["rays_d"].to(DEV), rays["rays_o"].to(DEV), DEV, "color_fine",
                                  gt_depth=rays["gt_depth"].to(DEV),
                                  npc_geo_feats={l: npc.get_geo_feats(l) for l in cl},
                                  npc_col_feats={l: npc.get_col_feats(l) for l in cl}, dynamic_r_query=rq)
    assert rq["fine"].shape == (rays["rays_d"].shape[0] * 5, 1)
    # opt-out (bench.py): the dict is left alone, results unchanged
    rq2 = {l: cl[l]["radius"].to(DEV).clone() for l in cl}
    renderer.write_back_radius = False
    try:
        with torch.no_grad():
            renderer.render_batch_ray(npc, dec, rays["rays_d"].to(DEV), rays["rays_o"].to(DEV), DEV, "color_fine",
                                      gt_depth=rays["gt_depth"].to(DEV),
                                      npc_geo_feats={l: npc.get_geo_feats(l) for l in cl},
                                      npc_col_feats={l: npc.get_col_feats(l) for l in cl}, dynamic_r_query=rq2)
    finally:
        renderer.write_back_radius = True
    assert rq2["fine"].shape == cl["fine"]["radius"].shape


def test_sampling_kernel_is_bit_identical_to_torch_ops(built):
    """hps_sample_rays reproduces the reference's torch expressions bit for bit (points, thresholds)."""
    scene, hps, npc, dec, renderer = built
    import ctypes
    g = torch.Generator().manual_seed(0)
    R, S = 1000, 5
    ro = (torch.randn(R, 3, generator=g) * 0.1).to(DEV)
    rd = torch.randn(R, 3, generator=g).to(DEV)
    gd = (torch.rand(R, generator=g) * 4 + 0.3).to(DEV)
    gd[::17] = 0.0
    rad = (0.04 + 0.12 * torch.rand(R, generator=g, dtype=torch.float64)).to(DEV)
    rad[::3] = 0.25
    zz = renderer._z_for_zero_depth(gd, DEV)
    z = torch.empty(R, S, device=DEV); pts = torch.empty(R * S, 3, device=DEV)
    t_up = torch.empty(R * S, device=DEV); t_dn = torch.empty(R * S, device=DEV)
    L = hps._lib
    hps.ops._call("hps_sample_rays", L.ptr(ro), L.ptr(rd), L.ptr(gd), L.ptr(rad), 0.0, L.ptr(zz), 0, R, S, 0.96, 1.04,
                  L.ptr(z), L.ptr(pts), L.ptr(t_up), L.ptr(t_dn), hps.ops._stream())
    z_ref, nz = renderer.surface_samples(gd, DEV)
    z_ref = torch.where(nz[:, None], z_ref, zz[None, :].expand_as(z_ref))
    pts_ref = (ro[..., None, :] + rd[..., None, :] * z_ref[..., :, None]).reshape(-1, 3)
    up_ref, dn_ref = hps.ops.radius_thresholds(rad.repeat_interleave(S), R * S, torch.device(DEV))
    assert torch.equal(z, z_ref) and torch.equal(pts, pts_ref)
    assert torch.equal(t_up, up_ref) and torch.equal(t_dn, dn_ref)
    zc, _ = O.surface_z_vals(gd.cpu())
    assert torch.equal(z.cpu(), zc)


def test_fused_losses_match_torch_expressions(built):
    scene, hps, npc, dec, renderer = built
    g = torch.Generator().manual_seed(1)
    R = 777
    depth = (torch.rand(R, generator=g) * 3 + 0.5).to(DEV)
    var = (torch.rand(R, generator=g) * 0.01).to(DEV)
    color = torch.randn(R, 3, generator=g).to(DEV)
    gt_d = (depth.cpu() + torch.randn(R, generator=g) * 0.05).to(DEV)
    gt_d[::11] = 0.0
    depth[5] = float("nan")
    gt_c = torch.rand(R, 3, generator=g).to(DEV)
    valid = (torch.rand(R, generator=g) > 0.1).to(DEV)
    aff = (torch.tensor([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0.0]) + 0.1 * torch.randn(12, generator=g)).to(DEV)
    for stage, affine in (("geometry_fine", None), ("color_fine", None), ("color_mid", aff)):
        outs = []
        for fn in (hps.mapper_loss, hps.mapper_loss_torch):
            d, c = depth.clone().requires_grad_(True), color.clone().requires_grad_(True)
            a = affine.clone().requires_grad_(True) if affine is not None else None
            loss = fn(d, c, valid, gt_d, gt_c, stage, 0.1, a)
            (loss * 1.7).backward()
            outs.append((loss, d.grad, c.grad if stage.startswith("color") else None, a.grad if a is not None else None))
        assert parity.rel(outs[0][0], outs[1][0]) < 1e-5
        assert parity.rel(torch.nan_to_num(outs[0][1]), torch.nan_to_num(outs[1][1])) < 1e-6
        for i in (2, 3):
            if outs[1][i] is not None:
                assert parity.rel(outs[0][i], outs[1][i]) < 1e-5
    depth[5] = 1.0
    outs = []
    for fn in (hps.tracker_loss, hps.tracker_loss_torch):
        d, c = depth.clone().requires_grad_(True), color.clone().requires_grad_(True)
        loss = fn(d, var, c, gt_d, gt_c, 0.5)
        loss.backward()
        outs.append((loss, d.grad, c.grad))
    assert parity.rel(outs[0][0], outs[1][0]) < 1e-5
    assert parity.rel(outs[0][1], outs[1][1]) < 1e-5 and parity.rel(outs[0][2], outs[1][2]) < 1e-6


def test_render_img_matches_batch_path(built):
    scene, hps, npc, dec, renderer = built
    cl = scene["clouds"]
    K = O.SCANNET_INTRINSICS
    import types
    small = hps.Renderer(hps.scannet_cfg(DEV), None, types.SimpleNamespace(H=24, W=32, fx=K["fx"] / 20, fy=K["fy"] / 20,
                                                                         cx=K["cx"] / 20, cy=K["cy"] / 20),
                         ray_batch_size=100)
    small.sigmoid_coefficient = 0.1
    small.min_render_chunk = 1            # first pass: the reference's chunking (8 chunks of 100 rays)
    depth_img = O.synthetic_depth()[::20, ::20].contiguous().to(DEV)
    rq = {l: torch.full((24, 32), 0.16 if l == "fine" else 0.96, dtype=torch.float64, device=DEV) for l in cl}
    c2w = torch.eye(4, device=DEV)[:3]
    d, u, c = small.render_img(npc, dec, c2w, DEV, "color_fine", gt_depth=depth_img,
                               npc_geo_feats={l: npc.get_geo_feats(l) for l in cl},
                               npc_col_feats={l: npc.get_col_feats(l) for l in cl}, dynamic_r_query=rq)
    assert d.shape == (24, 32) and c.shape == (24, 32, 3) and d.dtype == torch.float64
    small.min_render_chunk = 65536        # second pass: the default widening -> one chunk
    rq = {l: torch.full((24, 32), 0.16 if l == "fine" else 0.96, dtype=torch.float64, device=DEV) for l in cl}
    d2, u2, c2 = small.render_img(npc, dec, c2w, DEV, "color_fine", gt_depth=depth_img,
                                  npc_geo_feats={l: npc.get_geo_feats(l) for l in cl},
                                  npc_col_feats={l: npc.get_col_feats(l) for l in cl}, dynamic_r_query=rq)
    assert torch.equal(d, d2) and torch.equal(c, c2)       # chunking does not change per-ray results


def test_add_neural_points_density_test(built):
    scene, hps, npc0, dec, renderer = built
    cfg = hps.scannet_cfg(DEV)
    npc = hps.NeuralPointCloud(cfg)
    sc = O.synthetic_scene(8, 600, seed=4)
    ro, rd, gd, gc = (sc[k].to(DEV) for k in ("rays_o", "rays_d", "gt_depth", "gt_color"))
    gd[7] = 0.0
    r_add = torch.full((600,), 0.05, dtype=torch.float64, device=DEV)
    n1 = npc.add_neural_points(ro[:300], rd[:300], gd[:300], gc[:300], dynamic_radius=r_add[:300], level="fine", idx=0)
    assert int(n1) == 299 and npc.index_ntotal("fine") == 299 * 3 and npc.get_geo_feats("fine").shape == (897, 32)
    n2 = npc.add_neural_points(ro, rd, gd, gc, dynamic_radius=r_add, level="fine", idx=1)
    # oracle: a location is accepted iff no existing point lies within r_add (neural_point.py:262-268)
    pts_gt = (sc["rays_o"] + sc["rays_d"] * sc["gt_depth"][:, None])
    keep = torch.ones(600, dtype=torch.bool)
    keep[7] = False
    existing = npc.cloud_pos_tensor("fine")[:897].cpu()
    D, _ = O.exhaustive_knn(existing, pts_gt[keep], 8)
    want = int((D[:, 0] >= torch.tensor(0.05, dtype=torch.float64) ** 2).sum())
    assert int(n2) == want
    assert npc.index_ntotal("fine") == 897 + 3 * want
    import numpy as np
    assert np.array(npc.cloud_pos("fine")).shape == (897 + 3 * want, 3)


@pytest.mark.parametrize("tracker", [False, True])
def test_prepared_head_gives_identical_iteration(built, tracker):
    """Renderer.prepare_batch (sampling + neighbour search issued ahead of time) + render_batch_ray(prepared=...) is
    bit-identical to the one-call path in the forward and equal up to summation order in the gradients (mapper and
    tracker pose path), and the RayPrep can be refilled in place."""
    scene, hps, npc, dec, renderer = built
    cl, rays = scene["clouds"], scene["rays"]
    stage = "color_fine"

    def run(prepared_box):
        gf = {l: npc.get_geo_feats(l).detach().clone().requires_grad_(True) for l in cl}
        cf = {l: npc.get_col_feats(l).detach().clone().requires_grad_(True) for l in cl}
        ro = rays["rays_o"].to(DEV).clone().requires_grad_(tracker)
        rd = rays["rays_d"].to(DEV).clone().requires_grad_(tracker)
        gd = rays["gt_depth"].to(DEV)
        rq = {l: cl[l]["radius"].to(DEV).clone() for l in cl}
        prep = None
        if prepared_box is not None:
            prepared_box[0] = renderer.prepare_batch(npc, rd.detach(), ro.detach(), stage, gd, rq, out=prepared_box[0])
            prep = prepared_box[0]
        pos = {l: npc.cloud_pos_tensor(l) for l in cl}
        d, u, c, v = renderer.render_batch_ray(npc, dec, rd, ro, DEV, stage, gt_depth=gd, npc_geo_feats=gf,
                                               npc_col_feats=cf, is_tracker=tracker, cloud_pos=pos, dynamic_r_query=rq,
                                               prepared=prep)
        (d.sum() + c.sum()).backward()
        out = [d.detach(), u.detach(), c.detach(), v, gf["fine"].grad, cf["fine"].grad]
        if tracker:
            out += [ro.grad, rd.grad]
        return out

    for p in dec.parameters():
        p.grad = None
    ref = run(None)
    box = [None]
    for _ in range(2):          # second round refills the same RayPrep
        for p in dec.parameters():
            p.grad = None
        got = run(box)
        for a, b in zip(ref[:4], got[:4]):                    # forward: bit-identical
            assert torch.equal(a, b)
        for a, b in zip(ref[4:], got[4:]):                    # gradients: atomics reorder the sums (fp32, 1e-5 of the max)
            assert float((a - b).abs().max()) <= 1e-5 * max(float(a.abs().max()), 1e-30)


def test_frustum_rows_matches_oracle(built):
    """NeuralPointCloud.frustum_rows (Mapper.py:194-250 on the device) vs the numpy restatement, which is itself pinned
    against the reference's method run with the real cv2 (tests/test_oracle_golden.py): identical row sets except for
    points that sit within 1e-3 of one of the selection thresholds (float64 summation order / nothing else)."""
    import numpy as np
    scene, hps, npc, dec, renderer = built
    K = O.SCANNET_INTRINSICS
    depth = O.synthetic_depth()
    depth[::7, ::5] = 0.0                                   # zero-depth pixels take the maximum (Mapper.py:243-244)
    cam = torch.tensor([0.999, 0.02, -0.03, 0.01, 0.05, -0.02, 0.03])
    c2w = torch.cat([O.camera_from_tensor(cam), torch.tensor([[0.0, 0.0, 0.0, 1.0]])], 0)
    for level in ("mid", "fine"):
        pts = npc.cloud_pos_tensor(level).cpu()
        ref, margin = O.frustum_mask(pts.numpy(), c2w.numpy(), depth.numpy(), K["fx"], K["fy"], K["cx"], K["cy"], 20)
        got = npc.frustum_rows(level, c2w.to(DEV), depth.to(DEV), K["H"], K["W"], K["fx"], K["fy"], K["cx"], K["cy"], 20)
        got = got.cpu().numpy()
        assert 0 < len(ref) < len(pts)
        diff = np.setxor1d(ref, got)
        assert (margin[diff] < 1e-3).all(), "rows differ away from the thresholds"
        assert len(diff) <= 1e-3 * len(ref)


@pytest.mark.parametrize("optimiser", ["rows+torch", "mapper"])
def test_mapper_loop_five_iterations_tracks_the_oracle(hps, optimiser):
    """Five mapper iterations end to end (Mapper.py:783-874): frustum-selected live rows + RowAdam on the rows, Adam on
    the two fine decoders, vs the CPU oracle driven by torch.optim.Adam on dense tables.  The loss trajectory and the
    final feature rows must agree (fp32 rounding compounds over the steps: 2e-3 on the rows, 1e-3 on the losses)."""
    scene = parity.load("scene_small.pt")
    hpsm, npc, dec, renderer = parity.product(scene, DEV)
    cl, rays, cfg = scene["clouds"], scene["rays"], scene["cfg"]
    stage, level = "color_fine", "fine"
    K = O.SCANNET_INTRINSICS
    lr_rows, lr_dec, n_it = 5e-3, 1e-3, 5

    # ---- product
    depth_img = O.synthetic_depth()
    c2w = torch.eye(4)
    idx = npc.frustum_rows(level, c2w.to(DEV), depth_img.to(DEV), K["H"], K["W"], K["fx"], K["fy"], K["cx"], K["cy"], 0)
    assert idx.numel() > 100
    lg, lc = hpsm.LiveRows.pair(npc.get_geo_feats(level), npc.get_col_feats(level), idx)
    gf = {l: (lg if l == level else npc.get_geo_feats(l)) for l in cl}
    cf = {l: (lc if l == level else npc.get_col_feats(l)) for l in cl}
    dparams = list(dec.geo_decoder_fine.parameters()) + list(dec.color_decoder_fine.parameters())
    if optimiser == "mapper":      # the whole optimiser on this package's kernels (decoders through FusedAdam)
        opt_rows = hpsm.MapperOptimizer(hpsm.scannet_cfg(DEV)["mapping"], decoder_params=dparams,
                                        rows={"geometry_fine": lg, "color_fine": lc})
        opt_rows.set_learning_rates({"decoders": lr_dec, "geometry_mid": 0.0, "geometry_fine": lr_rows,
                                     "color_mid": 0.0, "color_fine": lr_rows})
        import types
        opt_dec = types.SimpleNamespace(zero_grad=lambda: None, step=lambda: None)     # opt_rows does both
    else:
        opt_rows = hpsm.RowAdam([(lg, lr_rows), (lc, lr_rows)])
        opt_dec = torch.optim.Adam(dparams, lr=lr_dec)
    ef = scene["exposure_feat"].to(DEV)
    gt_d, gt_c = rays["gt_depth"].to(DEV), rays["gt_color"].to(DEV)
    losses = []
    for _ in range(n_it):
        opt_rows.zero_grad()
        opt_dec.zero_grad()
        d, v, c, m = renderer.render_batch_ray(npc, dec, rays["rays_d"].to(DEV), rays["rays_o"].to(DEV), DEV, stage,
                                               gt_depth=gt_d, npc_geo_feats=gf, npc_col_feats=cf,
                                               dynamic_r_query={l: cl[l]["radius"].to(DEV).clone() for l in cl})
        aff = dec.color_decoder_fine.mlp_exposure(ef)
        loss = hpsm.mapper_loss(d, c, m, gt_d, gt_c, stage, cfg["w_color_mapper"], aff)
        loss.backward()
        opt_rows.step()
        opt_dec.step()
        losses.append(float(loss.detach()))

    # ---- oracle: dense tables, only the selected rows receive the Adam update (masked gradient, Mapper.py:862-874)
    odec = {k: {kk: vv.clone().requires_grad_(kk != "embedder._B" or k.startswith("geo")) for kk, vv in v.items()}
            for k, v in scene["decoders"].items()}
    ogf = {l: cl[l]["geo_feats"].clone() for l in cl}
    ocf = {l: cl[l]["col_feats"].clone() for l in cl}
    icpu = idx.cpu()
    rg = ogf[level][icpu].clone().requires_grad_(True)
    rc = ocf[level][icpu].clone().requires_grad_(True)
    oparams = [t for n in ("geo_decoder_fine", "color_decoder_fine") for t in odec[n].values() if t.requires_grad]
    o_rows = torch.optim.Adam([rg, rc], lr=lr_rows)
    o_dec = torch.optim.Adam(oparams, lr=lr_dec)
    olosses = []
    for _ in range(n_it):
        o_rows.zero_grad()
        o_dec.zero_grad()
        tg = {l: (ogf[l].index_put((icpu,), rg) if l == level else ogf[l]) for l in cl}
        tc = {l: (ocf[l].index_put((icpu,), rc) if l == level else ocf[l]) for l in cl}
        d, v, c, m = O.render_batch_ray({l: cl[l]["points"] for l in cl}, odec, rays["rays_d"], rays["rays_o"], stage,
                                        rays["gt_depth"], tg, tc, {l: cl[l]["radius"] for l in cl},
                                        scene["c_fill_geo"], scene["c_fill_col"])
        aff = O.exposure_affine(odec["color_decoder_fine"], scene["exposure_feat"])
        loss = O.mapper_loss(d, c, m, rays["gt_depth"], rays["gt_color"], stage, cfg["w_color_mapper"], aff)
        loss.backward()
        o_rows.step()
        o_dec.step()
        olosses.append(float(loss.detach()))

    assert olosses[-1] < olosses[0]                      # the loop optimises something
    for a, b in zip(losses, olosses):
        assert abs(a - b) <= 1e-3 * abs(b), (losses, olosses)
    assert parity.rel(lg.param, rg) < 2e-3 and parity.rel(lc.param, rc) < 2e-3
    assert torch.equal(npc.get_geo_feats(level)[idx], lg.param.detach())     # the table carries the optimised rows


def test_tracker_loop_pose_trajectory_tracks_the_oracle(hps):
    """Ten tracker iterations (Tracker.py:190-235: pose 7-vector + exposure latent under Adam, decoders frozen), the
    product's whole iteration replayed from ONE CUDA graph, vs the CPU oracle in float64 (the truth the fp32 reference is
    itself 1e-4..1e-3 away from on pose gradients)."""
    scene = parity.load("scene_small.pt")
    fx = parity.load("render_tracker_small.pt")["inputs"]
    hpsm, npc, dec, renderer = parity.product(scene, DEV)
    cl, cfg = scene["clouds"], scene["cfg"]
    K = O.SCANNET_INTRINSICS
    stage, n_it, lr = "color_fine", 10, 5e-4
    for prm in dec.parameters():
        prm.requires_grad_(False)

    # ---- product: graph-replayed iteration
    cam = fx["cam"].to(DEV).clone().requires_grad_(True)
    ef = scene["exposure_feat"].to(DEV).clone().requires_grad_(True)
    opt = torch.optim.Adam([cam, ef], lr=lr, capturable=True)
    cam.grad, ef.grad = torch.zeros_like(cam), torch.zeros_like(ef)
    pi, pj = fx["i"].to(DEV).float(), fx["j"].to(DEV).float()
    gt_d, gt_c = fx["gt_depth"].to(DEV), fx["gt_color"].to(DEV)
    pos = {l: npc.cloud_pos_tensor(l) for l in cl}
    gf = {l: npc.get_geo_feats(l) for l in cl}
    cf = {l: npc.get_col_feats(l) for l in cl}
    rq = {l: cl[l]["radius"].to(DEV) for l in cl}

    def it():
        opt.zero_grad(set_to_none=False)
        c2w = hpsm.get_camera_from_tensor(cam)
        ro, rd = hpsm.get_rays_from_uv(pi, pj, c2w, K["H"], K["W"], K["fx"], K["fy"], K["cx"], K["cy"], DEV)
        d, v, c, _ = renderer.render_batch_ray(npc, dec, rd, ro, DEV, stage, gt_depth=gt_d, npc_geo_feats=gf,
                                               npc_col_feats=cf, is_tracker=True, cloud_pos=pos,
                                               dynamic_r_query=dict(rq), exposure_feat=ef)
        loss = hpsm.tracker_loss(d, v, c, gt_d, gt_c, cfg["w_color_tracker"])
        loss.backward()
        opt.step()
        return loss

    cam0, ef0 = cam.detach().clone(), ef.detach().clone()
    graphed = hpsm.graphs.GraphedStep(it, {}, warmup=3)          # 3 eager warm-up iterations + the captured one ...
    with torch.no_grad():                                        # ... then restart from the initial state
        cam.copy_(cam0)
        ef.copy_(ef0)
        for st in opt.state.values():
            for k2, v2 in st.items():
                v2.zero_()
    traj = []
    for _ in range(n_it):
        graphed()
        traj.append(cam.detach().cpu().clone())

    # ---- oracle: float64 = truth, float32 = the reference's own arithmetic (yardstick for the fp32 noise floor)
    def oracle_traj(dtype):
        cv = lambda t: t.to(dtype) if t.is_floating_point() else t
        ocam = fx["cam"].to(dtype).clone().requires_grad_(True)
        oef = scene["exposure_feat"].to(dtype).clone().requires_grad_(True)
        oopt = torch.optim.Adam([ocam, oef], lr=lr)
        odec = {k: {kk: cv(vv) for kk, vv in v.items()} for k, v in scene["decoders"].items()}
        opos = {l: cv(cl[l]["points"]) for l in cl}
        out = []
        for _ in range(n_it):
            oopt.zero_grad()
            ro, rd = O.rays_from_uv(cv(fx["i"]), cv(fx["j"]), O.camera_from_tensor(ocam), K["fx"], K["fy"], K["cx"], K["cy"])
            d, v, c, _ = O.render_batch_ray(opos, odec, rd, ro, stage, cv(fx["gt_depth"]),
                                            {l: cv(cl[l]["geo_feats"]) for l in cl}, {l: cv(cl[l]["col_feats"]) for l in cl},
                                            {l: cl[l]["radius"] for l in cl}, cv(scene["c_fill_geo"]), cv(scene["c_fill_col"]),
                                            is_tracker=True, exposure_feat=oef)
            loss = O.tracker_loss(d, v, c, cv(fx["gt_depth"]), cv(fx["gt_color"]), cfg["w_color_tracker"])
            loss.backward()
            oopt.step()
            out.append(ocam.detach().double().clone())
        return out

    t64, t32 = oracle_traj(torch.float64), oracle_traj(torch.float32)
    moved = float((t64[-1] - fx["cam"].double()).abs().max())
    assert moved > 1e-3                                           # the pose did move
    dev_ours = [float((a.double() - b).abs().max()) for a, b in zip(traj, t64)]
    dev_ref = [float((a - b).abs().max()) for a, b in zip(t32, t64)]
    # The loss is piecewise smooth in the pose (neighbour sets and ReLU kinks switch as the 96 rays move), so two fp32
    # evaluations separate exponentially once one of them crosses a switch the other has not: the fp32 reference is
    # 7e-5 from the float64 truth after ten steps, the kernels 3e-4 (both doubling per step from ~iteration 3).  What
    # can be asserted: the first steps coincide (no systematic error) and the run stays within one Adam step (lr).
    assert max(dev_ours[:3]) <= 1e-5, (dev_ours, dev_ref)
    assert max(dev_ours) <= 2 * lr, (dev_ours, dev_ref)


def test_pixel_sampling_on_device(hps, golden):
    """get_samples / get_selected_index_with_grad / get_tensor_from_camera with CUDA inputs: nothing leaves the device,
    the rays are bit-identical to the reference expressions evaluated on the same pixels, and the fixture values of
    oracle/make_golden_sampling.py are reproduced where no RNG is involved."""
    g = golden("sampling_small.pt")
    fx, fy, cx, cy = g["intr"]
    H, W = g["H"], g["W"]
    depth, color, c2w = g["depth"].to(DEV), g["color"].to(DEV), g["c2w"].to(DEV)
    torch.manual_seed(0)
    ro, rd, sd, sc, i, j = hps.common.get_samples(4, H - 4, 6, W - 6, 500, H, W, fx, fy, cx, cy, c2w, depth, color, DEV,
                                                  depth_filter=True, return_index=True, depth_limit=5.0)
    for t in (ro, rd, sd, sc, i, j):
        assert t.is_cuda
    assert 0 < sd.shape[0] < 500 and bool(((sd > 0) & (sd < 5.0)).all())
    assert int(i.min()) >= 6 and int(i.max()) < W - 6 and int(j.min()) >= 4 and int(j.max()) < H - 4
    assert torch.equal(sd, depth[j, i]) and torch.equal(sc, color[j, i])
    fi, fj = i.float(), j.float()
    dirs = torch.stack([(fi - cx) / fx, -(fj - cy) / fy, -torch.ones_like(fi)], -1).reshape(-1, 1, 3)
    assert torch.equal(rd, torch.sum(dirs * c2w[:3, :3], -1))                    # common.py:112-117
    assert torch.equal(ro, c2w[:3, -1].expand(rd.shape))

    sel, mag = hps.common.get_selected_index_with_grad(4, H - 4, 6, W - 6, 20, color, gt_depth=depth, depth_limit=True,
                                                       as_numpy=False)
    assert sel.is_cuda and mag.is_cuda
    sel_c, mag_c = hps.common.get_selected_index_with_grad(4, H - 4, 6, W - 6, 20, color.cpu(), gt_depth=depth.cpu(),
                                                           depth_limit=True, as_numpy=False)
    assert torch.allclose(mag.cpu(), mag_c, rtol=0, atol=1e-14)
    assert set(sel.tolist()) == set(sel_c.tolist())

    for RT, (q, tq, _) in zip(g["mats"], g["quats"]):
        mine = hps.common.get_tensor_from_camera(RT.to(DEV))
        assert mine.is_cuda and torch.allclose(mine.cpu(), q, rtol=0, atol=2e-7)
        assert torch.allclose(hps.common.get_tensor_from_camera(RT.to(DEV), Tquad=True).cpu(), tq, rtol=0, atol=2e-7)


def test_sample_near_pcl_matches_reference(built, golden):
    """NeuralPointCloud.sample_near_pcl against the reference's own function (neural_point.py:443-495; fixture:
    oracle/make_golden_near_pcl.py): the not-near-the-cloud mask and the float32 sample depths, bit for bit, for the
    fixed query radius and for a radius per probe."""
    scene, hps, npc, dec, renderer = built
    g = golden("near_pcl_small.pt")
    none = ~(g["gt_depth"] > 0)
    ro, rd = scene["rays"]["rays_o"][none].to(DEV), g["rays_d"][none].to(DEV)
    keep = npc.radius_query
    try:
        for stage, s in g["stages"].items():
            level = "mid" if stage.endswith("mid") else "fine"
            npc.radius_query = g["radius_query"][level]
            z, invalid = npc.sample_near_pcl(ro, rd, s["near"], s["far"].to(DEV), 5, None, level)
            assert z.dtype == torch.float32 and z.shape == (int(none.sum()), 5)
            assert torch.equal(invalid.cpu(), s["invalid"]), stage
            assert torch.equal(z.cpu(), s["z"]), (stage, float((z.cpu() - s["z"]).abs().max()))
            z, invalid = npc.sample_near_pcl(ro, rd, s["near"], s["far"].to(DEV), 5, s["r_probe"].to(DEV), level)
            assert torch.equal(invalid.cpu(), s["invalid_dyn"]) and torch.equal(z.cpu(), s["z_dyn"]), stage
    finally:
        npc.radius_query = keep
    assert 0 < int(g["stages"]["color_fine"]["invalid"].sum()) < int(none.sum())      # the fixture has both kinds


@pytest.mark.parametrize("dynamic", [False, True])
def test_sample_near_pcl_render_mode(built, golden, dynamic):
    """rendering.sample_near_pcl = True (Renderer.py:196-206,247-251): depth-less rays are sampled near the cloud (one
    z row per ray through hps_sample_rays' z_zero_stride), keep their rendered depth, and are masked out when the
    cloud is not near.  The reference's call site cannot run (oracle/make_golden_near_pcl.py); the fused path is
    checked against the modular path and against sample_near_pcl's own output."""
    import types
    scene, hps, npc, dec, _ = built
    g = golden("near_pcl_small.pt")
    cfg = hps.scannet_cfg(DEV)
    cfg["rendering"]["sample_near_pcl"] = True
    cfg["use_dynamic_radius"] = dynamic
    rend = hps.Renderer(cfg, None, types.SimpleNamespace(**O.SCANNET_INTRINSICS))
    rend.sigmoid_coefficient = scene["cfg"]["coef"]
    gt = g["gt_depth"].to(DEV)
    ro, rd = scene["rays"]["rays_o"].to(DEV), g["rays_d"].to(DEV)
    cl = scene["clouds"]
    gf = {l: cl[l]["geo_feats"].to(DEV) for l in cl}
    cf = {l: cl[l]["col_feats"].to(DEV) for l in cl}
    keep = npc.radius_query
    try:
        for stage in ("geometry_mid", "color_fine"):
            level = "mid" if stage.endswith("mid") else "fine"
            npc.radius_query = g["radius_query"][level]
            rq = lambda: {l: cl[l]["radius"].to(DEV).clone() for l in cl} if dynamic else None
            with torch.no_grad():
                d1, u1, c1, v1 = rend.render_batch_ray(npc, dec, rd, ro, DEV, stage, gt_depth=gt, npc_geo_feats=gf,
                                                       npc_col_feats=cf, dynamic_r_query=rq())
                d2, u2, c2, v2 = rend.render_batch_ray_modular(npc, dec, rd, ro, DEV, stage, gt_depth=gt,
                                                               npc_geo_feats=gf, npc_col_feats=cf,
                                                               dynamic_r_query=rq())
            assert torch.equal(v1, v2)
            for a, b in ((d1, d2), (u1, u2), (c1, c2)):
                assert parity.rel(a, b) <= 1e-5
            none = ~(gt > 0)
            radius = cl[level]["radius"].to(DEV).reshape(-1)[none] if dynamic else None
            z0, not_near = npc.sample_near_pcl(ro[none], rd[none], rend.near_end, rend._far_end(gt), 5, radius, level)
            idx = torch.nonzero(none).squeeze(1)
            assert not bool(v1[idx[not_near]].any())                              # not near the cloud: never valid
            near_valid = v1[idx[~not_near]]
            if not dynamic:
                assert torch.equal(not_near.cpu(), g["stages"][stage]["invalid"])
                assert torch.equal(z0.cpu(), g["stages"][stage]["z"])
            if bool(near_valid.any()):
                dd = d1[idx[~not_near]][near_valid]
                zz = z0[~not_near][near_valid]
                assert bool((dd > 0).all())                                       # depth kept (Renderer.py:250)
                assert bool((dd <= zz[:, -1] + 1e-4).all())                       # sum of w*z with sum(w) <= 1
    finally:
        npc.radius_query = keep


def test_render_batch_ray_tracker_loss_equals_render_plus_loss(built):
    """renderer.render_batch_ray_tracker_loss (batch statistic accumulated by the compositing kernel, loss kernel,
    upstream gradient as a device scalar) against render_batch_ray + tracker_loss, with some rays stripped of their
    depth reading (they count with depth 0 in the statistic, Renderer.py:250-251) and an upstream gradient of 0.5."""
    scene, hps, npc, dec, renderer = built
    ref = parity.load("render_tracker_small.pt")
    inp = dict(ref["inputs"])
    inp["gt_depth"] = inp["gt_depth"].clone()
    inp["gt_depth"][::7] = 0.0
    K = O.SCANNET_INTRINSICS
    cl = scene["clouds"]
    outs = []
    for fused in (False, True):
        cam = inp["cam"].to(DEV).clone().requires_grad_(True)
        ef = scene["exposure_feat"].to(DEV).clone().requires_grad_(True)
        dec.zero_grad()
        c2w = hps.get_camera_from_tensor(cam)
        ro, rd = hps.get_rays_from_uv(inp["i"].to(DEV), inp["j"].to(DEV), c2w, K["H"], K["W"], K["fx"], K["fy"],
                                      K["cx"], K["cy"], DEV)
        gt_d, gt_c = inp["gt_depth"].to(DEV), inp["gt_color"].to(DEV)
        kw = dict(npc_geo_feats={l: npc.get_geo_feats(l) for l in cl}, npc_col_feats={l: npc.get_col_feats(l) for l in cl},
                  cloud_pos={l: npc.cloud_pos_tensor(l) for l in cl},
                  dynamic_r_query={l: cl[l]["radius"].to(DEV).clone() for l in cl}, exposure_feat=ef)
        if fused:
            loss, d, v, c, m = renderer.render_batch_ray_tracker_loss(npc, dec, rd, ro, DEV, "color_fine", gt_d, gt_c,
                                                                      0.5, **kw)
        else:
            d, v, c, m = renderer.render_batch_ray(npc, dec, rd, ro, DEV, "color_fine", gt_depth=gt_d, is_tracker=True,
                                                   **kw)
            loss = hps.tracker_loss(d, v, c, gt_d, gt_c, 0.5)
        (0.5 * loss).backward()
        outs.append(dict(loss=loss.detach(), d=d.detach(), v=v, c=c.detach(), m=m, g_cam=cam.grad, g_ef=ef.grad))
    a, b = outs
    assert torch.equal(a["d"], b["d"]) and torch.equal(a["v"], b["v"]) and torch.equal(a["c"], b["c"])
    assert torch.equal(a["m"], b["m"])
    assert parity.rel(b["loss"], a["loss"]) <= 1e-6
    assert parity.rel(b["g_cam"], a["g_cam"]) <= 1e-4 and parity.rel(b["g_ef"], a["g_ef"]) <= 1e-4


def test_render_batch_ray_loss_with_several_keyframes(built):
    """render_batch_ray_loss(affine=(F,12), frame_index=...) - a batch mixing rays of F keyframes, Mapper.py:805-822 -
    against render_batch_ray + mapper_loss with the same arguments; the affines come from the exposure MLP, so its
    parameters and the F latents receive their gradients through autograd."""
    scene, hps, npc, dec, renderer = built
    dev, stage, level, F = DEV, "color_fine", "fine", 3
    cl, rays = scene["clouds"], scene["rays"]
    R = rays["rays_o"].shape[0]
    frame = (torch.arange(R, device=dev) * F) // R
    gt_d, gt_c = rays["gt_depth"].to(dev), rays["gt_color"].to(dev)
    mlp = dec.color_decoder_fine.mlp_exposure
    outs = []
    for fused in (False, True):
        dec.zero_grad()
        gf = {l: cl[l]["geo_feats"].to(dev).clone().requires_grad_(True) for l in cl}
        cf = {l: cl[l]["col_feats"].to(dev).clone().requires_grad_(True) for l in cl}
        efs = [(scene["exposure_feat"].to(dev) * (1 + 0.3 * k)).clone().requires_grad_(True) for k in range(F)]
        affine = torch.stack([mlp(e) for e in efs])
        kw = dict(npc_geo_feats=gf, npc_col_feats=cf, cloud_pos={l: npc.cloud_pos_tensor(l) for l in cl},
                  dynamic_r_query={l: cl[l]["radius"].to(dev).clone() for l in cl})
        if fused:
            loss = renderer.render_batch_ray_loss(npc, dec, rays["rays_d"].to(dev), rays["rays_o"].to(dev), dev, stage,
                                                  gt_d, gt_c, 0.1, affine=affine, frame_index=frame, **kw)[0]
        else:
            d, v, c, m = renderer.render_batch_ray(npc, dec, rays["rays_d"].to(dev), rays["rays_o"].to(dev), dev, stage,
                                                   gt_depth=gt_d, is_tracker=False, exposure_feat=None, **kw)
            loss = hps.mapper_loss(d, c, m, gt_d, gt_c, stage, 0.1, affine, frame_index=frame)
        loss.backward()
        outs.append(dict(loss=loss.detach(), g_geo=gf[level].grad, g_col=cf[level].grad, g_ef=torch.stack([e.grad for e in efs]),
                         g_mlp={k: p.grad.clone() for k, p in mlp.named_parameters()},
                         g_dec={k: p.grad.clone() for k, p in dec.color_decoder_fine.named_parameters() if p.grad is not None}))
    a, b = outs
    assert parity.rel(b["loss"], a["loss"]) <= 1e-6
    assert float(a["g_ef"].abs().min(1).values.max()) > 0            # every frame's latent is reached
    for k in ("g_geo", "g_col", "g_ef"):
        assert parity.rel(b[k], a[k]) <= 2e-5, k
    assert max(parity.rel(b["g_mlp"][k], a["g_mlp"][k]) for k in a["g_mlp"]) <= 2e-5
    assert max(parity.rel(b["g_dec"][k], a["g_dec"][k]) for k in a["g_dec"]) <= 2e-5


@pytest.mark.parametrize("per_ray_z", [False, True])
def test_search_that_samples_on_the_fly_is_bit_identical(built, per_ray_z):
    """hps_sample_knn (the neighbour search computes its query points itself) against hps_sample_rays followed by
    hps_knn_query: sample depths, points, both thresholds, neighbour indices, distances and counts, bit for bit -
    with depth-less rays taking a shared z row or one row per ray."""
    scene, hps, npc, dec, renderer = built
    L, ops = hps._lib, hps.ops
    rays, cl = scene["rays"], scene["clouds"]["fine"]
    ro, rd = rays["rays_o"].to(DEV).contiguous(), rays["rays_d"].to(DEV).contiguous()
    gd = rays["gt_depth"].to(DEV).clone()
    gd[::4] = 0.0
    R, S = ro.shape[0], 5
    rad = cl["radius"].to(DEV).double().reshape(-1).contiguous()
    zz = (torch.rand(R, S, device=DEV) * 3 + 0.3).sort(1).values if per_ray_z else torch.linspace(0.3, 4.0, S, device=DEV)
    grid = npc.get_index("fine").grid_for(rad)
    f = lambda *s: torch.empty(*s, device=DEV)
    i32 = lambda *s: torch.empty(*s, device=DEV, dtype=torch.int32)
    a = dict(z=f(R, S), pts=f(R * S, 3), up=f(R * S), dn=f(R * S), I=i32(R * S, 8), D=f(R * S, 8), n=i32(R * S))
    b = {k: torch.empty_like(v) for k, v in a.items()}
    st = ops._stream()
    ops._call("hps_sample_rays", L.ptr(ro), L.ptr(rd), L.ptr(gd), L.ptr(rad), 0.0, L.ptr(zz), S if per_ray_z else 0, R, S,
              0.96, 1.04, L.ptr(a["z"]), L.ptr(a["pts"]), L.ptr(a["up"]), L.ptr(a["dn"]), st)
    ops._call("hps_knn_query", grid._h, L.ptr(a["pts"]), L.ptr(a["up"]), L.ptr(a["dn"]), R * S, L.ptr(a["I"]), L.ptr(a["D"]),
              L.ptr(a["n"]), None, 1, st)
    ops._call("hps_sample_knn", grid._h, L.ptr(ro), L.ptr(rd), L.ptr(gd), L.ptr(rad), 0.0, L.ptr(zz), S if per_ray_z else 0,
              R, S, 0.96, 1.04, L.ptr(b["z"]), L.ptr(b["pts"]), L.ptr(b["up"]), L.ptr(b["dn"]), L.ptr(b["I"]), L.ptr(b["D"]),
              L.ptr(b["n"]), None, st)
    torch.cuda.synchronize()
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert int(a["n"].sum()) > 0
    # scalar radius form
    ops._call("hps_sample_rays", L.ptr(ro), L.ptr(rd), L.ptr(gd), None, 0.08, None, 0, R, S, 0.96, 1.04, L.ptr(a["z"]),
              L.ptr(a["pts"]), L.ptr(a["up"]), L.ptr(a["dn"]), st)
    ops._call("hps_knn_query", grid._h, L.ptr(a["pts"]), L.ptr(a["up"]), L.ptr(a["dn"]), R * S, L.ptr(a["I"]), L.ptr(a["D"]),
              L.ptr(a["n"]), None, 1, st)
    ops._call("hps_sample_knn", grid._h, L.ptr(ro), L.ptr(rd), L.ptr(gd), None, 0.08, None, 0, R, S, 0.96, 1.04,
              L.ptr(b["z"]), L.ptr(b["pts"]), L.ptr(b["up"]), L.ptr(b["dn"]), L.ptr(b["I"]), L.ptr(b["D"]), L.ptr(b["n"]),
              None, st)
    torch.cuda.synchronize()
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_loss_kernels_match_the_reference_statements(hps, golden):
    """The fused loss kernels vs the fixture made by EXECUTING the reference's own loss statements (Mapper.py:793-856,
    Tracker.py:213-231; oracle/make_golden_losses.py): a mapper batch mixing rays of four keyframes with their own
    exposure affines (segments of 289 / 5 / ... rays), NaN depth, missing depth, invalid rays, all four stages; the
    tracker loss with outlier rays.  Loss and the gradients to depth, colour and every keyframe's affine."""
    fx = golden("losses_small.pt")
    ids = fx["frame_ids"][fx["inside_mask"]]
    frame = torch.unique_consecutive(ids, return_inverse=True)[1].to(DEV)        # keyframe slot 0..F-1 of every ray
    valid, gt_d, gt_c = fx["valid"].to(DEV), fx["gt_depth"].to(DEV), fx["gt_color"].to(DEV)
    for stage in ("geometry_mid", "geometry_fine", "color_mid", "color_fine"):
        want = fx["mapper"][(stage, True)]
        depth, color, aff = (fx[k].clone().to(DEV).requires_grad_(True) for k in ("depth", "color", "affine"))
        loss = hps.mapper_loss(depth, color, valid, gt_d, gt_c, stage, fx["w_color_mapper"], aff, frame_index=frame)
        loss.backward()
        assert parity.rel(loss.detach().cpu(), want["loss"]) < 1e-5, stage
        assert parity.rel(depth.grad.cpu(), want["g_depth"]) < 1e-6, stage
        if stage.startswith("color"):
            assert parity.rel(color.grad.cpu(), want["g_color"]) < 1e-5, stage
            assert parity.rel(aff.grad.cpu(), want["g_affine"]) < 1e-5, stage
    want = fx["tracker"][(True, True)]
    depth = fx["tracker_depth"].clone().to(DEV).requires_grad_(True)
    color = fx["tracker_color"].clone().to(DEV).requires_grad_(True)
    loss = hps.tracker_loss(depth, fx["uncertainty"].to(DEV), color, gt_d, gt_c, fx["w_color_tracker"])
    loss.backward()
    assert parity.rel(loss.detach().cpu(), want["loss"]) < 1e-5
    assert parity.rel(depth.grad.cpu(), want["g_depth"]) < 1e-5 and parity.rel(color.grad.cpu(), want["g_color"]) < 1e-6


def test_add_neural_points_matches_the_reference_class(hps, golden):
    """NeuralPointCloud.add_neural_points (density test + N_add points per accepted location, on the device) vs the
    fixture made by running the UNMODIFIED reference class on the CPU (oracle/make_golden_insertion.py): seven
    consecutive insertions into two levels - first frame, repeated + new rays under per-ray float64 radii, scalar
    radius_add with a missing-depth ray, radius_min (is_pts_grad), mid level incl. a frame that adds nothing.  After every
    call: the same count of accepted locations and the same cloud, point for point in the same order (<= 1e-6 m)."""
    import numpy as np
    fx = golden("insertion_small.pt")
    npc = hps.NeuralPointCloud(hps.scannet_cfg(DEV))
    assert npc.radius_add == fx["radius_add"] and npc.radius_min == fx["radius_min"]
    ro, rd, gc = fx["rays_o"].to(DEV), fx["rays_d"].to(DEV), fx["gt_color"].to(DEV)
    for c in fx["calls"]:
        a, b = c["rays"]
        depth = (fx["gt_depth_zero"] if c["zero"] else fx["gt_depth"])[a:b].to(DEV)
        rad = None
        if c["dynamic"]:
            rad = ((fx["r_fine"] if c["level"] == "fine" else fx["r_mid"])[a:b] * c.get("scale", 1.0)).to(DEV)
        n = npc.add_neural_points(ro[a:b], rd[a:b], depth, gc[a:b], train=False, is_pts_grad=c["is_pts_grad"],
                                  dynamic_radius=rad, level=c["level"], idx=c["idx"])
        tag = (c["level"], c["idx"])
        assert int(n) == c["count"], tag
        cloud = npc.cloud_pos_tensor(c["level"]).cpu()
        assert cloud.shape == c["cloud"].shape, tag
        # same fp32 expressions on both sides (expected bit-equal); the bar is 1e-6 m on positions of a few metres
        assert float((cloud - c["cloud"]).abs().max()) <= 1e-6, (tag, float((cloud - c["cloud"]).abs().max()))
        assert npc.index_ntotal(c["level"]) == c["cloud"].shape[0] == npc.pts_num()[c["level"]]
        assert npc.get_geo_feats(c["level"]).shape == (c["cloud"].shape[0], 32) == npc.get_col_feats(c["level"]).shape
        assert len(npc.input_pos()) == c["n_input_pos"]
    assert np.array(npc.cloud_pos("mid")).shape == (fx["calls"][-1]["cloud"].shape[0], 3)


@pytest.mark.parametrize("chunks", ["reference", "wide"])
def test_render_img_matches_the_reference_frame(built, golden, chunks):
    """Renderer.render_img vs the fixture made by running the reference's render_img on the CPU
    (oracle/make_golden_render_img.py): 24 x 32 frame, perturbed pose, float64 radius maps, exposure latent, 87 pixels
    without a depth reading.  'reference' = the reference's chunking (100 rays); 'wide' = the product's default widening,
    where the sampling range of the depth-less pixels still has to follow the reference's chunks (without
    Renderer._z_rows_for_zero_depth their colour is off by up to 0.6, tests/test_oracle_golden.py)."""
    import types
    scene, hps, npc, dec, renderer = built
    fx = golden("render_img_small.pt")
    cl = scene["clouds"]
    small = hps.Renderer(hps.scannet_cfg(DEV), None, types.SimpleNamespace(H=fx["H"], W=fx["W"], **fx["intrinsics"]),
                         ray_batch_size=fx["ray_batch_size"])
    small.sigmoid_coefficient = scene["cfg"]["coef"]
    small.min_render_chunk = 1 if chunks == "reference" else 65536
    missing = fx["depth"] == 0
    for stage, want in fx["stages"].items():
        d, u, c = small.render_img(npc, dec, fx["c2w"].to(DEV), DEV, stage, gt_depth=fx["depth"].to(DEV),
                                   npc_geo_feats={l: npc.get_geo_feats(l) for l in cl},
                                   npc_col_feats={l: npc.get_col_feats(l) for l in cl},
                                   dynamic_r_query={l: fx["radius"][l].clone().to(DEV) for l in cl},
                                   cloud_pos={l: npc.cloud_pos_tensor(l) for l in cl},
                                   exposure_feat=scene["exposure_feat"].to(DEV))
        assert d.dtype == torch.float64 and u.dtype == torch.float64 and d.shape == (fx["H"], fx["W"])
        d, u, c = d.cpu(), u.cpu(), c.cpu()
        assert bool((d[missing] == 0).all())
        assert parity.rel(d, want["depth"]) < parity.TOL, (stage, parity.rel(d, want["depth"]))
        assert parity.rel(u, want["uncertainty"]) < parity.TOL, (stage, parity.rel(u, want["uncertainty"]))
        assert float((c - want["color"]).abs().max()) < parity.TOL, (stage, float((c - want["color"]).abs().max()))


@pytest.mark.parametrize("separate_LR", [False, True])
def test_tracker_optimizer_on_the_device(hps, separate_LR):
    """TrackerOptimizer (Tracker.py:383-415 as one FusedAdam launch + the keep-the-best-pose rule :435-437 on the device)
    vs torch.optim.Adam over the same groups and the rule spelled out on the host, 14 iterations of a non-monotone loss.
    (Its CPU twin, tests/test_host_logic.py, builds the reference optimiser by executing the reference's source.)"""
    import copy
    import math
    cfg = hps.scannet_cfg(DEV)
    tracking = dict(cfg["tracking"], separate_LR=separate_LR, iters=14)
    torch.manual_seed(1219)
    dec_mine = hps.get_model(cfg).to(DEV)
    dec_ref = copy.deepcopy(dec_mine)
    g = torch.Generator().manual_seed(3)
    cam0 = (torch.tensor([0.99, 0.05, -0.08, 0.03, 0.4, -0.2, 0.1]) + 0.01 * torch.randn(7, generator=g)).to(DEV)
    ef_mine = (0.3 * torch.randn(8, generator=g)).to(DEV).requires_grad_(True)
    ef_ref = ef_mine.detach().clone().requires_grad_(True)
    target = cam0 + 0.02 * torch.randn(7, generator=g).to(DEV)
    opt = hps.TrackerOptimizer(tracking, cam0.clone(), ef_mine, dec_mine)
    lr = tracking["lr"]
    if separate_LR:
        quad, T = cam0[:4].clone().requires_grad_(True), cam0[-3:].clone().requires_grad_(True)
        groups = [{"params": [T], "lr": lr}, {"params": [quad], "lr": lr * 0.2}]
    else:
        rcam_leaf = cam0.clone().requires_grad_(True)
        groups = [{"params": [rcam_leaf], "lr": lr}]
    groups += [{"params": [ef_ref], "lr": 0.001},
               {"params": list(dec_ref.color_decoder_mid.mlp_exposure.parameters()), "lr": 0.001},
               {"params": list(dec_ref.color_decoder_fine.mlp_exposure.parameters()), "lr": 0.001}]
    topt = torch.optim.Adam(groups)

    def loss_of(cam, ef, dec, k):
        # like the tracker: mid level (and only its exposure MLP) in the first half, fine level in the second - the fine
        # MLP's first Adam step must use bias correction t = 1 (torch keeps state[p]['step'] per parameter)
        level = "mid" if opt.stage(k) == "color_mid" else "fine"
        mlps = list(getattr(dec, "color_decoder_" + level).mlp_exposure.parameters())
        wobble = 1.0 + 0.9 * math.sin(1.7 * k)
        return 100 * wobble * ((cam - target) ** 2).sum() + (ef ** 2).sum() + sum((p * p).sum() for p in mlps) * 0.1 + wobble

    best, candidate = float(1e20), None
    for k in range(opt.num_cam_iters):
        loss = loss_of(opt.camera_tensor(), ef_mine, dec_mine, k)
        loss.backward()
        opt.step(loss)
        rcam = torch.cat([quad, T], 0) if separate_LR else rcam_leaf
        topt.zero_grad()
        rloss = loss_of(rcam, ef_ref, dec_ref, k)
        rloss.backward()
        topt.step()
        if rloss.item() < best:
            best, candidate = rloss.item(), rcam.clone().detach()
        mine_now = torch.cat(opt.leaves, 0).detach() if separate_LR else opt.leaves[0].detach()
        ref_now = torch.cat([quad, T], 0).detach() if separate_LR else rcam_leaf.detach()
        assert float((mine_now - ref_now).abs().max()) < 5e-6, k
    assert abs(float(opt.best_loss) - best) <= 1e-4 * abs(best)
    assert float((opt.best_camera_tensor - candidate).abs().max()) < 5e-6
    assert float((ef_mine.detach() - ef_ref.detach()).abs().max()) < 5e-6
    for level in ("mid", "fine"):
        for a, b in zip(getattr(dec_mine, "color_decoder_" + level).mlp_exposure.parameters(),
                        getattr(dec_ref, "color_decoder_" + level).mlp_exposure.parameters()):
            assert float((a.detach() - b.detach()).abs().max()) < 5e-6, level
    steps = opt.opt.tensor_steps.tolist()
    assert max(steps) == opt.num_cam_iters and min(steps) == opt.num_cam_iters // 2


def _peer_worker(rank, world, port, out):
    import importlib
    import os
    import sys
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    hps = importlib.import_module("hierarchical-point-slam_b200")
    n = 100_003 * 4
    ex = hps.parallel.PeerAllReduce(n, dev, dist)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    res = {}
    for variant in ex.variants():
        mine = torch.randn(n, generator=g, device=dev)
        ref = mine.clone()
        dist.all_reduce(ref)
        for n_used in (n, 4096, 4):                       # whole buffer, a prefix, one vector
            ex.buffer.copy_(mine)
            dist.barrier()
            ex(variant, n_floats=n_used)
            torch.cuda.synchronize()
            ex.check()
            res[(variant, n_used)] = (float((ex.buffer[:n_used] - ref[:n_used]).abs().max() / ref.abs().max()),
                                      bool(torch.equal(ex.buffer[n_used:], mine[n_used:])))
            dist.barrier()
        # a sub-range (the early exchange of the feature-table gradients / the rest behind it): floats in front of and
        # behind it stay as they were
        for off, n_used in ((4096, 20_000), (n - 4000, 4000)):
            ex.buffer.copy_(mine)
            dist.barrier()
            ex(variant, n_floats=n_used, offset_floats=off)
            torch.cuda.synchronize()
            ex.check()
            keep = torch.ones(n, dtype=torch.bool, device=dev)
            keep[off:off + n_used] = False
            res[(variant, "range", off)] = (
                float((ex.buffer[off:off + n_used] - ref[off:off + n_used]).abs().max() / ref.abs().max()),
                bool(torch.equal(ex.buffer[keep], mine[keep])))
            dist.barrier()
        # bit-identical on every rank (the owner of a slice computes its sums once and broadcasts them)
        ex.buffer.copy_(mine)
        dist.barrier()
        ex(variant)
        torch.cuda.synchronize()
        gathered = [torch.empty_like(ex.buffer) for _ in range(world)]
        dist.all_gather(gathered, ex.buffer.clone())
        res[(variant, "identical")] = all(torch.equal(gathered[0], t) for t in gathered)
    # PeerGradExchange in two halves (features() behind the scatter, the rest from after_grads) = the exchange in one
    gx = hps.parallel.PeerGradExchange(n, dev, dist)
    mine = torch.randn(n, generator=g, device=dev)
    ref = mine.clone()
    dist.all_reduce(ref)
    used, feats = 300_000, 262_144
    for split in (False, True):
        gx.pool.tensor.copy_(mine)
        gx.pool.used, gx.pool.features = used, feats
        dist.barrier()
        if split:
            gx.features()
        gx()
        torch.cuda.synchronize()
        gx.check()
        res[("grad_exchange", "split" if split else "single")] = (
            float((gx.pool.tensor[:used] - ref[:used]).abs().max() / ref.abs().max()),
            bool(torch.equal(gx.pool.tensor[used:], mine[used:])))
        dist.barrier()
    if rank == 0:
        torch.save(res, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="the peer-memory exchange needs two GPUs on one box")
def test_peer_allreduce_sums_symmetric_buffers_across_gpus(hps, tmp_path):
    """hps_peer_allreduce (csrc/peer.cu; SURVEY 8e, the exchange step of the sharded mapper iteration): P2P and - where the
    fabric has NVLS - multimem variant against NCCL's all_reduce on two GPUs: the sum within fp32 reordering, untouched
    floats behind the reduced prefix, bit-identical results on both ranks, no timeouts."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "peer.pt")
    mp.spawn(_peer_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out, weights_only=False)
    assert any(k[0] == "p2p" for k in res)
    for key, val in res.items():
        if key[1] == "identical":
            assert val, key
        else:
            assert val[0] < 1e-6 and val[1], (key, val)
