"""Shared parity harness: product (CUDA, through the C ABI) vs oracle (CPU) on the committed small scene.
Used by the -m gpu tests and by __graft_entry__.smoke().  Tolerance: 1e-4 relative (max|a-b| / max|b|) for
fp32 quantities as BASELINE.json's north_star states; neighbour indices / masks bit-exact."""
import importlib
import os

import torch

from oracle import hps_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-4


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def load(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def product(scene, device):
    """npc / decoders / renderer of the product package holding the fixture's state."""
    hps = importlib.import_module("hierarchical-point-slam_b200")
    cfg = hps.scannet_cfg(device)
    npc = hps.NeuralPointCloud(cfg)
    for level, cl in scene["clouds"].items():
        npc.set_cloud(level, cl["points"].to(device), cl["geo_feats"].to(device), cl["col_feats"].to(device))
    dec = hps.get_model(cfg)
    for name, sd in scene["decoders"].items():
        m = getattr(dec, name)
        sd = dict(sd)
        B = sd.pop("embedder._B")
        if name.startswith("geo"):
            sd["embedder._B"] = B
        else:
            m.embedder._B = B.clone()
        m.load_state_dict(sd, strict=True)
    dec = dec.to(device)
    dec.c_fill_geo = scene["c_fill_geo"].to(device)
    dec.c_fill_col = scene["c_fill_col"].to(device)
    K = O.SCANNET_INTRINSICS
    import types
    renderer = hps.Renderer(cfg, None, types.SimpleNamespace(**K))
    renderer.sigmoid_coefficient = scene["cfg"]["coef"]
    return hps, npc, dec, renderer


def mapper_step(hps, npc, dec, renderer, scene, stage, device):
    cl, rays, cfg = scene["clouds"], scene["rays"], scene["cfg"]
    level = "mid" if stage.endswith("mid") else "fine"
    gf = {l: cl[l]["geo_feats"].to(device).clone().requires_grad_(True) for l in cl}
    cf = {l: cl[l]["col_feats"].to(device).clone().requires_grad_(True) for l in cl}
    pos = {l: npc.cloud_pos_tensor(l) for l in cl}
    dec.zero_grad()
    rq = {l: cl[l]["radius"].to(device).clone() for l in cl}
    gt_d, gt_c = rays["gt_depth"].to(device), rays["gt_color"].to(device)
    depth, var, color, valid = renderer.render_batch_ray(
        npc, dec, rays["rays_d"].to(device), rays["rays_o"].to(device), device, stage, gt_depth=gt_d,
        npc_geo_feats=gf, npc_col_feats=cf, is_tracker=False, cloud_pos=pos, dynamic_r_query=rq,
        exposure_feat=None)
    affine = None
    if stage.startswith("color"):
        affine = getattr(dec, "color_decoder_" + level).mlp_exposure(scene["exposure_feat"].to(device))
    loss = hps.mapper_loss(depth, color, valid, gt_d, gt_c, stage, cfg["w_color_mapper"], affine)
    loss.backward()
    return dict(depth=depth, var=var, color=color, valid=valid, loss=loss, g_geo=gf[level].grad,
                g_col=cf[level].grad)


def check_mapper(stage, out, ref, dec, tol=TOL):
    level = "mid" if stage.endswith("mid") else "fine"
    errs = {}
    assert torch.equal(out["valid"].cpu(), ref["valid"]), "valid_ray_mask differs"
    for k in ("depth", "var", "color", "loss", "g_geo"):
        errs[k] = rel(out[k], ref[k])
    if ref["g_col"] is not None:
        errs["g_col"] = rel(out["g_col"], ref["g_col"])
    cdec = getattr(dec, ("color_decoder_" if stage.startswith("color") else "geo_decoder_") + level)
    named = dict(cdec.named_parameters())
    worst = 0.0
    for k, g in ref["g_dec"].items():
        assert named[k].grad is not None, "no gradient for %s" % k
        worst = max(worst, rel(named[k].grad, g))
    errs["g_dec"] = worst
    if stage.startswith("color"):   # geometry decoder also receives gradients
        gnamed = dict(getattr(dec, "geo_decoder_" + level).named_parameters())
        errs["g_geo_dec"] = max(rel(gnamed[k].grad, g) for k, g in ref["g_geo_dec"].items())
    bad = {k: v for k, v in errs.items() if not v < tol}
    assert not bad, "stage %s over tolerance %g: %s (all: %s)" % (stage, tol, bad, errs)
    return errs


def tracker_step(hps, npc, dec, renderer, scene, ref_inputs, stage, device):
    cl, cfg = scene["clouds"], scene["cfg"]
    K = O.SCANNET_INTRINSICS
    cam = ref_inputs["cam"].to(device).clone().requires_grad_(True)
    ef = scene["exposure_feat"].to(device).clone().requires_grad_(True)
    dec.zero_grad()
    c2w = hps.get_camera_from_tensor(cam)
    ro, rd = hps.get_rays_from_uv(ref_inputs["i"].to(device), ref_inputs["j"].to(device), c2w, K["H"], K["W"],
                                  K["fx"], K["fy"], K["cx"], K["cy"], device)
    gt_d, gt_c = ref_inputs["gt_depth"].to(device), ref_inputs["gt_color"].to(device)
    rq = {l: cl[l]["radius"].to(device).clone() for l in cl}
    depth, var, color, _ = renderer.render_batch_ray(
        npc, dec, rd, ro, device, stage, gt_depth=gt_d,
        npc_geo_feats={l: npc.get_geo_feats(l) for l in cl}, npc_col_feats={l: npc.get_col_feats(l) for l in cl},
        is_tracker=True, cloud_pos={l: npc.cloud_pos_tensor(l) for l in cl}, dynamic_r_query=rq, exposure_feat=ef)
    loss = hps.tracker_loss(depth, var, color, gt_d, gt_c, cfg["w_color_tracker"])
    loss.backward()
    level = "mid" if stage.endswith("mid") else "fine"
    mexp = getattr(dec, "color_decoder_" + level).mlp_exposure
    return dict(depth=depth, var=var, color=color, loss=loss, g_cam=cam.grad, g_exposure=ef.grad,
                g_mlp_exposure={k: v.grad for k, v in mexp.named_parameters()})


def check_tracker(stage, out, ref, tol=TOL):
    errs = {k: rel(out[k], ref[k]) for k in ("depth", "var", "color", "loss", "g_cam", "g_exposure")}
    errs["g_mlp_exposure"] = max(rel(out["g_mlp_exposure"][k], g) for k, g in ref["g_mlp_exposure"].items())
    bad = {k: v for k, v in errs.items() if not v < tol}
    assert not bad, "tracker %s over tolerance %g: %s (all: %s)" % (stage, tol, bad, errs)
    return errs


def run_small_scene(device="cuda:0", stages=("geometry_mid", "color_mid", "geometry_fine", "color_fine"),
                    tracker=True):
    scene = load("scene_small.pt")
    hps, npc, dec, renderer = product(scene, device)
    report = {}
    ref = load("render_mapper_small.pt")
    for stage in stages:
        out = mapper_step(hps, npc, dec, renderer, scene, stage, device)
        report["mapper/" + stage] = check_mapper(stage, out, ref[stage], dec)
    if tracker:
        tref = load("render_tracker_small.pt")
        for stage in ("color_mid", "color_fine"):
            out = tracker_step(hps, npc, dec, renderer, scene, tref["inputs"], stage, device)
            report["tracker/" + stage] = check_tracker(stage, out, tref[stage])
    return report
