"""Debug aid (GPU): where does the tracker pose-gradient discrepancy come from?"""
import sys, os, importlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import hps_oracle as O
from tests import parity
import __graft_entry__
__graft_entry__.build()
DEV = "cuda:0"
scene = parity.load("scene_small.pt"); ref = parity.load("render_tracker_small.pt")
hps, npc, dec, renderer = parity.product(scene, DEV)
cl, cfg, inp = scene["clouds"], scene["cfg"], ref["inputs"]
K = O.SCANNET_INTRINSICS

def oracle(dt, stage):
    cvt = lambda t: t.to(dt) if t.is_floating_point() else t
    pos = {l: cvt(cl[l]["points"]) for l in cl}
    cam = cvt(inp["cam"]).clone().requires_grad_(True)
    ef = cvt(scene["exposure_feat"]).clone().requires_grad_(True)
    decd = {k: {kk: cvt(vv) for kk, vv in v.items()} for k, v in scene["decoders"].items()}
    ro, rd = O.rays_from_uv(cvt(inp["i"]), cvt(inp["j"]), O.camera_from_tensor(cam), K["fx"], K["fy"], K["cx"], K["cy"])
    ro.retain_grad(); rd.retain_grad()
    d, v, c, _ = O.render_batch_ray(pos, decd, rd, ro, stage, cvt(inp["gt_depth"]),
        {l: cvt(cl[l]["geo_feats"]) for l in cl}, {l: cvt(cl[l]["col_feats"]) for l in cl},
        {l: cl[l]["radius"] for l in cl}, cvt(scene["c_fill_geo"]), cvt(scene["c_fill_col"]), is_tracker=True, exposure_feat=ef)
    loss = O.tracker_loss(d, v, c, cvt(inp["gt_depth"]), cvt(inp["gt_color"]), cfg["w_color_tracker"])
    loss.backward()
    return cam.grad, rd.grad, ro.grad

for stage in ("color_mid", "color_fine"):
    g32, rd32, ro32 = oracle(torch.float32, stage)
    g64, rd64, ro64 = oracle(torch.float64, stage)
    cam = inp["cam"].to(DEV).clone().requires_grad_(True)
    ef = scene["exposure_feat"].to(DEV).clone().requires_grad_(True)
    c2w = hps.get_camera_from_tensor(cam)
    ro, rd = hps.get_rays_from_uv(inp["i"].to(DEV), inp["j"].to(DEV), c2w, K["H"], K["W"], K["fx"], K["fy"], K["cx"], K["cy"], DEV)
    ro.retain_grad(); rd.retain_grad()
    rq = {l: cl[l]["radius"].to(DEV).clone() for l in cl}
    depth, var, color, _ = renderer.render_batch_ray(npc, dec, rd, ro, DEV, stage, gt_depth=inp["gt_depth"].to(DEV),
        npc_geo_feats={l: npc.get_geo_feats(l) for l in cl}, npc_col_feats={l: npc.get_col_feats(l) for l in cl},
        is_tracker=True, cloud_pos={l: npc.cloud_pos_tensor(l) for l in cl}, dynamic_r_query=rq, exposure_feat=ef)
    loss = hps.tracker_loss(depth, var, color, inp["gt_depth"].to(DEV), inp["gt_color"].to(DEV), cfg["w_color_tracker"])
    loss.backward()
    print(stage)
    print("  g_cam   ours-vs-f64 %.2e  oracle32-vs-f64 %.2e  ours-vs-oracle32 %.2e" % (parity.rel(cam.grad, g64), parity.rel(g32, g64), parity.rel(cam.grad, g32)))
    print("  d/drays_d ours-vs-f64 %.2e  oracle32-vs-f64 %.2e" % (parity.rel(rd.grad, rd64), parity.rel(rd32, rd64)))
    print("  d/drays_o ours-vs-f64 %.2e  oracle32-vs-f64 %.2e" % (parity.rel(ro.grad, ro64), parity.rel(ro32, ro64)))
    e_ours = (rd.grad.cpu().double() - rd64).abs().sum(1); e_32 = (rd32.double() - rd64).abs().sum(1)
    k = torch.topk(e_ours, 5)
    print("  worst rays ours:", k.indices.tolist(), ["%.3g" % x for x in k.values.tolist()], "their oracle32 err:", ["%.3g" % x for x in e_32[k.indices].tolist()])
    print("  |rd64| rows:", ["%.3g" % x for x in rd64.abs().sum(1)[k.indices].tolist()])
    print("  sum abs err ours %.3g oracle32 %.3g ; total |grad| %.3g" % (e_ours.sum(), e_32.sum(), rd64.abs().sum()))
