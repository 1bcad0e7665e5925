"""Kernel timeline of one replayed mapper iteration (torch.profiler / CUPTI): start offset, duration, stream.
usage (GPU box): python scripts/step_timeline.py > gpurun_out/timeline.txt"""
import importlib
import os
import sys
import types

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

hps = importlib.import_module("hierarchical-point-slam_b200")
dev = "cuda:0"
scene = bench.make_scene()
cfg = hps.scannet_cfg(dev)
npc = hps.NeuralPointCloud(cfg)
for level in ("fine", "mid"):
    s = scene[level]
    npc.set_cloud(level, s["points"].to(dev), s["geo_feats"].to(dev), s["col_feats"].to(dev))
torch.manual_seed(1219)
dec = hps.get_model(cfg).to(dev)
dec.c_fill_geo, dec.c_fill_col = scene["c_fill_geo"].to(dev), scene["c_fill_col"].to(dev)
renderer = hps.Renderer(cfg, None, types.SimpleNamespace(**{k: cfg["cam"][k] for k in ("H", "W", "fx", "fy", "cx", "cy")}))
renderer.sigmoid_coefficient = 0.1
renderer.assume_positive_depth = True
renderer.write_back_radius = False
fine = scene["fine"]
R = bench.RAYS
st = {k: fine[k][:R].to(dev) for k in ("rays_o", "rays_d", "gt_depth", "gt_color", "radius")}
st["radius_mid"] = scene["radius_mid"][:R].to(dev)
exposure_feat = scene["exposure_feat"].to(dev)
K = cfg["cam"]
pf = npc.cloud_pos_tensor("fine")
zc = -pf[:, 2]
u = K["cx"] + K["fx"] * pf[:, 0] / zc
v = K["cy"] - K["fy"] * pf[:, 1] / zc
live_idx = ((zc > 0) & (u > -20) & (u < K["W"] + 20) & (v > -20) & (v < K["H"] + 20)).nonzero().flatten()
lg, lc = hps.LiveRows.pair(npc.get_geo_feats("fine"), npc.get_col_feats("fine"), live_idx)
gf = {"fine": lg, "mid": npc.get_geo_feats("mid")}
cf = {"fine": lc, "mid": npc.get_col_feats("mid")}
pos = {l: npc.cloud_pos_tensor(l) for l in ("fine", "mid")}
targets = [lg.param, lc.param] + list(dec.parameters())


def step():
    for t in targets:
        t.grad = None
    rq = {"fine": st["radius"], "mid": st["radius_mid"]}
    if bench.FUSED_LOSS:
        loss = renderer.render_batch_ray_loss(
            npc, dec, st["rays_d"], st["rays_o"], dev, bench.STAGE, st["gt_depth"], st["gt_color"], 0.1,
            npc_geo_feats=gf, npc_col_feats=cf, cloud_pos=pos, dynamic_r_query=rq, exposure_feat=exposure_feat)[0]
    else:
        depth, unc, color, valid = renderer.render_batch_ray(
            npc, dec, st["rays_d"], st["rays_o"], dev, bench.STAGE, gt_depth=st["gt_depth"], npc_geo_feats=gf,
            npc_col_feats=cf, is_tracker=False, cloud_pos=pos, dynamic_r_query=rq, exposure_feat=None)
        affine = dec.color_decoder_fine.mlp_exposure(exposure_feat)
        loss = hps.mapper_loss(depth, color, valid, st["gt_depth"], st["gt_color"], bench.STAGE, 0.1, affine)
    loss.backward()
    return loss


g = hps.graphs.GraphedStep(step, {}, warmup=3)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    g()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    flush.zero_()
    torch.cuda.synchronize()
    g()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Memset" not in e.name or
       (e.device_type == torch.autograd.DeviceType.CUDA)]
evs = sorted({(e.time_range.start, e.time_range.end, e.name, getattr(e, "device_resource_id", -1)) for e in evs})
evs = [e for e in evs if "FillFunctor<unsigned char>" not in e[2]]
t0 = evs[0][0]
print("%9s %8s %4s  %s" % ("start_us", "dur_us", "strm", "kernel"))
for s0, s1, name, strm in evs:
    print("%9.1f %8.1f %4s  %s" % (s0 - t0, s1 - s0, strm, name[:110]))
print("span %.1f us" % (evs[-1][1] - t0))
