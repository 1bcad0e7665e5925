"""Kernel timeline of one replayed tracker iteration (torch.profiler / CUPTI): start offset, duration, stream.
The iteration is bench.time_tracker's (pose -> rays -> render -> loss -> backward -> optimiser step, one CUDA graph).
usage (GPU box): python scripts/tracker_timeline.py > gpurun_out/tracker_timeline.txt"""
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
gf = {l: npc.get_geo_feats(l) for l in ("fine", "mid")}
cf = {l: npc.get_col_feats(l) for l in ("fine", "mid")}
pos = {l: npc.cloud_pos_tensor(l) for l in ("fine", "mid")}

made = []
Orig = hps.graphs.GraphedStep


class Keep(Orig):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        made.append(self)


hps.graphs.GraphedStep = Keep
res = bench.time_tracker(hps, npc, dec, renderer, scene, cfg, dev, gf, cf, pos, iters=20)
print("# %s" % res)
g = made[-1]
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    g()
    torch.cuda.synchronize()
evs = sorted({(e.time_range.start, e.time_range.end, e.name, getattr(e, "device_resource_id", -1))
              for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA})
t0 = evs[0][0]
print("%9s %8s %4s  %s" % ("start_us", "dur_us", "strm", "kernel"))
for s0, s1, name, strm in evs:
    print("%9.1f %8.1f %4s  %s" % (s0 - t0, s1 - s0, strm, name[:110]))
print("span %.1f us, %d kernels" % (evs[-1][1] - t0, len(evs)))
