"""Kernel timeline of one replayed mapper iteration (torch.profiler / CUPTI): start offset, duration, stream.
Default: the bench's overlapped form (rest of iteration i + head of iteration i+1 beside its backward tail);
HPS_BENCH_OVERLAP=0: the whole iteration alone.
usage (GPU box): python scripts/step_timeline.py > gpurun_out/timeline.txt"""
import importlib
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

hps = importlib.import_module("hierarchical-point-slam_b200")
dev = "cuda:0"
wl = bench.build_workload(hps, dev)
R = bench.RAYS
slots = [bench._batch_on_device(wl, b) for b in (0, 1)]
prep = [None, None]
ef = wl.scene["exposure_feat"].to(dev)
targets = [wl.lg.param, wl.lc.param] + list(wl.dec.parameters())


def head(k):
    b = slots[k]
    prep[k] = wl.renderer.prepare_batch(wl.npc, b["rays_d"], b["rays_o"], bench.STAGE, b["gt_depth"],
                                        {"fine": b["radius"], "mid": b["radius_mid"]}, out=prep[k])
    return prep[k].cnt


def step(k, prepared=None, tail=None):
    for t in targets:
        t.grad = None
    b = slots[k]
    loss = wl.renderer.render_batch_ray_loss(
        wl.npc, wl.dec, b["rays_d"], b["rays_o"], dev, bench.STAGE, b["gt_depth"], b["gt_color"], 0.1,
        npc_geo_feats=wl.gf, npc_col_feats=wl.cf, cloud_pos=wl.pos, dynamic_r_query={"fine": b["radius"], "mid": b["radius_mid"]},
        exposure_feat=ef, prepared=prepared, tail_work=tail)[0]
    loss.backward()
    return loss


if bench.OVERLAP_HEAD:
    head(0)
    head(1)
    graphs = [hps.graphs.GraphedStep(lambda k=k: step(k, prep[k], tail=lambda: head(1 - k)), {}, warmup=3) for k in (0, 1)]
else:
    graphs = [hps.graphs.GraphedStep(lambda: step(0), {}, warmup=3)] * 2
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for i in range(4):
    graphs[i % 2].graph.replay()
torch.cuda.synchronize()
def capture(which):
    """Three replays back to back (L2 flushed in front of each, as in the bench's timed loop); the MIDDLE one is printed:
    the host is then ahead of the device, as in the bench - an isolated replay shows the driver's node-by-node
    submission (root kernels of later branches start tens of microseconds late) on top of the device's schedule."""
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        torch.cuda.synchronize()
        for j in range(3):
            flush.zero_()
            graphs[(which + j) % 2].graph.replay()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs = sorted({(e.time_range.start, e.time_range.end, e.name, getattr(e, "device_resource_id", -1)) for e in evs})
    marks = [e[1] for e in evs if "FillFunctor<unsigned char>" in e[2]]      # ends of the three flushes
    if ISOLATED or len(marks) < 3:
        lo, hi = (marks[0] if marks else evs[0][0]), (evs[-1][1] + 1 if len(marks) < 2 else marks[1] - 1)
    else:
        lo, hi = marks[1], marks[2]
    evs = [e for e in evs if "FillFunctor<unsigned char>" not in e[2] and lo <= e[0] < hi]
    t0 = evs[0][0]
    print("%9s %8s %4s  %s" % ("start_us", "dur_us", "strm", "kernel"))
    for s0, s1, name, strm in evs:
        print("%9.1f %8.1f %4s  %s" % (s0 - t0, s1 - s0, strm, name[:110]))
    print("span %.1f us (flush end -> last kernel end %.1f us)" % (max(e[1] for e in evs) - t0, max(e[1] for e in evs) - lo))


ISOLATED = os.environ.get("HPS_TIMELINE_ISOLATED", "0") != "0"
print("# %s" % ("rest of iteration i + head of iteration i+1 beside the backward tail" if bench.OVERLAP_HEAD else "whole iteration"))
for rep in range(int(os.environ.get("HPS_TIMELINE_REPS", "1"))):
    capture(rep % 2)
