"""Which scheduling of the next batch's head (sampling + neighbour search) beside the backward tail is fastest?
Tries stream priority / launch position / one vs two launches, 60 replays each with the L2 flushed between them.
usage (GPU box): python scripts/tail_experiment.py"""
import importlib
import itertools
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

hps = importlib.import_module("hierarchical-point-slam_b200")
ops = hps.ops
dev = "cuda:0"
wl = bench.build_workload(hps, dev)
slots = [bench._batch_on_device(wl, b) for b in (0, 1)]
ef = wl.scene["exposure_feat"].to(dev)
targets = [wl.lg.param, wl.lc.param] + list(wl.dec.parameters())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def measure(overlap):
    prep = [None, None]

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
            wl.npc, wl.dec, b["rays_d"], b["rays_o"], dev, bench.STAGE, b["gt_depth"], b["gt_color"], 0.1, npc_geo_feats=wl.gf,
            npc_col_feats=wl.cf, cloud_pos=wl.pos, dynamic_r_query={"fine": b["radius"], "mid": b["radius_mid"]},
            exposure_feat=ef, prepared=prepared, tail_work=tail)[0]
        loss.backward()
        return loss
    if overlap:
        head(0)
        head(1)
        graphs = [hps.graphs.GraphedStep(lambda k=k: step(k, prep[k], tail=lambda: head(1 - k)), {}, warmup=3) for k in (0, 1)]
    else:
        graphs = [hps.graphs.GraphedStep(lambda: step(0), {}, warmup=3)] * 2
    for i in range(6):
        graphs[i % 2].graph.replay()
    torch.cuda.synchronize()
    evs = []
    for i in range(60):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graphs[i % 2].graph.replay()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return sum(ts) / len(ts), ts[len(ts) // 2]


print("no overlap (whole iteration in one graph): mean %.4f ms, median %.4f ms" % measure(False))
for prio, after, fused, wprio in itertools.product((-1, 0), (False, True), (True, False), (-2, 0)):
    ops.TAIL_PRIORITY, ops.TAIL_AFTER_SCATTER, ops.PREPARE_FUSED, ops.WGRAD_PRIORITY = prio, after, fused, wprio
    mean, med = measure(True)
    print("tail priority %2d, %-22s, %s, wgrad priority %2d: mean %.4f ms, median %.4f ms"
          % (prio, "behind colour wgrad" if after else "before the scatter", "1 launch " if fused else "2 launches", wprio, mean, med))
