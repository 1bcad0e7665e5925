"""Kernel timeline of one replayed mapper iteration on rank 0 of an N-GPU job (torchrun): the overlapped form with
this package's peer-memory gradient exchange inside the graph.  usage:
torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 scripts/step_timeline_mg.py > out.txt"""
import importlib
import os
import sys

import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = "cuda:%d" % lr
dist.init_process_group("nccl", device_id=torch.device(dev))
hps = importlib.import_module("hierarchical-point-slam_b200")
wl = bench.build_workload(hps, dev)
peer = hps.parallel.PeerGradExchange(2 * int(wl.live_idx.numel()) * 32 + 1_000_000, torch.device(dev), dist, zero_in_forward=True)
slots = [bench._batch_on_device(wl, (2 * rank + b) % bench.N_BATCHES) for b in (0, 1)]
prep = [None, None]
ef = wl.scene["exposure_feat"].to(dev)
targets = [wl.lg.param, wl.lc.param] + list(wl.dec.parameters())


def head(k):
    b = slots[k]
    prep[k] = wl.renderer.prepare_batch(wl.npc, b["rays_d"], b["rays_o"], bench.STAGE, b["gt_depth"],
                                        {"fine": b["radius"], "mid": b["radius_mid"]}, out=prep[k])
    return prep[k].cnt


def step(k):
    for t in targets:
        t.grad = None
    b = slots[k]
    loss = wl.renderer.render_batch_ray_loss(
        wl.npc, wl.dec, b["rays_d"], b["rays_o"], dev, bench.STAGE, b["gt_depth"], b["gt_color"], 0.1,
        npc_geo_feats=wl.gf, npc_col_feats=wl.cf, cloud_pos=wl.pos, dynamic_r_query={"fine": b["radius"], "mid": b["radius_mid"]},
        exposure_feat=ef, prepared=prep[k], grad_buffer=peer.pool,
        tail_work=(lambda: head(1 - k)) if bench.HEAD_AT == "tail" else None,
        after_feature_grads=peer.features if bench.SPLIT_EXCHANGE else None,
        after_grads=lambda: peer(beside=(lambda: head(1 - k)) if bench.HEAD_AT != "tail" else None))[0]
    loss.backward()
    return loss


head(0)
head(1)
graphs = [hps.graphs.GraphedStep(lambda k=k: step(k), {}, warmup=3) for k in (0, 1)]
for i in range(6):
    graphs[i % 2].graph.replay()
torch.cuda.synchronize()
dist.barrier()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(6):
        graphs[i % 2].graph.replay()
    torch.cuda.synchronize()
peer.check()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs = sorted({(e.time_range.start, e.time_range.end, e.name, getattr(e, "device_resource_id", -1)) for e in evs})
    t0 = evs[0][0]
    print("# rank 0 of %d: six back-to-back replays (rest of iteration i + head of i+1 + peer exchange)" % world)
    print("%9s %8s %4s  %s" % ("start_us", "dur_us", "strm", "kernel"))
    for s0, s1, name, strm in evs:
        print("%9.1f %8.1f %4s  %s" % (s0 - t0, s1 - s0, strm, name[:100]))
    print("span %.1f us for 6 iterations" % (max(e[1] for e in evs) - t0))
dist.barrier()
dist.destroy_process_group()
