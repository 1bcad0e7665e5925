"""Multi-GPU plumbing probes (run under torchrun, one mode per process group so that a hang costs one timeout):
  nccl_eager   time a 3.4 MB fp32 all-reduce issued eagerly (CUDA events, device time)
  nccl_graph   capture the same all-reduce inside a CUDA graph and replay it
  symm         torch.distributed._symmetric_memory: allocate, rendezvous, report the handles (multicast?), time the
               library's one-shot / two-shot / multimem all-reduce, capture them in a graph
  own          this package's own peer-memory all-reduce kernel (hps_peer_allreduce): correctness, time, graph replay
  iteration    a real mapper iteration per rank with its gradients in the symmetric pool; peer sum vs NCCL sum
usage: torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 scripts/multi_gpu_probe.py MODE"""
import os
import sys
import time

import torch
import torch.distributed as dist

mode = sys.argv[1]
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
N = 850_000          # 3.4 MB of fp32: live-row gradients of two tables + the decoders' gradients


def say(*a):
    if rank == 0:
        print("[%s, %d GPUs]" % (mode, world), *a, flush=True)


def timed(fn, n=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]) * 1e3


if mode == "nccl_eager":
    x = torch.ones(N, device=dev)
    say("eager all_reduce of %.1f MB: %.1f us per call (back to back)" % (N * 4 / 1e6, timed(lambda: dist.all_reduce(x))))
    y = torch.ones(8, device=dev)
    say("eager all_reduce of 32 B: %.1f us per call" % timed(lambda: dist.all_reduce(y)))
elif mode == "nccl_graph":
    x = torch.full((N,), float(rank + 1), device=dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            dist.all_reduce(x)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    say("warm-up done, capturing")
    g = torch.cuda.CUDAGraph()
    x.fill_(float(rank + 1))
    with torch.cuda.graph(g):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    say("captured")
    want = sum(range(1, world + 1))
    for i in range(3):
        x.fill_(float(rank + 1))
        g.replay()
        torch.cuda.synchronize()
        say("replay %d: x[0] = %.1f (want %.1f)" % (i, float(x[0]), want))
    say("graph replay of the all_reduce: %.1f us per replay" % timed(lambda: g.replay()))
elif mode == "symm":
    import torch.distributed._symmetric_memory as symm_mem
    group = dist.group.WORLD
    t = symm_mem.empty(N, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, group.group_name)
    say("rendezvous ok: world %d, buffer_ptrs %d, multicast_ptr %s, signal_pad_ptrs %d" % (
        hdl.world_size, len(hdl.buffer_ptrs), hex(hdl.multicast_ptr) if hdl.multicast_ptr else None, len(hdl.signal_pad_ptrs)))
    want = float(sum(range(1, world + 1)))
    for name in ("one_shot_all_reduce", "two_shot_all_reduce_", "multimem_all_reduce_", "multimem_one_shot_all_reduce"):
        op = getattr(torch.ops.symm_mem, name, None)
        if op is None:
            say(name, "not in this build")
            continue
        try:
            t.fill_(float(rank + 1))
            dist.barrier()
            out = op(t, "sum", group.group_name)
            torch.cuda.synchronize()
            res = out if out is not None else t
            say("%s: result[0] = %.1f (want %.1f), %.1f us per call" % (
                name, float(res[0]), want, timed(lambda: op(t, "sum", group.group_name))))
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                op(t, "sum", group.group_name)
            torch.cuda.synchronize()
            say("%s inside a CUDA graph: %.1f us per replay" % (name, timed(lambda: g.replay())))
        except Exception as e:      # noqa: BLE001
            say(name, "failed:", str(e).splitlines()[0][:200])
elif mode == "own":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import importlib
    hps = importlib.import_module("hierarchical-point-slam_b200")
    ex = hps.parallel.PeerAllReduce(N, dev, dist)
    want = float(sum(range(1, world + 1)))
    for variant in ex.variants():
        try:
            ex.buffer.fill_(float(rank + 1))
            dist.barrier()
            ex(variant)
            torch.cuda.synchronize()
            ok = bool((ex.buffer == want).all())
            say("own %s: result ok %s (x[0] = %.1f, x[-1] = %.1f), %.1f us per call" % (
                variant, ok, float(ex.buffer[0]), float(ex.buffer[-1]), timed(lambda: ex(variant))))
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                ex(variant)
            torch.cuda.synchronize()
            say("own %s inside a CUDA graph: %.1f us per replay" % (variant, timed(lambda: g.replay())))
        except Exception as e:      # noqa: BLE001
            say("own", variant, "failed:", str(e).splitlines()[0][:200])
elif mode == "iteration":
    # the real pipeline: every rank renders ITS ray batch (fused render + loss + backward, gradients laid out in the
    # symmetric pool), then (a) this package's peer all-reduce in place and (b) NCCL on a copy: the two sums must agree
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import importlib
    import bench
    hps = importlib.import_module("hierarchical-point-slam_b200")
    wl = bench.build_workload(hps, "cuda:%d" % lr)
    ex = hps.parallel.PeerGradExchange(2 * int(wl.live_idx.numel()) * 32 + 1_000_000, dev, dist)
    b = bench._batch_on_device(wl, rank % bench.N_BATCHES)
    targets = [wl.lg.param, wl.lc.param] + list(wl.dec.parameters())
    for variant in ex.peer.variants():
        for t in targets:
            t.grad = None
        loss = wl.renderer.render_batch_ray_loss(
            wl.npc, wl.dec, b["rays_d"], b["rays_o"], "cuda:%d" % lr, bench.STAGE, b["gt_depth"], b["gt_color"], 0.1,
            npc_geo_feats=wl.gf, npc_col_feats=wl.cf, cloud_pos=wl.pos, dynamic_r_query={"fine": b["radius"], "mid": b["radius_mid"]},
            exposure_feat=wl.scene["exposure_feat"].to(dev), grad_buffer=ex.pool)[0]
        loss.backward()
        torch.cuda.synchronize()
        used = ex.pool.used
        inside = [t for t in targets if t.grad is not None and
                  ex.pool.tensor.data_ptr() <= t.grad.data_ptr() < ex.pool.tensor.data_ptr() + 4 * used]
        outside = [t for t in targets if t.grad is not None and not any(t is u for u in inside)]
        ref = ex.pool.tensor[:used].clone()
        dist.all_reduce(ref)
        ex(variant)
        torch.cuda.synchronize()
        ex.check()
        got = ex.pool.tensor[:used]
        err = float((got - ref).abs().max() / ref.abs().max())
        say("iteration, %s: %d gradients inside the pool (%d floats), %d outside; peer sum vs NCCL sum: max rel diff %.2e; "
            "live-row gradient is a view of the pool: %s" % (variant, len(inside), used, len(outside), err,
                                                            wl.lg.param.grad.data_ptr() >= ex.pool.tensor.data_ptr()))
dist.barrier()
dist.destroy_process_group()
