"""Benchmark of the render-and-optimise hot path: mapping rays/sec (fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" = one mapper iteration of the reference's inner loop (Mapper.py:783-858) on one batch of
synthetic rays: render_batch_ray (stage color_fine) -> L1 depth + colour loss -> backward to the point
features of the level, the colour decoder and (as in the reference) the geometry decoder.
Workload at every N = BASELINE.json configs[1]: 5,000 rays (25,000 sample points) per GPU against a
1M-point fine cloud (+250k mid), ScanNet intrinsics, synthetic data, random-init decoders (seed 1219).
Weak scaling: every rank renders its own 5,000-ray shard of the global batch against the replicated cloud
and the gradients are all-reduced (NCCL) - SURVEY 8e.  At N > 1 the iteration is issued as two CUDA graphs (the
parameter-free head = sampling + neighbour search | the rest) so that the all-reduce of step i overlaps the head of
step i+1; it is started after the untimed L2 flush (it overlaps nothing untimed) and the last one is timed on its own.

JSON line: see the driver's contract; `value` = rays/s with inputs resident in HBM (CUDA-event timed per
step, L2 flushed between steps), `e2e` = same through the public API from pinned host buffers (double-buffered H2D on a copy stream) with the
loss read back every step, `roofline` = dominant kernel vs MEASURED_PEAKS.json, `cpu_baseline` = the
oracle port timed on the host cores on a bounded sample.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RAYS = 5000
N_FINE = 1_000_000
N_MID = 250_000
N_BATCHES = 8            # distinct ray batches cycled through
STAGE = "color_fine"
FUSED_LOSS = os.environ.get("HPS_BENCH_FUSED_LOSS", "1") != "0"   # renderer.render_batch_ray_loss vs render + mapper_loss
EXTENT = 10.0            # cloud spread (in image-plane widths): ~2,700 pts/m^2, ScanNet-like
CPU_SAMPLE_RAYS = 2500


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops"], tf_sus=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf=1590.0, tf_sus=1400.0, src="fallback")


def make_scene(seed=1219):
    import synthetic as S          # data generator only (the oracle is used by the CPU arm / cpu_baseline alone)
    fine = S.synthetic_scene(N_FINE, RAYS * N_BATCHES, (0.04, 0.16), seed=seed, extent_scale=EXTENT)
    mid = S.synthetic_scene(N_MID, 8, (0.24, 0.96), seed=seed + 1, extent_scale=EXTENT)
    g = torch.Generator().manual_seed(seed + 2)
    radius_mid = 0.24 + 0.72 * torch.rand(RAYS * N_BATCHES, generator=g, dtype=torch.float64)
    return dict(fine=fine, mid=mid, radius_mid=radius_mid,
                c_fill_geo=torch.randn(32, generator=g) * 0.01, c_fill_col=torch.randn(32, generator=g) * 0.01,
                exposure_feat=torch.randn(8, generator=g) * 0.1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region: ONE long-running `nvidia-smi -lms 100`
    process started before the region and stopped after it (B200_PROFILING.md recipe) - no per-sample fork that
    would perturb a host-bound step."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        rows = []
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            for line in out.strip().splitlines():
                parts = [x.strip() for x in line.split(",")]
                if len(parts) >= 6:
                    try:
                        float(parts[0])
                        rows.append(parts)
                    except ValueError:
                        pass
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(rows)}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_step_fn(scene, n_rays):
    from oracle import hps_oracle as O
    torch.manual_seed(1219)
    hps = importlib.import_module("hierarchical-point-slam_b200")
    dec_mod = hps.get_model(hps.scannet_cfg("cpu"))
    dec = {}
    for name in ("geo_decoder_fine", "color_decoder_fine"):
        m = getattr(dec_mod, name)
        d = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items() if v.is_floating_point()}
        d["embedder._B"] = m.embedder._B.detach().clone().requires_grad_(name.startswith("geo"))
        dec[name] = d
    knn = O.KDTreeNeighbors()
    fine = scene["fine"]
    pos = {"fine": fine["points"]}
    gf = {"fine": fine["geo_feats"].clone().requires_grad_(True)}
    cf = {"fine": fine["col_feats"].clone().requires_grad_(True)}

    def step(b):
        sl = slice(b * RAYS, b * RAYS + n_rays)
        for t in (gf["fine"], cf["fine"]):
            t.grad = None
        d, v, c, m = O.render_batch_ray(pos, dec, fine["rays_d"][sl], fine["rays_o"][sl], STAGE, fine["gt_depth"][sl],
                                        gf, cf, {"fine": fine["radius"][sl]}, scene["c_fill_geo"], scene["c_fill_col"],
                                        knn_fn=knn)
        aff = O.exposure_affine(dec["color_decoder_fine"], scene["exposure_feat"])
        loss = O.mapper_loss(d, c, m, fine["gt_depth"][sl], fine["gt_color"][sl], STAGE, 0.1, aff)
        loss.backward()
        return float(loss.detach())
    return step


def time_cpu(scene, n_rays, steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    step = cpu_step_fn(scene, n_rays)
    for i in range(warmup):
        step(i % N_BATCHES)
    t0 = time.perf_counter()
    for i in range(steps):
        step(i % N_BATCHES)
    dt = time.perf_counter() - t0
    return n_rays * steps / dt, dt / steps


def run_reference(args, rank):
    if rank != 0:
        return
    scene = make_scene()
    steps, warmup = max(1, min(args.steps, 10)), max(1, min(args.warmup, 2))
    rps, sps = time_cpu(scene, CPU_SAMPLE_RAYS, steps, warmup)
    cores = os.cpu_count() or 1
    sample = ("%d-ray mapper iteration (stage %s, fwd+loss+bwd) vs the 1M-point fine cloud; oracle port of the "
              "reference (torch CPU fp32) with scipy cKDTree(workers=-1) standing in for faiss-cpu" % (CPU_SAMPLE_RAYS, STAGE))
    line = {"impl": "reference", "metric": "mapping rays/sec (fwd+bwd)", "value": rps, "unit": "rays/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": sps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_REAL_STDOUT = None


def emit(line):
    """Print the result line on the process's original stdout."""
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return {"workload": "BASELINE configs[1]: ScanNet-shaped mapping step, %d sampled rays x 5 samples fwd+bwd (stage %s) "
                        "against a 1M-point fine cloud (+250k mid), per GPU" % (RAYS, STAGE),
            "rays_per_gpu": RAYS, "global_rays": RAYS * n_gpus, "fine_points": N_FINE, "mid_points": N_MID,
            "cloud_density": "extent_scale %.0f (~2.7k pts/m^2)" % EXTENT,
            "l2": "flushed between steps (256 MiB memset, untimed); %d distinct ray batches cycled" % N_BATCHES,
            "live_rows": "frustum-selected rows of the fine level are the optimised leaves (reference frustum_feature_selection)",
            "parallelism": "rays sharded x%d, cloud replicated, gradient all-reduce (NCCL)" % n_gpus if n_gpus > 1 else "single GPU"}


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import __graft_entry__
    dist = None
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        __graft_entry__.build()          # the other ranks wait, then load the finished library
    if world > 1:
        dist.barrier()
    if os.environ.get("HPS_BENCH_VERBOSE"):
        print("[rank %d] process group up, library built" % rank, file=sys.stderr, flush=True)
    hps = importlib.import_module("hierarchical-point-slam_b200")
    allreduce_grads = hps.parallel.allreduce_grads
    dev = "cuda:%d" % local_rank
    torch.cuda.set_device(local_rank)
    scene = make_scene()
    cfg = hps.scannet_cfg(dev)
    npc = hps.NeuralPointCloud(cfg)
    for level in ("fine", "mid"):
        s = scene[level]
        npc.set_cloud(level, s["points"].to(dev), s["geo_feats"].to(dev), s["col_feats"].to(dev))
    torch.manual_seed(1219)
    dec = hps.get_model(cfg).to(dev)
    dec.c_fill_geo, dec.c_fill_col = scene["c_fill_geo"].to(dev), scene["c_fill_col"].to(dev)
    import types
    renderer = hps.Renderer(cfg, None, types.SimpleNamespace(**{k: cfg["cam"][k] for k in ("H", "W", "fx", "fy", "cx", "cy")}))
    renderer.sigmoid_coefficient = 0.1
    fine = scene["fine"]
    # rank r takes every world-th batch offset so that ranks render different rays of the global batch
    host = {k: fine[k].pin_memory() for k in ("rays_o", "rays_d", "gt_depth", "gt_color", "radius")}
    host["radius_mid"] = scene["radius_mid"].pin_memory()
    devb = {k: v.to(dev) for k, v in host.items()}
    exposure_feat = scene["exposure_feat"].to(dev)
    # Frustum feature selection (Mapper.py:194-250, 512-536): the rows of the fine level that project into the
    # frame (+20 px) are the live, optimised rows; the others are frozen.  Identity pose, ScanNet intrinsics.
    K = cfg["cam"]
    pf = npc.cloud_pos_tensor("fine")
    zc = -pf[:, 2]
    u = K["cx"] + K["fx"] * pf[:, 0] / zc
    v = K["cy"] - K["fy"] * pf[:, 1] / zc
    live = (zc > 0) & (u > -20) & (u < K["W"] + 20) & (v > -20) & (v < K["H"] + 20)
    live_idx = live.nonzero().flatten()
    lg, lc = hps.LiveRows.pair(npc.get_geo_feats("fine"), npc.get_col_feats("fine"), live_idx)
    gf = {"fine": lg, "mid": npc.get_geo_feats("mid")}
    cf = {"fine": lc, "mid": npc.get_col_feats("mid")}
    pos = {l: npc.cloud_pos_tensor(l) for l in ("fine", "mid")}
    params = [p for p in dec.parameters()]
    grad_targets = [lg.param, lc.param] + params
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def batch_index(i):
        return (i * world + rank) % N_BATCHES

    def batch_slice(i):
        b = batch_index(i)
        return slice(b * RAYS, (b + 1) * RAYS)

    renderer.assume_positive_depth = True      # mapper batches are depth-filtered (common.py:248-254)
    renderer.write_back_radius = False         # nobody reads the repeated radii the reference writes into the dict
    # One ray batch travels as ONE packed byte buffer (fp32: rays_o | rays_d | gt_depth | gt_color, then fp64: the
    # two radius vectors) so that staging it costs one copy, not six; the iteration reads contiguous typed views.
    F32_KEYS = (("rays_o", 3), ("rays_d", 3), ("gt_depth", 1), ("gt_color", 3))
    N32 = sum(w for _, w in F32_KEYS) * RAYS                       # floats; 4 * N32 is a multiple of 8

    def pack(src, b):
        sl = slice(b * RAYS, (b + 1) * RAYS)
        f32 = torch.cat([src[k][sl].reshape(-1).float() for k, _ in F32_KEYS])
        f64 = torch.cat([src["radius"][sl].reshape(-1).double(), src["radius_mid"][sl].reshape(-1).double()])
        return torch.cat([f32.view(torch.uint8), f64.view(torch.uint8)])

    host_pk = {"bytes": torch.stack([pack(host, b) for b in range(N_BATCHES)]).pin_memory()}
    dev_pk = {k: v.to(dev) for k, v in host_pk.items()}
    static_pk = {"bytes": torch.empty_like(dev_pk["bytes"][0])}
    s32 = static_pk["bytes"][:4 * N32].view(torch.float32)
    s64 = static_pk["bytes"][4 * N32:].view(torch.float64)
    static, off = {}, 0
    for k, w in F32_KEYS:
        static[k] = s32[off:off + w * RAYS].view(devb[k][:RAYS].shape)
        off += w * RAYS
    static["radius"] = s64[:RAYS].view(devb["radius"][:RAYS].shape)
    static["radius_mid"] = s64[RAYS:].view(devb["radius_mid"][:RAYS].shape)

    prep_box = {"p": None}

    def head():
        """Parameter-free head of the iteration (sampling + neighbour search): its own graph when ranks > 1, so that it
        can run while the previous iteration's gradients are still being exchanged."""
        rq = {"fine": static["radius"], "mid": static["radius_mid"]}
        prep_box["p"] = renderer.prepare_batch(npc, static["rays_d"], static["rays_o"], STAGE, static["gt_depth"], rq,
                                               out=prep_box["p"])
        return prep_box["p"].cnt

    def step(prepared=None):
        """One mapper iteration on the rays held in `static` (reference Mapper.py:783-858)."""
        for t in grad_targets:
            t.grad = None
        rq = {"fine": static["radius"], "mid": static["radius_mid"]}
        if FUSED_LOSS:      # render + loss as one autograd node (compositing and loss in one kernel)
            loss = renderer.render_batch_ray_loss(
                npc, dec, static["rays_d"], static["rays_o"], dev, STAGE, static["gt_depth"], static["gt_color"], 0.1,
                npc_geo_feats=gf, npc_col_feats=cf, cloud_pos=pos, dynamic_r_query=rq, exposure_feat=exposure_feat,
                prepared=prepared)[0]
        else:               # the reference's seams: render_batch_ray, then the loss (HPS_BENCH_FUSED_LOSS=0)
            depth, unc, color, valid = renderer.render_batch_ray(
                npc, dec, static["rays_d"], static["rays_o"], dev, STAGE, gt_depth=static["gt_depth"],
                npc_geo_feats=gf, npc_col_feats=cf, is_tracker=False, cloud_pos=pos, dynamic_r_query=rq,
                exposure_feat=None, prepared=prepared)
            affine = dec.color_decoder_fine.mlp_exposure(exposure_feat)
            loss = hps.mapper_loss(depth, color, valid, static["gt_depth"], static["gt_color"], STAGE, 0.1, affine)
        loss.backward()
        return loss

    grad_exchange = hps.parallel.GradExchange(grad_targets, skip_missing=True)   # untouched (mid-level) decoders stay out
    comm = torch.cuda.Stream() if world > 1 else None
    exch = {"ev": None}

    def exchange():
        """Blocking form (instrumented pass): gradient all-reduce over NCCL on the iteration's stream."""
        if world > 1:
            grad_exchange(dist)

    def post():
        """Start the gradient all-reduce of the iteration just issued on the communication stream.  The fused
        backward's gradients are views of one buffer, so this is one grouped NCCL launch."""
        if world > 1:
            main = torch.cuda.current_stream()
            comm.wait_stream(main)
            with torch.cuda.stream(comm):
                grad_exchange(dist)
                exch["ev"] = torch.cuda.current_stream().record_event()

    def drain():
        if exch["ev"] is not None:
            torch.cuda.current_stream().wait_event(exch["ev"])
            exch["ev"] = None

    def log(msg):
        if os.environ.get("HPS_BENCH_VERBOSE"):
            print("[rank %d] %s" % (rank, msg), file=sys.stderr, flush=True)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- capture the iteration in a CUDA graph (falls back to eager issue if capture is impossible) ----
    for k in static_pk:
        static_pk[k].copy_(dev_pk[k][batch_index(0)])
    mode = "cuda-graph replay"
    try:
        hps.ops.LAUNCHES["n"] = 0
        if world > 1:
            # two graphs: head (no parameters involved) | rest.  Iteration i+1's head overlaps iteration i's exchange;
            # the rest waits for it (a real mapper would put its optimiser step there).
            g_head = hps.graphs.GraphedStep(head, static_pk, warmup=3)
            g_tail = hps.graphs.GraphedStep(lambda: step(prep_box["p"]), {}, warmup=3)
            mode = "cuda-graph replay (head | rest), exchange overlapped with the next head"

            def run(i, src):
                g_head(**{k: src[k][batch_index(i)] for k in static_pk})
                drain()
                return g_tail()

            def replay():                      # inputs already in static_pk
                g_head.graph.replay()
                drain()
                g_tail.graph.replay()
                return g_tail.outputs
        else:
            graphed = hps.graphs.GraphedStep(step, static_pk, warmup=3)

            def run(i, src):
                return graphed(**{k: src[k][batch_index(i)] for k in static_pk})

            def replay():
                graphed.graph.replay()
                return graphed.outputs
        launches_per_step = hps.ops.LAUNCHES["n"] // 4          # 3 warm-up passes + the captured one
    except Exception as e:                                        # pragma: no cover - reported in the JSON line
        mode = "eager (graph capture failed: %s)" % str(e).splitlines()[0][:120]
        hps.ops.LAUNCHES["n"] = 0
        step()
        launches_per_step = hps.ops.LAUNCHES["n"]

        def run(i, src):
            for k in static_pk:
                static_pk[k].copy_(src[k][batch_index(i)], non_blocking=True)
            drain()
            return step()

        def replay():
            drain()
            return step()

    log("iteration ready: %s" % mode)
    # ---- warm-up ----
    for i in range(args.warmup):
        run(i, dev_pk)
        post()
    drain()
    sync_all()
    # ---- timed: K steps, per-step CUDA events, L2 flushed (untimed) between steps; inputs resident in HBM.  With
    # ranks > 1 the exchange of step i is started after the flush that follows it, i.e. it overlaps the head of step
    # i+1 and nothing untimed; the last exchange is timed on its own. ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = []
    sync_all()
    for i in range(args.steps):
        flush.zero_()
        if i > 0:
            post()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run(args.warmup + i, dev_pk)
        b.record()
        evs.append((a, b))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    post()
    drain()
    b.record()
    evs.append((a, b))
    sync_all()
    log("timed region done")
    launches = launches_per_step * args.steps
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    # ---- e2e: every step's inputs come from pinned host memory and every step's loss is read back on the host, wall
    # clock around the synchronised region.  The loop is pipelined like a real mapper loop would be: step i+1's H2D
    # copy runs on a copy stream into the other of two staging buffers while step i computes, and the host reads the
    # loss of step i-1 (pinned, event-ordered) after it has issued step i - the GPU never waits for the host. ----
    copy_s = torch.cuda.Stream()
    main_s = torch.cuda.current_stream()
    stage = [torch.empty_like(static_pk["bytes"]) for _ in range(2)]
    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    ev_in, ev_used, ev_out = [None, None], [None, None], [None, None]
    seen = []

    def feed(i):
        """H2D copy of step i's ray batch (copy stream; waits until the staging buffer's previous reader is done)."""
        with torch.cuda.stream(copy_s):
            if ev_used[i % 2] is not None:
                copy_s.wait_event(ev_used[i % 2])
            stage[i % 2].copy_(host_pk["bytes"][batch_index(i)], non_blocking=True)
            ev_in[i % 2] = copy_s.record_event()

    def e2e_loop(n, first):
        feed(first)
        for k in range(n):
            i = first + k
            if k + 1 < n:
                feed(i + 1)
            main_s.wait_event(ev_in[i % 2])
            static_pk["bytes"].copy_(stage[i % 2], non_blocking=True)
            ev_used[i % 2] = main_s.record_event()
            loss = replay()
            post()                                                  # (ranks > 1) start the exchange
            loss_host[i % 2].copy_(loss.detach().reshape(()), non_blocking=True)      # D2H read of the loss
            ev_out[i % 2] = main_s.record_event()
            if k > 0:
                ev_out[(i - 1) % 2].synchronize()
                seen.append(float(loss_host[(i - 1) % 2]))
        ev_out[(first + n - 1) % 2].synchronize()
        seen.append(float(loss_host[(first + n - 1) % 2]))

    e2e_loop(3, 0)
    drain()
    sync_all()
    seen.clear()
    t0 = time.perf_counter()
    e2e_loop(args.steps, args.warmup)
    drain()
    sync_all()
    e2e_s = time.perf_counter() - t0
    assert len(seen) == args.steps and all(v == v for v in seen), "e2e: every step's loss must have been read back"
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s = float(t[0]), float(t[1])
    # ---- per-kernel timing pass (eager, instrumented, separate from the timed region) ----
    roof = None
    hps.ops.TIMERS = {}
    n_prof = min(args.steps, 20)
    for i in range(n_prof):              # every rank takes part (the step contains the all-reduce)
        flush.zero_()
        for k in static_pk:
            static_pk[k].copy_(dev_pk[k][batch_index(args.warmup + i)])
        step()
        exchange()
    torch.cuda.synchronize()
    log("profiling pass done")
    per = {k: sum(a.elapsed_time(b) for a, b in v) / n_prof for k, v in hps.ops.TIMERS.items()}
    hps.ops.TIMERS = None
    if rank == 0:
        roof = roofline(per, hps, npc, devb, batch_slice(0))
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    tracking = None
    if world == 1:
        tracking = time_tracker(hps, npc, dec, renderer, scene, cfg, dev, gf, cf, pos)
    h2d = sum(v[0].numel() * v.element_size() for v in host_pk.values())
    line = {"metric": "mapping rays/sec (fwd+bwd)", "value": RAYS * world * args.steps / (dev_ms * 1e-3), "unit": "rays/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(world), issue=mode, n_live_rows=int(live_idx.numel()),
                           api="renderer.render_batch_ray_loss + backward" if FUSED_LOSS else
                           "renderer.render_batch_ray + mapper_loss + backward"), "clocks": clocks,
            "e2e": {"value": RAYS * world * args.steps / e2e_s, "unit": "rays/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4,
                    "note": "every step: H2D of its ray batch from pinned memory (copy stream, two staging buffers) and "
                            "D2H of its loss (read by the host one step behind); no L2 flush"},
            "gpu_launches": launches, "roofline": roof["roofline"], "kernels_ms_per_step": roof["per_kernel"]}
    if tracking is not None:
        line["tracking"] = tracking
    if world == 1:
        rps, sps = time_cpu(scene, CPU_SAMPLE_RAYS, 3, 1)
        line["cpu_baseline"] = {"value": rps, "unit": "rays/s", "cores": os.cpu_count() or 1, "kind": "port",
                                "sample": "%d-ray mapper iteration of the same workload, oracle port (torch CPU fp32 + scipy "
                                          "cKDTree for the neighbour search), 3 timed repeats" % CPU_SAMPLE_RAYS}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def time_tracker(hps, npc, dec, renderer, scene, cfg, dev, gf, cf, pos, iters=100):
    """The second half of BASELINE.json's metric: tracking iterations / s (Tracker.py:190-235) on the same scene - R = 5000
    rays, stage color_fine, pose 7-vector + exposure latent under Adam, decoders frozen; the whole iteration (pose -> rays
    -> render -> loss -> backward -> Adam) replayed from one CUDA graph, timed with CUDA events.  Extra key of the JSON
    line, not part of `value`."""
    K = cfg["cam"]
    fine = scene["fine"]
    rd0 = fine["rays_d"][:RAYS]
    pi = (rd0[:, 0] * K["fx"] + K["cx"]).to(dev)
    pj = (K["cy"] - rd0[:, 1] * K["fy"]).to(dev)
    gt_d, gt_c = fine["gt_depth"][:RAYS].to(dev), fine["gt_color"][:RAYS].to(dev)
    rq = {"fine": fine["radius"][:RAYS].to(dev), "mid": scene["radius_mid"][:RAYS].to(dev)}
    feats_g = {l: (v.table if isinstance(v, hps.LiveRows) else v) for l, v in gf.items()}
    feats_c = {l: (v.table if isinstance(v, hps.LiveRows) else v) for l, v in cf.items()}
    req = [p.requires_grad for p in dec.parameters()]
    for p in dec.parameters():
        p.requires_grad_(False)
    try:
        cam = torch.tensor([1.0, 0.0, 0.0, 0.0, 0.005, -0.004, 0.003], device=dev).requires_grad_(True)
        ef = scene["exposure_feat"].to(dev).clone().requires_grad_(True)
        opt = hps.FusedAdam([([cam, ef], 5e-4)], own_grads=False)      # torch.optim.Adam semantics, one kernel

        def it():
            opt.zero_grad()
            c2w = hps.get_camera_from_tensor(cam)
            ro, rd = hps.get_rays_from_uv(pi, pj, c2w, K["H"], K["W"], K["fx"], K["fy"], K["cx"], K["cy"], dev)
            if FUSED_LOSS:
                loss = renderer.render_batch_ray_tracker_loss(
                    npc, dec, rd, ro, dev, STAGE, gt_d, gt_c, 0.5, npc_geo_feats=feats_g, npc_col_feats=feats_c,
                    cloud_pos=pos, dynamic_r_query=dict(rq), exposure_feat=ef)[0]
            else:
                d, v, c, _ = renderer.render_batch_ray(npc, dec, rd, ro, dev, STAGE, gt_depth=gt_d, npc_geo_feats=feats_g,
                                                       npc_col_feats=feats_c, is_tracker=True, cloud_pos=pos,
                                                       dynamic_r_query=dict(rq), exposure_feat=ef)
                loss = hps.tracker_loss(d, v, c, gt_d, gt_c, 0.5)
            loss.backward()
            opt.step()
            return loss

        g = hps.graphs.GraphedStep(it, {}, warmup=3)
        for _ in range(5):
            g()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(iters):
            g()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / iters
        return {"value": 1e3 / ms, "unit": "iters/s", "ms_per_iter": ms, "rays": RAYS, "stage": STAGE,
                "note": "pose (7) + exposure latent (8) under Adam, decoders frozen, whole iteration replayed from one "
                        "CUDA graph, back to back (no L2 flush), 1 GPU"}
    except Exception as e:                                    # pragma: no cover - reported, never fatal
        return {"value": None, "unit": "iters/s", "error": str(e).splitlines()[0][:160]}
    finally:
        for p, r in zip(dec.parameters(), req):
            p.requires_grad_(r)


def roofline(per, hps, npc, devb, sl):
    """Dominant kernel of the step vs the measured peak.  Algorithmic work per launch:
    colour MLP fwd 193,400 FLOP/pt, bwd (dgrad+wgrad) 2 x 193,400; geometry 30,958 / 2 x 30,958;
    gather 2,432 B/pt (two tables); search (84 + 12*cand) B/query."""
    P = RAYS * 5
    pk = peaks()
    work = {"hps_mlp_fwd[1]": ("tensor", 193400.0 * P), "hps_mlp_bwd[1]": ("tensor", 2 * 193400.0 * P),
            "hps_mlp_fwd[0]": ("tensor", 30958.0 * P), "hps_mlp_bwd[0]": ("tensor", 2 * 30958.0 * P),
            "hps_interp_fwd": ("hbm", 2432.0 * P), "hps_interp_bwd": ("hbm", 2432.0 * P)}
    # tensor-core entry points: same algorithmic (fp32-equivalent) FLOPs; the 3-product fp16 split is overhead
    for k in (0, 1):
        work["hps_mlp_tc_fwd[%d]" % k] = work["hps_mlp_fwd[%d]" % k]
        work["hps_mlp_tc_bwd[%d]" % k] = work["hps_mlp_bwd[%d]" % k]
    # measured mean candidate count for the search figure
    dev = devb["rays_o"].device
    z5 = devb["gt_depth"][sl][:, None, None] * torch.linspace(0.96, 1.04, 5, device=dev)[None, :, None]
    pts = (devb["rays_o"][sl][:, None, :] + devb["rays_d"][sl][:, None, :] * z5).reshape(-1, 3)
    _, cand = npc.query_neighbors(pts, step="query", dynamic_radius=devb["radius"][sl].repeat_interleave(5), level="fine",
                                  want_cand=True)
    mean_cand = float(cand.float().mean())
    work["hps_knn_query"] = ("hbm", (84.0 + 12.0 * mean_cand) * P)
    work["hps_sample_knn"] = ("hbm", (84.0 + 12.0 * mean_cand + 60.0) * P)     # + the sampling kernel's 36 B in, 24 B out
    top = max(per, key=per.get)
    bound, amount = work.get(top, ("hbm", 0.0))
    ms = per[top]
    if bound == "tensor":
        achieved, peak, unit = amount / (ms * 1e-3) / 1e12, pk["tf"], "TFLOP/s"
    else:
        achieved, peak, unit = amount / (ms * 1e-3) / 1e9, pk["hbm"], "GB/s"
    detail = {}
    for k, (b, amt) in work.items():
        if k in per and per[k] > 0:
            detail[k] = {"ms": per[k], "achieved": amt / (per[k] * 1e-3) / (1e12 if b == "tensor" else 1e9),
                         "unit": "TFLOP/s" if b == "tensor" else "GB/s",
                         "frac": amt / (per[k] * 1e-3) / (1e12 if b == "tensor" else 1e9) / (pk["tf"] if b == "tensor" else pk["hbm"])}
    traffic = None      # DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/)
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get(top)
    except (OSError, ValueError):
        pass
    return {"roofline": {"kernel": top, "bound": bound, "achieved": achieved, "peak": peak, "unit": unit,
                         "frac": achieved / peak, "traffic": traffic, "peak_source": pk["src"],
                         "mean_candidates_per_query": mean_cand, "all": detail},
            "per_kernel": per}


def main():
    # fd 1 carries the JSON line only: libraries (NCCL prints its version banner there) go to stderr meanwhile
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    if world == 1 and args.gpus > 1:
        # plain `python bench.py --gpus N`: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__), "--gpus", str(args.gpus),
               "--steps", str(args.steps), "--warmup", str(args.warmup)]
        os.dup2(_REAL_STDOUT, 1)                 # the ranks inherit the real stdout
        raise SystemExit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
