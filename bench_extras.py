"""Extra workloads of bench.py's JSON line (BASELINE.json configs[0], [2], [4] and the strong-scaling form of [3]).

Every function returns a small dict that becomes one extra key of the line; none of them enters `value`.  Each runs
under its own try/except in run(): a failing extra is reported as {"error": ...} and never costs the headline number.

    tracking        configs[2]  200-iteration coarse-to-fine pose optimisation (100 x color_mid, 100 x color_fine,
                                Tracker.py:426-434) on one GPU: iterations / s, and the pose after a short 3 + 3 run next to
                                the CPU oracle's trajectory of the same run
    single_frame    configs[0]  one 640x480 frame against a 50k / 200k cloud: render_img per level + 1 tracker iteration,
                                the CPU oracle timed on a 3000-ray chunk beside it, the chunk compared pixel by pixel
    full_frame      configs[4]  8M fine + 2M mid points: full-frame 307,200-ray render with the image rows dealt out to
                                the ranks, and the neighbour-search sweep over 1M .. 8M points
    strong_scaling  configs[3]  ONE reference-sized 10,000-ray batch (mapping.pixels) sharded over the ranks
"""
import importlib
import os
import time
import types

import torch

import synthetic as S

K_ITERS_TRACK = 200


def _err(e):
    return {"error": "%s: %s" % (type(e).__name__, str(e).splitlines()[0][:200] if str(e) else "")}


def _events():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _max_over_ranks(ms, dist, dev):
    if dist is None:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def _timed(fn, n, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = _events()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


# ----------------------------------------------------------------------------------------------------------------
# configs[2]: coarse-to-fine tracking
# ----------------------------------------------------------------------------------------------------------------
def _tracker_inputs(wl, n_batches, rays):
    scene, K, dev = wl.scene, wl.cfg["cam"], wl.dev
    fine = scene["fine"]
    out = []
    for b in range(n_batches):
        sl = slice(b * rays, (b + 1) * rays)
        rd0 = fine["rays_d"][sl]
        out.append(dict(pi=(rd0[:, 0] * K["fx"] + K["cx"]).to(dev), pj=(K["cy"] - rd0[:, 1] * K["fy"]).to(dev),
                        gt_d=fine["gt_depth"][sl].to(dev), gt_c=fine["gt_color"][sl].to(dev),
                        r_fine=fine["radius"][sl].to(dev), r_mid=scene["radius_mid"][sl].to(dev)))
    return out


CAM0 = [1.0, 0.0, 0.0, 0.0, 0.005, -0.004, 0.003]        # ~1 cm off the true (identity) pose


def tracking(hps, wl, bench):
    dev, cfg, npc, dec, renderer, K = wl.dev, wl.cfg, wl.npc, wl.dec, wl.renderer, wl.cfg["cam"]
    R = bench.RAYS
    batches = _tracker_inputs(wl, bench.N_BATCHES, R)
    static = {k: torch.empty_like(v) for k, v in batches[0].items()}
    feats_g = {l: (v.table if isinstance(v, hps.LiveRows) else v) for l, v in wl.gf.items()}
    feats_c = {l: (v.table if isinstance(v, hps.LiveRows) else v) for l, v in wl.cf.items()}
    req = [p.requires_grad for p in dec.parameters()]
    mlps = [dec.color_decoder_mid.mlp_exposure, dec.color_decoder_fine.mlp_exposure]
    saved_mlp = [[p.detach().clone() for p in m.parameters()] for m in mlps]
    for p in dec.parameters():
        p.requires_grad_(False)
    for m in mlps:                       # the tracker optimises the pose, the exposure latent and both exposure MLPs
        for p in m.parameters():         # (Tracker.py:383-415); every other decoder tensor is frozen
            p.requires_grad_(True)
            p.grad = None
    try:
        def make(iters):
            cam0 = torch.tensor(CAM0, device=dev)
            ef = wl.scene["exposure_feat"].to(dev).clone().requires_grad_(True)
            return hps.TrackerOptimizer(dict(cfg["tracking"], iters=iters), cam0, ef, dec), ef

        def iteration(opt, ef, stage, src):
            cam = opt.camera_tensor()
            c2w = hps.get_camera_from_tensor(cam)
            ro, rd = hps.get_rays_from_uv(src["pi"], src["pj"], c2w, K["H"], K["W"], K["fx"], K["fy"], K["cx"], K["cy"], dev)
            loss = renderer.render_batch_ray_tracker_loss(
                npc, dec, rd, ro, dev, stage, src["gt_d"], src["gt_c"], 0.5, npc_geo_feats=feats_g, npc_col_feats=feats_c,
                cloud_pos=wl.pos, dynamic_r_query={"fine": src["r_fine"], "mid": src["r_mid"]}, exposure_feat=ef)[0]
            loss.backward()
            opt.step(loss)
            return loss

        def restore():
            with torch.no_grad():
                for m, sv in zip(mlps, saved_mlp):
                    for p, q in zip(m.parameters(), sv):
                        p.copy_(q)
                        p.grad = None

        # ---- (1) a short 3 + 3 run next to the CPU oracle's trajectory (same rays every iteration) ----
        restore()
        opt, ef = make(6)
        for k in range(6):
            iteration(opt, ef, opt.stage(k), batches[0])
        cam_gpu = opt.leaves[0].detach().cpu().double()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cam_cpu = _oracle_tracker_run(wl, bench, 6, saved_mlp)
        cpu_s = time.perf_counter() - t0
        # compared as camera matrices: the 7-vector's quaternion is not normalised and the pose does not depend on its
        # length (common.py:304-328 divides by |q|^2), so d loss / d |q| is pure rounding noise that Adam turns into
        # +-lr steps of the quaternion's length on either side - invisible in [R | t]
        c2w = lambda cam: hps.get_camera_from_tensor(cam.float()).double()
        start = c2w(torch.tensor(CAM0, dtype=torch.float64))
        vs_oracle = {"iterations": "3 x color_mid + 3 x color_fine, %d rays, Adam on pose + exposure latent + both exposure MLPs" % R,
                     "max_abs_c2w_diff": float((c2w(cam_gpu) - c2w(cam_cpu)).abs().max()),
                     "c2w_moved_by": float((c2w(cam_cpu) - start).abs().max()),
                     "max_abs_7vector_diff": float((cam_gpu - cam_cpu).abs().max()),
                     "cpu_s_per_iter": cpu_s / 6}
        # ---- (2) the 200-iteration run, one CUDA graph per level ----
        restore()
        opt, ef = make(K_ITERS_TRACK)
        state0 = [t.detach().clone() for t in (opt.leaves[0], ef)]
        graphs = {}
        for stage in ("color_mid", "color_fine"):
            for t in opt.opt.params:
                t.grad = None
            graphs[stage] = hps.graphs.GraphedStep(lambda stage=stage: iteration(opt, ef, stage, static), {}, warmup=3)

        def reset():
            restore()
            with torch.no_grad():
                opt.leaves[0].copy_(state0[0])
                ef.copy_(state0[1])
                opt.opt._flat.zero_()
                opt.opt.tensor_steps.zero_()
                opt.opt.step_count.zero_()
                opt.best_loss.fill_(1e20)
                opt.best_camera_tensor.copy_(state0[0])

        def run_frame():
            for k in range(K_ITERS_TRACK):
                src = batches[k % len(batches)]
                for key, v in src.items():
                    static[key].copy_(v, non_blocking=True)
                graphs[opt.stage(k)].graph.replay()

        reset()
        run_frame()                      # warm-up frame
        torch.cuda.synchronize()
        reset()
        a, b = _events()
        a.record()
        run_frame()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        per = {}
        for stage in ("color_mid", "color_fine"):
            per[stage] = _timed(lambda stage=stage: graphs[stage].graph.replay(), 50, warm=3)
        return {"value": K_ITERS_TRACK * 1e3 / ms, "unit": "iters/s", "iterations": K_ITERS_TRACK, "ms_per_frame": ms,
                "ms_per_iter": {"color_mid": per["color_mid"], "color_fine": per["color_fine"]}, "rays": R,
                "final_loss": float(graphs["color_fine"].outputs.detach()), "best_loss": float(opt.best_loss),
                "pose_vs_cpu_oracle": vs_oracle,
                "note": "BASELINE configs[2]: %d x color_mid then %d x color_fine (Tracker.py:426-434), pose + exposure latent + "
                        "both exposure MLPs under TrackerOptimizer (one FusedAdam launch), a fresh ray batch every iteration "
                        "(%d cycled), one CUDA graph per level, back to back, 1 GPU; random-init decoders, so the pose does "
                        "not converge to the true one - parity is the pose next to the CPU oracle's after 3 + 3 iterations"
                        % (K_ITERS_TRACK // 2, K_ITERS_TRACK // 2, len(batches))}
    finally:
        with torch.no_grad():
            for m, sv in zip(mlps, saved_mlp):
                for p, q in zip(m.parameters(), sv):
                    p.copy_(q)
                    p.grad = None
        for p, r in zip(dec.parameters(), req):
            p.requires_grad_(r)


def _oracle_tracker_run(wl, bench, iters, saved_mlp):
    """The same short tracker run on the CPU oracle (torch.optim.Adam over the reference's groups, Tracker.py:383-415)."""
    from oracle import hps_oracle as O
    scene, K = wl.scene, wl.cfg["cam"]
    R = bench.RAYS
    fine, mid = scene["fine"], scene["mid"]
    dec = {}
    for name in ("geo_decoder_fine", "color_decoder_fine", "geo_decoder_mid", "color_decoder_mid"):
        m = getattr(wl.dec, name)
        d = {k: v.detach().cpu().clone() for k, v in m.state_dict().items() if v.is_floating_point()}
        d["embedder._B"] = m.embedder._B.detach().cpu().clone()
        dec[name] = d
    groups = []
    for name, sv in (("color_decoder_mid", saved_mlp[0]), ("color_decoder_fine", saved_mlp[1])):
        keys = ["mlp_exposure.linear1.weight", "mlp_exposure.linear1.bias", "mlp_exposure.linear2.weight",
                "mlp_exposure.linear2.bias"]
        for k, v in zip(keys, sv):
            dec[name][k] = v.detach().cpu().clone().requires_grad_(True)
        groups.append({"params": [dec[name][k] for k in keys], "lr": 0.001})
    cam = torch.tensor(CAM0).requires_grad_(True)
    ef = scene["exposure_feat"].clone().requires_grad_(True)
    opt = torch.optim.Adam([{"params": [cam], "lr": wl.cfg["tracking"]["lr"]}, {"params": [ef], "lr": 0.001}] + groups)
    knn = O.KDTreeNeighbors()
    pos = {"fine": fine["points"], "mid": mid["points"]}
    gf = {"fine": fine["geo_feats"], "mid": mid["geo_feats"]}
    cf = {"fine": fine["col_feats"], "mid": mid["col_feats"]}
    rd0 = fine["rays_d"][:R]
    pi, pj = rd0[:, 0] * K["fx"] + K["cx"], K["cy"] - rd0[:, 1] * K["fy"]
    gt_d, gt_c = fine["gt_depth"][:R], fine["gt_color"][:R]
    rq = {"fine": fine["radius"][:R], "mid": scene["radius_mid"][:R]}
    for k in range(iters):
        stage = "color_mid" if k / iters < 0.5 else "color_fine"
        opt.zero_grad()
        ro, rd = O.rays_from_uv(pi, pj, O.camera_from_tensor(cam), K["fx"], K["fy"], K["cx"], K["cy"])
        d, v, c, _ = O.render_batch_ray(pos, dec, rd, ro, stage, gt_d, gf, cf, rq, scene["c_fill_geo"], scene["c_fill_col"],
                                        is_tracker=True, exposure_feat=ef, knn_fn=knn)
        loss = O.tracker_loss(d, v, c, gt_d, gt_c, 0.5)
        loss.backward()
        opt.step()
    return cam.detach().double()


# ----------------------------------------------------------------------------------------------------------------
# configs[0]: one frame, 50k / 200k cloud
# ----------------------------------------------------------------------------------------------------------------
def _oracle_frame_chunk(wl, fine, mid, c2w, depth_map, rad, ef, chunk):
    """CPU leg of single_frame: the oracle port renders the frame's first `chunk` pixels (row-major, common.py:431-448)
    for both colour stages -> {stage: (seconds, depth, colour)}."""
    from oracle import hps_oracle as O
    K = wl.cfg["cam"]
    H, W = K["H"], K["W"]
    dec_cpu = {}
    for name in ("geo_decoder_fine", "color_decoder_fine", "geo_decoder_mid", "color_decoder_mid"):
        m = getattr(wl.dec, name)
        d = {k: v.detach().cpu().clone() for k, v in m.state_dict().items() if v.is_floating_point()}
        d["embedder._B"] = m.embedder._B.detach().cpu().clone()
        dec_cpu[name] = d
    knn = O.KDTreeNeighbors()
    ii, jj = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing="ij")
    ii, jj = ii.t().reshape(-1)[:chunk], jj.t().reshape(-1)[:chunk]
    ro_c, rd_c = O.rays_from_uv(ii, jj, c2w[:3], K["fx"], K["fy"], K["cx"], K["cy"])
    out = {}
    for stage in ("color_mid", "color_fine"):
        t0 = time.perf_counter()
        with torch.no_grad():
            d, v, c, _ = O.render_batch_ray({"fine": fine["points"], "mid": mid["points"]}, dec_cpu, rd_c, ro_c.contiguous(), stage,
                                            depth_map.reshape(-1)[:chunk], {"fine": fine["geo_feats"], "mid": mid["geo_feats"]},
                                            {"fine": fine["col_feats"], "mid": mid["col_feats"]},
                                            {k: r.reshape(-1)[:chunk] for k, r in rad.items()}, wl.scene["c_fill_geo"],
                                            wl.scene["c_fill_col"], exposure_feat=ef, knn_fn=knn)
        out[stage] = (time.perf_counter() - t0, d, c)
    return out


def single_frame(hps, wl, bench):
    dev, cfg, K = wl.dev, wl.cfg, wl.cfg["cam"]
    ext = bench.EXTENT * (200_000 / bench.N_FINE) ** 0.5          # the bench cloud's density
    fine = S.synthetic_scene(200_000, 8, (0.04, 0.16), seed=31, extent_scale=ext)
    mid = S.synthetic_scene(50_000, 8, (0.24, 0.96), seed=32, extent_scale=ext)
    npc = hps.NeuralPointCloud(cfg)
    for level, s in (("fine", fine), ("mid", mid)):
        npc.set_cloud(level, s["points"].to(dev), s["geo_feats"].to(dev), s["col_feats"].to(dev))
    renderer = hps.Renderer(cfg, None, types.SimpleNamespace(**{k: K[k] for k in ("H", "W", "fx", "fy", "cx", "cy")}))
    renderer.sigmoid_coefficient = 0.1
    H, W = K["H"], K["W"]
    depth_map = S.synthetic_depth()
    g = torch.Generator().manual_seed(5)
    depth_map[torch.rand(H, W, generator=g) < 0.05] = 0.0               # ~5 % pixels without a depth reading
    rad = {"fine": 0.04 + 0.12 * torch.rand(H, W, generator=g, dtype=torch.float64),
           "mid": 0.24 + 0.72 * torch.rand(H, W, generator=g, dtype=torch.float64)}
    c2w = torch.eye(4)
    ef = wl.scene["exposure_feat"]
    gf = {l: npc.get_geo_feats(l) for l in ("fine", "mid")}
    cf = {l: npc.get_col_feats(l) for l in ("fine", "mid")}
    pos = {l: npc.cloud_pos_tensor(l) for l in ("fine", "mid")}
    dmap_d, rad_d, c2w_d, ef_d = depth_map.to(dev), {k: v.to(dev) for k, v in rad.items()}, c2w.to(dev), ef.to(dev)
    res = {"cloud": "200k fine + 50k mid points", "frame": "640x480, ~5 % depth-less pixels"}
    chunk = 3000                                                       # the reference's ray_batch_size
    cpu = _oracle_frame_chunk(wl, fine, mid, c2w, depth_map, rad, ef, chunk)
    for stage in ("color_mid", "color_fine"):
        def frame():
            return renderer.render_img(npc, wl.dec, c2w_d, dev, stage, gt_depth=dmap_d, npc_geo_feats=gf, npc_col_feats=cf,
                                       dynamic_r_query={k: v.clone() for k, v in rad_d.items()}, cloud_pos=pos,
                                       exposure_feat=ef_d)
        ms = _timed(frame, 3, warm=1)
        d_img, u_img, c_img = frame()
        torch.cuda.synchronize()
        cpu_s, d, c = cpu[stage]
        res[stage] = {"gpu_ms_per_frame": ms, "gpu_rays_per_s": H * W / (ms * 1e-3),
                      "cpu_s_per_%d_ray_chunk" % chunk: cpu_s, "cpu_rays_per_s": chunk / cpu_s,
                      "depth_rel": bench._rel(d_img.reshape(-1)[:chunk].float(), d),
                      "color_rel": bench._rel(c_img.reshape(-1, 3)[:chunk], c)}
    res["note"] = ("BASELINE configs[0]: render_img per level on the GPU (65,536-ray chunks) and the CPU oracle on the frame's "
                   "first 3000-ray chunk (the reference's chunk size; its sampling range for depth-less pixels is that chunk's "
                   "statistic on both sides), compared pixel by pixel; the tracker iteration of this config is `tracking`")
    res["parity_ok"] = all(res[s]["depth_rel"] <= bench.PARITY_TOL and res[s]["color_rel"] <= bench.PARITY_TOL
                           for s in ("color_mid", "color_fine"))
    return res


# ----------------------------------------------------------------------------------------------------------------
# configs[4]: 8M + 2M points, full-frame render sharded over image rows + neighbour-search sweep
# ----------------------------------------------------------------------------------------------------------------
def _cloud_on_device(n, seed, extent_scale, dev):
    """synthetic.synthetic_scene's cloud, generated on the device with a seeded generator (identical on every rank)."""
    K = S.SCANNET_INTRINSICS
    g = torch.Generator(device=dev).manual_seed(seed)
    depth = S.synthetic_depth().to(dev)
    u = torch.rand(n, generator=g, device=dev) * (K["W"] - 1) * extent_scale - (extent_scale - 1) * K["W"] / 2
    v = torch.rand(n, generator=g, device=dev) * (K["H"] - 1) * extent_scale - (extent_scale - 1) * K["H"] / 2
    d = depth[v.clamp(0, K["H"] - 1).long(), u.clamp(0, K["W"] - 1).long()]
    d = d * (0.96 + 0.04 * torch.randint(0, 3, (n,), generator=g, device=dev).float())
    return torch.stack([(u - K["cx"]) / K["fx"] * d, -(v - K["cy"]) / K["fy"] * d, -d], -1).float().contiguous(), g


def full_frame(hps, wl, bench, rank, world, dist):
    dev, cfg, K = wl.dev, wl.cfg, wl.cfg["cam"]
    H, W = K["H"], K["W"]
    n_fine, n_mid = 8_000_000, 2_000_000
    ext = bench.EXTENT * (n_fine / bench.N_FINE) ** 0.5
    npc = hps.NeuralPointCloud(cfg)
    for level, n, seed in (("fine", n_fine, 41), ("mid", n_mid, 42)):
        pts, g = _cloud_on_device(n, seed, ext, dev)
        npc.set_cloud(level, pts, torch.randn(n, 32, generator=g, device=dev) * 0.1,
                      torch.randn(n, 32, generator=g, device=dev) * 0.1)
    renderer = hps.Renderer(cfg, None, types.SimpleNamespace(**{k: K[k] for k in ("H", "W", "fx", "fy", "cx", "cy")}))
    renderer.sigmoid_coefficient = 0.1
    depth_map = S.synthetic_depth()
    gcpu = torch.Generator().manual_seed(5)
    depth_map[torch.rand(H, W, generator=gcpu) < 0.05] = 0.0
    rad = {"fine": (0.04 + 0.12 * torch.rand(H, W, generator=gcpu, dtype=torch.float64)).to(dev),
           "mid": (0.24 + 0.72 * torch.rand(H, W, generator=gcpu, dtype=torch.float64)).to(dev)}
    dmap, c2w, ef = depth_map.to(dev), torch.eye(4, device=dev), wl.scene["exposure_feat"].to(dev)
    gf = {l: npc.get_geo_feats(l) for l in ("fine", "mid")}
    cf = {l: npc.get_col_feats(l) for l in ("fine", "mid")}
    pos = {l: npc.cloud_pos_tensor(l) for l in ("fine", "mid")}
    par = hps.parallel
    res = {"cloud": "8M fine + 2M mid points, replicated per GPU", "rays": H * W, "n_gpus": world,
           "sharding": "image rows dealt out to the ranks (parallel.render_img_sharded), row blocks all-gathered"}
    d_ = dist if world > 1 else None
    for stage in ("color_mid", "color_fine"):
        def frame(gather=True):
            return par.render_img_sharded(renderer, d_, npc, wl.dec, c2w, dev, stage, gather=gather, gt_depth=dmap,
                                          npc_geo_feats=gf, npc_col_feats=cf,
                                          dynamic_r_query={k: v.clone() for k, v in rad.items()}, cloud_pos=pos,
                                          exposure_feat=ef)
        frame()
        if d_ is not None:
            d_.barrier()
        torch.cuda.synchronize()
        a, b = _events()
        n = 3
        a.record()
        for _ in range(n):
            out = frame()
        b.record()
        torch.cuda.synchronize()
        ms = _max_over_ranks(a.elapsed_time(b) / n, d_, dev)
        entry = {"ms_per_frame": ms, "rays_per_s": H * W / (ms * 1e-3)}
        if world > 1:       # the gathered frame equals the frame one GPU renders alone (rank 0 checks, bit for bit)
            solo = renderer.render_img(npc, wl.dec, c2w, dev, stage, gt_depth=dmap, npc_geo_feats=gf, npc_col_feats=cf,
                                       dynamic_r_query={k: v.clone() for k, v in rad.items()}, cloud_pos=pos, exposure_feat=ef)
            entry["equals_single_gpu_frame"] = bool(all(torch.equal(x, y) for x, y in zip(out, solo)))
        res[stage] = entry
    # ---- neighbour search only: the frame's 1,536,000 sample points, dealt out by rows, against 1M .. 8M points ----
    r0, r1 = par.shard_bounds(H, rank, world)
    vv, uu = torch.meshgrid(torch.arange(r0, r1, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    dirs = torch.stack([(uu - K["cx"]) / K["fx"], -(vv - K["cy"]) / K["fy"], -torch.ones_like(uu)], -1).reshape(-1, 3).to(dev)
    z = S.synthetic_depth()[r0:r1].to(dev).reshape(-1, 1) * torch.linspace(0.96, 1.04, 5, device=dev)[None]
    q = (dirs[:, None, :] * z[..., None]).reshape(-1, 3).contiguous()
    radq = rad["fine"][r0:r1].reshape(-1).repeat_interleave(5)
    t_up, t_dn = hps.ops.radius_thresholds(radq, q.shape[0], dev)
    del npc, gf, cf, pos
    torch.cuda.empty_cache()
    sweep = {}
    for n_m in (1, 2, 4, 8):
        pts, _ = _cloud_on_device(n_m * 1_000_000, 7, bench.EXTENT * n_m ** 0.5, dev)
        idx = hps.GridIndex(dev)
        idx.reset(pts)
        idx.cell_hint = 0.08
        grid = idx._ensure_grid(None)
        grid.query(q, t_up, t_dn)
        if d_ is not None:
            d_.barrier()
        ms = _max_over_ranks(_timed(lambda: grid.query(q, t_up, t_dn), 5, warm=1), d_, dev)
        sweep["%dM" % n_m] = {"ms": ms, "queries_per_s": H * W * 5 / (ms * 1e-3)}
        del idx, grid, pts
    res["knn_sweep"] = dict(sweep, queries=H * W * 5, note="fine radii 0.04-0.16 m, queries dealt out by image rows")
    return res


# ----------------------------------------------------------------------------------------------------------------
# configs[3], strong form: one 10,000-ray batch sharded over the ranks
# ----------------------------------------------------------------------------------------------------------------
def strong_scaling(hps, wl, bench, rank, world, dist, steps):
    dev, npc, dec, renderer = wl.dev, wl.npc, wl.dec, wl.renderer
    total = wl.cfg["mapping"]["pixels"]                       # 10,000 (configs/ScanNet/scannet.yaml:34, Mapper.py:380)
    fine, scene = wl.scene["fine"], wl.scene
    lo, hi = hps.parallel.shard_bounds(total, rank, world)
    n_b = (fine["rays_o"].shape[0]) // total
    keys = ("rays_o", "rays_d", "gt_depth", "gt_color", "radius")
    batches = []
    for b in range(n_b):
        sl = slice(b * total + lo, b * total + hi)
        d = {k: fine[k][sl].to(dev).contiguous() for k in keys}
        d["radius_mid"] = scene["radius_mid"][sl].to(dev).contiguous()
        batches.append(d)
    static = {k: torch.empty_like(v) for k, v in batches[0].items()}
    ef = scene["exposure_feat"].to(dev)
    targets = [wl.lg.param, wl.lc.param] + list(dec.parameters())
    prep = {"p": None}

    def head():
        prep["p"] = renderer.prepare_batch(npc, static["rays_d"], static["rays_o"], bench.STAGE, static["gt_depth"],
                                           {"fine": static["radius"], "mid": static["radius_mid"]}, out=prep["p"])
        return prep["p"].cnt

    peer = getattr(wl, "peer", None) if world > 1 else None

    def rest():
        for t in targets:
            t.grad = None
        loss = renderer.render_batch_ray_loss(npc, dec, static["rays_d"], static["rays_o"], dev, bench.STAGE, static["gt_depth"],
                                              static["gt_color"], 0.1, npc_geo_feats=wl.gf, npc_col_feats=wl.cf, cloud_pos=wl.pos,
                                              dynamic_r_query={"fine": static["radius"], "mid": static["radius_mid"]},
                                              exposure_feat=ef, prepared=prep["p"],
                                              grad_buffer=peer.pool if peer is not None else None,
                                              after_grads=peer)[0]     # this package's peer-memory all-reduce, in the graph
        loss.backward()
        return loss

    for k, v in batches[0].items():
        static[k].copy_(v)
    g_head = hps.graphs.GraphedStep(head, {}, warmup=3)
    g_rest = hps.graphs.GraphedStep(rest, {}, warmup=3)
    ex = hps.parallel.GradExchange(targets, skip_missing=True)
    comm = torch.cuda.Stream()
    state = {"ev": None}

    def one(i):
        for k, v in batches[i % n_b].items():
            static[k].copy_(v, non_blocking=True)
        g_head.graph.replay()
        if state["ev"] is not None:
            torch.cuda.current_stream().wait_event(state["ev"])
            state["ev"] = None
        g_rest.graph.replay()
        if world > 1 and peer is None:
            comm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(comm):
                ex(dist)
                state["ev"] = comm.record_event()

    for i in range(3):
        one(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = _events()
    a.record()
    for i in range(steps):
        one(3 + i)
    if state["ev"] is not None:
        torch.cuda.current_stream().wait_event(state["ev"])
        state["ev"] = None
    b.record()
    torch.cuda.synchronize()
    ms = _max_over_ranks(a.elapsed_time(b) / steps, dist if world > 1 else None, dev)
    return {"global_rays": total, "rays_per_gpu": hi - lo, "n_gpus": world, "ms_per_step": ms, "value": total / (ms * 1e-3),
            "unit": "rays/s", "steps": steps, "scaling": "strong",
            "note": "ONE reference-sized batch (mapping.pixels = 10,000, Mapper.py:380) sharded contiguously over the ranks "
                    "(SURVEY 8e), gradients all-reduced (%s); back to back, no L2 flush; head | rest graphs"
                    % ("hps_peer_allreduce inside the rest graph" if peer is not None else "NCCL beside the next head")}


def mapping_with_optimizer(hps, wl, bench, steps):
    """The bench's mapper iteration INCLUDING the optimiser step (Mapper.py:783-860: zero_grad, render, loss, backward,
    optimizer.step with the stage's learning rates) - MapperOptimizer: RowAdam over the live rows of both fine tables fused
    with their write-back + one FusedAdam launch for decoders and exposure latent - captured like the bench's step (the next
    batch's head beside the backward tail), replayed back to back.  Parameters are restored afterwards."""
    dev, npc, dec, renderer = wl.dev, wl.npc, wl.dec, wl.renderer
    used = list(dec.geo_decoder_fine.parameters()) + list(dec.color_decoder_fine.parameters())
    ef = wl.scene["exposure_feat"].to(dev).clone().requires_grad_(True)
    owned = [wl.lg.param, wl.lc.param] + used
    keep = [t.detach().clone() for t in owned]
    mo = hps.MapperOptimizer(wl.cfg["mapping"], decoder_params=used, rows={"geometry_fine": wl.lg, "color_fine": wl.lc},
                             exposure_feat=ef)
    mo.set_learning_rates(hps.mapper_optim.stage_learning_rates(wl.cfg["mapping"], bench.STAGE))
    slots = [bench._batch_on_device(wl, b) for b in (0, 1)]
    prep = [None, None]

    def head(k):
        b = slots[k]
        prep[k] = renderer.prepare_batch(npc, b["rays_d"], b["rays_o"], bench.STAGE, b["gt_depth"],
                                         {"fine": b["radius"], "mid": b["radius_mid"]}, out=prep[k])
        return prep[k].cnt

    def step(k):
        mo.zero_grad()
        b = slots[k]
        loss = renderer.render_batch_ray_loss(
            npc, dec, b["rays_d"], b["rays_o"], dev, bench.STAGE, b["gt_depth"], b["gt_color"], 0.1, npc_geo_feats=wl.gf,
            npc_col_feats=wl.cf, cloud_pos=wl.pos, dynamic_r_query={"fine": b["radius"], "mid": b["radius_mid"]},
            exposure_feat=ef, prepared=prep[k], tail_work=lambda: head(1 - k))[0]
        loss.backward()
        mo.step()
        return loss

    try:
        head(0)
        head(1)
        graphs = [hps.graphs.GraphedStep(lambda k=k: step(k), {}, warmup=3) for k in (0, 1)]
        for i in range(4):
            graphs[i % 2].graph.replay()
        torch.cuda.synchronize()
        first = float(graphs[0].outputs.detach())      # (batch 0 after 5 optimiser steps on alternating batches)
        a, b = _events()
        a.record()
        for i in range(steps):
            graphs[i % 2].graph.replay()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / steps
        last = float(graphs[(steps - 1) % 2].outputs.detach())
    finally:
        with torch.no_grad():
            for t, k in zip(owned, keep):
                t.copy_(k)
                t.grad = None
        wl.lg.fresh = wl.lc.fresh = False          # the next render writes the restored leaves back into the tables
    return {"ms_per_step": ms, "value": bench.RAYS / (ms * 1e-3), "unit": "rays/s", "steps": steps,
            "iterations_per_s": 1e3 / ms, "loss_batch0_after_5_steps": first, "loss_last": last,
            "note": "zero_grad + render + loss + backward + MapperOptimizer.step() (RowAdam on %d live rows x 2 tables fused "
                    "with the table write-back, FusedAdam on the fine decoders + exposure latent) from one CUDA graph per "
                    "iteration, back to back, no L2 flush: compare with `sustained`" % int(wl.live_idx.numel())}


def run(hps, wl, args, rank, world, dist):
    bench = importlib.import_module("bench")
    out = {}
    only = set(filter(None, os.environ.get("HPS_BENCH_EXTRAS", "").split(",")))

    def want(name):
        return not only or name in only

    if world == 1 and want("tracking"):
        try:
            out["tracking"] = tracking(hps, wl, bench)
        except Exception as e:                                           # noqa: BLE001
            out["tracking"] = _err(e)
    if world == 1 and want("single_frame"):
        try:
            out["single_frame"] = single_frame(hps, wl, bench)
        except Exception as e:                                           # noqa: BLE001
            out["single_frame"] = _err(e)
    if world == 1 and want("mapping_with_optimizer"):
        try:
            out["mapping_with_optimizer"] = mapping_with_optimizer(hps, wl, bench, max(20, min(4 * args.steps, 400)))
        except Exception as e:                                           # noqa: BLE001
            out["mapping_with_optimizer"] = _err(e)
    if want("strong_scaling"):
        try:
            out["strong_scaling"] = strong_scaling(hps, wl, bench, rank, world, dist, max(10, min(args.steps, 100)))
        except Exception as e:                                           # noqa: BLE001
            out["strong_scaling"] = _err(e)
    if want("full_frame"):
        try:
            torch.cuda.empty_cache()
            out["full_frame"] = full_frame(hps, wl, bench, rank, world, dist)
        except Exception as e:                                           # noqa: BLE001
            out["full_frame"] = _err(e)
        torch.cuda.empty_cache()
    return out
