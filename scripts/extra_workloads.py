"""SURVEY 8(d) workloads beyond the bench line (configs 3 and 5), measured on one B200:
   (a) tracker iterations (R = 5000, pose + exposure latent optimised with Adam) - iterations / s, eager and graphed
   (b) full-frame 640x480 render (307,200 rays) through render_img, per level, at two ray batch sizes
   (c) neighbour-search-only sweep over cloud sizes 1M .. 8M points (1.536M queries = one full frame x 5 samples)
usage (GPU box): python scripts/extra_workloads.py > gpurun_out/extra_workloads.txt"""
import importlib
import os
import sys
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import synthetic as O  # noqa: E402  (scene generator; the oracle is not used here)

hps = importlib.import_module("hierarchical-point-slam_b200")
dev = "cuda:0"
SECTIONS = set(sys.argv[1:]) or {"a", "a2", "b", "c"}      # e.g. `python scripts/extra_workloads.py a2`
K = O.SCANNET_INTRINSICS


def timed(fn, n, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n, (time.perf_counter() - t0) * 1e3 / n


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
pos = {l: npc.cloud_pos_tensor(l) for l in ("fine", "mid")}
gf = {l: npc.get_geo_feats(l) for l in ("fine", "mid")}
cf = {l: npc.get_col_feats(l) for l in ("fine", "mid")}

# ------------------------------------------------------------------ (a) tracker iterations
print("== (a) tracker iteration, R = %d rays, pose (7) + exposure latent (8) under Adam, 1M fine / 250k mid points" % R)
rd0 = fine["rays_d"][:R]
pi = (rd0[:, 0] * K["fx"] + K["cx"]).to(dev)
pj = (K["cy"] - rd0[:, 1] * K["fy"]).to(dev)
gt_d, gt_c = fine["gt_depth"][:R].to(dev), fine["gt_color"][:R].to(dev)
rq_t = {"fine": fine["radius"][:R].to(dev), "mid": scene["radius_mid"][:R].to(dev)}
for prm in dec.parameters():          # the tracker optimises the pose and the exposure latent only (Tracker.py:190-235)
    prm.requires_grad_(False)
for stage in (("color_mid", "color_fine") if "a" in SECTIONS else ()):
    cam = torch.tensor([1.0, 0.0, 0.0, 0.0, 0.005, -0.004, 0.003], device=dev).requires_grad_(True)   # ~1 cm off
    ef = scene["exposure_feat"].to(dev).clone().requires_grad_(True)
    opt = torch.optim.Adam([cam, ef], lr=5e-4, capturable=True)

    def it():
        opt.zero_grad(set_to_none=False)
        c2w = hps.get_camera_from_tensor(cam)
        ro, rd = hps.get_rays_from_uv(pi, pj, c2w, K["H"], K["W"], K["fx"], K["fy"], K["cx"], K["cy"], dev)
        d, v, c, _ = renderer.render_batch_ray(npc, dec, rd, ro, dev, stage, gt_depth=gt_d, npc_geo_feats=gf,
                                               npc_col_feats=cf, is_tracker=True, cloud_pos=pos,
                                               dynamic_r_query=dict(rq_t), exposure_feat=ef)
        loss = hps.tracker_loss(d, v, c, gt_d, gt_c, 0.5)
        loss.backward()
        opt.step()
        return loss

    cam.grad, ef.grad = torch.zeros_like(cam), torch.zeros_like(ef)
    g_ms, w_ms = timed(it, 30)
    print("  %-10s eager : %.3f ms / iteration (GPU events), %.3f ms wall  -> %.0f iterations/s" % (stage, g_ms, w_ms, 1e3 / w_ms))
    try:
        graphed = hps.graphs.GraphedStep(it, {}, warmup=3)
        g_ms, w_ms = timed(lambda: graphed(), 100)
        print("  %-10s graph : %.3f ms / iteration (GPU events), %.3f ms wall  -> %.0f iterations/s  (loss %.4f)"
              % (stage, g_ms, w_ms, 1e3 / w_ms, float(graphed.outputs)))
        if os.environ.get("HPS_TIMELINE") and stage == "color_fine":
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                graphed()
                torch.cuda.synchronize()
            evs = sorted((e.time_range.start, e.time_range.end, e.name) for e in prof.events()
                         if e.device_type == torch.autograd.DeviceType.CUDA)
            for s0, s1, name in evs:
                print("      %8.1f %7.1f  %s" % (s0 - evs[0][0], s1 - s0, name[:90]))
    except Exception as e:   # noqa: BLE001
        print("  %-10s graph capture failed: %s" % (stage, str(e).splitlines()[0][:150]))

# ------------------------------------------------------------------ (a2) mapper iteration including the optimiser
print("== (a2) mapper iteration INCLUDING the optimiser step, stage color_fine, fused render+loss call")
for prm in dec.parameters():
    prm.requires_grad_(True)
Kc = cfg["cam"]
used = list(dec.geo_decoder_fine.parameters()) + list(dec.color_decoder_fine.parameters())
ro_m, rd_m = fine["rays_o"][:R].to(dev), fine["rays_d"][:R].to(dev)
live_idx = npc.frustum_rows("fine", torch.eye(4, device=dev), O.synthetic_depth().to(dev), Kc["H"], Kc["W"], Kc["fx"],
                            Kc["fy"], Kc["cx"], Kc["cy"], -20)
table0 = (npc.get_geo_feats("fine").clone(), npc.get_col_feats("fine").clone())
dec0 = [p.detach().clone() for p in used]


def mapper_variant(which):
    with torch.no_grad():                         # every variant starts from the same tables and decoders
        npc.get_geo_feats("fine").copy_(table0[0])
        npc.get_col_feats("fine").copy_(table0[1])
        for p, q in zip(used, dec0):
            p.copy_(q)
            p.grad = None
    lg, lc = hps.LiveRows.pair(npc.get_geo_feats("fine"), npc.get_col_feats("fine"), live_idx)
    gfm = {"fine": lg, "mid": npc.get_geo_feats("mid")}
    cfm = {"fine": lc, "mid": npc.get_col_feats("mid")}
    ef_m = scene["exposure_feat"].to(dev).clone().requires_grad_(True)
    if which == "none":
        opts = []
    elif which == "RowAdam + torch fused Adam":
        opts = [hps.RowAdam([(lg, 5e-3), (lc, 5e-3)]),
                torch.optim.Adam([{"params": used, "lr": 5e-3}, {"params": [ef_m], "lr": 1e-3}], fused=True,
                                 capturable=True)]
    else:                                         # everything on this package's kernels (Mapper.py:592-603, 646-660)
        mo = hps.MapperOptimizer(cfg["mapping"], decoder_params=used, rows={"geometry_fine": lg, "color_fine": lc},
                                 exposure_feat=ef_m)
        mo.set_learning_rates(hps.mapper_optim.stage_learning_rates(cfg["mapping"], "color_fine"))
        opts = [mo]

    def mapper_it():
        for o in opts:
            o.zero_grad()
        if not opts:
            for t in used + [lg.param, lc.param, ef_m]:
                t.grad = None
        loss = renderer.render_batch_ray_loss(npc, dec, rd_m, ro_m, dev, "color_fine", gt_d, gt_c, 0.1, npc_geo_feats=gfm,
                                              npc_col_feats=cfm, cloud_pos=pos, dynamic_r_query=dict(rq_t),
                                              exposure_feat=ef_m)[0]
        loss.backward()
        for o in opts:
            o.step()
        return loss

    try:
        gm = hps.graphs.GraphedStep(mapper_it, {}, warmup=3)
        traj = []                                  # loss after replay 1, 2, 5, 10, 20, 50 (iterations 4 .. 53)
        for k in range(1, 51):
            gm()
            if k in (1, 2, 5, 10, 20, 50):
                traj.append(float(gm.outputs.detach()))
        g_ms, w_ms = timed(lambda: gm(), 200)
        print("  %-28s: %.4f ms / iteration (GPU events) -> %.0f iterations/s = %.2f M rays/s; %d live rows; loss at "
              "iteration 4/5/8/13/23/53: %s" % (which, g_ms, 1e3 / w_ms, R / w_ms / 1e3, live_idx.numel(),
                                                " ".join("%.3f" % t for t in traj)))
    except Exception as e:   # noqa: BLE001
        print("  %s: graph capture failed: %s" % (which, str(e).splitlines()[0][:200]))


if "a2" in SECTIONS:
    for which in ("none", "RowAdam + torch fused Adam", "MapperOptimizer"):
        mapper_variant(which)

# ------------------------------------------------------------------ (b) full-frame render
print("== (b) full-frame render_img, 640x480 = 307,200 rays x 5 samples, no_grad")
depth_map = O.synthetic_depth().to(dev)
g = torch.Generator().manual_seed(5)
rad_maps = {"fine": (0.04 + 0.12 * torch.rand(K["H"], K["W"], generator=g, dtype=torch.float64)).to(dev),
            "mid": (0.24 + 0.72 * torch.rand(K["H"], K["W"], generator=g, dtype=torch.float64)).to(dev)}
c2w = torch.eye(4, device=dev)
ef0 = scene["exposure_feat"].to(dev)
for B in ((3000, 65536) if "b" in SECTIONS else ()):
    renderer.ray_batch_size = B
    renderer.min_render_chunk = B        # (the default widens chunks to 65536 rays; 3000 = the reference's chunking)
    for stage in ("color_mid", "color_fine"):
        def frame():
            return renderer.render_img(npc, dec, c2w, dev, stage, gt_depth=depth_map, npc_geo_feats=gf, npc_col_feats=cf,
                                       dynamic_r_query={k: v.clone() for k, v in rad_maps.items()}, cloud_pos=pos,
                                       exposure_feat=ef0)
        g_ms, w_ms = timed(frame, 3, warm=1)
        print("  ray batch %6d %-10s: %.2f ms / frame (wall %.2f ms) -> %.1f M rays/s" % (B, stage, g_ms, w_ms, 307200 / w_ms / 1e3))

# ------------------------------------------------------------------ (c) neighbour-search sweep
print("== (c) neighbour search only: 1,536,000 queries (full frame x 5 samples), fine radii, cloud size sweep")
H, W = K["H"], K["W"]
vv, uu = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
dirs = torch.stack([(uu - K["cx"]) / K["fx"], -(vv - K["cy"]) / K["fy"], -torch.ones_like(uu)], -1).reshape(-1, 3).to(dev)
z = depth_map.reshape(-1, 1) * torch.linspace(0.96, 1.04, 5, device=dev)[None]
q = (dirs[:, None, :] * z[..., None]).reshape(-1, 3).contiguous()
rad = rad_maps["fine"].reshape(-1).repeat_interleave(5)
t_up, t_dn = hps.ops.radius_thresholds(rad, q.shape[0], dev)
del npc, gf, cf, pos
torch.cuda.empty_cache()
for n_m in ((1, 2, 4, 8) if "c" in SECTIONS else ()):
    sc = O.synthetic_scene(n_m * 1_000_000, 8, (0.04, 0.16), seed=7, extent_scale=bench.EXTENT * n_m ** 0.5)
    idx = hps.GridIndex(dev)
    pts = sc["points"].to(dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    idx.reset(pts)
    grid = idx._ensure_grid(t_dn)
    torch.cuda.synchronize()
    build_ms = (time.perf_counter() - t0) * 1e3
    for group in (1, 5):
        g_ms, _ = timed(lambda: grid.query(q, t_up, t_dn, group=group), 5, warm=2)
        res = grid.query(q, t_up, t_dn, want_cand=True, group=group)
        print("  N = %dM  group %d: build %.1f ms, query %.2f ms -> %.0f M queries/s, mean candidates %.0f, mean in-radius %.1f"
              % (n_m, group, build_ms, g_ms, q.shape[0] / g_ms / 1e3, float(res[3].float().mean()), float(res[2].float().mean())))
    del idx, grid, pts
    torch.cuda.empty_cache()
