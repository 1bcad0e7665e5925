"""kNN query time, one warp per query (group=1) vs one warp per ray (group=5), over the query count, fine and mid level,
with mean candidates examined (GPU box)."""
import sys
import torch
sys.path.insert(0, ".")
import bench
import hps_b200 as hps

dev = "cuda:0"
scene = bench.make_scene()
idx = hps.GridIndex(dev)
idx.reset(scene["fine"]["points"].to(dev))
fine = scene["fine"]
for R in (1000, 5000, 20000, 40000, 307200):
    reps = (R + bench.RAYS * bench.N_BATCHES - 1) // (bench.RAYS * bench.N_BATCHES)
    ro = fine["rays_o"].repeat(reps, 1)[:R].to(dev)
    rd = fine["rays_d"].repeat(reps, 1)[:R].to(dev)
    gd = fine["gt_depth"].repeat(reps)[:R].to(dev)
    rad = fine["radius"].repeat(reps)[:R].to(dev)
    z = gd[:, None] * torch.linspace(0.96, 1.04, 5, device=dev)[None]
    q = (ro[:, None] + rd[:, None] * z[..., None]).reshape(-1, 3).contiguous()
    t_up, t_dn = hps.ops.radius_thresholds(rad.repeat_interleave(5), q.shape[0], dev)
    grid = idx._ensure_grid(t_dn)
    for group in (1, 5):
        for _ in range(3):
            grid.query(q, t_up, t_dn, group=group)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(20):
            grid.query(q, t_up, t_dn, group=group)
        b.record()
        torch.cuda.synchronize()
        cand = grid.query(q, t_up, t_dn, want_cand=True, group=group)[3]
        print("fine: rays %6d queries %7d group %d: %.1f us, mean candidates %.0f" % (R, q.shape[0], group, a.elapsed_time(b) * 1000 / 20,
                                                                                   float(cand.float().mean())))
# mid level: 250k points, radii 0.24 - 0.96 m
idx = hps.GridIndex(dev)
idx.reset(scene["mid"]["points"].to(dev))
idx.cell_hint = 0.48
R = 5000
ro, rd, gd = fine["rays_o"][:R].to(dev), fine["rays_d"][:R].to(dev), fine["gt_depth"][:R].to(dev)
z = gd[:, None] * torch.linspace(0.96, 1.04, 5, device=dev)[None]
q = (ro[:, None] + rd[:, None] * z[..., None]).reshape(-1, 3).contiguous()
t_up, t_dn = hps.ops.radius_thresholds(scene["radius_mid"][:R].to(dev).repeat_interleave(5), q.shape[0], dev)
grid = idx._ensure_grid(None)
for group in (1, 5):
    for _ in range(3):
        grid.query(q, t_up, t_dn, group=group)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(20):
        grid.query(q, t_up, t_dn, group=group)
    b.record()
    torch.cuda.synchronize()
    cand = grid.query(q, t_up, t_dn, want_cand=True, group=group)[3]
    print("mid : rays %6d queries %7d group %d: %.1f us, mean candidates %.0f" % (R, q.shape[0], group, a.elapsed_time(b) * 1000 / 20,
                                                                               float(cand.float().mean())))
