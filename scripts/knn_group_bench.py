"""kNN query time, per-query warps (group=1) vs ray-grouped warps (group=5), over the query count (GPU box)."""
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
for R in (5000, 20000, 40000):
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
        print("rays %6d queries %7d group %d: %.1f us" % (R, q.shape[0], group, a.elapsed_time(b) * 1000 / 20))
