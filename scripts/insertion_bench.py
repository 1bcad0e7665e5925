"""Cost of one point insertion (add_neural_points: density test + append of positions, feature rows and grid entries)
against clouds of 1M and 8M points: in-place append vs the round-1 path (torch.cat of three O(N) tensors + full grid
rebuild, forced here with HPS-internal switches).  usage (GPU box): python scripts/insertion_bench.py"""
import importlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import bench_extras  # noqa: E402
import synthetic as S  # noqa: E402

hps = importlib.import_module("hierarchical-point-slam_b200")
dev = "cuda:0"
for n_m in (1, 8):
    n = n_m * 1_000_000
    pts, g = bench_extras._cloud_on_device(n, 7, bench.EXTENT * n_m ** 0.5, dev)
    npc = hps.NeuralPointCloud(hps.scannet_cfg(dev))
    npc.set_cloud("fine", pts, torch.randn(n, 32, generator=g, device=dev) * 0.1, torch.randn(n, 32, generator=g, device=dev) * 0.1)
    sc = S.synthetic_scene(8, 30000, (0.04, 0.16), seed=11)
    # rays looking at a patch beside the cloud: ~all of them add points
    ro = sc["rays_o"].to(dev) + torch.tensor([bench.EXTENT * n_m ** 0.5 * 1.5, 0.0, 0.0], device=dev)
    rd, gd, gc = sc["rays_d"].to(dev), sc["gt_depth"].to(dev), sc["gt_color"].to(dev)
    npc.find_neighbors_faiss(pts[:10], step="query", level="fine")            # grid built
    index = npc.get_index("fine")
    times = []
    for k in range(6):
        sl = slice(5000 * k, 5000 * (k + 1))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        added = int(npc.add_neural_points(ro[sl] + torch.tensor([0.0, 3.0 * k, 0.0], device=dev), rd[sl], gd[sl], gc[sl], level="fine", idx=k + 1))
        npc.find_neighbors_faiss(pts[:10], step="query", level="fine")        # brings the grid up to date
        torch.cuda.synchronize()
        times.append((time.perf_counter() - t0) * 1e3)
    print("N = %dM: 5000-ray insertion + grid update: %s ms (first call grows the capacity buffers); %d locations x 3 points "
          "in the last call; grid appends %d, rebuilds %d" % (n_m, " ".join("%.2f" % t for t in times), added, index.appends, index.rebuilds))
    # the round-1 path for comparison: concatenate the three tensors and rebuild the grid
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    new = torch.randn(15000, 3, device=dev)
    p2 = torch.cat([npc.cloud_pos_tensor("fine"), new], 0)
    g2 = torch.cat([npc.get_geo_feats("fine"), torch.zeros(15000, 32, device=dev)], 0)
    c2 = torch.cat([npc.get_col_feats("fine"), torch.zeros(15000, 32, device=dev)], 0)
    grid = hps.ops.VoxelHashGrid()
    grid.build(p2, 0.08)
    torch.cuda.synchronize()
    print("N = %dM: round-1 path (three torch.cat + full rebuild): %.2f ms" % (n_m, (time.perf_counter() - t0) * 1e3))
    del npc, pts, p2, g2, c2, grid
    torch.cuda.empty_cache()
