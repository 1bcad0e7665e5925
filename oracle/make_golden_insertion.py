"""TEST INFRASTRUCTURE ONLY.  Generate tests/golden/insertion_small.pt by running the UNMODIFIED reference class
`NeuralPointCloud.add_neural_points` (src/neural_point.py:193-334), imported from /root/reference under the stubs of
oracle/ref_import.py (the faiss index is the exhaustive stand-in = the reference forced to exhaustive search), on the CPU:
seven consecutive insertions into two levels - first frame (no density test), a frame that repeats the first one's rays
plus new ones (density test with per-ray float64 radii), scalar radius_add with a missing-depth ray, radius_min
(is_pts_grad), and the mid level with its larger radii.  Stored per call: the returned count and the level's cloud.

Run in the build container only:   python oracle/make_golden_insertion.py
The fixture is committed; the GPU suite replays the calls on the product class and compares counts and positions.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
import synthetic  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    ref = ref_import.load_reference()
    cfg = ref_import.load_cfg(ref)
    npc = ref.neural_point.NeuralPointCloud(cfg)
    sc = synthetic.synthetic_scene(16, 900, (0.02, 0.08), seed=31)
    g = torch.Generator().manual_seed(32)
    ro, rd, gd, gc = sc["rays_o"], sc["rays_d"], sc["gt_depth"].clone(), sc["gt_color"]
    r_fine = sc["radius"]                                              # float64, 0.02 .. 0.08 (fine r_add range)
    r_mid = 0.12 + 0.36 * torch.rand(900, generator=g, dtype=torch.float64)
    gd_zero = gd.clone()
    gd_zero[650] = 0.0
    calls = [
        dict(level="fine", idx=0, rays=(0, 300), dynamic=True, is_pts_grad=False, zero=False),
        dict(level="fine", idx=1, rays=(0, 600), dynamic=True, is_pts_grad=False, zero=False),
        dict(level="fine", idx=2, rays=(600, 800), dynamic=False, is_pts_grad=False, zero=True),
        dict(level="fine", idx=3, rays=(700, 900), dynamic=False, is_pts_grad=True, zero=False),
        dict(level="mid", idx=0, rays=(0, 450), dynamic=True, is_pts_grad=False, zero=False),
        dict(level="mid", idx=1, rays=(300, 900), dynamic=True, is_pts_grad=False, zero=False),     # nothing accepted
        dict(level="mid", idx=2, rays=(300, 900), dynamic=True, is_pts_grad=False, zero=False, scale=0.25),
    ]
    for c in calls:
        a, b = c["rays"]
        depth = (gd_zero if c["zero"] else gd)[a:b]
        rad = (r_fine if c["level"] == "fine" else r_mid)[a:b] * c.get("scale", 1.0) if c["dynamic"] else None
        n = npc.add_neural_points(ro[a:b], rd[a:b], depth, gc[a:b], train=False, is_pts_grad=c["is_pts_grad"],
                                  dynamic_radius=rad, level=c["level"], idx=c["idx"])
        c["count"] = int(n)
        c["cloud"] = torch.tensor(npc.cloud_pos(c["level"]), dtype=torch.float32).reshape(-1, 3)
        c["n_input_pos"] = len(npc.input_pos())
        assert npc.index[c["level"]].ntotal == c["cloud"].shape[0] == npc.pts_num()[c["level"]]
        assert npc.get_geo_feats(c["level"]).shape == (c["cloud"].shape[0], 32)
        print("%-4s idx %d rays %3d..%3d dynamic=%-5s is_pts_grad=%-5s -> accepted %3d, cloud %d"
              % (c["level"], c["idx"], a, b, c["dynamic"], c["is_pts_grad"], c["count"], c["cloud"].shape[0]))
    out = dict(rays_o=ro, rays_d=rd, gt_depth=gd, gt_depth_zero=gd_zero, gt_color=gc, r_fine=r_fine, r_mid=r_mid,
               calls=calls, radius_add=cfg["pointcloud"]["radius_add"], radius_min=cfg["pointcloud"]["radius_min"])
    path = os.path.join(GOLD, "insertion_small.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
