"""TEST INFRASTRUCTURE ONLY.  Generate tests/golden/near_pcl_small.pt: NeuralPointCloud.sample_near_pcl
(neural_point.py:443-495 - the depth-less-ray sampler behind rendering.sample_near_pcl, default True in
configs/point_slam.yaml:168, False in configs/ScanNet/scannet.yaml:43), run UNMODIFIED (imported under
oracle/ref_import.py stubs, exhaustive index) on the committed small scene with a third of the rays stripped of their
depth reading; fixed query radius and a per-probe radius tensor.

Run in the build container only:   python oracle/make_golden_near_pcl.py

Only the function is pinned.  Its call site (Renderer.py:196-206) cannot run in the reference: it hands the
dynamic_r_query DICT to find_neighbors_faiss, whose shape assert (neural_point.py:411) needs a tensor, and the decoder
(decoder.py:204) needs the dict - so the first depth-less ray raises.  The product gives the call site its evident
meaning (per-ray radius repeated per probe) and tests it against its own seams.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
from oracle import make_golden as MG  # noqa: E402

RADIUS_QUERY = {"fine": 0.045, "mid": 0.2}


def main():
    ns = ref_import.load_reference()
    scene = torch.load(os.path.join(MG.GOLD, "scene_small.pt"), weights_only=False)
    clouds, rays = scene["clouds"], scene["rays"]
    gt_depth = rays["gt_depth"].clone()
    gt_depth[::3] = 0.0
    rays_d = rays["rays_d"].clone()
    rays_d[0] = torch.tensor([0.0, 1.0, 0.2])          # a depth-less ray that leaves the cloud: not near the cloud
    out = dict(gt_depth=gt_depth, rays_d=rays_d, radius_query=RADIUS_QUERY, stages={})
    for stage in ("geometry_mid", "color_fine"):
        level = "mid" if stage.endswith("mid") else "fine"
        cfg = ref_import.load_cfg(ns)
        cfg["use_dynamic_radius"] = False
        cfg["rendering"]["sample_near_pcl"] = True
        cfg["pointcloud"]["radius_query"] = RADIUS_QUERY[level]
        npc, decoders, renderer = MG.build_reference(ns, cfg, clouds)
        for name, sd in scene["decoders"].items():
            m = getattr(decoders, name)
            sd = dict(sd)
            B = sd.pop("embedder._B")
            if name.startswith("geo"):
                sd["embedder._B"] = B
            else:
                m.embedder._B = B.clone()
            m.load_state_dict(sd, strict=True)
        none = ~(gt_depth > 0)
        far = torch.clamp(torch.minimum(5 * gt_depth.mean(), torch.max(gt_depth * 1.2)), 0, torch.max(gt_depth * 1.2))
        z, invalid = npc.sample_near_pcl(rays["rays_o"][none].clone(), rays_d[none].clone(), renderer.near_end, far,
                                         renderer.N_surface, None, level)
        g = torch.Generator().manual_seed(3)
        lo, hi = MG.LEVEL_RADIUS[level]
        r_probe = lo + (hi - lo) * torch.rand(int(none.sum()) * 25, generator=g, dtype=torch.float64)
        z_dyn, invalid_dyn = npc.sample_near_pcl(rays["rays_o"][none].clone(), rays_d[none].clone(), renderer.near_end,
                                                 far, renderer.N_surface, r_probe, level)
        out["stages"][stage] = dict(z=z, invalid=invalid, near=renderer.near_end, far=far, r_probe=r_probe,
                                    z_dyn=z_dyn, invalid_dyn=invalid_dyn)
        default = torch.linspace(renderer.near_end, float(far), renderer.N_surface)
        print("%s: %d depth-less rays; fixed radius: %d not near the cloud, %d rows moved; per-probe radius: %d / %d"
              % (stage, int(none.sum()), int(invalid.sum()), int(((z - default).abs() > 1e-6).any(1).sum()),
                 int(invalid_dyn.sum()), int(((z_dyn - default).abs() > 1e-6).any(1).sum())))
    torch.save(out, os.path.join(MG.GOLD, "near_pcl_small.pt"))
    print("wrote near_pcl_small.pt")


if __name__ == "__main__":
    main()
