"""TEST INFRASTRUCTURE ONLY.  Generate tests/golden/render_img_small.pt by running the UNMODIFIED reference
`Renderer.render_img` (src/utils/Renderer.py:256-337: get_rays for every pixel, render_batch_ray chunk by chunk under
no_grad, float64 depth / uncertainty images), imported from /root/reference under the stubs of oracle/ref_import.py, on
the CPU: the cloud and decoders of tests/golden/scene_small.pt, a 24 x 32 frame with a perturbed pose, per-pixel float64
radius maps, ~8 % pixels without a depth reading (their sampling range depends on the 100-ray chunk they are rendered
in, Renderer.py:133-134, 148-153, 207-208), stages color_mid and color_fine with an exposure latent.

Run in the build container only:   python oracle/make_golden_render_img.py    (after oracle/make_golden.py)
The fixture is committed; the GPU suite renders the same frame through the product's render_img, with the reference's
chunking and with its own wide chunks.
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import hps_oracle as O  # noqa: E402
from oracle import make_golden as MG  # noqa: E402
from oracle import ref_import  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    ns = ref_import.load_reference()
    cfg = ref_import.load_cfg(ns)
    scene = torch.load(os.path.join(GOLD, "scene_small.pt"), weights_only=False)
    clouds = scene["clouds"]
    npc, decoders, _ = MG.build_reference(ns, cfg, clouds)
    for name, sd in scene["decoders"].items():                 # the fixture's decoder state (incl. the colour _B)
        m = getattr(decoders, name)
        sd = dict(sd)
        B = sd.pop("embedder._B")
        if name.startswith("geo"):
            sd["embedder._B"] = B
        else:
            m.embedder._B = B.clone()
        m.load_state_dict(sd, strict=True)
    K = O.SCANNET_INTRINSICS
    H, W, sc = 24, 32, 20
    slam = types.SimpleNamespace(nice=True, H=H, W=W, fx=K["fx"] / sc, fy=K["fy"] / sc, cx=K["cx"] / sc, cy=K["cy"] / sc)
    renderer = ns.renderer.Renderer(cfg, None, slam, ray_batch_size=100)
    renderer.sigmoid_coefficient = scene["cfg"]["coef"]
    g = torch.Generator().manual_seed(77)
    depth = O.synthetic_depth()[::sc, ::sc].contiguous().clone()
    depth[torch.rand(H, W, generator=g) < 0.08] = 0.0
    depth[5, :] = 0.0                                           # most of one chunk without a reading
    cam = torch.tensor([0.9995, 0.01, -0.02, 0.015, 0.03, -0.02, 0.025])
    c2w = O.camera_from_tensor(cam)                             # (3,4)
    radius = {"fine": 0.04 + 0.12 * torch.rand(H, W, generator=g, dtype=torch.float64),
              "mid": 0.24 + 0.72 * torch.rand(H, W, generator=g, dtype=torch.float64)}
    cloud_pos = {l: clouds[l]["points"] for l in clouds}
    out = dict(H=H, W=W, intrinsics={k: getattr(slam, k) for k in ("fx", "fy", "cx", "cy")}, depth=depth, c2w=c2w,
               radius=radius, ray_batch_size=100, stages={})
    n_chunks = (H * W + 99) // 100
    for stage in ("color_mid", "color_fine"):
        with MG.pinned_reference_rng([scene["c_fill_geo"], scene["c_fill_col"]] * n_chunks):
            d, u, c = renderer.render_img(npc, decoders, c2w, "cpu", stage, gt_depth=depth.clone(),
                                          npc_geo_feats={l: clouds[l]["geo_feats"] for l in clouds},
                                          npc_col_feats={l: clouds[l]["col_feats"] for l in clouds},
                                          dynamic_r_query={l: radius[l].clone() for l in radius}, cloud_pos=cloud_pos,
                                          exposure_feat=scene["exposure_feat"])
        assert d.dtype == torch.float64 and u.dtype == torch.float64 and d.shape == (H, W) and c.shape == (H, W, 3)
        out["stages"][stage] = dict(depth=d, uncertainty=u, color=c)
        print("%s: depth %.4f..%.4f, %d pixels without a reading (depth 0 there: %s), colour mean %.4f"
              % (stage, float(d[d > 0].min()), float(d.max()), int((depth == 0).sum()),
                 bool((d[depth == 0] == 0).all()), float(c.mean())))
    path = os.path.join(GOLD, "render_img_small.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
