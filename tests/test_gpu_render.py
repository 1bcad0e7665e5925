"""-m gpu: the drop-in render_batch_ray + losses + backward against fixtures produced by the imported
reference (tests/golden/, oracle/make_golden.py)."""
import pytest
import torch

from oracle import hps_oracle as O
from tests import parity

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def built(hps):
    scene = parity.load("scene_small.pt")
    return (scene,) + parity.product(scene, DEV)


@pytest.mark.parametrize("stage", ["geometry_mid", "color_mid", "geometry_fine", "color_fine"])
def test_mapper_iteration_matches_reference(built, stage):
    scene, hps, npc, dec, renderer = built
    ref = parity.load("render_mapper_small.pt")[stage]
    out = parity.mapper_step(hps, npc, dec, renderer, scene, stage, DEV)
    errs = parity.check_mapper(stage, out, ref, dec)
    print(stage, errs)


@pytest.mark.parametrize("stage", ["color_mid", "color_fine"])
def test_tracker_iteration_matches_reference(built, stage):
    scene, hps, npc, dec, renderer = built
    ref = parity.load("render_tracker_small.pt")
    out = parity.tracker_step(hps, npc, dec, renderer, scene, ref["inputs"], stage, DEV)
    errs = parity.check_tracker(stage, out, ref[stage], scene, ref["inputs"])
    print(stage, errs)


def test_render_batch_does_not_need_dict_copy(built):
    """Renderer.py:226-229 mutates dynamic_r_query in place; the drop-in keeps that visible behaviour."""
    scene, hps, npc, dec, renderer = built
    cl, rays = scene["clouds"], scene["rays"]
    rq = {l: cl[l]["radius"].to(DEV).clone() for l in cl}
    with torch.no_grad():
        renderer.render_batch_ray(npc, dec, rays["rays_d"].to(DEV), rays["rays_o"].to(DEV), DEV, "color_fine",
                                  gt_depth=rays["gt_depth"].to(DEV),
                                  npc_geo_feats={l: npc.get_geo_feats(l) for l in cl},
                                  npc_col_feats={l: npc.get_col_feats(l) for l in cl}, dynamic_r_query=rq)
    assert rq["fine"].shape == (rays["rays_d"].shape[0] * 5, 1)


def test_render_img_matches_batch_path(built):
    scene, hps, npc, dec, renderer = built
    cl = scene["clouds"]
    K = O.SCANNET_INTRINSICS
    import types
    small = hps.Renderer(hps.scannet_cfg(DEV), None, types.SimpleNamespace(H=24, W=32, fx=K["fx"] / 20, fy=K["fy"] / 20,
                                                                         cx=K["cx"] / 20, cy=K["cy"] / 20),
                         ray_batch_size=100)
    small.sigmoid_coefficient = 0.1
    depth_img = O.synthetic_depth()[::20, ::20].contiguous().to(DEV)
    rq = {l: torch.full((24, 32), 0.16 if l == "fine" else 0.96, dtype=torch.float64, device=DEV) for l in cl}
    c2w = torch.eye(4, device=DEV)[:3]
    d, u, c = small.render_img(npc, dec, c2w, DEV, "color_fine", gt_depth=depth_img,
                               npc_geo_feats={l: npc.get_geo_feats(l) for l in cl},
                               npc_col_feats={l: npc.get_col_feats(l) for l in cl}, dynamic_r_query=rq)
    assert d.shape == (24, 32) and c.shape == (24, 32, 3) and d.dtype == torch.float64
    small.ray_batch_size = 768
    rq = {l: torch.full((24, 32), 0.16 if l == "fine" else 0.96, dtype=torch.float64, device=DEV) for l in cl}
    d2, u2, c2 = small.render_img(npc, dec, c2w, DEV, "color_fine", gt_depth=depth_img,
                                  npc_geo_feats={l: npc.get_geo_feats(l) for l in cl},
                                  npc_col_feats={l: npc.get_col_feats(l) for l in cl}, dynamic_r_query=rq)
    assert torch.equal(d, d2) and torch.equal(c, c2)       # chunking does not change per-ray results


def test_add_neural_points_density_test(built):
    scene, hps, npc0, dec, renderer = built
    cfg = hps.scannet_cfg(DEV)
    npc = hps.NeuralPointCloud(cfg)
    sc = O.synthetic_scene(8, 600, seed=4)
    ro, rd, gd, gc = (sc[k].to(DEV) for k in ("rays_o", "rays_d", "gt_depth", "gt_color"))
    gd[7] = 0.0
    r_add = torch.full((600,), 0.05, dtype=torch.float64, device=DEV)
    n1 = npc.add_neural_points(ro[:300], rd[:300], gd[:300], gc[:300], dynamic_radius=r_add[:300], level="fine", idx=0)
    assert int(n1) == 299 and npc.index_ntotal("fine") == 299 * 3 and npc.get_geo_feats("fine").shape == (897, 32)
    n2 = npc.add_neural_points(ro, rd, gd, gc, dynamic_radius=r_add, level="fine", idx=1)
    # oracle: a location is accepted iff no existing point lies within r_add (neural_point.py:262-268)
    pts_gt = (sc["rays_o"] + sc["rays_d"] * sc["gt_depth"][:, None])
    keep = torch.ones(600, dtype=torch.bool)
    keep[7] = False
    existing = npc.cloud_pos_tensor("fine")[:897].cpu()
    D, _ = O.exhaustive_knn(existing, pts_gt[keep], 8)
    want = int((D[:, 0] >= torch.tensor(0.05, dtype=torch.float64) ** 2).sum())
    assert int(n2) == want
    assert npc.index_ntotal("fine") == 897 + 3 * want
    import numpy as np
    assert np.array(npc.cloud_pos("fine")).shape == (897 + 3 * want, 3)
