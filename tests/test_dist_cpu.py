"""world_size-2 gloo test (CPU) of the ray-sharding + gradient all-reduce logic: the sharded
oracle iteration, with gradients all-reduced, equals the single-process iteration."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import importlib
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from oracle import hps_oracle as O
    from tests import parity
    hps = importlib.import_module("hierarchical-point-slam_b200")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    scene = parity.load("scene_small.pt")
    cl, rays = scene["clouds"], scene["rays"]
    stage, level = "color_fine", "fine"
    gf = {l: cl[l]["geo_feats"].clone().requires_grad_(True) for l in cl}
    cf = {l: cl[l]["col_feats"].clone().requires_grad_(True) for l in cl}
    dec = {k: {kk: vv.clone().requires_grad_(True) for kk, vv in v.items()} for k, v in scene["decoders"].items()}
    ro, rd, gd, gc, rad = hps.parallel.shard_rays(rank, world, rays["rays_o"], rays["rays_d"], rays["gt_depth"],
                                                  rays["gt_color"], cl[level]["radius"])
    d, v, c, m = O.render_batch_ray({l: cl[l]["points"] for l in cl}, dec, rd, ro, stage, gd, gf, cf,
                                    {level: rad, "mid": rad}, scene["c_fill_geo"], scene["c_fill_col"])
    aff = O.exposure_affine(dec["color_decoder_fine"], scene["exposure_feat"])
    loss = O.mapper_loss(d, c, m, gd, gc, stage, 0.1, aff)
    loss.backward()
    targets = [gf[level], cf[level]] + [t for t in dec["color_decoder_fine"].values()]
    hps.parallel.allreduce_grads(targets, dist)
    tot = hps.parallel.allreduce_scalars([loss.detach()], dist)
    if rank == 0:
        torch.save(dict(g_geo=gf[level].grad, g_col=cf[level].grad, loss=tot[0],
                        g_dec={k: t.grad for k, t in dec["color_decoder_fine"].items()}), out)
    dist.destroy_process_group()


def test_sharded_iteration_equals_single_process(tmp_path):
    from tests import parity
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    ref = parity.load("render_mapper_small.pt")["color_fine"]
    assert parity.rel(got["loss"], ref["loss"]) < 1e-5
    assert parity.rel(got["g_geo"], ref["g_geo"]) < 1e-4 and parity.rel(got["g_col"], ref["g_col"]) < 1e-4
    for k, g in ref["g_dec"].items():
        assert parity.rel(got["g_dec"][k], g) < 1e-4, k


def test_shard_bounds_cover_everything():
    import importlib
    hps = importlib.import_module("hierarchical-point-slam_b200")
    for n in (0, 1, 7, 5000, 5003):
        for w in (1, 2, 4, 8):
            b = [hps.parallel.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
