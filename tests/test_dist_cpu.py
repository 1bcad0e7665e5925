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


def _exchange_worker(rank, world, port, out):
    import importlib
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    hps = importlib.import_module("hierarchical-point-slam_b200")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)
    # gradients laid out the way the fused backward hands them out: views of one flat buffer (with padding and a
    # scratch region in between), plus stragglers in their own storage, one of them float64, one missing
    flat = torch.randn(4096)
    shapes = [(0, (7, 5)), (36, (100,)), (200, (32, 32)), (2000, (3,))]        # (offset, shape); [1300, 2000) is scratch
    params = [torch.zeros(s) for _, s in shapes]
    for p, (off, s) in zip(params, shapes):
        n = p.numel()
        p.grad = flat[off:off + n].view(s)
    solo = [torch.zeros(70000), torch.zeros(11), torch.zeros(5, dtype=torch.float64), torch.zeros(4)]
    solo[0].grad, solo[1].grad, solo[2].grad = torch.randn(70000), torch.randn(11), torch.randn(5, dtype=torch.float64)
    targets = params + solo
    mine = [None if t.grad is None else t.grad.clone() for t in targets]
    ex = hps.parallel.GradExchange(targets)
    ex(dist)
    n_coll = len(ex.spans) + (1 if ex.bucket else 0)
    # second iteration on the same buffers (plan reused): refill with this rank's values and exchange again
    for t, m in zip(targets, mine):
        t.grad.copy_(m if m is not None else torch.zeros_like(t))
    ex(dist)
    torch.save(dict(mine=mine, got=[t.grad.clone() for t in targets], n_coll=n_coll), out % rank)
    dist.destroy_process_group()


def test_grad_exchange_coalesces_views_of_one_buffer(tmp_path):
    out = str(tmp_path / "ex%d.pt")
    mp.spawn(_exchange_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = [torch.load(out % k, weights_only=False) for k in range(2)]
    for i in range(len(r[0]["got"])):
        a = r[0]["mine"][i] if r[0]["mine"][i] is not None else torch.zeros_like(r[0]["got"][i])
        b = r[1]["mine"][i] if r[1]["mine"][i] is not None else torch.zeros_like(r[1]["got"][i])
        for k in range(2):
            assert torch.allclose(r[k]["got"][i], a + b, rtol=0, atol=1e-6), i
    # 4 views of one buffer -> 1 collective; the 70000-element gradient and the float64 one on their own; 2 small -> 1 bucket
    assert r[0]["n_coll"] == 4


class _RowRenderer:
    """Stand-in with Renderer.render_img's row_range contract (no kernels on the CPU): pixel values are a fixed function
    of the pixel coordinates, so the gathered frame can be compared with the one-process frame."""
    H, W = 37, 16          # 37 rows: uneven shards

    def render_img(self, scale, row_range=None):
        r0, r1 = row_range if row_range is not None else (0, self.H)
        v, u = torch.meshgrid(torch.arange(r0, r1, dtype=torch.float64), torch.arange(self.W, dtype=torch.float64), indexing="ij")
        depth = scale * (v * 100 + u)
        color = torch.stack([v, u, v + u], -1).float()
        return depth, depth * 0.5, color


def _render_worker(rank, world, port, out):
    import importlib
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    hps = importlib.import_module("hierarchical-point-slam_b200")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r = _RowRenderer()
    full = hps.parallel.render_img_sharded(r, dist, 2.0)
    local = hps.parallel.render_img_sharded(r, dist, 2.0, gather=False)
    torch.save(dict(full=full, local=local), out % rank)
    dist.destroy_process_group()


def test_render_img_sharded_over_rows_equals_the_whole_frame(tmp_path):
    """parallel.render_img_sharded (BASELINE configs[4]): rows dealt out to two ranks and all-gathered equal
    the one-process frame on every rank; gather=False hands back the rank's own rows."""
    out = str(tmp_path / "img%d.pt")
    mp.spawn(_render_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    want = _RowRenderer().render_img(2.0)
    bounds = [(0, 19), (19, 37)]
    for k in range(2):
        r = torch.load(out % k, weights_only=False)
        for got, w in zip(r["full"], want):
            assert got.dtype == w.dtype and torch.equal(got, w)
        d, u, c, rows = r["local"]
        assert tuple(rows) == bounds[k]
        assert torch.equal(d, want[0][rows[0]:rows[1]]) and torch.equal(c, want[2][rows[0]:rows[1]])
