"""Host-side logic of the product package that needs no GPU: SM split policy, zero pool, live rows bookkeeping."""
import importlib

import pytest
import torch

hps = importlib.import_module("hierarchical-point-slam_b200")
ops = hps.ops


def test_sm_split_policy():
    # one wave or less: no cap
    assert ops._sm_split(5000, "fwd") == (0, 0)
    for which in ("fwd", "bwd"):
        for P in (25000, 40000, 100000, 250000):
            nc, ng = ops._sm_split(P, which)
            assert nc > 0 and ng > 0 and nc + ng == ops.N_SM
            tiles = (P + 127) // 128
            r = ops._COST_RATIO[which]
            cost = max(-(-tiles // nc) * r, -(-tiles // ng))
            # no other split of the 148 SMs is better
            best = min(max(-(-tiles // a) * r, -(-tiles // (ops.N_SM - a))) for a in range(1, ops.N_SM))
            assert cost <= best + 1e-9
    assert ops._sm_split(25000, "fwd") == (98, 50)      # the bench workload: 2 colour tiles | 4 geometry tiles per CTA


def test_zero_pool_views_are_aligned_and_disjoint():
    zp = ops._ZeroPool("cpu")
    a = zp.want((7, 5))
    b = zp.want_like([torch.empty(3), torch.empty(4, 4), torch.empty(1)], skip_first=True)
    c = zp.want((2,))
    zp.alloc()
    ta, tb, tc = zp.get(a), zp.get(b), zp.get(c)
    assert tb[0] is None and tb[1].shape == (4, 4) and tb[2].shape == (1,)
    views = [ta, tb[1], tb[2], tc]
    base = zp.flat.data_ptr()
    spans = []
    for v in views:
        assert float(v.abs().sum()) == 0.0
        off = v.data_ptr() - base
        assert off % 16 == 0                      # red.global.add.v4 needs 16-byte rows
        spans.append((off, off + v.numel() * 4))
    spans.sort()
    assert all(spans[i][1] <= spans[i + 1][0] for i in range(len(spans) - 1))
    assert zp.get(None) is None


def test_live_rows_bookkeeping_on_cpu():
    table = torch.arange(40, dtype=torch.float32).reshape(10, 4).contiguous()
    idx = torch.tensor([7, 2, 5])
    rows = ops.LiveRows(table, idx)
    assert rows.param.shape == (3, 4) and rows.param.requires_grad and not rows.fresh
    assert rows.row_map.tolist() == [-1, -1, 1, -1, -1, 2, -1, 0, -1, -1]
    with torch.no_grad():
        rows.param += 100.0
    rows.refresh()                                 # index copy back into the table
    assert torch.equal(rows.table[idx], rows.param.detach())
    assert torch.equal(rows.table[0], table[0])
    d = rows.dense()
    assert d.shape == table.shape and d.requires_grad
    other = ops.LiveRows.pair(table.clone(), table.clone(), idx)
    assert other[0].row_map is other[1].row_map    # one map for the two tables of a level


def test_ops_refuse_cpu_tensors():
    # there is no CPU fallback: the kernels' host wrappers raise on non-CUDA tensors
    with pytest.raises(Exception):
        ops.exposure_mlp(torch.zeros(8), torch.zeros(128, 8), torch.zeros(128), torch.zeros(12, 128), torch.zeros(12))
    with pytest.raises(Exception):
        ops.camera_from_tensor(torch.zeros(7))


def test_dynamic_radius_maps_match_the_scipy_restatement():
    """frame.dynamic_radius (Tracker.py:297-325, Mapper.py:1026-1050 as tensor ops) vs the numpy/scipy oracle restatement
    of skimage 0.19.3's rgb2gray + sobel_h/sobel_v + interp1d: radii within 1e-9 m (float64 throughout, like the reference's numpy frames)."""
    from oracle import hps_oracle as O
    g = torch.Generator().manual_seed(9)
    H, W = 60, 80
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    img = torch.stack([0.5 + 0.5 * torch.sin(xx / 7), 0.5 + 0.5 * torch.cos(yy / 5), torch.rand(H, W, generator=g) * 0.3], -1)
    img[20:30, 30:50] = 0.9                      # hard edges -> the clip at the threshold is exercised
    cfg = hps.scannet_cfg("cpu")
    rh, ratio, thr = cfg["pointcloud"]["radius_hierarchy"], cfg["pointcloud"]["radius_query_ratio"], cfg["pointcloud"]["color_grad_threshold"]
    r_add, r_query = hps.frame.dynamic_radius(img, rh, ratio, thr)
    ref = O.dynamic_radius_maps(img.numpy(), rh, ratio, thr)
    for level in rh:
        a, q = ref[level]
        assert r_add[level].dtype == torch.float64 and r_add[level].shape == (H, W)
        assert float((r_add[level] - torch.from_numpy(a)).abs().max()) < 1e-9
        assert float((r_query[level] - torch.from_numpy(q)).abs().max()) < 1e-9
        vals = list(rh[level].values())
        assert float(r_add[level].max()) <= vals[0] + 1e-12 and float(r_add[level].min()) >= vals[1] - 1e-12
        assert float(r_add[level].min()) < vals[0] - 1e-4        # the map is not constant on this image
