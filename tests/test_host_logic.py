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


def test_backward_pool_layout_puts_the_feature_gradients_first():
    """_RenderRays._plan_pool: the layout of every zero-initialised buffer of the fused backward inside one allocation -
    feature-table gradients first (GradPool.features: what a multi-GPU exchange can take early), everything 16-byte
    aligned, a caller-owned GradPool told what was used; clear=False leaves the clearing to the caller's kernel."""
    import types
    n_p = ops.N_PARAM
    gprm = [torch.empty(3, 4)] + [torch.empty(5, 7) for _ in range(n_p - 1)]
    cprm = [torch.empty(3, 4)] + [torch.empty(2, 9) for _ in range(n_p - 1)]
    eprm = [torch.empty(8), torch.empty(16, 8), torch.empty(16), torch.empty(12, 16), torch.empty(12)]
    spec = types.SimpleNamespace(color_stage=True, tracker=False)
    # (spec, rays_o, rays_d, gt_depth, feats_geo, feats_col, affine, *geo params, *colour params, *exposure params)
    need = [False, False, False, False, True, True, False] + [False] + [True] * (n_p - 1) + [False] + [True] * (n_p - 1) \
        + [True] * 5
    zp, (h_dg, h_dcw, h_daff, h_dfg, h_dfc, h_exp, h_dpg, h_dpc) = ops._RenderRays._plan_pool(
        spec, need, gprm, cprm, None, 1000, 640, "cpu", eprm, None, "tc")
    assert zp.features == 2 * 1000 * ops.C_DIM and zp.total % 4 == 0 and zp.total > zp.features
    assert h_daff is None and h_dpg is None and h_dpc is None and h_exp is not None
    assert h_dg[0] is None and h_dcw[0] is None                       # the Fourier matrices are not parameters
    pool = ops.GradPool(torch.full((zp.total + 64,), 7.0))
    zp.alloc(into=pool, clear=False)
    assert pool.used == zp.total and pool.features == zp.features and not pool.zero_in_forward
    assert float(zp.flat.min()) == 7.0                                # untouched: the caller's kernel clears it
    dfg, dfc = zp.get(h_dfg), zp.get(h_dfc)
    base = pool.tensor.data_ptr()
    assert dfg.data_ptr() == base and dfc.data_ptr() == base + 4 * 1000 * ops.C_DIM
    for v in [dfg, dfc, zp.get(h_exp)] + [t for t in zp.get(h_dg) + zp.get(h_dcw) if t is not None]:
        off = v.data_ptr() - base
        assert off % 16 == 0 and off + 4 * v.numel() <= 4 * zp.total
        assert (off < 4 * zp.features) == (v is dfg or v is dfc)
    # a tracker call (pose gradients through the tensor-core dgrad kernels) adds the two (P,3) sums behind everything else
    spec_t = types.SimpleNamespace(color_stage=True, tracker=True)
    need_t = [False, True, True, False, False, False, False] + [False] * (2 * n_p)
    zt, hs = ops._RenderRays._plan_pool(spec_t, need_t, gprm, cprm, None, 1000, 640, "cpu", None, None, "tc")
    assert zt.features == 0 and hs[6] is not None and hs[7] is not None and zt.total == 2 * 640 * 3
    with pytest.raises(hps._lib.HpsError, match="grad_buffer"):
        zt.alloc(into=ops.GradPool(torch.zeros(16)))


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


# ---- pixel / ray sampling and pose conversion against the reference's own outputs -------------------

def test_get_samples_matches_reference_fixture(golden):
    """Same torch seed -> same torch.randint draw -> the very same rays, depths, colours and pixel indices as the
    reference's get_samples (common.py:236-258), bit for bit (fixture: oracle/make_golden_sampling.py)."""
    g = golden("sampling_small.pt")
    fx, fy, cx, cy = g["intr"]
    for case in g["cases"]:
        H0, H1, W0, W1 = case["window"]
        torch.manual_seed(case["seed"])
        out = hps.common.get_samples(H0, H1, W0, W1, case["n"], g["H"], g["W"], fx, fy, cx, cy, g["c2w"], g["depth"],
                                     g["color"], "cpu", depth_filter=case["depth_filter"], return_index=True,
                                     depth_limit=case["depth_limit"])
        assert len(out) == 6
        for name, a, b in zip(("rays_o", "rays_d", "depth", "color", "i", "j"), out, case["out"]):
            assert a.dtype == b.dtype and a.shape == b.shape, (case["seed"], name)
            assert torch.equal(a, b), (case["seed"], name)
        four = None
        torch.manual_seed(case["seed"])
        four = hps.common.get_samples(H0, H1, W0, W1, case["n"], g["H"], g["W"], fx, fy, cx, cy, g["c2w"], g["depth"],
                                      g["color"], "cpu", depth_filter=case["depth_filter"],
                                      depth_limit=case["depth_limit"])
        assert len(four) == 4 and torch.equal(four[1], case["out"][1])
    torch.manual_seed(9)
    uv = hps.common.get_sample_uv(3, 40, 2, 60, 100, g["depth"], g["color"], device="cpu")
    for a, b in zip(uv, g["uv"]):
        assert a.dtype == b.dtype and torch.equal(a, b)


def test_select_uv_draws_like_a_dense_gather():
    H, W = 12, 9
    i, j = torch.meshgrid(torch.arange(W).float(), torch.arange(H).float(), indexing="ij")
    i, j = i.t(), j.t()
    depth, color = torch.rand(H, W), torch.rand(H, W, 3)
    torch.manual_seed(3)
    si, sj, sd, sc = hps.common.select_uv(i, j, 40, depth, color, device="cpu")
    torch.manual_seed(3)
    pick = torch.randint(H * W, (40,))
    assert torch.equal(si, (pick % W).float()) and torch.equal(sj, (pick // W).float())
    assert torch.equal(sd, depth.reshape(-1)[pick]) and torch.equal(sc, color.reshape(-1, 3)[pick])


def test_get_tensor_from_camera_matches_reference_fixture(golden):
    """All four branches of the matrix -> quaternion conversion (largest of m00, m11, m22, trace) against the
    reference's scipy route (common.py:357-382); float64 arithmetic on both sides, so float32 results agree to the
    last bit or one ulp."""
    g = golden("sampling_small.pt")
    for RT, (q, tq, q_np) in zip(g["mats"], g["quats"]):
        mine = hps.common.get_tensor_from_camera(RT)
        assert mine.dtype == torch.float32 and mine.shape == (7,)
        assert torch.allclose(mine, q, rtol=0, atol=2e-7), (mine, q)
        assert torch.allclose(hps.common.get_tensor_from_camera(RT, Tquad=True), tq, rtol=0, atol=2e-7)
        assert torch.allclose(hps.common.get_tensor_from_camera(RT[:3].numpy()), q_np, rtol=0, atol=2e-7)
        back = hps.common.get_camera_from_tensor(mine)            # and round trip through the path's own helper
        assert torch.allclose(back, RT[:3], atol=2e-6)
    with pytest.raises(ValueError):
        hps.common.get_tensor_from_camera(torch.diag(torch.tensor([1.0, 1.0, -1.0, 1.0])))


def test_rotation_from_tensor_shapes():
    q = torch.randn(4, 4)
    R = hps.common.get_rotation_from_tensor(q)
    assert R.shape == (4, 3, 3)
    assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3).expand(4, 3, 3), atol=1e-5)
    assert torch.equal(hps.common.get_rotation_from_tensor(q[2]), R[2])


def test_gradient_guided_pixel_selection():
    """get_selected_index_with_grad / get_sample_uv_with_grad (common.py:174-233; PARITY UNPINNED - skimage is not
    installed) against the Sobel restatement the dynamic-radius oracle uses: the selected set is the top ratio*n
    gradient pixels, restricted to the window and to pixels with a (limited) depth reading."""
    from oracle import hps_oracle as O
    import numpy as np
    g = torch.Generator().manual_seed(5)
    H, W, n, ratio = 40, 56, 20, 15
    img = torch.rand(H, W, 3, generator=g, dtype=torch.float64)
    depth = 1 + 5 * torch.rand(H, W, generator=g)
    depth[torch.rand(H, W, generator=g) < 0.2] = 0
    sel, mag = hps.common.get_selected_index_with_grad(4, H - 4, 6, W - 6, n, img, ratio=ratio, gt_depth=depth,
                                                       depth_limit=True)
    assert isinstance(sel, np.ndarray) and isinstance(mag, np.ndarray) and mag.shape == (H, W)
    want_mag = O.sobel_magnitude(img.numpy())
    assert np.allclose(mag, want_mag, rtol=0, atol=1e-14)
    top = np.argpartition(want_mag, -ratio * n, axis=None)[-ratio * n:]
    hh, ww = np.unravel_index(top, (H, W))
    keep = (hh >= 4) & (hh < H - 4) & (ww >= 6) & (ww < W - 6)
    d = depth.numpy()[hh, ww]
    keep &= (d > 0) & (d <= 5.0)
    assert set(sel.tolist()) == set(top[keep].tolist())
    sel_t, mag_t = hps.common.get_selected_index_with_grad(4, H - 4, 6, W - 6, n, img, ratio=ratio, as_numpy=False)
    assert torch.is_tensor(sel_t) and sel_t.dtype == torch.int64 and torch.is_tensor(mag_t)
    keep0 = (hh >= 4) & (hh < H - 4) & (ww >= 6) & (ww < W - 6)
    assert set(sel_t.tolist()) == set(top[keep0].tolist())

    torch.manual_seed(0)
    picks = hps.common.get_sample_uv_with_grad(0, H, 0, W, n, img)
    assert picks.shape == (n,) and len(set(picks.tolist())) == n
    top5 = set(np.argpartition(want_mag, -5 * n, axis=None)[-5 * n:].tolist())
    assert set(picks.tolist()) <= top5

    c2w = torch.eye(4)
    out = hps.common.get_samples_with_pixel_grad(0, H, 0, W, (n, n), H, W, 50.0, 50.0, W / 2, H / 2, c2w,
                                                 depth, img.float(), "cpu", est_normal=img.flip(0))
    ro, rd, sd, sc, i, j = out
    assert i.dtype == torch.int64 and torch.equal(sd, depth[j, i]) and torch.equal(sc, img.float()[j, i])
    assert bool((sd > 0).all()) and torch.equal(torch.unique(j * W + i), j * W + i)      # sorted, distinct (union1d)
    want_o, want_d = hps.common.get_rays_from_uv(i.float(), j.float(), c2w, H, W, 50.0, 50.0, W / 2, H / 2, "cpu")
    assert torch.equal(rd, want_d) and torch.equal(ro, want_o)


def test_loss_buffer_layout_and_frame_checks():
    """Host-side bookkeeping of the fused losses: accumulator head sizes and the (affine, frame index) validation."""
    assert ops._loss_head_floats(0) == 16 and ops._loss_head_floats(12) == 16
    assert ops._loss_head_floats(13) == 32 and ops._loss_head_floats(12 * 5) == 64
    aff1, aff3 = torch.zeros(12), torch.zeros(3, 12)
    assert ops._loss_frames(None, None, 10) == (None, 0)
    assert ops._loss_frames(aff1, None, 10) == (None, 12)
    frm, n = ops._loss_frames(aff3, torch.tensor([0, 1, 2, 2]), 4)
    assert frm.dtype == torch.int32 and n == 36
    with pytest.raises(ValueError):
        ops._loss_frames(aff3, None, 4)                       # several affines need a frame index
    with pytest.raises(ValueError):
        ops._loss_frames(aff3, torch.tensor([0, 1]), 4)       # one entry per ray
    with pytest.raises(ValueError):
        ops._loss_frames(torch.zeros(13), None, 4)


def test_z_rows_for_depthless_rays_are_validated():
    assert ops._z_zero_stride(None, 7, 5) == 0
    assert ops._z_zero_stride(torch.zeros(5), 7, 5) == 0
    assert ops._z_zero_stride(torch.zeros(7, 5), 7, 5) == 5
    with pytest.raises(ValueError):
        ops._z_zero_stride(torch.zeros(4), 7, 5)
    with pytest.raises(ValueError):
        ops._z_zero_stride(torch.zeros(6, 5), 7, 5)


def test_mapper_loss_spec_modes():
    with pytest.raises(ValueError):
        ops.MapperLossSpec(None, 0.1, exposure="latent")
    s = ops.MapperLossSpec(None, 0.1, "affine", torch.zeros(3, dtype=torch.int64))
    assert s.exposure == "affine" and s.w_color == pytest.approx(0.1)
    assert ops.TrackerLossSpec(None, 0.5).zero_missing is False


def test_mapper_schedule_matches_the_reference_statements():
    """mapper_optim's stage selection, learning rates, BA pose learning rate and iteration counts against the fixture
    made by EXECUTING Mapper.py:605-612 and :633-666 of the reference (oracle/make_golden_schedule.py)."""
    import json
    import os
    GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    mo = hps.mapper_optim
    with open(os.path.join(GOLDEN, "mapper_schedule.json")) as f:
        gold = json.load(f)
    ref_cfg = gold["mapping_cfg"]
    mine = hps.scannet_cfg()["mapping"]
    for k, v in ref_cfg.items():                    # config.py carries the reference's merged yaml numbers
        assert mine[k] == v, k
    for case in gold["cases"]:
        n = case["num_joint_iters_in"]
        if case["more_iters_when_adding"] and case["idx"] > 0 and not case["color_refine"]:
            n = mo.joint_iters_when_adding(n, case["frame_pts_add"], ref_cfg["min_iter_ratio"])
        assert n == case["num_joint_iters"]
        assert mo.split_iters(n, case["mid_iter_ratio"]) == (case["num_mid_iters"], case["num_fine_iters"])
        last = case["idx"] == case["n_img"] - 1 and case["color_refine"]
        covered = 0
        for first, count, stage, *lrs in case["rows"]:
            for j in range(first, first + count):
                st = mo.mapper_stage(j, n, case["mid_iter_ratio"], case["geo_iter_ratio"], case["init"],
                                     ref_cfg["geo_iter_first"])
                assert st == stage, (j, st, stage)
                lr = mo.stage_learning_rates(ref_cfg, st, case["init"], last)
                got = [lr["decoders"], lr["geometry_mid"], lr["geometry_fine"], lr["color_mid"], lr["color_fine"]]
                cam = mo.ba_camera_lr(j, n, case["mid_iter_ratio"], case["geo_iter_ratio"], ref_cfg["BA_cam_lr"]) \
                    if case["BA"] else 0.0
                assert got + [cam] == lrs, (j, got, cam, lrs)
            covered += count
        assert covered == n


def test_mapper_optimizer_bookkeeping_with_emulated_kernels(monkeypatch):
    """MapperOptimizer's host side without a GPU: which tensors go into which launch, the per-tensor step-counter
    pointers, the scratch words, the learning rates handed to the kernels, the row write-back arguments.  The two
    Adam entry points are replaced by a numpy emulation of their documented semantics (include/hps_b200.h) acting on
    the raw pointers the host code passes, and the GPU parity test's own body runs against it on CPU tensors - so
    the comparison with the reference's single torch.optim.Adam (Mapper.py:592-603, 633-667) covers everything except
    the kernels' arithmetic, which `-m gpu` checks."""
    import contextlib
    import ctypes
    import numpy as np

    def arr(ptr, n, dt=np.float32):
        if isinstance(ptr, ctypes.c_void_p):
            ptr = ptr.value
        return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n,))

    def adam(p, g, m, v, lr, b1, b2, eps, t):
        b1, b2 = np.float32(b1), np.float32(b2)
        bc1, bc2 = np.float32(1 - float(b1) ** t), np.float32(1 - float(b2) ** t)
        m[:] = m + (g - m) * (np.float32(1) - b1)
        v[:] = v * b2 + (np.float32(1) - b2) * g * g
        p[:] = p - (np.float32(lr) / bc1) * (m / (np.sqrt(v) / np.sqrt(bc2) + np.float32(eps)))

    launches = []

    def fake_call(name, *a):
        if name == "hps_adam_rows":
            p, g, m, v, M, C, lr, b1, b2, eps, step, table, rows, _stream = a
            st = arr(step, 2, np.int32)
            P = arr(p, M * C)
            adam(P, arr(g, M * C), arr(m, M * C), arr(v, M * C), lr, b1, b2, eps, int(st[0]) + 1)
            st[0] += 1
            if table is not None:
                R = arr(rows, M, np.int64)
                arr(table, (int(R.max()) + 1) * C).reshape(-1, C)[R] = P.reshape(M, C)
        elif name == "hps_adam_multi_own_steps":
            n, ps, gs, ms, vs, numel, lrs, steps, b1, b2, eps, done, zero_grad, _stream = a
            launches.append(n)
            assert 0 < n <= 24 and arr(done, 1, np.int32)[0] == 0
            assert len({steps[i] for i in range(n)}) == n               # distinct counters
            for i in range(n):
                st = arr(steps[i], 1, np.int32)
                G = arr(gs[i], numel[i])
                adam(arr(ps[i], numel[i]), G, arr(ms[i], numel[i]), arr(vs[i], numel[i]), lrs[i], b1, b2, eps, int(st[0]) + 1)
                st[0] += 1
                if zero_grad:
                    G[:] = 0
        else:
            raise AssertionError("unexpected kernel call " + name)

    monkeypatch.setattr(ops, "_call", fake_call)
    monkeypatch.setattr(ops, "_need_cuda", lambda *t: None)
    monkeypatch.setattr(ops, "_stream", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    from tests import test_gpu_kernels as T
    monkeypatch.setattr(T, "DEV", "cpu")
    T.test_mapper_optimizer_follows_torch_adam_through_the_stages(hps)
    # geometry stages: 6 decoder tensors; colour stages: 24 decoder tensors + 2 poses + the latent = 24 + 3
    assert sorted(set(launches)) == [3, 6, 24]


def test_frustum_selection_arithmetic_matches_the_reference_method(golden):
    """The product's frustum feature selection (neural_point.frustum_select - the tensor ops NeuralPointCloud.frustum_rows
    runs on the device, here on CPU tensors) vs the reference's Mapper.get_mask_from_c2w executed with the real cv2
    (tests/golden/frustum_small.pt): identical index sets; 3x4 poses; empty cloud; and its depth sampler vs the
    oracle's cv2.remap restatement bit for bit."""
    import numpy as np
    from oracle import hps_oracle as O
    npm = hps.neural_point
    fx = golden("frustum_small.pt")
    K = fx["intrinsics"]
    args = (fx["depth"], K["H"], K["W"], K["fx"], K["fy"], K["cx"], K["cy"])
    for case in fx["cases"]:
        got = npm.frustum_select(case["points"], case["c2w"], *args, case["edge"])
        assert got.dtype == torch.long and torch.equal(got, case["indices"])
        assert torch.equal(npm.frustum_select(case["points"], case["c2w"][:3], *args, case["edge"]), got)
    assert npm.frustum_select(torch.zeros(0, 3), fx["cases"][0]["c2w"], *args, 0).numel() == 0
    g = torch.Generator().manual_seed(4)
    x = torch.rand(50_000, generator=g) * 700 - 30
    y = torch.rand(50_000, generator=g) * 540 - 30
    x[:4] = torch.tensor([float("nan"), float("inf"), -1e12, 639.0])
    mine = npm.remap_bilinear(fx["depth"], x, y)
    assert np.array_equal(mine.numpy(), O.remap_bilinear_cv2(fx["depth"].numpy(), x.numpy(), y.numpy()))
    try:
        import cv2
    except ImportError:
        return
    ref = np.concatenate([cv2.remap(fx["depth"].numpy(), x.numpy()[i:i + 30000], y.numpy()[i:i + 30000],
                                    interpolation=cv2.INTER_LINEAR)[:, 0] for i in range(0, 50_000, 30000)])
    assert np.array_equal(mine.numpy(), ref)                       # the product's sampler IS cv2.remap, bit for bit


def test_neural_point_cloud_reference_accessors():
    """The small accessors of the reference class that callers outside the render path use (checkpoint logger
    Logger.py:25-40, tools): present, with the reference's return conventions; no device needed."""
    import math
    npc = hps.NeuralPointCloud(hps.scannet_cfg("cpu"))
    for level in ("mid", "fine"):
        npc._init_level(level)
    assert npc.input_pos() == [] and npc.input_rgb() == [] and npc.input_normal() == [] and npc.input_normal_cartesian() == []
    assert npc.pts_num() == {"mid": 0, "fine": 0} and npc.index_ntotal("fine") == 0
    assert not npc.index_is_trained() and not npc.index_is_trained("mid")
    assert npc.index_train(torch.zeros(4, 3), level="mid") and npc.index_is_trained("mid") and not npc.index_is_trained()
    assert npc.index_train(torch.zeros(4, 3)) and npc.index_is_trained()
    with pytest.raises(AssertionError):
        npc.index_train([[0.0, 0.0, 0.0]])                      # 'use tensor to train FAISS index'
    assert npc.index_get_nprobe() == 4 and npc.index_get_nprobe("fine") == 4
    npc.index_set_nprobe(400)
    assert npc.index_get_nprobe("mid") == 400 and npc.get_index("fine").nprobe == 400
    npc.set_pts_num({"mid": 3, "fine": 5})
    assert npc.pts_num()["fine"] == 5
    assert npc.get_c_dim() == 32 and npc.get_radius_query() == 0.08 and npc.get_radius_add() == 0.04
    assert npc.merge_points([]) is None
    n = torch.tensor([[0.0, 0.0, 1.0], [1.0, 0.0, 0.0], [0.0, -1.0, 0.0], [0.6, 0.0, 0.8]])
    sph = npc.cart2sph(n)
    want = torch.tensor([[0.0, 0.0], [math.pi / 2, 0.0], [math.pi / 2, -math.pi / 2], [math.atan2(0.6, 0.8), 0.0]])
    assert torch.allclose(sph, want, atol=1e-7)


def test_zero_depth_sampling_rows_follow_the_reference_chunks():
    """render_img widens the reference's 3000-ray chunks; the sampling range of rays WITHOUT a depth reading is a statistic
    of the reference chunk they fall into (Renderer.py:133-134, 148-153, 207-208).  The segmented computation of the
    per-ray z rows must equal the chunk-by-chunk one (the expressions render_batch_ray uses for one chunk), including a
    ragged last chunk, an all-missing chunk and a chunk whose bound is clamped by 1.2 max(d)."""
    import types
    r = hps.Renderer(hps.scannet_cfg("cpu"), None, types.SimpleNamespace(H=4, W=4, fx=1.0, fy=1.0, cx=2.0, cy=2.0))
    g = torch.Generator().manual_seed(6)
    chunk, R = 100, 100 * 7 + 37
    gt = 0.5 + 4.0 * torch.rand(R, generator=g)
    gt[torch.rand(R, generator=g) < 0.2] = 0.0
    gt[200:300] = 0.0                                  # a chunk without any depth reading
    gt[300:400] = 0.01
    gt[333] = 9.0                                      # 5 mean(d) < 1.2 max(d): the mean term decides
    rows = r._z_rows_for_zero_depth(gt, chunk, "cpu")
    assert rows.shape == (R, 5) and rows.is_contiguous()
    for i in range(0, R, chunk):
        want = r._z_for_zero_depth(gt[i:i + chunk], "cpu")
        assert torch.allclose(rows[i:i + chunk], want.expand(min(chunk, R - i), 5), rtol=2e-6, atol=0), i
    assert torch.allclose(rows[200], torch.linspace(0.3, 0.0, 5), atol=1e-7)            # linspace(near_end, far = 0)
    assert float(rows[0, 0]) == pytest.approx(0.3) and float(rows[0, -1]) > 3.0


def test_depth_inside_mask_is_the_reference_statement():
    """common.depth_inside_mask vs the reference's own statement executed where it lies (Tracker.py:179-181,
    Mapper.py:766-768); skipped when the reference tree is absent (GPU box)."""
    import os
    import textwrap
    ref = os.environ.get("HPS_REFERENCE_ROOT", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "src")):
        pytest.skip("reference tree not present")
    g = torch.Generator().manual_seed(9)
    cases = [1.0 + 3.0 * torch.rand(1000, generator=g), torch.cat([0.2 * torch.rand(900, generator=g), 40 * torch.ones(100)]),
             torch.tensor([2.0, 2.0, 2.0, 50.0]), torch.rand(7, generator=g)]
    for name, first, last in (("Tracker.py", 179, 181), ("Mapper.py", 766, 768)):
        with open(os.path.join(ref, "src", name)) as f:
            src = textwrap.dedent("".join(f.readlines()[first - 1:last]))
        assert src.startswith("with torch.no_grad():") and "median" in src
        code = compile(src, "%s:%d-%d" % (name, first, last), "exec")
        for d in cases:
            ns = {"torch": torch, "batch_gt_depth": d}
            exec(code, ns)
            got = hps.common.depth_inside_mask(d)
            assert got.dtype == torch.bool and torch.equal(got, ns["inside_mask"])
    assert not bool(hps.common.depth_inside_mask(cases[1]).all())           # the 10 x median bound does bind there


def test_keyframe_dict_is_handed_out_as_clones():
    npc = hps.NeuralPointCloud(hps.scannet_cfg("cpu"))
    kf = [{"idx": torch.tensor(3), "color": torch.rand(4, 4, 3), "est_c2w": torch.eye(4)}]
    npc.set_keyframe_dict(kf)
    out = npc.get_keyframe_dict()
    assert len(out) == 1 and torch.equal(out[0]["color"], kf[0]["color"])
    out[0]["color"].zero_()
    assert float(kf[0]["color"].abs().sum()) > 0                    # the caller's copy does not alias ours
    again = hps.common.clone_kf_dict(kf)
    assert again[0]["est_c2w"].data_ptr() != kf[0]["est_c2w"].data_ptr()


def test_gradient_guided_paths_match_the_reference_code(golden):
    """Pixel pre-selection by colour gradient and the dynamic radius maps vs the fixture made by running the
    reference's own get_selected_index_with_grad (common.py:199-233) and executing Tracker.py:297-323
    (oracle/make_golden_gradient_sampling.py).  Only scikit-image's rgb2gray / sobel_h / sobel_v were stand-ins there
    (restated with scipy.ndimage; skimage is not installed) - the top-k selection, window and depth masks, clipping at
    color_grad_threshold and the interp1d maps are the reference's code with the real numpy / scipy."""
    import numpy as np
    fx = golden("gradient_small.pt")
    H0, H1, W0, W1 = fx["window"]
    img, depth, n, ratio = fx["image"], fx["depth"], fx["n"], fx["ratio"]
    for limit, want in fx["selected"].items():
        kw = {} if limit is None else dict(gt_depth=depth, depth_limit=limit)
        sel, mag = hps.common.get_selected_index_with_grad(H0, H1, W0, W1, n, img, ratio=ratio, **kw)
        assert np.array_equal(np.sort(sel), want.numpy()), limit
        assert np.allclose(mag, fx["grad_mag"].numpy(), rtol=0, atol=1e-14)
    r_add, r_query = hps.frame.dynamic_radius(img, fx["radius_hierarchy"], fx["radius_query_ratio"],
                                              fx["color_grad_threshold"])
    assert set(r_add) == set(fx["dynamic_r_add"]) == {"mid", "fine"}
    for level in r_add:
        for mine, want in ((r_add[level], fx["dynamic_r_add"][level]), (r_query[level], fx["dynamic_r_query"][level])):
            assert mine.dtype == torch.float64 and mine.shape == want.shape
            assert float((mine - want).abs().max()) <= 1e-15 + 4e-16 * float(want.abs().max())
        assert float(r_add[level].max()) > float(r_add[level].min())            # the flat patch gets the large radius


@pytest.mark.parametrize("separate_LR", [False, True])
def test_tracker_optimizer_follows_the_reference_construction(monkeypatch, separate_LR):
    """TrackerOptimizer vs the reference's own optimiser construction - Tracker.py:383-415 EXECUTED where it lies
    (pose as one 7-vector or as translation + quaternion with lr / 5, exposure latent and both exposure MLPs at 0.001) -
    over 14 iterations of a non-monotone loss, with the keep-the-best-pose rule of Tracker.py:435-437.  The Adam entry
    point is emulated (numpy on the raw pointers, documented semantics of hps_adam_multi); the kernel itself is checked
    in the GPU suite.  Skipped when the reference tree is absent."""
    import contextlib
    import copy
    import ctypes
    import math
    import os
    import textwrap
    import types
    import numpy as np
    from torch.autograd import Variable
    ref_root = os.environ.get("HPS_REFERENCE_ROOT", "/root/reference")
    if not os.path.isdir(os.path.join(ref_root, "src")):
        pytest.skip("reference tree not present")

    def arr(ptr, n, dt=np.float32):
        if isinstance(ptr, ctypes.c_void_p):
            ptr = ptr.value
        return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n,))

    def fake_call(name, *a):
        # per-parameter step counters (torch.optim.Adam's state[p]['step']): documented semantics of hps_adam_multi_own_steps
        assert name == "hps_adam_multi_own_steps", name
        n, ps, gs, ms, vs, numel, lrs, steps, b1, b2, eps, done, zero_grad, _stream = a
        b1, b2 = np.float32(b1), np.float32(b2)
        for i in range(n):
            st = arr(steps[i], 1, np.int32)
            t = int(st[0]) + 1
            bc1, bc2 = np.float32(1 - float(b1) ** t), np.float32(1 - float(b2) ** t)
            p, g, m, v = (arr(x[i], numel[i]) for x in (ps, gs, ms, vs))
            m[:] = m + (g - m) * (np.float32(1) - b1)
            v[:] = v * b2 + (np.float32(1) - b2) * g * g
            p[:] = p - (np.float32(lrs[i]) / bc1) * (m / (np.sqrt(v) / np.sqrt(bc2) + np.float32(eps)))
            if zero_grad:
                g[:] = 0
            st[0] = t

    monkeypatch.setattr(ops, "_call", fake_call)
    monkeypatch.setattr(ops, "_need_cuda", lambda *t: None)
    monkeypatch.setattr(ops, "_stream", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())

    cfg = hps.scannet_cfg("cpu")
    tracking = dict(cfg["tracking"], separate_LR=separate_LR, iters=14)
    torch.manual_seed(1219)
    dec_mine = hps.get_model(cfg)
    dec_ref = copy.deepcopy(dec_mine)
    g = torch.Generator().manual_seed(3)
    cam0 = torch.tensor([0.99, 0.05, -0.08, 0.03, 0.4, -0.2, 0.1]) + 0.01 * torch.randn(7, generator=g)
    ef_mine = (0.3 * torch.randn(8, generator=g)).requires_grad_(True)
    ef_ref = ef_mine.detach().clone().requires_grad_(True)
    target = cam0 + 0.02 * torch.randn(7, generator=g)

    # ---- the reference's construction, executed
    with open(os.path.join(ref_root, "src", "Tracker.py")) as f:
        src = textwrap.dedent("".join(f.readlines()[382:415]))
    assert src.startswith("if self.separate_LR:") and src.rstrip().endswith("optimizer_camera = torch.optim.Adam(optim_para_list)")
    me = types.SimpleNamespace(separate_LR=separate_LR, cam_lr=tracking["lr"], encode_exposure=True, exposure_feat=ef_ref,
                               decoders=dec_ref, quad=None, T=None)
    ns = dict(torch=torch, Variable=Variable, self=me, camera_tensor=cam0.clone(), device="cpu")
    exec(compile(src, "Tracker.py:383-415", "exec"), ns)
    topt = ns["optimizer_camera"]
    assert [grp["lr"] for grp in topt.param_groups] == ([tracking["lr"], tracking["lr"] * 0.2] if separate_LR else
                                                        [tracking["lr"]]) + [0.001] * 3

    opt = hps.TrackerOptimizer(tracking, cam0.clone(), ef_mine, dec_mine)
    assert [opt.stage(k) for k in (0, 6, 7, 13)] == ["color_mid", "color_mid", "color_fine", "color_fine"]

    def loss_of(cam, ef, dec, k):
        # as in the tracker (Tracker.py:190-212): the first half of the iterations renders the mid level, so only ITS
        # exposure MLP receives a gradient; the fine level's MLP gets its first gradient at k = iters / 2 and
        # torch.optim.Adam then starts that parameter's bias correction at step 1
        level = "mid" if opt.stage(k) == "color_mid" else "fine"
        mlps = list(getattr(dec, "color_decoder_" + level).mlp_exposure.parameters())
        wobble = 1.0 + 0.9 * math.sin(1.7 * k)                       # non-monotone: the best pose is an interior iteration
        return (100 * wobble * ((cam - target) ** 2).sum() + (ef ** 2).sum()
                + sum((p * p).sum() for p in mlps) * 0.1 + wobble)

    current_min_loss, candidate = float(1e20), None
    for k in range(opt.num_cam_iters):
        cam = opt.camera_tensor()
        loss = loss_of(cam, ef_mine, dec_mine, k)
        loss.backward()
        opt.step(loss)
        # reference iteration: Tracker.py:428-437 (+ the zero_grad / backward / step of optimize_cam_in_batch)
        rcam = torch.cat([me.quad, me.T], 0) if separate_LR else ns["camera_tensor"]
        topt.zero_grad()
        rloss = loss_of(rcam, ef_ref, dec_ref, k)
        rloss.backward()
        topt.step()
        if rloss.item() < current_min_loss:
            current_min_loss = rloss.item()
            candidate = rcam.clone().detach()
        mine_now = torch.cat(opt.leaves, 0).detach() if separate_LR else opt.leaves[0].detach()
        ref_now = torch.cat([me.quad, me.T], 0).detach() if separate_LR else ns["camera_tensor"].detach()
        assert float((mine_now - ref_now).abs().max()) < 2e-6, k
        assert float((ef_mine.detach() - ef_ref.detach()).abs().max()) < 2e-6
    assert abs(float(opt.best_loss) - current_min_loss) <= 1e-4 * abs(current_min_loss)
    assert float((opt.best_camera_tensor - candidate).abs().max()) < 2e-6
    for level in ("mid", "fine"):
        for a, b in zip(getattr(dec_mine, "color_decoder_" + level).mlp_exposure.parameters(),
                        getattr(dec_ref, "color_decoder_" + level).mlp_exposure.parameters()):
            assert float((a.detach() - b.detach()).abs().max()) < 2e-6, level
    # the fine MLP sat out the first half: its counter is iters / 2, the pose's is iters
    steps = opt.opt.tensor_steps.tolist()
    assert max(steps) == opt.num_cam_iters and min(steps) == opt.num_cam_iters // 2


def test_mapper_optimizer_groups_follow_the_reference_construction(monkeypatch):
    """MapperOptimizer's tensor -> learning-rate mapping vs the reference's own optimiser: Mapper.py:590-603 (the seven
    parameter groups) and :633-666 (per-iteration learning rates) EXECUTED where they lie, for every iteration of a
    600-iteration mapping with BA.  Every tensor must carry the learning rate of the reference group that holds it.
    Skipped when the reference tree is absent."""
    import os
    import textwrap
    import types
    ref_root = os.environ.get("HPS_REFERENCE_ROOT", "/root/reference")
    if not os.path.isdir(os.path.join(ref_root, "src")):
        pytest.skip("reference tree not present")
    monkeypatch.setattr(ops, "_need_cuda", lambda *t: None)
    with open(os.path.join(ref_root, "src", "Mapper.py")) as f:
        lines = f.readlines()
    build_src = textwrap.dedent("".join(lines[589:603]))
    loop_src = textwrap.dedent("".join(lines[632:666]))
    assert build_src.startswith("if self.nice:") and build_src.rstrip().endswith("optimizer = torch.optim.Adam(optim_para_list)")
    assert loop_src.startswith("if joint_iter <=")
    mapping = hps.scannet_cfg("cpu")["mapping"]
    mk = lambda *s: torch.randn(*s).requires_grad_(True)
    dec = [mk(4, 3), mk(4), mk(2, 4)]
    rows = {"geometry_mid": mk(10, 32), "geometry_fine": mk(12, 32), "color_mid": mk(10, 32), "color_fine": mk(12, 32)}
    cams, expo = [mk(7), mk(7)], mk(8)
    me = types.SimpleNamespace(nice=True, BA=True, encode_exposure=True, exposure_feat=expo, stage=None,
                               geo_iter_first=mapping["geo_iter_first"], geo_iter_ratio=mapping["geo_iter_ratio"],
                               n_img=100, color_refine=False, BA_cam_lr=mapping["BA_cam_lr"])
    ns = dict(torch=torch, self=me, decoders_para_list=dec, geo_pcl_para_mid=[rows["geometry_mid"]],
              geo_pcl_para_fine=[rows["geometry_fine"]], color_pcl_para_mid=[rows["color_mid"]],
              color_pcl_para_fine=[rows["color_fine"]], camera_tensor_list=cams, cfg={"mapping": mapping}, idx=5,
              init=False, color_refine=False)
    exec(compile(build_src, "Mapper.py:590-603", "exec"), ns)
    ref_opt = ns["optimizer"]
    opt = hps.MapperOptimizer(mapping, decoder_params=dec, rows=rows, cameras=cams, exposure_feat=expo)
    n_iters = mapping["iters"]
    ns["num_mid_iters"], ns["num_fine_iters"] = hps.mapper_optim.split_iters(n_iters, mapping["mid_iter_ratio"])
    loop_code = compile(loop_src, "Mapper.py:633-666", "exec")
    for joint_iter in range(n_iters):
        ns["joint_iter"] = joint_iter
        exec(loop_code, ns)
        assert opt.set_iteration(joint_iter, n_iters, BA=True) == me.stage
        want = {}
        for grp in ref_opt.param_groups:
            for p in grp["params"]:
                want[id(p)] = float(grp["lr"])
        got = {id(p): lr for p, lr in zip(opt.small_opt.params, opt.small_opt.lrs)}
        got.update({id(r.param): lr for r, lr in opt.row_opt.groups})
        assert got == want, joint_iter
    assert len(want) == len(dec) + 4 + len(cams) + 1


def test_host_side_sampling_equals_the_oracle_bit_for_bit():
    """Renderer.surface_samples + _z_for_zero_depth (the torch expressions the sampling kernel is held to in the GPU
    suite) vs the oracle's surface_z_vals, which the fixtures pin to the reference (Renderer.py:131-208) with zero
    error: positive depths, missing depths (their range is a batch statistic), an all-missing batch, a batch where
    5 mean(d) undercuts 1.2 max(d)."""
    import types
    from oracle import hps_oracle as O
    r = hps.Renderer(hps.scannet_cfg("cpu"), None, types.SimpleNamespace(H=4, W=4, fx=1.0, fy=1.0, cx=2.0, cy=2.0))
    g = torch.Generator().manual_seed(12)
    batches = [0.5 + 4.0 * torch.rand(300, generator=g), torch.zeros(40), torch.cat([torch.full((90,), 0.02), torch.tensor([7.0])])]
    batches[0][torch.rand(300, generator=g) < 0.15] = 0.0
    batches[2][::7] = 0.0
    for gt in batches:
        z, nz = r.surface_samples(gt, "cpu")
        z = torch.where(nz[:, None], z, r._z_for_zero_depth(gt, "cpu")[None, :].expand_as(z))
        want, want_nz = O.surface_z_vals(gt)
        assert torch.equal(nz, want_nz) and torch.equal(z, want), float((z - want).abs().max())


def test_host_side_pose_expressions_equal_the_oracle():
    """The product's torch expressions for the pose / ray helpers (common.py:103-119, 304-346; what the GPU suite holds the
    pose kernels to) vs the oracle, which the tracker fixtures pin to the reference (pose gradient error 0.0):
    pose matrix, ray origins / directions and the gradient to the 7-vector, bit for bit."""
    from oracle import hps_oracle as O
    g = torch.Generator().manual_seed(14)
    K = O.SCANNET_INTRINSICS
    for _ in range(4):
        cam = torch.cat([torch.randn(4, generator=g), torch.randn(3, generator=g)])
        cam[:4] = cam[:4] / cam[:4].norm() * (1.0 + 0.05 * float(torch.randn((), generator=g)))     # not exactly unit
        i = torch.randint(0, K["W"], (50,), generator=g).float()
        j = torch.randint(0, K["H"], (50,), generator=g).float()
        outs = []
        for fn_cam, fn_rays in ((hps.get_camera_from_tensor, lambda i_, j_, c: hps.get_rays_from_uv(
                i_, j_, c, K["H"], K["W"], K["fx"], K["fy"], K["cx"], K["cy"], "cpu")),
                (O.camera_from_tensor, lambda i_, j_, c: O.rays_from_uv(i_, j_, c, K["fx"], K["fy"], K["cx"], K["cy"]))):
            c = cam.clone().requires_grad_(True)
            c2w = fn_cam(c)
            ro, rd = fn_rays(i, j, c2w)
            ((rd * torch.arange(1, 4)).sum() + (ro * torch.tensor([0.3, -0.2, 0.5])).sum()).backward()
            outs.append((c2w.detach(), ro.detach(), rd.detach(), c.grad.clone()))
        for a, b in zip(*outs):
            assert a.shape == b.shape and torch.equal(a, b)


def test_point_insertion_host_logic_matches_the_reference_class(monkeypatch, golden):
    """NeuralPointCloud.add_neural_points on CPU tensors with the neighbour search answered by the oracle (the one part
    that is a CUDA kernel; it is checked bit-exact in the GPU suite): the masking, the float32 sample arithmetic, the
    order of the appended points and the returned counts reproduce the fixture made by running the unmodified reference
    class (oracle/make_golden_insertion.py) - the body of the GPU test, run here."""
    from oracle import hps_oracle as O
    npm = hps.neural_point

    def radius_search(self, q, radius, want_cand=False):
        q = q.detach().float().reshape(-1, 3)
        D, I, n = O.find_neighbors(self._pos, q, radius)
        t_up, t_dn = O.radius_thresholds(radius, q.shape[0])
        return npm.NeighborQuery(I.int(), D, n, t_up, t_dn)

    monkeypatch.setattr(npm.GridIndex, "radius_search", radius_search)
    monkeypatch.setattr(ops, "_need_cuda", lambda *t: None)
    from tests import test_gpu_render as T
    monkeypatch.setattr(T, "DEV", "cpu")
    T.test_add_neural_points_matches_the_reference_class(hps, golden)


@pytest.mark.parametrize("min_chunk", [1, 300, 65536])
def test_render_img_host_logic_matches_the_reference_frame(monkeypatch, golden, min_chunk):
    """Renderer.render_img's own code - ray generation, the chunk loop, the per-reference-chunk sampling rows handed to
    wide chunks, the float64 images - with render_batch_ray answered by the oracle (the fused kernels are what the GPU suite
    checks): the reference's frame (tests/golden/render_img_small.pt) comes out for the reference's chunking (100 rays),
    for 300-ray chunks that straddle reference chunks, and for one wide chunk."""
    import types
    from oracle import hps_oracle as O
    fx, sc = golden("render_img_small.pt"), golden("scene_small.pt")
    cl = sc["clouds"]
    seen = []

    def render_batch_ray(self, npc, decoders, rays_d, rays_o, device, stage, gt_depth=None, npc_geo_feats=None,
                         npc_col_feats=None, is_tracker=False, cloud_pos=None, normals=None, dynamic_r_query=None,
                         exposure_feat=None, prepared=None, _loss=None, _z_zero_rows=None):
        seen.append((rays_o.shape[0], _z_zero_rows is not None))
        orig = O.surface_z_vals

        def patched(gt, *a, **k):
            z, nz = orig(gt, *a, **k)
            if _z_zero_rows is not None:
                assert _z_zero_rows.shape == z.shape and _z_zero_rows.is_contiguous()
                z[~nz] = _z_zero_rows[~nz]
            return z, nz
        O.surface_z_vals = patched
        try:
            return O.render_batch_ray({l: cl[l]["points"] for l in cl}, sc["decoders"], rays_d, rays_o, stage, gt_depth,
                                      npc_geo_feats, npc_col_feats, {l: v.reshape(-1) for l, v in dynamic_r_query.items()},
                                      sc["c_fill_geo"], sc["c_fill_col"], exposure_feat=exposure_feat)
        finally:
            O.surface_z_vals = orig

    monkeypatch.setattr(hps.Renderer, "render_batch_ray", render_batch_ray)
    small = hps.Renderer(hps.scannet_cfg("cpu"), None, types.SimpleNamespace(H=fx["H"], W=fx["W"], **fx["intrinsics"]),
                         ray_batch_size=fx["ray_batch_size"])
    small.min_render_chunk = min_chunk
    for stage, want in fx["stages"].items():
        del seen[:]
        d, u, c = small.render_img(None, None, fx["c2w"], "cpu", stage, gt_depth=fx["depth"].clone(),
                                   npc_geo_feats={l: cl[l]["geo_feats"] for l in cl},
                                   npc_col_feats={l: cl[l]["col_feats"] for l in cl},
                                   dynamic_r_query={l: fx["radius"][l].clone() for l in cl}, exposure_feat=sc["exposure_feat"])
        B = max(fx["ray_batch_size"], min_chunk)
        assert [n for n, _ in seen] == [min(B, fx["H"] * fx["W"] - i) for i in range(0, fx["H"] * fx["W"], B)]
        assert all(rows == (B > fx["ray_batch_size"]) for _, rows in seen)
        assert d.dtype == torch.float64 and u.dtype == torch.float64 and d.shape == (fx["H"], fx["W"])
        if min_chunk == 1:
            assert torch.equal(d, want["depth"]) and torch.equal(c, want["color"]) and torch.equal(u, want["uncertainty"])
        else:
            assert float((d - want["depth"]).abs().max()) < 1e-5 and float((c - want["color"]).abs().max()) < 1e-5
            assert float((u - want["uncertainty"]).abs().max()) < 1e-5 * float(want["uncertainty"].max())


def test_peer_allreduce_range_arguments_are_validated_on_the_host():
    """PeerAllReduce.__call__ hands the kernel shifted peer addresses for a sub-range: offsets that are not a multiple of
    four floats, ranges that leave the buffer and a multicast request without a multicast address are refused before
    anything is launched (the object is built without symmetric memory here: only the argument checks run)."""
    par = hps.parallel
    ex = object.__new__(par.PeerAllReduce)
    ex._lib, ex.n, ex.rank, ex.world, ex.multicast_ptr, ex._shift_cache = hps._lib, 4096, 0, 2, 0, {}
    ex.buffer = torch.zeros(4)
    for kw in (dict(offset_floats=6, n_floats=16), dict(offset_floats=-4, n_floats=16), dict(offset_floats=4092, n_floats=8),
               dict(n_floats=5000)):
        with pytest.raises(hps._lib.HpsError, match="PeerAllReduce"):
            ex("p2p", **kw)
    with pytest.raises(hps._lib.HpsError, match="multicast"):
        ex("multimem", n_floats=16)
    assert ex.default_variant() == "p2p" and ex.variants() == ["p2p"]
    ex.multicast_ptr, ex.world = 1 << 20, 8
    assert ex.default_variant() == "multimem" and ex.variants() == ["p2p", "multimem"]
