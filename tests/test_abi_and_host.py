"""CPU-side checks: the C-ABI library builds, loads and exports every symbol include/hps_b200.h declares
(no compute calls without a GPU); host-side logic (thresholds, decoder construction, cfg)."""
import ctypes
import os
import random
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hps_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hps_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(hps):
    lib = ctypes.CDLL(hps._lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert set(names) == set(hps._lib.SIGNATURES), "ctypes table and header disagree"
    assert hps._lib.load().hps_version() >= 1
    assert hps._lib.load().hps_mlp_act_floats(1) == 1280 and hps._lib.load().hps_mlp_act_floats(0) == 320
    assert hps._lib.load().hps_mlp_packed_floats(1) == (40 + 128 * 3 + 168) * 128 + 5 * 32 * 128


def test_ops_refuse_cpu_tensors(hps):
    grid = None
    with pytest.raises(hps._lib.HpsError):
        hps.ops._need_cuda(torch.zeros(3))
    with pytest.raises(hps._lib.HpsError):
        hps.ops.composite(torch.zeros(2, 5, 4), torch.zeros(2, 5))


def test_radius_thresholds_match_oracle(hps):
    from oracle import hps_oracle as O
    r = torch.rand(1000, dtype=torch.float64) * 0.9 + 0.02
    r[::3] = 0.25
    up, dn = hps.ops.radius_thresholds(r.reshape(-1, 1), 1000, torch.device("cpu"))
    upo, dno = O.radius_thresholds(r)
    assert torch.equal(up, upo) and torch.equal(dn, dno)
    up, dn = hps.ops.radius_thresholds(0.08, 7, torch.device("cpu"))
    upo, dno = O.radius_thresholds(0.08, 7)
    assert torch.equal(up, upo) and torch.equal(dn, dno)
    with pytest.raises(ValueError):
        hps.ops.radius_thresholds(r, 5, torch.device("cpu"))


def test_decoders_have_reference_names_and_seeded_init(hps, golden):
    ref = golden("scene_small.pt")["decoders"]
    torch.manual_seed(1219)
    np.random.seed(1219)
    random.seed(1219)
    m = hps.get_model(hps.scannet_cfg("cpu"))
    assert sum(p.numel() for p in m.parameters()) == 260294
    for name, want in ref.items():
        dec = getattr(m, name)
        sd = dec.state_dict()
        assert set(sd) == set(want) - ({"embedder._B"} if name.startswith("color") else set())
        for k, v in want.items():
            if k.startswith("mlp_exposure.linear2"):
                continue      # the fixture rescales these on purpose
            mine = dec.embedder._B if k == "embedder._B" else sd[k]
            assert torch.equal(mine, v), (name, k)
        assert len(dec.kernel_params(torch.device("cpu"))) == hps.ops.N_PARAM


def test_unsupported_modes_fail_loudly(hps):
    cfg = hps.scannet_cfg("cpu")
    cfg["rendering"]["N_samples"] = 4
    with pytest.raises(NotImplementedError):
        hps.Renderer(cfg)
    cfg = hps.scannet_cfg("cpu")
    cfg["pointcloud"]["nn_weighting"] = "expo"
    with pytest.raises(NotImplementedError):
        hps.get_model(cfg)
