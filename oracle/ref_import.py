"""TEST INFRASTRUCTURE ONLY - imports the *unmodified* reference modules from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  Used by
oracle/make_golden.py to (a) pin the oracle restatement in oracle/hps_oracle.py against the
reference's own code and (b) generate the committed fixtures under tests/golden/.

The reference imports faiss / open3d / skimage / turtle at module top (neural_point.py:6-9,
common.py:5-11, decoder.py:3); none is installed here, none is used by the hot path except faiss,
which is replaced by ExhaustiveIndex below = "the reference's neighbour query forced to exhaustive
search" (nprobe == nlist, neural_point.py:67,214): every vector is scanned with the direct-difference
squared L2 and the k smallest are returned in ascending (D, index) order.
"""
import os
import sys
import types
import importlib

import numpy as np
import torch

REF_ROOT = os.environ.get("HPS_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "src"))


class ExhaustiveIndex:
    """Stand-in for faiss.IndexIVFFlat(IndexFlatL2(3), 3, nlist) with nprobe == nlist.

    D_ij = fl(fl(fl(dx*dx) + fl(dy*dy)) + fl(dz*dz)), dx = fl(q_x - p_x), all fp32, no FMA
    (torch elementwise ops never contract).  Order: ascending (D, index) (stable sort).
    """

    def __init__(self, *a, **k):
        self.is_trained = False
        self.nprobe = 1
        self._pts = torch.zeros(0, 3)

    @property
    def ntotal(self):
        return self._pts.shape[0]

    def train(self, x):
        self.is_trained = True

    def add(self, x):
        self._pts = torch.cat([self._pts, torch.as_tensor(x).detach().float().cpu().reshape(-1, 3)], 0)

    def search(self, q, k):
        q = torch.as_tensor(q).detach().float().cpu().reshape(-1, 3)
        p = self._pts
        Ds, Is = [], []
        chunk = max(1, (1 << 26) // max(1, p.shape[0]))
        for s in range(0, q.shape[0], chunk):
            qq = q[s:s + chunk]
            dx = qq[:, None, 0] - p[None, :, 0]
            dy = qq[:, None, 1] - p[None, :, 1]
            dz = qq[:, None, 2] - p[None, :, 2]
            D = (dx * dx + dy * dy) + dz * dz
            Dk, Ik = torch.sort(D, dim=1, stable=True)
            Ds.append(Dk[:, :k].contiguous())
            Is.append(Ik[:, :k].contiguous())
        return torch.cat(Ds, 0), torch.cat(Is, 0)


def _install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    if "faiss" not in sys.modules:
        f = mod("faiss", METRIC_L2=1,
                StandardGpuResources=lambda *a, **k: object(),
                IndexFlatL2=lambda d: ("flat", d),
                IndexIVFFlat=lambda *a, **k: ExhaustiveIndex(),
                index_cpu_to_gpu=lambda res, dev, idx: idx)
        c = mod("faiss.contrib")
        t = mod("faiss.contrib.torch_utils")
        f.contrib = c
        c.torch_utils = t
    for name in ("open3d", "skimage", "skimage.color", "skimage.filters", "colorama",
                 "torchmetrics", "torchmetrics.image", "torchmetrics.image.lpip", "pytorch_msssim"):
        if name not in sys.modules:
            mod(name)
    sys.modules["skimage.color"].rgb2gray = lambda x: x
    sys.modules["skimage"].color = sys.modules["skimage.color"]
    sys.modules["skimage"].filters = sys.modules["skimage.filters"]
    if "turtle" not in sys.modules or not hasattr(sys.modules["turtle"], "window_height"):
        mod("turtle", window_height=lambda: 0)


def load_reference():
    """Returns a namespace with the reference's hot-path modules (unmodified source, imported)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    ns = types.SimpleNamespace()
    ns.common = importlib.import_module("src.common")
    ns.neural_point = importlib.import_module("src.neural_point")
    ns.renderer = importlib.import_module("src.utils.Renderer")
    ns.decoder = importlib.import_module("src.conv_onet.models.decoder")
    ns.config = importlib.import_module("src.config")
    return ns


def load_cfg(ns, scene="configs/ScanNet/scene0181.yaml"):
    """Effective ScanNet cfg (inherit_from is cwd-relative, src/config.py:26-37)."""
    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        cfg = ns.config.load_config(scene, "configs/point_slam.yaml")
    finally:
        os.chdir(cwd)
    cfg["mapping"]["device"] = "cpu"
    cfg["tracking"]["device"] = "cpu"
    return cfg
