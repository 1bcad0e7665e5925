"""TEST INFRASTRUCTURE ONLY.  Generate tests/golden/frustum_small.pt by EXECUTING the reference's own
`Mapper.get_mask_from_c2w` (src/Mapper.py:194-250: frustum feature selection = numpy projection + cv2.remap of the depth
image + depth test) with the real cv2 (4.13 in the build container).  The method is read from the reference file where
it lies (/root/reference, never copied into the repo), dedented and executed against a stand-in `self`; importing the
whole src.Mapper module would need open3d / skimage / matplotlib, which are not installed.

Run in the build container only:   python oracle/make_golden_frustum.py
The fixture is committed; tests/test_oracle_golden.py pins oracle.frustum_mask against it and tests/test_host_logic.py
the product's selection arithmetic (the same torch ops that run on the device).
"""
import os
import sys
import textwrap
import types

import cv2
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("HPS_REFERENCE_ROOT", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle import hps_oracle as O  # noqa: E402


def reference_method():
    with open(os.path.join(REF, "src", "Mapper.py")) as f:
        lines = f.readlines()
    src = textwrap.dedent("".join(lines[193:250]))
    assert src.startswith("def get_mask_from_c2w(self, c2w, depth_np, level):"), src[:80]
    assert src.rstrip().endswith("return np.where(mask)[0].tolist()"), src[-80:]
    ns = {"np": np, "cv2": cv2, "torch": torch}
    exec(compile(src, "Mapper.py:194-250", "exec"), ns)
    return ns["get_mask_from_c2w"]


def main():
    fn = reference_method()
    K = O.SCANNET_INTRINSICS
    g = torch.Generator().manual_seed(1219)
    depth = O.synthetic_depth().clone()                       # (480, 640) f32, 1.4 - 2.6 m
    depth[::7, ::5] = 0.0                                     # missing depth: those pixels take the maximum (:243-244)
    depth[100:140, 200:260] = 0.0
    # cloud: surface samples over 1.6x the frame (so that the crop edges cut through it), depth jitter that straddles
    # the `depth + 0.5` test, plus points behind the camera and far outside
    n = 24_000
    ext = 1.6
    u = torch.rand(n, generator=g) * (K["W"] - 1) * ext - (ext - 1) * K["W"] / 2
    v = torch.rand(n, generator=g) * (K["H"] - 1) * ext - (ext - 1) * K["H"] / 2
    d0 = O.synthetic_depth()[v.clamp(0, K["H"] - 1).long(), u.clamp(0, K["W"] - 1).long()]
    d = d0 * (0.9 + 0.2 * torch.rand(n, generator=g)) + torch.where(torch.rand(n, generator=g) < 0.3,
                                                                     0.3 + 0.4 * torch.rand(n, generator=g), torch.zeros(n))
    pts = torch.stack([(u - K["cx"]) / K["fx"] * d, -(v - K["cy"]) / K["fy"] * d, -d], -1).float()
    pts[:200, 2] *= -1.0                                      # behind the camera
    pts[200:300] *= 40.0                                      # far outside
    cases = []
    cams = [torch.tensor([1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]),
            torch.tensor([0.999, 0.02, -0.03, 0.01, 0.05, -0.02, 0.03]),
            torch.tensor([0.98, -0.1, 0.12, 0.05, -0.2, 0.1, -0.15])]
    for cam, edge in zip(cams, (-4, 20, 0)):
        R_t = O.camera_from_tensor(cam / torch.cat([cam[:4].norm().expand(4), torch.ones(3)]))
        c2w = torch.cat([R_t, torch.tensor([[0.0, 0.0, 0.0, 1.0]])], 0).float()
        world = (pts.double() @ c2w[:3, :3].double().T + c2w[:3, 3].double()).float()      # cloud expressed in world space
        cloud = world.tolist()                                                               # the reference's list of lists
        me = types.SimpleNamespace(H=K["H"], W=K["W"], fx=K["fx"], fy=K["fy"], cx=K["cx"], cy=K["cy"], frustum_edge=edge,
                                   npc=types.SimpleNamespace(cloud_pos=lambda level, c=cloud: c))
        idx = fn(me, c2w, depth.numpy().copy(), "fine")
        cases.append(dict(points=world, c2w=c2w, edge=edge, indices=torch.tensor(idx, dtype=torch.long)))
        print("edge %3d: %d of %d points selected" % (edge, len(idx), n))
    out = dict(depth=depth, intrinsics={k: K[k] for k in ("H", "W", "fx", "fy", "cx", "cy")}, cases=cases,
               cv2_version=cv2.__version__)
    path = os.path.join(GOLD, "frustum_small.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
