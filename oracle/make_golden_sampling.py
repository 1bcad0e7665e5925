"""TEST INFRASTRUCTURE ONLY.  Generate tests/golden/sampling_small.pt by running the UNMODIFIED reference helpers
(src/common.py: get_samples :236-258, get_sample_uv :159-171, select_uv :122-137, get_tensor_from_camera :357-382), imported from /root/reference under the stubs of oracle/ref_import.py, on
seeded synthetic inputs on the CPU.

Run in the build container only:   python oracle/make_golden_sampling.py
The fixture is committed; tests/test_host_logic.py replays it (same torch seed -> same torch.randint draw).
The gradient-guided variants (:174-233, :261-301) need skimage, which is not installed: they stay PARITY UNPINNED
and are tested against the Sobel restatement of oracle/hps_oracle.py instead.
"""
import math
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def rot(axis, angle):
    x, y, z = axis
    n = math.sqrt(x * x + y * y + z * z)
    x, y, z = x / n, y / n, z / n
    c, s = math.cos(angle), math.sin(angle)
    C = 1 - c
    return torch.tensor([[c + x * x * C, x * y * C - z * s, x * z * C + y * s],
                         [y * x * C + z * s, c + y * y * C, y * z * C - x * s],
                         [z * x * C - y * s, z * y * C + x * s, c + z * z * C]], dtype=torch.float64)


def main():
    ref = ref_import.load_reference().common
    g = torch.Generator().manual_seed(1219)
    H, W = 48, 64
    depth = 1.0 + 3.0 * torch.rand(H, W, generator=g)
    depth[torch.rand(H, W, generator=g) < 0.1] = 0.0
    depth[torch.rand(H, W, generator=g) < 0.1] = 6.0
    color = torch.rand(H, W, 3, generator=g)
    fx, fy, cx, cy = 57.7, 57.8, 31.9, 24.3
    c2w = torch.eye(4)
    c2w[:3, :3] = rot((0.3, -1.0, 0.2), 0.4).float()
    c2w[:3, 3] = torch.tensor([0.5, -0.25, 1.5])

    cases = []
    for seed, (H0, H1, W0, W1), n, filt, limit in [
            (1, (0, H, 0, W), 200, False, None), (2, (0, H, 0, W), 200, True, None),
            (3, (5, H - 5, 7, W - 7), 333, True, 5.0), (4, (10, 11, 0, W), 50, True, None),
            (5, (0, H, 20, 21), 64, False, None)]:
        torch.manual_seed(seed)
        out = ref.get_samples(H0, H1, W0, W1, n, H, W, fx, fy, cx, cy, c2w, depth, color, "cpu",
                              depth_filter=filt, return_index=True, depth_limit=limit)
        cases.append(dict(seed=seed, window=(H0, H1, W0, W1), n=n, depth_filter=filt, depth_limit=limit,
                          out=[o.clone() for o in out]))
    torch.manual_seed(9)
    uv = ref.get_sample_uv(3, 40, 2, 60, 100, depth, color, device="cpu")

    mats, quats = [], []
    poses = [rot((1, 0, 0), 0.0), rot((0.2, 0.5, -0.1), 0.3), rot((1, 0.01, 0.02), 3.1), rot((0.02, 1, -0.01), 3.12),
             rot((0.01, -0.02, 1), 3.13), rot((1, 1, 1), 2.0), rot((-0.3, 0.8, 0.5), 1.7), rot((1, 0, 0), math.pi)]
    for k, R in enumerate(poses):
        RT = torch.eye(4)
        RT[:3, :3] = R.float()
        RT[:3, 3] = torch.tensor([0.1 * k, -0.2 * k, 0.05 * k * k])
        mats.append(RT)
        quats.append((ref.get_tensor_from_camera(RT), ref.get_tensor_from_camera(RT, Tquad=True),
                      ref.get_tensor_from_camera(RT[:3].numpy())))

    torch.save(dict(H=H, W=W, depth=depth, color=color, intr=(fx, fy, cx, cy), c2w=c2w, cases=cases,
                    uv=[t.clone() for t in uv], mats=mats, quats=quats),
               os.path.join(GOLD, "sampling_small.pt"))
    print("wrote", os.path.join(GOLD, "sampling_small.pt"))


if __name__ == "__main__":
    main()
