"""TEST INFRASTRUCTURE ONLY.  Generate tests/golden/gradient_small.pt by running the reference's own code for the two
places where the colour gradient steers the path:
  * `get_selected_index_with_grad` (src/common.py:199-233, imported unmodified): the tracker's pixel pre-selection in the
    ScanNet configuration (`sample_with_color_grad: True`, Tracker.py:326-332);
  * the per-frame dynamic radius maps (src/Tracker.py:297-323, statements executed where they lie in the reference tree).
Both call scikit-image (`rgb2gray`, `filters.sobel_h`, `filters.sobel_v`), which is NOT installed here: those three
functions - and only those - are stand-ins restated with scipy.ndimage after scikit-image 0.19.3 (the reference's pin,
env.yaml:136).  So the fixture pins the reference's selection / clipping / interpolation logic with the real numpy and
scipy; the three skimage functions themselves stay PARITY UNPINNED.

Run in the build container only:   python oracle/make_golden_gradient_sampling.py
"""
import os
import sys
import textwrap
import types

import numpy as np
import torch
from scipy import ndimage as ndi
from scipy.interpolate import interp1d

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def rgb2gray(img):                                   # skimage.color.rgb2gray: ITU-R 601-2 luma weights
    return np.asarray(img, dtype=np.float64) @ np.array([0.2125, 0.7154, 0.0721])


def _sobel(image, axis):                             # skimage.filters.sobel(image, axis=axis), mode='reflect'
    out = ndi.correlate1d(np.asarray(image, dtype=np.float64), [1, 0, -1], axis=axis, mode="reflect")
    return ndi.correlate1d(out, np.array([1, 2, 1]) / 4.0, axis=1 - axis, mode="reflect")


filters = types.SimpleNamespace(sobel_h=lambda im: _sobel(im, 0), sobel_v=lambda im: _sobel(im, 1))


def main():
    ref = ref_import.load_reference()
    cfg = ref_import.load_cfg(ref)
    ref.common.rgb2gray, ref.common.filters = rgb2gray, filters
    g = torch.Generator().manual_seed(5)
    H, W, n, ratio = 40, 56, 20, 15
    img = torch.rand(H, W, 3, generator=g, dtype=torch.float64)
    img[10:20, 10:30] = img[10:20, 10:30] * 0.05 + 0.4          # a flat patch: small gradients, large radii
    depth = 1 + 5 * torch.rand(H, W, generator=g)
    depth[torch.rand(H, W, generator=g) < 0.2] = 0
    out = dict(image=img, depth=depth, n=n, ratio=ratio, window=(4, H - 4, 6, W - 6), selected={})
    for limit in (False, True):
        sel, mag = ref.common.get_selected_index_with_grad(4, H - 4, 6, W - 6, n, img, ratio=ratio, gt_depth=depth,
                                                           depth_limit=limit)
        out["selected"][limit] = torch.from_numpy(np.sort(sel))
        out["grad_mag"] = torch.from_numpy(mag)
        print("depth_limit=%-5s: %d of the top %d pixels selected" % (limit, len(sel), ratio * n))
    sel, _ = ref.common.get_selected_index_with_grad(4, H - 4, 6, W - 6, n, img, ratio=ratio)
    out["selected"][None] = torch.from_numpy(np.sort(sel))

    # dynamic radius maps: Tracker.py:297-323 (the torch.save debug line :325 is left out)
    with open(os.path.join(ref_import.REF_ROOT, "src", "Tracker.py")) as f:
        src = textwrap.dedent("".join(f.readlines()[296:323]))
    assert src.startswith("if self.use_dynamic_radius:") and src.rstrip().endswith("self.device)"), (src[:60], src[-60:])
    pc = cfg["pointcloud"]
    me = types.SimpleNamespace(use_dynamic_radius=True, radius_query_ratio=pc["radius_query_ratio"],
                               color_grad_threshold=pc["color_grad_threshold"], radius_hierarchy=pc["radius_hierarchy"],
                               dynamic_r_add={}, dynamic_r_query={}, device="cpu")
    exec(compile(src, "Tracker.py:297-323", "exec"),
         dict(self=me, gt_color=img, rgb2gray=rgb2gray, filters=filters, np=np, interp1d=interp1d, torch=torch))
    out["radius_hierarchy"] = pc["radius_hierarchy"]
    out["radius_query_ratio"], out["color_grad_threshold"] = pc["radius_query_ratio"], pc["color_grad_threshold"]
    out["dynamic_r_add"], out["dynamic_r_query"] = me.dynamic_r_add, me.dynamic_r_query
    for level in me.dynamic_r_add:
        print("%s: r_add %.4f..%.4f  r_query %.4f..%.4f" % (level, float(me.dynamic_r_add[level].min()),
              float(me.dynamic_r_add[level].max()), float(me.dynamic_r_query[level].min()),
              float(me.dynamic_r_query[level].max())))
    path = os.path.join(GOLD, "gradient_small.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
