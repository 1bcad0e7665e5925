"""Per-frame preparation that feeds the render path: dynamic radius maps.

The reference builds, once per tracked / mapped frame and on the host, a per-pixel search radius from the colour
gradient (Tracker.py:297-325, Mapper.py:1026-1050): skimage rgb2gray -> sobel_h / sobel_v -> magnitude, clipped at
color_grad_threshold -> piecewise-linear map (scipy interp1d) [0, 0.01, thr] -> [r_max, r_max, r_min], evaluated in
float64, for every hierarchy level, once for insertion (r_add) and once for queries (ratio * r_add).  Here the same
arithmetic runs as a few tensor ops on the device the frame lives on; nothing is copied to the host.
"""
import torch
import torch.nn.functional as F

GRAY = (0.2125, 0.7154, 0.0721)          # skimage.color.rgb2gray


def color_gradient_magnitude(gt_color, threshold):
    """(H,W,3) colour image -> (H,W) float64 Sobel gradient magnitude clipped to [0, threshold] (the reference's frames
    are float64 numpy arrays at this point: datasets.py divides the uint8 image by 255.).
    skimage 0.19 semantics: sobel = correlate [1,0,-1] along the axis, [1,2,1]/4 across it, 'reflect' borders."""
    img = gt_color.to(torch.float64)
    gray = img[..., 0] * GRAY[0] + img[..., 1] * GRAY[1] + img[..., 2] * GRAY[2]
    g = gray[None, None]
    # scipy/skimage 'reflect' = symmetric padding (edge sample repeated) = torch 'replicate' for a 1-pixel halo
    g = F.pad(g, (1, 1, 1, 1), mode="replicate")
    edge = torch.tensor([1.0, 0.0, -1.0], device=img.device, dtype=torch.float64)
    smooth = torch.tensor([1.0, 2.0, 1.0], device=img.device, dtype=torch.float64) / 4.0
    k_h = (edge[:, None] * smooth[None, :])[None, None]       # derivative along rows (sobel_h)
    k_v = (smooth[:, None] * edge[None, :])[None, None]       # derivative along columns (sobel_v)
    gy = F.conv2d(g, k_h)[0, 0]
    gx = F.conv2d(g, k_v)[0, 0]
    return torch.sqrt(gx * gx + gy * gy).clamp(0.0, float(threshold))


def _interp3(x, x1, x2, y0, y2):
    """interp1d([0, x1, x2], [y0, y0, y2]) on 0 <= x <= x2, float64."""
    x = x.to(torch.float64)
    t = ((x - x1) / (x2 - x1)).clamp(min=0.0)
    return y0 + (y2 - y0) * t


def dynamic_radius(gt_color, radius_hierarchy, radius_query_ratio, color_grad_threshold):
    """-> (dynamic_r_add, dynamic_r_query): dicts level -> (H,W) float64 tensors on gt_color's device.
    radius_hierarchy[level] = {first key: r_add_max, second key: r_add_min} as in configs/point_slam.yaml:197-203."""
    mag = color_gradient_magnitude(gt_color, color_grad_threshold)
    r_add, r_query = {}, {}
    for level, d in radius_hierarchy.items():
        vals = list(d.values())
        r_max, r_min = float(vals[0]), float(vals[1])
        r_add[level] = _interp3(mag, 0.01, float(color_grad_threshold), r_max, r_min)
        r_query[level] = _interp3(mag, 0.01, float(color_grad_threshold), radius_query_ratio * r_max,
                                  radius_query_ratio * r_min)
    return r_add, r_query
