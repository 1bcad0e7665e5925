"""Mapper / Tracker losses of the hot path (reference Mapper.py:791-856, Tracker.py:213-231).

The default entry points run the fused loss+gradient kernels (ops.mapper_loss / ops.tracker_loss: no boolean
indexing, no host synchronisation).  The *_torch variants are the same formulas as small torch expressions
(used by tests as a cross-check of the kernels' autograd wiring)."""
import torch

from . import ops

mapper_loss = ops.mapper_loss
tracker_loss = ops.tracker_loss


def mapper_loss_torch(depth, color, valid_ray_mask, gt_depth, gt_color, stage, w_color_loss=0.1, affine=None,
                      frame_index=None):
    """affine: (12,) exposure tensor of the (single) frame the batch comes from, (F,12) with frame_index (R,) for a
    batch that mixes rays of F keyframes (Mapper.py:805-822), or None."""
    mask = (gt_depth > 0) & valid_ray_mask & (~torch.isnan(depth))
    loss = torch.abs(gt_depth[mask] - depth[mask]).sum()
    if stage.startswith("color"):
        if affine is not None and frame_index is not None:
            a = affine.reshape(-1, 12)[frame_index.long()]
            color = torch.einsum("rk,rkc->rc", color, a[:, :9].reshape(-1, 3, 3)) + a[:, 9:]
        elif affine is not None:
            color = torch.matmul(color, affine[:9].reshape(3, 3)) + affine[-3:]
        color = torch.sigmoid(color)
        loss = loss + w_color_loss * torch.abs(gt_color[mask] - color[mask]).sum()
    return loss


def tracker_loss_torch(depth, uncertainty, color, gt_depth, gt_color, w_color_loss=0.5):
    uncertainty = uncertainty.detach()
    nan_mask = (~torch.isnan(depth)) & (~torch.isnan(uncertainty))
    tmp = torch.abs(gt_depth - depth) / torch.sqrt(uncertainty + 1e-10)
    mask = (tmp < 10 * tmp.mean()) & (gt_depth > 0) & nan_mask
    geo_loss = torch.clamp(tmp, min=0.0, max=1e3)[mask].sum()
    color_loss = torch.abs(gt_color - color)[mask].sum()
    return geo_loss + w_color_loss * color_loss
