"""hps_b200: B200-native render-and-optimise path for Hierarchical Point-SLAM.

Host-side mirror of the reference's Python interface for this path (same class / function names and
argument meaning) over hand-written sm_100a kernels behind the C ABI in include/hps_b200.h.
Importable as `hps_b200` (repo-root alias module) or importlib.import_module("hierarchical-point-slam_b200").
"""
from . import _lib, frame, graphs, mapper_optim, ops, parallel, tracker_optim  # noqa: F401
from .config import scannet_cfg  # noqa: F401
from .common import (get_camera_from_tensor, get_rays, get_rays_from_uv, quad2rotation,  # noqa: F401
                     raw2outputs_nerf_color)
from .decoder import NICER, PointDecoder  # noqa: F401
from .losses import mapper_loss, mapper_loss_torch, tracker_loss, tracker_loss_torch  # noqa: F401
from .neural_point import GridIndex, NeuralPointCloud  # noqa: F401
from .mapper_optim import MapperOptimizer  # noqa: F401
from .tracker_optim import TrackerOptimizer  # noqa: F401
from .ops import FusedAdam, LiveRows, RowAdam  # noqa: F401
from .renderer import Renderer  # noqa: F401


def get_model(cfg):
    """config.get_model (reference src/conv_onet/config.py): NICER from a loaded cfg dict."""
    return NICER(cfg, dim=cfg["data"]["dim"], c_dim=cfg["model"]["c_dim"],
                 pos_embedding_method=cfg["model"]["pos_embedding_method"], use_normals=cfg["use_normals"],
                 use_view_direction=cfg["use_view_direction"])
