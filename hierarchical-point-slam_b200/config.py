"""Effective hot-path configuration of the reference's ScanNet setup, as plain data.

Values are what `load_config('configs/ScanNet/sceneXXXX.yaml', 'configs/point_slam.yaml')` yields in the
reference after the inherit/merge ("Config / flags" row - including the quirk that
scannet.yaml's first `pointcloud:` key is dropped by PyYAML).  Only keys the hot path reads are listed;
a full reference cfg dict can be passed to every class instead.
"""
import copy

# configs/point_slam.yaml:105-146 (the `init` and `stage` tables hold the same numbers)
_STAGE_LRS = {
    "geometry_mid": {"decoders_lr": 0.001, "geometry_mid_lr": 0.03, "geometry_fine_lr": 0.03, "color_lr": 0.0},
    "geometry_fine": {"decoders_lr": 0.001, "geometry_mid_lr": 0.03, "geometry_fine_lr": 0.03, "color_lr": 0.0},
    "color_mid": {"decoders_lr": 0.005, "geometry_mid_lr": 0.005, "geometry_fine_lr": 0.005, "color_lr": 0.005},
    "color_fine": {"decoders_lr": 0.005, "geometry_mid_lr": 0.005, "geometry_fine_lr": 0.005, "color_lr": 0.005},
}

_SCANNET = {
    "occupancy": True,
    "use_normals": False,
    "use_view_direction": False,
    "use_dynamic_radius": True,
    "use_bound": False,
    "pts_along_ray": False,
    "data": {"dim": 3},
    "model": {"c_dim": 32, "exposure_dim": 8, "pos_embedding_method": "fourier", "encode_rel_pos_in_col": False,
              "encode_rel_pos_in_geo": False, "encode_exposure": True, "encode_viewd": False},
    "pointcloud": {"nn_search_method": "faiss", "nn_num": 8, "min_nn_num": 2, "N_add": 3,
                   "nn_weighting": "distance", "radius_add": 0.04, "radius_min": 0.02, "radius_query": 0.08,
                   "radius_mesh": 0.08,
                   "radius_hierarchy": {"fine": {"radius_add_max_fine": 0.08, "radius_add_min_fine": 0.02},
                                        "mid": {"radius_add_max_mid": 0.48, "radius_add_min_mid": 0.12}},
                   "radius_query_ratio": 2, "color_grad_threshold": 0.15, "near_end_surface": 0.96,
                   "far_end_surface": 1.04, "add_along_normals": False,
                   "fix_interval_when_add_along_ray": False, "nlist": 400, "nprobe": 4},
    "rendering": {"N_samples": 0, "N_surface": 5, "N_importance": 0, "lindisp": False, "perturb": 0.0,
                  "near_end": 0.3, "near_end_surface": 0.96, "far_end_surface": 1.04,
                  "sigmoid_coef_tracker": 0.1, "sigmoid_coef_mapper": 0.1, "sample_near_pcl": False,
                  "skip_zero_depth_pixel": False},
    "cam": {"H": 480, "W": 640, "fx": 577.590698, "fy": 578.729797, "cx": 318.905426, "cy": 242.683609,
            "crop_edge": 10},
    "mapping": {"device": "cuda:0", "w_color_loss": 0.1, "pixels": 10000, "iters": 600, "iters_first": 500, "BA": False,
                "BA_cam_lr": 0.0002, "geo_iter_ratio": 0.3, "mid_iter_ratio": 0.5, "geo_iter_first": 200,
                "min_iter_ratio": 0.95, "fix_geo_decoder_mid": True, "fix_geo_decoder_fine": True,
                "fix_color_decoder": False, "frustum_feature_selection": True, "color_refine": False,
                "init": _STAGE_LRS, "stage": _STAGE_LRS},
    "tracking": {"device": "cuda:0", "w_color_loss": 0.5, "pixels": 5000, "iters": 100, "lr": 0.0005,
                 "handle_dynamic": True, "use_color_in_tracking": True},
}


def scannet_cfg(device="cuda:0"):
    cfg = copy.deepcopy(_SCANNET)
    cfg["mapping"]["device"] = device
    cfg["tracking"]["device"] = device
    return cfg
