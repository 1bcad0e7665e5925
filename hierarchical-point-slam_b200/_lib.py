"""ctypes binding of libhps_b200.so (the C ABI declared in include/hps_b200.h).

There is no CPU or PyTorch fallback: if the shared object is missing, cannot be loaded, or a tensor is
not on a CUDA device, the ops raise.  `python hierarchical-point-slam_b200/build.py` (or
__graft_entry__.build()) produces the library in-tree with nvcc for sm_100a.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhps_b200.so")

_vp, _i64, _i32, _f32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float

# name -> (restype, argtypes); mirrors include/hps_b200.h one to one
SIGNATURES = {
    "hps_last_error": (ctypes.c_char_p, []),
    "hps_version": (_i32, []),
    "hps_grid_create": (_i32, [ctypes.POINTER(_vp)]),
    "hps_grid_destroy": (_i32, [_vp]),
    "hps_grid_build": (_i32, [_vp, _vp, _i64, _f32, _vp]),
    "hps_grid_append": (_i32, [_vp, _vp, _i64, _i64, ctypes.POINTER(_i32), _vp]),
    "hps_grid_info": (_i32, [_vp, ctypes.POINTER(_i64), ctypes.POINTER(_i64), ctypes.POINTER(_i64),
                             ctypes.POINTER(_f32), ctypes.POINTER(_i32), _vp]),
    "hps_knn_query": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i32, _vp]),
    "hps_interp_fwd": (_i32, [_vp] * 10 + [_i64, _i32] + [_vp] * 5),
    "hps_interp_bwd": (_i32, [_vp] * 11 + [_i64] + [_vp] * 5),
    "hps_rows_store": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _vp]),
    "hps_mlp_act_floats": (_i64, [_i32]),
    "hps_mlp_packed_floats": (_i64, [_i32]),
    "hps_mlp_pack": (_i32, [_i32, _vp, _vp, _vp]),
    "hps_mlp_fwd": (_i32, [_i32, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "hps_mlp_tc_packed_bytes": (_i64, [_i32]),
    "hps_mlp_tc_pack": (_i32, [_i32, _vp, _vp, _vp]),
    "hps_mlp_tc_fwd": (_i32, [_i32, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i32, _vp]),
    "hps_mlp_tc_ws_bytes": (_i64, [_i32, _i64]),
    "hps_mlp_tc_bwd_packed_bytes": (_i64, [_i32]),
    "hps_mlp_tc_bwd_absmax_offset": (_i64, [_i32]),
    "hps_mlp_tc_bwd_pack": (_i32, [_i32, _vp, _vp, _vp]),
    "hps_mlp_tc_bwd": (_i32, [_i32, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "hps_tc_selftest": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "hps_tc_trace": (_i32, [_vp]),
    "hps_mlp_bwd": (_i32, [_i32, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "hps_composite_fwd": (_i32, [_vp, _i32, _vp, _i32, _vp, _vp, _i64, _i32, _f32, _vp, _vp, _vp, _vp, _i32, _vp,
                                 _vp, _vp, _i32, _vp]),
    "hps_composite_bwd": (_i32, [_vp, _i32, _vp, _i32, _vp, _vp, _i64, _i32, _f32, _vp, _vp, _vp, _vp, _i32, _vp,
                                 _i32, _vp, _vp, _vp]),
    "hps_composite_mapper_loss": (_i32, [_vp, _i32, _vp, _i32, _vp, _vp, _i64, _i32, _f32, _i32, _vp, _vp, _vp, _vp,
                                         _f32] + [_vp] * 13 + [_i64, _vp]),
    "hps_camera_from_tensor": (_i32, [_vp, _vp, _vp]),
    "hps_camera_from_tensor_bwd": (_i32, [_vp, _vp, _vp, _vp]),
    "hps_rays_from_uv": (_i32, [_vp, _vp, _vp, _i32, _i64, _f32, _f32, _f32, _f32, _vp, _vp, _vp]),
    "hps_rays_from_uv_bwd": (_i32, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _vp, _vp]),
    "hps_adam_rows": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _f32, _f32, _f32, _f32, _vp, _vp, _vp, _vp]),
    "hps_exposure_fwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "hps_exposure_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hps_sample_rays": (_i32, [_vp, _vp, _vp, _vp, ctypes.c_double, _vp, _i64, _i64, _i32, _f32, _f32, _vp, _vp, _vp, _vp,
                               _vp]),
    "hps_sample_knn": (_i32, [_vp, _vp, _vp, _vp, _vp, ctypes.c_double, _vp, _i64, _i64, _i32, _f32, _f32] + [_vp] * 9),
    "hps_sample_rays_bwd": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "hps_adam_multi": (_i32, [_i32, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _f32, _vp, _i32, _vp]),
    "hps_adam_multi_own_steps": (_i32, [_i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _f32, _vp, _i32, _vp]),
    "hps_mapper_loss": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _i64, _vp, _vp, _vp, _vp, _vp]),
    "hps_tracker_loss_stats": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "hps_peer_allreduce_ctas": (_i32, []),
    "hps_peer_allreduce_pad_words": (_i64, []),
    "hps_peer_allreduce": (_i32, [_vp, _vp, _vp, _i32, _i32, _i64, _vp]),
    "hps_tracker_loss": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_double, _f32, _i64, _vp, _vp, _vp, _vp]),
}

_lib = None


class HpsError(RuntimeError):
    pass


def load():
    """Load (once) and type the shared library.  Raises HpsError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HpsError("libhps_b200.so not built (%s missing): run `python hierarchical-point-slam_b200/build.py`; "
                       "there is no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().hps_last_error()
        raise HpsError("hps_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())
