"""ctypes binding of libplb200.so (include/plb200.h).

There is NO fallback: if the shared library is missing or cannot be loaded the
import of any product module raises.  The library is built in-tree by
`plb200/build.py` (nvcc, sm_100a) and travels to the GPU box with the repo.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libplb200.so")

MAX_SRC, MAX_SCALES, MAX_JOBS = 4, 4, 2
ROT_AXISANGLE, ROT_EULER = 0, 1
PHOTO_L1_MEAN, PHOTO_MIN_REPROJ = 0, 1
INPUT_DISP, INPUT_DEPTH, INPUT_LOGIT = 0, 1, 2
PHOTO_NO_SSIM, PHOTO_NO_AUTOMASK, PHOTO_CLIP = 1, 2, 4

_fp = C.c_void_p  # device pointers are passed as integers


class PhotoJob(C.Structure):
    _fields_ = [
        ("tgt", _fp),
        ("src", _fp * MAX_SRC),
        ("pose_index", C.c_int32 * MAX_SRC),
        ("pose_inv", C.c_int32 * MAX_SRC),
        ("disp", _fp * MAX_SCALES),
        ("dh", C.c_int32 * MAX_SCALES),
        ("dw", C.c_int32 * MAX_SCALES),
        ("g_disp", _fp * MAX_SCALES),
        ("g_src", _fp * MAX_SRC),
        ("g_tgt", _fp),
        ("n_src", C.c_int32),
        ("n_scales", C.c_int32),
        ("term_weight", C.c_float),
        ("mode", C.c_int32),
        ("flags", C.c_uint32),
        ("clip_loss", C.c_float),
    ]


class PhotoArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("n_jobs", C.c_int32),
        ("n_pose", C.c_int32),
        ("rotation_mode", C.c_int32),
        ("k_is_f64", C.c_int32),
        ("input_is_depth", C.c_int32),
        ("disp_a", C.c_float), ("disp_b", C.c_float),
        ("head_alpha", C.c_float), ("head_beta", C.c_float),
        ("want_grad", C.c_int32),
        ("deterministic", C.c_int32),
        ("sm_limit", C.c_int32),
        ("reserved2", C.c_int32),
        ("poses", _fp),
        ("K", _fp),
        ("g_poses", _fp),
        ("loss", _fp),
        ("upstream", _fp),
        ("skip_if_unit", _fp * 2),
        ("workspace", _fp),
        ("workspace_bytes", C.c_size_t),
        ("jobs", PhotoJob * MAX_JOBS),
    ]


class SmoothArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32),
        ("n_scales", C.c_int32),
        ("disp", _fp * MAX_SCALES),
        ("dh", C.c_int32 * MAX_SCALES),
        ("dw", C.c_int32 * MAX_SCALES),
        ("g_disp", _fp * MAX_SCALES),
        ("accumulate", C.c_int32),
        ("input_is_depth", C.c_int32),
        ("disp_a", C.c_float), ("disp_b", C.c_float),
        ("head_alpha", C.c_float), ("head_beta", C.c_float),
        ("scale_decay", C.c_float),
        ("want_grad", C.c_int32),
        ("loss", _fp),
        ("upstream", _fp),
        ("skip_if_unit", _fp * 2),
        ("workspace", _fp),
        ("workspace_bytes", C.c_size_t),
    ]


class EdgeArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("n_scales", C.c_int32),
        ("tgt", _fp),
        ("disp", _fp * MAX_SCALES),
        ("dh", C.c_int32 * MAX_SCALES),
        ("dw", C.c_int32 * MAX_SCALES),
        ("g_disp", _fp * MAX_SCALES),
        ("g_scratch", _fp * MAX_SCALES),
        ("accumulate", C.c_int32),
        ("normalize", C.c_int32),
        ("want_grad", C.c_int32),
        ("reserved", C.c_int32),
        ("loss", _fp),
        ("upstream", _fp),
        ("workspace", _fp),
        ("workspace_bytes", C.c_size_t),
        ("skip_if_unit", _fp * 2),
    ]


class WarpArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("rotation_mode", C.c_int32),
        ("pose_inv", C.c_int32),
        ("k_is_f64", C.c_int32),
        ("img", _fp),
        ("depth", _fp),
        ("pose", _fp),
        ("pose_stride", C.c_int32),
        ("reserved", C.c_int32),
        ("K", _fp),
        ("out", _fp),
        ("g_out", _fp),
        ("g_img", _fp),
        ("g_depth", _fp),
        ("g_pose", _fp),
        ("workspace", _fp),
        ("workspace_bytes", C.c_size_t),
    ]


class PhotomapArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("x", _fp),
        ("y", _fp),
        ("C1", C.c_float), ("C2", C.c_float),
        ("w_ssim", C.c_float), ("w_l1", C.c_float),
        ("clip", C.c_float),
        ("reserved", C.c_int32),
        ("out", _fp),
        ("threshold", _fp),
        ("g_out", _fp),
        ("g_x", _fp),
        ("g_y", _fp),
        ("workspace", _fp),
        ("workspace_bytes", C.c_size_t),
    ]


class CloudArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("sparsity", C.c_int32),
        ("depth", _fp),
        ("P", C.c_double * 12),
        ("Tinv", C.c_double * 16),
        ("cloud_f64", _fp),
        ("cloud_f32", _fp),
        ("index", _fp),
        ("valid", _fp),
        ("count", _fp),
        ("workspace", _fp),
        ("workspace_bytes", C.c_size_t),
    ]


class VeloArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("point_stride", C.c_int32), ("reserved", C.c_int32),
        ("points", _fp),
        ("counts", _fp),
        ("T", C.c_double * 16),
        ("P", C.c_double * 12),
        ("depth_f64", _fp),
        ("depth_f32", _fp),
        ("winner", _fp),
        ("workspace", _fp),
        ("workspace_bytes", C.c_size_t),
    ]


class PrepArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("in_h", C.c_int32), ("in_w", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("n_K", C.c_int32),
        ("frames", _fp),
        ("mean", C.c_float * 3), ("stdev", C.c_float * 3),
        ("out_planar", _fp), ("out_nhwc4", _fp),
        ("K_in", _fp), ("K_out", _fp),
        ("workspace", _fp), ("workspace_bytes", C.c_size_t),
    ]


# every symbol include/plb200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "plb_photo_workspace_bytes": (C.c_size_t, [C.POINTER(PhotoArgs)]),
    "plb_photo_loss": (C.c_int, [C.POINTER(PhotoArgs), C.c_void_p]),
    "plb_smooth_workspace_bytes": (C.c_size_t, [C.POINTER(SmoothArgs)]),
    "plb_smooth_loss": (C.c_int, [C.POINTER(SmoothArgs), C.c_void_p]),
    "plb_edge_smooth_workspace_bytes": (C.c_size_t, [C.POINTER(EdgeArgs)]),
    "plb_edge_smooth_loss": (C.c_int, [C.POINTER(EdgeArgs), C.c_void_p]),
    "plb_warp_workspace_bytes": (C.c_size_t, [C.POINTER(WarpArgs)]),
    "plb_warp_forward": (C.c_int, [C.POINTER(WarpArgs), C.c_void_p]),
    "plb_warp_backward": (C.c_int, [C.POINTER(WarpArgs), C.c_void_p]),
    "plb_reconstruct": (C.c_int, [_fp, _fp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _fp, C.c_void_p]),
    "plb_project": (C.c_int, [_fp, _fp, C.c_int32, _fp, C.c_int32, C.c_int32, C.c_int32, _fp, C.c_void_p]),
    "plb_pose_matrix": (C.c_int, [_fp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _fp, C.c_void_p]),
    "plb_pose_matrix_backward": (C.c_int, [_fp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _fp, _fp, C.c_void_p]),
    "plb_disp_to_depth": (C.c_int, [_fp, C.c_int64, C.c_float, C.c_float, _fp, C.c_void_p]),
    "plb_disp_to_depth_backward": (C.c_int, [_fp, _fp, C.c_int64, C.c_float, C.c_float, _fp, C.c_void_p]),
    "plb_photometric_map_workspace_bytes": (C.c_size_t, [C.POINTER(PhotomapArgs)]),
    "plb_photometric_map": (C.c_int, [C.POINTER(PhotomapArgs), C.c_void_p]),
    "plb_photometric_map_backward": (C.c_int, [C.POINTER(PhotomapArgs), C.c_void_p]),
    "plb_cloud_workspace_bytes": (C.c_size_t, [C.POINTER(CloudArgs)]),
    "plb_cloud_project": (C.c_int, [C.POINTER(CloudArgs), C.c_void_p]),
    "plb_velo_workspace_bytes": (C.c_size_t, [C.POINTER(VeloArgs)]),
    "plb_velo_project": (C.c_int, [C.POINTER(VeloArgs), C.c_void_p]),
    "plb_prep_workspace_bytes": (C.c_size_t, [C.POINTER(PrepArgs)]),
    "plb_prep_frames": (C.c_int, [C.POINTER(PrepArgs), C.c_void_p]),
    "plb_version": (C.c_char_p, []),
    "plb_launch_count": (C.c_uint64, []),
}

_ERRORS = {-1: "PLB_EINVAL: bad shape / count / flag", -2: "PLB_ENULL: a required pointer is NULL",
           -3: "PLB_EWORKSPACE: workspace missing or too small"}


class PlbError(RuntimeError):
    pass


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "libplb200.so not found at %s - build it with `python __graft_entry__.py` "
            "(or plb200/build.py); there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc, what):
    """Mirror how the reference surfaces errors: Python exceptions only."""
    if rc == 0:
        return
    if rc < 0:
        raise PlbError("%s: %s" % (what, _ERRORS.get(rc, "error %d" % rc)))
    raise PlbError("%s: CUDA error %d" % (what, rc))


def version():
    return lib.plb_version().decode()


def launch_count():
    return int(lib.plb_launch_count())
