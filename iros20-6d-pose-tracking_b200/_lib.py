"""ctypes binding of libse3tn.so (the C ABI in include/se3tn.h).  No torch types cross it."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libse3tn.so')

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_STATE, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
PREC_TF32, PREC_FP32, PREC_BF16X3, PREC_BF16 = 0, 1, 2, 3
RENDER_VISPY, RENDER_PYRENDER = 0, 1
WEIGHT_BLOB_FLOATS = 13528326

_vp, _i, _d, _sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t

# name -> (restype, argtypes); mirrors include/se3tn.h one to one
SIGNATURES = {
    'se3tn_workspace_bytes': (_sz, [_i]),
    'se3tn_create': (_i, [_i, _i, _vp, C.POINTER(_vp)]),
    'se3tn_destroy': (None, [_vp]),
    'se3tn_last_error': (C.c_char_p, [_vp]),
    'se3tn_load_weights': (_i, [_vp, _i, _vp, _sz]),
    'se3tn_set_stats': (_i, [_vp, _i, _vp, _vp, _i]),
    'se3tn_preprocess': (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    'se3tn_normalize': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    'se3tn_compute_bbox': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    'se3tn_crop_bbox': (_i, [_vp, _vp, _vp, _i, _i, _vp, _i, _i, _i, _vp, _vp, _vp]),
    'se3tn_forward': (_i, [_vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp]),
    'se3tn_forward_preprocessed': (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp]),
    'se3tn_pose_update': (_i, [_vp, _vp, _vp, _vp, _d, _d, _vp, _i, _vp]),
    'se3tn_so3_log': (_i, [_vp, _vp, _vp, _d, _d, _vp, _vp, _i, _vp]),
    'se3tn_track_batch': (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _d, _d, _i, _vp, _vp, _vp, _vp]),
    'se3tn_add_adi': (_i, [_vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    'se3tn_vocap': (_i, [_vp, _vp, _i, C.POINTER(_d), _vp]),
    'se3tn_allgather_poses': (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    'se3tn_upload_frame_window': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    'se3tn_track_host': (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _d, _d, _i, _vp, _vp, _vp, _vp]),
    'se3tn_fill_depth': (_i, [_vp, _vp, _i, _i, _d, _vp, _vp, _vp]),
    'se3tn_fill_depth_ex': (_i, [_vp, _vp, _i, _i, _d, _i, _i, _vp, _vp, _vp]),
    'se3tn_set_mesh': (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i]),
    'se3tn_render': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    'se3tn_render_ex': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    'se3tn_debug_buffer': (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(_sz)]),
    'se3tn_last_launch_count': (_i, [_vp]),
    'se3tn_get_trace': (_i, [_vp, _vp]),
    'se3tn_last_step_was_graph': (_i, [_vp]),
    'se3tn_set_profiling': (_i, [_vp, _i]),
    'se3tn_get_profile': (_i, [_vp, _vp]),
}

_lib = None


def build_library(force=False):
    from . import build as _build
    return _build.build(force=force)


def load():
    """Load (building first if the .so is missing and nvcc exists).  Never falls back to anything:
    a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build_library()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class Se3tnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__('libse3tn error %d: %s' % (code, msg))
        self.code = code


def check(code, ctx=None):
    if code != OK:
        msg = load().se3tn_last_error(ctx)
        raise Se3tnError(code, msg.decode() if msg else '')
