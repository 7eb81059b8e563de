"""ctypes binding of csrc/libg2s_b200.so (C ABI declared in include/g2s_b200.h).

There is NO fallback: if the shared library is missing or a call returns an error code, a RuntimeError is
raised (the reference's neural_renderer extension raises RuntimeError through AT_ASSERTM the same way).
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# G2S_LIB: an alternative build of the same library (kernel A/B experiments); the default is the in-tree build
LIB_PATH = os.environ.get("G2S_LIB") or os.path.join(HERE, "csrc", "libg2s_b200.so")

_c_int, _c_long, _c_float, _vp = ctypes.c_int, ctypes.c_long, ctypes.c_float, ctypes.c_void_p


class Camera(ctypes.Structure):
    """struct g2s_camera (include/g2s_b200.h)."""
    _fields_ = [("K", _c_float * 9), ("inv_K", _c_float * 9), ("rot_center_depth", _c_float),
                ("near_z", _c_float), ("far_z", _c_float), ("clamp_lo", _c_float), ("clamp_hi", _c_float),
                ("image_size", ctypes.c_int32), ("K_grid", _c_float * 9)]


_CAMP = ctypes.POINTER(Camera)


class PhotoLoss(ctypes.Structure):
    """g2s_photo_loss (include/g2s_b200.h): the fused masked photometric loss of model.py:265-274."""
    _fields_ = [("target", ctypes.c_void_p), ("view_mask", ctypes.c_void_p), ("depth_thresh", ctypes.c_float)]


_LOSSP = ctypes.POINTER(PhotoLoss)

# name -> (restype, argtypes); must list every symbol include/g2s_b200.h declares
SIGNATURES = {
    "g2s_version": (_c_int, []),
    "g2s_error_string": (ctypes.c_char_p, [_c_int]),
    "g2s_context_create": (_c_int, [ctypes.POINTER(_vp)]),
    "g2s_context_destroy": (None, [_vp]),
    "g2s_workspace_bytes": (ctypes.c_size_t, [_c_int, _c_int, _c_int]),
    "g2s_zbuffer_bytes": (ctypes.c_size_t, [_c_int, _c_int]),
    "g2s_zbuffer_init": (_c_int, [_vp, _c_int, _c_int, _c_float, _vp]),
    "g2s_warp_depth_fwd": (_c_int, [_CAMP, _vp, _c_long, _vp, _vp, _c_int, _vp, _vp, _vp, _vp]),
    "g2s_warp_depth_bwd": (_c_int, [_CAMP, _vp, _c_long, _vp, _vp, _c_int, _vp, _vp, _vp, _vp, _vp, _c_long, _vp,
                                    _vp, _vp]),
    "g2s_warp_grid_fwd": (_c_int, [_CAMP, _vp, _c_long, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _vp, _vp]),
    "g2s_warp_grid_bwd": (_c_int, [_CAMP, _vp, _c_long, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp,
                                   _vp]),
    "g2s_normal_fwd": (_c_int, [_CAMP, _vp, _c_int, _c_int, _c_int, _vp, _vp]),
    "g2s_normal_bwd": (_c_int, [_CAMP, _vp, _c_int, _c_int, _c_int, _vp, _vp, _c_int, _vp]),
    "g2s_sample_fwd": (_c_int, [_vp, _c_long, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                _vp, _vp]),
    "g2s_sample_bwd": (_c_int, [_vp, _c_long, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                _c_int, _vp, _c_long, _vp, _vp]),
    "g2s_chunk_views": (_c_int, [_c_int]),
    "g2s_chunk_views_bwd": (_c_int, [_c_int]),
    "g2s_render_fused_fwd": (_c_int, [_vp, _CAMP, _vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _vp, _c_int, _vp, _vp,
                                      _vp, _vp, _vp, _vp, _vp, _vp]),
    "g2s_render_fused_bwd": (_c_int, [_vp, _CAMP, _vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp,
                                      _vp, _vp, _c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "g2s_render_fused_loss_fwd": (_c_int, [_vp, _CAMP, _vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _vp, _c_int, _vp,
                                           _vp, _vp, _vp, _LOSSP, _vp, _vp, _vp, _vp]),
    "g2s_render_fused_loss_bwd": (_c_int, [_vp, _CAMP, _vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp,
                                           _vp, _LOSSP, _vp, _vp, _vp, _c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "g2s_render_rgb_fwd": (_c_int, [_CAMP, _vp, _vp, _c_long, _c_int, _c_int, _c_int, ctypes.POINTER(_c_float),
                                    _c_int, _vp, _vp, _vp, _vp]),
    "g2s_render_depth_fwd": (_c_int, [_CAMP, _vp, _c_int, _vp, _vp, _vp, _vp]),
    "g2s_render_depth_bwd": (_c_int, [_CAMP, _vp, _c_int, _vp, _vp, _vp, _vp, _vp]),
    "g2s_render_rgb_bwd": (_c_int, [_CAMP, _vp, _vp, _c_long, _c_int, _c_int, _c_int, ctypes.POINTER(_c_float), _c_int,
                                    _vp, _vp, _vp, _c_long, _vp, _vp, _vp, _vp]),
    "g2s_view_fwd": (_c_int, [_vp, _c_int, _c_int, _vp, _vp, _vp]),
    "g2s_view_bwd": (_c_int, [_vp, _c_int, _c_int, _vp, _vp, _vp, _vp]),
    "g2s_light_fwd": (_c_int, [_vp, _c_int, _vp, _vp]),
    "g2s_light_bwd": (_c_int, [_vp, _c_int, _vp, _vp, _vp]),
    "g2s_view_light_fwd": (_c_int, [_vp, _c_int, _vp, _c_int, _vp, _vp, _vp, _vp]),
    "g2s_view_light_bwd": (_c_int, [_vp, _c_int, _vp, _c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "g2s_reduce_ws_bytes": (ctypes.c_size_t, []),
    "g2s_clamped_depth_fwd": (_c_int, [_vp, _c_int, _c_int, _c_int, _c_int, _c_float, _c_float, _c_float, _c_int, _vp, _vp,
                                       _vp, _vp]),
    "g2s_clamped_depth_bwd": (_c_int, [_vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_float, _c_float, _c_int, _vp, _vp,
                                       _vp]),
    "g2s_shading_fwd": (_c_int, [_vp, _c_long, _vp, _vp, _c_long, _c_int, _c_int, _vp, _vp, _vp]),
    "g2s_shading_bwd": (_c_int, [_vp, _c_long, _vp, _vp, _c_long, _c_int, _c_int, _vp, _vp, _vp, _c_long, _vp, _vp,
                                 _c_long, _vp]),
    "g2s_photometric_fwd": (_c_int, [_vp, _vp, _c_long, _vp, _c_float, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _vp, _vp,
                                     _vp]),
    "g2s_photometric_bwd": (_c_int, [_vp, _vp, _c_long, _vp, _c_float, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _vp, _vp,
                                     _vp, _vp, _vp, _vp]),
    "g2s_smooth_fwd": (_c_int, [_vp, _c_int, _c_int, _c_int, _vp, _vp, _vp]),
    "g2s_smooth_bwd": (_c_int, [_vp, _c_int, _c_int, _c_int, _vp, _vp, _vp]),
    "g2s_launch_count": (_c_long, []),
    "g2s_selftest_division": (_c_int, [ctypes.c_ulonglong, ctypes.c_uint, _vp, _vp]),
    "g2s_selftest_face_vertices": (_c_int, [_c_int, _vp, _vp]),
    "g2s_selftest_index_math": (_c_int, [_c_int, _c_int, _c_long, _vp, _vp]),
    "g2s_selftest_raster": (_c_int, [ctypes.c_ulonglong, ctypes.c_uint, _c_int, _vp, _vp]),
    "g2s_profile_enable": (_c_int, [_c_int]),
    "g2s_profile_read": (_c_int, [_c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(_c_float),
                                  ctypes.POINTER(_c_int)]),
    "g2s_grid3d_fwd": (_c_int, [_CAMP, _vp, _c_long, _c_int, _c_int, _c_int, ctypes.POINTER(_c_int), _vp, _vp, _vp,
                                _vp, _vp, _vp, _vp]),
    "g2s_grid3d_bwd": (_c_int, [_CAMP, _vp, _c_long, _c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "g2s_grid_3d_to_2d_fwd": (_c_int, [_CAMP, _vp, _c_int, _c_int, _c_int, _vp, _vp]),
    "g2s_grid_3d_to_2d_bwd": (_c_int, [_CAMP, _vp, _c_int, _c_int, _c_int, _vp, _vp, _vp]),
}

# g2s_workspace_bytes kinds (include/g2s_b200.h)
WS_ZBUFFER, WS_RASTER_BWD, WS_TEX_BWD, WS_TEXELS, WS_GRAD_NORMAL, WS_RGB_MAP, WS_LOSS = range(7)

_lib = None


def load():
    """Loads the shared library (once) and sets the prototypes.  Raises RuntimeError when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libg2s_b200.so is not built (%s). Run `python gan-2d-to-3d_b200/build.py` or "
            "__graft_entry__.build(); there is no CPU / PyTorch fallback for this path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class Context:
    """A caller-owned g2s_context (the streams / events of the multi-lane forward) for the CURRENT device."""

    def __init__(self):
        h = _vp()
        check(load().g2s_context_create(ctypes.byref(h)), "g2s_context_create")
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None) and _lib is not None:
                _lib.g2s_context_destroy(self.handle)
        except Exception:
            pass


def ws_floats(kind, n, S):
    """number of fp32 elements of workspace `kind` for n views / images of side S (g2s_workspace_bytes)"""
    b = load().g2s_workspace_bytes(kind, n, S)
    if b == 0:
        raise RuntimeError("g2s_workspace_bytes: bad arguments (kind %d, n %d, S %d)" % (kind, n, S))
    return (b + 3) // 4


def check(code, what):
    if code != 0:
        msg = load().g2s_error_string(code).decode()
        raise RuntimeError("%s failed: %s (code %d)" % (what, msg, code))
