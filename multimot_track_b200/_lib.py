"""ctypes binding of include/orbx.h.  Fails loudly when liborbx.so is missing:
the product path has no CPU fallback."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_i32p = ctypes.POINTER(ctypes.c_int32)
c_f32p = ctypes.POINTER(ctypes.c_float)

ORBX_OK = 0
ERR_NAMES = {0: "ORBX_OK", -1: "ORBX_ERR_BAD_ARG", -2: "ORBX_ERR_CAPACITY", -3: "ORBX_ERR_CUDA",
             -4: "ORBX_ERR_OOM", -5: "ORBX_ERR_UNSUPPORTED", -6: "ORBX_ERR_STATE"}


class OrbxError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("%s (%d): %s" % (ERR_NAMES.get(code, "ORBX_ERR"), code, message))
        self.code = code


class OrbxConfig(ctypes.Structure):
    _fields_ = [("nfeatures", ctypes.c_int32), ("scale_factor", ctypes.c_float), ("nlevels", ctypes.c_int32),
                ("ini_th_fast", ctypes.c_int32), ("min_th_fast", ctypes.c_int32), ("max_width", ctypes.c_int32),
                ("max_height", ctypes.c_int32), ("max_batch", ctypes.c_int32), ("device_id", ctypes.c_int32)]


class OrbxPlan(ctypes.Structure):
    _fields_ = [("nlevels", ctypes.c_int32), ("max_keypoints", ctypes.c_int32)] + \
               [(n, ctypes.c_int32 * 16) for n in ("level_width", "level_height", "nfeatures_per_level", "cell_cols", "cell_rows",
                                                   "cell_w", "cell_h", "octree_roots", "max_candidates")] + \
               [(n, ctypes.c_float * 16) for n in ("scale", "inv_scale", "sigma2", "inv_sigma2", "keypoint_size")] + \
               [("umax", ctypes.c_int32 * 16)]


class OrbxPoolConfig(ctypes.Structure):
    _fields_ = [("extractor", OrbxConfig), ("ndevices", ctypes.c_int32), ("devices", ctypes.POINTER(ctypes.c_int32)), ("depth", ctypes.c_int32),
                ("compact_keypoints", ctypes.c_int32)]


class OrbxShardResult(ctypes.Structure):
    _fields_ = [("kps", ctypes.c_void_p), ("desc", ctypes.c_void_p), ("n", ctypes.c_void_p), ("nframes", ctypes.c_int32),
                ("first_frame", ctypes.c_int32), ("cap_per_frame", ctypes.c_int32), ("device", ctypes.c_int32), ("ckps", ctypes.c_void_p)]


# every symbol include/orbx.h declares: (restype, argtypes)
_VP, _I, _F, _SZ = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t
SIGNATURES = {
    "orbx_create": (_I, [ctypes.POINTER(OrbxConfig), ctypes.POINTER(_VP)]),
    "orbx_destroy": (None, [_VP]),
    "orbx_last_error": (ctypes.c_char_p, [_VP]),
    "orbx_get_tables": (_I, [_VP, _VP, _VP, _VP, _VP, _VP]),
    "orbx_max_keypoints": (_I, [_VP, _I, _I]),
    "orbx_make_plan": (_I, [ctypes.POINTER(OrbxConfig), _I, _I, ctypes.POINTER(OrbxPlan)]),
    "orbx_extract": (_I, [_VP, _VP, _I, _I, _I, _VP, _VP, _I, _VP]),
    "orbx_extract_color": (_I, [_VP, _VP, _I, _I, _I, _I, _I, _VP, _VP, _I, _VP]),
    "orbx_extract_batch": (_I, [_VP, _VP, _I, _I, _I, _I, _VP, _VP, _I, _VP]),
    "orbx_submit_device": (_I, [_VP, _VP, _I, _I, _I, _I, _SZ]),
    "orbx_submit_host": (_I, [_VP, _VP, _I, _I, _I, _I]),
    "orbx_collect": (_I, [_VP, _VP, _VP, _I, _VP]),
    "orbx_collect_view_compact": (_I, [_VP, ctypes.POINTER(_VP), ctypes.POINTER(_VP), ctypes.POINTER(_VP), ctypes.POINTER(_I)]),
    "orbx_expand_keypoints": (_I, [_VP, _VP, _I, _VP]),
    "orbx_pool_set_option": (_I, [_VP, _I, _I]),
    "orbx_collect_view": (_I, [_VP, ctypes.POINTER(_VP), ctypes.POINTER(_VP), ctypes.POINTER(_VP), ctypes.POINTER(_I)]),
    "orbx_get_level_size": (_I, [_VP, _I, ctypes.POINTER(_I), ctypes.POINTER(_I)]),
    "orbx_get_pyramid_level": (_I, [_VP, _I, _I, _VP, _I, _I]),
    "orbx_get_pyramid_levels": (_I, [_VP, _I, _I, _VP, _VP, _I]),
    "orbx_get_blurred_level": (_I, [_VP, _I, _I, _VP, _I]),
    "orbx_get_candidates": (_I, [_VP, _I, _I, _VP, _I, ctypes.POINTER(_I)]),
    "orbx_hamming256": (_I, [_VP, _VP]),
    "orbx_match": (_I, [_VP, _VP, _I, _VP, _I, _I, _F, _VP, _VP, _VP, _VP]),
    "orbx_match_device": (_I, [_VP, _VP, _I, _VP, _I, _I, _F, _VP, _VP, _VP, _VP]),
    "orbx_rotation_filter": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _I, _VP, _VP]),
    "orbx_rotation_filter_device": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "orbx_search_for_initialization": (_I, [_VP, _F, _F, _F, _F, _I, _VP, _VP, _VP, _I, _VP, _VP, _VP, _VP, _VP, _I, _F, _I, _VP]),
    "orbx_search_local_points": (_I, [_VP, _F, _F, _F, _F, _I, _VP, _VP, _VP, _VP, _VP, _VP, _I, _VP, _VP, _VP, _VP, _VP, _F, _F, _VP]),
    "orbx_search_by_bow": (_I, [_VP, _I, _VP, _VP, _VP, _I, _VP, _VP, _VP, _I, _VP, _VP, _I, _VP, _VP, _VP, _F, _I, _VP]),
    "orbx_search_by_bow_keyframes": (_I, [_VP, _I, _VP, _VP, _VP, _I, _VP, _VP, _VP, _I, _VP, _VP, _VP, _I, _VP, _VP, _VP, _F, _I, _VP]),
    "orbx_search_by_projection": (_I, [_VP, _VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _I, _VP, _VP, _VP, _VP, _VP, _F, _I, _I, _VP]),
    "orbx_voc_load_text": (_I, [_VP, ctypes.c_char_p, ctypes.POINTER(_VP)]),
    "orbx_voc_create": (_I, [_VP, _I, _I, _I, _I, _I, _VP, _VP, _VP, _VP, ctypes.POINTER(_VP)]),
    "orbx_voc_destroy": (None, [_VP]),
    "orbx_voc_info": (_I, [_VP, _VP, _VP, _VP, _VP]),
    "orbx_voc_transform": (_I, [_VP, _VP, _VP, _I, _I, _VP, _VP, _VP]),
    "orbx_voc_bow": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "orbx_distinctive_descriptors": (_I, [_VP, _VP, _VP, _I, _VP, _VP]),
    "orbx_stereo_match": (_I, [_VP, _VP, _I, _I, _F, _VP, _VP, _VP, _I, _VP]),
    "orbx_set_option": (_I, [_VP, _I, _I]),
    "orbx_pool_create": (_I, [ctypes.POINTER(OrbxPoolConfig), ctypes.POINTER(_VP)]),
    "orbx_pool_destroy": (None, [_VP]),
    "orbx_pool_last_error": (ctypes.c_char_p, [_VP]),
    "orbx_pool_devices": (_I, [_VP]),
    "orbx_pool_depth": (_I, [_VP]),
    "orbx_pool_handle": (_VP, [_VP, _I, _I]),
    "orbx_pool_shard_range": (None, [_I, _I, _I, ctypes.POINTER(_I), ctypes.POINTER(_I)]),
    "orbx_pool_submit_host": (ctypes.c_longlong, [_VP, _VP, _I, _I, _I, _I]),
    "orbx_pool_submit_device": (ctypes.c_longlong, [_VP, _VP, _VP, _I, _I, _I, _SZ]),
    "orbx_pool_collect": (_I, [_VP, ctypes.c_longlong, _VP]),
    "orbx_set_profiling": (_I, [_VP, _I]),
    "orbx_get_stage_ms": (_I, [_VP, _VP, _I]),
    "orbx_stage_name": (ctypes.c_char_p, [_I]),
    "orbx_sync": (_I, [_VP]),
    "orbx_stream": (_VP, [_VP]),
    "orbx_launch_count": (ctypes.c_longlong, [_VP]),
    "orbx_version": (ctypes.c_char_p, []),
}


def lib_path():
    return os.environ.get("ORBX_LIBRARY", os.path.join(_HERE, "liborbx.so"))


def load_library():
    """Load liborbx.so (built in-tree by `make -C multimot_track_b200/csrc` or __graft_entry__.build())."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise OrbxError(-3, "CUDA extension %s is missing; build it with `make -C multimot_track_b200/csrc` "
                            "(there is no CPU fallback)" % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == ABI mismatch with include/orbx.h
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(lib, handle, rc):
    if rc < 0:
        msg = lib.orbx_last_error(handle)
        raise OrbxError(rc, msg.decode() if msg else "")
    return rc


def make_plan(nfeatures, scale_factor, nlevels, ini_th, min_th, width, height):
    """Host-only geometry/tables for a shape (orbx_make_plan); needs no GPU."""
    lib = load_library()
    cfg = OrbxConfig(int(nfeatures), float(scale_factor), int(nlevels), int(ini_th), int(min_th), 0, 0, 0, -1)
    plan = OrbxPlan()
    rc = lib.orbx_make_plan(ctypes.byref(cfg), int(width), int(height), ctypes.byref(plan))
    if rc != 0:
        raise OrbxError(rc, (lib.orbx_last_error(None) or b"").decode())
    return plan
