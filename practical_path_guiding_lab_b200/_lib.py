"""ctypes binding of libsdtree.so (include/sdtree.h).

The CUDA library is the product: there is no fallback.  If libsdtree.so has not been
built (python -m practical_path_guiding_lab_b200.build) loading fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "libsdtree.so")

SDT_OK = 0
SDT_HOST_PTRS = 1
SDT_SYNC = 2
SDT_NO_WAIT = 16
SDT_REFINE_NO_KD = 4
SDT_REFINE_NO_QUAD = 8
SDT_TREE_PREV = 0
SDT_TREE_CURRENT = 1

c_f32p = C.POINTER(C.c_float)
c_u32p = C.POINTER(C.c_uint32)
c_u8p = C.POINTER(C.c_uint8)


class Vec3(C.Structure):
    _fields_ = [("x", C.c_void_p), ("y", C.c_void_p), ("z", C.c_void_p), ("stride", C.c_int64)]


class Vec2(C.Structure):
    _fields_ = [("x", C.c_void_p), ("y", C.c_void_p), ("stride", C.c_int64)]


class Config(C.Structure):
    _fields_ = [("bbox_min", C.c_float * 3), ("bbox_max", C.c_float * 3),
                ("kd_max_depth", C.c_int32), ("quad_max_depth", C.c_int32),
                ("store_nee", C.c_int32), ("device", C.c_int32),
                ("kd_capacity", C.c_uint32), ("quad_capacity", C.c_uint32)]


class Sizes(C.Structure):
    _fields_ = [(k, C.c_uint32) for k in ("n_kd", "n_quad", "n_roots", "n_interior", "n_levels",
                                         "kd_leaves", "error", "refine_count", "jump_trees", "jump2_tables")]


class Arrays(C.Structure):
    _fields_ = [("n_kd", C.c_uint32), ("n_quad", C.c_uint32), ("n_roots", C.c_uint32),
                ("kd_max_leaf_size", C.c_float), ("kd_max_depth", C.c_int32),
                ("quad_max_depth", C.c_int32), ("quad_store_nee", C.c_int32),
                ("kd_bbox_min", C.c_void_p), ("kd_bbox_max", C.c_void_p), ("kd_depth", C.c_void_p),
                ("kd_vert_count", C.c_void_p), ("kd_is_leaf", C.c_void_p), ("kd_quad_root", C.c_void_p),
                ("kd_child_left", C.c_void_p), ("kd_child_right", C.c_void_p),
                ("q_root_node", C.c_void_p), ("q_bbox_min", C.c_void_p), ("q_bbox_max", C.c_void_p),
                ("q_depth", C.c_void_p), ("q_irradiance", C.c_void_p), ("q_is_leaf", C.c_void_p),
                ("q_threshold", C.c_void_p), ("q_child", C.c_void_p * 4)]


class GuidedArgs(C.Structure):
    _fields_ = [("pos", Vec3), ("wo", Vec3), ("mode", C.c_void_p),
                ("u", C.c_void_p), ("u_stride", C.c_uint32), ("seed", C.c_uint32), ("lane_offset", C.c_uint32),
                ("bsdf_pdf", C.c_void_p), ("bsdf_value", Vec3), ("bsdf_sampling_fraction", C.c_double),
                ("dir", Vec3), ("sdtree_pdf", C.c_void_p), ("wo_pdf", C.c_void_p), ("weight", Vec3),
                ("em_dir", Vec3), ("em_active", C.c_void_p), ("sdtree_pdf_em", C.c_void_p)]


class Records(C.Structure):
    _fields_ = [("position", Vec3), ("direction", Vec2), ("radiance", C.c_void_p), ("wo_pdf", C.c_void_p),
                ("radiance_nee", Vec3), ("direction_nee", Vec2), ("active", C.c_void_p)]


class PathData(C.Structure):
    _fields_ = [("slots", C.c_uint32), ("max_depth", C.c_uint32), ("l_final", Vec3),
                ("throughput_radiance", Vec3), ("throughput_bsdf", Vec3), ("bsdf", Vec3),
                ("position", Vec3), ("direction", Vec2), ("wo_pdf", C.c_void_p),
                ("radiance_nee", Vec3), ("direction_nee", Vec2), ("active", C.c_void_p),
                ("radiance_out", C.c_void_p)]


# every symbol include/sdtree.h declares: name -> (restype, argtypes)
_H = C.c_void_p
_S = C.c_void_p
SYMBOLS = {
    "sdt_create": (C.c_int, [C.POINTER(Config), C.POINTER(_H)]),
    "sdt_destroy": (C.c_int, [_H]),
    "sdt_last_error": (C.c_char_p, [_H]),
    "sdt_upload": (C.c_int, [_H, C.POINTER(Arrays)]),
    "sdt_upload_stats": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "sdt_get_sizes": (C.c_int, [_H, C.POINTER(Sizes)]),
    "sdt_download": (C.c_int, [_H, C.c_int, C.POINTER(Arrays)]),
    "sdt_locate": (C.c_int, [_H, C.POINTER(Vec3), C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, _S]),
    "sdt_sample": (C.c_int, [_H, C.POINTER(Vec3), C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32,
                             C.c_uint32, C.POINTER(Vec3), C.c_void_p, C.c_void_p, C.c_uint32, _S]),
    "sdt_pdf": (C.c_int, [_H, C.POINTER(Vec3), C.POINTER(Vec3), C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                          C.c_uint32, _S]),
    "sdt_guided": (C.c_int, [_H, C.POINTER(GuidedArgs), C.c_uint32, C.c_uint32, _S]),
    "sdt_mis_nee": (C.c_int, [_H, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_double, C.c_int32, C.c_void_p, C.c_void_p, C.c_uint32, _S]),
    "sdt_mis_mixture": (C.c_int, [_H, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(Vec3), C.c_void_p, C.c_double,
                                  C.c_void_p, C.POINTER(Vec3), C.c_uint32, _S]),
    "sdt_dir_to_canonical": (C.c_int, [_H, C.POINTER(Vec3), C.c_uint32, C.c_void_p, C.c_uint32, _S]),
    "sdt_canonical_to_dir": (C.c_int, [_H, C.POINTER(Vec2), C.c_uint32, C.POINTER(Vec3), C.c_uint32, _S]),
    "sdt_splat_records": (C.c_int, [_H, C.POINTER(Records), C.c_uint32, C.c_uint32, _S]),
    "sdt_splat_path_data": (C.c_int, [_H, C.POINTER(PathData), C.c_uint32, _S]),
    "sdt_set_iteration_threshold": (C.c_int, [_H, C.c_int32]),
    "sdt_set_max_leaf_size": (C.c_int, [_H, C.c_float]),
    "sdt_refine": (C.c_int, [_H, C.c_uint32, _S]),
    "sdt_reset_stats": (C.c_int, [_H, _S]),
    "sdt_comm_unique_id": (C.c_int, [C.c_void_p]),
    "sdt_comm_init": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int32]),
    "sdt_allreduce": (C.c_int, [_H, _S]),
    "sdt_hint_records": (C.c_int, [_H, C.c_uint64]),
    "sdt_stat_buffers": (C.c_int, [_H, C.POINTER(C.c_void_p), C.POINTER(C.c_uint32), C.POINTER(C.c_void_p),
                                   C.POINTER(C.c_uint32)]),
    "sdt_set_tuning": (C.c_int, [_H, C.c_char_p, C.c_int64]),
    "sdt_synchronize": (C.c_int, [_H, C.c_void_p]),
    "sdt_kernel_launches": (C.c_uint64, [_H]),
    "sdt_sample_pdf": (C.c_int, [_H, C.POINTER(Vec3), C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                 C.POINTER(Vec3), C.c_void_p, C.POINTER(Vec3), C.c_void_p, C.c_uint32, _S]),
    "sdt_measure_l2": (C.c_int, [_H, C.c_uint64, C.c_uint32, C.POINTER(C.c_float), _S]),
    "sdt_measure_gather": (C.c_int, [_H, C.c_uint64, C.c_uint32, C.c_int32, C.POINTER(C.c_float), _S]),
}

_cache = {}


def load_library(path=None):
    """dlopen libsdtree.so and type every entry point.  No fallback: raises if missing."""
    path = os.path.abspath(path or DEFAULT_LIB)
    if path in _cache:
        return _cache[path]
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} not found: the CUDA library is not built. "
            "Run `python -m practical_path_guiding_lab_b200.build` (needs nvcc). There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _cache[path] = lib
    return lib
