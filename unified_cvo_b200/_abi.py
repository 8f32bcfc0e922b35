"""ctypes mirror of include/cvo_b200.h (types + library loader).

The library is the product: if libcvo_b200.so is missing or fails to load this
module raises — there is no CPU fallback of any kind.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libcvo_b200.so")

OK = 0
ERR_INVALID, ERR_CUDA, ERR_IO, ERR_STATE, ERR_NCCL, ERR_NOMEM = -2, -3, -4, -5, -6, -7

STOP_NONE, STOP_GRAD_SMALL, STOP_GRAD_ZERO, STOP_DIST_SMALL, STOP_MAX_ITER = 0, 1, 2, 4, 8
ELL_DECAYED = 16


class Params(C.Structure):
    """cvo_b200_params == cvo::CvoParams (CvoParams.hpp:12-73), field for field."""

    _fields_ = [
        ("ell_init_first_frame", C.c_float),
        ("ell_init", C.c_float),
        ("ell_min", C.c_float),
        ("min_ell_iter_limit", C.c_int),
        ("ell_max", C.c_float),
        ("dl", C.c_double),
        ("dl_step", C.c_double),
        ("sigma", C.c_float),
        ("sp_thres", C.c_float),
        ("c", C.c_float),
        ("d", C.c_float),
        ("c_ell", C.c_float),
        ("c_sigma", C.c_float),
        ("s_ell", C.c_float),
        ("s_sigma", C.c_float),
        ("MAX_ITER", C.c_int),
        ("eps", C.c_float),
        ("eps_2", C.c_float),
        ("min_step", C.c_float),
        ("max_step", C.c_float),
        ("step", C.c_float),
        ("nearest_neighbors_max", C.c_int),
        ("ell_decay_rate", C.c_float),
        ("ell_decay_rate_first_frame", C.c_float),
        ("ell_decay_start", C.c_int),
        ("ell_decay_start_first_frame", C.c_int),
        ("indicator_window_size", C.c_int),
        ("indicator_stable_threshold", C.c_float),
        ("is_pcl_visualization_on", C.c_int),
        ("is_using_least_square", C.c_int),
        ("is_ell_adaptive", C.c_int),
        ("is_full_ip_matrix", C.c_int),
        ("is_using_geometry", C.c_int),
        ("is_using_intensity", C.c_int),
        ("is_using_semantics", C.c_int),
        ("is_using_range_ell", C.c_int),
        ("is_using_kdtree", C.c_int),
        ("is_exporting_association", C.c_int),
        ("is_using_geometric_type", C.c_int),
        ("multiframe_using_cpu", C.c_int),
        ("multiframe_max_iters", C.c_int),
        ("multiframe_ell_init", C.c_float),
        ("multiframe_ell_min", C.c_float),
        ("multiframe_iter_per_ell", C.c_int),
        ("multiframe_ell_decay_rate", C.c_float),
        ("multiframe_iterations_per_ell", C.c_int),
        ("multiframe_iterations_per_solve", C.c_int),
        ("multiframe_expected_points", C.c_int),
        ("multiframe_downsample_voxel_size", C.c_float),
        ("multiframe_num_neighbors", C.c_int),
        ("multiframe_least_squares_num_threads", C.c_int),
        ("multiframe_min_nonzeros", C.c_int),
    ]

    def copy(self) -> "Params":
        out = Params()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(Params))
        return out

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


class IterTrace(C.Structure):
    _fields_ = [
        ("iter", C.c_int32),
        ("num_neighbors", C.c_int32),
        ("ell", C.c_float),
        ("max_row_nnz", C.c_uint32),
        ("nnz", C.c_uint64),
        ("omega_sum", C.c_double * 3),
        ("v_sum", C.c_double * 3),
        ("omega", C.c_float * 3),
        ("v", C.c_float * 3),
        ("B", C.c_double),
        ("C", C.c_double),
        ("D", C.c_double),
        ("E", C.c_double),
        ("step", C.c_float),
        ("flags", C.c_int32),
        ("dist", C.c_double),
        ("R", C.c_float * 9),
        ("T", C.c_float * 3),
        ("ell_next", C.c_float),
        ("num_neighbors_next", C.c_int32),
        ("a_sum", C.c_double),
        ("reserved", C.c_int32 * 6),
    ]


class AlignInfo(C.Structure):
    _fields_ = [
        ("ret", C.c_int32),
        ("iterations", C.c_int32),
        ("stop_reason", C.c_int32),
        ("final_num_neighbors", C.c_int32),
        ("final_ell", C.c_float),
        ("cell_query_fraction", C.c_float),
        ("registration_seconds", C.c_double),
        ("upload_seconds", C.c_double),
        ("pairs_tested", C.c_uint64),
    ]


class Edge(C.Structure):
    """cvo_b200_edge: one pose-graph edge of a batched update."""

    _fields_ = [("frame1", C.c_int32), ("frame2", C.c_int32), ("pose1", C.c_float * 12),
                ("pose2", C.c_float * 12), ("ell", C.c_float), ("num_neighbors", C.c_int32)]


_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)

# every symbol include/cvo_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "cvo_b200_abi_version": (C.c_int, []),
    "cvo_b200_device_count": (C.c_int, []),
    "cvo_b200_sizeof": (C.c_int, [C.c_int]),
    "cvo_b200_global_error": (C.c_char_p, []),
    "cvo_b200_params_default": (None, [C.POINTER(Params)]),
    "cvo_b200_params_read_yaml": (C.c_int, [C.c_char_p, C.POINTER(Params)]),
    "cvo_b200_create": (C.c_int, [C.POINTER(Params), C.c_int, C.POINTER(C.c_void_p)]),
    "cvo_b200_destroy": (None, [C.c_void_p]),
    "cvo_b200_write_params": (C.c_int, [C.c_void_p, C.POINTER(Params)]),
    "cvo_b200_get_params": (C.c_int, [C.c_void_p, C.POINTER(Params)]),
    "cvo_b200_last_error": (C.c_char_p, [C.c_void_p]),
    "cvo_b200_set_cloud": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int, _f32p, C.c_int, _f32p, C.c_int, _f32p, _f32p],
    ),
    "cvo_b200_set_row_range": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "cvo_b200_iterate": (
        C.c_int,
        [C.c_void_p, _f32p, _f32p, C.c_float, C.c_int, C.POINTER(IterTrace)],
    ),
    "cvo_b200_align": (
        C.c_int,
        [C.c_void_p, _f32p, _f32p, C.POINTER(AlignInfo), C.POINTER(IterTrace), C.c_int],
    ),
    "cvo_b200_align_host": (
        C.c_int,
        [C.c_void_p, C.c_int, _f32p, C.c_int, _f32p, C.c_int, _f32p, _f32p,
         C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.POINTER(AlignInfo)],
    ),
    "cvo_b200_inner_product": (C.c_int, [C.c_void_p, _f32p, C.c_float, _f32p]),
    "cvo_b200_function_angle": (C.c_int, [C.c_void_p, _f32p, C.c_float, C.c_int, _f32p]),
    "cvo_b200_association": (
        C.c_int,
        [C.c_void_p, _f32p, C.c_float, _f32p, C.POINTER(C.c_int64), _i32p, _i32p, _f32p],
    ),
    "cvo_b200_align_association": (
        C.c_int,
        [C.c_void_p, C.POINTER(C.c_int64), _i32p, _i32p, _f32p],
    ),
    "cvo_b200_frame_set": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int, _f32p, C.c_int, _f32p, C.c_int, _f32p, _f32p],
    ),
    "cvo_b200_frame_clear": (C.c_int, [C.c_void_p, C.c_int]),
    "cvo_b200_edge_update": (
        C.c_int,
        [C.c_void_p, C.c_int, _f32p, C.c_int, _f32p, C.c_float, C.c_int, C.POINTER(C.c_int64),
         _i32p, _i32p, _i32p, _f32p],
    ),
    "cvo_b200_edge_update_batch": (
        C.c_int,
        [C.c_void_p, C.c_int, C.POINTER(Edge), C.POINTER(C.c_int64), _i32p, _i32p, _i32p, _f32p],
    ),
    "cvo_b200_time_iterations": (
        C.c_int,
        [C.c_void_p, _f32p, _f32p, C.c_float, C.c_int, C.c_int, _f32p, _f32p],
    ),
    "cvo_b200_launch_count": (C.c_uint64, [C.c_void_p]),
    "cvo_b200_last_candidate_builds": (C.c_int, [C.c_void_p]),
    "cvo_b200_stream": (C.c_void_p, [C.c_void_p]),
    "cvo_b200_fma_peak": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "cvo_b200_comm_unique_id": (C.c_int, [C.c_char * 128]),
    "cvo_b200_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char * 128]),
    "cvo_b200_comm_mailbox_handle": (C.c_int, [C.c_void_p, C.c_char * 64]),
    "cvo_b200_comm_open_peers": (C.c_int, [C.c_void_p, C.c_char_p]),
    "cvo_b200_comm_shard_inner_products": (C.c_int, [C.c_void_p, C.c_int]),
    "cvo_b200_comm_destroy": (C.c_int, [C.c_void_p]),
}

_lib = None


def load_library(path: str | None = None) -> C.CDLL:
    """dlopen libcvo_b200.so and bind every declared symbol.  Raises if absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C unified_cvo_b200/csrc`). There is no CPU fallback."
        )
    lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib
