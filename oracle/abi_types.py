"""ctypes mirrors of the three ABI structs of include/cvo_b200.h, owned by the ORACLE.

TEST INFRASTRUCTURE.  The oracle must not depend on the product: importing it may not import
unified_cvo_b200 (whose library the reference arm of bench.py must never map).  The product has
its own mirrors (unified_cvo_b200/_abi.py); tests/test_abi.py asserts that the two sets describe
the same layout field for field.  Oracle entry points take any ctypes struct of that layout
(argtypes are void pointers), so tests can hand the product's Params straight to the oracle.
"""
import ctypes as C

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
