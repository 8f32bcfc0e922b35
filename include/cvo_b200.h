/*
 * cvo_b200.h — C-ABI of the B200-native Unified-CVO hot path.
 *
 * This is the drop-in boundary: plain C types, no Eigen / PCL / torch in any
 * signature.  The reference (UMich-CURLY/unified_cvo) has no FFI layer; its
 * boundary is the C++ class cvo::CvoGPU (include/UnifiedCvo/cvo/CvoGPU.hpp:33-232).
 * Every entry point below names the reference member/function it replaces, so a
 * maintainer can forward the C++ class to this library (see INTEGRATION.md and
 * include/UnifiedCvo/cvo/CvoGPU.hpp in this repo for the forwarding shim).
 *
 * Conventions
 *   - 4x4 poses are COLUMN-MAJOR float[16] (Eigen::Matrix4f::data() layout).
 *   - 3x3 matrices are column-major float[9].
 *   - features are ROW-major [n x F], label distributions row-major [n x C],
 *     geometric types [n x 2]; any of them may be NULL (treated as zeros, which
 *     is what the reference's zero-initialised device point holds,
 *     utils/PointSegmentedDistribution.hpp:63-76).
 *   - all functions return CVO_B200_OK (0) or a negative error code; nothing
 *     calls exit() (the reference does: cvo/CvoGPU_impl.cuh:27-36).
 *   - a handle is bound to one CUDA device and one stream; it is not
 *     thread-safe (neither is the reference: default stream + global syncs).
 */
#ifndef CVO_B200_H_
#define CVO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVO_B200_ABI_VERSION 1

/* error codes */
#define CVO_B200_OK 0
#define CVO_B200_ERR_INVALID (-2)   /* bad argument                         */
#define CVO_B200_ERR_CUDA (-3)      /* CUDA runtime failure (see last_error) */
#define CVO_B200_ERR_IO (-4)        /* file could not be read               */
#define CVO_B200_ERR_STATE (-5)     /* clouds not set / comm not ready      */
#define CVO_B200_ERR_NCCL (-6)      /* NCCL failure                         */
#define CVO_B200_ERR_NOMEM (-7)

/* Field-for-field mirror of cvo::CvoParams (include/UnifiedCvo/cvo/CvoParams.hpp:12-73):
 * same names, same order, same types, so the C++ shim can static_assert the
 * sizes equal and memcpy between the two. */
typedef struct cvo_b200_params {
  float ell_init_first_frame;
  float ell_init;
  float ell_min;
  int min_ell_iter_limit;
  float ell_max;
  double dl;
  double dl_step;
  float sigma;
  float sp_thres;
  float c;
  float d;
  float c_ell;
  float c_sigma;
  float s_ell;
  float s_sigma;
  int MAX_ITER;
  float eps;
  float eps_2;
  float min_step;
  float max_step;
  float step;

  int nearest_neighbors_max;
  float ell_decay_rate;

  float ell_decay_rate_first_frame;
  int ell_decay_start;
  int ell_decay_start_first_frame;

  int indicator_window_size;
  float indicator_stable_threshold;

  int is_pcl_visualization_on;
  int is_using_least_square;

  int is_ell_adaptive;
  int is_full_ip_matrix;

  int is_using_geometry;
  int is_using_intensity;
  int is_using_semantics;
  int is_using_range_ell;
  int is_using_kdtree;
  int is_exporting_association;
  int is_using_geometric_type;

  int multiframe_using_cpu;
  int multiframe_max_iters;
  float multiframe_ell_init;
  float multiframe_ell_min;
  int multiframe_iter_per_ell;
  float multiframe_ell_decay_rate;
  int multiframe_iterations_per_ell;
  int multiframe_iterations_per_solve;
  int multiframe_expected_points;
  float multiframe_downsample_voxel_size;
  int multiframe_num_neighbors;
  int multiframe_least_squares_num_threads;
  int multiframe_min_nonzeros;
} cvo_b200_params;

/* why an align() loop stopped (bit flags in cvo_b200_iter_trace.flags / align_info.stop_reason) */
#define CVO_B200_STOP_NONE 0
#define CVO_B200_STOP_GRAD_SMALL 1   /* |omega|<eps && |v|<eps   (CvoGPU.cu:1454-1458) */
#define CVO_B200_STOP_GRAD_ZERO 2    /* ... and both < 1e-8 => return -1              */
#define CVO_B200_STOP_DIST_SMALL 4   /* se(3) distance < eps_2   (CvoGPU.cu:1505-1508) */
#define CVO_B200_STOP_MAX_ITER 8
#define CVO_B200_ELL_DECAYED 16      /* ell was decayed at the end of this iteration  */

/* One CVO iteration, as the reference's debug log would describe it
 * (CvoGPU.cu:1387-1533).  Doubles as the parity probe. 232 bytes. */
typedef struct cvo_b200_iter_trace {
  int32_t iter;
  int32_t num_neighbors;     /* row cap used by this iteration                      */
  float ell;                 /* length-scale used by this iteration                 */
  uint32_t max_row_nnz;      /* max_i nonzeros[i]                                   */
  uint64_t nnz;              /* A_host.nonzero_sum                                  */
  double omega_sum[3];       /* sum_i omega_i/c   (double, before cast+normalise)   */
  double v_sum[3];           /* sum_i v_i/d                                         */
  float omega[3];            /* after joint normalisation (CvoGPU.cu:827-832)       */
  float v[3];
  double B, C, D, E;         /* step polynomial sums (CvoGPU.cu:1118-1121)          */
  float step;                /* clamped step (CvoGPU.cu:1151-1158)                  */
  int32_t flags;             /* CVO_B200_STOP_* | CVO_B200_ELL_DECAYED              */
  double dist;               /* || log(dRT) ||  (CvoGPU.cu:1473-1476)               */
  float R[9];                /* pose AFTER the update, column-major                 */
  float T[3];
  float ell_next;            /* ell for the next iteration                          */
  int32_t num_neighbors_next;
  double a_sum;              /* sum of stored kernel values (double accumulate)     */
  int32_t reserved[6];
} cvo_b200_iter_trace;

typedef struct cvo_b200_align_info {
  int32_t ret;               /* what CvoGPU::align returns: 0, or -1 (gradient vanished) */
  int32_t iterations;        /* the reference's "cvo # of iterations is k"          */
  int32_t stop_reason;       /* CVO_B200_STOP_*                                     */
  int32_t final_num_neighbors;
  float final_ell;
  float cell_query_fraction; /* share of the loop run with cell queries instead of the dense scan */
  double registration_seconds; /* CUDA-event time of the loop only (CvoGPU.cu:1534-1560) */
  double upload_seconds;       /* host->device of the clouds, when align is given host clouds */
  uint64_t pairs_tested;       /* N*M*iterations: the unit of the headline metric   */
} cvo_b200_align_info;

typedef struct cvo_b200_handle cvo_b200_handle;

/* ---- library / device --------------------------------------------------- */
int cvo_b200_abi_version(void);
/* number of CUDA devices visible, or a negative error code */
int cvo_b200_device_count(void);
/* sizeof of the ABI structs as compiled into the library, for binding self-checks:
 * which = 0 cvo_b200_params, 1 cvo_b200_iter_trace, 2 cvo_b200_align_info; -1 otherwise */
int cvo_b200_sizeof(int which);
/* thread-local message of the last failing call without a handle */
const char* cvo_b200_global_error(void);

/* ---- parameters (replaces CvoParams ctor + read_CvoParams_yaml,
 *      CvoParams.hpp:75-126 and :193-303) -------------------------------- */
void cvo_b200_params_default(cvo_b200_params* p);
/* Flat "key: number  # comment" YAML subset, '%YAML' / '---' lines ignored.
 * Duplicate keys: the FIRST occurrence wins, as with yaml-cpp. */
int cvo_b200_params_read_yaml(const char* path, cvo_b200_params* p);

/* ---- handle (replaces CvoGPU::CvoGPU / ~CvoGPU / write_params,
 *      CvoGPU.cu:64-83) ---------------------------------------------------- */
int cvo_b200_create(const cvo_b200_params* p, int device, cvo_b200_handle** out);
void cvo_b200_destroy(cvo_b200_handle* h);
int cvo_b200_write_params(cvo_b200_handle* h, const cvo_b200_params* p);
int cvo_b200_get_params(const cvo_b200_handle* h, cvo_b200_params* out);
const char* cvo_b200_last_error(const cvo_b200_handle* h);

/* ---- clouds (replaces CvoPointCloud_to_gpu, CvoGPU_impl.cu:206-285) ------
 * which: 0 = source (fixed cloud x, the rows of the kernel matrix),
 *        1 = target (moving cloud y).
 * Pointers are HOST pointers; the call packs SoA device buffers and returns
 * after the copy has been enqueued and completed.                            */
int cvo_b200_set_cloud(cvo_b200_handle* h, int which, int n, const float* xyz,
                       int F, const float* features, int C, const float* labels,
                       const float* geotype);
/* Source rows [row_begin,row_end) are the ones this handle scans.  Default: all rows
 * (row_end = -1).  For a single handle (tests, manual partitioning).  In a multi-GPU job
 * (after cvo_b200_comm_init) the range is NOT taken from here: every call derives the rank's
 * shard from (rank, world) and the size of the source cloud currently set - contiguous blocks
 * of ceil(N / world) rows rounded up to 64 - so re-uploading a source of another size keeps
 * the shards covering it.                                                     */
int cvo_b200_set_row_range(cvo_b200_handle* h, int row_begin, int row_end);

/* ---- the hot path --------------------------------------------------------
 * One CVO iteration (CvoGPU.cu:1389-1531) at an explicit state.  R,T are the
 * CURRENT pose blocks of T_target_to_source (column-major R); the kernels move
 * the target by its inverse exactly like update_tf (CvoGPU.cu:94-112).
 * The controller state (indicator queues) is NOT touched; trace->ell_next and
 * the decay flag are therefore not meaningful here.                          */
int cvo_b200_iterate(cvo_b200_handle* h, const float R[9], const float T[3], float ell,
                     int num_neighbors, cvo_b200_iter_trace* trace);

/* Full registration (replaces CvoGPU::align, CvoGPU.cu:1605-1632 + align_impl
 * :1338-1572) on the clouds previously set.  T_init = T_target_to_source.
 * T_out = [R^T, -R^T T; 0 0 0 1] (maps target points into the source frame).
 * trace may be NULL; at most trace_cap records are written (the first ones).
 * Returns CVO_B200_OK on success; info->ret carries the reference's return.   */
int cvo_b200_align(cvo_b200_handle* h, const float T_init[16], float T_out[16],
                   cvo_b200_align_info* info, cvo_b200_iter_trace* trace, int trace_cap);

/* Same, but uploads both clouds from host memory first and reads the pose back:
 * the end-to-end call a CvoGPU::align(const CvoPointCloud&, ...) makes.       */
int cvo_b200_align_host(cvo_b200_handle* h, int n_src, const float* src_xyz, int F,
                        const float* src_feat, int C, const float* src_labels,
                        const float* src_geotype, int n_tgt, const float* tgt_xyz,
                        const float* tgt_feat, const float* tgt_labels,
                        const float* tgt_geotype, const float T_init[16], float T_out[16],
                        cvo_b200_align_info* info);

/* sum_ij A_ij with the row cap at nearest_neighbors_max (replaces
 * CvoGPU::inner_product_gpu -> inner_product_impl -> A_sum,
 * CvoGPU.cu:1719-1794, SparseKernelMat.cu:62-66).                            */
int cvo_b200_inner_product(cvo_b200_handle* h, const float T[16], float ell, float* out);

/* cos overlap score (replaces CvoGPU::function_angle, CvoGPU.cu:1814-1846)   */
int cvo_b200_function_angle(cvo_b200_handle* h, const float T[16], float ell,
                            int is_approximate, float* out);

/* Soft data association (replaces CvoGPU::compute_association_gpu,
 * CvoGPU.cu:1876-1911 isotropic and :1975-1995 with a 3x3 kernel, and
 * gpu_association_to_cpu, CvoGPU_impl.cu:366-427).
 * kernel3x3 == NULL -> isotropic kernel with `ell`; else the Mahalanobis kernel
 * (column-major 3x3, inverted inside like CvoGPU.cu:1946).
 * Two-call protocol: call with cols/vals == NULL to get *nnz and row_ptr
 * (n_src+1 ints, may be NULL too), then with buffers of *nnz entries.
 * Output is CSR over (source row, target index) in the reference's insertion
 * order (ascending target index inside a row).                               */
int cvo_b200_association(cvo_b200_handle* h, const float T[16], float ell,
                         const float* kernel3x3, int64_t* nnz, int32_t* row_ptr,
                         int32_t* cols, float* vals);

/* The kernel matrix of the LAST iteration of the most recent cvo_b200_align on this handle
 * (replaces the export at the end of align_impl, CvoGPU.cu:1552-1556 ->
 * gpu_association_to_cpu, CvoGPU_impl.cu:366-427, done when is_exporting_association).
 * Same two-call CSR protocol as cvo_b200_association; row_ptr has n_src+1 entries; rows outside
 * this handle's row range are empty.  CVO_B200_ERR_STATE if another call has overwritten the
 * matrix since.  (The reference reads its matrix back with the row stride of the NEXT
 * iteration's cap when the loop ends on MAX_ITER — CvoGPU.cu:1518-1529 then :1553 — which
 * scrambles it; this call always returns the matrix as computed.)                          */
int cvo_b200_align_association(cvo_b200_handle* h, int64_t* nnz, int32_t* row_ptr,
                               int32_t* cols, float* vals);

/* ---- pose-graph edges (multi-frame registration) ----------
 * Frames stay resident on the device in the caller's layout (replaces the
 * points_init_gpu_ of a CvoFrameGPU, CvoFrameGPU.cu:7-30).  `frame` is a small
 * non-negative id chosen by the caller; setting an id again replaces its
 * contents.  Same array conventions as cvo_b200_set_cloud.                    */
int cvo_b200_frame_set(cvo_b200_handle* h, int frame, int n, const float* xyz, int F,
                       const float* features, int C, const float* labels,
                       const float* geotype);
/* frees one frame's device arrays (frame = -1: all of them)                  */
int cvo_b200_frame_clear(cvo_b200_handle* h, int frame);
/* One edge of the pose graph, one outer IRLS iteration: both frames are moved
 * by their own CURRENT poses — row-major 3x4 float, x' = P [x 1]^T (replaces
 * CvoFrameGPU::transform_pointcloud, CvoFrameGPU.cu:44-62 ->
 * transform_point_pose_vec, CvoGPU_impl.cu:84-150) — and the row-capped kernel
 * matrix between the moved clouds is filled with the edge's fixed `ell` and cap
 * `num_neighbors` and copied out (replaces BinaryStateGPU::update_inner_product,
 * IRLS_State_GPU.cu:43-79: clear_SparseKernelMat, fill_in_A_mat_gpu,
 * compute_nonzeros, copy_internal_SparseKernelMat_gpu_to_cpu).  Nothing but the
 * two poses goes to the device and nothing but the matrix comes back.
 * Rows = points of frame1, columns = points of frame2, CSR in the reference's
 * insertion order; *max_row_nnz (may be NULL) is what max_neighbors()
 * (SparseKernelMat.cu:48-53) returns for the caller's cap update
 * (IRLS_State_GPU.cu:45-47).  num_neighbors may exceed nearest_neighbors_max.
 * Two-call protocol like cvo_b200_association; the second call with the same
 * arguments reuses the matrix still on the device.  Overwrites the handle's two
 * cloud slots (cvo_b200_set_cloud).                                           */
int cvo_b200_edge_update(cvo_b200_handle* h, int frame1, const float pose1[12], int frame2,
                         const float pose2[12], float ell, int num_neighbors, int64_t* nnz,
                         int32_t* max_row_nnz, int32_t* row_ptr, int32_t* cols, float* vals);

/* All edges of one outer IRLS iteration in ONE call (the loop of CvoBatchIRLS::solve,
 * IRLS.cpp:111-121, over BinaryStateGPU::update_inner_product): every edge's kernels are enqueued
 * back to back - frame 1 moved by its pose, cell queries in frame 2's own cell table, CSR
 * compaction at a device-side running offset - and the host waits ONCE for all of them.
 * Same per-edge semantics and outputs as cvo_b200_edge_update.
 * nnz[e], max_row_nnz[e]: per edge.  row_ptr: the edges' row pointers back to back (edge e has
 * n(frame1_e) + 1 entries, each starting at 0).  cols / vals: the edges' entries back to back
 * (edge e's block starts at the sum of the earlier edges' nnz).  Two-call protocol: cols == NULL
 * computes everything and returns the sizes; the second call with the same edges copies the
 * entries without recomputing.  Needs is_using_geometry; every edge must be evaluable by cell
 * queries / tile cells (else CVO_B200_ERR_STATE: use cvo_b200_edge_update for that edge).     */
typedef struct cvo_b200_edge {
  int32_t frame1, frame2;
  float pose1[12], pose2[12]; /* row-major 3x4 */
  float ell;
  int32_t num_neighbors;
} cvo_b200_edge;
int cvo_b200_edge_update_batch(cvo_b200_handle* h, int n_edges, const cvo_b200_edge* edges,
                               int64_t* nnz, int32_t* max_row_nnz, int32_t* row_ptr, int32_t* cols,
                               float* vals);

/* How many iterations of the last cvo_b200_align built candidate cells (persistent tile mode: the
 * cells are reused while the pose has drifted less than their skin; 0 in the other modes). */
int cvo_b200_last_candidate_builds(const cvo_b200_handle* h);

/* ---- measurement helpers --------------------------------------------------
 * Runs `iters` iterations back to back at a FIXED state (pose, ell, cap),
 * timed with CUDA events on the handle's stream.  ms_total = whole iteration
 * chain; ms_pair_kernel = sum of the dense pairwise kernel's launches only
 * (events recorded around each launch).  Either output may be NULL.          */
int cvo_b200_time_iterations(cvo_b200_handle* h, const float R[9], const float T[3],
                             float ell, int num_neighbors, int iters, float* ms_total,
                             float* ms_pair_kernel);
/* number of kernel launches issued by this handle since creation            */
uint64_t cvo_b200_launch_count(const cvo_b200_handle* h);
/* the CUDA stream (cudaStream_t) all work of this handle is enqueued on      */
void* cvo_b200_stream(const cvo_b200_handle* h);
/* fp32 FMA-pipe microbenchmark on the handle's device: issues `iters` rounds
 * of independent FFMA chains, returns achieved lane-FMA/s (for the roofline
 * denominator).  kind: 0 = scalar FFMA, 1 = packed f32x2.     */
int cvo_b200_fma_peak(cvo_b200_handle* h, int kind, int iters, double* fma_per_s);

/* ---- multi-GPU (new capability: the reference is single-GPU) --------------
 * One process per GPU.  Rank 0 obtains a 128-byte NCCL unique id, the host
 * plumbing (torch.distributed / MPI / anything) broadcasts it, every rank calls
 * comm_init.  After that align/iterate shard the SOURCE rows by set_row_range
 * and all-reduce {omega,v,nnz,max} and {B,C,D,E} once each per iteration.     */
int cvo_b200_comm_unique_id(char id[128]);
int cvo_b200_comm_init(cvo_b200_handle* h, int rank, int world, const char id[128]);
/* Fused exchange (optional, after comm_init): every rank publishes the CUDA IPC handle of a small
 * mailbox in its HBM, the host plumbing all-gathers the 64-byte handles, every rank maps its
 * peers'.  From then on the cell-query and tile-cell modes run the WHOLE registration loop of a
 * sharded job in one persistent kernel per GPU whose two per-iteration exchanges are NVLink stores
 * into the peers' mailboxes + a spin on the own one (no NCCL call, no kernel boundary); the
 * dense-scan mode keeps the NCCL all-gathers.  handles = world x 64 bytes, rank order.                        */
int cvo_b200_comm_mailbox_handle(cvo_b200_handle* h, char out[64]);
int cvo_b200_comm_open_peers(cvo_b200_handle* h, const char* handles);
/* on != 0: cvo_b200_inner_product / cvo_b200_function_angle become COLLECTIVE calls of the job
 * (every rank must make them with the same arguments): each rank scans its shard of the source
 * rows, the ranks' sums of A are all-gathered and added in rank order (bit-identical result on
 * every rank).  Default off: every rank computes all rows on its own.  cvo_b200_association always
 * computes all rows; cvo_b200_align_association after a sharded align returns the rows of this
 * rank's shard (the other rows empty) - the host merges the disjoint row sets
 * (unified_cvo_b200/dist.py::gather_association).                                          */
int cvo_b200_comm_shard_inner_products(cvo_b200_handle* h, int on);
int cvo_b200_comm_destroy(cvo_b200_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* CVO_B200_H_ */
