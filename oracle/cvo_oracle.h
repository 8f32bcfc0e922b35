/*
 * cvo_oracle.h — CPU restatement of the reference's CvoGPU hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under unified_cvo_b200/ may include, link
 * or call this; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker / the CPU arm.
 *
 * PARITY: what is pinned against the reference itself, and what is not.
 *   PINNED (tier 1): fill_in_A_mat_gpu (CvoGPU.cu:477-593) with compute_range_ell,
 *   compute_geometric_type_ip and the gpu_utils.cuh helpers - the dominant kernel, which uses no
 *   Eigen/PCL/thrust - is compiled FROM THE REFERENCE'S OWN TEXT by oracle/make_ref.py into
 *   oracle/_ref/ (g++ host build, nvcc sm_100a builds); oracle_fill_A equals it bit for bit
 *   (values, indices, counts): tests/test_ref_pin.py (CPU) and tests/test_ref_pin_gpu.py (B200).
 *   PINNED MODULO EIGEN (tier 2): the dense-kernel variant K1b, compute_flow_gpu_no_eigen (K2),
 *   compute_step_size_xi / _poly_coeff (K3, K4) are compiled from the reference's text too, but
 *   against oracle/ref_mini_eigen.h, our stand-in for Eigen's fixed-size 3-vector primitives:
 *   formulas, float/double mix and statement order are the reference's, the evaluation order
 *   inside a 3-term product/reduction is our reading of Eigen 3.3 (c0 + (c1 + c2)).
 *   UNPINNED: the host controller (align_impl's schedule, Exp_SEK3, the cubic via Eigen's
 *   eigenvalue solver, Sophus' SE3 log) - needs Eigen/Sophus proper; checked only against an
 *   independent numpy restatement, closed forms (numpy.roots, scipy expm/logm) and invariants.
 *
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference).  Arithmetic types follow the C++ promotion rules of the
 * cited lines literally (float products, double exp, double row sums, ...);
 * compile with -ffp-contract=off: the only fused operations are the explicit fmaf() calls of the
 * device-arithmetic mode.
 */
#ifndef CVO_ORACLE_H_
#define CVO_ORACLE_H_

#include "../include/cvo_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_cloud {
  int n;                 /* points                                          */
  int F;                 /* feature dimension actually provided (0 ok)      */
  int C;                 /* classes actually provided (0 ok)                */
  const float* xyz;      /* n x 3                                           */
  const float* feat;     /* n x F row-major, may be NULL                    */
  const float* labels;   /* n x C row-major, may be NULL                    */
  const float* geotype;  /* n x 2, may be NULL (zeros)                      */
} oracle_cloud;

/* the ELL-style row-truncated kernel matrix (cvo/SparseKernelMat.hpp:11-19);
 * row stride is the num_neighbors of the fill call, like the reference. */
typedef struct oracle_sparse {
  int rows;
  int stride;            /* = num_neighbors of the last fill               */
  float* mat;            /* rows*capacity                                  */
  int* ind;              /* rows*capacity, -1 terminated                   */
  unsigned int* nonzeros;/* rows                                           */
  int capacity;          /* allocated columns                              */
  unsigned long long nonzero_sum;
} oracle_sparse;

oracle_sparse* oracle_sparse_new(int rows, int capacity);
void oracle_sparse_free(oracle_sparse* A);

/* CvoGPU.cu:94-112 update_tf */
void oracle_update_tf(const float R[9], const float T[3], float Rinv[9], float Tinv[3],
                      float transform16[16]);
/* CvoGPU_impl.cu:31-82 transform_point_R_T (xyz only matter) */
void oracle_transform(const float Rinv[9], const float Tinv[3], const float* y, int m,
                      float* y_out);
/* CvoGPU.cu:477-593 fill_in_A_mat_gpu (+ clear, SparseKernelMat.cu:90-98;
 * + compute_nonzeros :37-46).  y_moved = transformed target xyz. */
void oracle_fill_A(const cvo_b200_params* p, const oracle_cloud* src, const oracle_cloud* tgt,
                   const float* y_moved, int num_neighbors, float ell, oracle_sparse* A);
/* CvoGPU.cu:217-327 fill_in_A_mat_gpu_dense_mat_kernel; kernel_inv column-major */
void oracle_fill_A_dense_kernel(const cvo_b200_params* p, const oracle_cloud* src,
                                const oracle_cloud* tgt, const float* y_moved,
                                int num_neighbors, const float kernel_inv[9],
                                oracle_sparse* A);
/* CvoGPU.cu:729-848: per-row flow, double reduction, joint normalisation */
void oracle_compute_flow(const cvo_b200_params* p, const oracle_cloud* src,
                         const float* y_moved, const oracle_sparse* A, double omega_sum[3],
                         double v_sum[3], float omega[3], float v[3]);
/* CvoGPU.cu:953-1164: xi powers, B..E, cubic, clamp.  Returns step. */
float oracle_compute_step(const cvo_b200_params* p, const oracle_cloud* src,
                          const float* y_moved, int m, const oracle_sparse* A,
                          const float omega[3], const float v[3], float ell, double BCDE[4]);
/* the per-row values of the two passes above, before their reductions: omega_i/c, v_i/d
 * (rows x 3 doubles each) and B_i, C_i, D_i, E_i (rows x 4 doubles).  Test taps. */
void oracle_flow_rows(const cvo_b200_params* p, const oracle_cloud* src, const float* y_moved,
                      const oracle_sparse* A, double* omega_rows, double* v_rows);
void oracle_step_rows(const cvo_b200_params* p, const oracle_cloud* src, const float* y_moved, int m,
                      const oracle_sparse* A, const float omega[3], const float v[3], float ell,
                      double* bcde_rows);
/* LieGroup.cpp:245-274 Exp_SEK3 (K=1); out = 3x4 column-major float */
void oracle_exp_sek3(const float xi[6], float dt, float out12[12]);
/* LieGroup.cpp:309-325 poly_solver_order3 (double overload): roots of
 * c0 t^3 + c1 t^2 + c2 t + c3; returns 0 and fills re/im, or -1 if the
 * companion matrix is not finite (Eigen would return NaNs). */
int oracle_cubic_roots(const double coef[4], double re[3], double im[3]);
/* CvoGPU.cu:1167-1285 A_sparsity_indicator_ell_update over a sequence of indicators, the queues
 * empty at entry like in align_impl (:1377-1380); per call: decision, start sum, end sum.  Test tap. */
void oracle_indicator_sequence(const cvo_b200_params* params, int n, const float* indicators,
                               int* decrease, float* start_sums, float* end_sums);
/* Sophus::SE3d(dRT).log().norm(), call site CvoGPU.cu:1473-1476 */
double oracle_se3_log_norm(const double dR[9], const double dT[3]);

/* CvoGPU.cu:1338-1572 align_impl (+ :1605-1632).  Returns the reference's
 * return value (0 / -1).  trace may be NULL. */
int oracle_align(const cvo_b200_params* p, const oracle_cloud* src, const oracle_cloud* tgt,
                 const float T_init[16], float T_out[16], cvo_b200_align_info* info,
                 cvo_b200_iter_trace* trace, int trace_cap);
/* one iteration at an explicit state (what cvo_b200_iterate computes) */
void oracle_iterate(const cvo_b200_params* p, const oracle_cloud* src, const oracle_cloud* tgt,
                    const float R[9], const float T[3], float ell, int num_neighbors,
                    cvo_b200_iter_trace* trace, oracle_sparse* A_out /* may be NULL */);
/* CvoGPU.cu:1719-1778 inner_product_impl + SparseKernelMat.cu:62-66 A_sum.
 * A_out (optional) receives the association matrix. */
float oracle_inner_product(const cvo_b200_params* p, const oracle_cloud* src,
                           const oracle_cloud* tgt, const float T[16], float ell,
                           const float* kernel3x3, oracle_sparse* A_out);
/* CvoGPU.cu:1814-1846 function_angle (gpu branch) */
float oracle_function_angle(const cvo_b200_params* p, const oracle_cloud* src,
                            const oracle_cloud* tgt, const float T[16], float ell,
                            int is_approximate);
/* CvoGPU_impl.cu:84-150 transform_point_pose_vec; pose12 = row-major 3x4 */
void oracle_transform_pose_vec(const float pose12[12], const float* x, int n, float* x_out);
/* CvoFrameGPU.cu:44-62 + IRLS_State_GPU.cu:43-79: one edge update of the multi-frame IRLS */
unsigned long long oracle_edge_update(const cvo_b200_params* p, const oracle_cloud* f1,
                                      const float pose1[12], const oracle_cloud* f2,
                                      const float pose2[12], float ell, int num_neighbors,
                                      oracle_sparse* A);
int oracle_num_threads(void);
void oracle_set_num_threads(int n);
/* 1 (default): rows visit only the targets of the 27 grid cells around them, in ascending order -
 * outputs bit-identical to the dense loop; 0 (or ORACLE_DENSE=1): the literal dense N x M loop. */
void oracle_set_accel(int on);
/* Arithmetic of K1's float sums: 1 (default) = with the FMA contractions nvcc applies to the
 * reference's text under the reference's own flags (what its GPU path computes), 0 = every
 * multiply/add rounded separately (the text compiled by g++ / nvcc --fmad=false).  See the
 * comment above mul_add() in cvo_oracle.c. */
void oracle_set_device_arith(int on);
int oracle_device_arith(void);

#ifdef __cplusplus
}
#endif
#endif
