/*
 * cvo_oracle.h — CPU restatement of the reference's CvoGPU hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under unified_cvo_b200/ may include, link
 * or call this; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker / the CPU arm.
 *
 * PARITY UNPINNED: the reference ships no golden vectors, known-answer tests or
 * fixtures for this path and cannot be compiled in this
 * image (Eigen, Sophus, PCL, yaml-cpp, TBB absent), so this restatement could
 * not be checked against outputs of the reference itself.  It is pinned only
 * by (i) an independent numpy brute-force restatement agreeing with it
 * (tests/test_oracle.py), (ii) analytic invariants, and (iii) the committed
 * fixtures under tests/golden/ that freeze ITS outputs against regressions.
 *
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference).  Arithmetic types follow the C++ promotion rules of the
 * cited lines literally (float products, double exp, double row sums, ...);
 * compile with -ffp-contract=off so no FMA contraction sneaks in.
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
/* LieGroup.cpp:245-274 Exp_SEK3 (K=1); out = 3x4 column-major float */
void oracle_exp_sek3(const float xi[6], float dt, float out12[12]);
/* LieGroup.cpp:309-325 poly_solver_order3 (double overload): roots of
 * c0 t^3 + c1 t^2 + c2 t + c3; returns 0 and fills re/im, or -1 if the
 * companion matrix is not finite (Eigen would return NaNs). */
int oracle_cubic_roots(const double coef[4], double re[3], double im[3]);
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

#ifdef __cplusplus
}
#endif
#endif
