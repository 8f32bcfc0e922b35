/*
 * cvo_cpu_baseline.h — restated reference CPU path (class cvo::cvo, src/cvo/Cvo.cpp): the CPU
 * arm timed beside the GPU path.  TEST / MEASUREMENT INFRASTRUCTURE ONLY; see
 * cvo_cpu_baseline.c for what follows which reference lines.
 */
#ifndef CVO_CPU_BASELINE_H_
#define CVO_CPU_BASELINE_H_

#include "cvo_oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cpu_baseline_info {
  int ret;              /* cvo::align's return: 0, or -1 (gradient vanished)            */
  int iterations;       /* the reference's "cvo # of iterations is k"                    */
  int executed;         /* iterations actually run (k + 1 on a break, MAX_ITER otherwise) */
  int threads;          /* OpenMP threads used                                          */
  float final_ell;
  long long nnz_last;   /* A.nonZeros() of the last iteration                           */
  unsigned long long pairs; /* N * M * executed: the unit of the headline metric        */
  double seconds;       /* whole registration (the reference times transform/flow/step, Cvo.cpp:1031-1033) */
  double t_transform, t_kdtree_build, t_se_kernel, t_flow, t_step, t_rest;
} cpu_baseline_info;

/* cvo::set_pcd + cvo::align (Cvo.cpp:1182-1233, 885-1089).  T_init: column-major 4x4 initial
 * T_target_to_source or NULL (identity).  T_out = [R^T, -R^T T; 0 0 0 1] like CvoGPU::align. */
int cpu_baseline_align(const cvo_b200_params* p, const oracle_cloud* src, const oracle_cloud* tgt,
                       const float T_init[16], int use_semantics, float T_out[16],
                       cpu_baseline_info* info);

/* test tap of the kd-tree: per query point, how many of pts lie within squared distance r2 */
void cpu_baseline_radius_counts(const float* pts, int m, const float* q, int n, float r2, int* counts);

#ifdef __cplusplus
}
#endif
#endif
