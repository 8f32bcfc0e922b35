/*
 * cvo_oracle.c — CPU restatement of the reference's CvoGPU hot path.
 * TEST INFRASTRUCTURE ONLY; PARITY UNPINNED — see cvo_oracle.h.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC
 * (see oracle/Makefile).  Written from the behaviour of the cited reference
 * lines; expression shapes are kept so that C's usual arithmetic conversions
 * reproduce the float/double mix of the C++/CUDA original.
 *
 * Conventions the reference leaves to Eigen internals (unknowable here, Eigen
 * is not in the image) and which this oracle FIXES as normative:
 *   - 3-term inner products / squaredNorm / matrix-product coefficients are
 *     summed as c0 + (c1 + c2)   (Eigen 3.3 redux_novec_unroller, Length 3);
 *     the 6-term norm of the stacked twist as (c0+(c1+c2)) + (c3+(c4+c5)).
 *   - no FMA contraction anywhere.
 *   - thrust::reduce order = ascending row order, double.
 *
 * The dense N x M row loop of fill_in_A_mat_gpu is the normative statement; by default the rows
 * enumerate their candidates through a uniform grid (same targets that can pass, same ascending
 * order => bit-identical outputs, tests/test_oracle.py), which is what makes this file usable as
 * a CPU baseline.  ORACLE_DENSE=1 or oracle_set_accel(0) runs the literal loop.
 */
#include "cvo_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU baseline asks for the host's cores
 * explicitly (bench.py) instead of inheriting that. */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---------- small fixed-size helpers (float) ---------------------------- */
static inline float sum3f(float c0, float c1, float c2) { return c0 + (c1 + c2); }
static inline float dot3f(const float* a, const float* b) {
  return sum3f(a[0] * b[0], a[1] * b[1], a[2] * b[2]);
}
static inline double sum3d(double c0, double c1, double c2) { return c0 + (c1 + c2); }
/* column-major 3x3 times vector */
static inline void mat3f_vec(const float* M, const float* x, float* out) {
  for (int i = 0; i < 3; i++) out[i] = sum3f(M[i] * x[0], M[3 + i] * x[1], M[6 + i] * x[2]);
}
static inline void mat3f_mul(const float* A, const float* B, float* out) {
  float tmp[9];
  for (int j = 0; j < 3; j++)
    for (int i = 0; i < 3; i++)
      tmp[3 * j + i] = sum3f(A[i] * B[3 * j], A[3 + i] * B[3 * j + 1], A[6 + i] * B[3 * j + 2]);
  memcpy(out, tmp, sizeof(tmp));
}
/* gpu_utils.cuh:8-15 skew_gpu / LieGroup.cpp:11-19 skew, column-major */
static inline void skewf(const float* v, float* M) {
  M[0] = 0;     M[3] = -v[2]; M[6] = v[1];
  M[1] = v[2];  M[4] = 0;     M[7] = -v[0];
  M[2] = -v[1]; M[5] = v[0];  M[8] = 0;
}
static inline void cross3f(const float* a, const float* b, float* out) {
  out[0] = a[1] * b[2] - a[2] * b[1];
  out[1] = a[2] * b[0] - a[0] * b[2];
  out[2] = a[0] * b[1] - a[1] * b[0];
}

/* ---------- sparse matrix ----------------------------------------------- */
oracle_sparse* oracle_sparse_new(int rows, int capacity) {
  oracle_sparse* A = (oracle_sparse*)calloc(1, sizeof(oracle_sparse));
  if (!A) return NULL;
  A->rows = rows;
  A->capacity = capacity;
  A->stride = capacity;
  size_t n = (size_t)(rows > 0 ? rows : 1) * (size_t)(capacity > 0 ? capacity : 1);
  A->mat = (float*)calloc(n, sizeof(float));
  A->ind = (int*)malloc(n * sizeof(int));
  A->nonzeros = (unsigned int*)calloc((size_t)(rows > 0 ? rows : 1), sizeof(unsigned int));
  if (!A->mat || !A->ind || !A->nonzeros) {
    oracle_sparse_free(A);
    return NULL;
  }
  memset(A->ind, 0xFF, n * sizeof(int));
  return A;
}
void oracle_sparse_free(oracle_sparse* A) {
  if (!A) return;
  free(A->mat);
  free(A->ind);
  free(A->nonzeros);
  free(A);
}
/* SparseKernelMat.cu:90-98 clear_SparseKernelMat(A, num_neighbors) */
static void sparse_clear(oracle_sparse* A, int num_neighbors) {
  size_t n = (size_t)A->rows * (size_t)num_neighbors;
  A->nonzero_sum = 0;
  A->stride = num_neighbors;
  memset(A->mat, 0, n * sizeof(float));
  memset(A->ind, 0xFF, n * sizeof(int));
  memset(A->nonzeros, 0, (size_t)A->rows * sizeof(unsigned int));
}

/* ---------- pose plumbing ------------------------------------------------ */
/* CvoGPU.cu:94-112: R_inv = R^T; T_inv = -R_inv*T; transform = [R_inv T_inv;0 0 0 1] */
void oracle_update_tf(const float R[9], const float T[3], float Rinv[9], float Tinv[3],
                      float transform16[16]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Rinv[3 * j + i] = R[3 * i + j];
  float neg[9];
  for (int k = 0; k < 9; k++) neg[k] = -Rinv[k];
  mat3f_vec(neg, T, Tinv);
  if (transform16) {
    for (int j = 0; j < 3; j++) {
      for (int i = 0; i < 3; i++) transform16[4 * j + i] = Rinv[3 * j + i];
      transform16[4 * j + 3] = 0.f;
    }
    transform16[12] = Tinv[0];
    transform16[13] = Tinv[1];
    transform16[14] = Tinv[2];
    transform16[15] = 1.f;
  }
}

/* CvoGPU_impl.cu:46-53: trans = (*R) * input + (*T) */
void oracle_transform(const float Rinv[9], const float Tinv[3], const float* y, int m,
                      float* y_out) {
  for (int j = 0; j < m; j++) {
    float r[3];
    mat3f_vec(Rinv, y + 3 * j, r);
    y_out[3 * j + 0] = r[0] + Tinv[0];
    y_out[3 * j + 1] = r[1] + Tinv[1];
    y_out[3 * j + 2] = r[2] + Tinv[2];
  }
}

/* ---------- kernel matrix ------------------------------------------------ */
/* Arithmetic of the float sums of fill_in_A_mat_gpu (the Eigen-free kernel K1).
 *   0  "as written": every multiply and every add is rounded on its own - what g++ makes of the
 *      text on x86-64 and what nvcc makes of it with --fmad=false.  Pinned bit for bit against
 *      oracle/_ref/libcvo_ref_host.so (the reference's own text compiled by g++).
 *   1  (default) "as the reference's GPU computes": the reference builds Release with nvcc's
 *      default --fmad=true (CMakeLists.txt:29,79), so nvcc contracts a*b + c into fma(a,b,c).
 *      oracle/_ref/ref_harness.ptx (the reference's text compiled with its own flags) shows which:
 *        result += t*t                 -> result = fma(t, t, result)     (gpu_utils.cuh:24-41,97-104)
 *        dx*dx + dy*dy + dz*dz         -> fma(dz, dz, fma(dx, dx, dy*dy))  (gpu_utils.cuh:73-78, CvoGPU.cu:506)
 *      everything else of K1 is products, divisions and library calls (no contraction possible).
 *      Pinned bit for bit on a B200 against oracle/_ref/libcvo_ref_cuda.so (tests, -m gpu).
 * The Eigen-typed kernels (transform, K2-K4) are not affected: how Eigen's fixed-size
 * expressions contract is unknowable without Eigen itself, they keep c0 + (c1 + c2). */
static int g_device_arith = 1;
void oracle_set_device_arith(int on) { g_device_arith = on ? 1 : 0; }
int oracle_device_arith(void) { return g_device_arith; }
static inline float mul_add(int fused, float a, float b, float c) {
  return fused ? fmaf(a, b, c) : a * b + c;
}
/* gpu_utils.cuh:73-78 squared_dist(a, b) of two points, and the norm under CvoGPU.cu:506 */
static inline float sum_sq3(int fused, float dx, float dy, float dz) {
  if (fused) return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
  return dx * dx + dy * dy + dz * dz;
}
/* gpu_utils.cuh:24-41: serial accumulation starting from 0 */
static inline float dot_serial(int fused, const float* a, const float* b, int dim) {
  float result = 0;
  for (int i = 0; i < dim; i++) result = mul_add(fused, a[i], b[i], result);
  return result;
}
/* CvoGPU.cu:203-215 compute_geometric_type_ip */
static inline float geometric_type_ip(int fused, const float* ga, const float* gb) {
  float norm2_a = dot_serial(fused, ga, ga, 2);
  float norm2_b = dot_serial(fused, gb, gb, 2);
  float dot_ab = dot_serial(fused, ga, gb, 2);
  return dot_ab * dot_ab / (norm2_a * norm2_b);
}
/* CvoGPU.cu:86-90 compute_range_ell */
static inline float range_ell(float curr_ell, float dist_to_sensor) {
  float final_ell = ((dist_to_sensor) / 500.0 + 1.0) * curr_ell;
  return final_ell;
}

static const float kZero2[2] = {0.f, 0.f};

/* ---------- accelerated candidate enumeration (same outputs as the dense loop) ---------------
 * The reference's row loop visits every target j in ascending order and `continue`s on
 * d2 >= d2_thres (CvoGPU.cu:549-554).  With the isotropic geometric kernel on, a target farther
 * than h = max_i sqrt(d2_thres_i) from the row can therefore never change anything, so the row
 * may visit only the targets of the 27 cells of a uniform grid of edge h around it - in
 * ASCENDING j, so that the first-num_neighbors truncation (:526) and the order of every float
 * sum are those of the dense loop.  Outputs are bit-identical (tests/test_oracle.py checks it);
 * the dense loop stays the normative statement and is what ORACLE_DENSE=1 / oracle_set_accel(0)
 * run.  This is what makes the oracle a fair CPU baseline: the reference's own CPU path
 * (cvo::cvo::align, Cvo.cpp:349-456) also searches neighbours (nanoflann kd-tree) instead of
 * scanning N x M. */
static int g_accel = -1;
void oracle_set_accel(int on) { g_accel = on ? 1 : 0; }
static int accel_enabled(void) {
  if (g_accel < 0) {
    const char* e = getenv("ORACLE_DENSE");
    g_accel = (e && e[0] == '1') ? 0 : 1;
  }
  return g_accel;
}
typedef struct {
  int nx, ny, nz;
  float lo[3], inv_h;
  int* cell_start; /* ncell + 1 */
  int* items;      /* target indices, ascending inside a cell */
} cand_grid;
static int cmp_int(const void* a, const void* b) {
  const int x = *(const int*)a, y = *(const int*)b;
  return (x > y) - (x < y);
}
static int grid_cell_coord(float v, float lo, float inv_h, int n) {
  int c = (int)floorf((v - lo) * inv_h);
  if (c < 0) c = 0;
  if (c >= n) c = n - 1;
  return c;
}
/* returns 0 when the grid is not applicable (caller falls back to the dense loop) */
static int grid_build(cand_grid* g, const float* y, int m, float h) {
  memset(g, 0, sizeof(*g));
  if (!(h > 0.f) || !isfinite(h) || m <= 0) return 0;
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  int nfin = 0;
  for (int j = 0; j < m; j++) {
    const float* q = y + 3 * (size_t)j;
    if (!(isfinite(q[0]) && isfinite(q[1]) && isfinite(q[2]))) continue;
    for (int k = 0; k < 3; k++) {
      if (q[k] < lo[k]) lo[k] = q[k];
      if (q[k] > hi[k]) hi[k] = q[k];
    }
    nfin++;
  }
  if (nfin == 0) return 0;
  double dims[3];
  for (int k = 0; k < 3; k++) dims[k] = floor(((double)hi[k] - (double)lo[k]) / (double)h) + 1.0;
  if (dims[0] * dims[1] * dims[2] > 4.0e6) return 0;
  g->nx = (int)dims[0]; g->ny = (int)dims[1]; g->nz = (int)dims[2];
  for (int k = 0; k < 3; k++) g->lo[k] = lo[k];
  g->inv_h = 1.0f / h;
  const int ncell = g->nx * g->ny * g->nz;
  g->cell_start = (int*)calloc((size_t)ncell + 1, sizeof(int));
  g->items = (int*)malloc(sizeof(int) * (size_t)(nfin > 0 ? nfin : 1));
  int* cell_of = (int*)malloc(sizeof(int) * (size_t)m);
  for (int j = 0; j < m; j++) {
    const float* q = y + 3 * (size_t)j;
    if (!(isfinite(q[0]) && isfinite(q[1]) && isfinite(q[2]))) {
      cell_of[j] = -1; /* a non-finite target never passes d2 < thres */
      continue;
    }
    const int cx = grid_cell_coord(q[0], lo[0], g->inv_h, g->nx);
    const int cy = grid_cell_coord(q[1], lo[1], g->inv_h, g->ny);
    const int cz = grid_cell_coord(q[2], lo[2], g->inv_h, g->nz);
    cell_of[j] = (cz * g->ny + cy) * g->nx + cx;
    g->cell_start[cell_of[j] + 1]++;
  }
  for (int c = 0; c < ncell; c++) g->cell_start[c + 1] += g->cell_start[c];
  int* fill = (int*)malloc(sizeof(int) * (size_t)ncell);
  memcpy(fill, g->cell_start, sizeof(int) * (size_t)ncell);
  for (int j = 0; j < m; j++) /* ascending j inside every cell */
    if (cell_of[j] >= 0) g->items[fill[cell_of[j]]++] = j;
  free(fill);
  free(cell_of);
  return 1;
}
static void grid_free(cand_grid* g) {
  free(g->cell_start);
  free(g->items);
}
/* targets of the 27 cells around point p, ascending; returns their number */
static int grid_candidates(const cand_grid* g, const float* p, int* out) {
  if (!(isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]))) return 0; /* d2 is NaN/inf: no pass */
  /* unclamped cell coordinates: a row far outside the grid has no candidates */
  const float fx = floorf((p[0] - g->lo[0]) * g->inv_h), fy = floorf((p[1] - g->lo[1]) * g->inv_h),
              fz = floorf((p[2] - g->lo[2]) * g->inv_h);
  if (fx < -1.f || fy < -1.f || fz < -1.f || fx > (float)g->nx || fy > (float)g->ny || fz > (float)g->nz)
    return 0;
  const int cx = (int)fx, cy = (int)fy, cz = (int)fz;
  int n = 0;
  for (int dz = -1; dz <= 1; dz++) {
    const int z = cz + dz;
    if (z < 0 || z >= g->nz) continue;
    for (int dy = -1; dy <= 1; dy++) {
      const int yy = cy + dy;
      if (yy < 0 || yy >= g->ny) continue;
      for (int dx = -1; dx <= 1; dx++) {
        const int x = cx + dx;
        if (x < 0 || x >= g->nx) continue;
        const int c = (z * g->ny + yy) * g->nx + x;
        const int b = g->cell_start[c], e = g->cell_start[c + 1];
        memcpy(out + n, g->items + b, sizeof(int) * (size_t)(e - b));
        n += e - b;
      }
    }
  }
  qsort(out, (size_t)n, sizeof(int), cmp_int);
  return n;
}

/* shared body of the two fill kernels; kernel_inv == NULL -> isotropic */
static void fill_A_impl(const cvo_b200_params* p, const oracle_cloud* src,
                        const oracle_cloud* tgt, const float* y_moved, int num_neighbors,
                        float ell, const float* kernel_inv, oracle_sparse* A) {
  const int a_size = src->n, b_size = tgt->n;
  const int F = src->F < tgt->F ? src->F : tgt->F; /* missing dims are zeros on both sides */
  const int Fa = src->F, Fb = tgt->F;
  const int Ca = src->C, Cb = tgt->C;
  sparse_clear(A, num_neighbors);
  (void)F;
  const int fused = g_device_arith;

  /* accelerated candidate enumeration (see above): edge = the largest cut-off radius of any row */
  cand_grid grid;
  int use_grid = 0;
  if (accel_enabled() && !kernel_inv && p->is_using_geometry && a_size > 0) {
    float dmax = 0.f;
    int finite_rows = 1;
    for (int i = 0; i < a_size; i++) {
      const float* pa = src->xyz + 3 * (size_t)i;
      const float d = sqrtf(sum_sq3(fused, pa[0], pa[1], pa[2]));
      if (isfinite(d)) { if (d > dmax) dmax = d; } else finite_rows = 0;
    }
    (void)finite_rows; /* non-finite rows produce no candidates and no survivors either way */
    const float lmax = range_ell(ell, dmax);
    const float sigma2 = p->sigma * p->sigma;
    const float thres_max = -2.0 * lmax * lmax * logf(p->sp_thres / sigma2);
    if (thres_max > 0.f && isfinite(thres_max))
      use_grid = grid_build(&grid, y_moved, b_size, (float)(sqrt((double)thres_max) * (1.0 + 1e-5)) + 1e-6f);
  }

#pragma omp parallel
  {
  int* cand = use_grid ? (int*)malloc(sizeof(int) * (size_t)(b_size > 0 ? b_size : 1)) : NULL;
#pragma omp for schedule(dynamic, 16)
  for (int i = 0; i < a_size; i++) {
    /* CvoGPU.cu:497-501 */
    float sigma2 = p->sigma * p->sigma;
    float c2 = p->c_ell * p->c_ell;
    float c_sigma2 = p->c_sigma * p->c_sigma;
    float s_ell = p->s_ell;
    float s_sigma2 = p->s_sigma * p->s_sigma;
    const float* pa = src->xyz + 3 * (size_t)i;
    /* :506-507 */
    float a_to_sensor = sqrtf(sum_sq3(fused, pa[0], pa[1], pa[2]));
    float l = range_ell(ell, a_to_sensor);
    /* :509-515 (device log(float) resolves to logf, crt/math_functions.hpp) */
    float d2_thres = 1, d2_c_thres = 1, d2_s_thres = 1;
    if (!kernel_inv) {
      if (p->is_using_geometry) d2_thres = -2.0 * l * l * logf(p->sp_thres / sigma2);
    }
    if (p->is_using_intensity) d2_c_thres = -2.0 * c2 * logf(p->sp_thres / c_sigma2);
    if (kernel_inv) {
      /* :236-254 (dense-kernel variant squares s_ell first) */
      float s_ell_square = p->s_ell * p->s_ell;
      if (p->is_using_semantics) d2_s_thres = -2.0 * s_ell_square * logf(p->sp_thres / s_sigma2);
    } else {
      if (p->is_using_semantics) d2_s_thres = -2.0 * s_ell * s_ell * logf(p->sp_thres / s_sigma2);
    }
    const float* ga = src->geotype ? src->geotype + 2 * (size_t)i : kZero2;

    unsigned int num_inds = 0;
    float* Ai = A->mat + (size_t)i * num_neighbors;
    int* Ii = A->ind + (size_t)i * num_neighbors;
    const int n_visit = use_grid ? grid_candidates(&grid, pa, cand) : b_size;
    for (int jj = 0; jj < n_visit; jj++) {
      const int j = use_grid ? cand[jj] : jj; /* ascending j either way */
      if (num_inds == (unsigned int)num_neighbors) break; /* :526 */
      const float* pb = y_moved + 3 * (size_t)j;
      float a = 1, sk = 1, ck = 1, k = 1, geo_sim = 1;
      if (p->is_using_geometric_type) { /* :535-547 */
        const float* gb = tgt->geotype ? tgt->geotype + 2 * (size_t)j : kZero2;
        geo_sim = geometric_type_ip(fused, ga, gb);
        if (geo_sim < 0.01) continue;
      }
      if (p->is_using_geometry) {
        if (!kernel_inv) { /* :549-554, squared_dist(*p_b,*p_a) gpu_utils.cuh:73-78 */
          float dx = pb[0] - pa[0], dy = pb[1] - pa[1], dz = pb[2] - pa[2];
          float d2 = sum_sq3(fused, dx, dy, dz);
          if (d2 < d2_thres)
            k = sigma2 * exp(-d2 / (2.0 * l * l));
          else
            continue;
        } else { /* :290-296 + mahananobis_distance :152-171: dist^T * Kinv * dist, dist=a-b */
          float dist[3] = {pa[0] - pb[0], pa[1] - pb[1], pa[2] - pb[2]};
          /* (dist.transpose() * kernel_inv) is a 1x3 row, then * dist */
          float row[3];
          for (int c = 0; c < 3; c++)
            row[c] = sum3f(dist[0] * kernel_inv[3 * c], dist[1] * kernel_inv[3 * c + 1],
                           dist[2] * kernel_inv[3 * c + 2]);
          float d2 = dot3f(row, dist);
          k = sigma2 * exp(-d2 / 2.0);
        }
      }
      if (p->is_using_intensity) { /* :556-562: fixed-width device arrays, absent dims are 0 */
        float d2_color = 0;
        int Fm = Fa > Fb ? Fa : Fb;
        for (int f = 0; f < Fm; f++) {
          float fa = (src->feat && f < Fa) ? src->feat[(size_t)i * Fa + f] : 0.f;
          float fb = (tgt->feat && f < Fb) ? tgt->feat[(size_t)j * Fb + f] : 0.f;
          float tmp = (fa - fb);
          d2_color = mul_add(fused, tmp, tmp, d2_color);
        }
        if (d2_color < d2_c_thres)
          ck = c_sigma2 * exp(-d2_color / (2.0 * c2));
        else
          continue;
      }
      if (p->is_using_semantics) { /* :563-569 */
        float d2_semantic = 0;
        int Cm = Ca > Cb ? Ca : Cb;
        for (int c = 0; c < Cm; c++) {
          float la = (src->labels && c < Ca) ? src->labels[(size_t)i * Ca + c] : 0.f;
          float lb = (tgt->labels && c < Cb) ? tgt->labels[(size_t)j * Cb + c] : 0.f;
          float tmp = (la - lb);
          d2_semantic = mul_add(fused, tmp, tmp, d2_semantic);
        }
        if (d2_semantic < d2_s_thres) {
          if (kernel_inv) {
            float s_ell_square = p->s_ell * p->s_ell;
            sk = s_sigma2 * exp(-d2_semantic / (2.0 * s_ell_square));
          } else {
            sk = p->s_sigma * p->s_sigma * exp(-d2_semantic / (2.0 * s_ell * s_ell));
          }
        } else
          continue;
      }
      a = ck * k * sk * geo_sim; /* :570 */
      if (a > p->sp_thres) {     /* :576-589 */
        Ai[num_inds] = a;
        Ii[num_inds] = j;
        num_inds++;
      }
    }
    A->nonzeros[i] = num_inds; /* :592 */
  }
  free(cand);
  }
  if (use_grid) grid_free(&grid);
  /* SparseKernelMat.cu:37-46 compute_nonzeros */
  unsigned long long s = 0;
  for (int i = 0; i < a_size; i++) s += A->nonzeros[i];
  A->nonzero_sum = s;
}

void oracle_fill_A(const cvo_b200_params* p, const oracle_cloud* src, const oracle_cloud* tgt,
                   const float* y_moved, int num_neighbors, float ell, oracle_sparse* A) {
  fill_A_impl(p, src, tgt, y_moved, num_neighbors, ell, NULL, A);
}
void oracle_fill_A_dense_kernel(const cvo_b200_params* p, const oracle_cloud* src,
                                const oracle_cloud* tgt, const float* y_moved,
                                int num_neighbors, const float kernel_inv[9],
                                oracle_sparse* A) {
  fill_A_impl(p, src, tgt, y_moved, num_neighbors, 0.f, kernel_inv, A);
}

/* ---------- flow --------------------------------------------------------- */
static double* g_rows_out[2] = {NULL, NULL}; /* oracle_flow_rows / oracle_step_rows taps */
void oracle_compute_flow(const cvo_b200_params* p, const oracle_cloud* src,
                         const float* y_moved, const oracle_sparse* A, double omega_sum[3],
                         double v_sum[3], float omega[3], float v[3]) {
  const int rows = A->rows, nn = A->stride;
  double* om_all = (double*)malloc(sizeof(double) * 3 * (size_t)(rows > 0 ? rows : 1));
  double* v_all = (double*)malloc(sizeof(double) * 3 * (size_t)(rows > 0 ? rows : 1));
  /* CvoGPU.cu:729-790 compute_flow_gpu_no_eigen */
#pragma omp parallel for schedule(static)
  for (int i = 0; i < rows; i++) {
    const float* Ai = A->mat + (size_t)i * nn;
    const float* px = src->xyz + 3 * (size_t)i;
    float omega_i[3] = {0, 0, 0}, v_i[3] = {0, 0, 0};
    for (int j = 0; j < nn; j++) {
      int idx = A->ind[(size_t)i * nn + j];
      if (idx == -1) break;
      const float* py = y_moved + 3 * (size_t)idx;
      float cross_xy[3], diff_yx[3];
      cross3f(px, py, cross_xy);
      diff_yx[0] = py[0] - px[0];
      diff_yx[1] = py[1] - px[1];
      diff_yx[2] = py[2] - px[2];
      for (int k = 0; k < 3; k++) {
        omega_i[k] = omega_i[k] + cross_xy[k] * Ai[j];
        v_i[k] = v_i[k] + diff_yx[k] * Ai[j];
      }
    }
    for (int k = 0; k < 3; k++) {
      om_all[3 * (size_t)i + k] = (double)(omega_i[k] / p->c);
      v_all[3 * (size_t)i + k] = (double)(v_i[k] / p->d);
    }
  }
  /* :824-825 thrust::reduce (double) */
  for (int k = 0; k < 3; k++) omega_sum[k] = v_sum[k] = 0.0;
  for (int i = 0; i < rows; i++)
    for (int k = 0; k < 3; k++) {
      omega_sum[k] += om_all[3 * (size_t)i + k];
      v_sum[k] += v_all[3 * (size_t)i + k];
    }
  if (g_rows_out[0]) memcpy(g_rows_out[0], om_all, sizeof(double) * 3 * (size_t)rows);
  if (g_rows_out[1]) memcpy(g_rows_out[1], v_all, sizeof(double) * 3 * (size_t)rows);
  free(om_all);
  free(v_all);
  /* :824-832 cast to float, joint normalisation (Eigen normalize: z>0 ? /= sqrt(z)) */
  float ov[6];
  for (int k = 0; k < 3; k++) {
    ov[k] = (float)omega_sum[k];
    ov[3 + k] = (float)v_sum[k];
  }
  float z = sum3f(ov[0] * ov[0], ov[1] * ov[1], ov[2] * ov[2]) +
            sum3f(ov[3] * ov[3], ov[4] * ov[4], ov[5] * ov[5]);
  if (z > 0.f) {
    float nrm = sqrtf(z);
    for (int k = 0; k < 6; k++) ov[k] = ov[k] / nrm;
  }
  for (int k = 0; k < 3; k++) {
    omega[k] = ov[k];
    v[k] = ov[3 + k];
  }
}

/* ---------- cubic -------------------------------------------------------- */
static double poly3_eval(double p2, double p1, double p0, double t) {
  return ((t + p2) * t + p1) * t + p0;
}
static double poly3_polish(double p2, double p1, double p0, double t) {
  /* Newton from a closed-form start: quadratic convergence; stop on stagnation of |f| so a
   * root that sits between two doubles cannot ping-pong until the iteration cap */
  double fa = fabs(poly3_eval(p2, p1, p0, t));
  for (int it = 0; it < 24; it++) {
    if (fa == 0.0) break;
    double f = poly3_eval(p2, p1, p0, t);
    double df = (3.0 * t + 2.0 * p2) * t + p1;
    if (df == 0.0 || !isfinite(df)) break;
    double tn = t - f / df;
    if (!isfinite(tn) || tn == t) break;
    double fn = fabs(poly3_eval(p2, p1, p0, tn));
    if (!(fn < fa)) break;
    t = tn;
    fa = fn;
  }
  return t;
}
/* Roots of the companion matrix built in LieGroup.cpp:309-325 (double overload),
 * i.e. of t^3 + (c1/c0) t^2 + (c2/c0) t + c3/c0.  Closed form + Newton polish;
 * agrees with an eigenvalue solver to rounding (tests/test_oracle.py checks it
 * against numpy.roots). */
int oracle_cubic_roots(const double coef[4], double re[3], double im[3]) {
  double p2 = coef[1] / coef[0], p1 = coef[2] / coef[0], p0 = coef[3] / coef[0];
  if (!isfinite(p2) || !isfinite(p1) || !isfinite(p0)) {
    for (int i = 0; i < 3; i++) re[i] = im[i] = NAN;
    return -1;
  }
  /* scale t = s*u so the monic coefficients are O(1): s = max(|p2|, sqrt|p1|, cbrt|p0|) */
  double s = fabs(p2);
  if (sqrt(fabs(p1)) > s) s = sqrt(fabs(p1));
  if (cbrt(fabs(p0)) > s) s = cbrt(fabs(p0));
  if (s == 0.0) {
    for (int i = 0; i < 3; i++) re[i] = im[i] = 0.0;
    return 0;
  }
  double a2 = p2 / s, a1 = p1 / (s * s), a0 = p0 / (s * s * s);
  double q = (3.0 * a1 - a2 * a2) / 9.0;
  double r = (9.0 * a2 * a1 - 27.0 * a0 - 2.0 * a2 * a2 * a2) / 54.0;
  double disc = q * q * q + r * r;
  double x1;
  if (disc >= 0.0) {
    double sd = sqrt(disc);
    x1 = cbrt(r + sd) + cbrt(r - sd) - a2 / 3.0;
  } else {
    double th = acos(r / sqrt(-q * q * q));
    x1 = 2.0 * sqrt(-q) * cos(th / 3.0) - a2 / 3.0;
  }
  x1 = poly3_polish(a2, a1, a0, x1);
  /* deflate: u^2 + b u + c */
  /* deflate to u^2 + b u + c with the numerically safer of two formulas each:
   * c = product of the other two roots = -a0/x1 (no cancellation), and
   * b = -(their sum) = a2 + x1  or  (c - a1)/x1, whichever cancels less. */
  double b, c;
  if (x1 != 0.0) {
    c = -a0 / x1;
    double b1 = a2 + x1, b2 = (c - a1) / x1;
    double r1 = fabs(b1) / (fabs(a2) + fabs(x1));
    double den2 = fabs(c) + fabs(a1);
    double r2 = den2 > 0.0 ? fabs(c - a1) / den2 : 0.0;
    b = (r1 >= r2) ? b1 : b2;
  } else {
    b = a2;
    c = a1;
  }
  double d2 = b * b - 4.0 * c;
  double r2re, r2im, r3re, r3im;
  if (d2 >= 0.0) {
    double sq = sqrt(d2);
    double qq = -0.5 * (b + (b >= 0 ? sq : -sq));
    double u2 = qq, u3 = (qq != 0.0) ? c / qq : 0.0;
    if (qq == 0.0) u2 = 0.0;
    u2 = poly3_polish(a2, a1, a0, u2);
    u3 = poly3_polish(a2, a1, a0, u3);
    r2re = u2; r2im = 0.0; r3re = u3; r3im = 0.0;
  } else {
    r2re = -0.5 * b; r2im = 0.5 * sqrt(-d2);
    r3re = r2re;     r3im = -r2im;
  }
  re[0] = x1 * s;   im[0] = 0.0;
  re[1] = r2re * s; im[1] = r2im * s;
  re[2] = r3re * s; im[2] = r3im * s;
  return 0;
}

/* ---------- step size ---------------------------------------------------- */
float oracle_compute_step(const cvo_b200_params* p, const oracle_cloud* src,
                          const float* y_moved, int m, const oracle_sparse* A,
                          const float omega[3], const float v[3], float ell, double BCDE[4]) {
  const int rows = A->rows, nn = A->stride;
  const size_t mm = (size_t)(m > 0 ? m : 1);
  float* xiz = (float*)malloc(sizeof(float) * 3 * mm);
  float* xi2z = (float*)malloc(sizeof(float) * 3 * mm);
  float* xi3z = (float*)malloc(sizeof(float) * 3 * mm);
  float* xi4z = (float*)malloc(sizeof(float) * 3 * mm);
  float* normxiz2 = (float*)malloc(sizeof(float) * mm);
  float* xiz_dot_xi2z = (float*)malloc(sizeof(float) * mm);
  float* epsil_const = (float*)malloc(sizeof(float) * mm);
  /* CvoGPU.cu:953-998 compute_step_size_xi */
  float W[9], W2[9], W3[9], W4[9], Wv[3], W2v[3], W3v[3];
  skewf(omega, W);
  mat3f_mul(W, W, W2);   /* omega_hat*omega_hat */
  mat3f_mul(W2, W, W3);  /* (omega_hat*omega_hat)*omega_hat */
  mat3f_mul(W3, W, W4);
  mat3f_vec(W, v, Wv);
  mat3f_vec(W2, v, W2v);
  mat3f_vec(W3, v, W3v);
#pragma omp parallel for schedule(static)
  for (int j = 0; j < m; j++) {
    const float* y = y_moved + 3 * (size_t)j;
    float t[3];
    cross3f(omega, y, t);
    for (int k = 0; k < 3; k++) xiz[3 * (size_t)j + k] = t[k] + v[k];
    mat3f_vec(W2, y, t);
    for (int k = 0; k < 3; k++) xi2z[3 * (size_t)j + k] = t[k] + Wv[k];
    mat3f_vec(W3, y, t);
    for (int k = 0; k < 3; k++) xi3z[3 * (size_t)j + k] = t[k] + W2v[k];
    mat3f_vec(W4, y, t);
    for (int k = 0; k < 3; k++) xi4z[3 * (size_t)j + k] = t[k] + W3v[k];
    const float* a1 = xiz + 3 * (size_t)j;
    const float* a2 = xi2z + 3 * (size_t)j;
    const float* a3 = xi3z + 3 * (size_t)j;
    normxiz2[j] = dot3f(a1, a1);
    xiz_dot_xi2z[j] = (-dot3f(a1, a2));
    epsil_const[j] = dot3f(a2, a2) + 2 * dot3f(a1, a3);
  }
  /* CvoGPU.cu:1001-1082 compute_step_size_poly_coeff */
  double* Bv = (double*)calloc((size_t)(rows > 0 ? rows : 1) * 4, sizeof(double));
#pragma omp parallel for schedule(static)
  for (int i = 0; i < rows; i++) {
    double Bi = 0.0, Ci = 0.0, Di = 0.0, Ei = 0.0;
    const float* px = src->xyz + 3 * (size_t)i;
    float d2_sqrt = sqrtf(dot3f(px, px));
    float temp_ell = ell;
    if (p->is_using_range_ell) temp_ell = range_ell(ell, d2_sqrt);
    for (int j = 0; j < nn; j++) {
      int idx = A->ind[(size_t)i * nn + j];
      if (idx == -1) break;
      float temp_coef = 1 / (2.0 * temp_ell * temp_ell);
      const float* py = y_moved + 3 * (size_t)idx;
      float diff_xy[3] = {px[0] - py[0], px[1] - py[1], px[2] - py[2]};
      const float* z1 = xiz + 3 * (size_t)idx;
      const float* z2 = xi2z + 3 * (size_t)idx;
      const float* z3 = xi3z + 3 * (size_t)idx;
      const float* z4 = xi4z + 3 * (size_t)idx;
      float two_z2[3] = {2.0f * z2[0], 2.0f * z2[1], 2.0f * z2[2]};
      float neg_z3[3] = {-z3[0], -z3[1], -z3[2]};
      float two_z4[3] = {2.0f * z4[0], 2.0f * z4[1], 2.0f * z4[2]};
      float beta_ij = (-2.0 * temp_coef * dot3f(z1, diff_xy));
      float gamma_ij = (-temp_coef * (normxiz2[idx] + dot3f(two_z2, diff_xy)));
      float delta_ij = (2.0 * temp_coef * (xiz_dot_xi2z[idx] + dot3f(neg_z3, diff_xy)));
      float epsil_ij = (-temp_coef * (epsil_const[idx] + dot3f(two_z4, diff_xy)));
      float A_ij = A->mat[(size_t)i * nn + j];
      double bi = (double)(A_ij * beta_ij);
      Bi += bi;
      double ci = (double)(A_ij * (gamma_ij + beta_ij * beta_ij / 2.0));
      Ci += ci;
      double di =
          (double)(A_ij * (delta_ij + beta_ij * gamma_ij + beta_ij * beta_ij * beta_ij / 6.0));
      Di += di;
      double ei = (double)(A_ij * (epsil_ij + beta_ij * delta_ij +
                                   1 / 2.0 * beta_ij * beta_ij * gamma_ij +
                                   1 / 2.0 * gamma_ij * gamma_ij +
                                   1 / 24.0 * beta_ij * beta_ij * beta_ij * beta_ij));
      Ei += ei;
    }
    Bv[4 * (size_t)i + 0] = Bi;
    Bv[4 * (size_t)i + 1] = Ci;
    Bv[4 * (size_t)i + 2] = Di;
    Bv[4 * (size_t)i + 3] = Ei;
  }
  /* :1118-1121 */
  double B = 0, C = 0, D = 0, E = 0;
  for (int i = 0; i < rows; i++) {
    B += Bv[4 * (size_t)i + 0];
    C += Bv[4 * (size_t)i + 1];
    D += Bv[4 * (size_t)i + 2];
    E += Bv[4 * (size_t)i + 3];
  }
  if (g_rows_out[0]) memcpy(g_rows_out[0], Bv, sizeof(double) * 4 * (size_t)rows);
  free(Bv);
  free(xiz); free(xi2z); free(xi3z); free(xi4z);
  free(normxiz2); free(xiz_dot_xi2z); free(epsil_const);
  BCDE[0] = B; BCDE[1] = C; BCDE[2] = D; BCDE[3] = E;
  /* :1124-1158 */
  double coef[4] = {4.0 * E, 3.0 * D, 2.0 * C, B};
  double re[3], im[3];
  double temp_step = DBL_MAX;
  if (oracle_cubic_roots(coef, re, im) == 0) {
    for (int i = 0; i < 3; i++)
      if (re[i] > 0 && re[i] < temp_step && fabs(im[i]) < 1e-5) temp_step = re[i];
  }
  float step;
  if (temp_step > p->max_step)
    step = p->max_step;
  else if (temp_step < p->min_step)
    step = p->min_step;
  else
    step = (float)temp_step;
  return step;
}

/* per-row outputs of the two passes above (the reference keeps them in omega_gpu / v_gpu and
 * B..E device vectors before its thrust reductions): what tests/test_ref_pin*.py compare with the
 * reference's own K2 / K3+K4.  Not thread-safe (test taps). */
void oracle_flow_rows(const cvo_b200_params* p, const oracle_cloud* src, const float* y_moved,
                      const oracle_sparse* A, double* omega_rows, double* v_rows) {
  double os[3], vs[3];
  float o[3], v[3];
  g_rows_out[0] = omega_rows;
  g_rows_out[1] = v_rows;
  oracle_compute_flow(p, src, y_moved, A, os, vs, o, v);
  g_rows_out[0] = g_rows_out[1] = NULL;
}
void oracle_step_rows(const cvo_b200_params* p, const oracle_cloud* src, const float* y_moved, int m,
                      const oracle_sparse* A, const float omega[3], const float v[3], float ell,
                      double* bcde_rows /* rows x 4 */) {
  double bcde[4];
  g_rows_out[0] = bcde_rows;
  g_rows_out[1] = NULL;
  (void)oracle_compute_step(p, src, y_moved, m, A, omega, v, ell, bcde);
  g_rows_out[0] = NULL;
}

/* ---------- Lie group ---------------------------------------------------- */
/* LieGroup.cpp:245-274 (float; `using namespace std` makes sin/cos the float overloads) */
void oracle_exp_sek3(const float xi[6], float dt, float out12[12]) {
  const float TOLERANCE = 1e-6f;
  float R[9], Jl[9];
  const float* w = xi;
  float theta = sqrtf(dot3f(w, w));
  const float I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (theta < TOLERANCE) {
    memcpy(R, I, sizeof(I));
    memcpy(Jl, I, sizeof(I));
  } else {
    float A[9], A2[9];
    skewf(w, A);
    float theta2 = theta * theta;
    float stheta = sinf(dt * theta);
    float ctheta = cosf(dt * theta);
    float oneMinusCosTheta2 = (1 - ctheta) / (theta2);
    mat3f_mul(A, A, A2);
    float c1 = stheta / theta;
    float c3 = (dt * theta - stheta) / (theta2 * theta);
    for (int k = 0; k < 9; k++) {
      R[k] = (I[k] + c1 * A[k]) + oneMinusCosTheta2 * A2[k];
      Jl[k] = (dt * I[k] + oneMinusCosTheta2 * A[k]) + c3 * A2[k];
    }
  }
  memcpy(out12, R, sizeof(R));
  mat3f_vec(Jl, xi + 3, out12 + 9);
}

/* Sophus 1.0 SE3d(Matrix4d).log().norm() restated in closed form (library source
 * not in the reference tree; call site CvoGPU.cu:1473-1476).  Rotation matrix ->
 * unit quaternion (Shepperd/Eigen branch order) -> so(3) log -> V^{-1} t. */
double oracle_se3_log_norm(const double R[9], const double t[3]) {
#define M(i, j) R[3 * (j) + (i)]
  double q[4]; /* w x y z */
  double tr = M(0, 0) + M(1, 1) + M(2, 2);
  if (tr > 0.0) {
    double s = sqrt(tr + 1.0);
    q[0] = 0.5 * s;
    s = 0.5 / s;
    q[1] = (M(2, 1) - M(1, 2)) * s;
    q[2] = (M(0, 2) - M(2, 0)) * s;
    q[3] = (M(1, 0) - M(0, 1)) * s;
  } else {
    int i = 0;
    if (M(1, 1) > M(0, 0)) i = 1;
    if (M(2, 2) > M(i, i)) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    double s = sqrt(M(i, i) - M(j, j) - M(k, k) + 1.0);
    q[1 + i] = 0.5 * s;
    s = 0.5 / s;
    q[0] = (M(k, j) - M(j, k)) * s;
    q[1 + j] = (M(j, i) + M(i, j)) * s;
    q[1 + k] = (M(k, i) + M(i, k)) * s;
  }
#undef M
  double qn = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; i++) q[i] /= qn;
  const double eps = 1e-10;
  double sq_n = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  double n = sqrt(sq_n), w = q[0];
  double two_atan_nbyw_by_n;
  if (sq_n < eps * eps) {
    two_atan_nbyw_by_n = 2.0 / w - 2.0 / 3.0 * (sq_n) / (w * w * w);
  } else if (fabs(w) < eps) {
    two_atan_nbyw_by_n = (w > 0 ? M_PI : -M_PI) / n;
  } else {
    two_atan_nbyw_by_n = 2.0 * atan(n / w) / n;
  }
  double theta = two_atan_nbyw_by_n * n;
  double om[3] = {two_atan_nbyw_by_n * q[1], two_atan_nbyw_by_n * q[2],
                  two_atan_nbyw_by_n * q[3]};
  /* V^{-1} = I - 0.5*W + k*W^2 */
  double W[9] = {0, om[2], -om[1], -om[2], 0, om[0], om[1], -om[0], 0};
  double W2[9];
  for (int j = 0; j < 3; j++)
    for (int i = 0; i < 3; i++)
      W2[3 * j + i] = W[i] * W[3 * j] + W[3 + i] * W[3 * j + 1] + W[6 + i] * W[3 * j + 2];
  double kk;
  if (fabs(theta) < eps) {
    kk = 1.0 / 12.0;
  } else {
    double half = 0.5 * theta;
    kk = (1.0 - theta * cos(half) / (2.0 * sin(half))) / (theta * theta);
  }
  double up[3];
  for (int i = 0; i < 3; i++) {
    double s = 0;
    for (int j = 0; j < 3; j++) {
      double vij = (i == j ? 1.0 : 0.0) - 0.5 * W[3 * j + i] + kk * W2[3 * j + i];
      s += vij * t[j];
    }
    up[i] = s;
  }
  return sqrt(up[0] * up[0] + up[1] * up[1] + up[2] * up[2] + om[0] * om[0] + om[1] * om[1] +
              om[2] * om[2]);
}

/* ---------- indicator queues (CvoGPU.cu:1167-1285) ------------------------ */
typedef struct fqueue {
  float* buf;
  int cap, head, size;
} fqueue;
static void fq_init(fqueue* q, int cap) {
  q->cap = cap > 0 ? cap + 1 : 2;
  q->buf = (float*)malloc(sizeof(float) * (size_t)q->cap);
  q->head = q->size = 0;
}
static void fq_free(fqueue* q) { free(q->buf); }
static void fq_push(fqueue* q, float v) {
  q->buf[(q->head + q->size) % q->cap] = v;
  q->size++;
}
static float fq_front(const fqueue* q) { return q->buf[q->head]; }
static void fq_pop(fqueue* q) {
  q->head = (q->head + 1) % q->cap;
  q->size--;
}
static void fq_clear(fqueue* q) { q->head = q->size = 0; }

static int sparsity_indicator_ell_update(fqueue* start_q, fqueue* end_q, float* start_sum,
                                         float* end_sum, const float indicator,
                                         const cvo_b200_params* params) {
  int decrease = 0;
  int queue_len = params->indicator_window_size;
  if (start_q->size < queue_len) {
    fq_push(start_q, indicator);
    *start_sum += indicator;
  }
  if (start_q->size >= queue_len && end_q->size < queue_len) {
    fq_push(end_q, indicator);
    *end_sum += indicator;
  }
  if (start_q->size >= queue_len && end_q->size >= queue_len) {
    if (*end_sum / *start_sum > 1 - params->indicator_stable_threshold &&
        *end_sum / *start_sum < 1 + params->indicator_stable_threshold) {
      decrease = 1;
      fq_clear(start_q);
      fq_clear(end_q);
      *start_sum = 0;
      *end_sum = 0;
    } else {
      *end_sum -= fq_front(end_q);
      *start_sum += fq_front(end_q);
      fq_push(start_q, fq_front(end_q));
      fq_pop(end_q);
      *start_sum -= fq_front(start_q);
      fq_pop(start_q);
      fq_push(end_q, indicator);
      *end_sum += indicator;
    }
  }
  return decrease;
}

/* Test tap: the queues of one align() (both empty, sums 0 at entry, CvoGPU.cu:1377-1380) fed with
 * a sequence of indicator values; per call the decision and the two running sums after it. */
void oracle_indicator_sequence(const cvo_b200_params* params, int n, const float* indicators,
                               int* decrease, float* start_sums, float* end_sums) {
  fqueue start_q, end_q;
  fq_init(&start_q, params->indicator_window_size + 2);
  fq_init(&end_q, params->indicator_window_size + 2);
  float start_sum = 0, end_sum = 0;
  for (int k = 0; k < n; k++) {
    decrease[k] = sparsity_indicator_ell_update(&start_q, &end_q, &start_sum, &end_sum, indicators[k], params);
    start_sums[k] = start_sum;
    end_sums[k] = end_sum;
  }
  fq_free(&start_q);
  fq_free(&end_q);
}

/* ---------- one iteration / align ---------------------------------------- */
static double sparse_sum_double(const oracle_sparse* A) {
  double s = 0;
  for (int i = 0; i < A->rows; i++)
    for (unsigned int j = 0; j < A->nonzeros[i]; j++) s += A->mat[(size_t)i * A->stride + j];
  return s;
}
static unsigned int sparse_max_row(const oracle_sparse* A) {
  unsigned int mx = 0;
  for (int i = 0; i < A->rows; i++)
    if (A->nonzeros[i] > mx) mx = A->nonzeros[i];
  return mx;
}

/* pose update CvoGPU.cu:1460-1476; R,T updated in place; returns dist */
static double pose_update(const float omega[3], const float v[3], float step, float R[9],
                          float T[3]) {
  float vec_joined[6] = {omega[0], omega[1], omega[2], v[0], v[1], v[2]};
  float dtrans[12];
  oracle_exp_sek3(vec_joined, step, dtrans);
  double dR[9], dT[3], Rd[9], Td[3];
  for (int k = 0; k < 9; k++) {
    dR[k] = (double)dtrans[k];
    Rd[k] = (double)R[k];
  }
  for (int k = 0; k < 3; k++) {
    dT[k] = (double)dtrans[9 + k];
    Td[k] = (double)T[k];
  }
  for (int i = 0; i < 3; i++)
    T[i] = (float)(sum3d(Rd[i] * dT[0], Rd[3 + i] * dT[1], Rd[6 + i] * dT[2]) + Td[i]);
  for (int j = 0; j < 3; j++)
    for (int i = 0; i < 3; i++)
      R[3 * j + i] = (float)sum3d(Rd[i] * dR[3 * j], Rd[3 + i] * dR[3 * j + 1],
                                  Rd[6 + i] * dR[3 * j + 2]);
  return oracle_se3_log_norm(dR, dT);
}

void oracle_iterate(const cvo_b200_params* p, const oracle_cloud* src, const oracle_cloud* tgt,
                    const float R[9], const float T[3], float ell, int num_neighbors,
                    cvo_b200_iter_trace* tr, oracle_sparse* A_out) {
  memset(tr, 0, sizeof(*tr));
  oracle_sparse* A = A_out;
  int cap = num_neighbors > 1 ? num_neighbors : 1;
  if (!A || A->capacity < cap || A->rows != src->n) A = oracle_sparse_new(src->n, cap);
  float Rinv[9], Tinv[3];
  oracle_update_tf(R, T, Rinv, Tinv, NULL);
  float* y = (float*)malloc(sizeof(float) * 3 * (size_t)(tgt->n > 0 ? tgt->n : 1));
  oracle_transform(Rinv, Tinv, tgt->xyz, tgt->n, y);
  oracle_fill_A(p, src, tgt, y, num_neighbors, ell, A);
  tr->num_neighbors = num_neighbors;
  tr->ell = ell;
  tr->nnz = A->nonzero_sum;
  tr->max_row_nnz = sparse_max_row(A);
  tr->a_sum = sparse_sum_double(A);
  oracle_compute_flow(p, src, y, A, tr->omega_sum, tr->v_sum, tr->omega, tr->v);
  double bcde[4];
  tr->step = oracle_compute_step(p, src, y, tgt->n, A, tr->omega, tr->v, ell, bcde);
  tr->B = bcde[0]; tr->C = bcde[1]; tr->D = bcde[2]; tr->E = bcde[3];
  float Rn[9], Tn[3];
  memcpy(Rn, R, sizeof(Rn));
  memcpy(Tn, T, sizeof(Tn));
  tr->dist = pose_update(tr->omega, tr->v, tr->step, Rn, Tn);
  memcpy(tr->R, Rn, sizeof(Rn));
  memcpy(tr->T, Tn, sizeof(Tn));
  tr->ell_next = ell;
  int nn_next = (int)(tr->max_row_nnz * 1.2);
  tr->num_neighbors_next = p->nearest_neighbors_max < nn_next ? p->nearest_neighbors_max : nn_next;
  free(y);
  if (A != A_out) oracle_sparse_free(A);
}

int oracle_align(const cvo_b200_params* params, const oracle_cloud* src,
                 const oracle_cloud* tgt, const float T_init[16], float T_out[16],
                 cvo_b200_align_info* info, cvo_b200_iter_trace* trace, int trace_cap) {
  if (info) memset(info, 0, sizeof(*info));
  /* CvoGPU.cu:1614-1617: empty input -> return 0, transform untouched */
  if (src->n == 0 || tgt->n == 0) return 0;
  /* :1363-1364 */
  float R[9], T[3];
  for (int j = 0; j < 3; j++)
    for (int i = 0; i < 3; i++) R[3 * j + i] = T_init[4 * j + i];
  for (int i = 0; i < 3; i++) T[i] = T_init[12 + i];
  int ret = 0;
  float omega[3] = {0, 0, 0}, v[3] = {0, 0, 0};
  fqueue start_q, end_q;
  fq_init(&start_q, params->indicator_window_size + 2);
  fq_init(&end_q, params->indicator_window_size + 2);
  float start_sum = 0, end_sum = 0;
  float ell = params->ell_init; /* CvoState.cu:30 */
  int k = 0;
  int num_neighbors = params->nearest_neighbors_max; /* :1385 (is_using_kdtree unsupported) */
  int capA = params->nearest_neighbors_max > 1 ? params->nearest_neighbors_max : 1;
  oracle_sparse* A = oracle_sparse_new(src->n, capA);
  float* y = (float*)malloc(sizeof(float) * 3 * (size_t)tgt->n);
  float Rinv[9], Tinv[3];
  int stop_reason = CVO_B200_STOP_MAX_ITER;
  for (; k < params->MAX_ITER; k++) {
    cvo_b200_iter_trace trl;
    cvo_b200_iter_trace* tr = (trace && k < trace_cap) ? &trace[k] : &trl;
    memset(tr, 0, sizeof(*tr));
    tr->iter = k;
    tr->ell = ell;
    tr->num_neighbors = num_neighbors;
    oracle_update_tf(R, T, Rinv, Tinv, T_out);            /* :1393 */
    oracle_transform(Rinv, Tinv, tgt->xyz, tgt->n, y);     /* :1404 */
    oracle_fill_A(params, src, tgt, y, num_neighbors, ell, A); /* :1418 */
    tr->nnz = A->nonzero_sum;
    tr->max_row_nnz = sparse_max_row(A);
    tr->a_sum = sparse_sum_double(A);
    oracle_compute_flow(params, src, y, A, tr->omega_sum, tr->v_sum, omega, v); /* :1440 */
    memcpy(tr->omega, omega, sizeof(omega));
    memcpy(tr->v, v, sizeof(v));
    double bcde[4];
    float step = oracle_compute_step(params, src, y, tgt->n, A, omega, v, ell, bcde); /* :1448 */
    tr->B = bcde[0]; tr->C = bcde[1]; tr->D = bcde[2]; tr->E = bcde[3];
    tr->step = step;
    /* :1454-1458 */
    double on = sqrt(sum3d((double)omega[0] * omega[0], (double)omega[1] * omega[1],
                           (double)omega[2] * omega[2]));
    double vn = sqrt(sum3d((double)v[0] * v[0], (double)v[1] * v[1], (double)v[2] * v[2]));
    if (on < params->eps && vn < params->eps) {
      float onf = sqrtf(dot3f(omega, omega)), vnf = sqrtf(dot3f(v, v));
      stop_reason = CVO_B200_STOP_GRAD_SMALL;
      if (onf < 1e-8 && vnf < 1e-8) {
        ret = -1;
        stop_reason |= CVO_B200_STOP_GRAD_ZERO;
      }
      tr->flags = stop_reason;
      memcpy(tr->R, R, sizeof(R));
      memcpy(tr->T, T, sizeof(T));
      tr->ell_next = ell;
      tr->num_neighbors_next = num_neighbors;
      break;
    }
    /* :1460-1476 */
    double dist_this_iter = pose_update(omega, v, step, R, T);
    tr->dist = dist_this_iter;
    memcpy(tr->R, R, sizeof(R));
    memcpy(tr->T, T, sizeof(T));
    /* :1486-1493 */
    float ip_curr = (float)((double)A->nonzero_sum / sqrt((double)src->n * (double)tgt->n));
    int need_decay_ell =
        sparsity_indicator_ell_update(&start_q, &end_q, &start_sum, &end_sum, ip_curr, params);
    /* :1505-1508 */
    if (dist_this_iter < params->eps_2) {
      stop_reason = CVO_B200_STOP_DIST_SMALL;
      tr->flags = stop_reason;
      tr->ell_next = ell;
      tr->num_neighbors_next = num_neighbors;
      break;
    }
    /* :1509-1513 */
    if (k > params->ell_decay_start && need_decay_ell) {
      ell = ell * params->ell_decay_rate;
      if (ell < params->ell_min) ell = params->ell_min;
      tr->flags |= CVO_B200_ELL_DECAYED;
    }
    /* :1518-1529 */
    {
      unsigned int max_ind_val = tr->max_row_nnz;
      int cand = (int)(max_ind_val * 1.2);
      num_neighbors = params->nearest_neighbors_max < cand ? params->nearest_neighbors_max : cand;
    }
    tr->ell_next = ell;
    tr->num_neighbors_next = num_neighbors;
  }
  /* :1562 */
  oracle_update_tf(R, T, Rinv, Tinv, T_out);
  if (info) {
    info->ret = ret;
    info->iterations = k;
    info->stop_reason = (k >= params->MAX_ITER) ? CVO_B200_STOP_MAX_ITER : stop_reason;
    info->final_num_neighbors = num_neighbors;
    info->final_ell = ell;
    info->pairs_tested = (uint64_t)src->n * (uint64_t)tgt->n *
                         (uint64_t)(k < params->MAX_ITER ? k + 1 : k);
  }
  free(y);
  oracle_sparse_free(A);
  fq_free(&start_q);
  fq_free(&end_q);
  return ret;
}

/* ---------- inner product / angle --------------------------------------- */
/* 3x3 inverse via adjugate in float (Eigen's Matrix3f::inverse() uses the
 * cofactor formula for fixed size 3; CvoGPU.cu:1757/1946) */
static void mat3f_inverse(const float* m, float* out) {
#define A(i, j) m[3 * (j) + (i)]
  float c00 = A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1);
  float c10 = A(1, 2) * A(2, 0) - A(1, 0) * A(2, 2);
  float c20 = A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0);
  float det = sum3f(A(0, 0) * c00, A(0, 1) * c10, A(0, 2) * c20);
  float invdet = 1.0f / det;
  out[0] = c00 * invdet;
  out[1] = c10 * invdet;
  out[2] = c20 * invdet;
  out[3] = (A(0, 2) * A(2, 1) - A(0, 1) * A(2, 2)) * invdet;
  out[4] = (A(0, 0) * A(2, 2) - A(0, 2) * A(2, 0)) * invdet;
  out[5] = (A(2, 0) * A(0, 1) - A(0, 0) * A(2, 1)) * invdet;
  out[6] = (A(0, 1) * A(1, 2) - A(0, 2) * A(1, 1)) * invdet;
  out[7] = (A(1, 0) * A(0, 2) - A(0, 0) * A(1, 2)) * invdet;
  out[8] = (A(0, 0) * A(1, 1) - A(1, 0) * A(0, 1)) * invdet;
#undef A
}

float oracle_inner_product(const cvo_b200_params* p, const oracle_cloud* src,
                           const oracle_cloud* tgt, const float T16[16], float ell,
                           const float* kernel3x3, oracle_sparse* A_out) {
  float R[9], T[3], Rinv[9], Tinv[3];
  for (int j = 0; j < 3; j++)
    for (int i = 0; i < 3; i++) R[3 * j + i] = T16[4 * j + i];
  for (int i = 0; i < 3; i++) T[i] = T16[12 + i];
  oracle_update_tf(R, T, Rinv, Tinv, NULL);
  float* y = (float*)malloc(sizeof(float) * 3 * (size_t)(tgt->n > 0 ? tgt->n : 1));
  oracle_transform(Rinv, Tinv, tgt->xyz, tgt->n, y);
  int cap = p->nearest_neighbors_max;
  oracle_sparse* A = A_out;
  if (!A || A->capacity < cap || A->rows != src->n) A = oracle_sparse_new(src->n, cap > 1 ? cap : 1);
  if (kernel3x3) {
    /* CvoGPU.cu:1946-1949: inverse of the user kernel, geometric type switched off */
    float kinv[9];
    mat3f_inverse(kernel3x3, kinv);
    cvo_b200_params q = *p;
    q.is_using_geometric_type = 0;
    oracle_fill_A_dense_kernel(&q, src, tgt, y, cap, kinv, A);
  } else {
    oracle_fill_A(p, src, tgt, y, cap, ell, A);
  }
  /* SparseKernelMat.cu:62-66: float reduce over rows*cols (zeros included).
   * thrust's order is unspecified; ascending order in float here. */
  float s = 0.f;
  size_t n = (size_t)A->rows * (size_t)A->stride;
  for (size_t t = 0; t < n; t++) s += A->mat[t];
  free(y);
  if (A != A_out) oracle_sparse_free(A);
  return s;
}

/* ---------- pose-graph edge (multi-frame IRLS) ---------------------------- */
/* CvoGPU_impl.cu:84-150 transform_point_pose_vec (xyz only matter):
 * trans = T * [x y z 1]^T with T a row-major 3x4 float map of pose_vec.  Eigen's
 * unrolled coefficient redux over 4 terms sums (c0 + c1) + (c2 + c3). */
void oracle_transform_pose_vec(const float pose12[12], const float* x, int n, float* x_out) {
  for (int i = 0; i < n; i++) {
    const float* q = x + 3 * (size_t)i;
    for (int r = 0; r < 3; r++) {
      volatile float c0 = pose12[4 * r + 0] * q[0];
      volatile float c1 = pose12[4 * r + 1] * q[1];
      volatile float c2 = pose12[4 * r + 2] * q[2];
      volatile float c3 = pose12[4 * r + 3] * 1.0f;
      volatile float s01 = c0 + c1;
      volatile float s23 = c2 + c3;
      x_out[3 * (size_t)i + r] = s01 + s23;
    }
  }
}

/* IRLS.cpp:104-108 (every frame: transform_pointcloud, CvoFrameGPU.cu:44-62) followed by
 * IRLS_State_GPU.cu:43-79 BinaryStateGPU::update_inner_product for ONE edge:
 * clear_SparseKernelMat(num_neighbors), fill_in_A_mat_gpu on the two moved clouds with the
 * edge's ell, compute_nonzeros.  Returns nonzero_sum.  A needs capacity >= num_neighbors. */
unsigned long long oracle_edge_update(const cvo_b200_params* p, const oracle_cloud* f1,
                                      const float pose1[12], const oracle_cloud* f2,
                                      const float pose2[12], float ell, int num_neighbors,
                                      oracle_sparse* A) {
  float* x = (float*)malloc(sizeof(float) * 3 * (size_t)(f1->n > 0 ? f1->n : 1));
  float* y = (float*)malloc(sizeof(float) * 3 * (size_t)(f2->n > 0 ? f2->n : 1));
  oracle_transform_pose_vec(pose1, f1->xyz, f1->n, x);
  oracle_transform_pose_vec(pose2, f2->xyz, f2->n, y);
  oracle_cloud moved1 = *f1; /* features, labels, geometric types travel with the point */
  moved1.xyz = x;
  oracle_fill_A(p, &moved1, f2, y, num_neighbors, ell, A);
  unsigned long long s = A->nonzero_sum;
  free(x);
  free(y);
  return s;
}

float oracle_function_angle(const cvo_b200_params* p, const oracle_cloud* src,
                            const oracle_cloud* tgt, const float T[16], float ell,
                            int is_approximate) {
  if (src->n == 0 || tgt->n == 0) return 0;
  const float I16[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  float fxfz = oracle_inner_product(p, src, tgt, T, ell, NULL, NULL);
  float fx_norm, fz_norm;
  if (is_approximate) {
    fx_norm = sqrt(src->n);
    fz_norm = sqrt(tgt->n);
  } else {
    fx_norm = sqrt(oracle_inner_product(p, src, src, I16, ell, NULL, NULL));
    fz_norm = sqrt(oracle_inner_product(p, tgt, tgt, I16, ell, NULL, NULL));
  }
  return fxfz / (fx_norm * fz_norm);
}
