/*
 * cvo_cpu_baseline.c — restatement of the reference's CPU registration class cvo::cvo
 * (src/cvo/Cvo.cpp), the "reference CPU align()" BASELINE.json config 1 and SURVEY.md §8(d) name
 * as the CPU path to time beside the GPU path.
 *
 * TEST / MEASUREMENT INFRASTRUCTURE ONLY (bench.py's cpu_baseline and --impl reference legs,
 * tests/test_cpu_baseline.py).  It is a TIMING baseline, not a bit-parity oracle: cvo::cvo is a
 * different algorithm from CvoGPU (kd-tree radius search rebuilt every iteration, no row cap,
 * no gradient normalisation, step capped at 0.8, A.sum()-based length-scale indicator), needs
 * Eigen / TBB / nanoflann / PCL to build and so cannot be compiled here; label its numbers
 * "restated reference CPU path".
 *
 * What follows which lines:
 *   kd-tree build per iteration + radius search per source point   Cvo.cpp:363-378  (nanoflann:
 *       KDTreeVectorOfVectorsAdaptor, leaf size 10, buildIndex() is single-threaded, results
 *       sorted by distance) -> kd_build / kd_radius below (same structure: midpoint split of the
 *       widest dimension, leaves of <= 10 points; own code, nanoflann is not restated line by line)
 *   kernel values, thresholds, a > sp_thres                         Cvo.cpp:353-360, 384-447
 *   sparse A from triplets (rows sorted by column)                  Cvo.cpp:453-455
 *   flow: omega, v = sums over rows of (1/c) A_i cross, (1/d) A_i diff, double accumulation
 *                                                                   Cvo.cpp:557-698
 *   step: xi powers per target, B..E per row, cubic, cap 0.8        Cvo.cpp:701-830
 *   transform_pcd                                                   Cvo.cpp:832-842
 *   align loop: break tests, Exp_SEK3, pose update, dist_se3        Cvo.cpp:885-970
 *   compute_indicator (A.sum() windows, decrease / increase)        Cvo.cpp:1287-1381
 * TBB parallel_for -> OpenMP parallel for; tbb::spin_mutex accumulation -> OpenMP reduction.
 * cvo::cvo reads 27 numbers from a text file (Cvo.cpp:93-120) that the reference does not ship;
 * the parameters are mapped from CvoParams instead: ell_reduced_1 = ell_decay_rate,
 * ell_reduced_2 = indicator_window_size, ell_reduced_3 = indicator_stable_threshold.
 * Built WITHOUT IS_USING_SEMANTICS / IS_USING_NORMALS like the reference's default flags
 * (sk = nk = 1) unless use_semantics is set.
 */
#include "cvo_cpu_baseline.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ------------------------------------------------------------------ kd-tree */
typedef struct {
  int left, right;   /* children (node indices), -1 for a leaf */
  int begin, end;    /* leaf: range of the index array */
  int dim;
  float lo, hi;      /* split: left holds values <= lo .. , right >= hi (nanoflann's divlow/divhigh) */
} kd_node;

typedef struct {
  const float* pts;  /* m x 3 */
  int* idx;
  kd_node* nodes;
  int n_nodes, cap_nodes;
} kd_tree;

static int kd_new_node(kd_tree* t) {
  if (t->n_nodes == t->cap_nodes) {
    t->cap_nodes = t->cap_nodes ? 2 * t->cap_nodes : 1024;
    t->nodes = (kd_node*)realloc(t->nodes, sizeof(kd_node) * (size_t)t->cap_nodes);
  }
  return t->n_nodes++;
}

static int kd_build_rec(kd_tree* t, int begin, int end) {
  const int id = kd_new_node(t);
  if (end - begin <= 10) { /* max leaf, Cvo.cpp:368 */
    kd_node nd = {-1, -1, begin, end, 0, 0.f, 0.f};
    t->nodes[id] = nd;
    return id;
  }
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int k = begin; k < end; k++) {
    const float* p = t->pts + 3 * (size_t)t->idx[k];
    for (int d = 0; d < 3; d++) {
      if (p[d] < lo[d]) lo[d] = p[d];
      if (p[d] > hi[d]) hi[d] = p[d];
    }
  }
  int dim = 0;
  for (int d = 1; d < 3; d++)
    if (hi[d] - lo[d] > hi[dim] - lo[dim]) dim = d;
  const float cut = 0.5f * (lo[dim] + hi[dim]);
  /* partition: < cut | >= cut; degenerate -> split in the middle of the range */
  int i = begin, j = end - 1;
  while (i <= j) {
    while (i <= j && t->pts[3 * (size_t)t->idx[i] + dim] < cut) i++;
    while (i <= j && t->pts[3 * (size_t)t->idx[j] + dim] >= cut) j--;
    if (i < j) {
      int tmp = t->idx[i]; t->idx[i] = t->idx[j]; t->idx[j] = tmp;
      i++; j--;
    }
  }
  int mid = i;
  if (mid == begin || mid == end) mid = (begin + end) / 2;
  float llo = -FLT_MAX, rhi = FLT_MAX;
  for (int k = begin; k < mid; k++) { float v = t->pts[3 * (size_t)t->idx[k] + dim]; if (v > llo) llo = v; }
  for (int k = mid; k < end; k++) { float v = t->pts[3 * (size_t)t->idx[k] + dim]; if (v < rhi) rhi = v; }
  const int l = kd_build_rec(t, begin, mid);
  const int r = kd_build_rec(t, mid, end);
  kd_node nd = {l, r, begin, end, dim, llo, rhi};
  t->nodes[id] = nd;
  return id;
}

static void kd_build(kd_tree* t, const float* pts, int m) {
  t->pts = pts;
  t->n_nodes = 0;
  for (int i = 0; i < m; i++) t->idx[i] = i;
  if (m > 0) kd_build_rec(t, 0, m);
}

typedef struct { int idx; float d2; } kd_match;
typedef struct { kd_match* v; int n, cap; } kd_result;

static void kd_push(kd_result* r, int idx, float d2) {
  if (r->n == r->cap) {
    r->cap = r->cap ? 2 * r->cap : 64;
    r->v = (kd_match*)realloc(r->v, sizeof(kd_match) * (size_t)r->cap);
  }
  r->v[r->n].idx = idx;
  r->v[r->n].d2 = d2;
  r->n++;
}

static void kd_radius_rec(const kd_tree* t, int id, const float* q, float r2, kd_result* out) {
  const kd_node* nd = &t->nodes[id];
  if (nd->left < 0) {
    for (int k = nd->begin; k < nd->end; k++) {
      const int j = t->idx[k];
      const float* p = t->pts + 3 * (size_t)j;
      const float dx = q[0] - p[0], dy = q[1] - p[1], dz = q[2] - p[2];
      const float d2 = dx * dx + dy * dy + dz * dz;
      if (d2 < r2) kd_push(out, j, d2);
    }
    return;
  }
  const float v = q[nd->dim];
  const float dl = v - nd->lo, dr = nd->hi - v; /* distance to the slabs of the two children */
  if (dl <= 0.f || dl * dl < r2) kd_radius_rec(t, nd->left, q, r2, out);
  if (dr <= 0.f || dr * dr < r2) kd_radius_rec(t, nd->right, q, r2, out);
}

static int cmp_match_d2(const void* a, const void* b) {
  const float x = ((const kd_match*)a)->d2, y = ((const kd_match*)b)->d2;
  return (x > y) - (x < y);
}
static int cmp_entry_col(const void* a, const void* b) {
  return ((const kd_match*)a)->idx - ((const kd_match*)b)->idx;
}

/* test tap: number of points of pts[m] with squared distance < r2 from each of q[n] */
void cpu_baseline_radius_counts(const float* pts, int m, const float* q, int n, float r2, int* counts) {
  kd_tree t;
  memset(&t, 0, sizeof(t));
  t.idx = (int*)malloc(sizeof(int) * (size_t)(m > 0 ? m : 1));
  kd_build(&t, pts, m);
  kd_result res = {NULL, 0, 0};
  for (int i = 0; i < n; i++) {
    res.n = 0;
    if (m > 0) kd_radius_rec(&t, 0, q + 3 * (size_t)i, r2, &res);
    counts[i] = res.n;
  }
  free(res.v);
  free(t.idx);
  free(t.nodes);
}

/* ------------------------------------------------------------------ small math */
static void mat3_vec(const float* M /* column-major */, const float* x, float* out) {
  for (int i = 0; i < 3; i++) out[i] = M[i] * x[0] + M[3 + i] * x[1] + M[6 + i] * x[2];
}
static void mat3_mul(const float* A, const float* B, float* out) {
  float t[9];
  for (int j = 0; j < 3; j++)
    for (int i = 0; i < 3; i++) t[3 * j + i] = A[i] * B[3 * j] + A[3 + i] * B[3 * j + 1] + A[6 + i] * B[3 * j + 2];
  memcpy(out, t, sizeof(t));
}
static void cross3(const float* a, const float* b, float* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
static float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* LieGroup.cpp:368-378 dist_se3(float): Frobenius norm of the matrix logarithm of [dR dT; 0 1]
 * = sqrt(2 |w|^2 + |u|^2), (w, u) = se(3) log in closed form */
static double dist_se3f(const float* dR /* column-major */, const float* dT) {
  const double tr = (double)dR[0] + dR[4] + dR[8];
  double c = 0.5 * (tr - 1.0);
  if (c > 1.0) c = 1.0;
  if (c < -1.0) c = -1.0;
  const double th = acos(c);
  double w[3] = {dR[5] - dR[7], dR[6] - dR[2], dR[1] - dR[3]}; /* R32-R23, R13-R31, R21-R12 */
  const double s = sin(th);
  const double f = (th < 1e-8) ? 0.5 : th / (2.0 * s);
  for (int k = 0; k < 3; k++) w[k] *= f;
  /* u = V^-1 t,  V^-1 = I - W/2 + k W^2,  k = (1 - th cos(th/2) / (2 sin(th/2))) / th^2 */
  const double kk = (th < 1e-8) ? 1.0 / 12.0 : (1.0 - th * cos(0.5 * th) / (2.0 * sin(0.5 * th))) / (th * th);
  const double t[3] = {dT[0], dT[1], dT[2]};
  const double wxt[3] = {w[1] * t[2] - w[2] * t[1], w[2] * t[0] - w[0] * t[2], w[0] * t[1] - w[1] * t[0]};
  const double wxwxt[3] = {w[1] * wxt[2] - w[2] * wxt[1], w[2] * wxt[0] - w[0] * wxt[2], w[0] * wxt[1] - w[1] * wxt[0]};
  double u2 = 0.0, w2 = 0.0;
  for (int k = 0; k < 3; k++) {
    const double u = t[k] - 0.5 * wxt[k] + kk * wxwxt[k];
    u2 += u * u;
    w2 += w[k] * w[k];
  }
  return (double)(float)sqrt(2.0 * w2 + u2);
}

/* ------------------------------------------------------------------ the class */
typedef struct {
  int qs_head, qs_n, qe_head, qe_n, cap;
  float* qs;
  float* qe;
  float start_sum, end_sum;
} indicator_queues;

static void q_push(float* q, int* head, int* n, int cap, float v) { q[(*head + *n) % cap] = v; (*n)++; }
static float q_front(const float* q, int head) { return q[head]; }
static void q_pop(int* head, int* n, int cap) { *head = (*head + 1) % cap; (*n)--; }

int cpu_baseline_align(const cvo_b200_params* p, const oracle_cloud* src, const oracle_cloud* tgt,
                       const float T_init[16], int use_semantics, float T_out[16],
                       cpu_baseline_info* info) {
  const int N = src->n, M = tgt->n;
  memset(info, 0, sizeof(*info));
#ifdef _OPENMP
  info->threads = omp_get_max_threads();
#else
  info->threads = 1;
#endif
  for (int k = 0; k < 16; k++) T_out[k] = (k % 5 == 0) ? 1.f : 0.f;
  if (N == 0 || M == 0) return 0; /* Cvo.cpp:1186-1188 */
  const double t_begin = now_s();

  /* set_pcd, Cvo.cpp:1182-1233 */
  float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, T[3] = {0, 0, 0};
  if (T_init) {
    for (int j = 0; j < 3; j++)
      for (int i = 0; i < 3; i++) R[3 * j + i] = T_init[4 * j + i];
    for (int i = 0; i < 3; i++) T[i] = T_init[12 + i];
  }
  float ell = p->ell_init;
  const float ell_min = p->ell_min, ell_max = p->ell_max;
  const float ell_reduced_1 = p->ell_decay_rate;
  const int ell_reduced_2 = p->indicator_window_size > 0 ? p->indicator_window_size : 1;
  const float ell_reduced_3 = p->indicator_stable_threshold;
  const float sigma = p->sigma, sp_thres = p->sp_thres, c_ell = p->c_ell, c_sigma = p->c_sigma;
  const float s_ell = p->s_ell, s_sigma = p->s_sigma;
  const int F = src->F < tgt->F ? src->F : tgt->F;
  const int Cc = src->C < tgt->C ? src->C : tgt->C;

  float* cloud_y = (float*)malloc(sizeof(float) * 3 * (size_t)M);
  kd_tree tree;
  memset(&tree, 0, sizeof(tree));
  tree.idx = (int*)malloc(sizeof(int) * (size_t)M);
  /* sparse A, CSR rebuilt every iteration */
  long long* row_ptr = (long long*)malloc(sizeof(long long) * ((size_t)N + 1));
  kd_match** row_entries = (kd_match**)calloc((size_t)N, sizeof(kd_match*)); /* (col, a) per row */
  int* row_n = (int*)calloc((size_t)N, sizeof(int));
  int* row_cap = (int*)calloc((size_t)N, sizeof(int));
  float* xi = (float*)malloc(sizeof(float) * 15 * (size_t)M); /* xiz, xi2z, xi3z, xi4z, 3 scalars */
  indicator_queues Q;
  memset(&Q, 0, sizeof(Q));
  Q.cap = ell_reduced_2 + 1;
  Q.qs = (float*)malloc(sizeof(float) * (size_t)Q.cap);
  Q.qe = (float*)malloc(sizeof(float) * (size_t)Q.cap);
  int decrease = 0, increase = 0;

  int ret = 0, iter = p->MAX_ITER;
  float transform_R[9], transform_T[3];
  long long nnz = 0;
  for (int k = 0; k < p->MAX_ITER; k++) {
    /* update_tf, Cvo.cpp:282-288: transform = [R', -R'T] */
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) transform_R[3 * j + i] = R[3 * i + j];
    {
      float t[3];
      mat3_vec(transform_R, T, t);
      for (int i = 0; i < 3; i++) transform_T[i] = -t[i];
    }
    /* transform_pcd, Cvo.cpp:832-842 */
    double t0 = now_s();
#pragma omp parallel for schedule(static)
    for (int j = 0; j < M; j++) {
      float r[3];
      mat3_vec(transform_R, tgt->xyz + 3 * (size_t)j, r);
      for (int i = 0; i < 3; i++) cloud_y[3 * (size_t)j + i] = r[i] + transform_T[i];
    }
    double t1 = now_s();
    info->t_transform += t1 - t0;

    /* ---- compute_flow: se_kernel (Cvo.cpp:349-456) */
    const float s2 = sigma * sigma;
    const float l = ell;
    const float d2_thres = -2.0 * l * l * logf(sp_thres / s2);
    const float d2_c_thres = -2.0 * c_ell * c_ell * logf(sp_thres / c_sigma / c_sigma);
    kd_build(&tree, cloud_y, M); /* single-threaded like nanoflann's buildIndex() */
    double t_kd = now_s();
    info->t_kdtree_build += t_kd - t1;
#pragma omp parallel
    {
      kd_result res = {NULL, 0, 0};
#pragma omp for schedule(dynamic, 32)
      for (int i = 0; i < N; i++) {
        const float* pa = src->xyz + 3 * (size_t)i;
        res.n = 0;
        kd_radius_rec(&tree, 0, pa, d2_thres, &res);       /* search_radius = d2_thres (squared) */
        qsort(res.v, (size_t)res.n, sizeof(kd_match), cmp_match_d2); /* SearchParams.sorted = true */
        int cnt = 0;
        for (int jj = 0; jj < res.n; jj++) {
          const int idx = res.v[jj].idx;
          const float d2 = res.v[jj].d2;
          if (d2 < d2_thres) {
            float d2_color = 0.f;
            for (int f = 0; f < F; f++) {
              const float df = src->feat[(size_t)i * src->F + f] - tgt->feat[(size_t)idx * tgt->F + f];
              d2_color += df * df;
            }
            if (d2_color < d2_c_thres) {
              const float kk = s2 * exp(-d2 / (2.0 * l * l));
              const float ck = c_sigma * c_sigma * exp(-d2_color / (2.0 * c_ell * c_ell));
              float sk = 1.f;
              if (use_semantics && Cc > 0) { /* #ifdef IS_USING_SEMANTICS */
                float d2_sem = 0.f;
                for (int c = 0; c < Cc; c++) {
                  const float dc = src->labels[(size_t)i * src->C + c] - tgt->labels[(size_t)idx * tgt->C + c];
                  d2_sem += dc * dc;
                }
                sk = s_sigma * s_sigma * exp(-d2_sem / (2.0 * s_ell * s_ell));
              }
              const float a = ck * kk * sk * 1.f;
              if (a > sp_thres) {
                if (cnt == row_cap[i]) {
                  row_cap[i] = row_cap[i] ? 2 * row_cap[i] : 16;
                  row_entries[i] = (kd_match*)realloc(row_entries[i], sizeof(kd_match) * (size_t)row_cap[i]);
                }
                row_entries[i][cnt].idx = idx;
                row_entries[i][cnt].d2 = a; /* the value slot holds a */
                cnt++;
              }
            }
          }
        }
        /* setFromTriplets + makeCompressed: rows sorted by column */
        qsort(row_entries[i], (size_t)cnt, sizeof(kd_match), cmp_entry_col);
        row_n[i] = cnt;
      }
      free(res.v);
    }
    row_ptr[0] = 0;
    for (int i = 0; i < N; i++) row_ptr[i + 1] = row_ptr[i] + row_n[i];
    nnz = row_ptr[N];
    double t2 = now_s();
    info->t_se_kernel += t2 - t_kd;

    /* flow, Cvo.cpp:585-690 */
    double om0 = 0, om1 = 0, om2 = 0, v0 = 0, v1 = 0, v2 = 0, a_sum = 0;
#pragma omp parallel for schedule(static) reduction(+ : om0, om1, om2, v0, v1, v2, a_sum)
    for (int i = 0; i < N; i++) {
      const float* px = src->xyz + 3 * (size_t)i;
      float so[3] = {0, 0, 0}, sv[3] = {0, 0, 0};
      for (int e = 0; e < row_n[i]; e++) {
        const float* py = cloud_y + 3 * (size_t)row_entries[i][e].idx;
        const float a = row_entries[i][e].d2;
        float cr[3];
        cross3(px, py, cr);
        for (int q = 0; q < 3; q++) {
          so[q] += a * cr[q];
          sv[q] += a * (py[q] - px[q]);
        }
        a_sum += (double)a;
      }
      om0 += (double)(1 / p->c * so[0]); om1 += (double)(1 / p->c * so[1]); om2 += (double)(1 / p->c * so[2]);
      v0 += (double)(1 / p->d * sv[0]); v1 += (double)(1 / p->d * sv[1]); v2 += (double)(1 / p->d * sv[2]);
    }
    const float omega[3] = {(float)om0, (float)om1, (float)om2};
    const float v[3] = {(float)v0, (float)v1, (float)v2};
    double t3 = now_s();
    info->t_flow += t3 - t2;

    /* compute_step_size, Cvo.cpp:701-830 */
    float W[9] = {0, omega[2], -omega[1], -omega[2], 0, omega[0], omega[1], -omega[0], 0};
    float W2[9], W3[9], W4[9], Wv[3], W2v[3], W3v[3];
    mat3_mul(W, W, W2);
    mat3_mul(W2, W, W3);
    mat3_mul(W3, W, W4);
    mat3_vec(W, v, Wv);
    mat3_vec(W2, v, W2v);
    mat3_vec(W3, v, W3v);
#pragma omp parallel for schedule(static)
    for (int j = 0; j < M; j++) {
      const float* y = cloud_y + 3 * (size_t)j;
      float* o = xi + 15 * (size_t)j;
      float t[3];
      cross3(omega, y, t);
      for (int q = 0; q < 3; q++) o[q] = t[q] + v[q];
      mat3_vec(W2, y, t);
      for (int q = 0; q < 3; q++) o[3 + q] = t[q] + Wv[q];
      mat3_vec(W3, y, t);
      for (int q = 0; q < 3; q++) o[6 + q] = t[q] + W2v[q];
      mat3_vec(W4, y, t);
      for (int q = 0; q < 3; q++) o[9 + q] = t[q] + W3v[q];
      o[12] = dot3(o, o);
      o[13] = -dot3(o, o + 3);
      o[14] = dot3(o + 3, o + 3) + 2 * dot3(o, o + 6);
    }
    const float temp_coef = 1 / (2.0 * ell * ell);
    double B = 0, C = 0, D = 0, E = 0;
#pragma omp parallel for schedule(static) reduction(+ : B, C, D, E)
    for (int i = 0; i < N; i++) {
      const float* px = src->xyz + 3 * (size_t)i;
      double Bi = 0, Ci = 0, Di = 0, Ei = 0;
      for (int e = 0; e < row_n[i]; e++) {
        const int idx = row_entries[i][e].idx;
        const float* py = cloud_y + 3 * (size_t)idx;
        const float* o = xi + 15 * (size_t)idx;
        const float diff[3] = {px[0] - py[0], px[1] - py[1], px[2] - py[2]};
        const float beta = (float)(-2.0 * temp_coef * dot3(o, diff));
        const float gamma = (float)(-temp_coef * (o[12] + 2.0 * dot3(o + 3, diff)));
        const float delta = (float)(2.0 * temp_coef * (o[13] + (-dot3(o + 6, diff))));
        const float epsil = (float)(-temp_coef * (o[14] + 2.0 * dot3(o + 9, diff)));
        const float A_ij = row_entries[i][e].d2;
        Bi += (double)(A_ij * beta);
        Ci += (double)(A_ij * (gamma + beta * beta / 2.0));
        Di += (double)(A_ij * (delta + beta * gamma + beta * beta * beta / 6.0));
        Ei += (double)(A_ij * (epsil + beta * delta + 1 / 2.0 * beta * beta * gamma + 1 / 2.0 * gamma * gamma +
                               1 / 24.0 * beta * beta * beta * beta));
      }
      B += Bi; C += Ci; D += Di; E += Ei;
    }
    /* p_coef in float (Cvo.cpp:797-798), roots of the companion matrix, smallest positive real */
    const double coef[4] = {4.0 * (float)E, 3.0 * (float)D, 2.0 * (float)C, (double)(float)B};
    double re[3], im[3];
    float temp_step = FLT_MAX;
    if (oracle_cubic_roots(coef, re, im) == 0)
      for (int r = 0; r < 3; r++)
        if ((float)re[r] > 0 && (float)re[r] < temp_step && (float)im[r] == 0) temp_step = (float)re[r];
    float step = temp_step == FLT_MAX ? p->min_step : temp_step;
    step = step > 0.8 ? 0.8 : step;
    double t4 = now_s();
    info->t_step += t4 - t3;

    /* Cvo.cpp:932-939 */
    const double on = sqrt((double)omega[0] * omega[0] + (double)omega[1] * omega[1] + (double)omega[2] * omega[2]);
    const double vn = sqrt((double)v[0] * v[0] + (double)v[1] * v[1] + (double)v[2] * v[2]);
    if (on < p->eps && vn < p->eps) {
      iter = k;
      if (on < 1e-8 && vn < 1e-8) ret = -1;
      break;
    }
    /* Cvo.cpp:941-955 */
    const float xi6[6] = {omega[0], omega[1], omega[2], v[0], v[1], v[2]};
    float d12[12];
    oracle_exp_sek3(xi6, step, d12);
    float RdT[3], newR[9];
    mat3_vec(R, d12 + 9, RdT);
    for (int i = 0; i < 3; i++) T[i] = RdT[i] + T[i];
    mat3_mul(R, d12, newR);
    memcpy(R, newR, sizeof(R));
    const double dist = dist_se3f(d12, d12 + 9);
    if (dist < p->eps_2) { /* Cvo.cpp:966-970 */
      iter = k;
      break;
    }
    /* compute_indicator, Cvo.cpp:1287-1381 */
    {
      const float indicator = (float)a_sum;
      if (Q.qs_n < ell_reduced_2) {
        q_push(Q.qs, &Q.qs_head, &Q.qs_n, Q.cap, indicator);
        Q.start_sum += indicator;
      } else if (Q.qe_n < ell_reduced_2) {
        q_push(Q.qe, &Q.qe_head, &Q.qe_n, Q.cap, indicator);
        Q.end_sum += indicator;
      } else if (fabsf(1 - Q.end_sum / Q.start_sum) < ell_reduced_3) {
        decrease = 1;
        Q.qs_n = Q.qe_n = Q.qs_head = Q.qe_head = 0;
        Q.start_sum = Q.end_sum = 0;
      } else if (Q.end_sum / Q.start_sum < 0.7) {
        increase = 1;
        Q.qs_n = Q.qe_n = Q.qs_head = Q.qe_head = 0;
        Q.start_sum = Q.end_sum = 0;
      } else {
        const float f = q_front(Q.qe, Q.qe_head);
        Q.end_sum -= f;
        Q.start_sum += f;
        q_push(Q.qs, &Q.qs_head, &Q.qs_n, Q.cap, f);
        q_pop(&Q.qe_head, &Q.qe_n, Q.cap);
        Q.start_sum -= q_front(Q.qs, Q.qs_head);
        q_pop(&Q.qs_head, &Q.qs_n, Q.cap);
        q_push(Q.qe, &Q.qe_head, &Q.qe_n, Q.cap, indicator);
        Q.end_sum += indicator;
      }
      if (decrease && ell > ell_min) {
        ell = ell * ell_reduced_1;
        decrease = 0;
      }
      if (increase && ell < ell_max) {
        ell = ell * 1 / ell_reduced_1;
        increase = 0;
      }
    }
    info->t_rest += now_s() - t4;
  }
  /* final update_tf (Cvo.cpp:1040): T_out = [R', -R'T; 0 0 0 1], column-major */
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) T_out[4 * j + i] = R[3 * i + j];
  {
    float Rt[9], t[3];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) Rt[3 * j + i] = R[3 * i + j];
    mat3_vec(Rt, T, t);
    for (int i = 0; i < 3; i++) T_out[12 + i] = -t[i];
  }
  T_out[3] = T_out[7] = T_out[11] = 0.f;
  T_out[15] = 1.f;
  info->ret = ret;
  info->iterations = iter;
  info->executed = iter < p->MAX_ITER ? iter + 1 : p->MAX_ITER;
  info->final_ell = ell;
  info->nnz_last = nnz;
  info->pairs = (unsigned long long)N * (unsigned long long)M * (unsigned long long)info->executed;
  info->seconds = now_s() - t_begin;

  for (int i = 0; i < N; i++) free(row_entries[i]);
  free(row_entries); free(row_n); free(row_cap); free(row_ptr);
  free(cloud_y); free(tree.idx); free(tree.nodes); free(xi); free(Q.qs); free(Q.qe);
  return ret;
}
