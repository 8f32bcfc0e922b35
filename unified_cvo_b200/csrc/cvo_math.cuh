// cvo_math.cuh — small fixed-size math shared by the sparse kernels and the
// single-thread controller.  This translation unit is compiled with
// --fmad=false, so every expression below is evaluated without FMA contraction,
// in the float/double mix the cited reference lines use.  Three-term sums are
// c0 + (c1 + c2) (the order Eigen 3.3's unrolled redux produces for length 3);
// DESIGN.md records this as a convention the reference leaves to Eigen.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace cvo_b200 {

__device__ __forceinline__ float sum3f(float c0, float c1, float c2) { return c0 + (c1 + c2); }
__device__ __forceinline__ double sum3d(double c0, double c1, double c2) { return c0 + (c1 + c2); }
__device__ __forceinline__ float dot3f(const float* a, const float* b) {
  return sum3f(a[0] * b[0], a[1] * b[1], a[2] * b[2]);
}
// column-major 3x3 * vector
__device__ __forceinline__ void mat3f_vec(const float* M, const float* x, float* out) {
#pragma unroll
  for (int i = 0; i < 3; i++) out[i] = sum3f(M[i] * x[0], M[3 + i] * x[1], M[6 + i] * x[2]);
}
__device__ __forceinline__ void mat3f_mul(const float* A, const float* B, float* out) {
  float tmp[9];
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int i = 0; i < 3; i++)
      tmp[3 * j + i] = sum3f(A[i] * B[3 * j], A[3 + i] * B[3 * j + 1], A[6 + i] * B[3 * j + 2]);
#pragma unroll
  for (int k = 0; k < 9; k++) out[k] = tmp[k];
}
// gpu_utils.cuh:8-15 skew_gpu, column-major storage
__device__ __forceinline__ void skewf(const float* v, float* M) {
  M[0] = 0.f;   M[3] = -v[2]; M[6] = v[1];
  M[1] = v[2];  M[4] = 0.f;   M[7] = -v[0];
  M[2] = -v[1]; M[5] = v[0];  M[8] = 0.f;
}
__device__ __forceinline__ void cross3f(const float* a, const float* b, float* out) {
  out[0] = a[1] * b[2] - a[2] * b[1];
  out[1] = a[2] * b[0] - a[0] * b[2];
  out[2] = a[0] * b[1] - a[1] * b[0];
}

// ---- cubic: roots of c0 t^3 + c1 t^2 + c2 t + c3 (LieGroup.cpp:309-325 builds the
// companion matrix of the same polynomial and takes its eigenvalues) -------------
__device__ inline double poly3_eval(double p2, double p1, double p0, double t) {
  return ((t + p2) * t + p1) * t + p0;
}
__device__ inline double poly3_polish(double p2, double p1, double p0, double t) {
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
// returns false when the companion matrix would not be finite (no usable root)
__device__ inline bool cubic_roots(const double coef[4], double re[3], double im[3]) {
  double p2 = coef[1] / coef[0], p1 = coef[2] / coef[0], p0 = coef[3] / coef[0];
  if (!isfinite(p2) || !isfinite(p1) || !isfinite(p0)) return false;
  double s = fabs(p2);
  if (sqrt(fabs(p1)) > s) s = sqrt(fabs(p1));
  if (cbrt(fabs(p0)) > s) s = cbrt(fabs(p0));
  if (s == 0.0) {
    for (int i = 0; i < 3; i++) re[i] = im[i] = 0.0;
    return true;
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
  return true;
}

// LieGroup.cpp:245-274 Exp_SEK3 (K=1), float.  sin/cos: the reference calls the
// host float overloads; (float)sin((double)x) is the correctly rounded float in
// all but double-rounding corner cases, which is what glibc's sinf/cosf deliver.
__device__ inline void exp_sek3(const float xi[6], float dt, float out12[12]) {
  const float TOLERANCE = 1e-6f;
  float R[9], Jl[9];
  const float I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  float theta = sqrtf(dot3f(xi, xi));
  if (theta < TOLERANCE) {
    for (int k = 0; k < 9; k++) R[k] = Jl[k] = I[k];
  } else {
    float A[9], A2[9];
    skewf(xi, A);
    float theta2 = theta * theta;
    float arg = dt * theta;
    float stheta = (float)sin((double)arg);
    float ctheta = (float)cos((double)arg);
    float oneMinusCosTheta2 = (1 - ctheta) / (theta2);
    mat3f_mul(A, A, A2);
    float c1 = stheta / theta;
    float c3 = (dt * theta - stheta) / (theta2 * theta);
    for (int k = 0; k < 9; k++) {
      R[k] = (I[k] + c1 * A[k]) + oneMinusCosTheta2 * A2[k];
      Jl[k] = (dt * I[k] + oneMinusCosTheta2 * A[k]) + c3 * A2[k];
    }
  }
  for (int k = 0; k < 9; k++) out12[k] = R[k];
  mat3f_vec(Jl, xi + 3, out12 + 9);
}

// || Sophus::SE3d(dRT).log() || in closed form (call site CvoGPU.cu:1473-1476)
__device__ inline double se3_log_norm(const double R[9], const double t[3]) {
#define CVO_M(i, j) R[3 * (j) + (i)]
  double q[4];
  double tr = CVO_M(0, 0) + CVO_M(1, 1) + CVO_M(2, 2);
  if (tr > 0.0) {
    double s = sqrt(tr + 1.0);
    q[0] = 0.5 * s;
    s = 0.5 / s;
    q[1] = (CVO_M(2, 1) - CVO_M(1, 2)) * s;
    q[2] = (CVO_M(0, 2) - CVO_M(2, 0)) * s;
    q[3] = (CVO_M(1, 0) - CVO_M(0, 1)) * s;
  } else {
    int i = 0;
    if (CVO_M(1, 1) > CVO_M(0, 0)) i = 1;
    if (CVO_M(2, 2) > CVO_M(i, i)) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    double s = sqrt(CVO_M(i, i) - CVO_M(j, j) - CVO_M(k, k) + 1.0);
    q[1 + i] = 0.5 * s;
    s = 0.5 / s;
    q[0] = (CVO_M(k, j) - CVO_M(j, k)) * s;
    q[1 + j] = (CVO_M(j, i) + CVO_M(i, j)) * s;
    q[1 + k] = (CVO_M(k, i) + CVO_M(i, k)) * s;
  }
#undef CVO_M
  double qn = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; i++) q[i] /= qn;
  const double eps = 1e-10;
  const double kPi = 3.14159265358979323846;
  double sq_n = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  double n = sqrt(sq_n), w = q[0];
  double f;
  if (sq_n < eps * eps) {
    f = 2.0 / w - 2.0 / 3.0 * (sq_n) / (w * w * w);
  } else if (fabs(w) < eps) {
    f = (w > 0 ? kPi : -kPi) / n;
  } else {
    f = 2.0 * atan(n / w) / n;
  }
  double theta = f * n;
  double om[3] = {f * q[1], f * q[2], f * q[3]};
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

}  // namespace cvo_b200
