// ref_mini_eigen.h — OUR stand-in for the handful of Eigen fixed-size operations that the
// reference's kernels K1b, K2, K3, K4, its point transform and its host functions update_tf and
// Exp_SEK3 use (oracle/make_ref.py, tier 2).  TEST INFRASTRUCTURE.
//
// Eigen 3.3.9 (README.md:28 of the reference) is not in this image, so the tier-2 pin compiles the
// reference's own statements against this header instead.  What that pins: every formula, the
// float/double mix, the statement order, the scalar promotions (`2.0 * float_row` stays float:
// Eigen converts the literal to the expression's scalar).  What it CANNOT pin, and what DESIGN.md
// therefore still lists as a convention: the evaluation order INSIDE Eigen's fixed-size
// products / reductions, restated here as Eigen 3.3 implements them for sizes that are not
// vectorisable (3 floats):
//   * a reduction of n terms is the unrolled binary tree of redux_novec_unroller
//     (Eigen/src/Core/Redux.h): split at n/2, i.e. for n = 3: c0 + (c1 + c2);
//   * matrix * matrix / matrix * vector / row * column of these sizes are coefficient-based lazy
//     products: coeff(i,j) = that reduction over k of lhs(i,k) * rhs(k,j)
//     (Eigen/src/Core/ProductEvaluators.h, CoeffBasedProductMode);
//   * a chained product A*B*C*v is evaluated left to right, each factor materialised;
//   * cross(), squaredNorm(), dot(), norm() as in Eigen/src/Geometry/OrthoMethods.h and Dot.h.
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define ME_HD __host__ __device__ inline
#else
#define ME_HD inline
#endif

namespace Eigen {

template <typename T, int N, int Start, int Len>
struct redux_tree {
  template <typename F>
  ME_HD static T run(const F& f) {
    return redux_tree<T, N, Start, Len / 2>::run(f) +
           redux_tree<T, N, Start + Len / 2, Len - Len / 2>::run(f);
  }
};
template <typename T, int N, int Start>
struct redux_tree<T, N, Start, 1> {
  template <typename F>
  ME_HD static T run(const F& f) { return f(Start); }
};

enum { ColMajor = 0, RowMajor = 1 };  // Eigen/src/Core/util/Constants.h

template <typename T, int R, int C, int Opt = ColMajor>
struct Matrix;

// X.block<BR, BC>(i, j) = M  (Exp_SEK3, LieGroup.cpp:266-269; update_tf, CvoGPU.cu:102-103),
// X.block<BR, BC>(i, j) << a, b, ...  (update_tf, CvoGPU.cu:104: row-major fill of the block) and
// Matrix<...> R = T.block<3, 3>(0, 0)  (transform_point_pose_vec, CvoGPU_impl.cu:141).  P = the
// parent matrix type.
template <typename P, int BR, int BC>
struct BlockRef;
template <typename P, int BR, int BC>
struct BlockComma {
  BlockRef<P, BR, BC>* b;
  int k;
  ME_HD BlockComma& operator,(typename P::Scalar v) {
    b->set_rowmajor(k++, v);
    return *this;
  }
};
template <typename P, int BR, int BC>
struct BlockRef {
  typedef typename P::Scalar T;
  P* m;
  int i0, j0;
  template <int O>
  ME_HD BlockRef& operator=(const Matrix<T, BR, BC, O>& o) {
    for (int j = 0; j < BC; j++)
      for (int i = 0; i < BR; i++) (*m)(i0 + i, j0 + j) = o(i, j);
    return *this;
  }
  template <int O>
  ME_HD operator Matrix<T, BR, BC, O>() const {
    Matrix<T, BR, BC, O> r;
    for (int j = 0; j < BC; j++)
      for (int i = 0; i < BR; i++) r(i, j) = (*m)(i0 + i, j0 + j);
    return r;
  }
  ME_HD void set_rowmajor(int k, T v) { (*m)(i0 + k / BC, j0 + k % BC) = v; }
  ME_HD BlockComma<P, BR, BC> operator<<(T v) {
    set_rowmajor(0, v);
    return BlockComma<P, BR, BC>{this, 1};
  }
};

// Eigen::Map<M>(ptr): the caller's array viewed as an M in M's own storage order
// (transform_point_pose_vec, CvoGPU_impl.cu:117-119: a row-major 3x4 pose)
template <typename M>
struct Map {
  typename M::Scalar* p;
  ME_HD explicit Map(typename M::Scalar* q) : p(q) {}
  ME_HD operator M() const {
    M m;
    for (int i = 0; i < M::Size; i++) m.d[i] = p[i];
    return m;
  }
};

// Eigen::Ref<M> as a by-value function parameter (update_tf): a writable view of the caller's matrix
template <typename M>
using Ref = M&;

template <typename T, int R, int C>
struct CommaInit {
  Matrix<T, R, C>* m;
  int k;
  ME_HD CommaInit& operator,(T v) {
    m->set_rowmajor(k++, v);
    return *this;
  }
};

template <typename T, int R, int C, int Opt>
struct Matrix {
  typedef T Scalar;
  enum { Size = R * C };
  T d[R * C];  // in the storage order Opt names (vectors: contiguous)
  ME_HD Matrix() {}
  ME_HD Matrix(T a, T b, T c) { d[0] = a; d[1] = b; d[2] = c; }
  ME_HD static Matrix Zero() {
    Matrix m;
    for (int i = 0; i < R * C; i++) m.d[i] = T(0);
    return m;
  }
  ME_HD static Matrix Identity() {
    Matrix m = Zero();
    for (int i = 0; i < (R < C ? R : C); i++) m(i, i) = T(1);
    return m;
  }
  // v.head(3), v.segment<3>(k), X.block<3, 3>(i, j) = ... as Exp_SEK3 uses them (LieGroup.cpp:249-269)
  ME_HD Matrix<T, 3, 1> head(int n) const {
    Matrix<T, 3, 1> r;
    for (int i = 0; i < 3; i++) r.d[i] = (n == 3) ? d[i] : T(NAN);
    return r;
  }
  template <int N>
  ME_HD Matrix<T, N, 1> segment(int start) const {
    Matrix<T, N, 1> r;
    for (int i = 0; i < N; i++) r.d[i] = d[start + i];
    return r;
  }
  template <int BR, int BC>
  ME_HD BlockRef<Matrix, BR, BC> block(int i, int j) {
    return BlockRef<Matrix, BR, BC>{this, i, j};
  }
  ME_HD T& operator()(int i, int j) { return d[Opt == RowMajor ? i * C + j : j * R + i]; }
  ME_HD const T& operator()(int i, int j) const { return d[Opt == RowMajor ? i * C + j : j * R + i]; }
  ME_HD T& operator()(int i) { return d[i]; }
  ME_HD const T& operator()(int i) const { return d[i]; }
  ME_HD T& operator[](int i) { return d[i]; }
  ME_HD const T& operator[](int i) const { return d[i]; }
  ME_HD T* data() { return d; }
  ME_HD const T* data() const { return d; }
  ME_HD void set_rowmajor(int k, T v) { (*this)(k / C, k % C) = v; }
  ME_HD CommaInit<T, R, C> operator<<(T v) {
    set_rowmajor(0, v);
    return CommaInit<T, R, C>{this, 1};
  }
  ME_HD Matrix<T, C, R> transpose() const {
    Matrix<T, C, R> t;
    for (int i = 0; i < R; i++)
      for (int j = 0; j < C; j++) t(j, i) = (*this)(i, j);
    return t;
  }
  ME_HD Matrix operator+(const Matrix& o) const {
    Matrix r;
    for (int i = 0; i < R * C; i++) r.d[i] = d[i] + o.d[i];
    return r;
  }
  ME_HD Matrix operator-(const Matrix& o) const {
    Matrix r;
    for (int i = 0; i < R * C; i++) r.d[i] = d[i] - o.d[i];
    return r;
  }
  ME_HD Matrix operator-() const {
    Matrix r;
    for (int i = 0; i < R * C; i++) r.d[i] = -d[i];
    return r;
  }
  // scalar on the right; any arithmetic scalar is converted to the expression's scalar first
  template <typename S>
  ME_HD Matrix operator*(S s) const {
    Matrix r;
    const T st = (T)s;
    for (int i = 0; i < R * C; i++) r.d[i] = d[i] * st;
    return r;
  }
  template <typename S>
  ME_HD Matrix operator/(S s) const {
    Matrix r;
    const T st = (T)s;
    for (int i = 0; i < R * C; i++) r.d[i] = d[i] / st;
    return r;
  }
  template <int C2, int O2>
  ME_HD Matrix<T, R, C2> operator*(const Matrix<T, C, C2, O2>& o) const {
    Matrix<T, R, C2> r;
    for (int i = 0; i < R; i++)
      for (int j = 0; j < C2; j++) {
        const Matrix& a = *this;
        r(i, j) = redux_tree<T, C, 0, C>::run([&](int k) { return a(i, k) * o(k, j); });
      }
    return r;
  }
  ME_HD Matrix eval() const { return *this; }
  ME_HD T value() const {
    static_assert(R == 1 && C == 1, "value() needs a 1x1 expression");
    return d[0];
  }
  ME_HD T squaredNorm() const {
    const Matrix& a = *this;
    return redux_tree<T, R * C, 0, R * C>::run([&](int k) { return a.d[k] * a.d[k]; });
  }
  ME_HD T norm() const {
    using std::sqrt;
    return sqrt(squaredNorm());
  }
  ME_HD T dot(const Matrix& o) const {
    const Matrix& a = *this;
    return redux_tree<T, R * C, 0, R * C>::run([&](int k) { return a.d[k] * o.d[k]; });
  }
  ME_HD Matrix cross(const Matrix& o) const {
    static_assert(R * C == 3, "cross() needs 3-vectors");
    Matrix r;
    r.d[0] = d[1] * o.d[2] - d[2] * o.d[1];
    r.d[1] = d[2] * o.d[0] - d[0] * o.d[2];
    r.d[2] = d[0] * o.d[1] - d[1] * o.d[0];
    return r;
  }
  template <typename U>
  ME_HD Matrix<U, R, C> cast() const {
    Matrix<U, R, C> r;
    for (int i = 0; i < R * C; i++) r.d[i] = (U)d[i];
    return r;
  }
};

// scalar on the left (float or double literal): converted to the matrix' scalar, as Eigen does
template <typename T, int R, int C>
ME_HD Matrix<T, R, C> operator*(double s, const Matrix<T, R, C>& m) { return m * s; }
template <typename T, int R, int C>
ME_HD Matrix<T, R, C> operator*(float s, const Matrix<T, R, C>& m) { return m * s; }
template <typename T, int R, int C>
ME_HD Matrix<T, R, C> operator*(int s, const Matrix<T, R, C>& m) { return m * s; }

typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<float, 4, 1> Vector4f;
typedef Matrix<float, 1, 3> Vector3f_row;  // CvoState.cuh:12 of the reference

}  // namespace Eigen
