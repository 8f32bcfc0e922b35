// reference_api_stub.hpp — TEST-ONLY stand-ins for the third-party and reference types that
// shim/CvoGPU_b200.cpp and shim/IRLS_State_GPU_b200.cpp touch (Eigen, PCL, cvo::CvoPointCloud/CvoParams/Association/CvoGPU).
// The build container has neither Eigen nor PCL, so tests/test_shim_syntax.py compiles the shim
// against these declarations (-DCVO_SHIM_SYNTAX_CHECK -include this file) to keep it
// syntactically and type-wise honest, and tests/test_shim_runtime_gpu.py RUNS the shim against them
// on a B200 (the containers hold real data, column-major like Eigen's).  Nothing here is shipped; in
// a real build the shim includes the reference's own headers instead.
#pragma once
#include <cstddef>
#include <cstdint>
#include <list>
#include <memory>
#include <string>
#include <vector>

#include "cvo_b200.h"

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#ifndef NUM_CLASSES
#define NUM_CLASSES 19
#endif
#ifndef FEATURE_DIMENSIONS
#define FEATURE_DIMENSIONS 5
#endif

namespace Eigen {
constexpr int Dynamic = -1;
constexpr int RowMajor = 1;
// column-major storage like Eigen's default; only what the shim and the runtime driver touch
template <class T, int R, int C>
struct Matrix {
  std::vector<T> d;
  int r = R > 0 ? R : 0, c = C > 0 ? C : 0;
  Matrix() : d((size_t)(R > 0 ? R : 0) * (C > 0 ? C : 0)) {}
  Matrix(int rows, int cols) : d((size_t)rows * cols), r(rows), c(cols) {}
  explicit Matrix(int rows) : d((size_t)rows * (C > 0 ? C : 1)), r(rows), c(C > 0 ? C : 1) {}
  void resize(int rows, int cols) {
    r = rows;
    c = cols;
    d.assign((size_t)rows * cols, T(0));
  }
  long rows() const { return r; }
  long cols() const { return c; }
  long size() const { return (long)d.size(); }
  T* data() { return d.data(); }
  const T* data() const { return d.data(); }
  T& operator()(int i, int j) { return d[(size_t)j * r + i]; }
  const T& operator()(int i, int j) const { return d[(size_t)j * r + i]; }
  T& operator()(int i) { return d[i]; }
  const T& operator()(int i) const { return d[i]; }
  static Matrix Identity() {
    Matrix m;
    for (int i = 0; i < m.r && i < m.c; i++) m(i, i) = T(1);
    return m;
  }
};
using Matrix4f = Matrix<float, 4, 4>;
using Matrix3f = Matrix<float, 3, 3>;
using Vector3f = Matrix<float, 3, 1>;
using VectorXf = Matrix<float, Dynamic, 1>;
using MatrixXf = Matrix<float, Dynamic, Dynamic>;
template <class M>
struct Ref {
  M* m;
  Ref(M& x) : m(&x) {}
  Ref& operator=(const M& x) { *m = x; return *this; }
};
template <class T>
struct aligned_allocator : std::allocator<T> {
  template <class U> struct rebind { using other = aligned_allocator<U>; };
};
template <class T>
struct Triplet {
  int r, c;
  T v;
  Triplet(int r_, int c_, T v_) : r(r_), c(c_), v(v_) {}
  int row() const { return r; }
  int col() const { return c; }
  T value() const { return v; }
};
template <class T, int Opt>
struct SparseMatrix {
  int r = 0, c = 0;
  std::vector<Triplet<T>> t;  // row-major order of insertion is what the shim produces
  void resize(int rows, int cols) { r = rows; c = cols; t.clear(); }
  template <class It> void setFromTriplets(It b, It e) { t.assign(b, e); }
  void makeCompressed() {}
  long rows() const { return r; }
  long cols() const { return c; }
  long nonZeros() const { return (long)t.size(); }
};
}  // namespace Eigen

namespace pcl {
template <class P>
struct PointCloud {
  std::vector<P> points;
  size_t size() const { return points.size(); }
  const P& operator[](size_t i) const { return points[i]; }
};
}  // namespace pcl

namespace cvo {
struct CvoPoint {
  float x, y, z;
  float features[FEATURE_DIMENSIONS];
  float label_distribution[NUM_CLASSES];
  float geometric_type[2];
};
struct CvoParams : cvo_b200_params {};
struct Association {
  std::vector<int> source_inliers, target_inliers;
  Eigen::SparseMatrix<float, Eigen::RowMajor> pairs;
};
class CvoPointCloud {
 public:
  CvoPointCloud() {}
  CvoPointCloud(int feature_dimensions, int num_classes);
  ~CvoPointCloud() {}
  int num_points() const { return n_; }
  int size() const { return n_; }
  int num_classes() const { return nc_; }
  int num_features() const { return nf_; }
  const std::vector<Eigen::Vector3f, Eigen::aligned_allocator<Eigen::Vector3f>>& positions() const { return p_; }
  const Eigen::Matrix<float, Eigen::Dynamic, Eigen::Dynamic>& labels() const { return l_; }
  const Eigen::MatrixXf& features() const { return f_; }
  const std::vector<float>& geometric_types() const { return g_; }
  // the two fillers of utils/CvoPointCloud.hpp:172-173 (stand-ins for CvoPointCloud.cpp:1383-1420,
  // defined below the class): reserve sizes every container (geometric types to 2n ZEROS),
  // add_point writes one point
  void reserve(int num_points, int feature_dims, int num_classes);
  int add_point(int index, const Eigen::Vector3f& xyz, const Eigen::VectorXf& feature, const Eigen::VectorXf& label,
                const Eigen::VectorXf& geometric_type);

 private:
  int n_ = 0, nc_ = 0, nf_ = 0;
  std::vector<Eigen::Vector3f, Eigen::aligned_allocator<Eigen::Vector3f>> p_;
  Eigen::MatrixXf f_, l_;
  std::vector<float> g_;
};
inline CvoPointCloud::CvoPointCloud(int feature_dimensions, int num_classes)
    : nc_(num_classes), nf_(feature_dimensions) {}
inline void CvoPointCloud::reserve(int num_points, int feature_dims, int num_classes) {
  n_ = num_points;
  nf_ = feature_dims;
  nc_ = num_classes;
  p_.resize((size_t)n_);
  if (nf_) f_.resize(n_, nf_);
  if (nc_) l_.resize(n_, nc_);
  g_.assign((size_t)n_ * 2, 0.f);
}
inline int CvoPointCloud::add_point(int index, const Eigen::Vector3f& xyz, const Eigen::VectorXf& feature,
                                    const Eigen::VectorXf& label, const Eigen::VectorXf& geometric_type) {
  if (index >= n_ || geometric_type.size() != 2) return -1;
  p_[(size_t)index] = xyz;
  for (int j = 0; j < nf_; j++) f_(index, j) = feature(j);
  for (int j = 0; j < nc_; j++) l_(index, j) = label(j);
  g_[(size_t)index * 2] = geometric_type(0);
  g_[(size_t)index * 2 + 1] = geometric_type(1);
  return 0;
}
// ---- multi-frame types (cvo/CvoFrame.hpp, cvo/CvoFrameGPU.hpp, cvo/SparseKernelMat.hpp,
//      cvo/IRLS_State.hpp, cvo/IRLS_State_GPU.hpp), members in the reference's order
struct CvoFrame {
  typedef std::shared_ptr<CvoFrame> Ptr;
  CvoFrame(const CvoPointCloud* pts, const double poses[12]);  // reference's CvoFrame.cpp
  virtual ~CvoFrame() {}
  const CvoPointCloud* points;
  double pose_vec[12];
  virtual void transform_pointcloud();
};
class CvoFrameGPU_Impl;
struct CvoFrameGPU : public CvoFrame {
  CvoFrameGPU(const CvoPointCloud* pts, const double poses[12]);
  ~CvoFrameGPU();
  void transform_pointcloud();
  const CvoPoint* points_transformed_gpu() const;
  const float* pose_vec_gpu() const;

 private:
  std::unique_ptr<CvoFrameGPU_Impl> impl;
};
struct SparseKernelMat {
  int rows;
  int cols;
  unsigned int nonzero_sum;
  float* mat;
  int* ind_row2col;
  unsigned int* nonzeros;
};
void clear_SparseKernelMat_cpu(SparseKernelMat* A_cpu, int num_neighbors);
void init_internal_SparseKernelMat_cpu(int rows, int cols, SparseKernelMat* A_cpu);
void delete_internal_SparseKernelMat_cpu(SparseKernelMat* A_cpu);
}  // namespace cvo
namespace ceres {
class Problem;
}
namespace cvo {
class BinaryState {
 public:
  typedef std::shared_ptr<BinaryState> Ptr;
  virtual int update_inner_product() = 0;
  virtual void add_residual_to_problem(ceres::Problem& problem) = 0;
  virtual void update_ell() = 0;
};
class BinaryStateGPU : public BinaryState {
 public:
  typedef std::shared_ptr<BinaryStateGPU> Ptr;
  BinaryStateGPU(std::shared_ptr<CvoFrameGPU> pc1, std::shared_ptr<CvoFrameGPU> pc2,
                 const CvoParams* params_cpu, const CvoParams* params_gpu, unsigned int num_neighbor,
                 float init_ell);
  ~BinaryStateGPU();
  virtual int update_inner_product();
  void update_ell();                                        // reference's IRLS_State_GPU.cpp
  void add_residual_to_problem(ceres::Problem& problem);    // reference's IRLS_State_GPU.cpp

 private:
  std::shared_ptr<CvoFrameGPU> frame1_;
  std::shared_ptr<CvoFrameGPU> frame2_;
  CvoFrame* frame1();
  CvoFrame* frame2();
  unsigned int num_neighbors_;
  SparseKernelMat A_host_;
  SparseKernelMat* A_device_;
  SparseKernelMat A_result_cpu_;
  float ell_;
  int iter_;
  const unsigned int init_num_neighbors_;
  const CvoParams* params_gpu_;
  const CvoParams* params_cpu_;
};
class BinaryStateCPU : public BinaryState {  // cvo/IRLS_State_CPU.hpp:23-26 (reference's own source)
 public:
  typedef std::shared_ptr<BinaryStateCPU> Ptr;
  BinaryStateCPU(std::shared_ptr<CvoFrame> pc1, std::shared_ptr<CvoFrame> pc2, const CvoParams* params);
  virtual int update_inner_product();
  virtual void add_residual_to_problem(ceres::Problem& problem);
  virtual void update_ell();
};
class CvoBatchIRLS {  // cvo/IRLS.hpp:23-35 (reference's own source, Ceres)
 public:
  CvoBatchIRLS(const std::vector<std::shared_ptr<CvoFrame>>& frames, const std::vector<bool>& pivot_flags,
               const std::list<std::shared_ptr<BinaryState>>& states, const CvoParams* params);
  void solve();
};
// the members of cvo::CvoGPU that the shim defines (signatures as in the reference header)
class CvoGPU {
 private:
  CvoParams* params_gpu;
  CvoParams params;

 public:
  CvoGPU(const std::string& f);
  ~CvoGPU();
  CvoParams& get_params() { return params; }
  const CvoParams* get_params_gpu() { return params_gpu; }
  void write_params(const CvoParams* p_cpu);
  int align(const CvoPointCloud&, const CvoPointCloud&, const Eigen::Matrix4f&, Eigen::Ref<Eigen::Matrix4f>,
            Association* = nullptr, double* = nullptr) const;
  int align(const pcl::PointCloud<CvoPoint>&, const pcl::PointCloud<CvoPoint>&, const Eigen::Matrix4f&,
            Eigen::Ref<Eigen::Matrix4f>, Association* = nullptr, double* = nullptr) const;
  int align(std::vector<std::shared_ptr<CvoFrame>>& frames, const std::vector<bool>& frames_to_hold_const,
            const std::list<std::pair<std::shared_ptr<CvoFrame>, std::shared_ptr<CvoFrame>>>& edges,
            double* registration_seconds = nullptr) const;
  int align(std::vector<std::shared_ptr<CvoFrame>>& frames, const std::vector<bool>& frames_to_hold_const,
            const std::list<std::shared_ptr<BinaryState>>& edge_states,
            double* registration_seconds = nullptr) const;  // reference's CvoGPU.cpp:261
  float function_angle(const CvoPointCloud&, const CvoPointCloud&, const Eigen::Matrix4f&, float,
                       bool is_approximate = true, bool is_gpu = true) const;
  float function_angle(const pcl::PointCloud<CvoPoint>&, const pcl::PointCloud<CvoPoint>&,
                       const Eigen::Matrix4f&, float, bool is_approximate = true) const;
  void compute_association_gpu(const CvoPointCloud&, const CvoPointCloud&, const Eigen::Matrix4f&, float,
                               Association&) const;
  void compute_association_gpu(const CvoPointCloud&, const CvoPointCloud&, const Eigen::Matrix4f&,
                               const Eigen::Matrix3f&, Association&) const;
  float inner_product_gpu(const CvoPointCloud&, const CvoPointCloud&, const Eigen::Matrix4f&, float) const;
  float inner_product_gpu(const pcl::PointCloud<CvoPoint>&, const pcl::PointCloud<CvoPoint>&,
                          const Eigen::Matrix4f&, float) const;
  float inner_product_cpu(const CvoPointCloud&, const CvoPointCloud&, const Eigen::Matrix4f&, float) const;
};
}  // namespace cvo
