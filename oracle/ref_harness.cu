// ref_harness.cu — OUR harness around the reference's own kernels (oracle/make_ref.py extracts
// them, verbatim, into the git-ignored oracle/_ref/*.inc).  TEST INFRASTRUCTURE ONLY: nothing
// under unified_cvo_b200/, shim/, examples/ or include/ may link or load what this builds.
//
// Built three times (see make_ref.py): as host C++ (CVO_REF_HOST_BUILD: __global__ becomes a
// plain function and the grid a host loop) and twice by nvcc for sm_100a (reference flags, and
// --fmad=false).  The prelude below stands in for what the extracted text expects from headers
// that are not in this image: the PCL point macros (PCL_ADD_POINT4D / PCL_ADD_RGB — layout
// checked by static_assert against the offsets SURVEY.md §2 row 8 probed), `using namespace
// std` + <cmath> as CvoGPU.cu:45-46 has them, and Eigen's fixed-size types (ref_mini_eigen.h,
// tier 2 only).
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <vector>

#include "../include/cvo_b200.h"
#include "ref_mini_eigen.h"

#ifdef CVO_REF_HOST_BUILD
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __align__(n) alignas(n)
struct RefDim3 { unsigned x, y, z; };
static thread_local RefDim3 blockIdx, blockDim, threadIdx;
#else
#include <cuda_runtime.h>
#endif

using namespace std;  // CvoGPU.cu:46

// ---- PCL stand-ins (pcl/impl/point_types.hpp: PCL_ADD_POINT4D, PCL_ADD_RGB)
#define PCL_ADD_POINT4D \
  union alignas(16) {   \
    float data[4];      \
    struct {            \
      float x, y, z;    \
    };                  \
  };
#define PCL_ADD_RGB         \
  union {                   \
    union {                 \
      struct {              \
        uint8_t b, g, r, a; \
      };                    \
      float rgb;            \
    };                      \
    uint32_t rgba;          \
  };

namespace pcl {
#include "PointSegmentedDistribution.inc"
}  // namespace pcl

namespace cvo {
typedef pcl::PointSegmentedDistribution<FEATURE_DIMENSIONS, NUM_CLASSES> CvoPoint;  // utils/CvoPoint.hpp:9
static_assert(sizeof(CvoPoint) == 192 && alignof(CvoPoint) == 16, "CvoPoint layout");
static_assert(offsetof(CvoPoint, features) == 20 && offsetof(CvoPoint, label) == 40 &&
                  offsetof(CvoPoint, label_distribution) == 44 &&
                  offsetof(CvoPoint, geometric_type) == 44 + 4 * NUM_CLASSES,
              "CvoPoint offsets");

#include "CvoParams.inc"
#include "SparseKernelMat.inc"
static_assert(sizeof(CvoParams) == sizeof(cvo_b200_params), "cvo_b200_params mirrors CvoParams");

// ---- tier 1: no third-party arithmetic
#include "dot.inc"
#include "squared_dist_arr.inc"
#include "squared_dist_pt.inc"
#include "square_norm.inc"
#include "compute_range_ell.inc"
#include "compute_geometric_type_ip.inc"
#include "fill_in_A_mat_gpu.inc"

// ---- tier 2: Eigen primitives from ref_mini_eigen.h
#include "skew_gpu.inc"
#include "mahananobis_distance.inc"
#include "fill_in_A_mat_gpu_dense_mat_kernel.inc"
#include "compute_flow_gpu_no_eigen.inc"
#include "compute_step_size_xi.inc"
#include "compute_step_size_poly_coeff.inc"
#ifdef CVO_REF_HOST_BUILD
// ---- tier 1, host code of the reference: the indicator queues of align_impl (std::queue only)
#include "A_sparsity_indicator_ell_update.inc"
#endif
#ifdef CVO_REF_HOST_BUILD
}  // namespace cvo
namespace thrust {
template <typename Arg, typename Res>
struct unary_function {};
}  // namespace thrust
namespace cvo {
// ---- tier 2: the inverse pose (update_tf, a3) and the point transform (transform_point_R_T, a4)
typedef Eigen::Matrix<float, 3, 3> Mat33f;  // utils/data_type.hpp:95
typedef Eigen::Matrix<float, 3, 1> Vec3f;   // :98
typedef Eigen::Matrix<float, 4, 4> Mat44f;  // :137
struct CvoState {  // the two members update_tf writes (cvo/CvoState.cuh:50-51)
  Eigen::Matrix3f* R_gpu;
  Eigen::Vector3f* T_gpu;
};
enum { cudaMemcpyHostToDevice = 1 };
static inline int cudaMemcpy(void* dst, const void* src, size_t n, int) {
  memcpy(dst, src, n);
  return 0;
}
#include "update_tf.inc"
#include "transform_point_R_T.inc"
#include "transform_point_pose_vec.inc"
#endif
#ifdef CVO_REF_HOST_BUILD
// ---- tier 2, host code of the reference: the pose increment of align_impl (CvoGPU.cu:1462)
const float TOLERANCE = 1e-6;  // LieGroup.cpp:9
#include "skew.inc"
#include "Exp_SEK3.inc"
#endif
}  // namespace cvo

using cvo::CvoParams;
using cvo::CvoPoint;
using cvo::SparseKernelMat;

// ------------------------------------------------------------------------------------------
// packing: the callers hand SoA arrays (xyz[n*3], feat[n*F], lab[n*C], geo[n*2]); F <= 5,
// C <= NUM_CLASSES; missing channels stay zero as the reference's ctor leaves them
static void pack_points(std::vector<CvoPoint>& out, int n, const float* xyz, const float* feat, int F,
                        const float* lab, int C, const float* geo) {
  out.resize((size_t)n);
  for (int i = 0; i < n; i++) {
    CvoPoint& p = out[(size_t)i];
    p.x = xyz[3 * i];
    p.y = xyz[3 * i + 1];
    p.z = xyz[3 * i + 2];
    for (int f = 0; f < F && f < FEATURE_DIMENSIONS; f++) p.features[f] = feat[(size_t)i * F + f];
    for (int c = 0; c < C && c < NUM_CLASSES; c++) p.label_distribution[c] = lab[(size_t)i * C + c];
    if (geo) {
      p.geometric_type[0] = geo[2 * i];
      p.geometric_type[1] = geo[2 * i + 1];
    }
  }
}

#ifdef CVO_REF_HOST_BUILD
// ---- host "device": plain memory, the grid is a loop
template <typename T>
static T* dev_alloc(size_t n) { return (T*)malloc(n ? n * sizeof(T) : 1); }
template <typename T>
static T* dev_upload(const T* h, size_t n) {
  T* d = dev_alloc<T>(n);
  if (n) memcpy(d, h, n * sizeof(T));
  return d;
}
template <typename T>
static void dev_download(T* h, const T* d, size_t n) { if (n) memcpy(h, d, n * sizeof(T)); }
static void dev_set(void* d, int byte, size_t bytes) { if (bytes) memset(d, byte, bytes); }
static void dev_free(void* d) { free(d); }
static int dev_sync() { return 0; }
#define REF_LAUNCH(kernel, nthreads, ...)                                         \
  do {                                                                            \
    const long long n_blocks_ = (nthreads) / CUDA_BLOCK_SIZE + 1;                 \
    _Pragma("omp parallel for schedule(dynamic, 64)")                             \
    for (long long t_ = 0; t_ < n_blocks_ * CUDA_BLOCK_SIZE; t_++) {              \
      blockDim.x = CUDA_BLOCK_SIZE;                                               \
      blockIdx.x = (unsigned)(t_ / CUDA_BLOCK_SIZE);                              \
      threadIdx.x = (unsigned)(t_ % CUDA_BLOCK_SIZE);                             \
      kernel(__VA_ARGS__);                                                        \
    }                                                                             \
  } while (0)
#else
template <typename T>
static T* dev_alloc(size_t n) {
  T* d = nullptr;
  if (cudaMalloc((void**)&d, n ? n * sizeof(T) : 1) != cudaSuccess) return nullptr;
  return d;
}
template <typename T>
static T* dev_upload(const T* h, size_t n) {
  T* d = dev_alloc<T>(n);
  if (d && n) cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyHostToDevice);
  return d;
}
template <typename T>
static void dev_download(T* h, const T* d, size_t n) {
  if (n) cudaMemcpy(h, d, n * sizeof(T), cudaMemcpyDeviceToHost);
}
static void dev_set(void* d, int byte, size_t bytes) { if (bytes) cudaMemset(d, byte, bytes); }
static void dev_free(void* d) { cudaFree(d); }
static int dev_sync() {
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    fprintf(stderr, "cvo_ref: CUDA error %s\n", cudaGetErrorString(e));
    return -1;
  }
  return 0;
}
// the reference's launch shape: <<<n / CUDA_BLOCK_SIZE + 1, CUDA_BLOCK_SIZE>>> (CvoGPU.cu:665)
#define REF_LAUNCH(kernel, nthreads, ...) \
  kernel<<<(nthreads) / CUDA_BLOCK_SIZE + 1, CUDA_BLOCK_SIZE>>>(__VA_ARGS__)
#endif

namespace {
// the device-side SparseKernelMat as init_SparseKernelMat_gpu / clear_SparseKernelMat leave it
// (SparseKernelMat.cu:90-122): mat zeroed, ind_row2col all -1, nonzeros zeroed
struct DevA {
  SparseKernelMat host;  // device pointers inside
  SparseKernelMat* dev;
  DevA(int rows, int cols) {
    host.rows = rows;
    host.cols = cols;
    host.nonzero_sum = 0;
    const size_t n = (size_t)rows * (size_t)cols;
    host.mat = dev_alloc<float>(n);
    host.ind_row2col = dev_alloc<int>(n);
    host.nonzeros = dev_alloc<unsigned int>((size_t)rows);
    dev_set(host.mat, 0, n * sizeof(float));
    dev_set(host.ind_row2col, 0xFF, n * sizeof(int));
    dev_set(host.nonzeros, 0, (size_t)rows * sizeof(unsigned int));
    dev = dev_upload(&host, 1);
  }
  void fill_from(const float* mat, const int* ind, const unsigned int* nz) {
    const size_t n = (size_t)host.rows * (size_t)host.cols;
#ifdef CVO_REF_HOST_BUILD
    memcpy(host.mat, mat, n * sizeof(float));
    memcpy(host.ind_row2col, ind, n * sizeof(int));
    if (nz) memcpy(host.nonzeros, nz, (size_t)host.rows * sizeof(unsigned int));
#else
    cudaMemcpy(host.mat, mat, n * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(host.ind_row2col, ind, n * sizeof(int), cudaMemcpyHostToDevice);
    if (nz) cudaMemcpy(host.nonzeros, nz, (size_t)host.rows * sizeof(unsigned int), cudaMemcpyHostToDevice);
#endif
  }
  void read(float* mat, int* ind, unsigned int* nz) {
    const size_t n = (size_t)host.rows * (size_t)host.cols;
    dev_download(mat, host.mat, n);
    dev_download(ind, host.ind_row2col, n);
    dev_download(nz, host.nonzeros, (size_t)host.rows);
  }
  ~DevA() {
    dev_free(host.mat);
    dev_free(host.ind_row2col);
    dev_free(host.nonzeros);
    dev_free(dev);
  }
};

struct Cloud {
  CvoPoint* dev = nullptr;
  int n = 0;
  Cloud(int n_, const float* xyz, const float* feat, int F, const float* lab, int C, const float* geo) : n(n_) {
    std::vector<CvoPoint> h;
    pack_points(h, n, xyz, feat, F, lab, C, geo);
    dev = dev_upload(h.data(), (size_t)n);
  }
  ~Cloud() { dev_free(dev); }
};
}  // namespace

extern "C" {

// 1 = host build, 2 = nvcc with the reference's flags, 3 = nvcc --fmad=false
int cvo_ref_build_kind(void) {
#ifdef CVO_REF_HOST_BUILD
  return 1;
#elif defined(CVO_REF_NOFMA)
  return 3;
#else
  return 2;
#endif
}
#ifdef CVO_REF_HOST_BUILD
// The reference's A_sparsity_indicator_ell_update (CvoGPU.cu:1167-1285) over a sequence of
// indicators with the state align_impl gives it (:1377-1380: both queues empty, both sums 0).
int cvo_ref_indicator_sequence(const cvo_b200_params* params, int n, const float* indicators, int* decrease,
                               float* start_sums, float* end_sums) {
  CvoParams p;
  memcpy(&p, params, sizeof(p));
  std::queue<float> indicator_start_queue, indicator_end_queue;
  float indicator_start_sum = 0, indicator_end_sum = 0;
  for (int k = 0; k < n; k++) {
    decrease[k] = cvo::A_sparsity_indicator_ell_update(indicator_start_queue, indicator_end_queue, indicator_start_sum,
                                                       indicator_end_sum, indicators[k], p)
                      ? 1
                      : 0;
    start_sums[k] = indicator_start_sum;
    end_sums[k] = indicator_end_sum;
  }
  return 0;
}
#endif
#ifdef CVO_REF_HOST_BUILD
// The reference's update_tf (CvoGPU.cu:94-112) followed by its point transform
// (transform_point_R_T, CvoGPU_impl.cu:31-82, as transform_pointcloud_thrust applies it at
// CvoGPU.cu:1404 with update_normal_and_cov = false): R, out_Rinv column-major 3x3, transform16
// column-major 4x4, xyz / out n x 3.
int cvo_ref_update_tf_and_transform(const float R[9], const float T[3], float out_Rinv[9], float out_Tinv[3],
                                    float transform16[16], int n, const float* xyz, float* out) {
  cvo::Mat33f Rm, Rg;
  cvo::Vec3f Tm, Tg;
  for (int i = 0; i < 9; i++) Rm.d[i] = R[i];
  for (int i = 0; i < 3; i++) Tm.d[i] = T[i];
  cvo::CvoState st{&Rg, &Tg};
  cvo::Mat44f tf = cvo::Mat44f::Zero();
  cvo::update_tf(Rm, Tm, &st, tf);
  for (int i = 0; i < 9; i++) out_Rinv[i] = Rg.d[i];
  for (int i = 0; i < 3; i++) out_Tinv[i] = Tg.d[i];
  for (int i = 0; i < 16; i++) transform16[i] = tf.d[i];
  cvo::transform_point_R_T f(&Rg, &Tg, false);
  for (int j = 0; j < n; j++) {
    CvoPoint p;
    p.x = xyz[3 * j];
    p.y = xyz[3 * j + 1];
    p.z = xyz[3 * j + 2];
    const CvoPoint q = f(p);
    out[3 * j] = q.x;
    out[3 * j + 1] = q.y;
    out[3 * j + 2] = q.z;
  }
  return 0;
}
#endif
#ifdef CVO_REF_HOST_BUILD
// The reference's transform_point_pose_vec (CvoGPU_impl.cu:84-150; CvoFrameGPU::transform_pointcloud
// applies it with the frame's row-major 3x4 pose): xyz / out n x 3.
int cvo_ref_transform_pose_vec(const float pose12[12], int n, const float* xyz, float* out) {
  float pose[12];
  memcpy(pose, pose12, sizeof(pose));
  cvo::transform_point_pose_vec f(pose, false);
  for (int j = 0; j < n; j++) {
    CvoPoint p;
    p.x = xyz[3 * j];
    p.y = xyz[3 * j + 1];
    p.z = xyz[3 * j + 2];
    const CvoPoint q = f(p);
    out[3 * j] = q.x;
    out[3 * j + 1] = q.y;
    out[3 * j + 2] = q.z;
  }
  return 0;
}
#endif
#ifdef CVO_REF_HOST_BUILD
// The reference's Exp_SEK3(v, dt) (LieGroup.cpp:245-274): out12 = the 3x4 result, column-major
// (R's columns, then the translation).
int cvo_ref_exp_sek3(const float xi[6], float dt, float out12[12]) {
  Eigen::Matrix<float, 6, 1> v;
  for (int i = 0; i < 6; i++) v[i] = xi[i];
  const Eigen::Matrix<float, 3, 4> X = cvo::Exp_SEK3(v, dt);
  for (int i = 0; i < 12; i++) out12[i] = X.d[i];
  return 0;
}
#endif
int cvo_ref_num_classes(void) { return NUM_CLASSES; }
int cvo_ref_feature_dimensions(void) { return FEATURE_DIMENSIONS; }
int cvo_ref_sizeof_point(void) { return (int)sizeof(CvoPoint); }

// K1: the reference's fill_in_A_mat_gpu launched as se_kernel launches it (CvoGPU.cu:665-673)
// on the points given (cloud b = the ALREADY MOVED target).  Outputs are the three arrays of the
// SparseKernelMat with stride num_neighbors: mat, ind_row2col (-1 = unused), nonzeros.
int cvo_ref_fill_A(const cvo_b200_params* params, int n_a, const float* xyz_a, const float* feat_a,
                   const float* lab_a, const float* geo_a, int n_b, const float* xyz_b,
                   const float* feat_b, const float* lab_b, const float* geo_b, int F, int C,
                   int num_neighbors, float ell, float* mat, int* ind, unsigned int* nonzeros) {
  if (n_a <= 0 || num_neighbors < 0) return -2;
  CvoParams p;
  memcpy(&p, params, sizeof(p));
  CvoParams* p_dev = dev_upload(&p, 1);
  Cloud a(n_a, xyz_a, feat_a, F, lab_a, C, geo_a), b(n_b, xyz_b, feat_b, F, lab_b, C, geo_b);
  DevA A(n_a, num_neighbors > 0 ? num_neighbors : 1);
  REF_LAUNCH(cvo::fill_in_A_mat_gpu, n_a, p_dev, a.dev, n_a, b.dev, n_b, num_neighbors, ell, A.dev);
  const int rc = dev_sync();
  if (num_neighbors > 0) A.read(mat, ind, nonzeros);
  else dev_download(nonzeros, A.host.nonzeros, (size_t)n_a);
  dev_free(p_dev);
  return rc;
}

// K1b: fill_in_A_mat_gpu_dense_mat_kernel (CvoGPU.cu:217-327); kernel_inv column-major 3x3
int cvo_ref_fill_A_dense(const cvo_b200_params* params, int n_a, const float* xyz_a, const float* feat_a,
                         const float* lab_a, const float* geo_a, int n_b, const float* xyz_b,
                         const float* feat_b, const float* lab_b, const float* geo_b, int F, int C,
                         int num_neighbors, const float* kernel_inv, float* mat, int* ind,
                         unsigned int* nonzeros) {
  if (n_a <= 0 || num_neighbors <= 0) return -2;
  CvoParams p;
  memcpy(&p, params, sizeof(p));
  CvoParams* p_dev = dev_upload(&p, 1);
  Cloud a(n_a, xyz_a, feat_a, F, lab_a, C, geo_a), b(n_b, xyz_b, feat_b, F, lab_b, C, geo_b);
  Eigen::Matrix3f kh;
  for (int k = 0; k < 9; k++) kh.d[k] = kernel_inv[k];
  Eigen::Matrix3f* k_dev = dev_upload(&kh, 1);
  DevA A(n_a, num_neighbors);
  REF_LAUNCH(cvo::fill_in_A_mat_gpu_dense_mat_kernel, n_a, p_dev, a.dev, n_a, b.dev, n_b, num_neighbors,
             k_dev, A.dev);
  const int rc = dev_sync();
  A.read(mat, ind, nonzeros);
  dev_free(p_dev);
  dev_free(k_dev);
  return rc;
}

// K2: compute_flow_gpu_no_eigen (CvoGPU.cu:729-790) on a given sparse matrix: per-row omega_i / c
// and v_i / d as doubles, [n_a * 3] each
int cvo_ref_flow_rows(const cvo_b200_params* params, int n_a, const float* xyz_a, int n_b,
                      const float* xyz_b, int num_neighbors, const float* mat, const int* ind,
                      double* omega_rows, double* v_rows) {
  if (n_a <= 0 || num_neighbors <= 0) return -2;
  CvoParams p;
  memcpy(&p, params, sizeof(p));
  CvoParams* p_dev = dev_upload(&p, 1);
  Cloud a(n_a, xyz_a, nullptr, 0, nullptr, 0, nullptr), b(n_b, xyz_b, nullptr, 0, nullptr, 0, nullptr);
  DevA A(n_a, num_neighbors);
  A.fill_from(mat, ind, nullptr);
  Eigen::Vector3d* om = dev_alloc<Eigen::Vector3d>((size_t)n_a);
  Eigen::Vector3d* vv = dev_alloc<Eigen::Vector3d>((size_t)n_a);
  REF_LAUNCH(cvo::compute_flow_gpu_no_eigen, n_a, p_dev, a.dev, b.dev, A.dev, num_neighbors, om, vv);
  const int rc = dev_sync();
  static_assert(sizeof(Eigen::Vector3d) == 24, "Vector3d");
  dev_download((Eigen::Vector3d*)omega_rows, om, (size_t)n_a);
  dev_download((Eigen::Vector3d*)v_rows, vv, (size_t)n_a);
  dev_free(om);
  dev_free(vv);
  dev_free(p_dev);
  return rc;
}

// K3 + K4: compute_step_size_xi then compute_step_size_poly_coeff as compute_step_size launches
// them (CvoGPU.cu:1084-1116); per-row B, C, D, E as doubles [n_a] each
int cvo_ref_step_rows(const float* omega, const float* v, float ell, float ell_init,
                      int is_using_range_ell, int n_a, const float* xyz_a, int n_b,
                      const float* xyz_b, int num_neighbors, const float* mat, const int* ind,
                      double* B, double* C, double* D, double* E) {
  if (n_a <= 0 || n_b <= 0 || num_neighbors <= 0) return -2;
  Cloud a(n_a, xyz_a, nullptr, 0, nullptr, 0, nullptr), b(n_b, xyz_b, nullptr, 0, nullptr, 0, nullptr);
  DevA A(n_a, num_neighbors);
  A.fill_from(mat, ind, nullptr);
  Eigen::Vector3f oh(omega[0], omega[1], omega[2]), vh(v[0], v[1], v[2]);
  Eigen::Vector3f* o_dev = dev_upload(&oh, 1);
  Eigen::Vector3f* v_dev = dev_upload(&vh, 1);
  typedef Eigen::Vector3f_row Row;
  static_assert(sizeof(Row) == 12, "Vector3f_row");
  Row* xiz = dev_alloc<Row>((size_t)n_b);
  Row* xi2z = dev_alloc<Row>((size_t)n_b);
  Row* xi3z = dev_alloc<Row>((size_t)n_b);
  Row* xi4z = dev_alloc<Row>((size_t)n_b);
  float* normxiz2 = dev_alloc<float>((size_t)n_b);
  float* xiz_dot_xi2z = dev_alloc<float>((size_t)n_b);
  float* epsil_const = dev_alloc<float>((size_t)n_b);
  double* Bd = dev_alloc<double>((size_t)n_a);
  double* Cd = dev_alloc<double>((size_t)n_a);
  double* Dd = dev_alloc<double>((size_t)n_a);
  double* Ed = dev_alloc<double>((size_t)n_a);
  REF_LAUNCH(cvo::compute_step_size_xi, n_b, o_dev, v_dev, b.dev, n_b, num_neighbors, xiz, xi2z, xi3z, xi4z,
             normxiz2, xiz_dot_xi2z, epsil_const);
  REF_LAUNCH(cvo::compute_step_size_poly_coeff, n_a, ell, ell_init, is_using_range_ell, n_b, A.dev, a.dev,
             b.dev, xiz, xi2z, xi3z, xi4z, normxiz2, xiz_dot_xi2z, epsil_const, num_neighbors, Bd, Cd,
             Dd, Ed);
  const int rc = dev_sync();
  dev_download(B, Bd, (size_t)n_a);
  dev_download(C, Cd, (size_t)n_a);
  dev_download(D, Dd, (size_t)n_a);
  dev_download(E, Ed, (size_t)n_a);
  void* frees[] = {o_dev, v_dev, xiz, xi2z, xi3z, xi4z, normxiz2, xiz_dot_xi2z, epsil_const, Bd, Cd, Dd, Ed};
  for (void* f : frees) dev_free(f);
  return rc;
}

}  // extern "C"
