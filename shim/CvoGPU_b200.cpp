// CvoGPU_b200.cpp — the reference-side binding: cvo::CvoGPU forwarded to libcvo_b200.so.
//
// Compiled INSTEAD OF the reference's src/cvo/CvoGPU.cu + CvoGPU_impl.cu + CvoState.cu +
// SparseKernelMat.cu (CMakeLists.txt:176-192) into the same target `cvo_gpu_img_lib`
// (shim/CMakeLists.txt), against the reference's OWN headers
// (include/UnifiedCvo/cvo/CvoGPU.hpp:33-232, utils/CvoPointCloud.hpp:65-173,
// cvo/CvoParams.hpp:12-128, cvo/Association.hpp:7-11), so cvo_align_gpu_two_color_pcd and the
// KITTI / TUM drivers link unchanged.  It needs Eigen3, PCL and yaml-cpp-free: the YAML reader
// is inside libcvo_b200 (cvo_b200_params_read_yaml).  This translation unit cannot be compiled
// against the real headers in the build container (no Eigen / PCL there); tests/test_shim_syntax.py
// compiles it against the stand-in headers under shim/stubs/ and tests/test_shim_runtime_gpu.py
// runs it against them on a B200 (tests/shim_runtime/driver.cpp).
//
// The class layout is fixed by the reference header (CvoParams* params_gpu; CvoParams params;),
// so the device handle lives in a side table keyed by `this`.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <list>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

#ifndef CVO_SHIM_SYNTAX_CHECK
#include "cvo/CvoGPU.hpp"
#include "cvo/CvoFrameGPU.hpp"
#include "cvo/IRLS.hpp"
#include "cvo/IRLS_State.hpp"
#include "cvo/IRLS_State_CPU.hpp"
#include "cvo/IRLS_State_GPU.hpp"
#endif
#include "cvo_b200.h"
#include "shim_pack.hpp"

namespace cvo {
namespace {

static_assert(sizeof(CvoParams) == sizeof(cvo_b200_params),
              "cvo_b200_params must mirror cvo::CvoParams field for field (CvoParams.hpp:12-73)");

std::mutex g_mu;
std::unordered_map<const CvoGPU*, cvo_b200_handle*> g_handles;

cvo_b200_handle* handle_of(const CvoGPU* self) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_handles.find(self);
  return it == g_handles.end() ? nullptr : it->second;
}

[[noreturn]] void die(const cvo_b200_handle* h, const char* what, int rc) {
  // the reference's error behaviour for CUDA failures: message + exit (CvoGPU_impl.cuh:27-36)
  std::fprintf(stderr, "[cvo_b200] %s failed (%d): %s\n", what, rc,
               h ? cvo_b200_last_error(h) : cvo_b200_global_error());
  std::exit(EXIT_FAILURE);
}

using shim::Packed;
using shim::pack;

void set_clouds(cvo_b200_handle* h, const Packed& s, const Packed& t) {
  // both clouds must be packed with the same F / C (the reference compiles them in)
  int rc = cvo_b200_set_cloud(h, 0, s.n, s.xyz.data(), s.F, s.p_feat(), s.C, s.p_lab(), s.p_geo());
  if (rc != CVO_B200_OK) die(h, "cvo_b200_set_cloud(source)", rc);
  rc = cvo_b200_set_cloud(h, 1, t.n, t.xyz.data(), t.F, t.p_feat(), t.C, t.p_lab(), t.p_geo());
  if (rc != CVO_B200_OK) die(h, "cvo_b200_set_cloud(target)", rc);
}

// CSR from the library -> cvo::Association, filled like gpu_association_to_cpu
// (CvoGPU_impl.cu:366-427): source_inliers = rows with entries, target_inliers = one entry per
// stored pair, pairs = N x M row-major sparse matrix.
void export_association(cvo_b200_handle* h, int which /*0: compute, 1: last align iteration*/,
                        const float T16[16], float ell, const float* kernel3x3, int n_src,
                        int n_tgt, Association& out) {
  int64_t nnz = 0;
  std::vector<int32_t> row_ptr((size_t)n_src + 1, 0);
  int rc = which == 0
               ? cvo_b200_association(h, T16, ell, kernel3x3, &nnz, row_ptr.data(), nullptr, nullptr)
               : cvo_b200_align_association(h, &nnz, row_ptr.data(), nullptr, nullptr);
  if (rc != CVO_B200_OK) die(h, "cvo_b200_association", rc);
  if (nnz == 0) return;  // CvoGPU_impl.cu:377-378: output untouched
  std::vector<int32_t> cols((size_t)nnz);
  std::vector<float> vals((size_t)nnz);
  rc = which == 0 ? cvo_b200_association(h, T16, ell, kernel3x3, &nnz, row_ptr.data(), cols.data(),
                                         vals.data())
                  : cvo_b200_align_association(h, &nnz, row_ptr.data(), cols.data(), vals.data());
  if (rc != CVO_B200_OK) die(h, "cvo_b200_association", rc);
  out.pairs.resize(n_src, n_tgt);
  std::vector<Eigen::Triplet<float>> trip;
  trip.reserve((size_t)nnz);
  for (int i = 0; i < n_src; i++) {
    if (row_ptr[i + 1] > row_ptr[i]) out.source_inliers.push_back(i);
    for (int32_t k = row_ptr[i]; k < row_ptr[i + 1]; k++) {
      out.target_inliers.push_back(cols[k]);
      trip.emplace_back(i, cols[k], vals[k]);
    }
  }
  out.pairs.setFromTriplets(trip.begin(), trip.end());
  out.pairs.makeCompressed();
}

template <class Cloud>
int align_any(const CvoGPU* self, const CvoParams& params, int n_src, int n_tgt, const Cloud& src,
              const Cloud& tgt, const Eigen::Matrix4f& T_init, Eigen::Ref<Eigen::Matrix4f> transform,
              Association* association, double* registration_seconds) {
  if (n_src == 0 || n_tgt == 0) {  // CvoGPU.cu:1614-1617
    std::cout << "[align] point clouds inputs are empty\n";
    return 0;
  }
  cvo_b200_handle* h = handle_of(self);
  // The reference's controller reads the HOST params (mutated through get_params()), its
  // kernels the DEVICE copy of the last write_params (CvoGPU.cu:1338-1350); drivers keep the
  // two in sync by hand.  One upload of the host copy before every call gives the same
  // behaviour for every driver that does.
  int rc = cvo_b200_write_params(h, reinterpret_cast<const cvo_b200_params*>(&params));
  if (rc != CVO_B200_OK) die(h, "cvo_b200_write_params", rc);
  set_clouds(h, pack(src), pack(tgt));
  const Eigen::Matrix4f Ti = T_init;  // column-major float[16], as the C-ABI expects
  Eigen::Matrix4f To = Eigen::Matrix4f::Identity();
  cvo_b200_align_info info;
  rc = cvo_b200_align(h, Ti.data(), To.data(), &info, nullptr, 0);
  if (rc != CVO_B200_OK) die(h, "cvo_b200_align", rc);
  transform = To;
  std::cout << "cvo # of iterations is " << info.iterations << std::endl;  // CvoGPU.cu:1546
  if (params.is_exporting_association && association)  // CvoGPU.cu:1552-1556
    export_association(h, 1, nullptr, 0.f, nullptr, n_src, n_tgt, *association);
  if (registration_seconds) *registration_seconds = info.registration_seconds;
  return info.ret;
}

}  // namespace

// ---- CvoGPU.cu:64-83
CvoGPU::CvoGPU(const std::string& param_file) : params_gpu(nullptr) {
  static_assert(sizeof(params) == sizeof(cvo_b200_params), "CvoParams layout");
  cvo_b200_params p;
  cvo_b200_params_default(&p);
  int rc = cvo_b200_params_read_yaml(param_file.c_str(), &p);
  if (rc != CVO_B200_OK) die(nullptr, "cvo_b200_params_read_yaml", rc);
  std::memcpy(&params, &p, sizeof(p));
  std::printf("Some Cvo Params are: ell_init: %f, eps_2: %f\n", params.ell_init, params.eps_2);
  cvo_b200_handle* h = nullptr;
  rc = cvo_b200_create(&p, /*device=*/0, &h);  // CUDA_VISIBLE_DEVICES selects it, as upstream
  if (rc != CVO_B200_OK) die(nullptr, "cvo_b200_create", rc);
  std::lock_guard<std::mutex> lk(g_mu);
  g_handles[this] = h;
}

CvoGPU::~CvoGPU() {
  cvo_b200_handle* h = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_handles.find(this);
    if (it != g_handles.end()) {
      h = it->second;
      g_handles.erase(it);
    }
  }
  cvo_b200_destroy(h);
  shim::forget_edge_handle(&params);
}

void CvoGPU::write_params(const CvoParams* p_cpu) {
  // uploads only; the host copy is NOT replaced (CvoGPU.cu:73-77)
  cvo_b200_handle* h = handle_of(this);
  int rc = cvo_b200_write_params(h, reinterpret_cast<const cvo_b200_params*>(p_cpu));
  if (rc != CVO_B200_OK) die(h, "cvo_b200_write_params", rc);
}

// ---- CvoGPU.cu:1605-1632 and :1574-1603
int CvoGPU::align(const CvoPointCloud& source_points, const CvoPointCloud& target_points,
                  const Eigen::Matrix4f& T_target_frame_to_source_frame,
                  Eigen::Ref<Eigen::Matrix4f> transform, Association* association,
                  double* registration_seconds) const {
  return align_any(this, params, source_points.num_points(), target_points.num_points(),
                   source_points, target_points, T_target_frame_to_source_frame, transform,
                   association, registration_seconds);
}

int CvoGPU::align(const pcl::PointCloud<CvoPoint>& source_points,
                  const pcl::PointCloud<CvoPoint>& target_points,
                  const Eigen::Matrix4f& T_target_frame_to_source_frame,
                  Eigen::Ref<Eigen::Matrix4f> transform, Association* association,
                  double* registration_seconds) const {
  return align_any(this, params, (int)source_points.size(), (int)target_points.size(),
                   source_points, target_points, T_target_frame_to_source_frame, transform,
                   association, registration_seconds);
}

// ---- CvoGPU.cu:1780-1794 / 1796-1812
float CvoGPU::inner_product_gpu(const CvoPointCloud& source_points,
                                const CvoPointCloud& target_points,
                                const Eigen::Matrix4f& T_target_frame_to_source_frame,
                                float ell) const {
  if (source_points.num_points() == 0 || target_points.num_points() == 0) return 0;
  cvo_b200_handle* h = handle_of(this);
  set_clouds(h, pack(source_points), pack(target_points));
  const Eigen::Matrix4f T = T_target_frame_to_source_frame;
  float out = 0.f;
  int rc = cvo_b200_inner_product(h, T.data(), ell, &out);
  if (rc != CVO_B200_OK) die(h, "cvo_b200_inner_product", rc);
  return out;
}

float CvoGPU::inner_product_gpu(const pcl::PointCloud<CvoPoint>& source_points_pcl,
                                const pcl::PointCloud<CvoPoint>& target_points_pcl,
                                const Eigen::Matrix4f& init_guess_transform, float ell) const {
  if (source_points_pcl.size() == 0 || target_points_pcl.size() == 0) return 0;
  cvo_b200_handle* h = handle_of(this);
  set_clouds(h, pack(source_points_pcl), pack(target_points_pcl));
  const Eigen::Matrix4f T = init_guess_transform;
  float out = 0.f;
  int rc = cvo_b200_inner_product(h, T.data(), ell, &out);
  if (rc != CVO_B200_OK) die(h, "cvo_b200_inner_product", rc);
  return out;
}

// ---- CvoGPU.cu:1814-1846 / 1848-1874
float CvoGPU::function_angle(const CvoPointCloud& source_points, const CvoPointCloud& target_points,
                             const Eigen::Matrix4f& T_target_frame_to_source_frame, float ell,
                             bool is_approximate, bool is_gpu) const {
  if (source_points.num_points() == 0 || target_points.num_points() == 0) return 0;
  if (!is_gpu)  // the CPU variant stays the reference's own code (CvoGPU.cpp:96-213)
    return inner_product_cpu(source_points, target_points, T_target_frame_to_source_frame, ell) /
           (is_approximate
                ? std::sqrt((float)source_points.num_points()) * std::sqrt((float)target_points.num_points())
                : std::sqrt(inner_product_cpu(source_points, source_points, Eigen::Matrix4f::Identity(), ell)) *
                      std::sqrt(inner_product_cpu(target_points, target_points, Eigen::Matrix4f::Identity(), ell)));
  cvo_b200_handle* h = handle_of(this);
  set_clouds(h, pack(source_points), pack(target_points));
  const Eigen::Matrix4f T = T_target_frame_to_source_frame;
  float out = 0.f;
  int rc = cvo_b200_function_angle(h, T.data(), ell, is_approximate ? 1 : 0, &out);
  if (rc != CVO_B200_OK) die(h, "cvo_b200_function_angle", rc);
  return out;
}

float CvoGPU::function_angle(const pcl::PointCloud<CvoPoint>& source_points,
                             const pcl::PointCloud<CvoPoint>& target_points,
                             const Eigen::Matrix4f& T_target_frame_to_source_frame, float ell,
                             bool is_approximate) const {
  if (source_points.size() == 0 || target_points.size() == 0) return 0;
  cvo_b200_handle* h = handle_of(this);
  set_clouds(h, pack(source_points), pack(target_points));
  const Eigen::Matrix4f T = T_target_frame_to_source_frame;
  float out = 0.f;
  int rc = cvo_b200_function_angle(h, T.data(), ell, is_approximate ? 1 : 0, &out);
  if (rc != CVO_B200_OK) die(h, "cvo_b200_function_angle", rc);
  return out;
}

// ---- CvoGPU.cu:1876-1911 and :1975-1995
void CvoGPU::compute_association_gpu(const CvoPointCloud& source_points,
                                     const CvoPointCloud& target_points,
                                     const Eigen::Matrix4f& T_target_frame_to_source_frame,
                                     float lengthscale, Association& association) const {
  if (source_points.num_points() == 0 || target_points.num_points() == 0) return;
  cvo_b200_handle* h = handle_of(this);
  set_clouds(h, pack(source_points), pack(target_points));
  const Eigen::Matrix4f T = T_target_frame_to_source_frame;
  export_association(h, 0, T.data(), lengthscale, nullptr, source_points.num_points(),
                     target_points.num_points(), association);
}

void CvoGPU::compute_association_gpu(const CvoPointCloud& source_points,
                                     const CvoPointCloud& target_points,
                                     const Eigen::Matrix4f& T_target_frame_to_source_frame,
                                     const Eigen::Matrix3f& non_isotropic_kernel,
                                     Association& association) const {
  if (source_points.num_points() == 0 || target_points.num_points() == 0) return;
  cvo_b200_handle* h = handle_of(this);
  set_clouds(h, pack(source_points), pack(target_points));
  const Eigen::Matrix4f T = T_target_frame_to_source_frame;
  const Eigen::Matrix3f K = non_isotropic_kernel;  // column-major float[9]
  export_association(h, 0, T.data(), 0.f, K.data(), source_points.num_points(),
                     target_points.num_points(), association);
}

// ---- CvoGPU.cu:1637-1686: multi-frame registration from a list of frame pairs.  This overload
// is DEFINED IN THE REPLACED CvoGPU.cu (the one taking ready-made BinaryStates lives in the kept
// CvoGPU.cpp:261), and all main_multi_frame_irls_* / covisMap drivers call it: one edge state per
// pair - CPU (kd-tree) or GPU (this library: shim/IRLS_State_GPU_b200.cpp) as
// params.multiframe_using_cpu says - then the reference's own CvoBatchIRLS (IRLS.cpp, kept).
int CvoGPU::align(std::vector<CvoFrame::Ptr>& frames, const std::vector<bool>& frames_to_hold_const,
                  const std::list<std::pair<CvoFrame::Ptr, CvoFrame::Ptr>>& edges,
                  double* registration_seconds) const {
  auto start = std::chrono::system_clock::now();
  std::list<BinaryState::Ptr> binary_states;
  for (auto&& e : edges) {
    const CvoFrame::Ptr& f1 = e.first;
    const CvoFrame::Ptr& f2 = e.second;
    if (params.multiframe_using_cpu) {
      BinaryStateCPU::Ptr st(new BinaryStateCPU(f1, f2, &params));
      binary_states.push_back(std::dynamic_pointer_cast<BinaryState>(st));
    } else {
      BinaryStateGPU::Ptr st(new BinaryStateGPU(std::dynamic_pointer_cast<CvoFrameGPU>(f1),
                                                std::dynamic_pointer_cast<CvoFrameGPU>(f2), &params,
                                                params_gpu, params.multiframe_num_neighbors,
                                                params.multiframe_ell_init));
      binary_states.push_back(std::dynamic_pointer_cast<BinaryState>(st));
    }
  }
  CvoBatchIRLS batch_irls_problem(frames, frames_to_hold_const, binary_states, &params);
  batch_irls_problem.solve();
  auto end = std::chrono::system_clock::now();
  std::chrono::duration<double, std::milli> t_all = end - start;
  if (registration_seconds) *registration_seconds = (double)t_all.count() / 1000;
  return 0;
}

// Kept with the reference's own sources (they use only the public API above): the overload taking
// ready-made BinaryStates and inner_product_cpu (CvoGPU.cpp:261-289, :96-213), CvoPointCloud_to_pcl.

}  // namespace cvo
