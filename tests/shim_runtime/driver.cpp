// Runtime driver of the reference-facing C++ classes (tests/test_shim_runtime_gpu.py, TEST ONLY).
// Built against shim/stubs/reference_api_stub.hpp (Eigen / PCL / the reference headers are not in
// this image) together with shim/CvoGPU_b200.cpp and shim/IRLS_State_GPU_b200.cpp, linked to
// libcvo_b200.so, and run on a B200: the calls a reference driver makes
// (main_cvo_gpu_align_two_color_pcd.cpp, main_multi_frame_irls_*.cpp) go through cvo::CvoGPU,
// cvo::CvoFrameGPU and cvo::BinaryStateGPU exactly as declared in the reference's headers.  The
// members the reference's own kept sources define (CvoFrame.cpp, IRLS.cpp, IRLS_State_GPU.cpp's
// Ceres half) are stood in below; add_residual_to_problem walks A_result_cpu_ the way
// IRLS_State_GPU.cpp:14-45 does, so the layout update_inner_product fills is what gets checked.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "cvo_b200_batch.hpp"

namespace ceres {
class Problem {
 public:
  double checksum = 0.0;
  long entries = 0;
  long rows_with_entries = 0;
};
}  // namespace ceres

namespace cvo {
// ---- kept reference sources, stood in
float CvoGPU::inner_product_cpu(const CvoPointCloud&, const CvoPointCloud&, const Eigen::Matrix4f&, float) const { return 1.f; }
CvoFrame::CvoFrame(const CvoPointCloud* pts, const double poses[12]) : points(pts) {
  for (int i = 0; i < 12; i++) pose_vec[i] = poses[i];
}
void CvoFrame::transform_pointcloud() {}
void BinaryStateGPU::update_ell() {
  if (ell_ > params_cpu_->multiframe_ell_min) ell_ = ell_ * params_cpu_->multiframe_ell_decay_rate;
}
void BinaryStateGPU::add_residual_to_problem(ceres::Problem& problem) {
  const SparseKernelMat& A = A_result_cpu_;
  for (int r = 0; r < A.rows; r++) {
    bool any = false;
    for (unsigned int k = 0; k < num_neighbors_; k++) {
      const int idx = A.ind_row2col[(size_t)r * num_neighbors_ + k];
      if (idx == -1) break;
      problem.checksum += (double)(r + 1) * (double)(idx + 1) * (double)A.mat[(size_t)r * num_neighbors_ + k];
      problem.entries++;
      any = true;
    }
    if (any) problem.rows_with_entries++;
  }
}
BinaryStateCPU::BinaryStateCPU(std::shared_ptr<CvoFrame>, std::shared_ptr<CvoFrame>, const CvoParams*) {}
int BinaryStateCPU::update_inner_product() { return 0; }
void BinaryStateCPU::add_residual_to_problem(ceres::Problem&) {}
void BinaryStateCPU::update_ell() {}
namespace {
std::list<std::shared_ptr<BinaryState>> g_states;  // the stand-in CvoBatchIRLS has no members
}
CvoBatchIRLS::CvoBatchIRLS(const std::vector<std::shared_ptr<CvoFrame>>&, const std::vector<bool>&,
                           const std::list<std::shared_ptr<BinaryState>>& states, const CvoParams*) {
  g_states = states;
}
// two outer iterations of IRLS.cpp:77-215 without the Ceres solve: every edge refills its matrix
// (:111-121), hands it to the problem (:123-131) and decays its length-scale (:190-196)
void CvoBatchIRLS::solve() {
  const bool batched = std::getenv("SHIM_DRIVER_BATCH") != nullptr;  // the loop replaced as cvo_b200_batch.hpp shows
  for (int outer = 0; outer < 2; outer++) {
    if (batched) {
      std::vector<BinaryStateGPU*> gpu;
      for (auto&& s : g_states)
        if (auto* g = dynamic_cast<BinaryStateGPU*>(s.get())) gpu.push_back(g);
      std::printf("batch %d total %d\n", outer, update_inner_product_batch(gpu));
    }
    int e = 0;
    for (auto&& st : g_states) {
      const int nnz = batched ? -1 : st->update_inner_product();
      ceres::Problem problem;
      st->add_residual_to_problem(problem);
      std::printf("edge %d %d nnz %d entries %ld rows %ld checksum %.17g\n", outer, e, nnz, problem.entries,
                  problem.rows_with_entries, problem.checksum);
      st->update_ell();
      e++;
    }
  }
  g_states.clear();
}
int CvoGPU::align(std::vector<std::shared_ptr<CvoFrame>>&, const std::vector<bool>&,
                  const std::list<std::shared_ptr<BinaryState>>&, double*) const { return 0; }
}  // namespace cvo

namespace {
// blob: int32 n, F, C, has_geo; float xyz[n*3], feat[n*F], lab[n*C], geo[n*2 if has_geo]
bool load_cloud(const char* path, cvo::CvoPointCloud& pc) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  int32_t hdr[4];
  f.read(reinterpret_cast<char*>(hdr), sizeof(hdr));
  const int n = hdr[0], F = hdr[1], C = hdr[2];
  std::vector<float> xyz((size_t)n * 3), feat((size_t)n * F), lab((size_t)n * C), geo((size_t)n * 2, 0.f);
  f.read(reinterpret_cast<char*>(xyz.data()), xyz.size() * 4);
  f.read(reinterpret_cast<char*>(feat.data()), feat.size() * 4);
  f.read(reinterpret_cast<char*>(lab.data()), lab.size() * 4);
  if (hdr[3]) f.read(reinterpret_cast<char*>(geo.data()), geo.size() * 4);
  if (!f) return false;
  pc.reserve(n, F, C);
  for (int i = 0; i < n; i++) {
    Eigen::Vector3f p;
    p(0) = xyz[3 * (size_t)i]; p(1) = xyz[3 * (size_t)i + 1]; p(2) = xyz[3 * (size_t)i + 2];
    Eigen::VectorXf fe(F), la(C), ge(2);
    for (int j = 0; j < F; j++) fe(j) = feat[(size_t)i * F + j];
    for (int j = 0; j < C; j++) la(j) = lab[(size_t)i * C + j];
    ge(0) = geo[2 * (size_t)i]; ge(1) = geo[2 * (size_t)i + 1];
    pc.add_point(i, p, fe, la, ge);
  }
  return true;
}
void print_mat(const char* key, const Eigen::Matrix4f& T) {
  std::printf("%s", key);
  for (int i = 0; i < 16; i++) std::printf(" %a", (double)T.data()[i]);  // column-major, exact
  std::printf("\n");
}
double assoc_checksum(const cvo::Association& a) {
  double s = 0.0;
  for (const auto& t : a.pairs.t) s += (double)(t.row() + 1) * (double)(t.col() + 1) * (double)t.value();
  return s;
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 5) {
    std::fprintf(stderr, "usage: driver align <yaml> <src.bin> <tgt.bin> | driver edges <yaml> <poses.bin> <frame.bin>...\n");
    return 2;
  }
  const std::string mode = argv[1];
  cvo::CvoGPU gpu(argv[2]);
  if (mode == "align") {
    cvo::CvoPointCloud src, tgt;
    if (!load_cloud(argv[3], src) || !load_cloud(argv[4], tgt)) return 3;
    std::printf("points %d %d features %d classes %d\n", src.num_points(), tgt.num_points(), src.num_features(), src.num_classes());
    // what main_cvo_gpu_align_two_color_pcd.cpp does: params through get_params / write_params, then align
    cvo::CvoParams& p = gpu.get_params();
    p.is_exporting_association = 1;
    gpu.write_params(&p);
    Eigen::Matrix4f init = Eigen::Matrix4f::Identity(), result = Eigen::Matrix4f::Identity();
    cvo::Association assoc;
    double secs = -1.0;
    const int ret = gpu.align(src, tgt, init, result, &assoc, &secs);
    std::printf("align_ret %d seconds_positive %d\n", ret, secs > 0.0 ? 1 : 0);
    print_mat("transform", result);
    std::printf("association nnz %ld rows %zu cols %zu checksum %.17g\n", assoc.pairs.nonZeros(), assoc.source_inliers.size(),
                assoc.target_inliers.size(), assoc_checksum(assoc));
    // the three calls that reuse the pairwise pass, at the aligned pose (T_target_to_source = result^-1 is
    // what the drivers pass; the identity is enough to compare the marshalling)
    std::printf("inner_product %.9g\n", (double)gpu.inner_product_gpu(src, tgt, init, 0.8f));
    std::printf("function_angle %.9g\n", (double)gpu.function_angle(src, tgt, init, 0.8f, true, true));
    cvo::Association a2;
    gpu.compute_association_gpu(src, tgt, init, 0.8f, a2);
    std::printf("association2 nnz %ld rows %zu checksum %.17g\n", a2.pairs.nonZeros(), a2.source_inliers.size(), assoc_checksum(a2));
    Eigen::Matrix3f K;
    K(0, 0) = 0.30f; K(1, 1) = 0.20f; K(2, 2) = 0.50f; K(0, 1) = K(1, 0) = 0.02f; K(1, 2) = K(2, 1) = 0.01f;
    cvo::Association a3;
    gpu.compute_association_gpu(src, tgt, init, K, a3);
    std::printf("association3 nnz %ld rows %zu checksum %.17g\n", a3.pairs.nonZeros(), a3.source_inliers.size(), assoc_checksum(a3));
    return 0;
  }
  if (mode == "edges") {
    // poses.bin: double[12] per frame, row-major 3x4
    const int n_frames = argc - 4;
    std::ifstream pf(argv[3], std::ios::binary);
    std::vector<double> poses((size_t)n_frames * 12);
    pf.read(reinterpret_cast<char*>(poses.data()), poses.size() * 8);
    if (!pf) return 3;
    std::vector<cvo::CvoPointCloud> clouds((size_t)n_frames);
    std::vector<std::shared_ptr<cvo::CvoFrame>> frames;
    for (int k = 0; k < n_frames; k++) {
      if (!load_cloud(argv[4 + k], clouds[(size_t)k])) return 3;
      frames.push_back(std::shared_ptr<cvo::CvoFrame>(new cvo::CvoFrameGPU(&clouds[(size_t)k], &poses[(size_t)k * 12])));
    }
    std::list<std::pair<std::shared_ptr<cvo::CvoFrame>, std::shared_ptr<cvo::CvoFrame>>> edges;
    for (int k = 0; k < n_frames; k++) edges.push_back({frames[(size_t)k], frames[(size_t)((k + 1) % n_frames)]});
    std::vector<bool> hold_const((size_t)n_frames, false);
    hold_const[0] = true;
    double secs = -1.0;
    const int ret = gpu.align(frames, hold_const, edges, &secs);  // CvoGPU.cu:1637-1686 through the shim
    std::printf("multiframe_ret %d seconds_positive %d\n", ret, secs > 0.0 ? 1 : 0);
    return 0;
  }
  return 2;
}
