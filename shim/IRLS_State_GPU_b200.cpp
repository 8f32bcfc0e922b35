// IRLS_State_GPU_b200.cpp — the GPU half of the multi-frame registration forwarded to
// libcvo_b200.so: cvo::CvoFrameGPU and cvo::BinaryStateGPU.
//
// Compiled INSTEAD OF the reference's src/cvo/CvoFrameGPU.cu, src/cvo/IRLS_State_GPU.cu and the
// host helpers of src/cvo/SparseKernelMat.cu (CMakeLists.txt:176-192), against the reference's own
// headers (cvo/CvoFrameGPU.hpp, cvo/IRLS_State_GPU.hpp:21-95, cvo/SparseKernelMat.hpp).  The rest of
// the pose-graph path stays the reference's: CvoFrame.cpp, IRLS.cpp (CvoBatchIRLS::solve),
// IRLS_State_GPU.cpp (update_ell, add_residual_to_problem: Ceres) — the latter walks
// A_result_cpu_, which update_inner_product() below fills in the layout it expects.
// Like CvoGPU_b200.cpp this file is compiled against the stand-in headers under shim/stubs/ in the
// build container (tests/test_shim_syntax.py) and RUN against them on a B200 by
// tests/test_shim_runtime_gpu.py (tests/shim_runtime/driver.cpp).
//
// The class layouts are the reference's, so: the device handle of an edge is found through its
// params_cpu_ pointer (one handle per CvoParams object, created on first use, alive for the
// process); a frame registers itself with a handle the first time an edge on that handle uses
// it; A_host_.nonzero_sum carries the fullest row of the LAST matrix (what max_neighbors(&A_host_)
// returns upstream, IRLS_State_GPU.cu:45).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <unordered_map>
#include <vector>

#ifndef CVO_SHIM_SYNTAX_CHECK
#include "cvo/CvoFrame.hpp"
#include "cvo/CvoFrameGPU.hpp"
#include "cvo/CvoParams.hpp"
#include "cvo/IRLS_State_GPU.hpp"
#include "cvo/SparseKernelMat.hpp"
#include "utils/CvoPointCloud.hpp"
#endif
#include "cvo_b200.h"
#include "cvo_b200_batch.hpp"
#include "shim_pack.hpp"

namespace cvo {

// ---- host helpers of SparseKernelMat.cu:159-200 (plain C++, no CUDA) ------------------------
void clear_SparseKernelMat_cpu(SparseKernelMat* A_cpu, int num_neighbors) {
  A_cpu->nonzero_sum = 0;
  std::memset(A_cpu->mat, 0, sizeof(float) * (size_t)A_cpu->rows * num_neighbors);
  std::memset(A_cpu->ind_row2col, -1, sizeof(int) * (size_t)A_cpu->rows * num_neighbors);
  std::memset(A_cpu->nonzeros, 0, sizeof(unsigned int) * (size_t)A_cpu->rows);
}
void init_internal_SparseKernelMat_cpu(int rows, int cols, SparseKernelMat* A_cpu) {
  A_cpu->rows = rows;
  A_cpu->cols = cols;
  A_cpu->nonzero_sum = 0;
  A_cpu->mat = new float[(size_t)rows * cols]();
  A_cpu->ind_row2col = new int[(size_t)rows * cols]();
  std::memset(A_cpu->ind_row2col, -1, sizeof(int) * (size_t)rows * cols);
  A_cpu->nonzeros = new unsigned int[rows]();
}
void delete_internal_SparseKernelMat_cpu(SparseKernelMat* A_cpu) {
  delete[] A_cpu->mat;
  delete[] A_cpu->ind_row2col;
  delete[] A_cpu->nonzeros;
}

namespace {
std::mutex g_mu;
std::unordered_map<const CvoParams*, cvo_b200_handle*> g_edge_handles;
// frame ids per handle: released ids are reused (a sliding-window run constructs new CvoFrameGPU
// objects for ever; the library accepts ids < 65536)
struct IdPool {
  int next = 0;
  std::vector<int> free_ids;
};
std::unordered_map<cvo_b200_handle*, IdPool> g_ids;
int take_id(cvo_b200_handle* h) {
  std::lock_guard<std::mutex> lk(g_mu);
  IdPool& p = g_ids[h];
  if (!p.free_ids.empty()) {
    const int id = p.free_ids.back();
    p.free_ids.pop_back();
    return id;
  }
  return p.next++;
}
void give_id(cvo_b200_handle* h, int id) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_ids[h].free_ids.push_back(id);
}

[[noreturn]] void die(const cvo_b200_handle* h, const char* what, int rc) {
  std::fprintf(stderr, "[cvo_b200] %s failed (%d): %s\n", what, rc,
               h ? cvo_b200_last_error(h) : cvo_b200_global_error());
  std::exit(EXIT_FAILURE);
}

cvo_b200_handle* edge_handle(const CvoParams* params_cpu) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_edge_handles.find(params_cpu);
  if (it != g_edge_handles.end()) return it->second;
  static_assert(sizeof(CvoParams) == sizeof(cvo_b200_params), "CvoParams layout");
  cvo_b200_handle* h = nullptr;
  int rc = cvo_b200_create(reinterpret_cast<const cvo_b200_params*>(params_cpu), /*device=*/0, &h);
  if (rc != CVO_B200_OK) die(nullptr, "cvo_b200_create", rc);
  g_edge_handles[params_cpu] = h;
  return h;
}
}  // namespace

void shim::forget_edge_handle(const void* params_cpu) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_edge_handles.erase(static_cast<const CvoParams*>(params_cpu));
}

// ---- CvoFrameGPU (CvoFrameGPU.cu:7-100) ------------------------------------------------------
class CvoFrameGPU_Impl {
 public:
  explicit CvoFrameGPU_Impl(const CvoPointCloud* pts) : packed(shim::pack(*pts)) {}
  ~CvoFrameGPU_Impl() {
    for (auto& kv : ids) {
      cvo_b200_frame_clear(kv.first, kv.second);
      give_id(kv.first, kv.second);
    }
  }
  // the frame's id on handle h; the cloud goes to that device on first use (points_init_gpu_)
  int id_on(cvo_b200_handle* h) {
    auto it = ids.find(h);
    if (it != ids.end()) return it->second;
    const int id = take_id(h);
    int rc = cvo_b200_frame_set(h, id, packed.n, packed.xyz.data(), packed.F, packed.p_feat(), packed.C,
                                packed.p_lab(), packed.p_geo());
    if (rc != CVO_B200_OK) die(h, "cvo_b200_frame_set", rc);
    ids[h] = id;
    return id;
  }

 private:
  shim::Packed packed;
  std::unordered_map<cvo_b200_handle*, int> ids;
};

namespace {
// `impl` is private in the reference's class: edges reach a frame's implementation through this
// side table, filled by the constructor
std::unordered_map<const CvoFrameGPU*, CvoFrameGPU_Impl*> g_frames;

int frame_id_on(const CvoFrameGPU* f, cvo_b200_handle* h) {
  CvoFrameGPU_Impl* impl = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_frames.find(f);
    if (it != g_frames.end()) impl = it->second;
  }
  if (!impl) die(h, "frame lookup (CvoFrameGPU not constructed through this library)", CVO_B200_ERR_STATE);
  return impl->id_on(h);
}
}  // namespace

CvoFrameGPU::CvoFrameGPU(const CvoPointCloud* pts, const double poses[12])
    : CvoFrame(pts, poses), impl(new CvoFrameGPU_Impl(pts)) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_frames[this] = impl.get();
}
CvoFrameGPU::~CvoFrameGPU() {
  std::lock_guard<std::mutex> lk(g_mu);
  g_frames.erase(this);
}
// The edge update moves both frames on the device from pose_vec (cvo_b200_edge_update), so the
// per-iteration call of CvoBatchIRLS::solve (IRLS.cpp:106) has nothing left to do here.
void CvoFrameGPU::transform_pointcloud() {}
// Device pointers of the reference's own buffers; only the replaced IRLS_State_GPU.cu read them.
const CvoPoint* CvoFrameGPU::points_transformed_gpu() const { return nullptr; }
const float* CvoFrameGPU::pose_vec_gpu() const { return nullptr; }

// ---- BinaryStateGPU (IRLS_State_GPU.cu:16-89) ------------------------------------------------
// The members are private in the reference's class and a free function cannot be befriended from
// here, so every edge registers two closures (created in its constructor, i.e. with member
// access): `describe` = what the next update will ask the device for, `accept` = take the answer.
// update_inner_product() is describe -> cvo_b200_edge_update -> accept; update_inner_product_batch()
// (cvo_b200_batch.hpp) is describe x n -> ONE cvo_b200_edge_update_batch -> accept x n.
namespace {
struct EdgeHooks {
  std::function<cvo_b200_handle*(cvo_b200_edge&)> describe;
  std::function<void(int64_t, int32_t, const int32_t*, const int32_t*, const float*)> accept;
  int rows = 0;  // points of frame 1 = rows of the edge's matrix
};
std::unordered_map<const BinaryStateGPU*, EdgeHooks> g_hooks;
}  // namespace

BinaryStateGPU::BinaryStateGPU(std::shared_ptr<CvoFrameGPU> pc1, std::shared_ptr<CvoFrameGPU> pc2,
                               const CvoParams* params_cpu, const CvoParams* params_gpu,
                               unsigned int num_neighbor, float init_ell)
    : frame1_(pc1), frame2_(pc2), num_neighbors_(num_neighbor), ell_(init_ell), iter_(0),
      init_num_neighbors_(num_neighbor), params_gpu_(params_gpu), params_cpu_(params_cpu) {
  init_internal_SparseKernelMat_cpu(pc1->points->size(), num_neighbor, &A_result_cpu_);
  A_device_ = nullptr;  // the matrix lives in the handle's workspace
  A_host_.rows = pc1->points->size();
  A_host_.cols = (int)num_neighbor;
  A_host_.nonzero_sum = 0;  // fullest row of the last matrix
  A_host_.mat = nullptr;
  A_host_.ind_row2col = nullptr;
  A_host_.nonzeros = nullptr;
  EdgeHooks hk;
  hk.rows = (int)pc1->points->size();
  hk.describe = [this](cvo_b200_edge& e) -> cvo_b200_handle* {
    cvo_b200_handle* h = edge_handle(params_cpu_);
    // callers mutate *params_cpu_ between solves (CvoGPU::get_params()): upload it like write_params
    int rc = cvo_b200_write_params(h, reinterpret_cast<const cvo_b200_params*>(params_cpu_));
    if (rc != CVO_B200_OK) die(h, "cvo_b200_write_params", rc);
    const unsigned int last_num_neibors = A_host_.nonzero_sum;  // IRLS_State_GPU.cu:45-47
    if (last_num_neibors > 0)
      num_neighbors_ = std::min(init_num_neighbors_, (unsigned int)(last_num_neibors * 1.1));
    for (int i = 0; i < 12; i++) {  // CvoFrameGPU.cu:47-53: the double pose narrowed to float
      e.pose1[i] = static_cast<float>(frame1_->pose_vec[i]);
      e.pose2[i] = static_cast<float>(frame2_->pose_vec[i]);
    }
    e.frame1 = frame_id_on(frame1_.get(), h);
    e.frame2 = frame_id_on(frame2_.get(), h);
    e.ell = ell_;
    e.num_neighbors = (int32_t)num_neighbors_;
    return h;
  };
  hk.accept = [this](int64_t nnz, int32_t max_row, const int32_t* row_ptr, const int32_t* cols, const float* vals) {
    // CSR -> the row-strided layout add_residual_to_problem walks (IRLS_State_GPU.cpp:14-45:
    // stride num_neighbors_, rows end at the first -1)
    const int rows = A_result_cpu_.rows;
    clear_SparseKernelMat_cpu(&A_result_cpu_, (int)num_neighbors_);
    for (int r = 0; r < rows; r++) {
      const int32_t b = row_ptr[r], e = row_ptr[r + 1];
      for (int32_t k = b; k < e; k++) {
        A_result_cpu_.mat[(size_t)r * num_neighbors_ + (k - b)] = vals[k];
        A_result_cpu_.ind_row2col[(size_t)r * num_neighbors_ + (k - b)] = cols[k];
      }
      A_result_cpu_.nonzeros[r] = (unsigned int)(e - b);
    }
    A_result_cpu_.nonzero_sum = (unsigned int)nnz;
    A_host_.nonzero_sum = (unsigned int)max_row;
    iter_++;
  };
  std::lock_guard<std::mutex> lk(g_mu);
  g_hooks[this] = std::move(hk);
}

BinaryStateGPU::~BinaryStateGPU() {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_hooks.erase(this);
  }
  delete_internal_SparseKernelMat_cpu(&A_result_cpu_);
}

namespace {
const EdgeHooks& hooks_of(const BinaryStateGPU* s) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_hooks.find(s);
  if (it == g_hooks.end()) die(nullptr, "edge lookup (BinaryStateGPU not constructed through this library)", CVO_B200_ERR_STATE);
  return it->second;  // entries live as long as their edge; the map is node based
}
}  // namespace

int BinaryStateGPU::update_inner_product() {
  const EdgeHooks& hk = hooks_of(this);
  cvo_b200_edge e;
  cvo_b200_handle* h = hk.describe(e);
  const int rows = A_result_cpu_.rows;
  int64_t nnz = 0;
  int32_t max_row = 0;
  std::vector<int32_t> row_ptr((size_t)rows + 1, 0);
  int rc = cvo_b200_edge_update(h, e.frame1, e.pose1, e.frame2, e.pose2, e.ell, e.num_neighbors, &nnz, &max_row,
                                row_ptr.data(), nullptr, nullptr);
  if (rc != CVO_B200_OK) die(h, "cvo_b200_edge_update", rc);
  std::vector<int32_t> cols((size_t)nnz);
  std::vector<float> vals((size_t)nnz);
  if (nnz > 0) {  // same arguments: the matrix is still on the device, nothing is recomputed
    rc = cvo_b200_edge_update(h, e.frame1, e.pose1, e.frame2, e.pose2, e.ell, e.num_neighbors, &nnz, &max_row,
                              row_ptr.data(), cols.data(), vals.data());
    if (rc != CVO_B200_OK) die(h, "cvo_b200_edge_update", rc);
  }
  hk.accept(nnz, max_row, row_ptr.data(), cols.data(), vals.data());
  return (int)nnz;
}

// The whole edge loop of an outer IRLS iteration (IRLS.cpp:111-121) as one device call per handle:
// every matrix equals what update_inner_product() of that edge would have produced.
int update_inner_product_batch(const std::vector<BinaryStateGPU*>& states) {
  std::unordered_map<cvo_b200_handle*, std::vector<size_t>> by_handle;
  std::vector<cvo_b200_edge> desc(states.size());
  std::vector<const EdgeHooks*> hk(states.size());
  for (size_t k = 0; k < states.size(); k++) {
    hk[k] = &hooks_of(states[k]);
    by_handle[hk[k]->describe(desc[k])].push_back(k);
  }
  long total_all = 0;
  for (auto& kv : by_handle) {
    cvo_b200_handle* h = kv.first;
    const std::vector<size_t>& idx = kv.second;
    std::vector<cvo_b200_edge> edges;
    size_t rows_total = 0;
    for (size_t k : idx) {
      edges.push_back(desc[k]);
      rows_total += (size_t)hk[k]->rows + 1;
    }
    std::vector<int64_t> nnz(idx.size(), 0);
    std::vector<int32_t> mx(idx.size(), 0), row_ptr(rows_total, 0);
    int rc = cvo_b200_edge_update_batch(h, (int)edges.size(), edges.data(), nnz.data(), mx.data(), row_ptr.data(),
                                        nullptr, nullptr);
    if (rc == CVO_B200_ERR_STATE) {  // an edge outside the batched regimes: the per-edge calls serve every case
      for (size_t k : idx) total_all += states[k]->update_inner_product();
      continue;
    }
    if (rc != CVO_B200_OK) die(h, "cvo_b200_edge_update_batch", rc);
    int64_t total = 0;
    for (int64_t v : nnz) total += v;
    std::vector<int32_t> cols((size_t)std::max<int64_t>(total, 1));
    std::vector<float> vals((size_t)std::max<int64_t>(total, 1));
    if (total > 0) {
      rc = cvo_b200_edge_update_batch(h, (int)edges.size(), edges.data(), nnz.data(), mx.data(), row_ptr.data(),
                                      cols.data(), vals.data());
      if (rc != CVO_B200_OK) die(h, "cvo_b200_edge_update_batch", rc);
    }
    size_t ro = 0;
    int64_t eo = 0;
    for (size_t j = 0; j < idx.size(); j++) {
      hk[idx[j]]->accept(nnz[j], mx[j], row_ptr.data() + ro, cols.data() + eo, vals.data() + eo);
      ro += (size_t)hk[idx[j]]->rows + 1;
      eo += nnz[j];
    }
    total_all += (long)total;
  }
  return (int)total_all;
}

CvoFrame* BinaryStateGPU::frame1() { return frame1_.get(); }
CvoFrame* BinaryStateGPU::frame2() { return frame2_.get(); }

}  // namespace cvo
