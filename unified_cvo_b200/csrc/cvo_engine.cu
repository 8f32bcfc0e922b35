// cvo_engine.cu — host side of the C-ABI (include/cvo_b200.h): handle, device
// buffers, the iteration launcher, the device-resident align loop, inner product,
// association export, timing helpers and the optional NCCL plumbing.
//
// Replaces, on the host side: CvoGPU ctor/dtor/write_params (CvoGPU.cu:64-83),
// CvoPointCloud_to_gpu (CvoGPU_impl.cu:206-285), CvoState ctor/dtor (CvoState.cu:22-125),
// the loop skeleton of align_impl (CvoGPU.cu:1338-1572), inner_product_impl
// (:1719-1778), gpu_association_to_cpu (CvoGPU_impl.cu:366-427).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "cvo_device.cuh"
#include "cvo_upload.cuh"
#include "cvo_export.cuh"

namespace cvo_b200 {
void launch_prep(const IterArgs& A, int blocks, cudaStream_t s);
void launch_pair(const IterArgs& A, int blocks, cudaStream_t s);
void launch_tile(const IterArgs& A, int blocks, cudaStream_t s);
int tile_kernel_max_blocks_per_sm();
void launch_flow(const IterArgs& A, int blocks, cudaStream_t s);
void launch_step(const IterArgs& A, int blocks, cudaStream_t s);
void launch_finalize_flow(const IterArgs& A, const double* gathered, int stride, cudaStream_t s);
void launch_finalize_step(const IterArgs& A, const double* gathered, int stride, cudaStream_t s);
void launch_init_bound(const IterArgs& A, cudaStream_t s);
void launch_pose_source(const float4* xyz, int n, const float pose1[12], float4* out_xyz, float4* out_rowA,
                        cudaStream_t s);
void launch_fma_peak(int kind, int iters, int blocks, float* sink, cudaStream_t s);
int pair_kernel_max_blocks_per_sm();
int sparse_kernel_max_blocks_per_sm();
int grid_kernel_max_blocks_per_sm();
cudaError_t launch_align_grid(const IterArgs& A, int blocks, int threads, cudaStream_t s);
int align_grid_max_blocks_per_sm(int threads, int tile);
}  // namespace cvo_b200

using namespace cvo_b200;

namespace {
thread_local std::string g_error;

// ---- NCCL through dlopen: the single-GPU path has no NCCL dependency at all ------
typedef struct ncclComm* ncclComm_t;
struct UniqueId {  // ncclUniqueId: 128 opaque bytes, passed BY VALUE to ncclCommInitRank
  char internal[128];
};
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, UniqueId, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*CommAbort)(ncclComm_t) = nullptr;  // optional
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
const int kNcclFloat64 = 8;  // ncclDouble

bool load_nccl(std::string& err) {
  if (g_nccl.lib) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) {
    err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
    return false;
  }
  g_nccl.GetUniqueId = (int (*)(void*))dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.CommInitRank =
      (int (*)(ncclComm_t*, int, UniqueId, int))dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))dlsym(
      g_nccl.lib, "ncclAllGather");
  g_nccl.CommDestroy = (int (*)(ncclComm_t))dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.CommAbort = (int (*)(ncclComm_t))dlsym(g_nccl.lib, "ncclCommAbort");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy) {
    err = "libnccl is missing a required symbol";
    return false;
  }
  return true;
}

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  cudaError_t ensure(size_t n) {
    if (n <= cap && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = n < 16 ? 16 : n;
    cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct CloudDev {
  int n = 0, F = 0, C = 0;
  bool has_geo = false;
  // Morton-ordered arrays (source role, and target view 0)
  DevBuf<float4> xyz;
  DevBuf<float4> rowA;  // prefilter records (any cloud can play the source role)
  DevBuf<float> feat;   // n * Fp
  DevBuf<float> lab;    // n * Cp
  DevBuf<float2> geo;
  // original-order arrays (target view 1)
  DevBuf<float4> xyz_o;
  DevBuf<float> feat_o;
  DevBuf<float> lab_o;
  DevBuf<float2> geo_o;
  // pruning data over the Morton order
  DevBuf<float4> blk_sphere;   // per 256 points (target role)
  DevBuf<float4> tile_sphere;  // per 64 points (source role)
  DevBuf<float> tile_maxdist;
  DevBuf<int> inv;             // original index -> Morton position (device)
  std::vector<int> perm;       // Morton position -> original index (host copy, fetched lazily for exports)
  bool perm_on_host = false;
  DevBuf<int> perm_d;          // the same on the device
  DevBuf<unsigned long long> keys_d;  // sorted Morton keys (build only)
  // cell index over the Morton order (target role, grid mode): see GridView
  DevBuf<uint32_t> coarse;
  int cbits = 0, n_finite = 0;
  float lo[3] = {0, 0, 0};
  float key_scale = 0;         // lattice units per metre
  double extent = 0;           // edge of the key lattice's cube [m]
  double occupied_volume = 0;  // volume of the occupied coarse cells [m^3] (density estimate)
  float max_dist = 0;          // max_i |x_i| (largest range-scaled length-scale, source role)
  int Fp = 0, Cp = 0;   // strides the buffers were packed with
  float cx = 0, cy = 0, cz = 0;  // centroid (float)
  float radius = 0;              // max_i |x_i - centroid| (rounded up)
  bool set = false;
};

// A pose-graph frame resident on the device in the caller's layout (the points_init_gpu_ of a
// CvoFrameGPU, CvoFrameGPU.cu:7-30): edge updates move it by the frame's current pose and build
// the two cloud slots from it without touching the host.
struct FrameDev {
  int n = 0, F = 0, C = 0;
  bool has_geo = false, set = false;
  DevBuf<float> xyz, feat, lab, geo;
  // the frame as a cloud in its OWN coordinates (Morton order, cell table): edges are evaluated in
  // these tables at any pose, nothing is rebuilt per edge
  CloudDev cloud;
  void release() {
    xyz.release(); feat.release(); lab.release(); geo.release();
    CloudDev& c = cloud;
    c.xyz.release(); c.rowA.release(); c.feat.release(); c.lab.release(); c.geo.release();
    c.xyz_o.release(); c.feat_o.release(); c.lab_o.release(); c.geo_o.release();
    c.blk_sphere.release(); c.tile_sphere.release(); c.tile_maxdist.release(); c.inv.release();
    c.coarse.release(); c.perm_d.release(); c.keys_d.release();
    c.set = false;
    set = false;
    n = 0;
  }
};

// a cloud slot built from a frame at a pose (edges sharing a frame at the same pose reuse it)
struct SlotKey {
  bool valid = false;
  int frame = -1;
  unsigned long long gen = 0;
  float pose[12] = {};
};

// what the ELL matrix holds after cvo_b200_edge_update (second call of the two-call protocol)
struct EdgeKey {
  int f1, f2, cap;
  float ell;
  float p1[12], p2[12];
  unsigned long long gen1, gen2;  // generation of the frames' contents
};
}  // namespace

struct cvo_b200_handle {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  cvo_b200_params params;
  cvo_b200_params* d_params = nullptr;
  DevState* d_state = nullptr;
  CloudDev src, tgt;
  // per-iteration workspace
  DevBuf<float4> tgt_moved;
  DevBuf<float> px, py, pz, pw;
  DevBuf<uint32_t> pq;
  DevBuf<float4> rowrec;
  DevBuf<float2> row_lt;
  DevBuf<uint32_t> sat_list;
  DevBuf<uint32_t> cand, cand_cnt, ell_idx, row_nnz;
  DevBuf<float> ell_val;
  DevBuf<FlowPartial> flow_part, flow_part2;
  DevBuf<LLBoard> ll_board;  // barrier-free reductions of the persistent kernel
  DevBuf<StepPartial> step_part;
  DevBuf<float> zeros_f;    // stand-in for absent features / labels
  DevBuf<float2> zeros_g;   // stand-in for absent geometric types
  DevBuf<cvo_b200_iter_trace> d_trace;
  DevBuf<double> gathered;  // multi-GPU all-gather receive buffer
  DevBuf<unsigned long long> stamps;  // debug: per-block phase stamps (CVO_B200_STAMPS=1)
  // cloud build scratch (cvo_upload.cu)
  DevBuf<float> raw_xyz, raw_feat, raw_lab, raw_geo;
  DevBuf<unsigned long long> keys_in;
  DevBuf<int> idx_in;
  DevBuf<unsigned char> sort_temp;
  DevBuf<CloudStats> d_stats;
  // CSR export scratch (cvo_export.cu)
  DevBuf<int> csr_cnt, csr_ptr;
  DevBuf<unsigned char> csr_temp;
  DevBuf<int32_t> csr_cols;
  DevBuf<float> csr_vals;
  std::vector<int32_t> csr_rp_host;
  CloudStats* h_stats = nullptr;  // pinned
  int row_begin = 0, row_end = -1;
  // launch geometry
  int prep_blocks = 1, pair_blocks = 1, sparse_blocks = 1;
  // graph cache for the align loop: [0] dense scan (prep, pair, flow, step), [1] cell queries
  // (flow, step), [2] tile cells (tile, flow, step)
  cudaGraphExec_t graph_exec[3] = {nullptr, nullptr, nullptr};
  IterArgs graph_args[3];
  int graph_batch[3] = {0, 0, 0};
  int tile_blocks = 1;
  bool use_graph = true;
  int grid_blocks = 1;
  int persist_blocks = 1;   // cooperative grid of align_grid_kernel (all blocks co-resident)
  int persist_blocks_tile = 1;  // ... of its tile-cell instantiation
  int persist_blocks_brute = 1;  // ... when every row takes the exact one-warp walk (IterArgs::brute)
  int persist_threads = kPersistThreads;
  int persist_threads_tile = kPersistThreads;
  bool comm_broken = false;  // a multi-GPU exchange failed: tear the communicator down without waiting for peers
  bool use_persist = true;  // CVO_B200_PERSIST=0: one launch per phase even in cell-query mode
  int last_tile_builds = 0;   // iterations of the last align() that built candidate cells (persistent tile mode)
  float verlet_kappa = 0.1f;  // candidate-cell reuse of the persistent tile mode: skin / cut-off radius (CVO_B200_VERLET)
  int force_mode = -1;  // CVO_B200_MODE: -1 auto, 0 dense scan, 1 cell queries, 2 tile cells, 3 brute rows (where possible)
  // host poll buffer (pinned)
  int* h_poll = nullptr;
  // comm
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  bool shard_inner_products = false;  // cvo_b200_comm_shard_inner_products
  // fused exchange: own mailbox + the peers' (CUDA IPC), generation counter of persistent launches
  XMailbox* mailbox = nullptr;
  XMailbox* peers[kMaxWorld] = {};
  bool peers_ready = false;
  unsigned long long xgen = 0;
  std::string err;
  uint64_t launches = 0;
  // pose-graph frames (cvo_b200_frame_set) and the edge the ELL matrix currently holds
  std::vector<FrameDev> frames;
  std::vector<unsigned long long> frame_gen;
  unsigned long long frame_gen_next = 1;
  // cvo_b200_edge_update_batch: the edges' row pointers (device + host copy) and running offsets
  DevBuf<int> batch_ptr;
  DevBuf<long long> batch_base;
  std::vector<int> batch_rp_host;
  std::vector<cvo_b200_edge> batch_edges;
  std::vector<unsigned long long> batch_gens;
  bool batch_valid = false;
  bool edge_valid = false;
  bool edge_own_frame = false;            // the matrix came from the own-frame path (Morton columns)
  const int* edge_row_inv = nullptr;
  const int* edge_col_perm = nullptr;
  EdgeKey edge_key;
  SlotKey slot_key[2];  // which posed frame each cloud slot currently holds ([0] source, [1] target)
  IterArgs edge_args;
  DevBuf<float4> edge_src_xyz, edge_src_rowA;  // frame 1 of the current edge at its pose
  int cap_override = 0;  // ELL stride of the next prepare() when an edge asks for more than nearest_neighbors_max
  // the kernel matrix left behind by the last align() (cvo_b200_align_association)
  bool last_valid = false;
  int last_view = 1;
  IterArgs last_args;
};

namespace {
#define CVO_CUDA(h, expr)                                                               \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      (h)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                   \
      return CVO_B200_ERR_CUDA;                                                         \
    }                                                                                   \
  } while (0)

int fail(cvo_b200_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  g_error = msg;
  return code;
}
inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// Pick the work decomposition of the pair kernel and size every buffer.
int prepare(cvo_b200_handle* h, IterArgs& A, int mode, const float* kinv, const CloudDev* S = nullptr,
            const CloudDev* Tg = nullptr, bool sharded = true) {
  if ((!S && !h->src.set) || (!Tg && !h->tgt.set)) return fail(h, CVO_B200_ERR_STATE, "source/target cloud not set");
  h->last_valid = false;  // every caller of prepare() overwrites the ELL matrix
  h->edge_valid = false;
  // (a cached batch keeps its compacted entries in csr_cols / csr_vals: any later export reuses
  // those buffers, so the batch cache is dropped as well)
  h->batch_valid = false;
  const CloudDev& cs = S ? *S : h->src;
  const CloudDev& ct = Tg ? *Tg : h->tgt;
  const int N = cs.n, M = ct.n;
  int rb = sharded ? h->row_begin : 0;
  int re = (!sharded || h->row_end < 0 || h->row_end > N) ? N : h->row_end;
  if (sharded && h->world > 1) {
    // multi-GPU job: the shard follows (rank, world) and the CURRENT source size, so that a new
    // source cloud of another size can neither leave rows unscanned nor make ranks disagree
    // (contiguous blocks starting on a source tile, unified_cvo_b200/dist.py::shard_rows)
    const int per = round_up((N + h->world - 1) / h->world, kTileRows);
    rb = std::min(N, h->rank * per);
    re = std::min(N, rb + per);
  }
  if (rb < 0 || rb > re) return fail(h, CVO_B200_ERR_INVALID, "bad row range");
  const int n_rows = re - rb;
  const int Fp = std::max(cs.Fp, ct.Fp);
  const int Cp = std::max(cs.Cp, ct.Cp);
  if ((cs.Fp && ct.Fp && cs.Fp != ct.Fp) || (cs.Cp && ct.Cp && cs.Cp != ct.Cp))
    return fail(h, CVO_B200_ERR_INVALID, "source and target feature/class dimensions differ");
  const int cap_max = std::max(std::max(1, h->params.nearest_neighbors_max), h->cap_override);

  // ---- chunking: enough (row tile, target chunk) items to fill the machine a few times
  const int row_tiles = std::max(1, (n_rows + kTileRows - 1) / kTileRows);
  const int target_items = h->num_sms * 96;
  int nchunks = std::max(1, (target_items + row_tiles - 1) / row_tiles);
  const int max_chunks = std::max(1, (M + kJBlock - 1) / kJBlock);
  nchunks = std::min(nchunks, max_chunks);
  int chunk_len = round_up(std::max(1, (M + nchunks - 1) / nchunks), kJBlock);
  nchunks = std::max(1, (M + chunk_len - 1) / chunk_len);
  // cell capacity in candidate WORDS (one word = one lane's 8 consecutive targets)
  int L = std::min(chunk_len / 8, std::max(64, 2 * cap_max));
  // keep the candidate cells within ~8 GiB
  while ((size_t)std::max(n_rows, 1) * nchunks * L * 4 > ((size_t)8 << 30) && L > 32) L /= 2;
  const int M_pad = round_up(std::max(M, 1), kJBlock);
  if (M >= (1 << 24)) return fail(h, CVO_B200_ERR_INVALID, "target clouds of 2^24 points or more are not supported");
  // tile mode: ONE candidate cell per row; sized for the densest rows the mode is chosen for
  // (choose_mode), halved until the cells fit in ~4 GiB
  // ... shared by tile_parts warps per tile (small shards: enough items to fill the machine)
  int tile_parts = 1;
  // ~3.5 items per resident warp: items differ a lot in cost and are handed out dynamically, so the
  // build ends with its heaviest item (measured on C4: 2 parts 20.2 ms, 4 parts 19.4, 8 parts 19.7)
  while (tile_parts < 8 && 2 * row_tiles * tile_parts < 7 * h->num_sms * 24) tile_parts *= 2;
  if (const char* tp = getenv("CVO_B200_TILE_PARTS")) tile_parts = std::max(1, std::min(8, atoi(tp)));  // measurement aid
  int tile_L = 1024 / tile_parts;
  while ((size_t)std::max(n_rows, 1) * tile_parts * tile_L * 4 > ((size_t)4 << 30) && tile_L > 32) tile_L /= 2;

  const size_t n_zero = (size_t)std::max(N, M) * (size_t)std::max(std::max(Fp, Cp), 1);
  CVO_CUDA(h, h->tgt_moved.ensure((size_t)M));
  CVO_CUDA(h, h->px.ensure((size_t)M_pad));
  CVO_CUDA(h, h->py.ensure((size_t)M_pad));
  CVO_CUDA(h, h->pz.ensure((size_t)M_pad));
  CVO_CUDA(h, h->pw.ensure((size_t)M_pad));
  CVO_CUDA(h, h->pq.ensure((size_t)M_pad));
  CVO_CUDA(h, h->rowrec.ensure((size_t)std::max(n_rows, 1) * 2));
  CVO_CUDA(h, h->row_lt.ensure((size_t)std::max(n_rows, 1)));
  CVO_CUDA(h, h->sat_list.ensure((size_t)std::max(n_rows, 1)));
  CVO_CUDA(h, h->cand.ensure(std::max((size_t)std::max(n_rows, 1) * nchunks * L, (size_t)std::max(n_rows, 1) * tile_parts * tile_L)));
  CVO_CUDA(h, h->cand_cnt.ensure((size_t)std::max(n_rows, 1) * std::max(nchunks, tile_parts)));
  CVO_CUDA(h, h->ell_idx.ensure((size_t)std::max(n_rows, 1) * cap_max));
  CVO_CUDA(h, h->ell_val.ensure((size_t)std::max(n_rows, 1) * cap_max));
  CVO_CUDA(h, h->row_nnz.ensure((size_t)std::max(n_rows, 1)));
  if (h->zeros_f.cap < n_zero) {
    CVO_CUDA(h, h->zeros_f.ensure(n_zero));
    CVO_CUDA(h, cudaMemsetAsync(h->zeros_f.p, 0, h->zeros_f.cap * sizeof(float), h->stream));
  }
  if (h->zeros_g.cap < (size_t)std::max(N, M)) {
    CVO_CUDA(h, h->zeros_g.ensure((size_t)std::max(N, M)));
    CVO_CUDA(h, cudaMemsetAsync(h->zeros_g.p, 0, h->zeros_g.cap * sizeof(float2), h->stream));
  }

  // ---- launch geometry (fixed grids; kernels are grid-stride / work-stealing)
  int occ = pair_kernel_max_blocks_per_sm();
  if (occ < 1) occ = 1;
  const int n_items = row_tiles * nchunks;
  h->pair_blocks = std::max(1, std::min(h->num_sms * occ, (n_items + kPairWarps - 1) / kPairWarps));
  h->prep_blocks = std::max(1, std::min(h->num_sms * 4, (std::max(M_pad, n_rows) + 255) / 256));
  int tocc = tile_kernel_max_blocks_per_sm();
  if (tocc < 1) tocc = 1;
  // items are dealt to warp 0 of every block first: few items spread over all SMs instead of
  // filling a few blocks
  h->tile_blocks = std::max(1, std::min(h->num_sms * tocc, row_tiles * tile_parts));
  const int warps_per_block = kSparseThreads / 32;
  int socc = sparse_kernel_max_blocks_per_sm();
  if (socc < 1) socc = 1;
  // eight lanes per source row: one block covers 32 rows; grid-stride beyond one resident wave
  const int rows_per_block = warps_per_block * 4;
  h->sparse_blocks =
      std::max(1, std::min(h->num_sms * socc, (n_rows + rows_per_block - 1) / rows_per_block));
  int gocc = grid_kernel_max_blocks_per_sm();
  if (gocc < 1) gocc = 1;
  h->grid_blocks =
      std::max(1, std::min(h->num_sms * gocc, (n_rows + rows_per_block - 1) / rows_per_block));
  // wider blocks when the narrow ones would need a second pass over the rows (cvo_device.cuh)
  h->persist_threads = (n_rows > h->num_sms * (kPersistThreads / 32) * 4)        ? kPersistThreadsWide
                       : (n_rows <= h->num_sms * (kPersistThreadsSmall / 32) * 4) ? kPersistThreadsSmall
                                                                                  : kPersistThreads;
  int pocc = align_grid_max_blocks_per_sm(h->persist_threads, 0);
  if (pocc < 1) pocc = 1;
  const int rows_per_pblock = (h->persist_threads / 32) * 4;
  h->persist_blocks =
      std::max(1, std::min(h->num_sms * pocc, (n_rows + rows_per_pblock - 1) / rows_per_pblock));
  // few rows (the README demo: 523): filling block after block would put all the work on a handful
  // of SMs, 18 warps on 4 schedulers each, while the loop is a latency chain per row.  Instead the
  // 4-row groups are dealt warp-major over blocks of 4 busy warps (one per scheduler): more blocks
  // in the all-reduce (+ ~1 us), but the flow phase runs at a lone warp's speed
  // (measured on the demo's 523 rows only: 8 -> 33 blocks costs +0.4 us per all-reduce and buys 4 us
  // of flow phase; with thousands of rows the extra blocks of the all-reduce would cost more than the
  // lighter SMs buy, so the spread stops at 8 rows per SM)
  bool row_spread = n_rows <= 8 * h->num_sms;
  if (const char* rs = getenv("CVO_B200_ROW_SPREAD")) row_spread = row_spread && atoi(rs) != 0;  // measurement aid
  if (row_spread) h->persist_blocks = std::max(1, std::min(h->num_sms * pocc, ((n_rows + 3) / 4 + 3) / 4));
  // brute rows: one warp per row, 4 busy warps per block until every SM has a block
  h->persist_blocks_brute = std::max(1, std::min(std::min(kLLMaxBlocks, h->num_sms * pocc), (n_rows + 3) / 4));
  // tile mode: one block per SM (tile items and rows are dealt to its warps); the wide block where
  // it saves the second pass over the rows (cells are mostly reused: the row walk is the iteration)
  h->persist_threads_tile = (n_rows > h->num_sms * (kPersistThreads / 32) * 4 && n_rows <= h->num_sms * (kPersistThreadsWide / 32) * 4)
                                ? kPersistThreadsWide
                                : kPersistThreads;
  h->persist_blocks_tile =
      std::min(kLLMaxBlocks, h->num_sms * std::max(1, std::min(1, align_grid_max_blocks_per_sm(h->persist_threads_tile, 1))));
  CVO_CUDA(h, h->flow_part.ensure((size_t)std::max(h->sparse_blocks, std::max(h->grid_blocks, h->persist_blocks))));
  CVO_CUDA(h, h->flow_part2.ensure((size_t)h->persist_blocks));
  CVO_CUDA(h, h->ll_board.ensure(1));
  if (h->persist_blocks > kLLMaxBlocks) h->persist_blocks = kLLMaxBlocks;
  CVO_CUDA(h, h->step_part.ensure((size_t)std::max(h->sparse_blocks, std::max(h->grid_blocks, h->persist_blocks))));

  std::memset(&A, 0, sizeof(A));
  A.row_spread = row_spread ? 1 : 0;
  A.params = h->d_params;
  A.st = h->d_state;
  A.src_xyz = cs.xyz.p;
  A.src_rowA = cs.rowA.p;
  A.src_feat = cs.Fp ? cs.feat.p : h->zeros_f.p;
  A.src_lab = cs.Cp ? cs.lab.p : h->zeros_f.p;
  A.src_geo = cs.has_geo ? cs.geo.p : h->zeros_g.p;
  A.row_begin = rb;
  A.n_rows = n_rows;
  A.n_src_total = N;
  A.tv[0].xyz = ct.xyz.p;
  A.tv[0].feat = ct.Fp ? ct.feat.p : h->zeros_f.p;
  A.tv[0].lab = ct.Cp ? ct.lab.p : h->zeros_f.p;
  A.tv[0].geo = ct.has_geo ? ct.geo.p : h->zeros_g.p;
  A.tv[1].xyz = ct.xyz_o.p;
  A.tv[1].feat = ct.Fp ? ct.feat_o.p : h->zeros_f.p;
  A.tv[1].lab = ct.Cp ? ct.lab_o.p : h->zeros_f.p;
  A.tv[1].geo = ct.has_geo ? ct.geo_o.p : h->zeros_g.p;
  // the Morton view needs tile spheres aligned with this shard and a geometric cut-off to prune on
  static const bool no_prune = getenv("CVO_B200_NO_PRUNE") && getenv("CVO_B200_NO_PRUNE")[0] == '1';
  A.prune = (!no_prune && mode == 0 && h->params.is_using_geometry && (rb % kTileRows) == 0) ? 1 : 0;
  A.tgt_inv = ct.inv.p;
  A.sat_list = h->sat_list.p;
  A.blk_sphere = ct.blk_sphere.p;
  A.tile_sphere = cs.tile_sphere.p;
  A.tile_maxdist = cs.tile_maxdist.p;
  A.tgt_moved = h->tgt_moved.p;
  A.px = h->px.p; A.py = h->py.p; A.pz = h->pz.p; A.pw = h->pw.p;
  A.pq = h->pq.p;
  A.M = M;
  A.Fp = Fp;
  A.Cp = Cp;
  A.cx = cs.cx; A.cy = cs.cy; A.cz = cs.cz;
  A.tcx = ct.cx; A.tcy = ct.cy; A.tcz = ct.cz;
  A.trad = ct.radius;
  A.M_pad = M_pad;
  A.rowrec = h->rowrec.p;
  A.row_lt = h->row_lt.p;
  A.cand = h->cand.p;
  A.cand_cnt = h->cand_cnt.p;
  A.nchunks = nchunks;
  A.chunk_len = chunk_len;
  A.L = L;
  A.ell_idx = h->ell_idx.p;
  A.ell_val = h->ell_val.p;
  A.row_nnz = h->row_nnz.p;
  A.cap_max = cap_max;
  A.flow_part = h->flow_part.p;
  A.flow_part2 = h->flow_part2.p;
  A.ll = h->ll_board.p;
  A.step_part = h->step_part.p;
  A.mode = mode;
  if (kinv)
    for (int k = 0; k < 9; k++) A.kinv[k] = kinv[k];
  A.world = sharded ? h->world : 1;
  A.n_items = n_items;
  A.stamps = nullptr;
  if (getenv("CVO_B200_STAMPS")) {
    if (h->stamps.cap < (size_t)8 * 4096) {
      CVO_CUDA(h, h->stamps.ensure((size_t)8 * 4096));
      CVO_CUDA(h, cudaMemsetAsync(h->stamps.p, 0, (size_t)8 * 4096 * sizeof(unsigned long long), h->stream));
    }
    A.stamps = h->stamps.p;
  }
  A.colour = h->params.is_using_intensity ? 1 : 0;
  A.xfused = 0;
  A.xrank = h->rank;
  A.xworld = h->world;
  A.xgen = 0;
  for (int r = 0; r < kMaxWorld; r++) A.xpeer[r] = h->peers[r];
  A.posevec = 0;
  A.edge_slack = 0.f;
  A.verlet_kappa = h->verlet_kappa;
  A.src_rmax = cs.max_dist;
  A.grid = 0;
  A.tile = 0;
  A.tile_L = tile_L;
  A.tile_parts = tile_parts;
  A.src_keys = cs.keys_d.p;
  {
    // cut level of the tiles: cube nodes of the source's octree of edge >= 2 tile edges
    const double rho_s = (cs.n_finite > 0 && cs.occupied_volume > 0.0) ? (double)cs.n_finite / cs.occupied_volume : 1.0;
    const double D = 2.0 * std::cbrt((double)kTileRows / rho_s);
    int b = 0;
    while (b < 21 && (double)(1u << b) < D * (double)cs.key_scale) b++;
    A.tile_cut_bits = 3 * b;
  }
  A.gv.coarse = ct.coarse.p;
  A.gv.cbits = ct.cbits;
  A.gv.n_finite = ct.n_finite;
  A.gv.lo[0] = ct.lo[0]; A.gv.lo[1] = ct.lo[1]; A.gv.lo[2] = ct.lo[2];
  A.gv.scale = ct.key_scale;
  if (h->world > 1) CVO_CUDA(h, h->gathered.ensure((size_t)h->world * 16));
  return CVO_B200_OK;
}

// One iteration's worth of launches.  `stage`: 3 = everything, 2 = stop after flow.
int enqueue_iteration(cvo_b200_handle* h, const IterArgs& A, int stage, cudaEvent_t pair_begin,
                      cudaEvent_t pair_end) {
  cudaStream_t s = h->stream;
  // CVO_B200_DEBUG_SKIP (bit mask 1 prep, 2 pair, 4 flow, 8 step): measurement aid only — the
  // marginal cost of one kernel inside the real pipeline; results are meaningless when set
  static const int skip = getenv("CVO_B200_DEBUG_SKIP") ? atoi(getenv("CVO_B200_DEBUG_SKIP")) : 0;
  const int sparse_blocks = A.grid ? h->grid_blocks : h->sparse_blocks;
  if (A.grid) {  // cell queries: no O(M) prep, no O(N*M) scan
    if (pair_begin) cudaEventRecord(pair_begin, s);
    if (!(skip & 4)) launch_flow(A, sparse_blocks, s);
    if (pair_end) cudaEventRecord(pair_end, s);
    h->launches += 1;
  } else if (A.tile) {  // tile cells: no O(M) prep either; the targets are moved on the fly
    if (pair_begin) cudaEventRecord(pair_begin, s);
    if (!(skip & 2)) launch_tile(A, h->tile_blocks, s);
    if (pair_end) cudaEventRecord(pair_end, s);
    if (!(skip & 4)) launch_flow(A, h->sparse_blocks, s);
    h->launches += 2;
  } else {
    if (!(skip & 1)) launch_prep(A, h->prep_blocks, s);
    if (pair_begin) cudaEventRecord(pair_begin, s);
    if (!(skip & 2)) launch_pair(A, h->pair_blocks, s);
    if (pair_end) cudaEventRecord(pair_end, s);
    if (!(skip & 4)) launch_flow(A, h->sparse_blocks, s);
    h->launches += 3;
  }
  if (A.world > 1) {
    DevState* st = h->d_state;
    int rc = g_nccl.AllGather(&st->local_flow[0], h->gathered.p, 9, kNcclFloat64, h->comm, s);
    if (rc != 0) h->comm_broken = true;
    if (rc != 0) return fail(h, CVO_B200_ERR_NCCL, "ncclAllGather(flow) failed");
    launch_finalize_flow(A, h->gathered.p, 9, s);
    h->launches += 1;
  }
  if (stage >= 3 && !(skip & 8)) {
    launch_step(A, sparse_blocks, s);
    h->launches += 1;
    if (A.world > 1) {
      DevState* st = h->d_state;
      int rc = g_nccl.AllGather(&st->local_step[0], h->gathered.p, 4, kNcclFloat64, h->comm, s);
      if (rc != 0) return fail(h, CVO_B200_ERR_NCCL, "ncclAllGather(step) failed");
      launch_finalize_step(A, h->gathered.p, 4, s);
      h->launches += 1;
    }
  }
  return CVO_B200_OK;
}

// One launch of the persistent kernel: its reduction board and its two monotone counters start
// from zero (stream-ordered memsets, no synchronisation).
cudaError_t launch_persistent(cvo_b200_handle* h, const IterArgs& A) {
  cudaError_t e = cudaMemsetAsync(h->ll_board.p, 0, sizeof(LLBoard), h->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(&h->d_state->work_counter, 0, sizeof(unsigned int), h->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(&h->d_state->n_sat, 0, sizeof(unsigned int), h->stream);
  if (e == cudaSuccess && A.tile) e = cudaMemsetAsync(&h->d_state->item_counter, 0, sizeof(unsigned int), h->stream);
  if (e != cudaSuccess) return e;
  return launch_align_grid(A, A.tile ? h->persist_blocks_tile : (A.brute ? h->persist_blocks_brute : h->persist_blocks),
                           A.tile ? h->persist_threads_tile : h->persist_threads, h->stream);
}

void host_update_tf(const float R[9], const float T[3], float Rinv[9], float Tinv[3]) {
  // CvoGPU.cu:94-112 on the host, for the initial pose only (same arithmetic as the
  // device controller: -R^T applied with c0 + (c1 + c2)).
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Rinv[3 * j + i] = R[3 * i + j];
  for (int i = 0; i < 3; i++) {
    volatile float c0 = (-Rinv[i]) * T[0];
    volatile float c1 = (-Rinv[3 + i]) * T[1];
    volatile float c2 = (-Rinv[6 + i]) * T[2];
    volatile float s12 = c1 + c2;
    Tinv[i] = c0 + s12;
  }
}

// fill_in_A_mat_gpu's scalar prologue (CvoGPU.cu:495-515, and :235-254 for the dense-kernel
// variant), in the same float arithmetic, once per call instead of once per thread
KernConsts make_consts(const cvo_b200_params& p) {
  KernConsts k;
  std::memset(&k, 0, sizeof(k));
  volatile float sigma2 = p.sigma * p.sigma;
  volatile float c2 = p.c_ell * p.c_ell;
  volatile float c_sigma2 = p.c_sigma * p.c_sigma;
  volatile float s_sigma2 = p.s_sigma * p.s_sigma;
  volatile float s_ell_square = p.s_ell * p.s_ell;
  k.sigma2 = sigma2; k.c2 = c2; k.c_sigma2 = c_sigma2; k.s_ell = p.s_ell; k.s_sigma2 = s_sigma2;
  k.s_ell_square = s_ell_square;
  k.sp_thres = p.sp_thres;
  k.use_geo_type = p.is_using_geometric_type;
  k.use_geometry = p.is_using_geometry;
  k.use_intensity = p.is_using_intensity;
  k.use_semantics = p.is_using_semantics;
  k.use_range_ell = p.is_using_range_ell;
  k.c_div = p.c;
  k.d_div = p.d;
  volatile float q_geo = p.sp_thres / sigma2;
  k.log_geo = logf(q_geo);
  k.d2_c_thres = 1.f;
  k.d2_s_thres = 1.f;
  k.d2_s_thres_dense = 1.f;
  if (k.use_intensity) {
    volatile float q = p.sp_thres / c_sigma2;
    k.d2_c_thres = (float)(-2.0 * (double)c2 * (double)logf(q));
  }
  if (k.use_semantics) {
    volatile float q = p.sp_thres / s_sigma2;
    const double lg = (double)logf(q);
    k.d2_s_thres = (float)(-2.0 * (double)p.s_ell * (double)p.s_ell * lg);
    k.d2_s_thres_dense = (float)(-2.0 * (double)s_ell_square * lg);
  }
  return k;
}

int init_state(cvo_b200_handle* h, const IterArgs& A, const float R[9], const float T[3], float ell,
               int cap, int controller_on, int max_iter, cvo_b200_iter_trace* d_trace,
               int trace_cap, bool allow_morton = true) {
  static thread_local DevState hs;  // ~8.5 KB; keep it off the stack of small callers
  std::memset(&hs, 0, sizeof(hs));
  std::memcpy(hs.R, R, sizeof(hs.R));
  std::memcpy(hs.T, T, sizeof(hs.T));
  host_update_tf(R, T, hs.Rinv, hs.Tinv);
  hs.ell = ell;
  hs.num_neighbors = cap;
  hs.max_iter = max_iter;
  hs.controller_on = controller_on;
  hs.trace = d_trace;
  hs.trace_cap = trace_cap;
  hs.kc = make_consts(h->params);
  hs.prune_on = (allow_morton && A.prune) ? 1 : 0;
  hs.view = hs.prune_on ? 0 : 1;
  CVO_CUDA(h, cudaMemcpyAsync(h->d_state, &hs, sizeof(hs), cudaMemcpyHostToDevice, h->stream));
  // the source of an async copy from pageable memory is staged before the call returns
  launch_init_bound(A, h->stream);  // Rinv/Tinv + the |y'-c| bound, same code as the controller
  h->launches += 1;
  return CVO_B200_OK;
}

void split_pose(const float T16[16], float R[9], float T[3]) {
  for (int j = 0; j < 3; j++)
    for (int i = 0; i < 3; i++) R[3 * j + i] = T16[4 * j + i];
  for (int i = 0; i < 3; i++) T[i] = T16[12 + i];
}

void destroy_graph(cvo_b200_handle* h) {
  for (int m = 0; m < 3; m++) {
    if (h->graph_exec[m]) cudaGraphExecDestroy(h->graph_exec[m]);
    h->graph_exec[m] = nullptr;
    h->graph_batch[m] = 0;
  }
}

// Capture `batch` iterations into one graph.  All per-iteration values (pose, ell, cap,
// flags) live in DevState, so the graph is parameter-free and is rebuilt only when the
// buffers or the decomposition change.
int ensure_graph(cvo_b200_handle* h, const IterArgs& A, int batch) {
  const int m = A.grid ? 1 : (A.tile ? 2 : 0);
  if (h->graph_exec[m] && h->graph_batch[m] == batch && std::memcmp(&h->graph_args[m], &A, sizeof(A)) == 0)
    return CVO_B200_OK;
  if (h->graph_exec[m]) cudaGraphExecDestroy(h->graph_exec[m]);
  h->graph_exec[m] = nullptr;
  h->graph_batch[m] = 0;
  cudaGraph_t graph = nullptr;
  CVO_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
  const uint64_t before = h->launches;
  int rc = CVO_B200_OK;
  for (int b = 0; b < batch && rc == CVO_B200_OK; b++) rc = enqueue_iteration(h, A, 3, nullptr, nullptr);
  h->launches = before;  // counted when the graph is launched
  cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
  if (rc != CVO_B200_OK) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess) return fail(h, CVO_B200_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
  e = cudaGraphInstantiate(&h->graph_exec[m], graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) {
    h->graph_exec[m] = nullptr;
    return fail(h, CVO_B200_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(e));
  }
  h->graph_args[m] = A;
  h->graph_batch[m] = batch;
  return CVO_B200_OK;
}

int launches_per_iteration(const cvo_b200_handle* h, const IterArgs& A) {
  return (A.grid ? 2 : (A.tile ? 3 : 4)) + (h->world > 1 ? 2 : 0);
}

// Candidate-generator policy: a cost model calibrated on B200 (r01: C2, KITTI-sized and 200k
// clouds at ell = 0.1 .. 1.5).  Per source row and iteration
//   cell queries  test the points of <= 27 cube cells of edge h in [r, 2r) (r = largest cut-off
//                 radius, CvoGPU.cu:506-511): tests_g = 27 h^3 rho_t, ~2.4 ps each (one lane, one
//                 test) inside the persistent kernel;
//   dense scan    tests every target of the 256-point Morton blocks whose bounding sphere is
//                 within r of the row's 64-row tile: tests_d = rho_t 4/3 pi (r + r_tile + r_blk)^3,
//                 ~0.62 ps each (packed FMA prefilter), plus ~45 us per iteration for four
//                 launches with last-block tails instead of one persistent kernel.
// Densities come from the occupied coarse cells, so slab- or surface-like clouds are not
// mistaken for sparse ones.  The choice never changes a result.
bool grid_profitable(const cvo_b200_handle* h, const CloudDev& cs, const CloudDev& ct, float ell,
                     int n_rows = -1) {
  const cvo_b200_params& p = h->params;
  if (!p.is_using_geometry) return false;  // no geometric cut-off: every pair is a candidate
  if (h->force_mode >= 0) return h->force_mode == 1;
  if (ct.n_finite == 0 || !(ct.occupied_volume > 0.0)) return false;
  const double lmax = ((double)cs.max_dist / 500.0 + 1.0) * (double)ell;
  const double q = (double)p.sp_thres / ((double)p.sigma * (double)p.sigma);
  if (!(q > 0.0) || !(q < 1.0)) return !(q > 0.0) ? false : true;  // q >= 1: nothing survives
  const double r = lmax * std::sqrt(-2.0 * std::log(q));
  if (!(r > 0.0) || !std::isfinite(r)) return false;
  double hcell = ct.extent;
  const double hmin = ct.extent / (double)(1 << ct.cbits);  // query cells are never finer than the coarse table
  while (hcell * 0.5 >= r && hcell * 0.5 >= hmin) hcell *= 0.5;  // finest usable level with h >= r
  const double density = (double)ct.n_finite / ct.occupied_volume;
  const double tests_g = std::min((double)ct.n, 27.0 * hcell * hcell * hcell * density);
  const double rho_s = (cs.n_finite > 0 && cs.occupied_volume > 0.0) ? (double)cs.n_finite / cs.occupied_volume : density;
  const double r_tile = 0.85 * std::cbrt((double)kTileRows / rho_s);
  const double r_blk = 0.85 * std::cbrt((double)kJBlock / density);
  const double reach = r + r_tile + r_blk;
  const double tests_d = std::min((double)ct.n, density * 4.18879 * reach * reach * reach);
  const double rows = (double)(n_rows >= 0 ? n_rows : cs.n);
  const double us_grid = rows * tests_g * 2.4e-6;
  const double us_dense = rows * tests_d * 0.62e-6 + 45.0;
  return us_grid < us_dense;
}

// The third generator: tile cells (tile_kernel).  A tile of 64 Morton-consecutive rows (edge
// ~cbrt(64 / rho_s)) is swept against the cube cells covering its box inflated by the cut-off
// radius: tests_t = rho_t * 1.5 (e_tile + 2 r)^3 per row at ~0.3 ps each (packed FMA prefilter,
// measured r02), three launches per iteration.  It needs the Morton view (geometry on, isotropic
// kernel, tile-aligned shard) and rows whose candidate runs fit the per-row cell (tile_L words).
// Sets A.grid / A.tile.  The choice never changes a result.
void choose_mode(const cvo_b200_handle* h, IterArgs& A, const CloudDev& cs, const CloudDev& ct, float ell,
                 int rows_policy, bool sat_recent) {
  A.tile = 0;
  A.grid = (grid_profitable(h, cs, ct, ell, rows_policy) && (!sat_recent || h->force_mode == 1)) ? 1 : 0;
  if (h->force_mode == 0 || h->force_mode == 1) return;  // dense | grid forced
  const cvo_b200_params& p = h->params;
  if (!A.prune || !p.is_using_geometry || ct.n_finite == 0 || !(ct.occupied_volume > 0.0) || sat_recent) return;
  const double lmax = ((double)cs.max_dist / 500.0 + 1.0) * (double)ell;
  const double q = (double)p.sp_thres / ((double)p.sigma * (double)p.sigma);
  if (!(q > 0.0) || !(q < 1.0)) return;
  const double r = lmax * std::sqrt(-2.0 * std::log(q));
  if (!(r > 0.0) || !std::isfinite(r)) return;
  const double density = (double)ct.n_finite / ct.occupied_volume;
  const double rho_s = (cs.n_finite > 0 && cs.occupied_volume > 0.0) ? (double)cs.n_finite / cs.occupied_volume : density;
  const double e_tile = std::cbrt((double)kTileRows / rho_s);
  const double box = e_tile + 2.0 * r;
  const double tests_t = std::min((double)ct.n, density * 1.5 * box * box * box);
  // candidate runs a row can produce: the points of its own ball, eight per run at worst one each
  const double ball = density * 4.18879 * r * r * r;
  // rows whose candidates overflow their cells are redone exhaustively (O(M) each): keep them rare.
  // With a colour cut only a fraction of the ball becomes candidates (measured: 1 % on random
  // colours); the loop leaves the mode when rows do overflow (sat_recent).
  // (reused cells are built with the radius enlarged by up to 2 kappa: verlet_decide)
  const double skin = (h->use_persist && (h->world == 1 || h->peers_ready)) ? 1.0 + 1.5 * (double)h->verlet_kappa : 1.0;  // = `reuse` below
  const double cand_est = ball * skin * skin * skin * (p.is_using_intensity ? 0.25 : 1.0);
  if (h->force_mode != 2 && cand_est > 0.5 * (double)A.tile_L * (double)A.tile_parts) return;
  const double rows = (double)rows_policy;
  // inside the persistent kernel the cells are reused over several iterations (verlet_decide): a
  // build is amortised (measured: 4-10 iterations per build where the mode pays off) but costs
  // ~12 us of latency per iteration on average whatever its size
  const bool reuse = h->use_persist && (h->world == 1 || h->peers_ready) && h->verlet_kappa > 0.f;
  const double us_tile = reuse ? rows * tests_t * 0.3e-6 / 4.0 + 12.0 : rows * tests_t * 0.3e-6 + 30.0;
  double hcell = ct.extent;
  const double hmin = ct.extent / (double)(1 << ct.cbits);
  while (hcell * 0.5 >= r && hcell * 0.5 >= hmin) hcell *= 0.5;
  const double tests_g = std::min((double)ct.n, 27.0 * hcell * hcell * hcell * density);
  const double r_tile = 0.85 * e_tile, r_blk = 0.85 * std::cbrt((double)kJBlock / density);
  const double reach = r + r_tile + r_blk;
  const double tests_d = std::min((double)ct.n, density * 4.18879 * reach * reach * reach);
  const double us_other = A.grid ? rows * tests_g * 2.4e-6 : rows * tests_d * 0.62e-6 + 45.0;
  if (h->force_mode == 2 || us_tile < us_other) {
    A.tile = 1;
    A.grid = 0;
  }
  static const bool dbg = getenv("CVO_B200_DEBUG_POLICY") != nullptr;
  if (dbg)
    fprintf(stderr, "[policy] ell %.3f rows %d: r %.3f ball %.0f tests/row tile %.0f grid %.0f dense %.0f -> %s\n", ell,
            rows_policy, r, ball, tests_t, tests_g, tests_d, A.tile ? "tile" : (A.grid ? "grid" : "dense"));
}

// The fourth way through the flow phase, for the align loop only: a few hundred rows against a small
// target (the README demo: 523 x 1 080, rows cut at their cap for most of the registration).  No
// candidate generator pays off there - 8 lanes per row walk the whole target serially and the cut
// rows are redone anyway - so the persistent kernel gives every row a whole warp that walks ALL
// targets in the caller's order with the exact arithmetic (redo_row: the literal loop of
// CvoGPU.cu:524-591).  Called after choose_mode; the choice never changes a result.
void choose_brute(const cvo_b200_handle* h, IterArgs& A) {
  A.brute = 0;
  A.brute_cap = 0;
  if (h->force_mode != -1 && h->force_mode != 3) return;
  if (!h->use_persist || A.world > 1 || A.mode != 0 || !h->params.is_using_geometry) return;
  const bool small = A.n_rows <= 4 * h->num_sms && A.M <= 2048;  // <= 4 busy warps per SM, <= 64 passes per row
  if (h->force_mode == 3 || small) {
    A.brute = 1;
    A.grid = 1;  // launched and exported like a cell-query run (Morton column indices)
    A.tile = 0;
    // small targets: a team of warps per row (brute_team_rows) - the survivors of a team warp's part
    // of the targets wait in shared memory, at most min(row cap, part) of them
    const int part = ((A.M + kBruteTeam - 1) / kBruteTeam + 31) & ~31;
    const int slots = std::min(A.cap_max, part);
    if (A.M <= 2048 && slots > 0 && (size_t)(h->persist_threads / 32) * (size_t)slots * 8 <= (size_t)96 * 1024) {
      A.brute = 2;
      A.brute_cap = slots;
    }
  }
}

// Builds a cloud's resident representation (Morton order, SoA packing, cell table, bounding
// spheres: cvo_upload.cu) from raw arrays that are ALREADY on the device, laid out as the caller's
// (xyz n x 3, features n x F, labels n x C, geotype n x 2; null = absent).  One small read-back
// returns the scalars the host needs.
int build_cloud(cvo_b200_handle* h, CloudDev& c, int n, int F, const float* d_xyz, const float* d_feat,
                int C, const float* d_lab, const float* d_geo);

// Replaces CvoPointCloud_to_gpu (CvoGPU_impl.cu:206-285).  The caller's arrays go to the device
// as they are (four copies); everything else happens there (build_cloud).
int upload_cloud(cvo_b200_handle* h, CloudDev& c, int n, const float* xyz, int F,
                 const float* features, int C, const float* labels, const float* geotype) {
  if (n < 0 || F < 0 || C < 0 || (n > 0 && !xyz)) return fail(h, CVO_B200_ERR_INVALID, "bad cloud arguments");
  const int Fe = features ? F : 0, Ce = labels ? C : 0;
  cudaStream_t s = h->stream;
  const size_t nn = (size_t)n;
  if (n > 0) {
    CVO_CUDA(h, h->raw_xyz.ensure(nn * 3));
    CVO_CUDA(h, cudaMemcpyAsync(h->raw_xyz.p, xyz, nn * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
    if (Fe) {
      CVO_CUDA(h, h->raw_feat.ensure(nn * Fe));
      CVO_CUDA(h, cudaMemcpyAsync(h->raw_feat.p, features, nn * Fe * sizeof(float), cudaMemcpyHostToDevice, s));
    }
    if (Ce) {
      CVO_CUDA(h, h->raw_lab.ensure(nn * Ce));
      CVO_CUDA(h, cudaMemcpyAsync(h->raw_lab.p, labels, nn * Ce * sizeof(float), cudaMemcpyHostToDevice, s));
    }
    if (geotype) {
      CVO_CUDA(h, h->raw_geo.ensure(nn * 2));
      CVO_CUDA(h, cudaMemcpyAsync(h->raw_geo.p, geotype, nn * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
    }
  }
  // build_cloud ends with a stream synchronisation: the caller's arrays may be released after it
  return build_cloud(h, c, n, Fe, h->raw_xyz.p, Fe ? h->raw_feat.p : nullptr, Ce,
                     Ce ? h->raw_lab.p : nullptr, geotype ? h->raw_geo.p : nullptr);
}

int build_cloud(cvo_b200_handle* h, CloudDev& c, int n, int F, const float* d_xyz, const float* d_feat,
                int C, const float* d_lab, const float* d_geo) {
  c.n = n;
  c.F = d_feat ? F : 0;
  c.C = d_lab ? C : 0;
  c.Fp = round_up(c.F, 4);
  c.Cp = round_up(c.C, 4);
  c.has_geo = d_geo != nullptr;
  c.set = (n == 0);  // a non-empty cloud counts as set only once its build has succeeded (below)
  c.perm.clear();
  c.perm_on_host = false;
  c.max_dist = 0.f;
  c.n_finite = 0;
  h->last_valid = false;
  h->edge_valid = false;
  h->slot_key[&c == &h->tgt ? 1 : 0].valid = false;
  if (n == 0) return CVO_B200_OK;
  cudaStream_t s = h->stream;
  const size_t nn = (size_t)n;
  const bool geotype = d_geo != nullptr;
  // ---- cell table resolution: ~1 point per 16 cells on a slab-like cloud, 4..7 bits per axis
  //      (<= 8 MB), so late iterations (cut-off radius far below the cell) test a handful of points
  int cb = 4;
  while (cb < 7 && ((size_t)1 << (3 * cb)) < (size_t)16 * nn) cb++;
  c.cbits = cb;
  const int db = std::max(1, cb - 2);  // density estimate: cells with tens of points
  const size_t ncell = ((size_t)1 << (3 * cb)) + 1;
  const int nblk = (n + kJBlock - 1) / kJBlock, ntile = (n + kTileRows - 1) / kTileRows;
  const size_t temp_bytes = cloud_sort_temp_bytes(n);
  CVO_CUDA(h, h->keys_in.ensure(nn));
  CVO_CUDA(h, h->idx_in.ensure(nn));
  CVO_CUDA(h, h->sort_temp.ensure(temp_bytes));
  CVO_CUDA(h, h->d_stats.ensure(1));
  CVO_CUDA(h, c.keys_d.ensure(nn));
  CVO_CUDA(h, c.perm_d.ensure(nn));
  CVO_CUDA(h, c.inv.ensure(nn));
  CVO_CUDA(h, c.xyz.ensure(nn));
  CVO_CUDA(h, c.xyz_o.ensure(nn));
  CVO_CUDA(h, c.rowA.ensure(nn));
  if (c.Fp) { CVO_CUDA(h, c.feat.ensure(nn * c.Fp)); CVO_CUDA(h, c.feat_o.ensure(nn * c.Fp)); }
  if (c.Cp) { CVO_CUDA(h, c.lab.ensure(nn * c.Cp)); CVO_CUDA(h, c.lab_o.ensure(nn * c.Cp)); }
  if (geotype) { CVO_CUDA(h, c.geo.ensure(nn)); CVO_CUDA(h, c.geo_o.ensure(nn)); }
  CVO_CUDA(h, c.coarse.ensure(ncell));
  CVO_CUDA(h, c.blk_sphere.ensure((size_t)std::max(nblk, 1)));
  CVO_CUDA(h, c.tile_sphere.ensure((size_t)std::max(ntile, 1)));
  CVO_CUDA(h, c.tile_maxdist.ensure((size_t)std::max(ntile, 1)));
  CloudBuild B;
  std::memset(&B, 0, sizeof(B));
  B.n = n; B.F = c.F; B.C = c.C; B.Fp = c.Fp; B.Cp = c.Cp; B.cbits = cb; B.dbits = db;
  B.xyz3 = d_xyz;
  B.feat_in = c.F ? d_feat : nullptr;
  B.lab_in = c.C ? d_lab : nullptr;
  B.geo_in = d_geo;
  B.keys_in = h->keys_in.p; B.idx_in = h->idx_in.p;
  B.sort_temp = h->sort_temp.p; B.sort_temp_bytes = temp_bytes;
  B.stats = h->d_stats.p;
  B.keys = c.keys_d.p; B.perm = c.perm_d.p; B.inv = c.inv.p;
  B.xyz = c.xyz.p; B.xyz_o = c.xyz_o.p; B.rowA = c.rowA.p;
  B.feat = c.feat.p; B.feat_o = c.feat_o.p; B.lab = c.lab.p; B.lab_o = c.lab_o.p;
  B.geo = c.geo.p; B.geo_o = c.geo_o.p;
  B.coarse = c.coarse.p;
  B.blk_sphere = c.blk_sphere.p; B.tile_sphere = c.tile_sphere.p; B.tile_maxdist = c.tile_maxdist.p;
  CVO_CUDA(h, build_cloud_device(B, s));
  h->launches += 8;
  CVO_CUDA(h, cudaMemcpyAsync(h->h_stats, h->d_stats.p, sizeof(CloudStats), cudaMemcpyDeviceToHost, s));
  CVO_CUDA(h, cudaStreamSynchronize(s));
  const CloudStats& st = *h->h_stats;
  c.n_finite = st.n_finite;
  c.cx = st.centroid[0]; c.cy = st.centroid[1]; c.cz = st.centroid[2];
  c.lo[0] = st.lo[0]; c.lo[1] = st.lo[1]; c.lo[2] = st.lo[2];
  c.key_scale = (float)st.scale;
  c.extent = st.extent;
  c.max_dist = st.max_dist;
  {
    float r2;
    std::memcpy(&r2, &st.radius2_bits, sizeof(float));
    c.radius = (float)(std::sqrt((double)r2) * (1.0 + 1e-6)) + 1e-6f;  // +inf stays +inf
  }
  const double hd = st.extent / (double)(1 << db);
  c.occupied_volume = (double)st.occupied_cells * hd * hd * hd;
  c.set = true;
  return CVO_B200_OK;
}

// Morton position -> original index, fetched from the device on first use (exports only)
int fetch_perm(cvo_b200_handle* h, CloudDev& c) {
  if (c.perm_on_host) return CVO_B200_OK;
  c.perm.resize((size_t)c.n);
  if (c.n > 0)
    CVO_CUDA(h, cudaMemcpy(c.perm.data(), c.perm_d.p, sizeof(int) * (size_t)c.n, cudaMemcpyDeviceToHost));
  c.perm_on_host = true;
  return CVO_B200_OK;
}

// Runs the device-resident loop to completion.  Returns when DevState.done is set.
// The candidate generator (dense scan / cell queries) is chosen per batch of iterations from the
// polled device state; it never changes a result, only the cost of an iteration.
int run_loop(cvo_b200_handle* h, IterArgs A, int max_iter, float ell0, float* grid_fraction) {
  const int batch = 32;
  int batches = 0, grid_batches = 0;
  if (grid_fraction) *grid_fraction = 0.f;
  int rc;
  int launched_iters = 0;
  float ell = ell0;
  unsigned int sat_seen = 0;
  bool sat_recent = false;
  while (true) {
    // rows that reach their cap are redone exhaustively (O(M) each) by one block in the
    // Morton-ordered modes: leave cell queries while that happens a lot
    // multi-GPU: every rank must take the same decision (a rank in the persistent kernel and a
    // peer in the NCCL graph would wait for each other forever), so the policy only sees
    // replicated inputs there: the nominal shard size and the (replicated) length-scale
    const int rows_policy = A.world > 1 ? (A.n_src_total + A.world - 1) / A.world : A.n_rows;
    if (A.world > 1) sat_recent = false;
    choose_mode(h, A, h->src, h->tgt, ell, rows_policy, sat_recent);
    choose_brute(h, A);
    if ((A.grid || A.tile) && h->use_persist && (A.world == 1 || h->peers_ready)) {
      // the whole loop in one cooperative launch (align_grid_kernel); it returns when done.
      // world > 1: the two per-iteration exchanges are NVLink stores into the peers' mailboxes
      A.xfused = A.world > 1 ? 1 : 0;
      A.xgen = ++h->xgen;
      CVO_CUDA(h, launch_persistent(h, A));
      h->launches += 1;
      CVO_CUDA(h, cudaMemcpyAsync(h->h_poll, &h->d_state->iter, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      CVO_CUDA(h, cudaStreamSynchronize(h->stream));
      batches++;
      grid_batches++;
      if (A.stamps && h->h_poll[0] > 0) {  // CVO_B200_STAMPS=1: per-phase time of block 0 / thread 0 over the whole loop
        unsigned long long acc[10];
        cudaMemcpy(acc, A.stamps, sizeof(acc), cudaMemcpyDeviceToHost);
        const char* names[10] = {"flow rows", "flow allreduce (incl. wait)", "tile phase (when built)", "redo of cut rows", "finalize",
                                 "step rows", "step allreduce (incl. wait)", "tile hand-over", "-", "controller"};
        for (int k = 0; k < 10; k++)
          fprintf(stderr, "[align phases] %-28s %7.2f us/iter\n", names[k], (double)acc[k] / 1e3 / (double)h->h_poll[0]);
      }
      if (h->h_poll[1] /*done*/) {
        if (h->h_poll[2] /*ret*/ == CVO_B200_ERR_NCCL) h->comm_broken = true;
        if (h->h_poll[2] /*ret*/ == CVO_B200_ERR_NCCL)
          return fail(h, CVO_B200_ERR_NCCL, "fused exchange: a peer's record did not arrive (peer gone?)");
        break;
      }
      return fail(h, CVO_B200_ERR_STATE, "persistent align kernel returned without finishing");
    }
    batches++;
    grid_batches += A.grid ? 1 : 0;
    if (h->use_graph) {
      rc = ensure_graph(h, A, batch);
      if (rc != CVO_B200_OK) return rc;
      CVO_CUDA(h, cudaGraphLaunch(h->graph_exec[A.grid ? 1 : (A.tile ? 2 : 0)], h->stream));
      h->launches += (uint64_t)batch * launches_per_iteration(h, A);
    } else {
      for (int b = 0; b < batch; b++) {
        rc = enqueue_iteration(h, A, 3, nullptr, nullptr);
        if (rc != CVO_B200_OK) return rc;
      }
    }
    launched_iters += batch;
    // iter, done, ret, stop_reason | ell | sat_total
    CVO_CUDA(h, cudaMemcpyAsync(h->h_poll, &h->d_state->iter, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CVO_CUDA(h, cudaMemcpyAsync(h->h_poll + 4, &h->d_state->ell, sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CVO_CUDA(h, cudaMemcpyAsync(h->h_poll + 5, &h->d_state->sat_total, sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
    CVO_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->h_poll[1] /*done*/) break;
    std::memcpy(&ell, h->h_poll + 4, sizeof(float));
    unsigned int sat_now;
    std::memcpy(&sat_now, h->h_poll + 5, sizeof(unsigned int));
    sat_recent = (sat_now - sat_seen) > 16u * (unsigned)batch;
    sat_seen = sat_now;
    if (launched_iters > max_iter + batch) return fail(h, CVO_B200_ERR_STATE, "align loop did not terminate");
  }
  if (grid_fraction && batches > 0) *grid_fraction = (float)grid_batches / (float)batches;
  return CVO_B200_OK;
}

}  // namespace

// =============================================================================== C-ABI
extern "C" {

int cvo_b200_abi_version(void) { return CVO_B200_ABI_VERSION; }

int cvo_b200_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(cvo_b200_params);
    case 1: return (int)sizeof(cvo_b200_iter_trace);
    case 2: return (int)sizeof(cvo_b200_align_info);
    default: return -1;
  }
}

int cvo_b200_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    g_error = cudaGetErrorString(e);
    cudaGetLastError();
    return CVO_B200_ERR_CUDA;
  }
  return n;
}

const char* cvo_b200_global_error(void) { return g_error.c_str(); }

int cvo_b200_create(const cvo_b200_params* p, int device, cvo_b200_handle** out) {
  if (!p || !out) return fail(nullptr, CVO_B200_ERR_INVALID, "null argument");
  *out = nullptr;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(nullptr, CVO_B200_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
  cvo_b200_handle* h = new (std::nothrow) cvo_b200_handle();
  if (!h) return fail(nullptr, CVO_B200_ERR_NOMEM, "out of host memory");
  h->device = device;
  h->params = *p;
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
    delete h;
    return fail(nullptr, CVO_B200_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
  }
  h->num_sms = prop.multiProcessorCount;
  if (prop.major < 10) {
    delete h;
    return fail(nullptr, CVO_B200_ERR_CUDA,
                "this library is built for sm_100a (B200) only; device is sm_" +
                    std::to_string(prop.major) + std::to_string(prop.minor));
  }
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMalloc((void**)&h->d_params, sizeof(cvo_b200_params)) != cudaSuccess ||
      cudaMalloc((void**)&h->d_state, sizeof(DevState)) != cudaSuccess ||
      cudaMallocHost((void**)&h->h_poll, 16 * sizeof(int)) != cudaSuccess ||
      cudaMallocHost((void**)&h->h_stats, sizeof(CloudStats)) != cudaSuccess) {
    std::string msg = std::string("handle allocation: ") + cudaGetErrorString(cudaGetLastError());
    cvo_b200_destroy(h);
    return fail(nullptr, CVO_B200_ERR_CUDA, msg);
  }
  cudaMemcpy(h->d_params, p, sizeof(*p), cudaMemcpyHostToDevice);
  cudaMemset(h->d_state, 0, sizeof(DevState));
  const char* ng = getenv("CVO_B200_NO_GRAPH");
  h->use_graph = !(ng && ng[0] == '1');
  const char* fm = getenv("CVO_B200_MODE");  // dense | grid | tile | brute (anything else: automatic)
  if (fm && std::strcmp(fm, "dense") == 0) h->force_mode = 0;
  if (fm && std::strcmp(fm, "grid") == 0) h->force_mode = 1;
  if (fm && std::strcmp(fm, "tile") == 0) h->force_mode = 2;
  if (fm && std::strcmp(fm, "brute") == 0) h->force_mode = 3;
  const char* pe = getenv("CVO_B200_PERSIST");
  h->use_persist = !(pe && pe[0] == '0');
  if (const char* vk = getenv("CVO_B200_VERLET")) h->verlet_kappa = std::max(0.f, std::min(1.f, (float)atof(vk)));
  *out = h;
  return CVO_B200_OK;
}

void cvo_b200_destroy(cvo_b200_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  destroy_graph(h);
  // after a failed exchange the peers may never join a collective teardown: abort instead of waiting
  if (h->comm && h->comm_broken && g_nccl.CommAbort) g_nccl.CommAbort(h->comm);
  else if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  for (int r = 0; r < kMaxWorld; r++)
    if (h->peers[r] && h->peers[r] != h->mailbox) cudaIpcCloseMemHandle(h->peers[r]);
  if (h->mailbox) cudaFree(h->mailbox);
  for (CloudDev* c : {&h->src, &h->tgt}) {
    c->xyz.release(); c->rowA.release(); c->feat.release(); c->lab.release(); c->geo.release();
    c->xyz_o.release(); c->feat_o.release(); c->lab_o.release(); c->geo_o.release();
    c->blk_sphere.release(); c->tile_sphere.release(); c->tile_maxdist.release(); c->inv.release();
    c->coarse.release(); c->perm_d.release(); c->keys_d.release();
  }
  h->tgt_moved.release(); h->pq.release(); h->px.release(); h->py.release(); h->pz.release(); h->pw.release();
  h->rowrec.release(); h->row_lt.release(); h->sat_list.release(); h->flow_part2.release();
  h->ll_board.release(); h->edge_src_xyz.release(); h->edge_src_rowA.release();
  h->batch_ptr.release(); h->batch_base.release();
  h->cand.release(); h->cand_cnt.release(); h->ell_idx.release(); h->row_nnz.release();
  h->ell_val.release(); h->flow_part.release(); h->step_part.release(); h->zeros_f.release();
  h->zeros_g.release(); h->d_trace.release(); h->gathered.release(); h->stamps.release();
  h->raw_xyz.release(); h->raw_feat.release(); h->raw_lab.release(); h->raw_geo.release();
  h->keys_in.release(); h->idx_in.release(); h->sort_temp.release(); h->d_stats.release();
  for (FrameDev& f : h->frames) f.release();
  h->csr_cnt.release(); h->csr_ptr.release(); h->csr_temp.release(); h->csr_cols.release(); h->csr_vals.release();
  if (h->d_params) cudaFree(h->d_params);
  if (h->d_state) cudaFree(h->d_state);
  if (h->h_poll) cudaFreeHost(h->h_poll);
  if (h->h_stats) cudaFreeHost(h->h_stats);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int cvo_b200_write_params(cvo_b200_handle* h, const cvo_b200_params* p) {
  if (!h || !p) return fail(h, CVO_B200_ERR_INVALID, "null argument");
  cudaSetDevice(h->device);
  // The reference keeps a host copy (read by the controller) and a device copy (read by
  // the kernels) that drivers sync by hand (CvoGPU.cu:73-77); here one call updates both.
  h->params = *p;
  CVO_CUDA(h, cudaMemcpyAsync(h->d_params, &h->params, sizeof(*p), cudaMemcpyHostToDevice, h->stream));
  CVO_CUDA(h, cudaStreamSynchronize(h->stream));
  return CVO_B200_OK;
}

int cvo_b200_get_params(const cvo_b200_handle* h, cvo_b200_params* out) {
  if (!h || !out) return CVO_B200_ERR_INVALID;
  *out = h->params;
  return CVO_B200_OK;
}

const char* cvo_b200_last_error(const cvo_b200_handle* h) { return h ? h->err.c_str() : g_error.c_str(); }

int cvo_b200_set_cloud(cvo_b200_handle* h, int which, int n, const float* xyz, int F,
                       const float* features, int C, const float* labels, const float* geotype) {
  if (!h || (which != 0 && which != 1)) return fail(h, CVO_B200_ERR_INVALID, "bad handle / which");
  cudaSetDevice(h->device);
  return upload_cloud(h, which == 0 ? h->src : h->tgt, n, xyz, F, features, C, labels, geotype);
}

int cvo_b200_set_row_range(cvo_b200_handle* h, int row_begin, int row_end) {
  if (!h || row_begin < 0 || (row_end >= 0 && row_end < row_begin)) return fail(h, CVO_B200_ERR_INVALID, "bad row range");
  h->row_begin = row_begin;
  h->row_end = row_end;
  return CVO_B200_OK;
}

int cvo_b200_iterate(cvo_b200_handle* h, const float R[9], const float T[3], float ell,
                     int num_neighbors, cvo_b200_iter_trace* trace) {
  if (!h || !R || !T || !trace) return fail(h, CVO_B200_ERR_INVALID, "null argument");
  cudaSetDevice(h->device);
  IterArgs A;
  int rc = prepare(h, A, 0, nullptr);
  if (rc != CVO_B200_OK) return rc;
  if (h->src.n == 0 || h->tgt.n == 0) return fail(h, CVO_B200_ERR_STATE, "empty cloud");
  if (num_neighbors > A.cap_max) return fail(h, CVO_B200_ERR_INVALID, "num_neighbors exceeds nearest_neighbors_max");
  CVO_CUDA(h, h->d_trace.ensure(1));
  // multi-GPU: replicated inputs only, so that every rank takes the same decision
  choose_mode(h, A, h->src, h->tgt, ell, A.world > 1 ? (A.n_src_total + A.world - 1) / A.world : A.n_rows, false);
  choose_brute(h, A);
  rc = init_state(h, A, R, T, ell, num_neighbors, 0, 1, h->d_trace.p, 1);
  if (rc != CVO_B200_OK) return rc;
  if ((A.grid || A.tile) && h->use_persist && (A.world == 1 || h->peers_ready)) {
    A.xfused = A.world > 1 ? 1 : 0;
    A.xgen = ++h->xgen;
    CVO_CUDA(h, launch_persistent(h, A));
    h->launches += 1;
  } else {
    rc = enqueue_iteration(h, A, 3, nullptr, nullptr);
    if (rc != CVO_B200_OK) return rc;
  }
  CVO_CUDA(h, cudaMemcpyAsync(trace, h->d_trace.p, sizeof(*trace), cudaMemcpyDeviceToHost, h->stream));
  CVO_CUDA(h, cudaStreamSynchronize(h->stream));
  CVO_CUDA(h, cudaGetLastError());
  return CVO_B200_OK;
}

int cvo_b200_align(cvo_b200_handle* h, const float T_init[16], float T_out[16],
                   cvo_b200_align_info* info, cvo_b200_iter_trace* trace, int trace_cap) {
  if (!h || !T_init || !T_out) return fail(h, CVO_B200_ERR_INVALID, "null argument");
  cudaSetDevice(h->device);
  if (info) std::memset(info, 0, sizeof(*info));
  if (!h->src.set || !h->tgt.set) return fail(h, CVO_B200_ERR_STATE, "source/target cloud not set");
  // CvoGPU.cu:1614-1617: empty input -> return 0, output untouched
  if (h->src.n == 0 || h->tgt.n == 0) return CVO_B200_OK;
  if (h->params.is_using_kdtree) return fail(h, CVO_B200_ERR_INVALID, "is_using_kdtree is not supported (off in every shipped yaml)");
  if (h->params.indicator_window_size > kQueueCap - 2)
    return fail(h, CVO_B200_ERR_INVALID, "indicator_window_size too large");
  IterArgs A;
  int rc = prepare(h, A, 0, nullptr);
  if (rc != CVO_B200_OK) return rc;
  float R[9], T[3];
  split_pose(T_init, R, T);
  if (trace_cap < 0) trace_cap = 0;
  if (!trace) trace_cap = 0;
  if (trace_cap > 0) CVO_CUDA(h, h->d_trace.ensure((size_t)trace_cap));
  const int max_iter = h->params.MAX_ITER;
  if (max_iter <= 0) {  // loop body never runs: transform = update_tf(init)
    float Rinv[9], Tinv[3];
    host_update_tf(R, T, Rinv, Tinv);
    for (int j = 0; j < 3; j++) {
      for (int i = 0; i < 3; i++) T_out[4 * j + i] = Rinv[3 * j + i];
      T_out[4 * j + 3] = 0.f;
    }
    T_out[12] = Tinv[0]; T_out[13] = Tinv[1]; T_out[14] = Tinv[2]; T_out[15] = 1.f;
    if (info) info->stop_reason = CVO_B200_STOP_MAX_ITER;
    return CVO_B200_OK;
  }
  rc = init_state(h, A, R, T, h->params.ell_init, h->params.nearest_neighbors_max, 1, max_iter,
                  trace_cap > 0 ? h->d_trace.p : nullptr, trace_cap);
  if (rc != CVO_B200_OK) return rc;
  cudaEvent_t ev0, ev1;
  CVO_CUDA(h, cudaEventCreate(&ev0));
  CVO_CUDA(h, cudaEventCreate(&ev1));
  if (h->use_graph) {  // instantiate outside the timed region, like the reference's CvoState setup
    IterArgs Ag = A;
    choose_mode(h, Ag, h->src, h->tgt, h->params.ell_init, A.world > 1 ? (A.n_src_total + A.world - 1) / A.world : A.n_rows, false);
    choose_brute(h, Ag);
    if (!((Ag.grid || Ag.tile) && h->use_persist && (A.world == 1 || h->peers_ready))) {
      rc = ensure_graph(h, Ag, 32);
      if (rc != CVO_B200_OK) {
        cudaEventDestroy(ev0);
        cudaEventDestroy(ev1);
        return rc;
      }
    }
  }
  cudaEventRecord(ev0, h->stream);
  float grid_fraction = 0.f;
  rc = run_loop(h, A, max_iter, h->params.ell_init, &grid_fraction);
  cudaEventRecord(ev1, h->stream);
  if (rc != CVO_B200_OK) {
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    return rc;
  }
  static thread_local DevState hs;
  CVO_CUDA(h, cudaMemcpyAsync(&hs, h->d_state, sizeof(hs), cudaMemcpyDeviceToHost, h->stream));
  CVO_CUDA(h, cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ev0, ev1);
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  CVO_CUDA(h, cudaGetLastError());
  for (int j = 0; j < 3; j++) {
    for (int i = 0; i < 3; i++) T_out[4 * j + i] = hs.Rinv[3 * j + i];
    T_out[4 * j + 3] = 0.f;
  }
  T_out[12] = hs.Tinv[0]; T_out[13] = hs.Tinv[1]; T_out[14] = hs.Tinv[2]; T_out[15] = 1.f;
  const int executed = (hs.stop_reason == CVO_B200_STOP_MAX_ITER) ? hs.iter : hs.iter + 1;
  if (info) {
    info->ret = hs.ret;
    info->iterations = hs.iter;
    info->stop_reason = hs.stop_reason;
    info->final_num_neighbors = hs.num_neighbors;
    info->final_ell = hs.ell;
    info->cell_query_fraction = grid_fraction;
    info->registration_seconds = (double)ms / 1000.0;
    info->pairs_tested = (uint64_t)h->src.n * (uint64_t)h->tgt.n * (uint64_t)executed;
  }
  h->last_valid = true;
  h->last_view = hs.last_grid ? 0 : hs.last_view;
  h->last_args = A;
  h->last_tile_builds = (int)hs.tile_builds;
  if (trace_cap > 0) {
    const int nrec = std::min(trace_cap, executed);
    CVO_CUDA(h, cudaMemcpy(trace, h->d_trace.p, sizeof(cvo_b200_iter_trace) * (size_t)nrec, cudaMemcpyDeviceToHost));
  }
  return CVO_B200_OK;
}

int cvo_b200_align_host(cvo_b200_handle* h, int n_src, const float* src_xyz, int F,
                        const float* src_feat, int C, const float* src_labels,
                        const float* src_geotype, int n_tgt, const float* tgt_xyz,
                        const float* tgt_feat, const float* tgt_labels,
                        const float* tgt_geotype, const float T_init[16], float T_out[16],
                        cvo_b200_align_info* info) {
  if (!h) return CVO_B200_ERR_INVALID;
  cudaSetDevice(h->device);
  cudaEvent_t e0, e1;
  CVO_CUDA(h, cudaEventCreate(&e0));
  CVO_CUDA(h, cudaEventCreate(&e1));
  cudaEventRecord(e0, h->stream);
  int rc = upload_cloud(h, h->src, n_src, src_xyz, F, src_feat, C, src_labels, src_geotype);
  if (rc == CVO_B200_OK) rc = upload_cloud(h, h->tgt, n_tgt, tgt_xyz, F, tgt_feat, C, tgt_labels, tgt_geotype);
  cudaEventRecord(e1, h->stream);
  if (rc != CVO_B200_OK) {
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return rc;
  }
  rc = cvo_b200_align(h, T_init, T_out, info, nullptr, 0);
  float ms = 0.f;
  cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (info) info->upload_seconds = (double)ms / 1000.0;
  return rc;
}

static int inner_product_common(cvo_b200_handle* h, const float T16[16], float ell,
                                const float* kernel3x3, double* a_sum, IterArgs* A_out,
                                const CloudDev* S = nullptr, const CloudDev* Tg = nullptr,
                                int cap = -1) {
  IterArgs A;
  float kinv[9];
  int mode = 0;
  if (kernel3x3) {
    // Matrix3f::inverse() for fixed size 3: cofactors / determinant, float (CvoGPU.cu:1946)
    const float* m = kernel3x3;
#define CVO_K(i, j) m[3 * (j) + (i)]
    volatile float c00 = CVO_K(1, 1) * CVO_K(2, 2) - CVO_K(1, 2) * CVO_K(2, 1);
    volatile float c10 = CVO_K(1, 2) * CVO_K(2, 0) - CVO_K(1, 0) * CVO_K(2, 2);
    volatile float c20 = CVO_K(1, 0) * CVO_K(2, 1) - CVO_K(1, 1) * CVO_K(2, 0);
    volatile float t1 = CVO_K(0, 1) * c10, t2 = CVO_K(0, 2) * c20;
    volatile float t12 = t1 + t2;
    const float det = CVO_K(0, 0) * c00 + t12;
    const float invdet = 1.0f / det;
    kinv[0] = c00 * invdet;
    kinv[1] = c10 * invdet;
    kinv[2] = c20 * invdet;
    kinv[3] = (CVO_K(0, 2) * CVO_K(2, 1) - CVO_K(0, 1) * CVO_K(2, 2)) * invdet;
    kinv[4] = (CVO_K(0, 0) * CVO_K(2, 2) - CVO_K(0, 2) * CVO_K(2, 0)) * invdet;
    kinv[5] = (CVO_K(2, 0) * CVO_K(0, 1) - CVO_K(0, 0) * CVO_K(2, 1)) * invdet;
    kinv[6] = (CVO_K(0, 1) * CVO_K(1, 2) - CVO_K(0, 2) * CVO_K(1, 1)) * invdet;
    kinv[7] = (CVO_K(1, 0) * CVO_K(0, 2) - CVO_K(0, 0) * CVO_K(1, 2)) * invdet;
    kinv[8] = (CVO_K(0, 0) * CVO_K(1, 1) - CVO_K(1, 0) * CVO_K(0, 1)) * invdet;
#undef CVO_K
    mode = 1;
  }
  // Associations are never sharded (every row is needed where the matrix is read).  Scalar
  // inner products (inner_product_gpu / function_angle) are sharded over the ranks of a
  // multi-GPU job when the caller opted in (cvo_b200_comm_shard_inner_products): every rank
  // scans its rows, one all-gather of the ranks' sums of A - a COLLECTIVE call then.
  const bool shard = h->world > 1 && h->shard_inner_products && A_out == nullptr && h->comm != nullptr;
  int rc = prepare(h, A, mode, kernel3x3 ? kinv : nullptr, S, Tg, shard);
  if (rc != CVO_B200_OK) return rc;
  float R[9], T[3];
  split_pose(T16, R, T);
  // inner_product_impl uses num_neighbors = nearest_neighbors_max (CvoGPU.cu:1752-1754)
  rc = init_state(h, A, R, T, ell, cap >= 0 ? cap : A.cap_max, 0, 1, nullptr, 0, false);
  if (rc != CVO_B200_OK) return rc;
  rc = enqueue_iteration(h, A, 2, nullptr, nullptr);
  if (rc != CVO_B200_OK) return rc;
  double s = 0.0;
  CVO_CUDA(h, cudaMemcpyAsync(&s, &h->d_state->a_sum, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CVO_CUDA(h, cudaStreamSynchronize(h->stream));
  CVO_CUDA(h, cudaGetLastError());
  *a_sum = s;
  if (A_out) *A_out = A;
  return CVO_B200_OK;
}

int cvo_b200_inner_product(cvo_b200_handle* h, const float T[16], float ell, float* out) {
  if (!h || !T || !out) return fail(h, CVO_B200_ERR_INVALID, "null argument");
  cudaSetDevice(h->device);
  if (!h->src.set || !h->tgt.set) return fail(h, CVO_B200_ERR_STATE, "source/target cloud not set");
  if (h->src.n == 0 || h->tgt.n == 0) {
    *out = 0.f;
    return CVO_B200_OK;
  }
  double s = 0.0;
  int rc = inner_product_common(h, T, ell, nullptr, &s, nullptr);
  if (rc != CVO_B200_OK) return rc;
  *out = (float)s;
  return CVO_B200_OK;
}

int cvo_b200_function_angle(cvo_b200_handle* h, const float T[16], float ell, int is_approximate,
                            float* out) {
  if (!h || !T || !out) return fail(h, CVO_B200_ERR_INVALID, "null argument");
  cudaSetDevice(h->device);
  if (!h->src.set || !h->tgt.set) return fail(h, CVO_B200_ERR_STATE, "source/target cloud not set");
  // CvoGPU.cu:1821-1823
  if (h->src.n == 0 || h->tgt.n == 0) {
    *out = 0.f;
    return CVO_B200_OK;
  }
  float fxfz = 0.f;
  int rc = cvo_b200_inner_product(h, T, ell, &fxfz);
  if (rc != CVO_B200_OK) return rc;
  float fx_norm, fz_norm;
  if (is_approximate) {  // :1831-1833
    fx_norm = std::sqrt((double)h->src.n);
    fz_norm = std::sqrt((double)h->tgt.n);
  } else {               // :1835-1837: self inner products at the identity
    const float I16[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    double sxx = 0.0, szz = 0.0;
    rc = inner_product_common(h, I16, ell, nullptr, &sxx, nullptr, &h->src, &h->src);
    if (rc != CVO_B200_OK) return rc;
    rc = inner_product_common(h, I16, ell, nullptr, &szz, nullptr, &h->tgt, &h->tgt);
    if (rc != CVO_B200_OK) return rc;
    fx_norm = std::sqrt((float)sxx);
    fz_norm = std::sqrt((float)szz);
  }
  *out = fxfz / (fx_norm * fz_norm);
  return CVO_B200_OK;
}

// The ELL matrix of an exact-view (original target indices) run as CSR in the caller's row order.
// Device rows are in the source cloud's Morton order; un-permuting, the prefix sum and the
// compaction happen on the device (cvo_export.cu), so only row_ptr and the nnz entries are
// copied to the host.  Two-call protocol (cols/vals may be null).  Replaces
// gpu_association_to_cpu (CvoGPU_impl.cu:366-427) / copy_internal_SparseKernelMat_gpu_to_cpu.
static int export_csr(cvo_b200_handle* h, const IterArgs& A, int64_t* nnz, int32_t* max_row_nnz,
                      int32_t* row_ptr, int32_t* cols, float* vals, bool reuse_row_ptr = false,
                      const int* row_inv = nullptr, const int* col_perm = nullptr) {
  const int n_rows = A.n_rows;
  cudaStream_t s = h->stream;
  CsrExport E;
  std::memset(&E, 0, sizeof(E));
  E.n_rows = n_rows;
  E.cap_max = A.cap_max;
  E.row_nnz = A.row_nnz;
  E.ell_idx = A.ell_idx;
  E.ell_val = A.ell_val;
  E.inv = row_inv ? row_inv : h->src.inv.p;  // caller's row -> Morton position
  E.col_perm = col_perm;                      // Morton-view matrix: Morton position -> caller's column
  E.scan_temp_bytes = csr_scan_temp_bytes(n_rows);
  CVO_CUDA(h, h->csr_cnt.ensure((size_t)n_rows + 1));
  CVO_CUDA(h, h->csr_ptr.ensure((size_t)n_rows + 1));
  CVO_CUDA(h, h->csr_temp.ensure(E.scan_temp_bytes));
  E.cnt = h->csr_cnt.p;
  E.row_ptr = h->csr_ptr.p;
  E.scan_temp = h->csr_temp.p;
  // second call of the two-call protocol on the same matrix: the prefix sums are still on the
  // device (csr_ptr) and on the host (csr_rp_host)
  std::vector<int32_t>& rp_host = h->csr_rp_host;
  if (!(reuse_row_ptr && rp_host.size() == (size_t)n_rows + 1)) {
    CVO_CUDA(h, csr_row_ptr_device(E, s));
    h->launches += 2;
    rp_host.resize((size_t)n_rows + 1);
    CVO_CUDA(h, cudaMemcpyAsync(rp_host.data(), E.row_ptr, sizeof(int32_t) * ((size_t)n_rows + 1),
                                cudaMemcpyDeviceToHost, s));
    CVO_CUDA(h, cudaStreamSynchronize(s));
  }
  const int32_t* rp = rp_host.data();
  if (row_ptr) std::memcpy(row_ptr, rp, sizeof(int32_t) * ((size_t)n_rows + 1));
  const int64_t total = rp[n_rows];
  int32_t mx = 0;
  for (int i = 0; i < n_rows; i++) mx = std::max(mx, rp[i + 1] - rp[i]);
  *nnz = total;
  if (max_row_nnz) *max_row_nnz = mx;
  if (!cols || !vals || total == 0) return CVO_B200_OK;
  CVO_CUDA(h, h->csr_cols.ensure((size_t)total));
  CVO_CUDA(h, h->csr_vals.ensure((size_t)total));
  E.cols = h->csr_cols.p;
  E.vals = h->csr_vals.p;
  CVO_CUDA(h, csr_gather_device(E, s));
  h->launches += 1;
  CVO_CUDA(h, cudaMemcpyAsync(cols, E.cols, sizeof(int32_t) * (size_t)total, cudaMemcpyDeviceToHost, s));
  CVO_CUDA(h, cudaMemcpyAsync(vals, E.vals, sizeof(float) * (size_t)total, cudaMemcpyDeviceToHost, s));
  CVO_CUDA(h, cudaStreamSynchronize(s));
  return CVO_B200_OK;
}

int cvo_b200_association(cvo_b200_handle* h, const float T[16], float ell, const float* kernel3x3,
                         int64_t* nnz, int32_t* row_ptr, int32_t* cols, float* vals) {
  if (!h || !T || !nnz) return fail(h, CVO_B200_ERR_INVALID, "null argument");
  cudaSetDevice(h->device);
  if (!h->src.set || !h->tgt.set) return fail(h, CVO_B200_ERR_STATE, "source/target cloud not set");
  *nnz = 0;
  if (h->src.n == 0 || h->tgt.n == 0) return CVO_B200_OK;  // CvoGPU.cu:1884-1885
  double s = 0.0;
  IterArgs A;
  int rc = inner_product_common(h, T, ell, kernel3x3, &s, &A);
  if (rc != CVO_B200_OK) return rc;
  return export_csr(h, A, nnz, nullptr, row_ptr, cols, vals);
}

int cvo_b200_align_association(cvo_b200_handle* h, int64_t* nnz, int32_t* row_ptr, int32_t* cols,
                               float* vals) {
  if (!h || !nnz) return fail(h, CVO_B200_ERR_INVALID, "null argument");
  cudaSetDevice(h->device);
  *nnz = 0;
  if (!h->last_valid) return fail(h, CVO_B200_ERR_STATE, "no align() result is resident");
  const IterArgs& A = h->last_args;
  const int N = A.n_src_total, n_rows = A.n_rows, rb = A.row_begin;
  std::vector<uint32_t> cnt((size_t)std::max(n_rows, 1));
  CVO_CUDA(h, cudaMemcpy(cnt.data(), A.row_nnz, sizeof(uint32_t) * (size_t)n_rows, cudaMemcpyDeviceToHost));
  // device rows are Morton positions of the source cloud; this shard holds [rb, rb + n_rows)
  int rcp = fetch_perm(h, h->src);
  if (rcp == CVO_B200_OK) rcp = fetch_perm(h, h->tgt);
  if (rcp != CVO_B200_OK) return rcp;
  const std::vector<int>& sperm = h->src.perm;  // Morton position -> original row
  std::vector<int> srow((size_t)N, -1);         // original row -> local device row
  for (int s = 0; s < n_rows; s++) srow[sperm[rb + s]] = s;
  int64_t total = 0;
  if (row_ptr) row_ptr[0] = 0;
  for (int i = 0; i < N; i++) {
    if (srow[i] >= 0) total += cnt[srow[i]];
    if (row_ptr) row_ptr[i + 1] = (int32_t)total;
  }
  *nnz = total;
  if (!cols || !vals || total == 0) return CVO_B200_OK;
  std::vector<uint32_t> idx((size_t)n_rows * A.cap_max);
  std::vector<float> val((size_t)n_rows * A.cap_max);
  CVO_CUDA(h, cudaMemcpy(idx.data(), A.ell_idx, idx.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  CVO_CUDA(h, cudaMemcpy(val.data(), A.ell_val, val.size() * sizeof(float), cudaMemcpyDeviceToHost));
  const std::vector<int>& tperm = h->tgt.perm;  // Morton position -> original target index
  std::vector<std::pair<int32_t, float>> row;
  int64_t o = 0;
  for (int i = 0; i < N; i++) {
    const int s = srow[i];
    if (s < 0) continue;
    row.clear();
    for (uint32_t k = 0; k < cnt[s]; k++) {
      const uint32_t j = idx[(size_t)s * A.cap_max + k];
      row.emplace_back(h->last_view == 0 ? (int32_t)tperm[j] : (int32_t)j, val[(size_t)s * A.cap_max + k]);
    }
    std::sort(row.begin(), row.end());  // ascending target index, the reference's insertion order
    for (const auto& e : row) {
      cols[o] = e.first;
      vals[o] = e.second;
      o++;
    }
  }
  return CVO_B200_OK;
}

// ---- pose-graph edges ----------------------------------------------------
namespace {
constexpr int kMaxFrames = 1 << 16;
}

int cvo_b200_frame_set(cvo_b200_handle* h, int frame, int n, const float* xyz, int F,
                       const float* features, int C, const float* labels, const float* geotype) {
  if (!h || frame < 0 || frame >= kMaxFrames || n < 0 || F < 0 || C < 0 || (n > 0 && !xyz))
    return fail(h, CVO_B200_ERR_INVALID, "bad frame arguments");
  cudaSetDevice(h->device);
  if ((size_t)frame >= h->frames.size()) {
    h->frames.resize((size_t)frame + 1);
    h->frame_gen.resize((size_t)frame + 1, 0);
  }
  FrameDev& f = h->frames[(size_t)frame];
  f.set = false;  // stays unset if an allocation or a copy below fails
  f.n = n;
  f.F = features ? F : 0;
  f.C = labels ? C : 0;
  f.has_geo = geotype != nullptr;
  h->frame_gen[(size_t)frame] = h->frame_gen_next++;
  const size_t nn = (size_t)n;
  cudaStream_t s = h->stream;
  if (n > 0) {
    CVO_CUDA(h, f.xyz.ensure(nn * 3));
    CVO_CUDA(h, cudaMemcpyAsync(f.xyz.p, xyz, nn * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
    if (f.F) {
      CVO_CUDA(h, f.feat.ensure(nn * f.F));
      CVO_CUDA(h, cudaMemcpyAsync(f.feat.p, features, nn * f.F * sizeof(float), cudaMemcpyHostToDevice, s));
    }
    if (f.C) {
      CVO_CUDA(h, f.lab.ensure(nn * f.C));
      CVO_CUDA(h, cudaMemcpyAsync(f.lab.p, labels, nn * f.C * sizeof(float), cudaMemcpyHostToDevice, s));
    }
    if (geotype) {
      CVO_CUDA(h, f.geo.ensure(nn * 2));
      CVO_CUDA(h, cudaMemcpyAsync(f.geo.p, geotype, nn * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
    }
    CVO_CUDA(h, cudaStreamSynchronize(s));  // the caller's arrays may be released after this call
  }
  // the frame's own-frame cloud (Morton order, cell table): built once, used by every edge
  int rc = build_cloud(h, f.cloud, n, f.F, f.xyz.p, f.F ? f.feat.p : nullptr, f.C, f.C ? f.lab.p : nullptr,
                       f.has_geo ? f.geo.p : nullptr);
  if (rc != CVO_B200_OK) return rc;
  f.set = true;
  return CVO_B200_OK;
}

int cvo_b200_frame_clear(cvo_b200_handle* h, int frame) {
  if (!h || frame < -1) return fail(h, CVO_B200_ERR_INVALID, "bad frame");
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  h->edge_valid = false;
  h->slot_key[0].valid = h->slot_key[1].valid = false;
  for (size_t k = 0; k < h->frames.size(); k++)
    if (frame < 0 || (size_t)frame == k) h->frames[k].release();
  return CVO_B200_OK;
}

// frame `f` moved by its pose becomes cloud slot `c` (everything on the device)
static int build_posed_frame(cvo_b200_handle* h, CloudDev& c, int frame, const float pose[12]) {
  const FrameDev& f = h->frames[(size_t)frame];
  SlotKey& key = h->slot_key[&c == &h->tgt ? 1 : 0];
  if (key.valid && key.frame == frame && key.gen == h->frame_gen[(size_t)frame] &&
      std::memcmp(key.pose, pose, sizeof(key.pose)) == 0)
    return CVO_B200_OK;  // the slot already holds this frame at this pose
  if (f.n > 0) {
    CVO_CUDA(h, h->raw_xyz.ensure((size_t)f.n * 3));
    PoseVec P;
    std::memcpy(P.m, pose, sizeof(P.m));
    CVO_CUDA(h, pose_vec_transform_device(f.xyz.p, h->raw_xyz.p, f.n, P, h->stream));
    h->launches += 1;
  }
  int rc = build_cloud(h, c, f.n, f.F, h->raw_xyz.p, f.F ? f.feat.p : nullptr, f.C,
                       f.C ? f.lab.p : nullptr, f.has_geo ? f.geo.p : nullptr);
  if (rc != CVO_B200_OK) return rc;
  key.valid = true;
  key.frame = frame;
  key.gen = h->frame_gen[(size_t)frame];
  std::memcpy(key.pose, pose, sizeof(key.pose));
  return CVO_B200_OK;
}

// One pose-graph edge in the frames' OWN cell tables (enqueued, not waited for): frame 1's rows are
// moved by pose 1 (one small kernel), frame 2 stays where its table was built and its points are
// moved by pose 2 on the fly, exactly as transform_point_pose_vec moves them; the cell queries map
// the moved rows back with (R, T) ~ pose2^-1.  No sort, no cloud build, no O(N M) scan per edge.
// *done = false: the edge is in a regime neither cell queries nor tile cells serve (huge
// length-scale): nothing was enqueued, the caller takes the rebuild path.
static int enqueue_edge_own_frame(cvo_b200_handle* h, const FrameDev& f1, const float pose1[12], const FrameDev& f2,
                                  const float pose2[12], float ell, int num_neighbors, IterArgs& A, bool* done) {
  *done = false;
  h->cap_override = num_neighbors;
  int rc = prepare(h, A, 0, nullptr, &f1.cloud, &f2.cloud, false);
  h->cap_override = 0;
  if (rc != CVO_B200_OK) return rc;
  if (num_neighbors > A.cap_max) return fail(h, CVO_B200_ERR_INVALID, "num_neighbors exceeds the row stride");
  A.posevec = 1;
  std::memcpy(A.pose2, pose2, sizeof(A.pose2));
  // state pose = pose2^-1 up to rounding: R = R2^T (so that Rinv = R^T = R2 exactly), T = -R2^T t2
  float R[9], T[3];
  double t2n = 0.0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[3 * j + i] = pose2[4 * j + i];  // column-major R(i,j) = R2(j,i)
  for (int i = 0; i < 3; i++) {
    double acc = 0.0;
    for (int j = 0; j < 3; j++) acc -= (double)pose2[4 * j + i] * (double)pose2[4 * j + 3];
    T[i] = (float)acc;
    t2n += std::fabs((double)pose2[4 * i + 3]);
  }
  // |y - q| bound of the cell queries assumes y' = fl(R^T y - R^T T); here y' = fl(P2 [y 1]): the
  // two differ by the rounding of T and of the 4-term sums, a few ulp of |t2| + |y| (2x safety)
  A.edge_slack = (float)(4e-6 * (1.0 + t2n + (double)f2.cloud.radius + std::fabs((double)f2.cloud.cx) +
                                std::fabs((double)f2.cloud.cy) + std::fabs((double)f2.cloud.cz)));
  choose_mode(h, A, f1.cloud, f2.cloud, ell, A.n_rows, false);
  if (!(A.grid || A.tile)) return CVO_B200_OK;
  CVO_CUDA(h, h->edge_src_xyz.ensure((size_t)f1.n));
  CVO_CUDA(h, h->edge_src_rowA.ensure((size_t)f1.n));
  launch_pose_source(f1.cloud.xyz.p, f1.n, pose1, h->edge_src_xyz.p, h->edge_src_rowA.p, h->stream);
  h->launches += 1;
  A.src_xyz = h->edge_src_xyz.p;
  A.src_rowA = h->edge_src_rowA.p;
  rc = init_state(h, A, R, T, ell, num_neighbors, 0, 1, nullptr, 0, true);
  if (rc != CVO_B200_OK) return rc;
  rc = enqueue_iteration(h, A, 2, nullptr, nullptr);
  if (rc != CVO_B200_OK) return rc;
  *done = true;
  return CVO_B200_OK;
}

int cvo_b200_edge_update(cvo_b200_handle* h, int frame1, const float pose1[12], int frame2,
                         const float pose2[12], float ell, int num_neighbors, int64_t* nnz,
                         int32_t* max_row_nnz, int32_t* row_ptr, int32_t* cols, float* vals) {
  if (!h || !pose1 || !pose2 || !nnz || num_neighbors < 0)
    return fail(h, CVO_B200_ERR_INVALID, "null argument / negative num_neighbors");
  cudaSetDevice(h->device);
  if (frame1 < 0 || frame2 < 0 || (size_t)frame1 >= h->frames.size() ||
      (size_t)frame2 >= h->frames.size() || !h->frames[(size_t)frame1].set ||
      !h->frames[(size_t)frame2].set)
    return fail(h, CVO_B200_ERR_STATE, "frame not set");
  const FrameDev& f1 = h->frames[(size_t)frame1];
  const FrameDev& f2 = h->frames[(size_t)frame2];
  *nnz = 0;
  if (max_row_nnz) *max_row_nnz = 0;
  if (f1.n == 0 || f2.n == 0) {
    if (row_ptr)
      for (int i = 0; i <= f1.n; i++) row_ptr[i] = 0;
    return CVO_B200_OK;
  }
  EdgeKey key;
  std::memset(&key, 0, sizeof(key));
  key.f1 = frame1; key.f2 = frame2; key.cap = num_neighbors; key.ell = ell;
  std::memcpy(key.p1, pose1, sizeof(key.p1));
  std::memcpy(key.p2, pose2, sizeof(key.p2));
  key.gen1 = h->frame_gen[(size_t)frame1];
  key.gen2 = h->frame_gen[(size_t)frame2];
  // second call of the two-call protocol: the matrix of this very edge is still on the device
  const bool cached = h->edge_valid && std::memcmp(&key, &h->edge_key, sizeof(key)) == 0;
  static const bool rebuild_env = getenv("CVO_B200_EDGE_REBUILD") && getenv("CVO_B200_EDGE_REBUILD")[0] == '1';
  bool own_frame = !cached && !rebuild_env && h->params.is_using_geometry && !h->params.is_using_kdtree &&
                   f1.cloud.set && f2.cloud.set;
  if (own_frame) {
    IterArgs A;
    bool done = false;
    int rc = enqueue_edge_own_frame(h, f1, pose1, f2, pose2, ell, num_neighbors, A, &done);
    if (rc != CVO_B200_OK) return rc;
    if (done) {
      h->edge_key = key;
      h->edge_args = A;
      h->edge_valid = true;
      h->edge_own_frame = true;
      h->edge_row_inv = f1.cloud.inv.p;
      h->edge_col_perm = f2.cloud.perm_d.p;
      h->slot_key[0].valid = h->slot_key[1].valid = false;
    } else {
      own_frame = false;  // dense regime (huge length-scale): the rebuild path below
    }
  }
  if (h->edge_valid && h->edge_own_frame && (cached || own_frame))
    return export_csr(h, h->edge_args, nnz, max_row_nnz, row_ptr, cols, vals, cached, h->edge_row_inv,
                      h->edge_col_perm);
  if (!cached) {
    h->edge_own_frame = false;
    // a frame shared with the previous edge may sit in the other slot (ring / chain graphs):
    // exchange the slots when that saves a build
    auto holds = [&](int slot, int frame, const float* pose) {
      const SlotKey& k = h->slot_key[slot];
      return (k.valid && k.frame == frame && k.gen == h->frame_gen[(size_t)frame] &&
              std::memcmp(k.pose, pose, sizeof(k.pose)) == 0) ? 1 : 0;
    };
    if (holds(1, frame1, pose1) + holds(0, frame2, pose2) > holds(0, frame1, pose1) + holds(1, frame2, pose2)) {
      std::swap(h->src, h->tgt);
      std::swap(h->slot_key[0], h->slot_key[1]);
    }
    int rc = build_posed_frame(h, h->src, frame1, pose1);
    if (rc != CVO_B200_OK) return rc;
    rc = build_posed_frame(h, h->tgt, frame2, pose2);
    if (rc != CVO_B200_OK) return rc;
    // both clouds are already where fill_in_A_mat_gpu sees them: the pairwise pass runs at the
    // identity (1*y + (0*y + 0*y) + (-0) is exact), with the edge's own cap and a fixed ell
    const float I16[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    double s = 0.0;
    IterArgs A;
    h->cap_override = num_neighbors;
    rc = inner_product_common(h, I16, ell, nullptr, &s, &A, nullptr, nullptr, num_neighbors);
    h->cap_override = 0;
    if (rc != CVO_B200_OK) return rc;
    h->edge_key = key;
    h->edge_args = A;
    h->edge_valid = true;
  }
  return export_csr(h, h->edge_args, nnz, max_row_nnz, row_ptr, cols, vals, cached);
}

int cvo_b200_edge_update_batch(cvo_b200_handle* h, int n_edges, const cvo_b200_edge* edges, int64_t* nnz,
                               int32_t* max_row_nnz, int32_t* row_ptr, int32_t* cols, float* vals) {
  if (!h || n_edges < 0 || (n_edges > 0 && (!edges || !nnz))) return fail(h, CVO_B200_ERR_INVALID, "null argument");
  cudaSetDevice(h->device);
  if (!h->params.is_using_geometry || h->params.is_using_kdtree)
    return fail(h, CVO_B200_ERR_STATE, "batched edges need the geometric kernel (use cvo_b200_edge_update)");
  size_t rows_total = 0, worst = 0;
  std::vector<unsigned long long> gens;
  for (int e = 0; e < n_edges; e++) {
    const cvo_b200_edge& ed = edges[e];
    if (ed.frame1 < 0 || ed.frame2 < 0 || (size_t)ed.frame1 >= h->frames.size() || (size_t)ed.frame2 >= h->frames.size() ||
        !h->frames[(size_t)ed.frame1].set || !h->frames[(size_t)ed.frame2].set || ed.num_neighbors < 0)
      return fail(h, CVO_B200_ERR_STATE, "frame not set / negative num_neighbors");
    rows_total += (size_t)h->frames[(size_t)ed.frame1].n + 1;
    worst += (size_t)h->frames[(size_t)ed.frame1].n * (size_t)std::max(ed.num_neighbors, 1);
    gens.push_back(h->frame_gen[(size_t)ed.frame1]);
    gens.push_back(h->frame_gen[(size_t)ed.frame2]);
  }
  // only the SECOND call of the two-call protocol (entries wanted) may be answered from the device-side
  // copy of the first; a size query always recomputes, so repeated rounds are never served stale
  const bool cached = cols != nullptr && h->batch_valid && h->batch_edges.size() == (size_t)n_edges && gens == h->batch_gens &&
                      (n_edges == 0 || std::memcmp(h->batch_edges.data(), edges, sizeof(cvo_b200_edge) * (size_t)n_edges) == 0);
  cudaStream_t s = h->stream;
  if (!cached) {
    h->batch_valid = false;
    CVO_CUDA(h, h->batch_ptr.ensure(std::max(rows_total, (size_t)1)));
    CVO_CUDA(h, h->batch_base.ensure((size_t)n_edges + 1));
    CVO_CUDA(h, h->csr_cols.ensure(std::max(worst, (size_t)1)));
    CVO_CUDA(h, h->csr_vals.ensure(std::max(worst, (size_t)1)));
    CVO_CUDA(h, cudaMemsetAsync(h->batch_base.p, 0, sizeof(long long), s));
    size_t off = 0;
    for (int e = 0; e < n_edges; e++) {
      const cvo_b200_edge& ed = edges[e];
      const FrameDev& f1 = h->frames[(size_t)ed.frame1];
      const FrameDev& f2 = h->frames[(size_t)ed.frame2];
      if (f1.n == 0 || f2.n == 0) {  // no entries: row pointers all zero, the running offset moves on unchanged
        CVO_CUDA(h, cudaMemsetAsync(h->batch_ptr.p + off, 0, sizeof(int) * ((size_t)f1.n + 1), s));
        CVO_CUDA(h, cudaMemcpyAsync(h->batch_base.p + e + 1, h->batch_base.p + e, sizeof(long long),
                                    cudaMemcpyDeviceToDevice, s));
        off += (size_t)f1.n + 1;
        continue;
      }
      IterArgs A;
      bool done = false;
      int rc = enqueue_edge_own_frame(h, f1, ed.pose1, f2, ed.pose2, ed.ell, ed.num_neighbors, A, &done);
      if (rc != CVO_B200_OK) return rc;
      if (!done)
        return fail(h, CVO_B200_ERR_STATE, "an edge of the batch is outside the cell-query / tile regimes: use cvo_b200_edge_update");
      CsrExport E;
      std::memset(&E, 0, sizeof(E));
      E.n_rows = A.n_rows;
      E.cap_max = A.cap_max;
      E.row_nnz = A.row_nnz;
      E.ell_idx = A.ell_idx;
      E.ell_val = A.ell_val;
      E.inv = f1.cloud.inv.p;
      E.col_perm = f2.cloud.perm_d.p;
      E.scan_temp_bytes = csr_scan_temp_bytes(A.n_rows);
      CVO_CUDA(h, h->csr_cnt.ensure((size_t)A.n_rows + 1));
      CVO_CUDA(h, h->csr_temp.ensure(E.scan_temp_bytes));
      E.cnt = h->csr_cnt.p;
      E.row_ptr = h->batch_ptr.p + off;
      E.scan_temp = h->csr_temp.p;
      E.base = h->batch_base.p + e;
      E.base_next = h->batch_base.p + e + 1;
      E.cols = h->csr_cols.p;
      E.vals = h->csr_vals.p;
      CVO_CUDA(h, csr_row_ptr_device(E, s));
      CVO_CUDA(h, csr_gather_device(E, s));
      h->launches += 4;
      off += (size_t)f1.n + 1;
    }
    // ONE wait for the whole batch: the row pointers of every edge
    h->batch_rp_host.resize(std::max(rows_total, (size_t)1));
    if (rows_total)
      CVO_CUDA(h, cudaMemcpyAsync(h->batch_rp_host.data(), h->batch_ptr.p, sizeof(int) * rows_total, cudaMemcpyDeviceToHost, s));
    CVO_CUDA(h, cudaStreamSynchronize(s));
    CVO_CUDA(h, cudaGetLastError());
    h->batch_edges.assign(edges, edges + n_edges);
    h->batch_gens = gens;
    h->batch_valid = true;
    h->edge_valid = false;  // the single-edge cache shares the ELL matrix
  }
  size_t off = 0;
  int64_t total = 0;
  for (int e = 0; e < n_edges; e++) {
    const int n1 = h->frames[(size_t)edges[e].frame1].n;
    const int* rp = h->batch_rp_host.data() + off;
    int32_t mx = 0;
    for (int i = 0; i < n1; i++) mx = std::max(mx, (int32_t)(rp[i + 1] - rp[i]));
    nnz[e] = rp[n1];
    if (max_row_nnz) max_row_nnz[e] = mx;
    total += rp[n1];
    off += (size_t)n1 + 1;
  }
  if (row_ptr && rows_total) std::memcpy(row_ptr, h->batch_rp_host.data(), sizeof(int32_t) * rows_total);
  if (cols && vals && total > 0) {
    CVO_CUDA(h, cudaMemcpyAsync(cols, h->csr_cols.p, sizeof(int32_t) * (size_t)total, cudaMemcpyDeviceToHost, s));
    CVO_CUDA(h, cudaMemcpyAsync(vals, h->csr_vals.p, sizeof(float) * (size_t)total, cudaMemcpyDeviceToHost, s));
    CVO_CUDA(h, cudaStreamSynchronize(s));
  }
  return CVO_B200_OK;
}

int cvo_b200_last_candidate_builds(const cvo_b200_handle* h) { return h ? h->last_tile_builds : CVO_B200_ERR_INVALID; }

int cvo_b200_time_iterations(cvo_b200_handle* h, const float R[9], const float T[3], float ell,
                             int num_neighbors, int iters, float* ms_total, float* ms_pair_kernel) {
  if (!h || !R || !T || iters <= 0) return fail(h, CVO_B200_ERR_INVALID, "bad argument");
  cudaSetDevice(h->device);
  IterArgs A;
  int rc = prepare(h, A, 0, nullptr);
  if (rc != CVO_B200_OK) return rc;
  if (h->src.n == 0 || h->tgt.n == 0) return fail(h, CVO_B200_ERR_STATE, "empty cloud");
  if (num_neighbors > A.cap_max) num_neighbors = A.cap_max;
  // multi-GPU: replicated inputs only, so that every rank takes the same decision
  choose_mode(h, A, h->src, h->tgt, ell, A.world > 1 ? (A.n_src_total + A.world - 1) / A.world : A.n_rows, false);
  rc = init_state(h, A, R, T, ell, num_neighbors, 2, iters, nullptr, 0);
  if (rc != CVO_B200_OK) return rc;
  const bool persist = (A.grid || A.tile) && h->use_persist && (A.world == 1 || h->peers_ready);
  if (persist) {
    A.xfused = A.world > 1 ? 1 : 0;
    A.xgen = ++h->xgen;
  }
  // per-kernel events only exist when every phase is its own launch
  const int n_ev = (ms_pair_kernel && !persist) ? iters : 0;
  std::vector<cudaEvent_t> ea((size_t)n_ev), eb((size_t)n_ev);
  for (int i = 0; i < n_ev; i++) {
    cudaEventCreate(&ea[i]);
    cudaEventCreate(&eb[i]);
  }
  cudaEvent_t e0, e1;
  CVO_CUDA(h, cudaEventCreate(&e0));
  CVO_CUDA(h, cudaEventCreate(&e1));
  cudaEventRecord(e0, h->stream);
  if (persist) {  // one launch runs all `iters` iterations; no per-kernel events
    cudaError_t le = launch_persistent(h, A);
    if (le != cudaSuccess) rc = fail(h, CVO_B200_ERR_CUDA, std::string("cooperative launch: ") + cudaGetErrorString(le));
    h->launches += 1;
  } else {
    for (int i = 0; i < iters && rc == CVO_B200_OK; i++)
      rc = enqueue_iteration(h, A, 3, n_ev ? ea[i] : nullptr, n_ev ? eb[i] : nullptr);
  }
  cudaEventRecord(e1, h->stream);
  cudaError_t se = cudaStreamSynchronize(h->stream);
  float ms = 0.f, msp = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  for (int i = 0; i < n_ev; i++) {
    float t = 0.f;
    cudaEventElapsedTime(&t, ea[i], eb[i]);
    msp += t;
    cudaEventDestroy(ea[i]);
    cudaEventDestroy(eb[i]);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (rc != CVO_B200_OK) return rc;
  CVO_CUDA(h, se);
  CVO_CUDA(h, cudaGetLastError());
  if (getenv("CVO_B200_DEBUG_TAILS")) {  // %globaltimer stamps of the LAST iteration (ns)
    unsigned long long d[16];
    cudaMemcpy(d, (char*)h->d_state + offsetof(DevState, dbg), sizeof(d), cudaMemcpyDeviceToHost);
    fprintf(stderr,
            "[tails] prep_start 0 | pair_start %+lld | flow_start %+lld tail_begin %+lld reduced %+lld "
            "finalized %+lld | step_start %+lld tail_begin %+lld reduced %+lld controller_done %+lld (ns)\n",
            (long long)(d[9] - d[8]), (long long)(d[0] - d[8]), (long long)(d[1] - d[8]),
            (long long)(d[2] - d[8]), (long long)(d[3] - d[8]), (long long)(d[4] - d[8]),
            (long long)(d[5] - d[8]), (long long)(d[6] - d[8]), (long long)(d[7] - d[8]));
    fprintf(stderr, "[tails] controller: cubic done %+lld exp+pose done %+lld se3log done %+lld end %+lld (ns after reduced)\n",
            (long long)(d[13] - d[6]), (long long)(d[14] - d[6]), (long long)(d[15] - d[6]), (long long)(d[7] - d[6]));
    fprintf(stderr, "[tails] flow reduce: loads done %+lld shuffles done %+lld smem done %+lld (ns after tail_begin)\n",
            (long long)(d[10] - d[1]), (long long)(d[11] - d[1]), (long long)(d[12] - d[1]));
  }
  if (A.stamps && A.tile) {
    unsigned long long t3[3];
    cudaMemcpy(t3, A.stamps + 32000, sizeof(t3), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[tiles] %llu tile sweeps, targets swept per tile: mean %.0f max %llu (x 64 rows = pair tests)\n", t3[2],
            t3[2] ? (double)t3[0] / (double)t3[2] : 0.0, t3[1]);
    cudaMemset(A.stamps + 32000, 0, sizeof(t3));
  }
  if (A.stamps && persist) {  // inside the controller of the LAST iteration (thread 0 / thread 32 stamps)
    unsigned long long d[16];
    cudaMemcpy(d, h->d_state->dbg, sizeof(d), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[controller] cubic done %+lld | exp+pose (t0) done %+lld | se3log (t32) done %+lld | end %+lld ns after entry\n",
            (long long)(d[13] - d[12]), (long long)(d[14] - d[12]), (long long)(d[15] - d[12]), (long long)(d[11] - d[12]));
  }
  if (A.stamps && persist) {  // per-phase time of block 0 / thread 0, averaged over the iterations
    unsigned long long acc[10];
    cudaMemcpy(acc, A.stamps, sizeof(acc), cudaMemcpyDeviceToHost);
    const char* names[10] = {"flow rows", "flow allreduce (incl. wait)", "-", "redo of cut rows", "finalize", "step rows",
                             "step allreduce (incl. wait)", "-", "-", "controller"};
    for (int k = 0; k < 10; k++)
      fprintf(stderr, "[phases] %-28s %7.2f us/iter\n", names[k], (double)acc[k] / 1e3 / (double)iters);
  }
  if (A.stamps && !persist) {  // per-block phase stamps of the LAST flow launch (ns, relative to the first block)
    const int nb = A.grid ? h->grid_blocks : h->sparse_blocks;
    std::vector<unsigned long long> st8((size_t)8 * nb);
    cudaMemcpy(st8.data(), A.stamps, st8.size() * 8, cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull;
    for (int b = 0; b < nb; b++) t0 = std::min(t0, st8[8 * (size_t)b]);
    const char* names[6] = {"entry", "hot loaded", "rows done", "partial written", "fence done", "atomic done"};
    for (int k = 0; k < 6; k++) {
      std::vector<long long> v;
      for (int b = 0; b < nb; b++) v.push_back((long long)(st8[8 * (size_t)b + k] - t0));
      std::sort(v.begin(), v.end());
      fprintf(stderr, "[stamps] %-16s min %6lld  p50 %6lld  p90 %6lld  max %6lld ns\n", names[k], v[0],
              v[v.size() / 2], v[v.size() * 9 / 10], v.back());
    }
  }
  if (ms_total) *ms_total = ms;
  if (ms_pair_kernel) *ms_pair_kernel = msp;
  return CVO_B200_OK;
}

uint64_t cvo_b200_launch_count(const cvo_b200_handle* h) { return h ? h->launches : 0; }
void* cvo_b200_stream(const cvo_b200_handle* h) { return h ? (void*)h->stream : nullptr; }

int cvo_b200_fma_peak(cvo_b200_handle* h, int kind, int iters, double* fma_per_s) {
  if (!h || !fma_per_s || iters <= 0) return fail(h, CVO_B200_ERR_INVALID, "bad argument");
  cudaSetDevice(h->device);
  CVO_CUDA(h, h->zeros_f.ensure(64));
  const int blocks = h->num_sms * 8;
  launch_fma_peak(kind, iters / 8 + 1, blocks, h->zeros_f.p, h->stream);  // warm-up
  cudaEvent_t e0, e1;
  CVO_CUDA(h, cudaEventCreate(&e0));
  CVO_CUDA(h, cudaEventCreate(&e1));
  cudaEventRecord(e0, h->stream);
  launch_fma_peak(kind, iters, blocks, h->zeros_f.p, h->stream);
  cudaEventRecord(e1, h->stream);
  h->launches += 2;
  CVO_CUDA(h, cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  CVO_CUDA(h, cudaGetLastError());
  const double fmas = (double)blocks * 256.0 * (double)iters * 16.0 * (kind == 1 ? 2.0 : 1.0);
  *fma_per_s = fmas / ((double)ms * 1e-3);
  return CVO_B200_OK;
}

int cvo_b200_comm_unique_id(char id[128]) {
  std::string err;
  if (!load_nccl(err)) return fail(nullptr, CVO_B200_ERR_NCCL, err);
  UniqueId u;
  int rc = g_nccl.GetUniqueId(&u);
  if (rc != 0) return fail(nullptr, CVO_B200_ERR_NCCL, "ncclGetUniqueId failed");
  std::memcpy(id, u.internal, 128);
  return CVO_B200_OK;
}

int cvo_b200_comm_init(cvo_b200_handle* h, int rank, int world, const char id[128]) {
  if (!h || world < 1 || rank < 0 || rank >= world) return fail(h, CVO_B200_ERR_INVALID, "bad rank/world");
  cudaSetDevice(h->device);
  if (world == 1) {
    h->rank = 0;
    h->world = 1;
    return CVO_B200_OK;
  }
  std::string err;
  if (!load_nccl(err)) return fail(h, CVO_B200_ERR_NCCL, err);
  UniqueId u;
  std::memcpy(u.internal, id, 128);
  int rc = g_nccl.CommInitRank(&h->comm, world, u, rank);
  if (rc != 0)
    return fail(h, CVO_B200_ERR_NCCL,
                std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
  h->rank = rank;
  h->world = world;
  destroy_graph(h);
  return CVO_B200_OK;
}

int cvo_b200_comm_mailbox_handle(cvo_b200_handle* h, char out[64]) {
  if (!h || !out) return fail(h, CVO_B200_ERR_INVALID, "null argument");
  cudaSetDevice(h->device);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (!h->mailbox) {
    CVO_CUDA(h, cudaMalloc((void**)&h->mailbox, sizeof(XMailbox)));
    CVO_CUDA(h, cudaMemset(h->mailbox, 0, sizeof(XMailbox)));
  }
  cudaIpcMemHandle_t hd;
  CVO_CUDA(h, cudaIpcGetMemHandle(&hd, h->mailbox));
  std::memcpy(out, &hd, 64);
  return CVO_B200_OK;
}

int cvo_b200_comm_open_peers(cvo_b200_handle* h, const char* handles) {
  if (!h || !handles) return fail(h, CVO_B200_ERR_INVALID, "null argument");
  if (h->world < 2 || h->world > kMaxWorld) return fail(h, CVO_B200_ERR_STATE, "comm_init first (2..16 ranks)");
  if (!h->mailbox) return fail(h, CVO_B200_ERR_STATE, "call cvo_b200_comm_mailbox_handle first");
  cudaSetDevice(h->device);
  for (int r = 0; r < h->world; r++) {
    if (r == h->rank) {
      h->peers[r] = h->mailbox;
      continue;
    }
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, handles + 64 * (size_t)r, 64);
    void* p = nullptr;
    CVO_CUDA(h, cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    h->peers[r] = (XMailbox*)p;
  }
  h->peers_ready = true;
  h->xgen = 0;
  return CVO_B200_OK;
}

int cvo_b200_comm_shard_inner_products(cvo_b200_handle* h, int on) {
  if (!h) return CVO_B200_ERR_INVALID;
  h->shard_inner_products = on != 0;
  return CVO_B200_OK;
}

int cvo_b200_comm_destroy(cvo_b200_handle* h) {
  if (!h) return CVO_B200_ERR_INVALID;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  destroy_graph(h);
  // after a failed exchange the peers may never join a collective teardown: abort instead of waiting
  if (h->comm && h->comm_broken && g_nccl.CommAbort) g_nccl.CommAbort(h->comm);
  else if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  h->comm = nullptr;
  for (int r = 0; r < kMaxWorld; r++) {
    if (h->peers[r] && h->peers[r] != h->mailbox) cudaIpcCloseMemHandle(h->peers[r]);
    h->peers[r] = nullptr;
  }
  h->peers_ready = false;
  if (h->mailbox) cudaFree(h->mailbox);
  h->mailbox = nullptr;
  h->world = 1;
  h->rank = 0;
  return CVO_B200_OK;
}

}  // extern "C"
