// cvo_device.cuh — device-side state and launch arguments of the CVO hot path.
//
// Data layout in HBM (all SoA; the reference's 192-byte AoS CvoPoint,
// utils/PointSegmentedDistribution.hpp:147-229, is NOT used on the device):
//   Both clouds are stored in MORTON (Z-curve) order of their own coordinates.  Source rows are
//   always used in that order (row order never matters: every accumulated quantity is a sum or
//   max over rows).  The target has two views: view 0 = Morton order (spatially compact 256-target
//   blocks that whole source tiles can skip), view 1 = the caller's original order, needed
//   whenever a row reaches its cap, because the reference keeps the FIRST num_neighbors survivors
//   in ORIGINAL target order (CvoGPU.cu:524-526).
//   source cloud (the rows of the kernel matrix, N points)
//     src_xyz   float4[N]      exact coordinates (x,y,z,0)
//     src_rowA  float4[N]      prefilter record (-2(x-c), dist_to_sensor); c = source centroid
//     src_feat  float [N*Fp]   row-major, Fp = F rounded up to 4, zero padded
//     src_lab   float [N*Cp]   row-major label distributions
//     src_geo   float2[N]      geometric type
//   target cloud (the moving cloud, M points): tgt_xyz / tgt_feat / tgt_lab / tgt_geo
//     plus, rewritten every iteration by prep_kernel:
//     tgt_moved float4[M]      y' = Rinv*y + Tinv, exact reference arithmetic
//     px,py,pz,pw float[M]     y'-c and |y'-c|^2: the streamed operand of the pair kernel
//   candidate cells  uint32[N * nchunks * L] + uint32 counts[N * nchunks]
//   kernel matrix    ELL: ell_idx uint32[N*cap_max], ell_val float[N*cap_max], row_nnz[N]
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/cvo_b200.h"

namespace cvo_b200 {

constexpr int kTileRows = 64;          // source rows staged per warp tile
constexpr int kJQ = 8;                 // targets held per lane
constexpr int kJBlock = 32 * kJQ;      // targets per warp sweep
constexpr int kPairWarps = 8;          // warps per CTA of the pair kernel
constexpr int kSparseThreads = 256;    // CTA size of the sparse kernels
// CTA sizes of the persistent kernel: ONE block per SM, so the all-to-all reduction after a grid
// barrier reads 148 partials.  The kernel is latency bound and register hungry, so the block is
// as small as one pass over the source rows allows (4 rows per warp):
//   18 warps, 96 registers   up to 10 656 rows   (C2: 28.9 -> 25.0 us per iteration vs 24 warps)
//   24 warps, 80 registers   up to 14 208 rows
//   28 warps, 72 registers   beyond              (KITTI-sized 16 384 rows: one pass instead of
//                                                 two, 32.8 -> 28.5 us per iteration)
constexpr int kPersistThreadsSmall = 576;  // 18 warps, up to 112 registers: no spilling; covers 10 656 rows
constexpr int kPersistThreads = 768;
constexpr int kPersistThreadsWide = 896;
constexpr int kQueueCap = 1024;        // max indicator_window_size supported

// per-block partial sums of the flow pass: omega[3], v[3], a_sum, nnz, max_row (all as doubles:
// nnz and max_row are exact integers < 2^53)
struct FlowPartial {
  double v[9];
};
// The scalar prologue of fill_in_A_mat_gpu (CvoGPU.cu:495-515), evaluated once per
// write_params on the host (same float arithmetic) instead of once per thread.
struct KernConsts {
  float sigma2, c2, c_sigma2, s_ell, s_sigma2, s_ell_square, sp_thres;
  float log_geo;      // logf(sp_thres / sigma2)
  float d2_c_thres, d2_s_thres, d2_s_thres_dense;
  int use_geo_type, use_geometry, use_intensity, use_semantics;
  int use_range_ell;   // CvoParams::is_using_range_ell (step kernel)
  float c_div, d_div;  // CvoParams::c, d: divisors of the per-row flow (CvoGPU.cu:785-788)
};
struct StepPartial {
  double b, c, d, e;
};

// Device-resident controller state: everything align_impl keeps in host
// variables (CvoGPU.cu:1363-1386) lives here so the loop needs no host round trip.
// Layout: the first kHot1Bytes are what every block of every kernel reads at its start (one
// cooperative load into shared memory = one L2 round trip instead of a chain of dependent
// loads); the next block is the step kernel's input.  The host polls iter..stop_reason.
struct DevState {
  // ---- hot block 1: pose, schedule, constants
  float R[9], T[3];         // current T_target_to_source blocks (column-major R)
  float Rinv[9], Tinv[3];   // update_tf output; Rinv/Tinv are also the final transform
  float ell;
  int num_neighbors;
  int iter;
  int done;
  int ret;
  int stop_reason;
  int max_iter;
  int controller_on;        // 1: align loop; 0: single iterate() call (queues/ell/cap untouched);
                            // 2: fixed-state timing loop (pose restored, runs to max_iter)
  float ymax2_bound;        // upper bound of max_j |y'_j - c|^2 for the CURRENT Rinv/Tinv
  float smax;               // upper bound of sigma_max(Rinv) (scales block radii)
  float grid_slack;         // absolute slack [m] of a cell query in the target's own frame
  // candidate-cell reuse (persistent tile mode): the cells are built with every row's query radius
  // enlarged by vl_str + vl_srot |x| and stay valid while the pose has drifted less than that
  float vl_str, vl_srot;    // translation budget [m], rotation budget [Frobenius norm of R - R_build]
  int prune_on;             // this run may use the Morton view
  int view;                 // target view of the CURRENT iteration: 0 Morton (pruned), 1 original
  unsigned int sat_base;    // persistent kernel: value of the (monotone) n_sat counter at the start of this iteration
  unsigned int work_base;   // persistent kernel: likewise for the (monotone) work_counter
  KernConsts kc;
  // ---- hot block 2: flow result = step kernel input
  float omega[3], v[3];
  float W2[9], W3[9], W4[9], Wv[3], W2v[3], W3v[3];  // omega_hat powers (CvoGPU.cu:970-980)
  // ---- the rest
  double omega_sum[3], v_sum[3];
  double a_sum;
  unsigned long long nnz;
  unsigned int max_row_nnz;
  int last_grid;            // the LAST executed iteration used cell queries (Morton index space)
  int last_view;            // view the LAST executed iteration used (its ELL matrix is in that index space)
  // candidate-cell reuse: the state the current cells were built at
  float vl_R[9], vl_T[3];   // R, T (source -> target frame)
  float vl_ls;              // ell * smax
  float vl_slack;           // grid_slack
  int vl_valid;             // cells exist for this launch
  int tile_rebuild;         // decision of the current iteration (same in every block)
  unsigned int tile_builds; // how many iterations built cells (statistics)
  unsigned int tile_item_base, tile_item_next;  // value of the monotone item counter at the start of the current / next build
  // step
  double B, C, D, E;
  float step;
  double dist;
  // indicator queues (CvoGPU.cu:1377-1380)
  float q_start[kQueueCap];
  float q_end[kQueueCap];
  int qs_head, qs_size, qe_head, qe_size;
  float start_sum, end_sum;
  // scheduling scratch
  unsigned int work_counter;   // dynamic hand-out of row groups (flow_rows)
  unsigned int item_counter;   // dynamic hand-out of (tile, chunk) items (pair_kernel)
  unsigned int flow_blocks_done;
  unsigned int step_blocks_done;
  unsigned int n_sat;       // view 0: rows that reached their cap (redone exactly by the flow tail)
  unsigned int n_capped;    // view 1: rows that reached their cap (keeps the run on view 1)
  unsigned int sat_total;   // running sum of n_sat + n_capped over the iterations (host policy)
  unsigned int bar_count;   // grid barrier of the persistent kernel (monotone arrival counter)
  // trace
  cvo_b200_iter_trace* trace;
  int trace_cap;
  int pad0;
  // multi-GPU: totals of the LOCAL row shard, written before the collective
  double local_flow[9];     // omega[3], v[3], a_sum, (double)nnz, (double)max_row_nnz
  double local_step[4];
  unsigned long long dbg[16];  // %globaltimer stamps of the tails (tools/gpu_tails.py)
  // fused multi-GPU exchange: block 0 gathers the ranks' records and hands the job totals to the
  // other blocks of its GPU here ([phase][parity], same self-validating words; tag ~0 = a peer
  // timed out)
  unsigned long long xll[2][2][32];
};
constexpr int kHot1Words = (int)(offsetof(DevState, omega) / 4);
constexpr int kHot2Words = (int)((offsetof(DevState, omega_sum) - offsetof(DevState, omega)) / 4);


// Cell index of the target cloud in its OWN frame (built once per upload; a rigid motion of the
// target does not invalidate it: source rows are mapped INTO that frame to query it).
// The Morton-ordered target (view 0) is a linear octree: every cube cell of every level is a
// contiguous range of the sorted 63-bit keys.
struct GridView {
  const uint32_t* coarse;          // [2^(3*cbits) + 1] first Morton position of each coarse cell
                                   // (non-finite points sort last and belong to no cell)
  int cbits;                       // bits per axis of the coarse table
  int n_finite;                    // points with a finite key
  float lo[3];                     // origin of the key lattice
  float scale;                     // lattice units per metre ((2^21 - 1) / extent)
};

// Fused multi-GPU exchange (persistent kernel, one process per GPU): every rank owns a mailbox in
// its HBM that its peers map through CUDA IPC.  A rank all-gathers a small record by STORING it
// into every peer's mailbox over NVLink and polling its own mailbox - no NCCL call, no kernel
// boundary, and no memory fence either: every double travels as two self-validating 8-byte
// words (tag << 32 | half), so a word is either the old one or completely the new one and the
// receiver simply polls until both tags match (the LL protocol of collective libraries).
// Slots are indexed by (phase, iteration parity, source rank); a rank cannot run two iterations
// ahead of a peer (it needs the peer's records to advance), so two parities suffice.
constexpr int kMaxWorld = 16;
struct XMailbox {
  unsigned long long ll[2][2][kMaxWorld][32];  // [phase][parity][rank][2 words per value]
};

// Barrier-free all-reduce board of the persistent kernel: every block PUBLISHES its per-block
// values as self-validating 8-byte words (tag << 32 | half of a double, the LL protocol again) and
// every block POLLS all blocks' words - arrival is data validity, so one reduction costs one L2
// store + one L2 load round trip instead of fence + atomic + spin + fence + reload.  The board is
// zeroed before every launch; tags are the reduction's sequence number (>= 1) inside the launch;
// two parities suffice (a block cannot run two reductions ahead of another one).
constexpr int kBruteTeam = 4;     // warps per source row of the team walk (IterArgs::brute == 2)
constexpr int kLLMaxBlocks = 160;  // >= the SM count of the part (B200: 148)
constexpr int kLLValues = 11;
struct LLBoard {
  unsigned long long w[2][kLLMaxBlocks][2 * kLLValues];
};

struct TargetView {
  const float4* xyz;
  const float* feat;
  const float* lab;
  const float2* geo;
};

struct IterArgs {
  const cvo_b200_params* params;  // device copy (the reference's params_gpu)
  DevState* st;
  // source shard
  const float4* src_xyz;
  const float4* src_rowA;
  const float* src_feat;
  const float* src_lab;
  const float2* src_geo;
  int row_begin;   // first global source row of this shard
  int n_rows;      // rows in this shard
  int n_src_total; // N (all shards) — the indicator uses N*M
  // target: view 0 = Morton order (prunable), view 1 = original order (exact row cap)
  TargetView tv[2];
  int prune;       // view 0 may skip (source tile, target block) pairs by bounding spheres
  const int* tgt_inv;         // original target index -> Morton position
  uint32_t* sat_list;         // [n_rows] rows of the Morton view that reached their cap
  const float4* blk_sphere;   // [M_pad/256] view-0 target blocks: centre xyz, radius (own frame)
  const float4* tile_sphere;  // [ceil(N/64)] source tiles: centre xyz, radius
  const float* tile_maxdist;  // [ceil(N/64)] max distance-to-sensor inside the tile
  float4* tgt_moved;
  float* px; float* py; float* pz; float* pw;
  uint32_t* pq;    // [M_pad] packed colour summaries of the targets (view order), for the emission path
  int M;
  int Fp, Cp;      // padded feature / class dims (same for both clouds)
  float cx, cy, cz;        // prefilter centre (source centroid)
  float tcx, tcy, tcz;     // target centroid and radius max_j |y_j - tc| (static per cloud)
  float trad;
  int M_pad;               // M rounded up to kJBlock: px..pw are padded with (0,0,0,+inf)
  float4* rowrec;          // [n_rows][2]: (ax,ax,ay,ay) (az,az,t,t), rewritten every iteration
  float2* row_lt;          // [n_rows]: (l_i, d2_thres_i) of the exact test
  // candidate cells
  uint32_t* cand;
  uint32_t* cand_cnt;
  int nchunks, chunk_len, L;
  // ELL kernel matrix
  uint32_t* ell_idx;
  float* ell_val;
  uint32_t* row_nnz;
  int cap_max;
  // partial sums
  FlowPartial* flow_part;
  FlowPartial* flow_part2;  // persistent kernel: partials of the exact redo of cut rows
  StepPartial* step_part;
  // kernel variant
  int mode;        // 0 isotropic (fill_in_A_mat_gpu), 1 Mahalanobis (.._dense_mat_kernel)
  float kinv[9];   // column-major inverse kernel for mode 1
  LLBoard* ll;     // persistent kernel: barrier-free reduction board (zeroed before every launch)
  unsigned long long* stamps;  // debug (CVO_B200_STAMPS=1): [blocks][8] %globaltimer of block phases
  // fused multi-GPU exchange (xfused = 1): peers' mailboxes (own included), this launch's generation
  XMailbox* xpeer[kMaxWorld];
  unsigned long long xgen;
  int xfused, xrank, xworld;
  int colour;      // is_using_intensity (selects the persistent kernel with the stage-1 colour cut)
  int grid;        // 1: candidates come from cell queries (flow_kernel_t<1>; no prep/pair launch)
  int tile;        // 1: candidates come from tile_kernel (source tile x the octree cells of the target
                   //    its rows can reach; Morton view only; no prep launch)
  int tile_L;      // tile mode: candidate words per (row, part) cell
  int tile_parts;  // tile mode: warps sharing one tile (each sweeps every tile_parts-th cell group)
  int tile_cut_bits;  // tile mode: a tile is cut where neighbouring rows' Morton keys differ above this bit
  const unsigned long long* src_keys;  // Morton keys of the source rows (sorted; source's own lattice)
  GridView gv;
  // pose-graph edges evaluated in the frames' OWN cell tables (no posed cloud is rebuilt): the
  // source rows are the posed copy of frame 1, the target is moved on the fly by frame 2's pose
  // with the arithmetic of transform_point_pose_vec (CvoGPU_impl.cu:84-150) instead of R, T
  int posevec;
  float pose2[12];   // row-major 3x4
  float edge_slack;  // extra slack of the cell queries: the state's (R, T) only approximates pose2^-1
  float verlet_kappa;  // candidate-cell reuse: skin as a fraction of the cut-off radius (0 = rebuild every iteration)
  float src_rmax;      // max |x| over the source (scales the rotation budget)
  int world;       // >1: tails only publish local totals, finalize kernels run after the collective
  int n_items;     // pair-kernel work items = row_tiles * nchunks
  int brute;       // persistent kernel: no candidate generator - one warp per row walks ALL targets in the caller's
                   // order with the exact arithmetic (redo_row); a handful of rows against a small target.
                   // 2: a TEAM of kBruteTeam warps per row (brute_team_rows), each warp a contiguous part of the targets
  int brute_cap;   // brute == 2: survivor slots per warp buffer (>= min(row cap, targets per team warp))
  int row_spread;  // few rows: 4-row groups are dealt warp-major (group k -> block k % blocks, warp k / blocks),
                   // so that every SM gets a few busy warps instead of the first blocks getting them all
};

}  // namespace cvo_b200
