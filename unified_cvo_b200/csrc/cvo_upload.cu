// cvo_upload.cu — device-side build of a cloud's resident representation.
//
// Replaces CvoPointCloud_to_gpu (CvoGPU_impl.cu:206-285): the caller's arrays are copied to the
// device as they are and everything derived from them is built there, on the handle's stream:
//   bounding cube + centroid -> 63-bit Morton keys -> radix sort (cub) -> Morton-ordered and
//   original-order SoA arrays, prefilter records, permutation and inverse -> cell table
//   (lower bounds of the cube cells at `cbits` bits per axis) -> bounding spheres of the 64-row
//   tiles / 256-target blocks -> scalar statistics for the host (one small D2H copy).
// The host never touches the points: the end-to-end call (cvo_b200_align_host) costs a few
// kernel launches per cloud instead of milliseconds of host sorting and packing.
//
// Compiled with --fmad=false like cvo_kernels.cu: rowA.w is the reference's a_to_sensor
// (CvoGPU.cu:506), evaluated uncontracted in float.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <cub/device/device_radix_sort.cuh>

#include "cvo_device.cuh"
#include "cvo_upload.cuh"

namespace cvo_b200 {

namespace {
constexpr int kStatThreads = 1024;

__device__ __forceinline__ bool finite3(float x, float y, float z) {
  return isfinite(x) && isfinite(y) && isfinite(z);
}
__device__ __forceinline__ unsigned long long spread21(unsigned long long v) {
  v &= 0x1fffffull;
  v = (v | (v << 32)) & 0x1f00000000ffffull;
  v = (v | (v << 16)) & 0x1f0000ff0000ffull;
  v = (v | (v << 8)) & 0x100f00f00f00f00full;
  v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
  v = (v | (v << 2)) & 0x1249249249249249ull;
  return v;
}

// ---- pass 1: bounding box, centroid, max |x| over the finite points.  One block, fixed
//      thread assignment and a fixed tree: deterministic.
__global__ void __launch_bounds__(kStatThreads) cloud_stats_kernel(const float* __restrict__ xyz3, int n,
                                                                   CloudStats* st) {
  __shared__ float s_lo[3][kStatThreads / 32], s_hi[3][kStatThreads / 32], s_md[kStatThreads / 32];
  __shared__ double s_sum[3][kStatThreads / 32];
  __shared__ int s_cnt[kStatThreads / 32];
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY}, md = 0.f;
  double sum[3] = {0, 0, 0};
  int cnt = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float x = xyz3[3 * (size_t)i], y = xyz3[3 * (size_t)i + 1], z = xyz3[3 * (size_t)i + 2];
    if (finite3(x, y, z)) {
      lo[0] = fminf(lo[0], x); hi[0] = fmaxf(hi[0], x);
      lo[1] = fminf(lo[1], y); hi[1] = fmaxf(hi[1], y);
      lo[2] = fminf(lo[2], z); hi[2] = fmaxf(hi[2], z);
      sum[0] += x; sum[1] += y; sum[2] += z;
      cnt++;
    }
    const float d = sqrtf(__fmaf_rn(z, z, __fmaf_rn(x, x, y * y)));  // a_to_sensor (CvoGPU.cu:506)
    if (d > md) md = d;                               // NaN compares false
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
      sum[k] += __shfl_xor_sync(0xffffffffu, sum[k], o);
    }
    md = fmaxf(md, __shfl_xor_sync(0xffffffffu, md, o));
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    for (int k = 0; k < 3; k++) { s_lo[k][w] = lo[k]; s_hi[k][w] = hi[k]; s_sum[k][w] = sum[k]; }
    s_md[w] = md;
    s_cnt[w] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < nw; i++) {
      for (int k = 0; k < 3; k++) {
        s_lo[k][0] = fminf(s_lo[k][0], s_lo[k][i]);
        s_hi[k][0] = fmaxf(s_hi[k][0], s_hi[k][i]);
        s_sum[k][0] += s_sum[k][i];
      }
      s_md[0] = fmaxf(s_md[0], s_md[i]);
      s_cnt[0] += s_cnt[i];
    }
    const int nf = s_cnt[0];
    st->n_finite = nf;
    st->max_dist = s_md[0];
    for (int k = 0; k < 3; k++) {
      st->lo[k] = nf ? s_lo[k][0] : 0.f;
      st->centroid[k] = nf ? (float)(s_sum[k][0] / (double)nf) : 0.f;
    }
    double ext = 1e-30;
    if (nf)
      for (int k = 0; k < 3; k++) ext = fmax(ext, (double)s_hi[k][0] - (double)s_lo[k][0]);
    st->extent = ext;
    st->scale = 2097151.0 / ext;  // one isotropic lattice: cells are cubes
    st->radius2_bits = 0u;
    st->occupied_cells = 0ull;
  }
}

// ---- pass 2: Morton key per point (non-finite points: ~0, i.e. last) and max |x - centroid|
__global__ void cloud_keys_kernel(const float* __restrict__ xyz3, int n, CloudStats* st,
                                  unsigned long long* __restrict__ keys, int* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float r2 = 0.f;
  if (i < n) {
    const float x = xyz3[3 * (size_t)i], y = xyz3[3 * (size_t)i + 1], z = xyz3[3 * (size_t)i + 2];
    unsigned long long key = ~0ull;
    if (st->n_finite > 0 && finite3(x, y, z)) {
      const double sc = st->scale;
      const unsigned long long qx = (unsigned long long)(((double)x - (double)st->lo[0]) * sc);
      const unsigned long long qy = (unsigned long long)(((double)y - (double)st->lo[1]) * sc);
      const unsigned long long qz = (unsigned long long)(((double)z - (double)st->lo[2]) * sc);
      key = spread21(qx) | (spread21(qy) << 1) | (spread21(qz) << 2);
    }
    keys[i] = key;
    idx[i] = i;
    const double dx = (double)x - (double)st->centroid[0], dy = (double)y - (double)st->centroid[1],
                 dz = (double)z - (double)st->centroid[2];
    r2 = __double2float_ru(dx * dx + dy * dy + dz * dz);  // NaN / inf propagate below
  }
  // max over the block, then one atomicMax on the float's bit pattern (non-negative floats order
  // like unsigned ints; NaN -> +inf: every pair stays a candidate)
  if (isnan(r2)) r2 = INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r2 = fmaxf(r2, __shfl_xor_sync(0xffffffffu, r2, o));
  if ((threadIdx.x & 31) == 0 && r2 > 0.f) atomicMax(&st->radius2_bits, __float_as_uint(r2));
}

// ---- pass 3 (after the sort): Morton-ordered and original-order SoA arrays
struct GatherArgs {
  int n, F, C, Fp, Cp;
  const float* xyz3; const float* feat_in; const float* lab_in; const float* geo_in;
  const int* perm;                    // Morton position -> original index (sorted values)
  const unsigned long long* keys;     // sorted
  const CloudStats* st;
  float4* xyz; float4* xyz_o; float4* rowA;
  float* feat; float* feat_o; float* lab; float* lab_o;
  float2* geo; float2* geo_o;
  int* inv;
};
__global__ void cloud_gather_kernel(GatherArgs G) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= G.n) return;
  const int i = G.perm[s];
  G.inv[i] = s;
  const float x = G.xyz3[3 * (size_t)i], y = G.xyz3[3 * (size_t)i + 1], z = G.xyz3[3 * (size_t)i + 2];
  // .w of a point = its first four feature channels quantised to 8 bits each (clamped to [0,1]):
  // a 4-byte summary from which eval_pair derives a LOWER bound of the colour distance and
  // rejects most colour mismatches without touching the 32-byte feature rows
  auto pack_colour = [&](int idx) -> float {
    unsigned int q = 0u;
    for (int k = 0; k < 4 && k < G.F; k++) {
      const float f = G.feat_in[(size_t)idx * G.F + k];
      const float c = fminf(fmaxf(f, 0.f), 1.f);  // NaN -> 0; the full test decides then
      q |= (unsigned int)(c * 255.f) << (8 * k);
    }
    return __uint_as_float(q);
  };
  G.xyz[s] = make_float4(x, y, z, pack_colour(i));
  G.xyz_o[s] = make_float4(G.xyz3[3 * (size_t)s], G.xyz3[3 * (size_t)s + 1], G.xyz3[3 * (size_t)s + 2],
                           pack_colour(s));
  // CvoGPU.cu:506 as the reference's GPU build contracts it (see eval_pair, cvo_kernels.cu)
  const float dist = sqrtf(__fmaf_rn(z, z, __fmaf_rn(x, x, y * y)));
  const float cx = G.st->centroid[0], cy = G.st->centroid[1], cz = G.st->centroid[2];
  // prefilter record: a = -2 (x - c) and the reference's a_to_sensor
  G.rowA[s] = make_float4(-2.f * (x - cx), -2.f * (y - cy), -2.f * (z - cz), dist);
  for (int k = 0; k < G.Fp; k++) {
    G.feat[(size_t)s * G.Fp + k] = (k < G.F) ? G.feat_in[(size_t)i * G.F + k] : 0.f;
    G.feat_o[(size_t)s * G.Fp + k] = (k < G.F) ? G.feat_in[(size_t)s * G.F + k] : 0.f;
  }
  for (int k = 0; k < G.Cp; k++) {
    G.lab[(size_t)s * G.Cp + k] = (k < G.C) ? G.lab_in[(size_t)i * G.C + k] : 0.f;
    G.lab_o[(size_t)s * G.Cp + k] = (k < G.C) ? G.lab_in[(size_t)s * G.C + k] : 0.f;
  }
  if (G.geo_in) {
    G.geo[s] = make_float2(G.geo_in[2 * (size_t)i], G.geo_in[2 * (size_t)i + 1]);
    G.geo_o[s] = make_float2(G.geo_in[2 * (size_t)s], G.geo_in[2 * (size_t)s + 1]);
  }
}

// ---- pass 4: cell table (first Morton position of every cube cell at cbits bits per axis) and
//      the number of occupied cells two levels up (density estimate of the mode policy)
__global__ void cell_table_kernel(const unsigned long long* __restrict__ keys, int n, int cbits,
                                  CloudStats* st, uint32_t* __restrict__ coarse) {
  const unsigned long long ncell = 1ull << (3 * cbits);
  const unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c > ncell) return;
  const int nf = st->n_finite;
  if (c == ncell) {
    coarse[c] = (uint32_t)nf;
    return;
  }
  const unsigned long long k = c << (3 * (21 - cbits));
  int lo = 0, hi = nf;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (keys[mid] < k) lo = mid + 1; else hi = mid;
  }
  coarse[c] = (uint32_t)lo;
}
__global__ void occupied_cells_kernel(const unsigned long long* __restrict__ keys, int dbits, CloudStats* st) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const int nf = st->n_finite;
  int flag = 0;
  if (s < nf) {
    const int dsh = 3 * (21 - dbits);
    flag = (s == 0) || ((keys[s] >> dsh) != (keys[s - 1] >> dsh));
  }
  const unsigned b = __ballot_sync(0xffffffffu, flag);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(&st->occupied_cells, (unsigned long long)__popc(b));
}

// ---- pass 5: bounding spheres over the Morton order, one warp per group of `group` points:
//      centre = centre of the bounding box, radius = max distance to it (rounded up); a group
//      with a non-finite point gets an infinite sphere (never skipped)
__global__ void spheres_kernel(const float4* __restrict__ xyz, const float4* __restrict__ rowA, int n,
                               int group, float4* __restrict__ out, float* __restrict__ maxdist) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int ng = (n + group - 1) / group;
  if (g >= ng) return;
  const int b = g * group, e = min(n, b + group);
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY}, md = 0.f;
  bool fin = true;
  for (int s = b + lane; s < e; s += 32) {
    const float4 q = xyz[s];
    fin = fin && finite3(q.x, q.y, q.z);
    lo[0] = fminf(lo[0], q.x); hi[0] = fmaxf(hi[0], q.x);
    lo[1] = fminf(lo[1], q.y); hi[1] = fmaxf(hi[1], q.y);
    lo[2] = fminf(lo[2], q.z); hi[2] = fmaxf(hi[2], q.z);
    md = fmaxf(md, rowA[s].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
    }
    md = fmaxf(md, __shfl_xor_sync(0xffffffffu, md, o));
  }
  fin = __all_sync(0xffffffffu, fin);
  const double cx = 0.5 * ((double)lo[0] + (double)hi[0]), cy = 0.5 * ((double)lo[1] + (double)hi[1]),
               cz = 0.5 * ((double)lo[2] + (double)hi[2]);
  double r2 = 0.0;
  for (int s = b + lane; s < e; s += 32) {
    const float4 q = xyz[s];
    const double dx = q.x - cx, dy = q.y - cy, dz = q.z - cz;
    r2 = fmax(r2, dx * dx + dy * dy + dz * dz);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r2 = fmax(r2, __shfl_xor_sync(0xffffffffu, r2, o));
  if (lane == 0) {
    if (fin) {
      out[g] = make_float4((float)cx, (float)cy, (float)cz,
                           (float)(sqrt(r2) * (1.0 + 1e-5)) +
                               1e-5f * (float)(fabs(cx) + fabs(cy) + fabs(cz)) + 1e-6f);
      if (maxdist) maxdist[g] = md;
    } else {
      out[g] = make_float4(0.f, 0.f, 0.f, INFINITY);
      if (maxdist) maxdist[g] = INFINITY;
    }
  }
}

// CvoFrameGPU::transform_pointcloud -> transform_point_pose_vec (CvoFrameGPU.cu:44-62,
// CvoGPU_impl.cu:84-150): x' = P [x y z 1]^T with P a row-major 3x4 float pose.  Eigen's unrolled
// 4-term redux sums (c0 + c1) + (c2 + c3); intrinsics keep it uncontracted whatever the flags.
__global__ void pose_vec_transform_kernel(const float* __restrict__ in, float* __restrict__ out, int n,
                                          PoseVec P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = in[3 * (size_t)i], y = in[3 * (size_t)i + 1], z = in[3 * (size_t)i + 2];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    const float c0 = __fmul_rn(P.m[4 * r], x), c1 = __fmul_rn(P.m[4 * r + 1], y);
    const float c2 = __fmul_rn(P.m[4 * r + 2], z), c3 = __fmul_rn(P.m[4 * r + 3], 1.0f);
    out[3 * (size_t)i + r] = __fadd_rn(__fadd_rn(c0, c1), __fadd_rn(c2, c3));
  }
}
}  // namespace

cudaError_t pose_vec_transform_device(const float* xyz3_in, float* xyz3_out, int n, const PoseVec& P,
                                      cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  pose_vec_transform_kernel<<<(n + 255) / 256, 256, 0, s>>>(xyz3_in, xyz3_out, n, P);
  return cudaGetLastError();
}

size_t cloud_sort_temp_bytes(int n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long*)nullptr,
                                  (unsigned long long*)nullptr, (const int*)nullptr, (int*)nullptr, n);
  return bytes;
}

cudaError_t build_cloud_device(const CloudBuild& B, cudaStream_t s) {
  const int n = B.n;
  cloud_stats_kernel<<<1, kStatThreads, 0, s>>>(B.xyz3, n, B.stats);
  const int tb = 256, nb = (n + tb - 1) / tb;
  cloud_keys_kernel<<<nb, tb, 0, s>>>(B.xyz3, n, B.stats, B.keys_in, B.idx_in);
  size_t bytes = B.sort_temp_bytes;
  // Only the bits the cell table resolves are sorted: 7 bits per axis (bits 42..62; cell queries
  // never use cells finer than the coarse table, cbits <= 7) plus bit 63, which only the ~0 key of
  // a non-finite point has, so those still sort strictly last.  The sort is stable, so ties keep
  // the caller's order.  3 radix passes instead of 8: a third of the build of a KITTI-sized cloud
  // (profiles/r01h_edge_loop_summary.txt).
  constexpr int kSortBeginBit = 3 * (21 - 7);
  cudaError_t e = cub::DeviceRadixSort::SortPairs(B.sort_temp, bytes, B.keys_in, B.keys, B.idx_in, B.perm, n,
                                                  kSortBeginBit, 64, s);
  if (e != cudaSuccess) return e;
  GatherArgs G;
  G.n = n; G.F = B.F; G.C = B.C; G.Fp = B.Fp; G.Cp = B.Cp;
  G.xyz3 = B.xyz3; G.feat_in = B.feat_in; G.lab_in = B.lab_in; G.geo_in = B.geo_in;
  G.perm = B.perm; G.keys = B.keys; G.st = B.stats;
  G.xyz = B.xyz; G.xyz_o = B.xyz_o; G.rowA = B.rowA;
  G.feat = B.feat; G.feat_o = B.feat_o; G.lab = B.lab; G.lab_o = B.lab_o;
  G.geo = B.geo; G.geo_o = B.geo_o; G.inv = B.inv;
  cloud_gather_kernel<<<nb, tb, 0, s>>>(G);
  const unsigned long long ncell = (1ull << (3 * B.cbits)) + 1ull;
  cell_table_kernel<<<(unsigned)((ncell + tb - 1) / tb), tb, 0, s>>>(B.keys, n, B.cbits, B.stats, B.coarse);
  occupied_cells_kernel<<<nb, tb, 0, s>>>(B.keys, B.dbits, B.stats);
  const int nblk = (n + kJBlock - 1) / kJBlock, ntile = (n + kTileRows - 1) / kTileRows;
  spheres_kernel<<<(nblk * 32 + tb - 1) / tb, tb, 0, s>>>(B.xyz, B.rowA, n, kJBlock, B.blk_sphere, nullptr);
  spheres_kernel<<<(ntile * 32 + tb - 1) / tb, tb, 0, s>>>(B.xyz, B.rowA, n, kTileRows, B.tile_sphere,
                                                            B.tile_maxdist);
  return cudaGetLastError();
}

}  // namespace cvo_b200
