// cvo_kernels.cu — sm_100a kernels of one CVO iteration.
//
// Per iteration (reference call stack: align_impl, CvoGPU.cu:1387-1533):
//   prep_kernel      update_tf + transform_pointcloud_thrust (CvoGPU.cu:94-112,
//                    CvoGPU_impl.cu:31-82,164-173): y' = Rinv*y + Tinv, the centred SoA
//                    operand of the pair kernel, and the per-row prefilter records.
//   pair_kernel      the dense N x M part of fill_in_A_mat_gpu (CvoGPU.cu:477-593):
//                    a conservative fp32 prefilter |x-y'|^2 < thres_i on packed
//                    f32x2 FMAs that emits, per (row, target chunk), the ORDERED
//                    list of candidate targets.  Source tiles are TMA-staged into
//                    shared memory (double buffered), targets are streamed into registers.
//   flow_kernel      the exact per-pair arithmetic of fill_in_A_mat_gpu on the
//                    candidates (row cap = first num_neighbors survivors in target
//                    order), the ELL kernel matrix, compute_flow_gpu_no_eigen
//                    (:729-790) and the flow reduction + normalisation (:824-838).
//   step_kernel      compute_step_size_xi + compute_step_size_poly_coeff (:953-1082),
//                    the B..E reduction, cubic, clamp (:1118-1158) and, in its last
//                    block, the whole controller of align_impl (:1452-1531).
//
// This file is compiled with --fmad=false: all C++ float/double expressions are
// evaluated uncontracted, in the reference's mixed precision.  The only fused
// arithmetic is the explicit fma.rn.f32x2 of the prefilter, whose rounding error is
// covered by the candidate margin (see prefilter_threshold()).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "cvo_device.cuh"
#include "cvo_math.cuh"

namespace cvo_b200 {

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
// packed fp32x2 FMA (sm_100+): two pair tests per issue slot
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b,
                                                   unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float min3(float a, float b, float c) {
  float d;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
  } while (!ok);
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// CvoGPU.cu:86-90 compute_range_ell
__device__ __forceinline__ float range_ell(float curr_ell, float dist_to_sensor) {
  float final_ell = ((dist_to_sensor) / 500.0 + 1.0) * curr_ell;
  return final_ell;
}

// ================================================================== prep_kernel
// Prefilter identity: |x~ - y~|^2 = |y~|^2 - 2 x~.y~ + |x~|^2 with x~ = x - c, y~ = y' - c.
// pair_kernel evaluates s = w + ax*yx + ay*yy + az*yz (a = -2 x~, w = |y~|^2) with three packed
// FMAs per TWO pairs and tests s < t_i, t_i = thres_i - |x~_i|^2 + margin_i.  margin_i bounds
// every rounding error between s and the reference's float d2 (gpu_utils.cuh:73-78):
//   * centring x~ = fl(x-c), y~ = fl(y'-c): distance error <= u(|x~|+|y~|)
//   * w = fl(|y~|^2) and the three FMAs: <= 6u(|x~|+|y~|)^2
//   * the reference's own d2 rounding: <= 5u*thres
// (u = 2^-24) with a 2x safety factor, so the prefilter can only ADD candidates; the exact test
// in flow_kernel removes them again.  |y~| is bounded by ymax2_bound (controller).
__device__ __forceinline__ float prefilter_threshold(float d2_thres, const float4& a, float ymax2) {
  const double u = 5.9604644775390625e-08;  // 2^-24
  const double th = (double)d2_thres;
  if (!(th > 0.0)) return -INFINITY;  // d2 < thres can never hold
  const double nx =
      0.25 * ((double)a.x * (double)a.x + (double)a.y * (double)a.y + (double)a.z * (double)a.z);
  const double sN = sqrt(nx) + sqrt((double)ymax2);
  const double margin = 16.0 * u * sN * sN + 4.0 * u * sqrt(th) * sN + 1e-6 * th;
  float t = __double2float_ru(th - nx + margin);
  if (isnan(a.x) || isnan(a.y) || isnan(a.z) || isnan(a.w)) t = -INFINITY;
  return t;
}

__global__ void __launch_bounds__(256) prep_kernel(IterArgs A) {
  DevState* st = A.st;
  if (st->done) return;
  const int view = st->view;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int stride = gridDim.x * blockDim.x;
  if (gid == 0) st->dbg[8] = gtime();
  float Ri[9], Ti[3];
#pragma unroll
  for (int k = 0; k < 9; k++) Ri[k] = st->Rinv[k];
#pragma unroll
  for (int k = 0; k < 3; k++) Ti[k] = st->Tinv[k];
  // ---- targets: exact moved coordinates + the centred SoA operand (padded to M_pad)
  for (int j = gid; j < A.M_pad; j += stride) {
    if (j < A.M) {
      const float4 y = A.tv[view].xyz[j];
      const float yv[3] = {y.x, y.y, y.z};
      float r[3];
      mat3f_vec(Ri, yv, r);  // (*R) * input, CvoGPU_impl.cu:46-50
      const float m0 = r[0] + Ti[0], m1 = r[1] + Ti[1], m2 = r[2] + Ti[2];
      A.tgt_moved[j] = make_float4(m0, m1, m2, y.w);  // .w: the point's packed colour summary
      const float ux = m0 - A.cx, uy = m1 - A.cy, uz = m2 - A.cz;
      A.px[j] = ux;
      A.py[j] = uy;
      A.pz[j] = uz;
      A.pw[j] = ux * ux + uy * uy + uz * uz;
      A.pq[j] = __float_as_uint(y.w);
    } else {
      A.px[j] = 0.f;
      A.py[j] = 0.f;
      A.pz[j] = 0.f;
      A.pw[j] = INFINITY;  // padding can never be a candidate
      A.pq[j] = 0u;
    }
  }
  // ---- source rows: range-scaled length-scale, exact threshold, prefilter record
  const float ell = st->ell;
  const float ymax2 = st->ymax2_bound;
  const int use_geometry = st->kc.use_geometry;
  const float log_geo = st->kc.log_geo;
  for (int r = gid; r < A.n_rows; r += stride) {
    const float4 a = A.src_rowA[A.row_begin + r];
    const float l = range_ell(ell, a.w);  // CvoGPU.cu:506-507
    float d2_thres = 1.f;
    float t;
    if (use_geometry && A.mode == 0) {
      d2_thres = -2.0 * l * l * log_geo;  // CvoGPU.cu:511
      t = prefilter_threshold(d2_thres, a, ymax2);
    } else {
      t = INFINITY;  // no geometric cut in the reference either: every pair is a candidate
    }
    A.row_lt[r] = make_float2(l, d2_thres);
    A.rowrec[2 * r] = make_float4(a.x, a.x, a.y, a.y);
    // .w: the row's packed colour summary (emission path of pair_kernel)
    A.rowrec[2 * r + 1] = make_float4(a.z, a.z, t, A.src_xyz[A.row_begin + r].w);
  }
}

// ================================================================== pair_kernel
struct __align__(16) PairSmemWarp {
  float4 rec[2][2 * kTileRows];  // double-buffered TMA destination
  uint32_t cnt[kTileRows];
  uint64_t bar[2];
};

// candidate word: the first target index of the lane's run of 8 consecutive targets in the high
// 24 bits (any alignment; clouds of up to 2^24 points), the 8-bit mask of candidate targets inside
// the run in the low bits
__device__ __forceinline__ uint32_t make_word(int j_base, uint32_t qmask) {
  return ((uint32_t)j_base << 8) | qmask;
}

__global__ void __launch_bounds__(kPairWarps * 32, 3) pair_kernel(IterArgs A) {
  DevState* st = A.st;
  if (st->done) return;
  const int view = st->view;
  __shared__ PairSmemWarp smem[kPairWarps];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  PairSmemWarp& S = smem[warp];
  const int L = A.L;
  const unsigned lt_mask = (1u << lane) - 1u;
  // colour cut of the emission path: integer threshold on sum_k max(|qa_k - qb_k| - 1, 0)^2
  const bool colour_cut = st->kc.use_intensity != 0;
  unsigned int lb_thr = 0xffffffffu;
  if (colour_cut) {
    const float th = st->kc.d2_c_thres;
    // d2_color >= lb / 255^2; reject when lb * 0.999 / 65025 >= th.  th <= 0 or NaN: the full test
    // rejects every pair, so does lb >= 0.
    lb_thr = (th > 0.f) ? (th * (65025.f / 0.999f) < 4.0e9f ? (unsigned int)ceilf(th * (65025.f / 0.999f)) : 0xffffffffu) : 0u;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) st->dbg[9] = gtime();

  if (lane == 0) {
    mbar_init(&S.bar[0], 1);
    mbar_init(&S.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t phase0 = 0, phase1 = 0;

  // (tile, chunk) items: the first one of every warp is static; the rest is static round-robin
  // on small problems (a global work counter costs ~10^4 same-address L2 atomics per launch, more
  // than the sweep itself on 10k x 10k clouds) and handed out dynamically on large ones, where
  // sphere pruning makes the items very unequal (27 % of the 200k sweep was tail otherwise)
  const int gwarp = blockIdx.x * kPairWarps + warp;
  const int nwarps = gridDim.x * kPairWarps;
  // measured on the 200k clouds: heavy-first (diagonal-major) order + dynamic hand-out is 6 %
  // faster than static; on KITTI-sized clouds the ~10^4 atomics would cost more than the sweep
  const bool dynamic_items = A.n_items > 4 * nwarps && (double)A.n_rows * (double)A.M > 2.0e9;
  int next_item = gwarp;
  auto fetch_item = [&]() -> int {
    const int it = next_item;
    if (dynamic_items) {
      int nx = 0;
      if (lane == 0) nx = nwarps + (int)atomicAdd(&st->item_counter, 1u);
      next_item = __shfl_sync(0xffffffffu, nx, 0);
    } else {
      next_item += nwarps;
    }
    return it;
  };
  // Item k -> (source tile, target chunk) in DIAGONAL-MAJOR order.  Both clouds are Morton ordered,
  // so tile t mostly meets the chunks around c0(t) = t * nchunks / ntiles; everything else is
  // pruned at once.  Enumerating the diagonals c0, c0+1, c0-1, c0+2, ... tile by tile puts the
  // heavy items first and deals them round-robin to all warps, instead of giving a warp a run of
  // chunks of ONE tile (a few heavy, most empty).
  const int ntiles = A.n_items / A.nchunks;
  const long long ntiles_global = ((long long)A.n_src_total + kTileRows - 1) / kTileRows;
  auto decode_item = [&](int item, int& rt, int& jc) {
    rt = item % ntiles;
    const int d = item / ntiles;
    const int c0 = (int)(((long long)(A.row_begin / kTileRows + rt) * A.nchunks) / ntiles_global);
    const int off = (d & 1) ? (d + 1) / 2 : -(d / 2);
    jc = ((c0 + off) % A.nchunks + A.nchunks) % A.nchunks;
  };
  auto issue_tile = [&](int item, int buf) {
    if (item >= A.n_items) return;
    int rt, jc_unused;
    decode_item(item, rt, jc_unused);
    const int row0 = rt * kTileRows;
    const int nrows = min(kTileRows, A.n_rows - row0);
    if (lane == 0) {
      fence_proxy_async();
      const uint32_t bytes = (uint32_t)nrows * 2u * (uint32_t)sizeof(float4);
      mbar_expect_tx(&S.bar[buf], bytes);
      tma_bulk_g2s(S.rec[buf], A.rowrec + 2 * (size_t)row0, bytes, &S.bar[buf]);
    }
  };

  int item = fetch_item();
  int buf = 0;
  issue_tile(item, 0);
  while (item < A.n_items) {
    const int next = fetch_item();
    issue_tile(next, buf ^ 1);  // prefetch the next source tile while this one is swept

    int rt, jc;
    decode_item(item, rt, jc);
    const int row0 = rt * kTileRows;
    const int nrows = min(kTileRows, A.n_rows - row0);
    if (buf == 0) {
      mbar_wait(&S.bar[0], phase0);
      phase0 ^= 1u;
    } else {
      mbar_wait(&S.bar[1], phase1);
      phase1 ^= 1u;
    }
    float4* rec = S.rec[buf];
#pragma unroll
    for (int h = 0; h < kTileRows / 32; h++) {
      const int r = lane + 32 * h;
      S.cnt[r] = 0u;
      if (r >= nrows) {  // rows past the end of a ragged tile: stale smem, neutralise
        rec[2 * r] = make_float4(0.f, 0.f, 0.f, 0.f);
        rec[2 * r + 1] = make_float4(0.f, 0.f, -INFINITY, -INFINITY);
      }
    }
    __syncwarp();

    const int j_begin = jc * A.chunk_len;
    const int j_end = min(A.M_pad, j_begin + A.chunk_len);
    uint32_t* cell0 = A.cand + ((size_t)row0 * A.nchunks + jc) * (size_t)L;
    const size_t cell_stride = (size_t)A.nchunks * (size_t)L;
    // ---- pruning data of this source tile: bounding sphere and the largest cut-off radius of
    //      its rows, sqrt(max_i d2_thres_i) with l_i = ell (1 + dist_i/500)  (CvoGPU.cu:506-511)
    float4 tsph = make_float4(0.f, 0.f, 0.f, 0.f);
    float reach = INFINITY;
    const bool prune = A.prune && view == 0;
    if (prune) {
      const int gt = (A.row_begin + row0) / kTileRows;
      tsph = A.tile_sphere[gt];
      const double lmax = ((double)A.tile_maxdist[gt] / 500.0 + 1.0) * (double)st->ell;
      const double th = -2.0 * lmax * lmax * (double)st->kc.log_geo * (1.0 + 1e-5);
      reach = (float)(sqrt(fmax(th, 0.0)) * (1.0 + 1e-5)) + tsph.w + 1e-4f;
    }

    for (int jb = j_begin; jb < j_end; jb += kJBlock) {
      if (prune) {
        // conservative skip: no pair of this (tile, block) can pass d2 < thres when the moved
        // block sphere and the tile sphere are farther apart than the largest cut-off
        const float4 b = A.blk_sphere[jb / kJBlock];
        const float bx = st->Rinv[0] * b.x + st->Rinv[3] * b.y + st->Rinv[6] * b.z + st->Tinv[0];
        const float by = st->Rinv[1] * b.x + st->Rinv[4] * b.y + st->Rinv[7] * b.z + st->Tinv[1];
        const float bz = st->Rinv[2] * b.x + st->Rinv[5] * b.y + st->Rinv[8] * b.z + st->Tinv[2];
        const float dx = bx - tsph.x, dy = by - tsph.y, dz = bz - tsph.z;
        const float lim = reach + b.w * st->smax * 1.0001f + 1e-4f * (fabsf(bx) + fabsf(by) + fabsf(bz));
        if (dx * dx + dy * dy + dz * dz > lim * lim) continue;
      }
      // ---- stream 256 targets into registers: lane owns targets jb + 8*lane .. +7 (two float4
      //      per coordinate, a warp reads 1 KB contiguous per array), already packed in pairs
      const int jl = jb + 8 * lane;
      unsigned long long X[4], Y[4], Z[4], W[4];
      {
        const float4 xa = __ldg(reinterpret_cast<const float4*>(A.px + jl));
        const float4 xb = __ldg(reinterpret_cast<const float4*>(A.px + jl + 4));
        const float4 ya = __ldg(reinterpret_cast<const float4*>(A.py + jl));
        const float4 yb = __ldg(reinterpret_cast<const float4*>(A.py + jl + 4));
        const float4 za = __ldg(reinterpret_cast<const float4*>(A.pz + jl));
        const float4 zb = __ldg(reinterpret_cast<const float4*>(A.pz + jl + 4));
        const float4 wa = __ldg(reinterpret_cast<const float4*>(A.pw + jl));
        const float4 wb = __ldg(reinterpret_cast<const float4*>(A.pw + jl + 4));
        X[0] = pack2(xa.x, xa.y); X[1] = pack2(xa.z, xa.w); X[2] = pack2(xb.x, xb.y); X[3] = pack2(xb.z, xb.w);
        Y[0] = pack2(ya.x, ya.y); Y[1] = pack2(ya.z, ya.w); Y[2] = pack2(yb.x, yb.y); Y[3] = pack2(yb.z, yb.w);
        Z[0] = pack2(za.x, za.y); Z[1] = pack2(za.z, za.w); Z[2] = pack2(zb.x, zb.y); Z[3] = pack2(zb.z, zb.w);
        W[0] = pack2(wa.x, wa.y); W[1] = pack2(wa.z, wa.w); W[2] = pack2(wb.x, wb.y); W[3] = pack2(wb.z, wb.w);
      }
      // min over the lane's 8 pair tests of one row
      auto row_min = [&](int r, float& t) -> float {
        const float4 ra = rec[2 * r];
        const float4 rb = rec[2 * r + 1];
        const unsigned long long AX = pack2(ra.x, ra.y), AY = pack2(ra.z, ra.w),
                                 AZ = pack2(rb.x, rb.y);
        t = rb.z;
        float s[8];
#pragma unroll
        for (int p = 0; p < 4; p++) {
          unsigned long long v = fma2(AX, X[p], W[p]);
          v = fma2(AY, Y[p], v);
          v = fma2(AZ, Z[p], v);
          unpack2(v, s[2 * p], s[2 * p + 1]);
        }
        const float m = fminf(min3(s[0], s[1], s[2]), min3(s[3], s[4], s[5]));
        return min3(m, s[6], s[7]);
      };
      // rare path: some lane of the warp holds a candidate of row r -> ordered emission
      auto emit_row = [&](int r, unsigned lanes) {
        (void)lanes;
        const float4 ra = rec[2 * r];
        const float4 rb = rec[2 * r + 1];
        const unsigned long long AX = pack2(ra.x, ra.y), AY = pack2(ra.z, ra.w),
                                 AZ = pack2(rb.x, rb.y);
        const float t = rb.z;
        uint32_t qmask = 0;
#pragma unroll
        for (int p = 0; p < 4; p++) {
          unsigned long long v = fma2(AX, X[p], W[p]);
          v = fma2(AY, Y[p], v);
          v = fma2(AZ, Z[p], v);
          float s0, s1;
          unpack2(v, s0, s1);
          qmask |= (s0 < t ? 1u : 0u) << (2 * p);
          qmask |= (s1 < t ? 1u : 0u) << (2 * p + 1);
        }
        if (colour_cut && qmask != 0u) {
          // lower bound of the colour distance from the 8-bit summaries (see eval_pair): a target
          // whose bound already fails d2_color < d2_c_thres is dropped here, so that it never
          // becomes a candidate (on colour-rich clouds 99 % of the geometric candidates)
          const unsigned int qa = __float_as_uint(rb.w);
          const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(A.pq + jl));
          const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(A.pq + jl + 4));
          const unsigned int qb[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
          for (int q = 0; q < 8; q++) {
            const unsigned int d = __vsubus4(__vabsdiffu4(qa, qb[q]), 0x01010101u);
            if (__dp4a(d, d, 0u) >= lb_thr) qmask &= ~(1u << q);
          }
        }
        lanes = __ballot_sync(0xffffffffu, qmask != 0u);
        uint32_t c = S.cnt[r];
        const uint32_t pos = c + __popc(lanes & lt_mask);
        if (qmask != 0u && pos < (uint32_t)L)
          cell0[(size_t)r * cell_stride + pos] = make_word(jl, qmask);
        c += __popc(lanes);
        __syncwarp();
        if (lane == 0) {
          S.cnt[r] = c;
          if (c > (uint32_t)L) {  // cell overflowed (flow rescans it): stop looking at this row in this chunk
            rec[2 * r + 1].z = -INFINITY;
            rec[2 * r + 1].w = -INFINITY;
          }
        }
        __syncwarp();
      };
      // ---- sweep the staged source rows, four rows per warp vote
#pragma unroll 1
      for (int r = 0; r < kTileRows; r += 4) {
        float t0, t1, t2, t3;
        const float m0 = row_min(r, t0);
        const float m1 = row_min(r + 1, t1);
        const float m2 = row_min(r + 2, t2);
        const float m3 = row_min(r + 3, t3);
        const bool f = (m0 < t0) | (m1 < t1) | (m2 < t2) | (m3 < t3);
        if (__any_sync(0xffffffffu, f)) {
          unsigned b;
          b = __ballot_sync(0xffffffffu, m0 < t0);
          if (b) emit_row(r, b);
          b = __ballot_sync(0xffffffffu, m1 < t1);
          if (b) emit_row(r + 1, b);
          b = __ballot_sync(0xffffffffu, m2 < t2);
          if (b) emit_row(r + 2, b);
          b = __ballot_sync(0xffffffffu, m3 < t3);
          if (b) emit_row(r + 3, b);
        }
      }
    }
    // ---- publish the per-cell word counts (count > L means "overflowed": flow_kernel rescans)
#pragma unroll
    for (int h = 0; h < kTileRows / 32; h++) {
      const int r = lane + 32 * h;
      if (r < nrows) A.cand_cnt[(size_t)(row0 + r) * A.nchunks + jc] = S.cnt[r];
    }
    __syncwarp();
    item = next;
    buf ^= 1;
  }
}

// ================================================================== exact pair arithmetic
// The reference's device exp(double) (CvoGPU.cu:552,567,580).  One out-of-line copy: the
// sparse kernels inline eval_pair at several call sites and each holds three exps; inlined,
// the kernels grow past 200 KB of SASS and stall on instruction fetch.
__device__ __noinline__ double exp_ref(double x) { return exp(x); }

struct RowCtx {
  float px[3];
  unsigned int qa;  // packed colour summary of the source point (cvo_upload.cu)
  float l;          // range-scaled length-scale of this row
  float d2_thres;
  float ga[2];
};

// The body of the j-loop of fill_in_A_mat_gpu (CvoGPU.cu:534-589) / of the dense-kernel
// variant (:279-321) for one (i, j).  Returns true if the pair is stored; a = A_ij.
// pb = the moved target point y'_j; j indexes the target arrays of `view`.
__device__ __forceinline__ bool eval_pair(const IterArgs& A, const KernConsts& kc,
                                          const RowCtx& rc, int i_global, int view, int j,
                                          const float4& pb, float& a_out) {
  // ARITHMETIC: this is the reference's kernel AS ITS GPU BUILD COMPUTES IT.  The reference compiles
  // Release with nvcc's default --fmad=true (CMakeLists.txt:29,79), so the multiply-adds of its
  // text are contracted; oracle/_ref/ref_harness.ptx (the reference's own text under its own
  // flags, built by oracle/make_ref.py) shows which: `result += t*t` -> fma(t,t,result) in
  // dot / squared_dist / square_norm, and dx*dx + dy*dy + dz*dz -> fma(dz,dz,fma(dx,dx,dy*dy)).
  // This file is compiled with --fmad=false, so the contractions are spelled out here and nothing
  // else is fused.  tests/test_ref_pin_gpu.py holds the result to the reference kernel bit for bit.
  // The reference interleaves tests and kernel values (geometry test -> k = exp -> colour test ->
  // ck = exp -> ...).  Every test only rejects the pair (no side effect), so all the cheap float
  // tests run first and the double-precision exps only for pairs that pass them all: the same
  // pairs are stored with the same values, and a pair that fails the colour test (most
  // geometric survivors do) never pays for an exp.
  float a = 1, sk = 1, ck = 1, k = 1, geo_sim = 1;
  float d2 = 0.f, d2_color = 0.f, d2_semantic = 0.f;
  if (kc.use_geo_type && A.mode == 0) {  // mode 1 switches it off, CvoGPU.cu:1948-1949
    const float2 gb = A.tv[view].geo[j];
    const float norm2_a = __fmaf_rn(rc.ga[1], rc.ga[1], __fmaf_rn(rc.ga[0], rc.ga[0], 0.f));
    const float norm2_b = __fmaf_rn(gb.y, gb.y, __fmaf_rn(gb.x, gb.x, 0.f));
    const float dot_ab = __fmaf_rn(rc.ga[1], gb.y, __fmaf_rn(rc.ga[0], gb.x, 0.f));
    geo_sim = dot_ab * dot_ab / (norm2_a * norm2_b);
    if (geo_sim < 0.01) return false;
  }
  if (kc.use_geometry) {
    if (A.mode == 0) {
      const float dx = pb.x - rc.px[0], dy = pb.y - rc.px[1], dz = pb.z - rc.px[2];
      d2 = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, dy * dy));  // gpu_utils.cuh:73-78 as nvcc contracts it
      if (!(d2 < rc.d2_thres)) return false;
    } else {
      const float dist[3] = {rc.px[0] - pb.x, rc.px[1] - pb.y, rc.px[2] - pb.z};
      float row[3];
#pragma unroll
      for (int c = 0; c < 3; c++)
        row[c] = sum3f(dist[0] * A.kinv[3 * c], dist[1] * A.kinv[3 * c + 1],
                       dist[2] * A.kinv[3 * c + 2]);
      d2 = dot3f(row, dist);
    }
  }
  if (kc.use_intensity) {
    // lower bound of the colour distance from the 8-bit summaries: per channel
    // |fa - fb| >= (|qa - qb| - 1) / 255 (clamping to [0,1] is 1-Lipschitz, so it only weakens the
    // bound).  If even the bound fails the reference's test d2_color < d2_c_thres, the pair is
    // rejected exactly as the full test would reject it - without loading the feature rows.
    {
      const unsigned int d = __vsubus4(__vabsdiffu4(rc.qa, __float_as_uint(pb.w)), 0x01010101u);
      const unsigned int lb = __dp4a(d, d, 0u);
      if ((float)lb * (0.999f / 65025.f) >= kc.d2_c_thres) return false;
    }
    const float* fa = A.src_feat + (size_t)i_global * A.Fp;
    const float* fb = A.tv[view].feat + (size_t)j * A.Fp;
    for (int f = 0; f < A.Fp; f += 4) {
      const float4 va = *reinterpret_cast<const float4*>(fa + f);
      const float4 vb = *reinterpret_cast<const float4*>(fb + f);
      float tmp = va.x - vb.x;
      d2_color = __fmaf_rn(tmp, tmp, d2_color);
      tmp = va.y - vb.y;
      d2_color = __fmaf_rn(tmp, tmp, d2_color);
      tmp = va.z - vb.z;
      d2_color = __fmaf_rn(tmp, tmp, d2_color);
      tmp = va.w - vb.w;
      d2_color = __fmaf_rn(tmp, tmp, d2_color);
    }
    if (!(d2_color < kc.d2_c_thres)) return false;
  }
  if (kc.use_semantics) {
    const float* la = A.src_lab + (size_t)i_global * A.Cp;
    const float* lb = A.tv[view].lab + (size_t)j * A.Cp;
    for (int c = 0; c < A.Cp; c += 4) {
      const float4 va = *reinterpret_cast<const float4*>(la + c);
      const float4 vb = *reinterpret_cast<const float4*>(lb + c);
      float tmp = va.x - vb.x;
      d2_semantic = __fmaf_rn(tmp, tmp, d2_semantic);
      tmp = va.y - vb.y;
      d2_semantic = __fmaf_rn(tmp, tmp, d2_semantic);
      tmp = va.z - vb.z;
      d2_semantic = __fmaf_rn(tmp, tmp, d2_semantic);
      tmp = va.w - vb.w;
      d2_semantic = __fmaf_rn(tmp, tmp, d2_semantic);
    }
    const float thr = (A.mode == 1) ? kc.d2_s_thres_dense : kc.d2_s_thres;
    if (!(d2_semantic < thr)) return false;
  }
  // ---- kernel values (CvoGPU.cu:552,567,580 and :297,311,318 for the dense-kernel variant)
  if (kc.use_geometry)
    k = (A.mode == 0) ? kc.sigma2 * exp_ref(-d2 / (2.0 * rc.l * rc.l)) : kc.sigma2 * exp_ref(-d2 / 2.0);
  if (kc.use_intensity) ck = kc.c_sigma2 * exp_ref(-d2_color / (2.0 * kc.c2));
  if (kc.use_semantics) {
    if (A.mode == 1)
      sk = kc.s_sigma2 * exp_ref(-d2_semantic / (2.0 * kc.s_ell_square));
    else
      sk = kc.s_sigma2 * exp_ref(-d2_semantic / (2.0 * kc.s_ell * kc.s_ell));
  }
  a = ck * k * sk * geo_sim;
  a_out = a;
  return a > kc.sp_thres;
}

// ================================================================== flow finalisation
// thrust::reduce results -> float, joint normalisation (CvoGPU.cu:824-838), and the omega_hat
// powers of compute_step_size_xi (CvoGPU.cu:970-980), evaluated once per iteration, left to
// right like the Eigen expressions ((W*W)*W)*W and (W*W)*v.
__device__ void finalize_flow_scalar(DevState* st, const double tot[9]) {
  for (int k = 0; k < 3; k++) {
    st->omega_sum[k] = tot[k];
    st->v_sum[k] = tot[3 + k];
  }
  st->a_sum = tot[6];
  st->nnz = (unsigned long long)(tot[7] + 0.5);
  st->max_row_nnz = (unsigned int)(tot[8] + 0.5);
  float ov[6];
  for (int k = 0; k < 6; k++) ov[k] = (float)tot[k];
  const float z = sum3f(ov[0] * ov[0], ov[1] * ov[1], ov[2] * ov[2]) +
                  sum3f(ov[3] * ov[3], ov[4] * ov[4], ov[5] * ov[5]);
  if (z > 0.f) {
    const float nrm = sqrtf(z);
    for (int k = 0; k < 6; k++) ov[k] = ov[k] / nrm;
  }
  float W[9], W2[9], W3[9], W4[9], t3[3];
  for (int k = 0; k < 3; k++) {
    st->omega[k] = ov[k];
    st->v[k] = ov[3 + k];
  }
  skewf(ov, W);
  mat3f_mul(W, W, W2);
  mat3f_mul(W2, W, W3);
  mat3f_mul(W3, W, W4);
  for (int k = 0; k < 9; k++) {
    st->W2[k] = W2[k];
    st->W3[k] = W3[k];
    st->W4[k] = W4[k];
  }
  mat3f_vec(W, ov + 3, t3);
  for (int k = 0; k < 3; k++) st->Wv[k] = t3[k];
  mat3f_vec(W2, ov + 3, t3);
  for (int k = 0; k < 3; k++) st->W2v[k] = t3[k];
  mat3f_vec(W3, ov + 3, t3);
  for (int k = 0; k < 3; k++) st->W3v[k] = t3[k];
}

// deterministic block-wide reduction of per-block partials (value-major: part[k * nparts + b]).
// One WARP per value: its lanes stride over the partials in a fixed order, then a fixed
// xor-shuffle tree.  The first NSUM values are summed, the rest are max-reduced.  All NV values
// are reduced concurrently when the block has >= NV warps, so the latency is one load round
// plus one shuffle tree regardless of NV.  out[] is valid on thread 0.
template <int NV, int NSUM>
__device__ void block_reduce_partials(const double* __restrict__ part, int nparts,
                                      double* out /* NV, valid on thread 0 */,
                                      double* sh /* >= NV */,
                                      unsigned long long* dbg = nullptr) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();  // sh may still be in use by the caller's previous phase
  for (int k = w; k < NV; k += nw) {  // one pass when the block has >= NV warps
    double acc = 0.0;
    for (int b = lane; b < nparts; b += 32) {
      const double x = __ldcg(part + (size_t)k * nparts + b);
      acc = (k < NSUM) ? acc + x : fmax(acc, x);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double x = __shfl_xor_sync(0xffffffffu, acc, o);
      acc = (k < NSUM) ? acc + x : fmax(acc, x);
    }
    if (lane == 0) sh[k] = acc;
  }
  if (dbg && threadIdx.x == 0) dbg[0] = dbg[1] = dbg[2] = gtime();
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) out[k] = sh[k];
  }
}

// ================================================================== cell queries (grid mode)
// y' = Rinv*y + Tinv with the arithmetic of prep_kernel (CvoGPU_impl.cu:46-50): grid mode has no
// prep launch, every consumer of a moved target point recomputes it from the static cloud.
__device__ __forceinline__ float4 move_point(const float* Ri, const float* Ti, const float4 y) {
  const float yv[3] = {y.x, y.y, y.z};
  float r[3];
  mat3f_vec(Ri, yv, r);
  return make_float4(r[0] + Ti[0], r[1] + Ti[1], r[2] + Ti[2], y.w);  // .w: packed colour summary
}
// transform_point_pose_vec (CvoGPU_impl.cu:84-150): x' = P [x y z 1]^T, P row-major 3x4; Eigen's
// unrolled 4-term redux sums (c0 + c1) + (c2 + c3)
__device__ __forceinline__ float4 move_point_pose(const float* P, const float4 y) {
  float r[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const float c0 = P[4 * i] * y.x, c1 = P[4 * i + 1] * y.y, c2 = P[4 * i + 2] * y.z, c3 = P[4 * i + 3] * 1.0f;
    r[i] = (c0 + c1) + (c2 + c3);
  }
  return make_float4(r[0], r[1], r[2], y.w);
}
// the moved target point of the "on the fly" generators: by the state's inverse pose, or (pose-graph
// edges) by frame 2's pose vector
__device__ __forceinline__ float4 move_target(const IterArgs& A, const float* Ri, const float* Ti, const float4 y) {
  return A.posevec ? move_point_pose(A.pose2, y) : move_point(Ri, Ti, y);
}
// 21 bits -> every third bit (the host's spread21, cvo_engine.cu)
__device__ __forceinline__ unsigned long long spread21_dev(unsigned int a) {
  unsigned long long v = a & 0x1fffffull;
  v = (v | (v << 32)) & 0x1f00000000ffffull;
  v = (v | (v << 16)) & 0x1f0000ff0000ffull;
  v = (v | (v << 8)) & 0x100f00f00f00f00full;
  v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
  v = (v | (v << 2)) & 0x1249249249249249ull;
  return v;
}
// Ranges [start, start+len) of the Morton-ordered target covered by two cube cells of edge
// 2^lvl lattice units with first keys k0, k1 (v0/v1: the lane has such a cell).  A cell's keys
// are [k, k + 8^lvl).  Query cells are never finer than the cells of the coarse table
// (lvl >= 21 - cbits), so both bounds of a range are table entries: four independent loads,
// no search.  The table is sized for <1 point per coarse cell, so late iterations (cut-off
// radius far below the coarse cell) still test only a handful of points per row.
__device__ __forceinline__ void cell_ranges2(const GridView& G, int lvl3, bool v0,
                                             unsigned long long k0, bool v1, unsigned long long k1,
                                             uint32_t& s0, uint32_t& l0, uint32_t& s1, uint32_t& l1) {
  const int csh = 3 * (21 - G.cbits);
  s0 = l0 = s1 = l1 = 0u;
  if (v0) {
    s0 = __ldg(G.coarse + (k0 >> csh));
    l0 = __ldg(G.coarse + ((k0 + (1ull << lvl3)) >> csh)) - s0;
  }
  if (v1) {
    s1 = __ldg(G.coarse + (k1 >> csh));
    l1 = __ldg(G.coarse + ((k1 + (1ull << lvl3)) >> csh)) - s1;
  }
}


// ================================================================== tile_kernel
// Candidate generator of the dense regimes (large cut-off radius against the point spacing, e.g.
// the 200k x 200k clouds at ell = 1.5, where a sixth of the tested pairs pass the geometric cut):
// one warp takes a TILE of 64 Morton-consecutive source rows and tests it against exactly the
// targets its rows can reach - the points of the octree cells of the target (in the target's OWN
// frame, where the cell table stays valid under any pose) that intersect the ball around the tile.
// Per tile: (1) the tile's source points are TMA-staged into shared memory (cp.async.bulk +
// mbarrier) and turned into prefilter records in that frame (q = R x + T, conservative radius as
// in the cell queries); every lane OWNS two rows and keeps their records in registers; (2) <= 4 x 4
// x 4 cube cells covering the tile's bounding box are looked up as contiguous ranges of the
// Morton-ordered target; (3) the ranges are staged through shared memory, 128 targets at a time
// (coalesced float4 loads, centred SoA), and every lane runs its two rows past them:
// s = w + ax*yx + ay*yy + az*yz < t_i, two TARGETS per fma.rn.f32x2 (same identity and margin
// analysis as pair_kernel: the prefilter can only ADD candidates, flow_rows<2> re-tests them with
// the reference's arithmetic), then the colour lower bound on the hits.  A lane appends its rows'
// candidates itself - no ballots, no ordered emission: with one pair in six passing, pair_kernel's
// "rare hit" path (warp votes, re-evaluation, ordered emission) ran for every row of every sweep.
// pair_kernel also tests every 256-target block whose bounding sphere is near the tile's (6-10x
// more points than necessary on these clouds); this kernel tests ~1.5x the points of the tile's ball.
// Output: per row ONE candidate cell of tile_L words (target << 8 | 1) + its count; a count above
// tile_L marks an overflowed row (exact redo).  Order inside a row is free: the Morton view counts
// one survivor past the cap and redoes cut rows exhaustively in original order.
constexpr int kTileChunk = 128;  // targets staged per sweep
struct __align__(16) TileSmemWarp {
  union {
    float4 stage[2 * kTileRows];  // TMA destination: src_xyz[64], src_rowA[64]
    struct {
      float X[kTileChunk], Y[kTileChunk], Z[kTileChunk], W[kTileChunk];  // centred targets, |y~|^2
    } t;
  };
  uint32_t col[kTileChunk];  // packed colour summaries of the staged targets (4 x 8 bit)
  uint32_t nb[kTileChunk];   // their squared norms: sum of the four squared channels
  uint32_t idx[kTileChunk];  // their Morton positions
  uint32_t cstart[32], cend[32];  // the current group of 32 cube cells: range of Morton positions
  uint32_t cpre[33];              // exclusive prefix of their run counts (runs of 4 targets)
};

// counter != nullptr: every warp takes item `gwarp` first and the rest from a MONOTONE global counter
// (items differ a lot in cost: a static deal left a third of a build waiting for the slowest block);
// over one phase the counter advances by exactly the number of items (every warp that had a first
// item ends with one failing fetch), so the caller tracks its base without resetting it.
__device__ __forceinline__ void tile_phase(const IterArgs& A, const DevState* hs, TileSmemWarp& S,
                                           uint64_t* bar, int gwarp, int nwarps, uint32_t& bar_phase,
                                           unsigned int* counter = nullptr, unsigned int counter_base = 0u) {
  const int lane = threadIdx.x & 31;
  const GridView& G = A.gv;
  const uint32_t L = (uint32_t)A.tile_L;
  const float ell_now = hs->ell;
  const float log_geo = hs->kc.log_geo;
  const float g_smax = hs->smax, g_slack = hs->grid_slack + A.edge_slack;
  const float skin_tr = hs->vl_str, skin_rot = hs->vl_srot;
  const float* Rf = hs->R;
  const float* Tf = hs->T;
  const bool colour_cut = hs->kc.use_intensity != 0;
  unsigned int lb_thr = 0xffffffffu;
  if (colour_cut) {
    const float th = hs->kc.d2_c_thres;
    lb_thr = (th > 0.f) ? (th * (65025.f / 0.999f) < 4.0e9f ? (unsigned int)ceilf(th * (65025.f / 0.999f)) : 0xffffffffu) : 0u;
  }
  // The per-hit colour test of this kernel: with q the 8-bit summaries, per channel
  // |fa - fb| >= (|dq| - 1) / 255, and max(|d| - 1, 0)^2 >= d^2 - 2|d|, sum|d| <= 2 sqrt(S) for four
  // channels (S = sum d^2), so  65025 d2_color >= S - 4 sqrt(S).  S = |qa|^2 + |qb|^2 - 2 qa.qb is one
  // dp4a and two adds per pair (the video-SIMD absolute differences of eval_pair's bound are
  // emulated on this architecture, ~20 instructions).  A pair with S >= s_thr = (2 + sqrt(4 + T))^2
  // has S - 4 sqrt(S) >= T = lb_thr, fails d2_color < d2_c_thres and never becomes a candidate.
  unsigned int s_thr = 0xffffffffu;
  if (colour_cut && lb_thr != 0xffffffffu) {
    if (lb_thr == 0u) {
      s_thr = 0u;
    } else {
      const double rt2 = 2.0 + sqrt(4.0 + (double)lb_thr);
      const double v = rt2 * rt2 * (1.0 + 1e-9) + 2.0;
      s_thr = v < 4.0e9 ? (unsigned int)v : 0xffffffffu;
    }
  }
  // |y - tc| <= trad for every target (own frame, static): the prefilter's bound on |y~|
  const float ymax2 = A.trad * A.trad * 1.000004f + 1e-12f;
  const int ntiles = (A.n_rows + kTileRows - 1) / kTileRows;
  const float top = 2097151.f;
  const int csh = 3 * (21 - G.cbits);
  // items = (tile, part): the cell groups of a tile are dealt round-robin to `tile_parts` warps, each
  // with its own candidate cell per row, so that small shards still fill the machine
  const int P = A.tile_parts;
  for (int item = gwarp; item < ntiles * P;) {
    const int this_item = item;
    if (counter) {  // fetch the next item now: the atomic's latency hides behind this item's work
      unsigned int nxt = 0u;
      if (lane == 0) nxt = atomicAdd(counter, 1u) - counter_base;
      item = nwarps + (int)__shfl_sync(0xffffffffu, nxt, 0);
    } else {
      item += nwarps;
    }
    const int tile = this_item / P, part = this_item - tile * P;
    const int row0 = tile * kTileRows;
    const int nrows = min(kTileRows, A.n_rows - row0);
    // ---- (1) TMA-stage the tile's source points and range records
    __syncwarp();
    if (lane == 0) {
      fence_proxy_async();
      const uint32_t bytes = (uint32_t)nrows * (uint32_t)sizeof(float4);
      mbar_expect_tx(bar, 2u * bytes);
      tma_bulk_g2s(S.stage, A.src_xyz + A.row_begin + row0, bytes, bar);
      tma_bulk_g2s(S.stage + kTileRows, A.src_rowA + A.row_begin + row0, bytes, bar);
    }
    mbar_wait(bar, bar_phase);
    bar_phase ^= 1u;
    // ---- the lane's two rows: prefilter records in registers, the row's ball in lattice units
    float rax[2], ray[2], raz[2], rt[2], fq[2][3], frq[2];
    unsigned int rqa[2], rna[2];
    unsigned int cut_after = 0u;  // bit hh: the tile is cut after row lane + 32 hh (Morton key jump)
#pragma unroll
    for (int hh = 0; hh < 2; hh++) {
      const int r = lane + 32 * hh;
      rax[hh] = ray[hh] = raz[hh] = 0.f;
      rt[hh] = -INFINITY;
      rqa[hh] = rna[hh] = 0u;
      frq[hh] = -1.f;
      fq[hh][0] = fq[hh][1] = fq[hh][2] = 0.f;
      if (r < nrows) {
        const float4 x4 = S.stage[r];
        const float4 a4 = S.stage[kTileRows + r];
        const float l = range_ell(ell_now, a4.w);          // CvoGPU.cu:506-507
        const float d2_thres = -2.0 * l * l * log_geo;     // CvoGPU.cu:511
        if (d2_thres > 0.f) {
          const float xv[3] = {x4.x, x4.y, x4.z};
          float q[3];
          mat3f_vec(Rf, xv, q);
          q[0] += Tf[0]; q[1] += Tf[1]; q[2] += Tf[2];
          // |y - q| <= smax |y' - x| + slack (update_tf_device); same radius as the cell queries
          float rq = sqrtf(d2_thres) * g_smax * 1.00001f + g_slack +
                     2e-6f * (fabsf(xv[0]) + fabsf(xv[1]) + fabsf(xv[2]) + fabsf(q[0]) + fabsf(q[1]) + fabsf(q[2]));
          // candidate-cell reuse (verlet_decide): the cells stay valid while q has moved by less than
          // the skin, |q - q_build| <= |R - R_build|_F |x| + |T - T_build|
          if (skin_tr > 0.f)
            rq += skin_tr + skin_rot * (sqrtf(xv[0] * xv[0] + xv[1] * xv[1] + xv[2] * xv[2]) * 1.000001f) +
                  4e-6f * sqrtf(d2_thres) * g_smax;
          const float th = rq * rq * 1.000001f;
          const float4 a = make_float4(-2.f * (q[0] - A.tcx), -2.f * (q[1] - A.tcy), -2.f * (q[2] - A.tcz), 0.f);
          const float t = prefilter_threshold(th, a, ymax2);
          const float f0 = (q[0] - G.lo[0]) * G.scale, f1 = (q[1] - G.lo[1]) * G.scale, f2 = (q[2] - G.lo[2]) * G.scale;
          const float fr = rq * G.scale + 2.f;  // + the float rounding of the lattice coordinates
          // a NaN anywhere makes every comparison false: no candidates, like d2 < thres
          const bool hit = f0 + fr >= -2.f && f1 + fr >= -2.f && f2 + fr >= -2.f && f0 - fr <= top + 2.f &&
                           f1 - fr <= top + 2.f && f2 - fr <= top + 2.f && t > -INFINITY;
          if (hit) {
            rax[hh] = a.x; ray[hh] = a.y; raz[hh] = a.z;
            rt[hh] = t;
            rqa[hh] = __float_as_uint(x4.w);  // the row's packed colour summary
            rna[hh] = __dp4a(rqa[hh], rqa[hh], 0u);
            fq[hh][0] = f0; fq[hh][1] = f1; fq[hh][2] = f2;
            frq[hh] = fr;
          }
        }
        // segments: Morton-consecutive rows can be far apart (the Z curve jumps); the tile is cut
        // wherever two neighbours lie in different cube nodes of the source's octree at the level
        // A.tile_cut_bits (edge ~2 tile edges), so every segment is spatially compact
        if (r + 1 < nrows) {
          const unsigned long long k0 = __ldg(A.src_keys + A.row_begin + row0 + r);
          const unsigned long long k1 = __ldg(A.src_keys + A.row_begin + row0 + r + 1);
          if (((k0 ^ k1) >> A.tile_cut_bits) != 0ull) cut_after |= 1u << hh;
        }
      }
    }
    __syncwarp();  // the staging area is reused for the target chunks from here on
    const unsigned long long cuts = (unsigned long long)__ballot_sync(0xffffffffu, cut_after & 1u) |
                                    ((unsigned long long)__ballot_sync(0xffffffffu, cut_after & 2u) << 32);
    // row-duplicated operands of the packed FMAs: (a, a) x (y_k, y_k+1) tests two targets at once
    const unsigned long long AX0 = pack2(rax[0], rax[0]), AY0 = pack2(ray[0], ray[0]), AZ0 = pack2(raz[0], raz[0]);
    const unsigned long long AX1 = pack2(rax[1], rax[1]), AY1 = pack2(ray[1], ray[1]), AZ1 = pack2(raz[1], raz[1]);
    uint32_t cnt0 = 0u, cnt1 = 0u;
    uint32_t* cell_a = A.cand + ((size_t)(row0 + lane) * (size_t)P + (size_t)part) * (size_t)L;
    uint32_t* cell_b = A.cand + ((size_t)(row0 + lane + 32) * (size_t)P + (size_t)part) * (size_t)L;
    unsigned long long n_swept = 0ull;
    for (int seg0 = 0; seg0 < nrows;) {
      // the segment [seg0, seg1): up to and including the first cut at or after seg0
      const unsigned long long rest = cuts >> seg0;
      const int seg1 = rest ? min(nrows, seg0 + __ffsll((long long)rest)) : nrows;
      const bool act0 = lane >= seg0 && lane < seg1 && frq[0] >= 0.f;
      const bool act1 = lane + 32 >= seg0 && lane + 32 < seg1 && frq[1] >= 0.f;
      const float t0 = act0 ? rt[0] : -INFINITY, t1 = act1 ? rt[1] : -INFINITY;
      // ---- the segment's box and ball in lattice units
      float lo3[3] = {INFINITY, INFINITY, INFINITY}, hi3[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int k = 0; k < 3; k++) {
        if (act0) { lo3[k] = fminf(lo3[k], fq[0][k] - frq[0]); hi3[k] = fmaxf(hi3[k], fq[0][k] + frq[0]); }
        if (act1) { lo3[k] = fminf(lo3[k], fq[1][k] - frq[1]); hi3[k] = fmaxf(hi3[k], fq[1][k] + frq[1]); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          lo3[k] = fminf(lo3[k], __shfl_xor_sync(0xffffffffu, lo3[k], o));
          hi3[k] = fmaxf(hi3[k], __shfl_xor_sync(0xffffffffu, hi3[k], o));
        }
      }
      seg0 = seg1;
      if (!(hi3[0] >= lo3[0])) continue;  // no active row
      const float c3[3] = {0.5f * (lo3[0] + hi3[0]), 0.5f * (lo3[1] + hi3[1]), 0.5f * (lo3[2] + hi3[2])};
      float brad = 0.f;  // radius of the ball around c3 that holds every active row's ball
      if (act0) {
        const float dx = fq[0][0] - c3[0], dy = fq[0][1] - c3[1], dz = fq[0][2] - c3[2];
        brad = fmaxf(brad, sqrtf(dx * dx + dy * dy + dz * dz) * 1.000001f + frq[0]);
      }
      if (act1) {
        const float dx = fq[1][0] - c3[0], dy = fq[1][1] - c3[1], dz = fq[1][2] - c3[2];
        brad = fmaxf(brad, sqrtf(dx * dx + dy * dy + dz * dz) * 1.000001f + frq[1]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) brad = fmaxf(brad, __shfl_xor_sync(0xffffffffu, brad, o));
      brad += 1.f;
      const float brad2 = brad * brad * 1.00001f;
      const int bx0 = (int)fminf(fmaxf(floorf(lo3[0]) - 1.f, 0.f), top), bx1 = (int)fminf(fmaxf(floorf(hi3[0]) + 1.f, 0.f), top);
      const int by0 = (int)fminf(fmaxf(floorf(lo3[1]) - 1.f, 0.f), top), by1 = (int)fminf(fmaxf(floorf(hi3[1]) + 1.f, 0.f), top);
      const int bz0 = (int)fminf(fmaxf(floorf(lo3[2]) - 1.f, 0.f), top), bz1 = (int)fminf(fmaxf(floorf(hi3[2]) + 1.f, 0.f), top);
      // ---- (2) cube cells covering the box: <= 12 per axis, never finer than the cell table
      int lvl = 21 - G.cbits;
      while (((bx1 >> lvl) - (bx0 >> lvl)) > 11 || ((by1 >> lvl) - (by0 >> lvl)) > 11 ||
             ((bz1 >> lvl) - (bz0 >> lvl)) > 11)
        lvl++;
      const int icx0 = bx0 >> lvl, icy0 = by0 >> lvl, icz0 = bz0 >> lvl;
      const int nx = (bx1 >> lvl) - icx0 + 1, ny = (by1 >> lvl) - icy0 + 1, nz = (bz1 >> lvl) - icz0 + 1;
      const int ncell = nx * ny * nz;
      const float ch = (float)(1 << lvl);
      for (int cbase = 32 * part; cbase < ncell; cbase += 32 * P) {
        // one cell per lane: dropped unless the segment's ball reaches it
        uint32_t s0 = 0u, e0 = 0u;
        const int c = cbase + lane;
        if (c < ncell) {
          const int cz = c / (nx * ny), cy = (c - cz * nx * ny) / nx, cx = c - cz * nx * ny - cy * nx;
          const float bx = (float)((icx0 + cx) << lvl), by = (float)((icy0 + cy) << lvl), bz = (float)((icz0 + cz) << lvl);
          const float ex = fmaxf(0.f, fmaxf(bx - c3[0], c3[0] - (bx + ch)));
          const float ey = fmaxf(0.f, fmaxf(by - c3[1], c3[1] - (by + ch)));
          const float ez = fmaxf(0.f, fmaxf(bz - c3[2], c3[2] - (bz + ch)));
          if ((ex * ex + ey * ey + ez * ez) <= brad2) {
            const unsigned long long key =
                (spread21_dev((unsigned)(icx0 + cx)) | (spread21_dev((unsigned)(icy0 + cy)) << 1) |
                 (spread21_dev((unsigned)(icz0 + cz)) << 2))
                << (3 * lvl);
            s0 = __ldg(G.coarse + (key >> csh));
            e0 = __ldg(G.coarse + ((key + (1ull << (3 * lvl))) >> csh));
          }
        }
        const uint32_t nrun = (e0 - s0 + 3u) >> 2;
        uint32_t inc = nrun;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t tt = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += tt;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
        if (total == 0u) continue;
        __syncwarp();  // the previous group's tables have been consumed
        S.cstart[lane] = s0;
        S.cend[lane] = e0;
        S.cpre[lane] = inc - nrun;
        if (lane == 0) S.cpre[32] = total;
        __syncwarp();
        n_swept += total;
        // ---- (3) sweep: 128 targets per chunk (one run of 4 consecutive targets per lane), staged
        //      in shared memory; every lane runs its two rows past them
        for (uint32_t sb = 0; sb < total; sb += 32u) {
          const uint32_t slot = sb + (uint32_t)lane;
          const bool valid = slot < total;
          int cc = 0;  // largest cell with cpre[cc] <= slot (empty cells share their successor's prefix)
#pragma unroll
          for (int stp = 16; stp > 0; stp >>= 1)
            if (S.cpre[cc + stp] <= slot) cc += stp;
          const uint32_t j0 = valid ? S.cstart[cc] + 4u * (slot - S.cpre[cc]) : 0u;
          const uint32_t jend = valid ? S.cend[cc] : 0u;
          __syncwarp();  // the previous chunk has been consumed
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const uint32_t j = j0 + (uint32_t)q;
            const bool v = j < jend;
            const float4 y = v ? __ldg(A.tv[0].xyz + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float ux = y.x - A.tcx, uy = y.y - A.tcy, uz = y.z - A.tcz;
            const int k = 4 * lane + q;
            S.t.X[k] = ux;
            S.t.Y[k] = uy;
            S.t.Z[k] = uz;
            S.t.W[k] = v ? (ux * ux + uy * uy + uz * uz) : INFINITY;  // padding: never a candidate
            S.col[k] = __float_as_uint(y.w);
            S.nb[k] = __dp4a(__float_as_uint(y.w), __float_as_uint(y.w), 0u);
            S.idx[k] = j;
          }
          __syncwarp();
          const int kmax = (int)min((uint32_t)kTileChunk, 4u * (total - sb));
          // geometric prefilter of the whole chunk first: 8 hit bits per step (4 targets x the
          // lane's 2 rows) collected in registers ...
          uint32_t hm[kTileChunk / 16];
#pragma unroll
          for (int wq = 0; wq < kTileChunk / 16; wq++) {
            uint32_t acc = 0u;
            if (16 * wq < kmax) {  // warp-uniform: a short last chunk stops early
#pragma unroll
              for (int it = 0; it < 4; it++) {
                const int k = 16 * wq + 4 * it;
                // four targets: two packed pairs per coordinate
                const float4 x4 = *reinterpret_cast<const float4*>(&S.t.X[k]);
                const float4 y4 = *reinterpret_cast<const float4*>(&S.t.Y[k]);
                const float4 z4 = *reinterpret_cast<const float4*>(&S.t.Z[k]);
                const float4 w4 = *reinterpret_cast<const float4*>(&S.t.W[k]);
                const unsigned long long Xa = pack2(x4.x, x4.y), Xb = pack2(x4.z, x4.w);
                const unsigned long long Ya = pack2(y4.x, y4.y), Yb = pack2(y4.z, y4.w);
                const unsigned long long Za = pack2(z4.x, z4.y), Zb = pack2(z4.z, z4.w);
                const unsigned long long Wa = pack2(w4.x, w4.y), Wb = pack2(w4.z, w4.w);
                float sv[8];
                unpack2(fma2(AZ0, Za, fma2(AY0, Ya, fma2(AX0, Xa, Wa))), sv[0], sv[1]);
                unpack2(fma2(AZ0, Zb, fma2(AY0, Yb, fma2(AX0, Xb, Wb))), sv[2], sv[3]);
                unpack2(fma2(AZ1, Za, fma2(AY1, Ya, fma2(AX1, Xa, Wa))), sv[4], sv[5]);
                unpack2(fma2(AZ1, Zb, fma2(AY1, Yb, fma2(AX1, Xb, Wb))), sv[6], sv[7]);
                unsigned int m = 0u;  // bits 0..3: row a x targets k..k+3, bits 4..7: row b
#pragma unroll
                for (int b = 0; b < 4; b++) {
                  m |= (sv[b] < t0 ? 1u : 0u) << b;
                  m |= (sv[4 + b] < t1 ? 1u : 0u) << (4 + b);
                }
                acc |= m << (8 * it);
              }
            }
            hm[wq] = acc;
          }
          // ... then every lane drains ITS hits: the colour lower bound from the 8-bit summaries
          // (see eval_pair: a target whose bound already fails d2_color < d2_c_thres never becomes
          // a candidate) and the append to the row's cell.  With one pair in six passing the
          // geometric cut, testing colour per hit costs a third of testing it per pair, and a lane's
          // hit count varies little over a chunk (no warp-wide "any lane hit" coupling).
#pragma unroll
          for (int wq = 0; wq < kTileChunk / 16; wq++) {
            uint32_t m = hm[wq];
            while (m) {
              const int b = __ffs(m) - 1;
              m &= m - 1u;
              const int kk = 16 * wq + 4 * (b >> 3) + (b & 3);
              const bool second = (b & 4) != 0;
              if (colour_cut) {
                const unsigned int S2 = (second ? rna[1] : rna[0]) + S.nb[kk] -
                                        2u * __dp4a(second ? rqa[1] : rqa[0], S.col[kk], 0u);
                if (S2 >= s_thr) continue;
              }
              const uint32_t word = (S.idx[kk] << 8) | 1u;
              if (!second) {
                if (cnt0 < L) cell_a[cnt0] = word;
                cnt0++;
              } else {
                if (cnt1 < L) cell_b[cnt1] = word;
                cnt1++;
              }
            }
          }
        }
      }
    }
    if (lane < nrows) A.cand_cnt[(size_t)(row0 + lane) * P + part] = cnt0;
    if (lane + 32 < nrows) A.cand_cnt[(size_t)(row0 + lane + 32) * P + part] = cnt1;
    if (A.stamps && lane == 0) {  // debug (CVO_B200_STAMPS=1): targets swept per tile
      atomicAdd(&A.stamps[32000], n_swept * 4ull);
      atomicMax(&A.stamps[32001], n_swept * 4ull);
      atomicAdd(&A.stamps[32002], 1ull);
    }
  }
}

__global__ void __launch_bounds__(kPairWarps * 32, 3) tile_kernel(IterArgs A) {
  DevState* st = A.st;
  __shared__ TileSmemWarp smem[kPairWarps];
  __shared__ uint32_t s_hot[kHot1Words];
  for (int i = threadIdx.x; i < kHot1Words; i += blockDim.x)
    s_hot[i] = __ldcg(reinterpret_cast<const uint32_t*>(st) + i);
  __syncthreads();
  const DevState* hs = reinterpret_cast<const DevState*>(s_hot);
  if (hs->done) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  TileSmemWarp& S = smem[warp];
  __shared__ uint64_t bars[kPairWarps];
  if (lane == 0) {
    mbar_init(&bars[warp], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t bar_phase = 0u;
  // warps of one block take tiles that are far apart in the Morton order
  tile_phase(A, hs, S, &bars[warp], warp * gridDim.x + blockIdx.x, gridDim.x * kPairWarps, bar_phase);
}

// ================================================================== flow_kernel
// Two candidate generators feed the same exact per-pair arithmetic:
//   kGrid = false  the ordered candidate cells written by pair_kernel (dense scan);
//   kGrid = true   cell queries: the source row is mapped into the target's own frame
//                  (q = R x + T), the cube cells of the target's linear octree that the ball
//                  |y - q| <= r_i touches (<= 3 per axis, level chosen per row) are looked up as
//                  contiguous ranges of the Morton-ordered target, and every point of those
//                  ranges gets the exact test.  Nothing of the iteration is O(N*M) then.
// Eight lanes per source row (four rows per warp): candidates are ~10 per row in tracking
// regimes, so a full warp per row would idle two thirds of its lanes and quadruple the
// per-row bookkeeping.  Every group walks its row's candidate cells in target order, expands
// the candidate words into a small shared-memory list and drains it eight candidates at a time.
constexpr int kGroup = 8;                    // lanes per source row
constexpr int kRowsPerWarp = 32 / kGroup;    // 4
constexpr int kGroupList = 80;               // pending (<8) + one batch of 8 words (<=64)
constexpr int kGridList = 48;                // cell queries: pending (<8) + one quad pass (<=32)

// Exact redo of ONE source row by one warp: all targets in the caller's original order with the
// reference's arithmetic (the literal loop of CvoGPU.cu:524-591), for rows that were cut at their
// cap in a Morton-ordered candidate walk.  Adds the row's contribution to f (lane 0):
// omega[3], v[3], a_sum, nnz, max.  ELL indices are Morton positions (tgt_inv), like the rest of
// the matrix in these modes.
template <bool kFly>
__device__ __forceinline__ void redo_row(const IterArgs& A, const KernConsts& kc, const float* Ri,
                                         const float* Ti, float ell_now, int cap, int row, int lane,
                                         double (&f)[9]) {
  const float c_div = kc.c_div, d_div = kc.d_div;
  const unsigned lt32 = (1u << lane) - 1u;
    const int ig = A.row_begin + row;
    RowCtx rc;
    {
      const float4 pa = A.src_xyz[ig];
      rc.px[0] = pa.x; rc.px[1] = pa.y; rc.px[2] = pa.z;
      rc.qa = __float_as_uint(pa.w);
      if (kFly) {
        rc.l = range_ell(ell_now, A.src_rowA[ig].w);
        rc.d2_thres = -2.0 * rc.l * rc.l * kc.log_geo;
      } else {
        const float2 lt = A.row_lt[row];
        rc.l = lt.x;
        rc.d2_thres = lt.y;
      }
      rc.ga[0] = rc.ga[1] = 0.f;
      if (kc.use_geo_type) {
        const float2 gg = A.src_geo[ig];
        rc.ga[0] = gg.x; rc.ga[1] = gg.y;
      }
    }
    float om[3] = {0.f, 0.f, 0.f}, vv[3] = {0.f, 0.f, 0.f};
    double asum = 0.0;
    int count = 0;
    uint32_t* out_idx = A.ell_idx + (size_t)row * A.cap_max;
    float* out_val = A.ell_val + (size_t)row * A.cap_max;
    for (int jb = 0; jb < A.M && count < cap; jb += 32) {
      const int j = jb + lane;
      float a = 0.f;
      float4 pb = make_float4(0.f, 0.f, 0.f, 0.f);
      bool surv = false;
      if (j < A.M) {
        pb = move_target(A, Ri, Ti, A.tv[1].xyz[j]);  // same arithmetic as prep_kernel / the edge's pose
        surv = eval_pair(A, kc, rc, ig, 1, j, pb, a);
      }
      const unsigned mask = __ballot_sync(0xffffffffu, surv);
      const int pos = count + __popc(mask & lt32);
      if (surv && pos < cap) {
        out_idx[pos] = (uint32_t)A.tgt_inv[j];  // tgt_moved of this iteration is in Morton order
        out_val[pos] = a;
        const float py[3] = {pb.x, pb.y, pb.z};
        float cr[3];
        cross3f(rc.px, py, cr);
#pragma unroll
        for (int q = 0; q < 3; q++) {
          om[q] = om[q] + cr[q] * a;
          vv[q] = vv[q] + (py[q] - rc.px[q]) * a;
        }
        asum += (double)a;
      }
      count = min(cap, count + __popc(mask));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int q = 0; q < 3; q++) {
        om[q] += __shfl_xor_sync(0xffffffffu, om[q], o);
        vv[q] += __shfl_xor_sync(0xffffffffu, vv[q], o);
      }
      asum += __shfl_xor_sync(0xffffffffu, asum, o);
    }
    if (lane == 0) {
      A.row_nnz[row] = (uint32_t)count;
#pragma unroll
      for (int q = 0; q < 3; q++) {
        f[q] += (double)(om[q] / c_div);
        f[3 + q] += (double)(vv[q] / d_div);
      }
      f[6] += asum;
      f[7] += (double)count;
      f[8] = fmax(f[8], (double)count);
    }
}

// IterArgs::brute == 2 - the exact walk of redo_row by a TEAM of kBruteTeam warps per source row.
// One warp's walk over all M targets is a serial chain of M / 32 passes, each ~1 us of dependent
// double-precision latency (the three exps of a surviving pair), and on the README demo (523 rows,
// 1 080 targets, rows cut at their cap) that chain IS the iteration.  Here warp t of a team takes
// the t-th contiguous part of the targets (caller's order), evaluates it with the reference's
// arithmetic and parks its survivors (target, a) in shared memory in target order; after a block
// barrier the survivors' positions in the row follow from the team's counts, and every warp stores
// and accumulates those of ITS survivors that are among the row's first `cap` in target order -
// the reference's first-K truncation (CvoGPU.cu:524-526), exactly.  A warp stops early once it
// holds `cap` survivors itself.  Every thread of the block must call it (block barriers); f (lane 0
// of every warp) as in redo_row.  s_raw: >= 8 * brute_cap bytes per team warp; s_cnt: 32 ints.
__device__ __forceinline__ void brute_team_rows(const IterArgs& A, const DevState* hs, unsigned char* s_raw,
                                                int* s_cnt, double (&f)[9]) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int teams = (int)(blockDim.x >> 5) / kBruteTeam;
  const int team = w / kBruteTeam, tw = w - team * kBruteTeam;
  const bool in_team = team < teams;
  const int E = A.brute_cap;
  uint32_t* buf_j = reinterpret_cast<uint32_t*>(s_raw) + (size_t)(in_team ? w : 0) * E;
  float* buf_a = reinterpret_cast<float*>(s_raw) + (size_t)(teams * kBruteTeam) * E + (size_t)(in_team ? w : 0) * E;
  const KernConsts& kc = hs->kc;
  const int cap = hs->num_neighbors;
  const float c_div = kc.c_div, d_div = kc.d_div;
  const unsigned lt32 = (1u << lane) - 1u;
  const int Q = ((A.M + kBruteTeam - 1) / kBruteTeam + 31) & ~31;  // targets per team warp
  const int j0 = tw * Q, j1 = min(A.M, j0 + Q);
  for (int base = 0; base < A.n_rows; base += (int)gridDim.x * teams) {  // block-uniform trip count
    const int row = base + team * (int)gridDim.x + (int)blockIdx.x;
    const bool rvalid = in_team && row < A.n_rows;
    const int ig = A.row_begin + (rvalid ? row : 0);
    RowCtx rc;
    int cnt = 0;
    if (rvalid) {
      const float4 pa = A.src_xyz[ig];
      rc.px[0] = pa.x; rc.px[1] = pa.y; rc.px[2] = pa.z;
      rc.qa = __float_as_uint(pa.w);
      rc.l = range_ell(hs->ell, A.src_rowA[ig].w);
      rc.d2_thres = -2.0 * rc.l * rc.l * kc.log_geo;
      rc.ga[0] = rc.ga[1] = 0.f;
      if (kc.use_geo_type) {
        const float2 gg = A.src_geo[ig];
        rc.ga[0] = gg.x; rc.ga[1] = gg.y;
      }
      for (int jb = j0; jb < j1 && cnt < cap; jb += 32) {
        const int j = jb + lane;
        float a = 0.f;
        bool surv = false;
        if (j < j1) {
          const float4 pb = move_target(A, hs->Rinv, hs->Tinv, A.tv[1].xyz[j]);
          surv = eval_pair(A, kc, rc, ig, 1, j, pb, a);
        }
        const unsigned mask = __ballot_sync(0xffffffffu, surv);
        const int e = cnt + __popc(mask & lt32);
        if (surv && e < cap && e < E) {
          buf_j[e] = (uint32_t)j;
          buf_a[e] = a;
        }
        cnt = min(cap, cnt + __popc(mask));
      }
      cnt = min(cnt, E);
    }
    if (lane == 0) s_cnt[w] = cnt;
    __syncthreads();
    if (rvalid) {
      int prefix = 0, total = 0;
#pragma unroll
      for (int t = 0; t < kBruteTeam; t++) {
        const int c = s_cnt[team * kBruteTeam + t];
        if (t < tw) prefix += c;
        total += c;
      }
      total = min(total, cap);
      float om[3] = {0.f, 0.f, 0.f}, vv[3] = {0.f, 0.f, 0.f};
      double asum = 0.0;
      uint32_t* out_idx = A.ell_idx + (size_t)row * A.cap_max;
      float* out_val = A.ell_val + (size_t)row * A.cap_max;
      __syncwarp();  // the warp's own buffer writes
      for (int e = lane; e < cnt; e += 32) {
        const int pos = prefix + e;
        if (pos < cap) {
          const int j = (int)buf_j[e];
          const float a = buf_a[e];
          const float4 pb = move_target(A, hs->Rinv, hs->Tinv, A.tv[1].xyz[j]);
          out_idx[pos] = (uint32_t)A.tgt_inv[j];  // Morton positions, like the rest of the matrix
          out_val[pos] = a;
          const float py[3] = {pb.x, pb.y, pb.z};
          float cr[3];
          cross3f(rc.px, py, cr);  // compute_flow_gpu_no_eigen, CvoGPU.cu:765-781
#pragma unroll
          for (int q = 0; q < 3; q++) {
            om[q] = om[q] + cr[q] * a;
            vv[q] = vv[q] + (py[q] - rc.px[q]) * a;
          }
          asum += (double)a;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q = 0; q < 3; q++) {
          om[q] += __shfl_xor_sync(0xffffffffu, om[q], o);
          vv[q] += __shfl_xor_sync(0xffffffffu, vv[q], o);
        }
        asum += __shfl_xor_sync(0xffffffffu, asum, o);
      }
      if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 3; q++) {
          f[q] += (double)(om[q] / c_div);
          f[3 + q] += (double)(vv[q] / d_div);
        }
        f[6] += asum;
        if (tw == 0) {
          A.row_nnz[row] = (uint32_t)total;
          f[7] += (double)total;
          f[8] = fmax(f[8], (double)total);
        }
      }
    }
    __syncthreads();  // buffers and counts are reused by the next round
  }
}

// The source rows of this block (grid-stride over 8-lane row groups): candidates -> exact pair
// arithmetic -> ELL rows + per-row flow; returns the warp's partial sums in bp (valid on lane 0):
// omega[3], v[3], a_sum, nnz, max row count.  hs = this block's shared-memory copy of the hot
// state; st = the global state (saturation counters only).
// kGen: 0 = candidate cells of pair_kernel (moved targets / row thresholds from prep_kernel),
//       1 = cell queries, 2 = candidate cells of tile_kernel (Morton view; moved targets and row
//       thresholds computed on the fly like the cell queries: no prep launch)
template <int kGen, bool kColour = true>
__device__ __forceinline__ void flow_rows(const IterArgs& A, DevState* st, const DevState* hs,
                                          uint32_t* list, double (&bp)[9], double* queued = nullptr) {
  constexpr bool kGrid = (kGen == 1);
  constexpr bool kFly = (kGen != 0);
  const int view = kFly ? 0 : hs->view;  // cell queries / tile cells index the Morton-ordered target
  const float* s_pose = hs->Rinv;          // Rinv[9], Tinv[3] are contiguous
  const float ell_now = hs->ell;
  const float g_smax = hs->smax, g_slack = hs->grid_slack + A.edge_slack;
  unsigned int g_lb_thr = 0xffffffffu;  // integer threshold of the colour lower bound (stage 1)
  if (kColour && hs->kc.use_intensity) {
    const float th = hs->kc.d2_c_thres;
    g_lb_thr = (th > 0.f) ? (th * (65025.f / 0.999f) < 4.0e9f ? (unsigned int)ceilf(th * (65025.f / 0.999f)) : 0xffffffffu) : 0u;
  }

  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  const int g = lane >> 3;          // group inside the warp
  const int gl = lane & 7;          // lane inside the group
  const int gshift = g * kGroup;
  const unsigned lt8 = (1u << gl) - 1u;
  const KernConsts& kc = hs->kc;
  const int cap = hs->num_neighbors;
  // Cell queries visit a row's candidates in Morton order, so a row that was cut at its cap
  // holds the wrong survivors and is redone in original target order by the tail.  Counting one
  // survivor PAST the cap tells a complete row with exactly `cap` survivors (nothing to redo;
  // common when the cap has adapted down to 1.2 * max row count = a handful) from a cut one.
  const int cap_stop = (kGrid || view == 0) ? cap + 1 : cap;  // both Morton-ordered walks
  const float c_div = kc.c_div, d_div = kc.d_div;  // divisors (CvoGPU.cu:785-788)
  const int L = (kGen == 2) ? A.tile_L : A.L;
  const int nch = (kGen == 2) ? A.tile_parts : A.nchunks;
  if (blockIdx.x == 0 && threadIdx.x == 0) st->dbg[0] = gtime();

  double w_om[3] = {0, 0, 0}, w_v[3] = {0, 0, 0}, w_asum = 0.0;  // group sums (leader lane)
  double w_nnz = 0.0, w_max = 0.0;
  double w_sat = 0.0, w_capped = 0.0;  // rows this group queued for the exact redo / found capped

  // Row groups: the first pass is static (warp w of block b takes rows 4(b*W+w)..+3); when there
  // are more rows than resident row slots the rest is handed out dynamically, four
  // rows per warp from a global counter: rows differ a lot in cost in the dense regimes (a cloud's
  // boundary rows see half the neighbours) and a static split left 20 % of C4's flow phase idle.
  const int total_slots = gridDim.x * warps_per_block * kRowsPerWarp;
  const bool dynamic_rows = A.n_rows > total_slots;
  // (A.row_spread: few rows - the groups are dealt warp-major, a few busy warps on every SM; step_rows
  // uses the same map, so a warp reads back the ELL rows it wrote itself)
  int row_base = (A.row_spread ? warp_in_block * (int)gridDim.x + (int)blockIdx.x
                               : (int)blockIdx.x * warps_per_block + warp_in_block) * kRowsPerWarp;
  for (bool first_pass = true;; first_pass = false) {
    if (!first_pass) {
      if (dynamic_rows) {
        int nb = 0;
        if (lane == 0)
          nb = total_slots + (int)(atomicAdd(&st->work_counter, (unsigned)kRowsPerWarp) - hs->work_base);
        row_base = __shfl_sync(0xffffffffu, nb, 0);
      } else {
        row_base += total_slots;
      }
    }
    const int row = row_base + g;
    const bool rvalid = row < A.n_rows;
    if (!__any_sync(0xffffffffu, rvalid)) break;
    const int ig = A.row_begin + (rvalid ? row : 0);
    RowCtx rc;
    {
      const float4 pa = A.src_xyz[ig];
      rc.px[0] = pa.x; rc.px[1] = pa.y; rc.px[2] = pa.z;
      rc.qa = __float_as_uint(pa.w);
      if (kFly) {  // what prep_kernel writes to row_lt (CvoGPU.cu:506-511)
        rc.l = range_ell(ell_now, A.src_rowA[ig].w);
        rc.d2_thres = -2.0 * rc.l * rc.l * kc.log_geo;
      } else {
        const float2 lt = A.row_lt[rvalid ? row : 0];
        rc.l = lt.x;
        rc.d2_thres = lt.y;
      }
      rc.ga[0] = rc.ga[1] = 0.f;
      if (kc.use_geo_type) {
        const float2 gg = A.src_geo[ig];
        rc.ga[0] = gg.x; rc.ga[1] = gg.y;
      }
    }
    float om[3] = {0.f, 0.f, 0.f}, vv[3] = {0.f, 0.f, 0.f};
    double asum = 0.0;
    int count = 0;   // survivors stored so far (group-uniform)
    int nlist = 0;   // pending candidates in the group's list (group-uniform)
    uint32_t* out_idx = A.ell_idx + (size_t)(rvalid ? row : 0) * A.cap_max;
    float* out_val = A.ell_val + (size_t)(rvalid ? row : 0) * A.cap_max;
    const uint32_t* cnt_row = A.cand_cnt + (size_t)(rvalid ? row : 0) * nch;
    const uint32_t* cell_row = A.cand + (size_t)(rvalid ? row : 0) * nch * (size_t)L;

    // <=8 candidates per group -> exact test, ordered capped store, flow accumulation
    auto consume8 = [&](bool valid, int j) {
      float a = 0.f;
      float4 pb = make_float4(0.f, 0.f, 0.f, 0.f);
      bool surv = false;
      if (valid) {
        pb = kFly ? move_target(A, s_pose, s_pose + 9, A.tv[0].xyz[j]) : A.tgt_moved[j];
        surv = eval_pair(A, kc, rc, ig, view, j, pb, a);
      }
      const unsigned bits = (__ballot_sync(0xffffffffu, surv) >> gshift) & 0xffu;
      const int pos = count + __popc(bits & lt8);
      if (surv && pos < cap) {
        out_idx[pos] = (uint32_t)j;
        out_val[pos] = a;
        // compute_flow_gpu_no_eigen, CvoGPU.cu:765-781
        const float py[3] = {pb.x, pb.y, pb.z};
        float cr[3];
        cross3f(rc.px, py, cr);
#pragma unroll
        for (int k = 0; k < 3; k++) {
          om[k] = om[k] + cr[k] * a;
          vv[k] = vv[k] + (py[k] - rc.px[k]) * a;
        }
        asum += (double)a;
      }
      count = min(cap_stop, count + __popc(bits));
    };
    // drain the list while some group holds at least `need` pending candidates
    auto drain = [&](int need) {
      int head = 0;
      while (true) {
        const bool go = (nlist - head) >= need && (nlist - head) > 0 && count < cap_stop;
        if (!__any_sync(0xffffffffu, go)) break;
        const bool v = go && (head + gl) < nlist;
        consume8(v, v ? (int)list[head + gl] : 0);
        if (go) head += kGroup;
      }
      // compact what is left to the front (at most 7 entries, group-uniform branch)
      const int rest = max(0, nlist - head);
      uint32_t keep = 0;
      if (gl < rest) keep = list[head + gl];
      __syncwarp();
      if (gl < rest) list[gl] = keep;
      nlist = (count < cap_stop) ? rest : 0;
      __syncwarp();
    };
    // append the candidates of one word per lane (ascending target order) to the list
    auto append_words = [&](bool valid, uint32_t word) {
      const uint32_t qm = valid ? (word & 0xffu) : 0u;
      const int nb = __popc(qm);
      int incl = nb;
#pragma unroll
      for (int o = 1; o < kGroup; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o, kGroup);
        if (gl >= o) incl += t;
      }
      const int total = __shfl_sync(0xffffffffu, incl, kGroup - 1, kGroup);
      int off = nlist + incl - nb;
      const int jb8 = (int)(word >> 8);
      uint32_t m = qm;
      while (m) {
        const int q = __ffs(m) - 1;
        m &= m - 1;
        list[off++] = (uint32_t)(jb8 + q);
      }
      nlist += total;
      __syncwarp();
    };

    if constexpr (kGrid) {
      // ---- the cube cells of the target's octree that the ball around q = R x + T touches
      const GridView& G = A.gv;
      int nx = 0, ny = 0, ncell = 0, lvl = 0;
      int icx0 = 0, icy0 = 0, icz0 = 0;              // first cell of the block (cell units)
      float qx = 0.f, qy = 0.f, qz = 0.f, rq2 = -1.f;  // the row in the target's frame, query radius^2
      float fq0 = 0.f, fq1 = 0.f, fq2 = 0.f, frq2 = 0.f;  // ball centre / radius^2 in lattice units
      unsigned long long kx0 = 0, ky0 = 0, kz0 = 0;  // dilated coordinates of the first cell
      const unsigned long long mx = 0x1249249249249249ull;
      if (rvalid && rc.d2_thres > 0.f && cap > 0) {
        const float* Rf = hs->R;
        const float* Tf = hs->T;
        float q[3];
        mat3f_vec(Rf, rc.px, q);
        q[0] += Tf[0]; q[1] += Tf[1]; q[2] += Tf[2];
        // |y - q| <= smax |y' - x| + slack (see update_tf_device); d2 < thres in float
        const float rq = sqrtf(rc.d2_thres) * g_smax * 1.00001f + g_slack +
                         2e-6f * (fabsf(rc.px[0]) + fabsf(rc.px[1]) + fabsf(rc.px[2]) + fabsf(q[0]) +
                                  fabsf(q[1]) + fabsf(q[2]));
        const float top = 2097151.f;
        const float fx0 = (q[0] - rq - G.lo[0]) * G.scale, fx1 = (q[0] + rq - G.lo[0]) * G.scale;
        const float fy0 = (q[1] - rq - G.lo[1]) * G.scale, fy1 = (q[1] + rq - G.lo[1]) * G.scale;
        const float fz0 = (q[2] - rq - G.lo[2]) * G.scale, fz1 = (q[2] + rq - G.lo[2]) * G.scale;
        // a NaN anywhere makes every comparison false: no candidates, like d2 < thres
        const bool hit = fx1 >= -2.f && fy1 >= -2.f && fz1 >= -2.f && fx0 <= top + 2.f &&
                         fy0 <= top + 2.f && fz0 <= top + 2.f;
        if (hit) {
          // one lattice unit of padding on both sides covers the float rounding of f*0/f*1
          const int ix0 = (int)fminf(fmaxf(floorf(fx0) - 1.f, 0.f), top);
          const int iy0 = (int)fminf(fmaxf(floorf(fy0) - 1.f, 0.f), top);
          const int iz0 = (int)fminf(fmaxf(floorf(fz0) - 1.f, 0.f), top);
          const int ix1 = (int)fminf(fmaxf(floorf(fx1) + 1.f, 0.f), top);
          const int iy1 = (int)fminf(fmaxf(floorf(fy1) + 1.f, 0.f), top);
          const int iz1 = (int)fminf(fmaxf(floorf(fz1) + 1.f, 0.f), top);
          const int span = max(ix1 - ix0, max(iy1 - iy0, iz1 - iz0));
          lvl = span <= 2 ? 0 : (31 - __clz(span)) - 1;
          lvl = max(lvl, 21 - G.cbits);  // never finer than the coarse table (cell_ranges2)
          while (((ix1 >> lvl) - (ix0 >> lvl)) > 2 || ((iy1 >> lvl) - (iy0 >> lvl)) > 2 ||
                 ((iz1 >> lvl) - (iz0 >> lvl)) > 2)
            lvl++;
          nx = (ix1 >> lvl) - (ix0 >> lvl) + 1;
          ny = (iy1 >> lvl) - (iy0 >> lvl) + 1;
          ncell = nx * ny * ((iz1 >> lvl) - (iz0 >> lvl) + 1);
          icx0 = ix0 >> lvl; icy0 = iy0 >> lvl; icz0 = iz0 >> lvl;
          kx0 = spread21_dev((unsigned)icx0);
          ky0 = spread21_dev((unsigned)icy0);
          kz0 = spread21_dev((unsigned)icz0);
          fq0 = (q[0] - G.lo[0]) * G.scale;
          fq1 = (q[1] - G.lo[1]) * G.scale;
          fq2 = (q[2] - G.lo[2]) * G.scale;
          qx = q[0]; qy = q[1]; qz = q[2];
          rq2 = rq * rq * 1.000001f;
          const float frq = rq * G.scale + 3.f;  // + the padding of the box above
          frq2 = frq * frq * 1.00001f;
        }
      }
      // first key of cell number `cell` (= cxi + nx (cyi + ny czi)) of the row's cell block
      auto cell_key = [&](int cell, bool& touches) -> unsigned long long {
        // nx, ny in {1,2,3}, cell < 27: divisions by table
        const int t2 = nx == 1 ? cell : (nx == 2 ? (cell >> 1) : ((cell * 22) >> 6));
        const int cxi = cell - t2 * nx;
        const int czi = ny == 1 ? t2 : (ny == 2 ? (t2 >> 1) : ((t2 * 22) >> 6));
        const int cyi = t2 - czi * ny;
        // dilated increments ((k | ~mask) + 1) & mask instead of a bit spread per cell
        unsigned long long sx = kx0, sy = ky0, sz = kz0;
        if (cxi >= 1) sx = ((sx | ~mx) + 1ull) & mx;
        if (cxi >= 2) sx = ((sx | ~mx) + 1ull) & mx;
        if (cyi >= 1) sy = ((sy | ~mx) + 1ull) & mx;
        if (cyi >= 2) sy = ((sy | ~mx) + 1ull) & mx;
        if (czi >= 1) sz = ((sz | ~mx) + 1ull) & mx;
        if (czi >= 2) sz = ((sz | ~mx) + 1ull) & mx;
        // cells of the block that the ball does not reach (its corners, mostly) are skipped
        const float ch = (float)(1 << lvl);
        const float bx = (float)((icx0 + cxi) << lvl), by = (float)((icy0 + cyi) << lvl),
                    bz = (float)((icz0 + czi) << lvl);
        const float ex = fmaxf(0.f, fmaxf(bx - fq0, fq0 - (bx + ch)));
        const float ey = fmaxf(0.f, fmaxf(by - fq1, fq1 - (by + ch)));
        const float ez = fmaxf(0.f, fmaxf(bz - fq2, fq2 - (bz + ch)));
        touches = (ex * ex + ey * ey + ez * ez) <= frq2;
        return (sx | (sy << 1) | (sz << 2)) << (3 * lvl);
      };
      for (int cbase = 0;; cbase += 2 * kGroup) {
        const bool wact = cbase < ncell && count < cap_stop;
        if (!__any_sync(0xffffffffu, wact)) break;
        // ---- two cells per lane: their ranges of the Morton-ordered target.  The order in
        //      which a row's candidates are visited is free (a cut row is redone anyway).
        uint32_t start0 = 0, len0 = 0, start1 = 0, len1 = 0;
        {
          const int c0 = cbase + gl, c1 = cbase + kGroup + gl;
          bool v0 = wact && c0 < ncell, v1 = wact && c1 < ncell;
          bool t0 = false, t1 = false;
          const unsigned long long k0 = v0 ? cell_key(c0, t0) : 0ull;
          const unsigned long long k1 = v1 ? cell_key(c1, t1) : 0ull;
          v0 = v0 && t0;
          v1 = v1 && t1;
          cell_ranges2(G, 3 * lvl, v0, k0, v1, k1, start0, len0, start1, len1);
        }
        // ---- flatten in units of QUADS (4 consecutive targets of one range): a lane tests four
        //      points per pass, so the owner search is amortised and the four loads overlap
        const uint32_t nq0 = (len0 + 3u) >> 2, nq1 = (len1 + 3u) >> 2;
        const uint32_t nq = nq0 + nq1;
        uint32_t incl = nq;
#pragma unroll
        for (int o = 1; o < kGroup; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o, kGroup);
          if (gl >= o) incl += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, kGroup - 1, kGroup);
        const uint32_t excl = incl - nq;
        // ---- stage 1: the geometric cut alone on every point of the ranges (cheap); the
        //      survivors are queued so that the full kernel (colour, semantics, double exp)
        //      runs on packed lanes (stage 2 = drain -> consume8)
        for (uint32_t wb = 0;; wb += kGroup) {
          const bool bact = wb < total && count < cap_stop;
          if (!__any_sync(0xffffffffu, bact)) break;
          const uint32_t b = wb + gl;
          const bool valid = bact && b < total;
          int o = 0;  // owner lane = number of lanes whose inclusive count is <= b
#pragma unroll
          for (int stp = 4; stp > 0; stp >>= 1) {
            const uint32_t e = __shfl_sync(0xffffffffu, incl, (o + stp - 1) & 7, kGroup);
            if (e <= b) o += stp;
          }
          o = min(o, kGroup - 1);
          const uint32_t oex = __shfl_sync(0xffffffffu, excl, o, kGroup);
          const uint32_t oq0 = __shfl_sync(0xffffffffu, nq0, o, kGroup);
          const uint32_t ol0 = __shfl_sync(0xffffffffu, len0, o, kGroup);
          const uint32_t os0 = __shfl_sync(0xffffffffu, start0, o, kGroup);
          const uint32_t ol1 = __shfl_sync(0xffffffffu, len1, o, kGroup);
          const uint32_t os1 = __shfl_sync(0xffffffffu, start1, o, kGroup);
          const uint32_t rr = b - oex;
          const bool first = rr < oq0;
          const uint32_t jb = first ? os0 + 4u * rr : os1 + 4u * (rr - oq0);
          const uint32_t je = first ? os0 + ol0 : os1 + ol1;
          // stage-1 test in the target's OWN frame: |y - q|^2 <= rq^2 with q = R x + T and the
          // conservative query radius rq (every pair with the reference's d2 < d2_thres is inside,
          // see update_tf_device) - one load, three subtractions, three FMAs; the exact test of
          // the reference's arithmetic follows in stage 2
          auto test_slot = [&](uint32_t j, bool vt) -> bool {
            if (!vt) return false;
            const float4 y = A.tv[0].xyz[j];
            const float dx = y.x - qx, dy = y.y - qy, dz = y.z - qz;
            const bool near = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx)) <= rq2;
            if constexpr (!kColour) return near;
            if (!near) return false;
            // the point's packed colour summary rides in .w: drop targets whose colour-distance
            // LOWER bound already fails the reference's colour test (see eval_pair)
            const unsigned int d = __vsubus4(__vabsdiffu4(rc.qa, __float_as_uint(y.w)), 0x01010101u);
            return __dp4a(d, d, 0u) < g_lb_thr;
          };
          auto queue_slot = [&](uint32_t j, bool pass) {
            const unsigned bits = (__ballot_sync(0xffffffffu, pass) >> gshift) & 0xffu;
            if (pass) list[nlist + __popc(bits & lt8)] = j;
            nlist += __popc(bits);
          };
          // how many slots of its quad the busiest lane of the warp needs: short ranges (the
          // sparse regimes) take the one-slot path, long ranges four independent loads + tests
          const int need = valid ? (int)min(4u, je - jb) : 0;
          if (__reduce_max_sync(0xffffffffu, need) <= 1) {
            queue_slot(jb, test_slot(jb, need >= 1));
          } else {
            bool pass[4];
#pragma unroll
            for (int t = 0; t < 4; t++) pass[t] = test_slot(jb + (uint32_t)t, t < need);
#pragma unroll
            for (int t = 0; t < 4; t++) queue_slot(jb + (uint32_t)t, pass[t]);
          }
          __syncwarp();
          drain(kGroup);
        }
      }
    } else {
    if (kGen == 2) {
      // a tile cell that overflowed its word list (more candidates than tile_L): the row is
      // handled like one cut at its cap - exact redo in original target order
      bool ov = false;
      if (rvalid && cap > 0)
        for (int c = gl; c < nch; c += kGroup) ov |= cnt_row[c] > (uint32_t)L;
      if ((__ballot_sync(0xffffffffu, ov) >> gshift) & 0xffu) count = cap_stop;
    }
    for (int cbase = 0;; cbase += 4 * kGroup) {
      const bool wact = rvalid && cbase < nch && count < cap_stop;
      if (!__any_sync(0xffffffffu, wact)) break;
      // ---- four cell counts per lane: a window of 32 cells per group
      uint32_t c4[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int c = cbase + 4 * gl + k;
        c4[k] = (wact && c < nch) ? cnt_row[c] : 0u;
      }
      const bool over = (c4[0] > (uint32_t)L) | (c4[1] > (uint32_t)L) | (c4[2] > (uint32_t)L) |
                        (c4[3] > (uint32_t)L);
      if (!__any_sync(0xffffffffu, over)) {
        // ---- flatten the window's words: prefix over (lane, k)
        const uint32_t p1 = c4[0], p2 = p1 + c4[1], p3 = p2 + c4[2], s4 = p3 + c4[3];
        uint32_t incl = s4;
#pragma unroll
        for (int o = 1; o < kGroup; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o, kGroup);
          if (gl >= o) incl += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, kGroup - 1, kGroup);
        const uint32_t excl = incl - s4;
        for (uint32_t wb = 0;; wb += kGroup) {
          const bool bact = wb < total && count < cap_stop;
          if (!__any_sync(0xffffffffu, bact)) break;
          const uint32_t b = wb + gl;
          const bool valid = bact && b < total;
          // owner lane = number of lanes whose inclusive count is <= b
          int o = 0;
#pragma unroll
          for (int stp = 4; stp > 0; stp >>= 1) {
            const uint32_t e = __shfl_sync(0xffffffffu, incl, (o + stp - 1) & 7, kGroup);
            if (e <= b) o += stp;
          }
          o = min(o, kGroup - 1);
          const uint32_t oex = __shfl_sync(0xffffffffu, excl, o, kGroup);
          const uint32_t q1 = __shfl_sync(0xffffffffu, p1, o, kGroup);
          const uint32_t q2 = __shfl_sync(0xffffffffu, p2, o, kGroup);
          const uint32_t q3 = __shfl_sync(0xffffffffu, p3, o, kGroup);
          const uint32_t rr = b - oex;
          const int k = (rr >= q1) + (rr >= q2) + (rr >= q3);
          const uint32_t kbase = k == 0 ? 0u : (k == 1 ? q1 : (k == 2 ? q2 : q3));
          uint32_t word = 0;
          if (valid) word = cell_row[(size_t)(cbase + 4 * o + k) * (size_t)L + (rr - kbase)];
          if constexpr (kGen == 2) {
            // a tile-cell word names exactly ONE target (target << 8 | 1): eight words are eight
            // candidates, evaluated in place - no expansion into the group's list, no drain
            consume8(valid, (int)(word >> 8));
          } else {
            append_words(valid, word);
            drain(kGroup);
          }
        }
      } else {
        // ---- rare: a cell overflowed its word list -> walk the window cell by cell and
        //      rescan the overflowed chunks exhaustively (still in ascending target order)
        for (int cc = 0; cc < 4 * kGroup; cc++) {
          const uint32_t n = __shfl_sync(0xffffffffu, c4[cc & 3], cc >> 2, kGroup);
          const bool cact = wact && (cbase + cc) < nch && count < cap_stop;
          if (!__any_sync(0xffffffffu, cact && n > 0u)) continue;
          const bool is_over = cact && n > (uint32_t)L;
          if (__any_sync(0xffffffffu, is_over)) drain(1);  // flush: order before the rescan
          const int j_begin = (cbase + cc) * A.chunk_len;
          const int j_end = min(A.M, j_begin + A.chunk_len);
          const uint32_t* cell = cell_row + (size_t)(cbase + cc) * (size_t)L;
          for (uint32_t s = 0;; s += kGroup) {
            const bool list_go = cact && !is_over && s < n && count < cap_stop;
            const bool scan_go = is_over && (j_begin + (int)s) < j_end && count < cap_stop;
            if (!__any_sync(0xffffffffu, list_go | scan_go)) break;
            if (__any_sync(0xffffffffu, scan_go)) {
              const int j = j_begin + (int)s + gl;
              consume8(scan_go && j < j_end, j);
            }
            const bool v = list_go && (s + gl) < n;
            append_words(v, v ? cell[s + gl] : 0u);
            drain(kGroup);
          }
        }
      }
    }
    }
    drain(1);  // whatever is still pending
    // ---- row epilogue: omega_i / c, v_i / d in float, then double (CvoGPU.cu:785-788)
#pragma unroll
    for (int o = kGroup / 2; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        om[k] += __shfl_xor_sync(0xffffffffu, om[k], o, kGroup);
        vv[k] += __shfl_xor_sync(0xffffffffu, vv[k], o, kGroup);
      }
      asum += __shfl_xor_sync(0xffffffffu, asum, o, kGroup);
    }
    // A row that reached its cap: in the Morton view it may hold the wrong survivors (the
    // reference keeps the first ones in ORIGINAL target order), so it is queued for the exact
    // redo in this kernel's tail and contributes nothing here.
    const bool capped = rvalid && count >= cap_stop && !((kGrid || view == 0) && cap == 0);
    count = min(count, cap);
    if (capped && gl == 0) {
      if (view == 0) {
        A.sat_list[atomicAdd(&st->n_sat, 1u) - hs->sat_base] = (uint32_t)row;
        w_sat += 1.0;
      } else {
        atomicAdd(&st->n_capped, 1u);
        w_capped += 1.0;
      }
    }
    if (gl == 0 && rvalid && !(capped && view == 0)) {
      A.row_nnz[row] = (uint32_t)count;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        w_om[k] += (double)(om[k] / c_div);
        w_v[k] += (double)(vv[k] / d_div);
      }
      w_asum += asum;
      w_nnz += (double)count;
      w_max = fmax(w_max, (double)count);
    }
  }
  // ---- block partial: fixed xor tree over the four group leaders of a warp, then over warps
  bp[0] = w_om[0]; bp[1] = w_om[1]; bp[2] = w_om[2]; bp[3] = w_v[0]; bp[4] = w_v[1]; bp[5] = w_v[2];
  bp[6] = w_asum; bp[7] = w_nnz; bp[8] = w_max;
#pragma unroll
  for (int o = 8; o <= 16; o <<= 1) {
#pragma unroll
    for (int k = 0; k < 9; k++) {
      const double x = __shfl_xor_sync(0xffffffffu, bp[k], o);
      bp[k] = (k < 8) ? bp[k] + x : fmax(bp[k], x);
    }
    w_sat += __shfl_xor_sync(0xffffffffu, w_sat, o);
    w_capped += __shfl_xor_sync(0xffffffffu, w_capped, o);
  }
  if (queued) {
    queued[0] = w_sat;
    queued[1] = w_capped;
  }
}

template <int kGen>
__global__ void __launch_bounds__(kSparseThreads, 3) flow_kernel_t(IterArgs A) {
  constexpr bool kFly = (kGen != 0);
  DevState* st = A.st;
  __shared__ double sh[kSparseThreads * 9];
  __shared__ uint32_t s_hot[kHot1Words];  // pose, schedule, constants: one cooperative load
  __shared__ uint32_t s_list[kSparseThreads / kGroup][kGroupList];
  __shared__ bool is_last;
  unsigned long long* stamp = (A.stamps && threadIdx.x == 0) ? A.stamps + 8 * (size_t)blockIdx.x : nullptr;
  if (stamp) stamp[0] = gtime();
  for (int i = threadIdx.x; i < kHot1Words; i += blockDim.x)
    s_hot[i] = __ldcg(reinterpret_cast<const uint32_t*>(st) + i);
  __syncthreads();
  const DevState* hs = reinterpret_cast<const DevState*>(s_hot);
  if (hs->done) return;
  if (stamp) stamp[1] = gtime();
  double bp[9];
  flow_rows<kGen>(A, st, hs, s_list[threadIdx.x >> 3], bp);
  const int view = kFly ? 0 : hs->view;
  const float ell_now = hs->ell;
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  const KernConsts& kc = hs->kc;
  const int cap = hs->num_neighbors;
  const float c_div = kc.c_div, d_div = kc.d_div;
  if (stamp) stamp[2] = gtime();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 9; k++) sh[warp_in_block * 9 + k] = bp[k];
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    const int k = threadIdx.x;
    double r = sh[k];
    for (int w = 1; w < warps_per_block; w++) r = (k < 8) ? r + sh[w * 9 + k] : fmax(r, sh[w * 9 + k]);
    reinterpret_cast<double*>(A.flow_part)[(size_t)k * gridDim.x + blockIdx.x] = r;
  }
  __syncthreads();
  if (stamp) stamp[3] = gtime();
  if (threadIdx.x == 0) {
    __threadfence();  // release: this block's partial (written before the barrier above)
    if (stamp) stamp[4] = gtime();
    const unsigned int prev = atomicAdd(&st->flow_blocks_done, 1u);
    is_last = (prev == gridDim.x - 1);
    if (is_last) __threadfence();  // acquire for the whole block (readers use ld.cg after the barrier)
    if (stamp) stamp[5] = gtime();
  }
  __syncthreads();
  if (!is_last) return;
  // ---- last block: reduce all block partials in a fixed order
  if (threadIdx.x == 0) st->dbg[1] = gtime();
  double tot[9];
  block_reduce_partials<9, 8>(reinterpret_cast<const double*>(A.flow_part), (int)gridDim.x, tot, sh, &st->dbg[10]);
  // ---- exact redo of the rows that reached their cap in the Morton view: one warp per row scans
  //      ALL targets in the caller's original order with the reference's arithmetic (the literal
  //      loop of CvoGPU.cu:524-591).  Rare in tracking regimes; when it is not, the controller
  //      moves the run to the original-order view.
  const unsigned int n_sat = (view == 0) ? *(volatile unsigned int*)&st->n_sat : 0u;
  if (n_sat > 0u) {
    double f[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    float Ri[9], Ti[3];
#pragma unroll
    for (int q = 0; q < 9; q++) Ri[q] = st->Rinv[q];
#pragma unroll
    for (int q = 0; q < 3; q++) Ti[q] = st->Tinv[q];
    for (unsigned int si = warp_in_block; si < n_sat; si += warps_per_block)
      redo_row<kFly>(A, kc, Ri, Ti, ell_now, cap, (int)__ldcg(&A.sat_list[si]), lane, f);
    // fold the redone rows into the totals (fixed order over the warps of this block)
    __syncthreads();
    if (lane == 0) {
      double* d = sh + warp_in_block * 9;
#pragma unroll
      for (int q = 0; q < 9; q++) d[q] = f[q];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 0; w < warps_per_block; w++) {
        for (int q = 0; q < 8; q++) tot[q] += sh[w * 9 + q];
        tot[8] = fmax(tot[8], sh[w * 9 + 8]);
      }
    }
  }
  if (threadIdx.x == 0) {
    st->dbg[2] = gtime();
    st->flow_blocks_done = 0u;
    if (A.world > 1) {
      for (int k = 0; k < 9; k++) st->local_flow[k] = tot[k];
    } else {
      finalize_flow_scalar(st, tot);
    }
    st->dbg[3] = gtime();
  }
}

// ================================================================== step_kernel + controller
__device__ int indicator_update(DevState* st, float indicator, const cvo_b200_params* params) {
  // A_sparsity_indicator_ell_update, CvoGPU.cu:1167-1285, with ring buffers for the queues
  int decrease = 0;
  const int queue_len = params->indicator_window_size;
  if (st->qs_size < queue_len) {
    st->q_start[(st->qs_head + st->qs_size) % kQueueCap] = indicator;
    st->qs_size++;
    st->start_sum += indicator;
  }
  if (st->qs_size >= queue_len && st->qe_size < queue_len) {
    st->q_end[(st->qe_head + st->qe_size) % kQueueCap] = indicator;
    st->qe_size++;
    st->end_sum += indicator;
  }
  if (st->qs_size >= queue_len && st->qe_size >= queue_len) {
    if (st->end_sum / st->start_sum > 1 - params->indicator_stable_threshold &&
        st->end_sum / st->start_sum < 1 + params->indicator_stable_threshold) {
      decrease = 1;
      st->qs_head = st->qs_size = 0;
      st->qe_head = st->qe_size = 0;
      st->start_sum = 0;
      st->end_sum = 0;
    } else {
      const float e_front = st->q_end[st->qe_head];
      st->end_sum -= e_front;
      st->start_sum += e_front;
      st->q_start[(st->qs_head + st->qs_size) % kQueueCap] = e_front;
      st->qs_size++;
      st->qe_head = (st->qe_head + 1) % kQueueCap;
      st->qe_size--;
      st->start_sum -= st->q_start[st->qs_head];
      st->qs_head = (st->qs_head + 1) % kQueueCap;
      st->qs_size--;
      st->q_end[(st->qe_head + st->qe_size) % kQueueCap] = indicator;
      st->qe_size++;
      st->end_sum += indicator;
    }
  }
  return decrease;
}

// CvoGPU.cu:94-112: R_inv = R^T, T_inv = -R_inv * T; plus the bound on |y' - c| the next
// prep_kernel needs: |Rinv (y - tc) + (Rinv tc + Tinv - c)| <= sigma_max(Rinv)*trad + |...|,
// sigma_max^2 <= max row sum of |Rinv^T Rinv| (Gershgorin).
__device__ void update_tf_device(const IterArgs& A, DevState* st) {
  float neg[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) st->Rinv[3 * j + i] = st->R[3 * i + j];
  for (int k = 0; k < 9; k++) neg[k] = -st->Rinv[k];
  float tv[3];
  mat3f_vec(neg, st->T, tv);
  for (int k = 0; k < 3; k++) st->Tinv[k] = tv[k];
  // ---- conservative bounds for the candidate generators.  They sit on the critical path of every
  //      iteration, so they are evaluated in float (a quarter of the latency of the double
  //      chain); every float rounding (relative 2^-24 per operation, a handful of operations)
  //      is covered by the explicit inflation factors below.
  float G[9];  // Rinv^T Rinv = R R^T
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      G[3 * i + j] = st->Rinv[3 * i] * st->Rinv[3 * j] + st->Rinv[3 * i + 1] * st->Rinv[3 * j + 1] +
                     st->Rinv[3 * i + 2] * st->Rinv[3 * j + 2];
  float smax2 = 0.f;  // sigma_max^2 <= max row sum (Gershgorin)
  for (int i = 0; i < 3; i++)
    smax2 = fmaxf(smax2, fabsf(G[3 * i]) + fabsf(G[3 * i + 1]) + fabsf(G[3 * i + 2]));
  const float sm = sqrtf(smax2) * 1.00001f;
  // |y' - c| = |Rinv (y - tc) + (Rinv tc + Tinv - c)| <= sm * trad + |Rinv tc + Tinv - c|
  float off2 = 0.f;
  for (int i = 0; i < 3; i++) {
    const float o = st->Rinv[i] * A.tcx + st->Rinv[3 + i] * A.tcy + st->Rinv[6 + i] * A.tcz + st->Tinv[i];
    const float oc = fabsf(o - (i == 0 ? A.cx : (i == 1 ? A.cy : A.cz))) +
                     4e-7f * (fabsf(o) + fabsf(st->Tinv[i]) + sm * (fabsf(A.tcx) + fabsf(A.tcy) + fabsf(A.tcz)));
    off2 += oc * oc;
  }
  const float ymax = (sm * A.trad + sqrtf(off2)) * 1.00002f + 1e-6f;
  st->ymax2_bound = ymax * ymax * 1.000001f;
  st->smax = sm;
  // Slack of a cell query (flow_rows<true>).  With y' = fl(R^T y + Tinv), Tinv = fl(-R^T T),
  // q = fl(R x + T), E = R R^T - I and u = 2^-24:
  //   |y - q| <= smax |y' - x| + |E| (|y| + |T|) + 8u smax^2 (|y| + |T|) + 4u (smax |x| + |T|)
  // (the |x| term is added per row).  |y| <= |tc| + trad; |E|_2 <= |E|_F, each entry of G known
  // to 4u in float.  2x safety + 1 um.
  {
    float e2 = 0.f;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        const float d = fabsf(G[3 * i + j] - (i == j ? 1.f : 0.f)) + 3e-7f * smax2;
        e2 += d * d;
      }
    const float tn = sqrtf(st->T[0] * st->T[0] + st->T[1] * st->T[1] + st->T[2] * st->T[2]) * 1.000001f;
    const float yn = (sqrtf(A.tcx * A.tcx + A.tcy * A.tcy + A.tcz * A.tcz) + A.trad) * 1.000001f;
    const float u = 5.9604644775390625e-08f;
    st->grid_slack =
        (2.f * (sqrtf(e2) * (yn + tn) + 8.f * u * sm * sm * (yn + tn) + 4.f * u * tn) + 1e-6f) * 1.00001f;
  }
  // target view of the next iteration: leave the Morton view when many rows reach their cap
  // (each costs an O(M) exact redo), come back once no row does
  st->last_view = A.tile ? 0 : st->view;  // tile cells index the Morton view whatever st->view said
  st->last_grid = A.grid;
  int next_view = 1;
  if (st->prune_on) next_view = (st->view == 0) ? (st->n_sat > 16u ? 1 : 0) : (st->n_capped == 0u ? 0 : 1);
  if (A.tile) next_view = 0;  // tile cells exist in the Morton view only; cut rows are redone exactly
  st->view = next_view;
  st->sat_total += st->n_sat + st->n_capped;
  st->n_sat = 0u;
  st->n_capped = 0u;
}

// Everything align_impl does on the host after the reductions (CvoGPU.cu:1124-1158,
// 1452-1531).  It is a chain of dependent double-precision scalar work (cubic, Exp, log), i.e.
// pure latency, and it sits on the critical path of every iteration, so it is split over TWO
// warps of the calling block (threads 0 and 32 run on different schedulers):
//   phase 1   A: cubic -> clamped step            B: gradient test, indicator queues
//   phase 2   A: Exp_SEK3, pose update, update_tf B: Exp_SEK3 (same arithmetic), se(3) log norm,
//                                                    eps_2 test, ell decay, row-cap update
//   phase 3   A: iteration counter, stop flags, trace record
// Every thread of the block must call it (it contains block barriers); bcde (the global sums
// B, C, D, E) needs to be valid on thread 0 only.  `st` may live in shared or global memory.
struct CtrlScratch {
  int grad_small, need_decay, finished, flags;
  float ell_used;
  int cap_used;
  double dist;
};

__device__ void controller_step(const IterArgs& A, DevState* st, const double bcde[4], CtrlScratch* sc) {
  const cvo_b200_params* params = A.params;
  const int tid = threadIdx.x;
  // ---------------------------------------------------------------- phase 1
  if (tid == 0) {
    st->B = bcde[0]; st->C = bcde[1]; st->D = bcde[2]; st->E = bcde[3];
    // step size (CvoGPU.cu:1124-1158)
    const double coef[4] = {4.0 * bcde[3], 3.0 * bcde[2], 2.0 * bcde[1], bcde[0]};
    double re[3], im[3];
    double temp_step = 1.7976931348623157e308;  // numeric_limits<double>::max()
    if (cubic_roots(coef, re, im)) {
      for (int i = 0; i < 3; i++)
        if (re[i] > 0 && re[i] < temp_step && fabs(im[i]) < 1e-5) temp_step = re[i];
    }
    st->dbg[13] = gtime();
    float step;
    if (temp_step > params->max_step)
      step = params->max_step;
    else if (temp_step < params->min_step)
      step = params->min_step;
    else
      step = (float)temp_step;
    st->step = step;
  } else if (tid == 32) {
    sc->ell_used = st->ell;
    sc->cap_used = st->num_neighbors;
    sc->flags = 0;
    sc->finished = 0;
    sc->need_decay = 0;
    sc->dist = 0.0;
    const float* omega = st->omega;
    const float* v = st->v;
    // gradient test (CvoGPU.cu:1454-1458)
    const double on = sqrt(sum3d((double)omega[0] * omega[0], (double)omega[1] * omega[1],
                                 (double)omega[2] * omega[2]));
    const double vn = sqrt(sum3d((double)v[0] * v[0], (double)v[1] * v[1], (double)v[2] * v[2]));
    int gs = 0;
    if (on < params->eps && vn < params->eps && st->controller_on != 2) {  // mode 2: timing loop
      const float onf = sqrtf(dot3f(omega, omega)), vnf = sqrtf(dot3f(v, v));
      int reason = CVO_B200_STOP_GRAD_SMALL;
      if (onf < 1e-8 && vnf < 1e-8) {
        st->ret = -1;
        reason |= CVO_B200_STOP_GRAD_ZERO;
      }
      st->stop_reason = reason;
      sc->flags = reason;
      sc->finished = 1;
      gs = 1;
    } else if (st->controller_on == 1) {
      // indicator (CvoGPU.cu:1486-1493): does not depend on the step
      const float ip_curr =
          (float)((double)st->nnz / sqrt((double)A.n_src_total * (double)A.M));
      sc->need_decay = indicator_update(st, ip_curr, params);
    }
    sc->grad_small = gs;
  }
  __syncthreads();
  // ---------------------------------------------------------------- phase 2
  const bool moving = !sc->grad_small;
  if ((tid == 0 || tid == 32) && moving) {
    // pose update (CvoGPU.cu:1460-1476)
    const float vec_joined[6] = {st->omega[0], st->omega[1], st->omega[2], st->v[0], st->v[1], st->v[2]};
    // dist = |log(dRT)| with dRT the FLOAT-rounded Exp_SEK3(step * twist) (CvoGPU.cu:1473-1476) only
    // decides the eps_2 stop test.  In exact arithmetic log(Exp(step xi)) = step xi, and the float
    // rounding of the increment moves the norm by < 1e-6 (entries of R to 6e-8, the small-angle
    // terms far less), so whenever step |xi| is further than that from eps_2 the test is decided
    // without the double-precision Exp + quaternion log chain (~1.4 us on the critical path of every
    // iteration).  The exact value is still computed when it could matter or is recorded (trace).
    bool need_exact = true;
    double d_fast = 0.0;
    if (tid == 32) {
      double n2 = 0.0;
      for (int q = 0; q < 6; q++) n2 += (double)vec_joined[q] * (double)vec_joined[q];
      d_fast = (double)st->step * sqrt(n2);
      // (trace_cap, not the trace pointer: the persistent kernel clears the pointer in every block but
      //  block 0, and all blocks must take the same path here)
      const bool traced = st->iter < st->trace_cap;
      const double margin = 2e-6 + 1e-4 * d_fast;
      need_exact = traced || st->controller_on != 1 || !(fabs(d_fast - (double)params->eps_2) > margin) ||
                   !(d_fast < 1.0);  // (rotation angles near pi: the log is not step*xi any more)
    }
    float dtrans[12];
    double dR[9], dT[3];
    if (need_exact) {  // thread 0 always; both threads evaluate the same Exp_SEK3
      exp_sek3(vec_joined, st->step, dtrans);
      for (int q = 0; q < 9; q++) dR[q] = (double)dtrans[q];
      for (int q = 0; q < 3; q++) dT[q] = (double)dtrans[9 + q];
    }
    if (tid == 0) {
      double Rd[9], Td[3];
      for (int q = 0; q < 9; q++) Rd[q] = (double)st->R[q];
      for (int q = 0; q < 3; q++) Td[q] = (double)st->T[q];
      float Rn[9], Tn[3];
      for (int i = 0; i < 3; i++)
        Tn[i] = (float)(sum3d(Rd[i] * dT[0], Rd[3 + i] * dT[1], Rd[6 + i] * dT[2]) + Td[i]);
      for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++)
          Rn[3 * j + i] = (float)sum3d(Rd[i] * dR[3 * j], Rd[3 + i] * dR[3 * j + 1],
                                       Rd[6 + i] * dR[3 * j + 2]);
      // the trace reports the pose AFTER the update; the fixed-state timing loop (mode 2) then
      // keeps the pose where it was
      cvo_b200_iter_trace* tr = (st->trace && st->iter < st->trace_cap) ? &st->trace[st->iter] : nullptr;
      if (tr) {
        for (int q = 0; q < 9; q++) tr->R[q] = Rn[q];
        for (int q = 0; q < 3; q++) tr->T[q] = Tn[q];
      }
      if (st->controller_on != 2) {
        for (int q = 0; q < 9; q++) st->R[q] = Rn[q];
        for (int q = 0; q < 3; q++) st->T[q] = Tn[q];
      }
      st->dbg[14] = gtime();
      update_tf_device(A, st);  // next iteration's Rinv/Tinv and bounds (or the final transform)
    } else {
      const double dist_this_iter = need_exact ? se3_log_norm(dR, dT) : d_fast;
      st->dbg[15] = gtime();
      st->dist = dist_this_iter;
      sc->dist = dist_this_iter;
      if (st->controller_on == 1) {
        if (dist_this_iter < params->eps_2) {  // :1505-1508
          st->stop_reason = CVO_B200_STOP_DIST_SMALL;
          sc->flags = CVO_B200_STOP_DIST_SMALL;
          sc->finished = 1;
        } else {
          if (st->iter > params->ell_decay_start && sc->need_decay) {  // :1509-1513
            st->ell = st->ell * params->ell_decay_rate;
            if (st->ell < params->ell_min) st->ell = params->ell_min;
            sc->flags |= CVO_B200_ELL_DECAYED;
          }
          // :1518-1529
          const int cand = (int)(st->max_row_nnz * 1.2);
          st->num_neighbors = params->nearest_neighbors_max < cand ? params->nearest_neighbors_max : cand;
        }
      }
    }
  }
  __syncthreads();
  // ---------------------------------------------------------------- phase 3
  if (tid == 0) {
    const int k = st->iter;
    cvo_b200_iter_trace* tr = (st->trace && k < st->trace_cap) ? &st->trace[k] : nullptr;
    bool finished = sc->finished != 0;
    if (!moving) update_tf_device(A, st);  // gradient vanished: pose untouched, final transform
    int cap_next = st->num_neighbors;
    if (st->controller_on == 0) {
      const int cand = (int)(st->max_row_nnz * 1.2);
      cap_next = params->nearest_neighbors_max < cand ? params->nearest_neighbors_max : cand;
      finished = true;
    }
    if (tr) {
      tr->iter = k;
      tr->num_neighbors = sc->cap_used;
      tr->ell = sc->ell_used;
      tr->max_row_nnz = st->max_row_nnz;
      tr->nnz = st->nnz;
      for (int q = 0; q < 3; q++) {
        tr->omega_sum[q] = st->omega_sum[q];
        tr->v_sum[q] = st->v_sum[q];
        tr->omega[q] = st->omega[q];
        tr->v[q] = st->v[q];
      }
      tr->B = st->B; tr->C = st->C; tr->D = st->D; tr->E = st->E;
      tr->step = st->step;
      tr->flags = sc->flags;
      tr->dist = sc->dist;
      tr->a_sum = st->a_sum;
      for (int q = 0; q < 6; q++) tr->reserved[q] = 0;
      if (!moving) {  // pose untouched
        for (int q = 0; q < 9; q++) tr->R[q] = st->R[q];
        for (int q = 0; q < 3; q++) tr->T[q] = st->T[q];
      }
      tr->ell_next = st->ell;
      tr->num_neighbors_next = cap_next;
    }
    if (!finished) {
      st->iter = k + 1;
      if (st->iter >= st->max_iter) {
        st->stop_reason = CVO_B200_STOP_MAX_ITER;
        finished = true;
      }
    }
    st->work_counter = 0u;
    st->item_counter = 0u;
    if (finished) st->done = 1;
  }
}

// compute_step_size_xi + compute_step_size_poly_coeff (CvoGPU.cu:953-1082) over the ELL rows of
// this block; returns the warp's B, C, D, E sums (all lanes).  Row data written by other blocks
// (exact redo) is read past L1.
template <bool kFly>
__device__ __forceinline__ void step_rows(const IterArgs& A, const DevState* hs, double& wB,
                                          double& wC, double& wD, double& wE) {
  const float* s_pose = hs->Rinv;  // Rinv[9], Tinv[3] are contiguous
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  const int g = lane >> 3, gl = lane & 7;
  const float ell = hs->ell;
  const int use_range_ell = hs->kc.use_range_ell;

  // compute_step_size_xi prologue (CvoGPU.cu:970-980): precomputed by the flow finaliser
  const float* omega = hs->omega;
  const float* v = hs->v;
  const float* W2 = hs->W2;
  const float* W3 = hs->W3;
  const float* W4 = hs->W4;
  const float* Wv = hs->Wv;
  const float* W2v = hs->W2v;
  const float* W3v = hs->W3v;

  wB = wC = wD = wE = 0.0;
  // eight lanes per source row, four rows per warp (rows hold ~10 entries in tracking regimes)
  const int slot0 = (A.row_spread ? warp_in_block * (int)gridDim.x + (int)blockIdx.x
                                  : (int)blockIdx.x * warps_per_block + warp_in_block) * kRowsPerWarp + g;  // flow_rows' map
  const int slot_stride = gridDim.x * warps_per_block * kRowsPerWarp;
  for (int row = slot0; row < A.n_rows; row += slot_stride) {
    const int n = (int)__ldcg(A.row_nnz + row);
    if (n == 0) continue;
    const int ig = A.row_begin + row;
    const float4 pa = A.src_xyz[ig];
    const float px[3] = {pa.x, pa.y, pa.z};
    float temp_ell = ell;
    if (use_range_ell) {
      const float d2_sqrt = sqrtf(dot3f(px, px));
      temp_ell = range_ell(ell, d2_sqrt);
    }
    const float temp_coef = 1 / (2.0 * temp_ell * temp_ell);
    const uint32_t* idx = A.ell_idx + (size_t)row * A.cap_max;
    const float* val = A.ell_val + (size_t)row * A.cap_max;
    for (int e = gl; e < n; e += kGroup) {
      const int j = (int)__ldcg(idx + e);
      const float A_ij = __ldcg(val + e);
      const float4 yb = kFly ? move_target(A, s_pose, s_pose + 9, A.tv[0].xyz[j]) : A.tgt_moved[j];
      const float y[3] = {yb.x, yb.y, yb.z};
      // compute_step_size_xi, CvoGPU.cu:974-983
      float z1[3], z2[3], z3[3], z4[3], t[3];
      cross3f(omega, y, t);
#pragma unroll
      for (int k = 0; k < 3; k++) z1[k] = t[k] + v[k];
      mat3f_vec(W2, y, t);
#pragma unroll
      for (int k = 0; k < 3; k++) z2[k] = t[k] + Wv[k];
      mat3f_vec(W3, y, t);
#pragma unroll
      for (int k = 0; k < 3; k++) z3[k] = t[k] + W2v[k];
      mat3f_vec(W4, y, t);
#pragma unroll
      for (int k = 0; k < 3; k++) z4[k] = t[k] + W3v[k];
      const float normxiz2 = dot3f(z1, z1);
      const float xiz_dot_xi2z = (-dot3f(z1, z2));
      const float epsil_const = dot3f(z2, z2) + 2 * dot3f(z1, z3);
      // compute_step_size_poly_coeff, CvoGPU.cu:1056-1078
      const float diff_xy[3] = {px[0] - y[0], px[1] - y[1], px[2] - y[2]};
      const float two_z2[3] = {2.0f * z2[0], 2.0f * z2[1], 2.0f * z2[2]};
      const float neg_z3[3] = {-z3[0], -z3[1], -z3[2]};
      const float two_z4[3] = {2.0f * z4[0], 2.0f * z4[1], 2.0f * z4[2]};
      const float beta_ij = (-2.0 * temp_coef * dot3f(z1, diff_xy));
      const float gamma_ij = (-temp_coef * (normxiz2 + dot3f(two_z2, diff_xy)));
      const float delta_ij = (2.0 * temp_coef * (xiz_dot_xi2z + dot3f(neg_z3, diff_xy)));
      const float epsil_ij = (-temp_coef * (epsil_const + dot3f(two_z4, diff_xy)));
      const double bi = double(A_ij * beta_ij);
      wB += bi;
      const double ci = double(A_ij * (gamma_ij + beta_ij * beta_ij / 2.0));
      wC += ci;
      const double di =
          double(A_ij * (delta_ij + beta_ij * gamma_ij + beta_ij * beta_ij * beta_ij / 6.0));
      wD += di;
      const double ei = double(A_ij * (epsil_ij + beta_ij * delta_ij +
                                       1 / 2.0 * beta_ij * beta_ij * gamma_ij +
                                       1 / 2.0 * gamma_ij * gamma_ij +
                                       1 / 24.0 * beta_ij * beta_ij * beta_ij * beta_ij));
      wE += ei;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    wB += __shfl_xor_sync(0xffffffffu, wB, o);
    wC += __shfl_xor_sync(0xffffffffu, wC, o);
    wD += __shfl_xor_sync(0xffffffffu, wD, o);
    wE += __shfl_xor_sync(0xffffffffu, wE, o);
  }
}

template <bool kFly>
__global__ void __launch_bounds__(kSparseThreads, 3) step_kernel_t(IterArgs A) {
  DevState* st = A.st;
  __shared__ double sh[kSparseThreads * 4];
  __shared__ uint32_t s_hot[kHot1Words + kHot2Words];  // pose + constants, flow result
  __shared__ bool is_last;
  for (int i = threadIdx.x; i < kHot1Words + kHot2Words; i += blockDim.x)
    s_hot[i] = __ldcg(reinterpret_cast<const uint32_t*>(st) + i);
  __syncthreads();
  const DevState* hs = reinterpret_cast<const DevState*>(s_hot);
  if (hs->done) return;
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  if (blockIdx.x == 0 && threadIdx.x == 0) st->dbg[4] = gtime();
  double wB, wC, wD, wE;
  step_rows<kFly>(A, hs, wB, wC, wD, wE);
  if (lane == 0) {
    sh[warp_in_block * 4 + 0] = wB;
    sh[warp_in_block * 4 + 1] = wC;
    sh[warp_in_block * 4 + 2] = wD;
    sh[warp_in_block * 4 + 3] = wE;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double r = sh[threadIdx.x];
    for (int w = 1; w < warps_per_block; w++) r += sh[w * 4 + threadIdx.x];
    reinterpret_cast<double*>(A.step_part)[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = r;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(&st->step_blocks_done, 1u);
    is_last = (prev == gridDim.x - 1);
    if (is_last) __threadfence();
  }
  __syncthreads();
  if (!is_last) return;
  if (threadIdx.x == 0) st->dbg[5] = gtime();
  double tot[4];
  block_reduce_partials<4, 4>(reinterpret_cast<const double*>(A.step_part), (int)gridDim.x, tot, sh);
  if (threadIdx.x == 0) {
    st->dbg[6] = gtime();
    st->step_blocks_done = 0u;
    if (A.world > 1) {
      for (int k = 0; k < 4; k++) st->local_step[k] = tot[k];
    }
  }
  if (A.world <= 1) {  // block-uniform: the controller runs on two warps of this (last) block
    __shared__ CtrlScratch s_ctrl;
    __syncthreads();
    controller_step(A, st, tot, &s_ctrl);
    if (threadIdx.x == 0) st->dbg[7] = gtime();
  }
}

// ================================================================== persistent align kernel
// One cooperative launch runs the whole registration loop in cell-query mode: every block keeps
// its own copy of the controller state in shared memory, the two cross-row reductions of an
// iteration are grid barriers after which EVERY block reduces the block partials in the same
// fixed order and runs the (deterministic) controller itself, so all copies stay bit-identical
// and nothing is broadcast.  Compared with one launch per phase this removes, per iteration,
// two kernel boundaries, two last-block hand-offs and the reloads of the state, and keeps the
// static clouds in L1 across iterations.  Block 0 records the trace and writes the state back.
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// all blocks are co-resident (cooperative launch); counter is monotone: barrier #e completes when
// it reaches e * gridDim.x
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& epoch) {
  __syncthreads();
  epoch += 1u;
  if (threadIdx.x == 0) {
    __threadfence();  // release this block's writes (cumulative over the barrier above)
    atomicAdd(counter, 1u);
    const unsigned int target = epoch * gridDim.x;
    while (ld_acquire_u32(counter) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}
// ---- fused multi-GPU all-gather of one small record (see XMailbox).  tot[] is this rank's total
// (valid on thread 0 of every block; identical in all blocks); on return it holds the sum (first
// NSUM values) / max (rest) over all ranks, reduced in RANK ORDER on every rank and in every
// block, so all copies of the controller state stay bit-identical across the whole job.
// Returns false on thread 0 if a peer's record did not arrive within ~10 s (a dead peer must not
// hang the GPU).
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// one double as two tagged words / back (polls until both halves carry `tag`; false on timeout
// or on the poison tag)
__device__ __forceinline__ void ll_store(unsigned long long* w, double x, unsigned int tag) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  st_volatile_u64(w, ((unsigned long long)tag << 32) | (b & 0xffffffffull));
  st_volatile_u64(w + 1, ((unsigned long long)tag << 32) | (b >> 32));
}
__device__ __forceinline__ bool ll_load(const unsigned long long* w, unsigned int tag, long long t0,
                                        long long budget, double& x) {
  unsigned long long lo, hi;
  while (true) {
    lo = ld_volatile_u64(w);
    hi = ld_volatile_u64(w + 1);
    if ((unsigned int)(lo >> 32) == tag && (unsigned int)(hi >> 32) == tag) break;
    if ((unsigned int)(lo >> 32) == 0xffffffffu || (budget > 0 && clock64() - t0 > budget)) return false;
  }
  x = __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull)));
  return true;
}
template <int NV, int NSUM>
__device__ bool xgpu_allgather(const IterArgs& A, DevState* gst, double (&tot)[NV], int phase,
                               unsigned long long epoch) {
  bool ok = true;
  if (threadIdx.x == 0) {
    const int par = (int)(epoch & 1ull);
    // 32-bit tag: launch generation and iteration; never 0 (fresh mailbox) or ~0 (poison)
    const unsigned int tag = ((((unsigned int)(epoch >> 32) & 0x7fffu) + 1u) << 16) | ((unsigned int)epoch & 0xffffu);
    if (blockIdx.x == 0) {
      // ---- publish: 2 NV independent 8-byte stores per peer over NVLink (own mailbox included)
      for (int r = 0; r < A.xworld; r++) {
        unsigned long long* dst = A.xpeer[r]->ll[phase][par][A.xrank];
#pragma unroll
        for (int k = 0; k < NV; k++) ll_store(dst + 2 * k, tot[k], tag);
      }
      // ---- gather in rank order from the own mailbox
      XMailbox* me = A.xpeer[A.xrank];
      double acc[NV];
#pragma unroll
      for (int k = 0; k < NV; k++) acc[k] = 0.0;
      const long long t0 = clock64();
      for (int r = 0; r < A.xworld && ok; r++) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
          double x = 0.0;
          ok = ok && ll_load(me->ll[phase][par][r] + 2 * k, tag, t0, 20000000000ll /* ~10 s */, x);
          acc[k] = (k < NSUM) ? acc[k] + x : fmax(acc[k], x);
        }
      }
      // ---- hand the job totals (or the failure) to the other blocks of this GPU
#pragma unroll
      for (int k = 0; k < NV; k++) {
        tot[k] = acc[k];
        ll_store(gst->xll[phase][par] + 2 * k, acc[k], ok ? tag : 0xffffffffu);
      }
    } else {
#pragma unroll
      for (int k = 0; k < NV; k++) {
        double x = 0.0;
        ok = ok && ll_load(gst->xll[phase][par] + 2 * k, tag, 0, 0, x);
        tot[k] = x;
      }
    }
  }
  return ok;
}

// Barrier-free all-reduce over the cooperative grid (see LLBoard).  v: the warp's values (valid on
// lane 0 of every warp).  On return out[] (valid on EVERY thread) holds the sum (k < NSUM) / max
// over all blocks, reduced in block order by every block, so all blocks get bit-identical totals.
// release: this block wrote data other blocks will read after the reduction (queued rows, redone
// ELL rows) - the publishing lanes fence before their stores; acquire: readers fence after the
// poll.  sh_w: >= 32 * NV doubles; sh_all: >= NV * kLLMaxBlocks doubles.
template <int NV, int NSUM>
__device__ __forceinline__ void ll_allreduce(LLBoard* board, unsigned int seq, const double (&v)[NV],
                                             double* sh_w, double* sh_all, double (&out)[NV], bool release,
                                             bool acquire) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int par = (int)(seq & 1u);
  __syncthreads();  // sh_w / sh_all may still be read by the previous phase; orders the block's writes
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) sh_w[w * NV + k] = v[k];
  }
  __syncthreads();
  // one warp per value: the block's warps in a fixed xor tree, lane 0 publishes
  for (int k = w; k < NV; k += nw) {
    double r = (lane < nw) ? sh_w[lane * NV + k] : 0.0;  // 0 is neutral for the sums and the max (>= 0)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double x = __shfl_xor_sync(0xffffffffu, r, o);
      r = (k < NSUM) ? r + x : fmax(r, x);
    }
    if (lane == 0) {
      if (release) __threadfence();
      ll_store(&board->w[par][blockIdx.x][2 * k], r, seq);
    }
  }
  // every thread fetches its share of the nblocks x NV values; a word is either old or complete.
  // (Issuing a thread's loads as one batch of 16-byte loads was measured slower: +1.2 us per
  // reduction on C2 - the values mostly arrive while the first ones are being polled.)
  const int total = (int)gridDim.x * NV;
  for (int id = threadIdx.x; id < total; id += blockDim.x) {
    const int b = id / NV, k = id - b * NV;
    double x = 0.0;
    (void)ll_load(&board->w[par][b][2 * k], seq, 0, 0, x);
    sh_all[k * kLLMaxBlocks + b] = x;
  }
  if (acquire) __threadfence();
  __syncthreads();
  for (int k = w; k < NV; k += nw) {
    double acc = 0.0;
    for (int b = lane; b < (int)gridDim.x; b += 32) {
      const double x = sh_all[k * kLLMaxBlocks + b];
      acc = (k < NSUM) ? acc + x : fmax(acc, x);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double x = __shfl_xor_sync(0xffffffffu, acc, o);
      acc = (k < NSUM) ? acc + x : fmax(acc, x);
    }
    if (lane == 0) sh_w[k] = acc;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; k++) out[k] = sh_w[k];
}

// block partial (NV values per warp in v, valid on lane 0) -> part[k * gridDim.x + blockIdx.x]
template <int NV, int NSUM>
__device__ __forceinline__ void publish_block_partial(const double (&v)[NV], double* sh, double* part) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();  // sh may still be read by the previous phase
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) sh[w * NV + k] = v[k];
  }
  __syncthreads();
  // one warp per value: lanes = the block's warps (<= 32), fixed xor tree
  for (int k = w; k < NV; k += nw) {
    double r = (lane < nw) ? sh[lane * NV + k] : 0.0;  // 0 is neutral for the sums and the max (>= 0)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double x = __shfl_xor_sync(0xffffffffu, r, o);
      r = (k < NSUM) ? r + x : fmax(r, x);
    }
    if (lane == 0) part[(size_t)k * gridDim.x + blockIdx.x] = r;
  }
}

// dynamic shared memory of the persistent kernel: one buffer, three lives - the candidate queues
// of the flow phase (per 8-lane row group: cell queries 48 words, tile cells 80), all blocks'
// values of a reduction, and (tile mode) the per-warp staging of the tile phase.  Every reduction
// starts and ends with a block barrier; the phases run between two of them.
constexpr size_t align_grid_smem_bytes(int threads, int gen) {
  const size_t list = sizeof(uint32_t) * (size_t)(threads / kGroup) * (size_t)(gen == 2 ? kGroupList : kGridList);
  const size_t all = sizeof(double) * kLLValues * kLLMaxBlocks;
  const size_t tile = gen == 2 ? sizeof(TileSmemWarp) * (size_t)(threads / 32) : 0;
  return (list > all ? (list > tile ? list : tile) : (all > tile ? all : tile)) + 16;
}

// Candidate-cell reuse of the persistent tile mode (a Verlet list with a skin).  The cells of
// tile_phase hold, per source row x, every target y with |y - q_b| <= rq_b(x) + skin(x), q_b = R_b x
// + T_b the row in the target's frame at BUILD time and skin(x) = s_tr + s_rot |x| (+ a relative
// 4e-6 of the radius).  An iteration at pose (R, T) needs every y with |y - q| <= rq(x); since
// |q - q_b| <= |R - R_b|_F |x| + |T - T_b|, the old cells still contain them while
//   |R - R_b|_F |x| + |T - T_b| + (slack - slack_b)+ <= s_rot |x| + s_tr for every row (checked at
//   |x| = 0 and at max|x|: the expression is linear in |x|),   ell smax <= ell_b smax_b (1 + 2e-6)
// (the cut-off radius is proportional to ell * smax, CvoGPU.cu:506-511; flow_rows<2> re-tests every
// candidate with the reference's arithmetic at the CURRENT pose, so a superset changes nothing).
// Budgets: s_tr = kappa * (cut-off radius at range 0), s_rot = s_tr / max|x|.  Called by thread 0 of
// every block at the top of an iteration: same inputs, same decision everywhere.
__device__ void verlet_decide(const IterArgs& A, DevState* st) {
  bool rebuild = st->vl_valid == 0 || st->controller_on == 2 || !(A.verlet_kappa > 0.f);
  if (!rebuild) {
    float dr2 = 0.f, dt2 = 0.f;
    for (int i = 0; i < 9; i++) {
      const float d = st->R[i] - st->vl_R[i];
      dr2 += d * d;
    }
    for (int i = 0; i < 3; i++) {
      const float d = st->T[i] - st->vl_T[i];
      dt2 += d * d;
    }
    const float dr = sqrtf(dr2) * 1.0001f, dt = sqrtf(dt2) * 1.0001f + 1e-7f;
    const float ds = fmaxf(0.f, st->grid_slack - st->vl_slack) * 1.0001f;
    const float ls = st->ell * st->smax;
    // (NaN anywhere: every comparison is false -> rebuild)
    // every row has |x| <= xm, and (dr - s_rot) |x| + (dt + ds - s_tr) is linear in |x|: it is <= 0 for
    // all rows iff it is at |x| = 0 and at |x| = xm - rotation may use what translation left over
    const float xm = A.src_rmax * 1.0001f;
    const bool keep = dt + ds <= st->vl_str && dr * xm + dt + ds <= st->vl_srot * xm + st->vl_str &&
                      ls <= st->vl_ls * 1.000002f;
    rebuild = !keep;
  }
  if (rebuild) {
    const float r0 = sqrtf(fmaxf(0.f, (float)(-2.0 * st->ell * st->ell * st->kc.log_geo)));
    const bool on = A.verlet_kappa > 0.f && st->controller_on != 2 && r0 > 0.f && r0 < 1e30f;
    st->vl_str = on ? A.verlet_kappa * r0 : 0.f;
    st->vl_srot = on ? st->vl_str / fmaxf(A.src_rmax, 1e-3f) : 0.f;
    for (int i = 0; i < 9; i++) st->vl_R[i] = st->R[i];
    for (int i = 0; i < 3; i++) st->vl_T[i] = st->T[i];
    st->vl_ls = st->ell * st->smax;
    st->vl_slack = st->grid_slack;
    st->vl_valid = 1;
    st->tile_builds += 1u;
    // dynamic hand-out of the build's (tile, part) items: tile_phase advances the counter by exactly this many
    st->tile_item_base = st->tile_item_next;
    st->tile_item_next += (unsigned int)(((A.n_rows + kTileRows - 1) / kTileRows) * A.tile_parts);
  }
  st->tile_rebuild = rebuild ? 1 : 0;
}

// kGen: 1 = cell queries, 2 = tile cells (a tile phase + a grid-wide hand-over in front of the flow
// phase)
template <int kThreads, bool kFused, bool kColour, int kGen = 1>
__global__ void __launch_bounds__(kThreads, 1) align_grid_kernel(IterArgs A) {
  __shared__ double sh[32 * kLLValues];              // per-warp values of a reduction
  extern __shared__ __align__(16) unsigned char s_raw[];
  constexpr int kListWords = (kGen == 2) ? kGroupList : kGridList;
  uint32_t(*s_list)[kListWords] = reinterpret_cast<uint32_t(*)[kListWords]>(s_raw);
  double* sh_all = reinterpret_cast<double*>(s_raw);
  __shared__ __align__(16) DevState s_st;
  __shared__ CtrlScratch s_ctrl;
  __shared__ int s_team_cnt[32];  // brute_team_rows: survivors per warp
  // debug (CVO_B200_STAMPS=1): time spent per phase by thread 0 of block 0, summed over the loop
  __shared__ unsigned long long s_acc[12];
  unsigned long long t_prev = 0ull;
  const bool stamping = A.stamps != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  if (stamping) {
    for (int q = 0; q < 12; q++) s_acc[q] = 0ull;
    t_prev = gtime();
  }
#define CVO_PHASE(q)                    \
  if (stamping) {                       \
    const unsigned long long t = gtime(); \
    s_acc[q] += t - t_prev;             \
    t_prev = t;                         \
  }
  DevState* gst = A.st;
  for (int i = threadIdx.x; i < (int)(sizeof(DevState) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(&s_st)[i] = __ldcg(reinterpret_cast<const uint32_t*>(gst) + i);
  __syncthreads();
  if (threadIdx.x == 0) {
    if (blockIdx.x != 0) s_st.trace = nullptr;  // block 0 records the trace
    s_st.sat_base = 0u;                          // the host zeroes the monotone counters per launch
    s_st.work_base = 0u;
    s_st.tile_item_base = s_st.tile_item_next = 0u;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  uint32_t tile_bar_phase = 0u;
  __shared__ uint64_t s_tile_bar[kGen == 2 ? kThreads / 32 : 1];  // outside the overlaid buffer
  if (kGen == 2) {
    if (lane == 0) {
      mbar_init(&s_tile_bar[warp_in_block], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  unsigned int seq = 0;  // sequence number of the grid-wide reductions of this launch
  // dynamic row hand-out (more rows than resident row slots): every warp ends its flow phase
  // with exactly one failing fetch, so the monotone counter advances by a known amount
  const int total_slots = (int)gridDim.x * warps_per_block * kRowsPerWarp;
  const unsigned int work_per_iter =
      A.n_rows > total_slots
          ? (unsigned)kRowsPerWarp * (unsigned)((A.n_rows - total_slots + kRowsPerWarp - 1) / kRowsPerWarp +
                                                (int)gridDim.x * warps_per_block)
          : 0u;

  while (!s_st.done) {  // the same value in every block
    // ---- flow phase (fill_in_A_mat_gpu + compute_flow_gpu_no_eigen on this block's rows)
    double tot[9];
    unsigned int n_sat = 0u;
    if (kGen == 2) {
      // ---- tile phase: candidate cells for ALL rows of this rank, items dealt round-robin to the
      //      warps of the grid; the rows are evaluated by other blocks, so the cells are handed
      //      over through one grid-wide reduction with release / acquire.  Skipped while the cells
      //      of an earlier iteration still cover the current pose (verlet_decide).
      if (threadIdx.x == 0) verlet_decide(A, &s_st);
      __syncthreads();
      if (s_st.tile_rebuild) {  // the same value in every block
        TileSmemWarp& S = reinterpret_cast<TileSmemWarp*>(s_raw)[warp_in_block];
        tile_phase(A, &s_st, S, &s_tile_bar[kGen == 2 ? warp_in_block : 0], warp_in_block * (int)gridDim.x + (int)blockIdx.x,
                   (int)gridDim.x * warps_per_block, tile_bar_phase, &gst->item_counter, s_st.tile_item_base);
        CVO_PHASE(2)
        double one[1] = {1.0}, got[1];
        ll_allreduce<1, 1>(A.ll, ++seq, one, sh, sh_all, got, true, true);
        CVO_PHASE(7)
      }
    }
    // (A.brute, cell-query instantiation only: the generator is skipped and EVERY row goes through
    //  the exact walk below, one warp per row - a few hundred rows against a small target)
    const bool brute = (kGen == 1) && A.brute != 0;
    if (!brute) {
      double bp[9], queued[2], v[kLLValues], r[kLLValues];
      flow_rows<kGen, kColour>(A, gst, &s_st, s_list[threadIdx.x >> 3], bp, queued);
      CVO_PHASE(0)
#pragma unroll
      for (int k = 0; k < 8; k++) v[k] = bp[k];
      v[8] = queued[0];
      v[9] = queued[1];
      v[10] = bp[8];
      // release only when this block queued rows for the exact redo (sat_list entries must be
      // visible to whoever redoes them); block-uniform decision
      // ... or when rows were handed out dynamically: step_rows walks the rows by the static map, so
      // it reads ELL rows other blocks wrote (release here, acquire after the poll)
      const bool dyn = work_per_iter != 0u;
      const bool rel = (__syncthreads_or(lane == 0 && queued[0] > 0.0) != 0) || dyn;
      ll_allreduce<kLLValues, 10>(A.ll, ++seq, v, sh, sh_all, r, rel, dyn);
      CVO_PHASE(1)
#pragma unroll
      for (int k = 0; k < 8; k++) tot[k] = r[k];
      tot[8] = r[10];
      n_sat = (unsigned int)(r[8] + 0.5);
    } else {
#pragma unroll
      for (int k = 0; k < 9; k++) tot[k] = 0.0;
    }
    if (brute || n_sat > 0u) {
      // exact redo of the cut rows (brute: of every row), one warp per row over the whole grid, behind
      // one more reduction
      if (!brute) {
        __threadfence();  // acquire: the queued row indices of the other blocks
        __syncthreads();
      }
      double f[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, v[kLLValues], r[kLLValues];
      if (brute && A.brute == 2) {  // block-uniform: teams of warps per row (block barriers inside)
        brute_team_rows(A, &s_st, s_raw, s_team_cnt, f);
      } else {
        const unsigned int n_exact = brute ? (unsigned int)A.n_rows : n_sat;
        for (unsigned int si = (brute || A.row_spread) ? warp_in_block * gridDim.x + blockIdx.x : blockIdx.x * warps_per_block + warp_in_block;
             si < n_exact; si += gridDim.x * warps_per_block)
          redo_row<true>(A, s_st.kc, s_st.Rinv, s_st.Tinv, s_st.ell, s_st.num_neighbors,
                         brute ? (int)si : (int)__ldcg(&A.sat_list[si]), lane, f);
      }
#pragma unroll
      for (int k = 0; k < 8; k++) v[k] = f[k];
      v[8] = v[9] = 0.0;
      v[10] = f[8];
      // redone ELL rows are read by other blocks (step_rows' static row map): release + acquire
      ll_allreduce<kLLValues, 10>(A.ll, ++seq, v, sh, sh_all, r, true, true);
#pragma unroll
      for (int k = 0; k < 8; k++) tot[k] += r[k];
      tot[8] = fmax(tot[8], r[10]);
    }
    CVO_PHASE(3)
    // ---- multi-GPU: this rank's totals -> the job's totals (NVLink stores + local spin)
    const unsigned long long xepoch = (A.xgen << 32) | (unsigned long long)(unsigned)(s_st.iter + 1);
    bool xok = true;
    if (kFused) xok = xgpu_allgather<9, 8>(A, gst, tot, 0, xepoch);
    // ---- normalisation, omega_hat powers
    if (threadIdx.x == 0) {
      s_st.n_sat = n_sat;  // update_tf_device's bookkeeping
      s_st.sat_base += n_sat;
      s_st.work_base += work_per_iter;
      finalize_flow_scalar(&s_st, tot);
      if (!xok) {  // a peer is gone: stop this rank's loop with an error instead of spinning
        s_st.ret = CVO_B200_ERR_NCCL;
        s_st.stop_reason = CVO_B200_STOP_NONE;
        s_st.done = 1;
      }
    }
    __syncthreads();
    CVO_PHASE(4)
    // ---- step phase (compute_step_size_xi + _poly_coeff on this block's ELL rows)
    {
      double w4[4], r4[4];
      step_rows<true>(A, &s_st, w4[0], w4[1], w4[2], w4[3]);
      CVO_PHASE(5)
      ll_allreduce<4, 4>(A.ll, ++seq, w4, sh, sh_all, r4, false, false);
      CVO_PHASE(6)
      bool xok2 = true;
      if (kFused) xok2 = xgpu_allgather<4, 4>(A, gst, r4, 1, xepoch);
      if (stamping) s_st.dbg[12] = gtime();
      controller_step(A, &s_st, r4, &s_ctrl);
      if (stamping) s_st.dbg[11] = gtime();
      if (threadIdx.x == 0 && !xok2) {
        s_st.ret = CVO_B200_ERR_NCCL;
        s_st.done = 1;
      }
    }
    __syncthreads();
    CVO_PHASE(9)
  }
  if (stamping)
    for (int q = 0; q < 10; q++) A.stamps[q] = s_acc[q];
#undef CVO_PHASE
  // ---- block 0 hands the final state (pose, flags, results) back.  The scheduling scratch
  //      (work counters, exchange flags) is NOT written back: other blocks may still use it.
  if (blockIdx.x == 0) {
    const int w0 = (int)(offsetof(DevState, work_counter) / 4), w1 = (int)(offsetof(DevState, trace) / 4),
              w2 = (int)(offsetof(DevState, xll) / 4);
    if (threadIdx.x == 0) s_st.sat_base = s_st.work_base = 0u;  // meaningful inside a launch only
    __syncthreads();
    for (int i = threadIdx.x; i < w2; i += blockDim.x)
      if (i < w0 || i >= w1)
        reinterpret_cast<uint32_t*>(gst)[i] = reinterpret_cast<const uint32_t*>(&s_st)[i];
  }
}

// ================================================================== multi-GPU finalisers
// After the all-gather of every rank's local totals (gathered[r*stride ..]); each rank
// reduces in rank order, so all ranks compute bit-identical omega, v, step and pose.
__global__ void finalize_flow_kernel(IterArgs A, const double* gathered, int stride) {
  DevState* st = A.st;
  if (st->done) return;
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double tot[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int r = 0; r < A.world; r++) {
    const double* g = gathered + (size_t)r * stride;
    for (int k = 0; k < 8; k++) tot[k] += g[k];
    tot[8] = fmax(tot[8], g[8]);
  }
  finalize_flow_scalar(st, tot);
}
__global__ void finalize_step_kernel(IterArgs A, const double* gathered, int stride) {
  DevState* st = A.st;
  __shared__ CtrlScratch s_ctrl;
  if (st->done) return;  // uniform: one block
  double tot[4] = {0, 0, 0, 0};
  if (threadIdx.x == 0) {
    for (int r = 0; r < A.world; r++) {
      const double* g = gathered + (size_t)r * stride;
      for (int k = 0; k < 4; k++) tot[k] += g[k];
    }
  }
  controller_step(A, st, tot, &s_ctrl);
}
// host-initialised state needs the same Rinv/Tinv/bound update_tf_device computes
// The log-dependent constants of fill_in_A_mat_gpu's prologue (CvoGPU.cu:509-515, :246-254 for the
// dense-kernel variant) with the DEVICE overloads the reference's threads call: log(float) is
// logf, whose device implementation is not glibc's (they may differ in the last bit), the products
// are double.  The host (make_consts) fills everything else.
__device__ void device_consts(KernConsts& k) {
  k.log_geo = logf(k.sp_thres / k.sigma2);
  k.d2_c_thres = k.d2_s_thres = k.d2_s_thres_dense = 1.f;
  if (k.use_intensity) k.d2_c_thres = -2.0 * k.c2 * logf(k.sp_thres / k.c_sigma2);
  if (k.use_semantics) {
    k.d2_s_thres = -2.0 * k.s_ell * k.s_ell * logf(k.sp_thres / k.s_sigma2);
    k.d2_s_thres_dense = -2.0 * k.s_ell_square * logf(k.sp_thres / k.s_sigma2);
  }
}
__global__ void init_bound_kernel(IterArgs A) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  device_consts(A.st->kc);
  update_tf_device(A, A.st);
}

// ================================================================== fp32 pipe microbenchmark
// Independent FMA chains; kind 0 = scalar FFMA, 1 = packed FFMA2.  Reports lane-FMAs.
__global__ void __launch_bounds__(256) fma_peak_kernel(int kind, int iters, float* sink) {
  float a = 1.0000001f, b = 1e-9f * (float)threadIdx.x;
  if (kind == 0) {
    float x[16];
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = (float)k;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int k = 0; k < 16; k++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[k]) : "f"(a), "f"(b));
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; k++) s += x[k];
    if (s == 123.456f) sink[0] = s;
  } else {
    unsigned long long x[16];
    const unsigned long long a2 = pack2(a, a), b2 = pack2(b, b);
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = pack2((float)k, (float)k + 0.5f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int k = 0; k < 16; k++)
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[k]) : "l"(a2), "l"(b2));
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; k++) {
      float lo, hi;
      unpack2(x[k], lo, hi);
      s += lo + hi;
    }
    if (s == 123.456f) sink[0] = s;
  }
}

// ================================================================== host-side launchers
void launch_prep(const IterArgs& A, int blocks, cudaStream_t s) {
  prep_kernel<<<blocks, 256, 0, s>>>(A);
}
void launch_pair(const IterArgs& A, int blocks, cudaStream_t s) {
  pair_kernel<<<blocks, kPairWarps * 32, 0, s>>>(A);
}
void launch_flow(const IterArgs& A, int blocks, cudaStream_t s) {
  if (A.grid)
    flow_kernel_t<1><<<blocks, kSparseThreads, 0, s>>>(A);
  else if (A.tile)
    flow_kernel_t<2><<<blocks, kSparseThreads, 0, s>>>(A);
  else
    flow_kernel_t<0><<<blocks, kSparseThreads, 0, s>>>(A);
}
void launch_step(const IterArgs& A, int blocks, cudaStream_t s) {
  if (A.grid || A.tile)
    step_kernel_t<true><<<blocks, kSparseThreads, 0, s>>>(A);
  else
    step_kernel_t<false><<<blocks, kSparseThreads, 0, s>>>(A);
}
void launch_tile(const IterArgs& A, int blocks, cudaStream_t s) {
  tile_kernel<<<blocks, kPairWarps * 32, 0, s>>>(A);
}
int tile_kernel_max_blocks_per_sm() {
  int n = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, tile_kernel, kPairWarps * 32, 0);
  return n;
}
void launch_finalize_flow(const IterArgs& A, const double* gathered, int stride, cudaStream_t s) {
  finalize_flow_kernel<<<1, 32, 0, s>>>(A, gathered, stride);
}
void launch_finalize_step(const IterArgs& A, const double* gathered, int stride, cudaStream_t s) {
  finalize_step_kernel<<<1, 64, 0, s>>>(A, gathered, stride);
}
void launch_init_bound(const IterArgs& A, cudaStream_t s) { init_bound_kernel<<<1, 32, 0, s>>>(A); }
// Pose-graph edges: the source rows of an edge = frame 1 (Morton order of its own frame) moved by
// its pose (transform_point_pose_vec), with the records the generators read: xyz + colour summary,
// and rowA.w = the reference's a_to_sensor of the MOVED point (CvoGPU.cu:506, contracted as its GPU
// build does).  The prefilter centre of pair_kernel is not used on this path.
struct PoseArg {
  float m[12];
};
__global__ void pose_source_kernel(const float4* __restrict__ xyz, int n, PoseArg P, float4* __restrict__ out_xyz,
                                   float4* __restrict__ out_rowA) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 x = move_point_pose(P.m, xyz[i]);
  out_xyz[i] = x;
  const float dist = sqrtf(__fmaf_rn(x.z, x.z, __fmaf_rn(x.x, x.x, x.y * x.y)));
  out_rowA[i] = make_float4(-2.f * x.x, -2.f * x.y, -2.f * x.z, dist);
}
void launch_pose_source(const float4* xyz, int n, const float pose1[12], float4* out_xyz, float4* out_rowA,
                        cudaStream_t s) {
  PoseArg P;
  for (int k = 0; k < 12; k++) P.m[k] = pose1[k];
  if (n > 0) pose_source_kernel<<<(n + 255) / 256, 256, 0, s>>>(xyz, n, P, out_xyz, out_rowA);
}
void launch_fma_peak(int kind, int iters, int blocks, float* sink, cudaStream_t s) {
  fma_peak_kernel<<<blocks, 256, 0, s>>>(kind, iters, sink);
}
// the instantiations of the persistent kernel: block size x single GPU / fused multi-GPU
// exchange x colour cut in stage 1 (each kept out of the code that does not need it: the kernel is
// register bound) for the cell queries; two block sizes for the tile cells
template <int kThreads>
static const void* align_grid_fn_t(bool fused, bool colour) {
  if (fused)
    return colour ? (const void*)align_grid_kernel<kThreads, true, true>
                  : (const void*)align_grid_kernel<kThreads, true, false>;
  return colour ? (const void*)align_grid_kernel<kThreads, false, true>
                : (const void*)align_grid_kernel<kThreads, false, false>;
}
static const void* align_grid_fn(int threads, bool fused, bool colour, bool tile) {
  if (tile) {
    if (threads == kPersistThreadsWide)
      return fused ? (const void*)align_grid_kernel<kPersistThreadsWide, true, true, 2>
                   : (const void*)align_grid_kernel<kPersistThreadsWide, false, true, 2>;
    return fused ? (const void*)align_grid_kernel<kPersistThreads, true, true, 2>
                 : (const void*)align_grid_kernel<kPersistThreads, false, true, 2>;
  }
  if (threads == kPersistThreadsSmall) return align_grid_fn_t<kPersistThreadsSmall>(fused, colour);
  if (threads == kPersistThreadsWide) return align_grid_fn_t<kPersistThreadsWide>(fused, colour);
  return align_grid_fn_t<kPersistThreads>(fused, colour);
}
static int align_grid_threads(int threads, bool tile) {
  if (tile) return threads == kPersistThreadsWide ? kPersistThreadsWide : kPersistThreads;
  return (threads == kPersistThreadsWide || threads == kPersistThreadsSmall) ? threads : kPersistThreads;
}
cudaError_t launch_align_grid(const IterArgs& A, int blocks, int threads, cudaStream_t s) {
  IterArgs a = A;
  void* args[] = {&a};
  const bool tile = A.tile != 0;
  const int t = align_grid_threads(threads, tile);
  const void* fn = align_grid_fn(t, A.xfused != 0, A.colour != 0, tile);
  size_t smem = align_grid_smem_bytes(t, tile ? 2 : 1);
  if (A.brute == 2) {  // the team walk parks (target, a) per survivor: 8 bytes x brute_cap per warp
    const size_t need = (size_t)(t / 32) * (size_t)A.brute_cap * 8 + 16;
    if (need > smem) smem = need;
  }
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  return cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(t), args, smem, s);
}
int align_grid_max_blocks_per_sm(int threads, int tile) {
  int n = 0, m = 0;
  const int t = align_grid_threads(threads, tile != 0);
  const size_t smem = align_grid_smem_bytes(t, tile ? 2 : 1);
  const void* f0 = align_grid_fn(t, false, true, tile != 0);
  const void* f1 = align_grid_fn(t, true, true, tile != 0);
  cudaFuncSetAttribute(f0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(f1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, f0, t, smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&m, f1, t, smem);
  return n < m ? n : m;
}
int pair_kernel_max_blocks_per_sm() {
  int n = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, pair_kernel, kPairWarps * 32, 0);
  return n;
}
int sparse_kernel_max_blocks_per_sm() {
  int a = 0, b = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, flow_kernel_t<0>, kSparseThreads, 0);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, step_kernel_t<false>, kSparseThreads, 0);
  int c = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, flow_kernel_t<2>, kSparseThreads, 0);
  if (c < a) a = c;
  return a < b ? a : b;
}
int grid_kernel_max_blocks_per_sm() {
  int a = 0, b = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, flow_kernel_t<1>, kSparseThreads, 0);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, step_kernel_t<true>, kSparseThreads, 0);
  return a < b ? a : b;
}

}  // namespace cvo_b200
