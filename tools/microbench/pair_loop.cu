// Microbenchmark of the pair-prefilter inner loop: which structure reaches the FMA pipe?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o pair_loop pair_loop.cu
// Each variant sweeps N=M=10240 (1.05e8 pair tests) with thresholds that never fire (pure fast
// path) and reports us per launch and pair tests per second.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ unsigned long long pack2(float a, float b) {
  unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(unsigned long long v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float min3(float a, float b, float c) {
  float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

constexpr int TR = 64;

// MODE 0: packed FFMA2 + min3 tree; MODE 1: scalar FFMA + min3; MODE 2: FFMA2 + FSETP/or chain
template <int RPV, int MINB, int MODE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, MINB)
pair_loop(const float4* __restrict__ rowrec, const float* __restrict__ px, const float* __restrict__ py,
          const float* __restrict__ pz, const float* __restrict__ pw, int n_tiles, int n_blocks_j,
          unsigned* __restrict__ hits) {
  __shared__ float4 s_rec[WARPS][2 * TR];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4* rec = s_rec[warp];
  const int gw = blockIdx.x * WARPS + warp, nw = gridDim.x * WARPS;
  unsigned found = 0;
  const int n_items = n_tiles * n_blocks_j;
  for (int item = gw; item < n_items; item += nw) {
    const int rt = item / n_blocks_j, jb = (item - rt * n_blocks_j) * 256;
    __syncwarp();
    for (int r = lane; r < 2 * TR; r += 32) rec[r] = rowrec[(size_t)rt * 2 * TR + r];
    __syncwarp();
    const int jl = jb + 8 * lane;
    const float4 xa = __ldg((const float4*)(px + jl)), xb = __ldg((const float4*)(px + jl + 4));
    const float4 ya = __ldg((const float4*)(py + jl)), yb = __ldg((const float4*)(py + jl + 4));
    const float4 za = __ldg((const float4*)(pz + jl)), zb = __ldg((const float4*)(pz + jl + 4));
    const float4 wa = __ldg((const float4*)(pw + jl)), wb = __ldg((const float4*)(pw + jl + 4));
    unsigned long long X[4] = {pack2(xa.x, xa.y), pack2(xa.z, xa.w), pack2(xb.x, xb.y), pack2(xb.z, xb.w)};
    unsigned long long Y[4] = {pack2(ya.x, ya.y), pack2(ya.z, ya.w), pack2(yb.x, yb.y), pack2(yb.z, yb.w)};
    unsigned long long Z[4] = {pack2(za.x, za.y), pack2(za.z, za.w), pack2(zb.x, zb.y), pack2(zb.z, zb.w)};
    unsigned long long W[4] = {pack2(wa.x, wa.y), pack2(wa.z, wa.w), pack2(wb.x, wb.y), pack2(wb.z, wb.w)};
#pragma unroll 1
    for (int r0 = 0; r0 < TR; r0 += RPV) {
      bool f = false;
#pragma unroll
      for (int rr = 0; rr < RPV; rr++) {
        const float4 ra = rec[2 * (r0 + rr)], rb = rec[2 * (r0 + rr) + 1];
        const float t = rb.z;
        float s[8];
        if (MODE == 1) {
          float xs[8], ys[8], zs[8], ws[8];
#pragma unroll
          for (int p = 0; p < 4; p++) { unpack2(X[p], xs[2*p], xs[2*p+1]); unpack2(Y[p], ys[2*p], ys[2*p+1]);
                                        unpack2(Z[p], zs[2*p], zs[2*p+1]); unpack2(W[p], ws[2*p], ws[2*p+1]); }
#pragma unroll
          for (int q = 0; q < 8; q++) s[q] = fmaf(rb.x, zs[q], fmaf(ra.z, ys[q], fmaf(ra.x, xs[q], ws[q])));
        } else {
          const unsigned long long AX = pack2(ra.x, ra.y), AY = pack2(ra.z, ra.w), AZ = pack2(rb.x, rb.y);
#pragma unroll
          for (int p = 0; p < 4; p++) {
            unsigned long long v = fma2(AX, X[p], W[p]);
            v = fma2(AY, Y[p], v);
            v = fma2(AZ, Z[p], v);
            unpack2(v, s[2 * p], s[2 * p + 1]);
          }
        }
        if (MODE == 2) {
#pragma unroll
          for (int q = 0; q < 8; q++) f |= (s[q] < t);
        } else {
          const float m = fminf(min3(s[0], s[1], s[2]), min3(s[3], s[4], s[5]));
          f |= (min3(m, s[6], s[7]) < t);
        }
      }
      if (__any_sync(0xffffffffu, f)) found += 1;
    }
  }
  if (found) atomicAdd(hits, found);
}

template <int RPV, int MINB, int MODE, int WARPS>
void run(const char* name, const float4* rowrec, const float* px, const float* py, const float* pz,
         const float* pw, int n, unsigned* hits, int sms) {
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pair_loop<RPV, MINB, MODE, WARPS>, WARPS * 32, 0);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, pair_loop<RPV, MINB, MODE, WARPS>);
  const int n_tiles = n / TR, nbj = n / 256;
  const int blocks = sms * occ;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 5; w++) pair_loop<RPV, MINB, MODE, WARPS><<<blocks, WARPS * 32>>>(rowrec, px, py, pz, pw, n_tiles, nbj, hits);
  cudaEventRecord(e0);
  const int reps = 400;
  for (int w = 0; w < reps; w++) pair_loop<RPV, MINB, MODE, WARPS><<<blocks, WARPS * 32>>>(rowrec, px, py, pz, pw, n_tiles, nbj, hits);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double us = ms / reps * 1e3;
  printf("%-34s regs %3d occ %d blocks %4d : %7.2f us/launch  %.3e pairs/s  (%s)\n", name, fa.numRegs, occ, blocks, us,
         (double)n * n / (us * 1e-6), cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int n = 10240;
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  std::vector<float4> rec(2 * n);
  std::vector<float> x(n), y(n), z(n), w(n);
  srand(1);
  for (int i = 0; i < n; i++) {
    float a = rand() / (float)RAND_MAX * 20 - 10, b = rand() / (float)RAND_MAX * 4 - 2, c = rand() / (float)RAND_MAX * 28 - 14;
    rec[2 * i] = make_float4(-2 * a, -2 * a, -2 * b, -2 * b);
    rec[2 * i + 1] = make_float4(-2 * c, -2 * c, -1e30f, -1e30f);  // never a candidate: pure fast path
    x[i] = rand() / (float)RAND_MAX * 20 - 10; y[i] = rand() / (float)RAND_MAX * 4 - 2; z[i] = rand() / (float)RAND_MAX * 28 - 14;
    w[i] = x[i] * x[i] + y[i] * y[i] + z[i] * z[i];
  }
  float4* d_rec; float *dx, *dy, *dz, *dw; unsigned* hits;
  cudaMalloc(&d_rec, rec.size() * 16); cudaMalloc(&dx, n * 4); cudaMalloc(&dy, n * 4); cudaMalloc(&dz, n * 4); cudaMalloc(&dw, n * 4);
  cudaMalloc(&hits, 4); cudaMemset(hits, 0, 4);
  cudaMemcpy(d_rec, rec.data(), rec.size() * 16, cudaMemcpyHostToDevice);
  cudaMemcpy(dx, x.data(), n * 4, cudaMemcpyHostToDevice); cudaMemcpy(dy, y.data(), n * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dz, z.data(), n * 4, cudaMemcpyHostToDevice); cudaMemcpy(dw, w.data(), n * 4, cudaMemcpyHostToDevice);
  const int sms = prop.multiProcessorCount;
  printf("%s, %d SMs\n", prop.name, sms);
  {  // ramp the clocks: ~1.5 s of work before anything is timed
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); float ms = 0; cudaEventRecord(a);
    while (ms < 1500.f) { for (int k = 0; k < 20; k++) pair_loop<4, 3, 0, 8><<<sms * 3, 256>>>(d_rec, dx, dy, dz, dw, n / TR, n / 256, hits);
      cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b); }
  }
#define RUN(RPV, MINB, MODE, WARPS) run<RPV, MINB, MODE, WARPS>("RPV=" #RPV " MINB=" #MINB " MODE=" #MODE " WARPS=" #WARPS, d_rec, dx, dy, dz, dw, n, hits, sms)
  RUN(1, 3, 0, 8); RUN(2, 3, 0, 8); RUN(4, 3, 0, 8); RUN(8, 3, 0, 8);
  RUN(4, 2, 0, 8); RUN(4, 4, 0, 8); RUN(4, 6, 0, 8); RUN(4, 8, 0, 8);
  RUN(8, 4, 0, 8); RUN(8, 6, 0, 8);
  RUN(4, 3, 1, 8); RUN(4, 4, 1, 8); RUN(4, 3, 2, 8); RUN(4, 4, 2, 8);
  RUN(4, 6, 0, 4); RUN(4, 8, 0, 4); RUN(4, 12, 0, 4); RUN(4, 16, 0, 4);
  return 0;
}
