// Microbenchmark 2: how much shared-memory operand traffic can the pair loop afford?
//   LD=2 : duplicated row record, two LDS.128 per row (32 B)            [v2 kernel]
//   LD=1 : packed row record (ax,ay,az,t), one LDS.128 + register moves  (16 B)
//   LD=0 : no shared-memory operand at all (row values live in registers) -> compute bound
//   JQ   : targets per lane (8 or 16)
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

template <int JQ, int LD, int RPV, int MINB>
__global__ void __launch_bounds__(256, MINB)
pair_loop(const float4* __restrict__ rowrec, const float4* __restrict__ rowpk, const float* __restrict__ px,
          const float* __restrict__ py, const float* __restrict__ pz, const float* __restrict__ pw,
          int n_tiles, int n_blocks_j, unsigned* __restrict__ hits) {
  __shared__ float4 s_rec[8][2 * TR];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4* rec = s_rec[warp];
  const int gw = blockIdx.x * 8 + warp, nw = gridDim.x * 8;
  unsigned found = 0;
  const int n_items = n_tiles * n_blocks_j;
  constexpr int NP = JQ / 2;
  for (int item = gw; item < n_items; item += nw) {
    const int rt = item / n_blocks_j, jb = (item - rt * n_blocks_j) * (32 * JQ);
    __syncwarp();
    if (LD == 2) for (int r = lane; r < 2 * TR; r += 32) rec[r] = rowrec[(size_t)rt * 2 * TR + r];
    else for (int r = lane; r < TR; r += 32) rec[r] = rowpk[(size_t)rt * TR + r];
    __syncwarp();
    const int jl = jb + JQ * lane;
    unsigned long long X[NP], Y[NP], Z[NP], W[NP];
#pragma unroll
    for (int h = 0; h < JQ / 4; h++) {
      const float4 xa = __ldg((const float4*)(px + jl + 4 * h)), ya = __ldg((const float4*)(py + jl + 4 * h));
      const float4 za = __ldg((const float4*)(pz + jl + 4 * h)), wa = __ldg((const float4*)(pw + jl + 4 * h));
      X[2*h] = pack2(xa.x, xa.y); X[2*h+1] = pack2(xa.z, xa.w); Y[2*h] = pack2(ya.x, ya.y); Y[2*h+1] = pack2(ya.z, ya.w);
      Z[2*h] = pack2(za.x, za.y); Z[2*h+1] = pack2(za.z, za.w); W[2*h] = pack2(wa.x, wa.y); W[2*h+1] = pack2(wa.z, wa.w);
    }
    float4 fixed = rec[lane & 1];
#pragma unroll 1
    for (int r0 = 0; r0 < TR; r0 += RPV) {
      bool f = false;
#pragma unroll
      for (int rr = 0; rr < RPV; rr++) {
        unsigned long long AX, AY, AZ; float t;
        if (LD == 2) {
          const float4 ra = rec[2 * (r0 + rr)], rb = rec[2 * (r0 + rr) + 1];
          AX = pack2(ra.x, ra.y); AY = pack2(ra.z, ra.w); AZ = pack2(rb.x, rb.y); t = rb.z;
        } else if (LD == 1) {
          const float4 ra = rec[r0 + rr];
          AX = pack2(ra.x, ra.x); AY = pack2(ra.y, ra.y); AZ = pack2(ra.z, ra.z); t = ra.w;
        } else {
          fixed.x += 1.0f;  // keep the compiler from hoisting; no memory operand
          AX = pack2(fixed.x, fixed.x); AY = pack2(fixed.y, fixed.y); AZ = pack2(fixed.z, fixed.z); t = fixed.w;
        }
        float s[JQ];
#pragma unroll
        for (int p = 0; p < NP; p++) {
          unsigned long long v = fma2(AX, X[p], W[p]);
          v = fma2(AY, Y[p], v);
          v = fma2(AZ, Z[p], v);
          unpack2(v, s[2 * p], s[2 * p + 1]);
        }
        float m = fminf(min3(s[0], s[1], s[2]), min3(s[3], s[4], s[5]));
        m = min3(m, s[6], s[7]);
        if (JQ == 16) {
          float m2 = fminf(min3(s[8], s[9], s[10]), min3(s[11], s[12], s[13]));
          m = min3(m, m2, fminf(s[14], s[15]));
        }
        f |= (m < t);
      }
      if (__any_sync(0xffffffffu, f)) found += 1;
    }
  }
  if (found) atomicAdd(hits, found);
}

template <int JQ, int LD, int RPV, int MINB>
void run(const char* name, const float4* rowrec, const float4* rowpk, const float* px, const float* py, const float* pz,
         const float* pw, int n, unsigned* hits, int sms) {
  int occ = 0;
  auto k = pair_loop<JQ, LD, RPV, MINB>;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 256, 0);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k);
  const int n_tiles = n / TR, nbj = n / (32 * JQ);
  const int blocks = sms * occ;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 20; w++) k<<<blocks, 256>>>(rowrec, rowpk, px, py, pz, pw, n_tiles, nbj, hits);
  cudaEventRecord(e0);
  const int reps = 400;
  for (int w = 0; w < reps; w++) k<<<blocks, 256>>>(rowrec, rowpk, px, py, pz, pw, n_tiles, nbj, hits);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double us = ms / reps * 1e3;
  printf("%-28s regs %3d occ %d : %7.2f us/launch  %.3e pairs/s  (%s)\n", name, fa.numRegs, occ, us,
         (double)n * n / (us * 1e-6), cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int n = 20480;  // 4.2e8 pair tests per launch: many items per warp, little tail quantisation
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  std::vector<float4> rec(2 * n), pk(n);
  std::vector<float> x(n), y(n), z(n), w(n);
  srand(1);
  for (int i = 0; i < n; i++) {
    float a = rand() / (float)RAND_MAX * 20 - 10, b = rand() / (float)RAND_MAX * 4 - 2, c = rand() / (float)RAND_MAX * 28 - 14;
    rec[2 * i] = make_float4(-2 * a, -2 * a, -2 * b, -2 * b);
    rec[2 * i + 1] = make_float4(-2 * c, -2 * c, -1e30f, -1e30f);
    pk[i] = make_float4(-2 * a, -2 * b, -2 * c, -1e30f);
    x[i] = rand() / (float)RAND_MAX * 20 - 10; y[i] = rand() / (float)RAND_MAX * 4 - 2; z[i] = rand() / (float)RAND_MAX * 28 - 14;
    w[i] = x[i] * x[i] + y[i] * y[i] + z[i] * z[i];
  }
  float4 *d_rec, *d_pk; float *dx, *dy, *dz, *dw; unsigned* hits;
  cudaMalloc(&d_rec, rec.size() * 16); cudaMalloc(&d_pk, pk.size() * 16);
  cudaMalloc(&dx, n * 4); cudaMalloc(&dy, n * 4); cudaMalloc(&dz, n * 4); cudaMalloc(&dw, n * 4);
  cudaMalloc(&hits, 4); cudaMemset(hits, 0, 4);
  cudaMemcpy(d_rec, rec.data(), rec.size() * 16, cudaMemcpyHostToDevice);
  cudaMemcpy(d_pk, pk.data(), pk.size() * 16, cudaMemcpyHostToDevice);
  cudaMemcpy(dx, x.data(), n * 4, cudaMemcpyHostToDevice); cudaMemcpy(dy, y.data(), n * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dz, z.data(), n * 4, cudaMemcpyHostToDevice); cudaMemcpy(dw, w.data(), n * 4, cudaMemcpyHostToDevice);
  const int sms = prop.multiProcessorCount;
  printf("%s, %d SMs, n=%d\n", prop.name, sms, n);
  { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); float ms = 0; cudaEventRecord(a);
    while (ms < 1500.f) { for (int k = 0; k < 5; k++) pair_loop<8, 2, 4, 3><<<sms * 3, 256>>>(d_rec, d_pk, dx, dy, dz, dw, n / TR, n / 256, hits);
      cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b); } }
#define RUN(JQ, LD, RPV, MINB) run<JQ, LD, RPV, MINB>("JQ=" #JQ " LD=" #LD " RPV=" #RPV " MINB=" #MINB, d_rec, d_pk, dx, dy, dz, dw, n, hits, sms)
  RUN(8, 2, 4, 3); RUN(8, 2, 4, 4); RUN(8, 1, 4, 3); RUN(8, 1, 4, 4); RUN(8, 0, 4, 3); RUN(8, 0, 4, 4);
  RUN(16, 2, 4, 2); RUN(16, 2, 2, 2); RUN(16, 1, 4, 2); RUN(16, 1, 2, 2); RUN(16, 1, 2, 3); RUN(16, 0, 4, 2); RUN(16, 0, 2, 2);
  return 0;
}
