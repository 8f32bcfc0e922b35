// cvo_export.cu — the kernel matrix leaves the device as CSR, not as the padded ELL array.
//
// Replaces copy_internal_SparseKernelMat_gpu_to_cpu (IRLS_State_GPU.cu:70-71) and the D2H half of
// gpu_association_to_cpu (CvoGPU_impl.cu:366-427), which copy rows x cap x 8 bytes whatever the
// fill (KITTI-sized, cap 256: 33 MB for a few thousand entries).  Here: row counts gathered into
// the caller's row order -> exclusive scan (cub) -> one warp per row copies its entries to their
// final place; only row_ptr and the nnz entries cross PCIe.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/device/device_scan.cuh>

#include "cvo_export.cuh"

namespace cvo_b200 {
namespace {
__global__ void csr_counts_kernel(CsrExport E) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > E.n_rows) return;
  E.cnt[i] = i < E.n_rows ? (int)E.row_nnz[E.inv[i]] : 0;
}

__global__ void csr_gather_kernel(CsrExport E) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < E.n_rows; i += warps) {
    const int s = E.inv[i];
    const int n = (int)E.row_nnz[s];
    const long long o = (E.base ? *E.base : 0ll) + (long long)E.row_ptr[i];
    const uint32_t* idx = E.ell_idx + (size_t)s * E.cap_max;
    const float* val = E.ell_val + (size_t)s * E.cap_max;
    if (!E.col_perm) {
      for (int k = lane; k < n; k += 32) {
        E.cols[o + k] = (int32_t)idx[k];
        E.vals[o + k] = val[k];
      }
    } else {
      // Morton-view matrix: map to the caller's column indices and write every entry at its RANK
      // (columns are unique inside a row), i.e. in ascending column order
      for (int k = lane; k < n; k += 32) {
        const int key = E.col_perm[idx[k]];
        int rank = 0;
        for (int q = 0; q < n; q++) rank += (E.col_perm[__ldg(idx + q)] < key) ? 1 : 0;
        E.cols[o + rank] = (int32_t)key;
        E.vals[o + rank] = val[k];
      }
    }
  }
}
}  // namespace

size_t csr_scan_temp_bytes(int n_rows) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int*)nullptr, (int*)nullptr, n_rows + 1);
  return bytes;
}

cudaError_t csr_row_ptr_device(const CsrExport& E, cudaStream_t s) {
  const int n1 = E.n_rows + 1;
  csr_counts_kernel<<<(n1 + 255) / 256, 256, 0, s>>>(E);
  size_t bytes = E.scan_temp_bytes;
  cudaError_t e = cub::DeviceScan::ExclusiveSum(E.scan_temp, bytes, E.cnt, E.row_ptr, n1, s);
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

namespace {
__global__ void csr_base_next_kernel(CsrExport E) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *E.base_next = (E.base ? *E.base : 0ll) + (long long)E.row_ptr[E.n_rows];
}
}  // namespace

cudaError_t csr_gather_device(const CsrExport& E, cudaStream_t s) {
  const int blocks = (E.n_rows + 7) / 8;  // 8 warps per block
  csr_gather_kernel<<<blocks < 1 ? 1 : (blocks > 148 * 8 ? 148 * 8 : blocks), 256, 0, s>>>(E);
  if (E.base_next) csr_base_next_kernel<<<1, 32, 0, s>>>(E);
  return cudaGetLastError();
}

}  // namespace cvo_b200
