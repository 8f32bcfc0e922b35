// cvo_export.cuh — device-side compaction of the ELL kernel matrix into CSR (cvo_export.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace cvo_b200 {

struct CsrExport {
  int n_rows;               // rows of the ELL matrix (all source rows)
  int cap_max;              // ELL row stride
  const uint32_t* row_nnz;  // per DEVICE row (Morton position of the source point)
  const uint32_t* ell_idx;
  const float* ell_val;
  const int* inv;           // caller's row -> device row
  const int* col_perm;      // null: ell_idx holds the caller's column indices in insertion order;
                            // else: ell_idx holds Morton positions of the target in any order ->
                            // columns = col_perm[idx], rows sorted ascending (the reference's order)
  int* cnt;                 // scratch, n_rows + 1
  int* row_ptr;             // out, n_rows + 1 (caller's row order)
  void* scan_temp;
  size_t scan_temp_bytes;
  const long long* base;    // null or: device pointer to the offset of this matrix' block in cols / vals
  long long* base_next;     // null or: receives *base + row_ptr[n_rows] (the next matrix' offset)
  int32_t* cols;            // out, compacted (row_ptr[n_rows] entries)
  float* vals;
};

size_t csr_scan_temp_bytes(int n_rows);
// row_ptr[i] = number of entries in the caller's rows before i; row_ptr[n_rows] = total
cudaError_t csr_row_ptr_device(const CsrExport& E, cudaStream_t s);
// cols / vals of every row, in the row's insertion order, at row_ptr[i]
cudaError_t csr_gather_device(const CsrExport& E, cudaStream_t s);

}  // namespace cvo_b200
