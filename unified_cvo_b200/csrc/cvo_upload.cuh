// cvo_upload.cuh — interface of the device-side cloud build (cvo_upload.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace cvo_b200 {

// scalar results of the build, read back by the host with one small copy
struct CloudStats {
  float lo[3];          // origin of the key lattice (bounding box minimum of the finite points)
  float centroid[3];
  float max_dist;       // max_i |x_i| (a_to_sensor)
  unsigned int radius2_bits;  // bit pattern of max_i |x_i - centroid|^2 (float, rounded up)
  int n_finite;
  int pad;
  double extent;        // edge of the key lattice's cube [m]
  double scale;         // lattice units per metre
  unsigned long long occupied_cells;  // occupied cube cells at `dbits` bits per axis
};

struct CloudBuild {
  int n, F, C, Fp, Cp, cbits, dbits;
  // raw input as the caller laid it out, already on the device
  const float* xyz3; const float* feat_in; const float* lab_in; const float* geo_in;
  // scratch
  unsigned long long* keys_in; int* idx_in; void* sort_temp; size_t sort_temp_bytes;
  // outputs
  CloudStats* stats;
  unsigned long long* keys; int* perm; int* inv;
  float4* xyz; float4* xyz_o; float4* rowA;
  float* feat; float* feat_o; float* lab; float* lab_o;
  float2* geo; float2* geo_o;
  uint32_t* coarse;
  float4* blk_sphere; float4* tile_sphere; float* tile_maxdist;
};

// a frame pose as CvoFrame::pose_vec holds it: row-major 3x4 (CvoFrame.hpp), float on the device
struct PoseVec {
  float m[12];
};
// out = P [in 1]^T per point (n x 3 floats, packed); in and out are distinct buffers
cudaError_t pose_vec_transform_device(const float* xyz3_in, float* xyz3_out, int n, const PoseVec& P,
                                      cudaStream_t s);

size_t cloud_sort_temp_bytes(int n);
cudaError_t build_cloud_device(const CloudBuild& B, cudaStream_t s);

}  // namespace cvo_b200
