// shim_pack.hpp — reference containers -> the plain arrays of the C-ABI (shared by the shim's
// translation units).  Include AFTER the reference's headers.
#pragma once
#include <cstring>
#include <vector>

namespace cvo {
namespace shim {

// ~CvoGPU: the pose-graph edges of a CvoGPU find their device handle through the address of its
// CvoParams (IRLS_State_GPU_b200.cpp).  When the object dies that address may be reused by another
// CvoGPU: the entry is dropped so the newcomer gets a handle of its own (the old handle stays alive
// for the frames that still hold ids on it).
void forget_edge_handle(const void* params_cpu);

// CvoPointCloud (column-major Eigen matrices) -> the row-major arrays of cvo_b200_set_cloud.
// Mirrors what CvoPointCloud_to_gpu reads (CvoGPU_impl.cu:206-263).
struct Packed {
  int n = 0, F = 0, C = 0;
  std::vector<float> xyz, feat, lab, geo;
  const float* p_feat() const { return feat.empty() ? nullptr : feat.data(); }
  const float* p_lab() const { return lab.empty() ? nullptr : lab.data(); }
  const float* p_geo() const { return geo.empty() ? nullptr : geo.data(); }
};

inline Packed pack(const CvoPointCloud& pc) {
  Packed o;
  o.n = pc.num_points();
  o.xyz.resize((size_t)o.n * 3);
  const auto& pos = pc.positions();
  for (int i = 0; i < o.n; i++)
    for (int k = 0; k < 3; k++) o.xyz[3 * (size_t)i + k] = pos[i](k);
  const auto& f = pc.features();
  if (f.rows() == o.n && f.cols() > 0) {
    o.F = (int)f.cols();
    o.feat.resize((size_t)o.n * o.F);
    for (int i = 0; i < o.n; i++)
      for (int j = 0; j < o.F; j++) o.feat[(size_t)i * o.F + j] = f(i, j);
  }
  if (pc.num_classes() > 0) {
    const auto& l = pc.labels();
    o.C = pc.num_classes();
    o.lab.resize((size_t)o.n * o.C);
    for (int i = 0; i < o.n; i++)
      for (int j = 0; j < o.C; j++) o.lab[(size_t)i * o.C + j] = l(i, j);
  }
  const auto& g = pc.geometric_types();
  if ((int)g.size() >= 2 * o.n) o.geo.assign(g.begin(), g.begin() + 2 * (size_t)o.n);
  return o;
}

inline Packed pack(const pcl::PointCloud<CvoPoint>& pc) {
  Packed o;
  o.n = (int)pc.size();
  o.F = FEATURE_DIMENSIONS;
  o.C = NUM_CLASSES;
  o.xyz.resize((size_t)o.n * 3);
  o.feat.resize((size_t)o.n * o.F);
  o.lab.resize((size_t)o.n * o.C);
  o.geo.resize((size_t)o.n * 2);
  for (int i = 0; i < o.n; i++) {
    const CvoPoint& p = pc[i];
    o.xyz[3 * (size_t)i] = p.x; o.xyz[3 * (size_t)i + 1] = p.y; o.xyz[3 * (size_t)i + 2] = p.z;
    std::memcpy(&o.feat[(size_t)i * o.F], p.features, sizeof(float) * o.F);
    std::memcpy(&o.lab[(size_t)i * o.C], p.label_distribution, sizeof(float) * o.C);
    o.geo[2 * (size_t)i] = p.geometric_type[0];
    o.geo[2 * (size_t)i + 1] = p.geometric_type[1];
  }
  return o;
}

}  // namespace shim
}  // namespace cvo
