// cvo_b200_batch.hpp - the one ADDITION of the binding to the reference's API surface (optional).
//
// CvoBatchIRLS::solve (src/cvo/IRLS.cpp:111-121, kept reference code) refills the edges one by one:
//     for (auto&& state : *states_) state->update_inner_product();
// Each call waits for the device.  Replacing that loop by
//     std::vector<cvo::BinaryStateGPU*> gpu;
//     for (auto&& s : *states_) if (auto* g = dynamic_cast<cvo::BinaryStateGPU*>(s.get())) gpu.push_back(g);
//     cvo::update_inner_product_batch(gpu);
// enqueues the kernels of every edge back to back and waits once (cvo_b200_edge_update_batch); every
// edge ends up with exactly the matrix, cap and iteration count its own update_inner_product() would
// have produced.  Include AFTER cvo/IRLS_State_GPU.hpp.  Returns the total number of stored entries.
#pragma once
#include <vector>

namespace cvo {
class BinaryStateGPU;
int update_inner_product_batch(const std::vector<BinaryStateGPU*>& states);
}  // namespace cvo
