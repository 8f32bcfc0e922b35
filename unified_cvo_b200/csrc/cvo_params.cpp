// cvo_params.cpp — parameter defaults and the YAML-subset reader of the C-ABI.
//
// Replaces cvo::CvoParams::CvoParams() (CvoParams.hpp:75-126) and
// read_CvoParams_yaml (CvoParams.hpp:193-303).  The reference parses with yaml-cpp
// (absent from this image); every shipped cvo_params/*.yaml is a flat list of
// "key: number  # comment" lines after a "%YAML:1.0" / "---" header, which is all
// this reader accepts.  Unknown keys are ignored (the reference only looks up the
// keys it knows).  Duplicate keys: the FIRST occurrence wins, which is what yaml-cpp's
// map lookup (linear search, first equal key) returns — e.g.
// cvo_intensity_params_img_gpu0.yaml sets nearest_neighbors_max twice (256 at :22,
// 512 at :38) and the reference therefore runs with 256; see DESIGN.md.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/cvo_b200.h"

extern "C" void cvo_b200_params_default(cvo_b200_params* p) {
  if (!p) return;
  std::memset(p, 0, sizeof(*p));
  p->ell_init_first_frame = 0.5f;
  p->ell_init = 0.5f;
  p->ell_min = 0.05f;
  p->min_ell_iter_limit = 1;
  p->ell_max = 1.2f;
  p->dl = 0;
  p->dl_step = 0.3;
  p->sigma = 0.1f;
  p->sp_thres = 0.0006f;
  p->c = 7.0f;
  p->d = 7.0f;
  p->c_ell = 0.15f;
  p->c_sigma = 0.6f;
  p->s_ell = 0.1f;
  p->s_sigma = 0.8f;
  p->MAX_ITER = 10000;
  p->min_step = 2e-5f;
  p->eps = 0.00005f;
  p->eps_2 = 0.000012f;
  // max_step and step have NO default in the reference (uninitialised memory unless the
  // yaml sets max_step); 0 here, every shipped yaml sets max_step.
  p->max_step = 0.f;
  p->step = 0.f;
  p->ell_decay_rate = 0.9f;
  p->ell_decay_rate_first_frame = 0.99f;
  p->ell_decay_start = 30;
  p->ell_decay_start_first_frame = 300;
  p->indicator_window_size = 15;
  p->indicator_stable_threshold = 0.2f;
  p->is_pcl_visualization_on = 0;
  p->is_using_least_square = 0;
  p->is_ell_adaptive = 0;
  p->is_full_ip_matrix = 0;
  p->is_using_geometry = 1;
  p->is_using_intensity = 0;
  p->is_using_semantics = 0;
  p->is_using_range_ell = 0;
  p->is_using_kdtree = 0;
  p->is_using_geometric_type = 0;
  p->is_exporting_association = 0;
  p->multiframe_using_cpu = 1;
  p->multiframe_max_iters = 200;
  p->nearest_neighbors_max = 512;
  p->multiframe_ell_init = 0.15f;
  p->multiframe_ell_min = 0.05f;
  p->multiframe_iter_per_ell = 10;
  p->multiframe_ell_decay_rate = 0.7f;
  p->multiframe_iterations_per_ell = 50;
  p->multiframe_iterations_per_solve = 8;
  p->multiframe_downsample_voxel_size = 0.5f;
  p->multiframe_expected_points = 1000;
  p->multiframe_num_neighbors = 128;
  p->multiframe_min_nonzeros = 300;
  p->multiframe_least_squares_num_threads = 24;
}

namespace {
struct Field {
  const char* name;
  char type;  // 'f' float, 'i' int, 'd' double
  size_t offset;
};
#define CVO_F(n) {#n, 'f', offsetof(cvo_b200_params, n)}
#define CVO_I(n) {#n, 'i', offsetof(cvo_b200_params, n)}
#define CVO_D(n) {#n, 'd', offsetof(cvo_b200_params, n)}
// exactly the keys read_CvoParams_yaml looks up (CvoParams.hpp:197-296)
const Field kFields[] = {
    CVO_F(ell_init_first_frame), CVO_F(ell_init), CVO_F(ell_min), CVO_I(min_ell_iter_limit),
    CVO_F(ell_max), CVO_D(dl), CVO_D(dl_step), CVO_F(sigma), CVO_F(sp_thres), CVO_F(c), CVO_F(d),
    CVO_F(c_ell), CVO_F(c_sigma), CVO_F(s_ell), CVO_F(s_sigma), CVO_I(MAX_ITER), CVO_F(eps),
    CVO_F(eps_2), CVO_F(min_step), CVO_F(max_step), CVO_F(ell_decay_rate),
    CVO_F(ell_decay_rate_first_frame), CVO_I(ell_decay_start), CVO_I(ell_decay_start_first_frame),
    CVO_I(indicator_window_size), CVO_F(indicator_stable_threshold),
    CVO_I(is_pcl_visualization_on), CVO_I(is_using_least_square), CVO_I(is_full_ip_matrix),
    CVO_I(is_using_geometry), CVO_I(is_using_intensity), CVO_I(is_using_semantics),
    CVO_I(is_using_range_ell), CVO_I(is_using_kdtree), CVO_I(is_using_geometric_type),
    CVO_I(is_exporting_association), CVO_I(nearest_neighbors_max), CVO_I(multiframe_using_cpu),
    CVO_F(multiframe_ell_init), CVO_I(multiframe_max_iters), CVO_F(multiframe_ell_min),
    CVO_F(multiframe_ell_decay_rate), CVO_I(multiframe_iterations_per_ell),
    CVO_I(multiframe_iterations_per_solve), CVO_F(multiframe_downsample_voxel_size),
    CVO_I(multiframe_expected_points), CVO_I(multiframe_num_neighbors),
    CVO_I(multiframe_min_nonzeros), CVO_I(multiframe_least_squares_num_threads),
};
std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n");
  if (a == std::string::npos) return "";
  size_t b = s.find_last_not_of(" \t\r\n");
  return s.substr(a, b - a + 1);
}
}  // namespace

extern "C" int cvo_b200_params_read_yaml(const char* path, cvo_b200_params* p) {
  if (!path || !p) return CVO_B200_ERR_INVALID;
  FILE* f = std::fopen(path, "r");
  if (!f) return CVO_B200_ERR_IO;
  char buf[1024];
  bool seen[sizeof(kFields) / sizeof(kFields[0])] = {false};
  while (std::fgets(buf, sizeof(buf), f)) {
    std::string line(buf);
    size_t hash = line.find('#');
    if (hash != std::string::npos) line = line.substr(0, hash);
    line = trim(line);
    if (line.empty() || line[0] == '%' || line.rfind("---", 0) == 0) continue;
    size_t colon = line.find(':');
    if (colon == std::string::npos) continue;
    std::string key = trim(line.substr(0, colon));
    std::string val = trim(line.substr(colon + 1));
    if (val.empty()) continue;
    for (size_t fi = 0; fi < sizeof(kFields) / sizeof(kFields[0]); fi++) {
      const Field& fd = kFields[fi];
      if (key != fd.name) continue;
      if (seen[fi]) break;  // first occurrence wins
      seen[fi] = true;
      char* end = nullptr;
      errno = 0;
      char* base = reinterpret_cast<char*>(p) + fd.offset;
      if (fd.type == 'f') {
        float v = std::strtof(val.c_str(), &end);
        if (end != val.c_str()) *reinterpret_cast<float*>(base) = v;
      } else if (fd.type == 'd') {
        double v = std::strtod(val.c_str(), &end);
        if (end != val.c_str()) *reinterpret_cast<double*>(base) = v;
      } else {
        double v = std::strtod(val.c_str(), &end);  // "1999", "1e5" both occur in the wild
        if (end != val.c_str()) *reinterpret_cast<int*>(base) = (int)v;
      }
      if (end == val.c_str()) {
        std::fclose(f);
        return CVO_B200_ERR_INVALID;
      }
      break;
    }
  }
  std::fclose(f);
  return CVO_B200_OK;
}
