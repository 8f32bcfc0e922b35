// cvo_align_gpu_two_color_pcd — the reference's README demo (BASELINE configs[0]) as a
// dependency-free C++17 program over the C-ABI (include/cvo_b200.h): no Eigen, PCL or boost.
//
// Same command line and the same steps as src/experiments/main_cvo_gpu_align_two_color_pcd.cpp:
//   cvo_align_gpu_two_color_pcd source.pcd target.pcd cvo_params.yaml [ell_init]
//   :40-50  load the two PCDs, centroids, dist = |mean(source) - mean(target)|
//   :56-66  ell_init = dist (or argv[4]); first-frame decay rate / start; write_params
//   :70-82  init guess = identity; align; print the transform
//   :86-108 before_align.pcd / after_align.pcd = source + target moved by identity / result
// Where Eigen and PCL exist, the reference's own driver links against shim/ instead
// (INTEGRATION.md); this file is the same demo for boxes that have neither.
//
// Build: make -C examples   (needs only libcvo_b200.so)
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "cvo_b200.h"

namespace {

// What CvoPointCloud(const pcl::PointCloud<pcl::PointXYZRGB>&) holds (CvoPointCloud.cpp:570-594):
// positions, features = (r, g, b) / 255, 0, 0 (FEATURE_DIMENSIONS = 5), geometric type (0, 1).
// XYZ-only files give what the PointXYZ constructor holds (:634-652): no features, type (1, 0).
struct Cloud {
  int n = 0, F = 0;
  std::vector<float> xyz, feat, geo;
  std::vector<uint32_t> rgb;  // packed 0x00RRGGBB, kept for the output files
};

bool load_pcd(const std::string& path, Cloud& c, std::string& err) {
  std::ifstream in(path);
  if (!in) { err = "cannot open " + path; return false; }
  std::vector<std::string> fields;
  long points = -1;
  std::string line;
  bool data = false;
  while (std::getline(in, line)) {
    std::istringstream ss(line);
    std::string key;
    if (!(ss >> key) || key[0] == '#') continue;
    if (key == "FIELDS") { std::string f; while (ss >> f) fields.push_back(f); }
    else if (key == "POINTS") ss >> points;
    else if (key == "DATA") {
      std::string kind; ss >> kind;
      if (kind != "ascii") { err = path + ": only ASCII PCD files are supported"; return false; }
      data = true;
      break;
    }
  }
  int ix = -1, iy = -1, iz = -1, ic = -1;
  for (size_t k = 0; k < fields.size(); k++) {
    if (fields[k] == "x") ix = (int)k;
    if (fields[k] == "y") iy = (int)k;
    if (fields[k] == "z") iz = (int)k;
    if (fields[k] == "rgb") ic = (int)k;
  }
  if (!data || points < 0 || ix < 0 || iy < 0 || iz < 0) { err = path + ": malformed PCD header"; return false; }
  c.n = (int)points;
  c.F = ic >= 0 ? 5 : 0;
  c.xyz.resize((size_t)c.n * 3);
  c.geo.resize((size_t)c.n * 2);
  if (c.F) { c.feat.assign((size_t)c.n * 5, 0.f); c.rgb.resize((size_t)c.n); }
  std::vector<std::string> tok(fields.size());
  for (int i = 0; i < c.n; i++) {
    if (!std::getline(in, line)) { err = path + ": fewer points than POINTS says"; return false; }
    std::istringstream ss(line);
    for (auto& t : tok)
      if (!(ss >> t)) { err = path + ": short data line"; return false; }
    c.xyz[3 * (size_t)i + 0] = (float)std::strtod(tok[ix].c_str(), nullptr);
    c.xyz[3 * (size_t)i + 1] = (float)std::strtod(tok[iy].c_str(), nullptr);
    c.xyz[3 * (size_t)i + 2] = (float)std::strtod(tok[iz].c_str(), nullptr);
    if (c.F) {
      // TYPE U (packed integer) or TYPE F (the same 32 bits printed as a float)
      uint32_t packed;
      if (tok[ic].find_first_of(".eE") != std::string::npos) {
        float f = std::strtof(tok[ic].c_str(), nullptr);
        std::memcpy(&packed, &f, sizeof(packed));
      } else {
        packed = (uint32_t)std::strtoull(tok[ic].c_str(), nullptr, 10);
      }
      c.rgb[i] = packed & 0xffffffu;
      c.feat[5 * (size_t)i + 0] = (float)(((packed >> 16) & 255) / 255.0);
      c.feat[5 * (size_t)i + 1] = (float)(((packed >> 8) & 255) / 255.0);
      c.feat[5 * (size_t)i + 2] = (float)((packed & 255) / 255.0);
      c.geo[2 * (size_t)i + 0] = 0.f; c.geo[2 * (size_t)i + 1] = 1.f;
    } else {
      c.geo[2 * (size_t)i + 0] = 1.f; c.geo[2 * (size_t)i + 1] = 0.f;
    }
  }
  return true;
}

// get_pc_mean (main_cvo_gpu_align_two_color_pcd.cpp:26-32): running float sum, then / n
void cloud_mean(const Cloud& c, float m[3]) {
  m[0] = m[1] = m[2] = 0.f;
  for (int i = 0; i < c.n; i++)
    for (int k = 0; k < 3; k++) m[k] = m[k] + c.xyz[3 * (size_t)i + k];
  for (int k = 0; k < 3; k++) m[k] = m[k] / (float)c.n;
}

// source + target moved by the column-major 4x4 T, as an ASCII PCD (:86-105)
bool save_sum_pcd(const std::string& path, const Cloud& src, const Cloud& tgt, const float T[16]) {
  std::FILE* fp = std::fopen(path.c_str(), "w");
  if (!fp) return false;
  const bool colour = src.F && tgt.F;
  const long n = (long)src.n + tgt.n;
  std::fprintf(fp, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z%s\nSIZE 4 4 4%s\n"
                   "TYPE F F F%s\nCOUNT 1 1 1%s\nWIDTH %ld\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %ld\nDATA ascii\n",
               colour ? " rgb" : "", colour ? " 4" : "", colour ? " U" : "", colour ? " 1" : "", n, n);
  for (int j = 0; j < tgt.n; j++) {
    const float* y = &tgt.xyz[3 * (size_t)j];
    float p[3];
    for (int r = 0; r < 3; r++) p[r] = T[r] * y[0] + T[4 + r] * y[1] + T[8 + r] * y[2] + T[12 + r];
    if (colour) std::fprintf(fp, "%.8g %.8g %.8g %u\n", p[0], p[1], p[2], tgt.rgb[j]);
    else std::fprintf(fp, "%.8g %.8g %.8g\n", p[0], p[1], p[2]);
  }
  for (int i = 0; i < src.n; i++) {
    const float* x = &src.xyz[3 * (size_t)i];
    if (colour) std::fprintf(fp, "%.8g %.8g %.8g %u\n", x[0], x[1], x[2], src.rgb[i]);
    else std::fprintf(fp, "%.8g %.8g %.8g\n", x[0], x[1], x[2]);
  }
  std::fclose(fp);
  return true;
}

int fail(const char* what, int rc, const cvo_b200_handle* h) {
  std::fprintf(stderr, "%s failed (%d): %s\n", what, rc, h ? cvo_b200_last_error(h) : cvo_b200_global_error());
  return 2;
}

}  // namespace

int main(int argc, char* argv[]) {
  if (argc < 4) {
    std::fprintf(stderr, "usage: %s source.pcd target.pcd cvo_params.yaml [ell_init]\n", argv[0]);
    return 1;
  }
  Cloud source, target;
  std::string err;
  if (!load_pcd(argv[1], source, err) || !load_pcd(argv[2], target, err)) {
    std::fprintf(stderr, "%s\n", err.c_str());
    return 1;
  }
  float ms[3], mt[3];
  cloud_mean(source, ms);
  cloud_mean(target, mt);
  const float dx = ms[0] - mt[0], dy = ms[1] - mt[1], dz = ms[2] - mt[2];
  const float dist = std::sqrt(dx * dx + dy * dy + dz * dz);
  std::printf("source mean is %g %g %g, target mean is %g %g %g, dist is %g\n", ms[0], ms[1], ms[2], mt[0], mt[1],
              mt[2], dist);

  cvo_b200_params p;
  cvo_b200_params_default(&p);
  int rc = cvo_b200_params_read_yaml(argv[3], &p);
  if (rc != CVO_B200_OK) return fail("cvo_b200_params_read_yaml", rc, nullptr);
  p.ell_init = argc > 4 ? std::strtof(argv[4], nullptr) : dist;   // :57-59
  p.ell_decay_rate = p.ell_decay_rate_first_frame;                 // :60
  p.ell_decay_start = p.ell_decay_start_first_frame;               // :61
  if (!source.F || !target.F) p.is_using_intensity = 0;            // main_cvo_gpu_align_two_pcd.cpp:66
  std::printf("write ell! ell init is %g\n", p.ell_init);

  cvo_b200_handle* h = nullptr;
  rc = cvo_b200_create(&p, /*device=*/0, &h);
  if (rc != CVO_B200_OK) return fail("cvo_b200_create", rc, nullptr);

  const float I16[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};  // init guess (:70-76)
  float result[16];
  cvo_b200_align_info info;
  std::memset(&info, 0, sizeof(info));
  std::printf("Start align... num_fixed is %d, num_moving is %d\n", source.n, target.n);
  std::fflush(stdout);
  rc = cvo_b200_align_host(h, source.n, source.xyz.data(), source.F, source.F ? source.feat.data() : nullptr, 0,
                           nullptr, source.geo.data(), target.n, target.xyz.data(),
                           target.F ? target.feat.data() : nullptr, nullptr, target.geo.data(), I16, result, &info);
  if (rc != CVO_B200_OK) {
    int code = fail("cvo_b200_align_host", rc, h);
    cvo_b200_destroy(h);
    return code;
  }
  std::printf("Transform is\n");
  for (int r = 0; r < 4; r++)  // column-major storage, printed row by row like Eigen's operator<<
    std::printf("%11.6g %11.6g %11.6g %11.6g\n", result[r], result[4 + r], result[8 + r], result[12 + r]);
  std::printf("\nalign returned %d after %d iterations\n", info.ret, info.iterations);
  if (!save_sum_pcd("before_align.pcd", source, target, I16) || !save_sum_pcd("after_align.pcd", source, target, result))
    std::fprintf(stderr, "could not write before_align.pcd / after_align.pcd\n");
  std::printf("num of points before and after alignment is %d, %d\n", source.n + target.n, source.n + target.n);
  std::printf("Average registration time is %g\n", info.registration_seconds);
  cvo_b200_destroy(h);
  return 0;
}
