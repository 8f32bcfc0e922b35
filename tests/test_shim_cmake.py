"""shim/CMakeLists.txt as a package: the standalone build configures, builds, installs and a
find_package(UnifiedCvo) consumer links UnifiedCvo::cvo_gpu_img_lib (README.md:85-100 of the
reference; its install/export block is CMakeLists.txt:778-839).  Eigen/PCL are not in this image,
so the build runs with CVO_B200_SHIM_STUBS=ON (the stand-in declarations of shim/stubs); the
members the reference's own kept sources define are supplied by a small extra source."""
import os
import shutil
import subprocess

import pytest

from helpers import ROOT

CMAKE = shutil.which("cmake")
LIB = os.path.join(ROOT, "unified_cvo_b200", "csrc", "libcvo_b200.so")

KEPT = r"""
namespace cvo {
float CvoGPU::inner_product_cpu(const CvoPointCloud&, const CvoPointCloud&, const Eigen::Matrix4f&, float) const { return 0.f; }
CvoFrame::CvoFrame(const CvoPointCloud* pts, const double poses[12]) : points(pts) { for (int i = 0; i < 12; i++) pose_vec[i] = poses[i]; }
void CvoFrame::transform_pointcloud() {}
void BinaryStateGPU::update_ell() {}
void BinaryStateGPU::add_residual_to_problem(ceres::Problem&) {}
BinaryStateCPU::BinaryStateCPU(std::shared_ptr<CvoFrame>, std::shared_ptr<CvoFrame>, const CvoParams*) {}
int BinaryStateCPU::update_inner_product() { return 0; }
void BinaryStateCPU::add_residual_to_problem(ceres::Problem&) {}
void BinaryStateCPU::update_ell() {}
CvoBatchIRLS::CvoBatchIRLS(const std::vector<std::shared_ptr<CvoFrame>>&, const std::vector<bool>&, const std::list<std::shared_ptr<BinaryState>>&, const CvoParams*) {}
void CvoBatchIRLS::solve() {}
int CvoGPU::align(std::vector<std::shared_ptr<CvoFrame>>&, const std::vector<bool>&, const std::list<std::shared_ptr<BinaryState>>&, double*) const { return 0; }
}
"""

CONSUMER_CMAKE = """cmake_minimum_required(VERSION 3.18)
project(consumer LANGUAGES CXX)
find_package(UnifiedCvo REQUIRED)
add_executable(consumer main.cpp)
target_compile_options(consumer PRIVATE -include {stub})
target_include_directories(consumer PRIVATE {inc})
target_link_libraries(consumer PRIVATE UnifiedCvo::cvo_gpu_img_lib)
"""

CONSUMER_MAIN = r"""
#include <cstdio>
int main(int argc, char** argv) {
  if (argc > 1) { cvo::CvoGPU g(argv[1]); std::printf("ell_init %f\n", g.get_params().ell_init); }
  std::printf("NUM_CLASSES %d FEATURE_DIMENSIONS %d\n", NUM_CLASSES, FEATURE_DIMENSIONS);
  return 0;
}
"""


@pytest.mark.skipif(CMAKE is None or not os.path.exists(LIB), reason="cmake or libcvo_b200.so missing")
def test_standalone_package_installs_and_is_found(tmp_path):
    kept = tmp_path / "kept.cpp"
    kept.write_text(KEPT)
    build, prefix = tmp_path / "build", tmp_path / "prefix"
    env = dict(os.environ, CC="/usr/bin/gcc", CXX="/usr/bin/g++")

    def run(*cmd, cwd=None):
        out = subprocess.run(cmd, capture_output=True, text=True, cwd=cwd, env=env)
        assert out.returncode == 0, (cmd, out.stdout[-2000:], out.stderr[-3000:])
        return out.stdout

    run(CMAKE, "-S", os.path.join(ROOT, "shim"), "-B", str(build), "-DCVO_B200_SHIM_STUBS=ON",
        f"-DCVO_B200_KEEP_SOURCES={kept}", f"-DCMAKE_INSTALL_PREFIX={prefix}", "-DCMAKE_BUILD_TYPE=Release")
    run(CMAKE, "--build", str(build), "-j4")
    run(CMAKE, "--install", str(build))
    for rel in ("cmake/UnifiedCvoConfig.cmake", "cmake/UnifiedCvoConfigVersion.cmake",
                "cmake/UnifiedCvoTargets.cmake", "lib/UnifiedCvo-0.1/libcvo_gpu_img_lib.so",
                "lib/UnifiedCvo-0.1/libcvo_b200.so", "include/UnifiedCvo-0.1/cvo_b200.h"):
        assert (prefix / rel).exists(), rel
    targets = (prefix / "cmake" / "UnifiedCvoTargets.cmake").read_text()
    assert "UnifiedCvo::cvo_gpu_img_lib" in targets
    assert "NUM_CLASSES=19" in targets and "FEATURE_DIMENSIONS=5" in targets  # PUBLIC definitions travel
    assert "cvo_b200" not in targets.replace("libcvo_b200", "")  # no foreign target leaks into the export
    # a find_package consumer (README.md:85-100)
    cons = tmp_path / "consumer"
    cons.mkdir()
    (cons / "CMakeLists.txt").write_text(CONSUMER_CMAKE.format(
        stub=os.path.join(ROOT, "shim", "stubs", "reference_api_stub.hpp"), inc=os.path.join(ROOT, "include")))
    (cons / "main.cpp").write_text(CONSUMER_MAIN)
    run(CMAKE, "-S", str(cons), "-B", str(cons / "b"), f"-DUnifiedCvo_DIR={prefix / 'cmake'}")
    run(CMAKE, "--build", str(cons / "b"))
    out = run(str(cons / "b" / "consumer"))
    assert "NUM_CLASSES 19 FEATURE_DIMENSIONS 5" in out
    # the installed library finds libcvo_b200.so next to itself ($ORIGIN), not in the build tree
    ldd = subprocess.run(["ldd", str(prefix / "lib/UnifiedCvo-0.1/libcvo_gpu_img_lib.so")],
                         capture_output=True, text=True).stdout
    assert str(prefix / "lib/UnifiedCvo-0.1/libcvo_b200.so") in ldd, ldd
