#!/usr/bin/env python3
"""Recipe that builds oracle/_ref: the REFERENCE'S OWN kernels of the hot path, compiled from
the sources where they lie under /root/reference.  TEST INFRASTRUCTURE ONLY.

Nothing of the reference is copied into the repository: this script extracts the definitions
listed in WANTED (located by their signature and brace matching, line numbers are only checked
against the ones DESIGN.md cites) into the git-ignored oracle/_ref/, and compiles them together
with the committed harness oracle/ref_harness.cu (ours: a prelude that stands in for the
PCL/thrust/yaml-cpp headers the extracted text names, and extern "C" entry points).

Tier 1 (no third-party arithmetic at all, reference text verbatim):
  CvoParams POD                         include/UnifiedCvo/cvo/CvoParams.hpp:12-128
  SparseKernelMat POD                   include/UnifiedCvo/cvo/SparseKernelMat.hpp:5-19
  PointSegmentedDistribution (CvoPoint) include/UnifiedCvo/utils/PointSegmentedDistribution.hpp:17-99
  dot / squared_dist x2 / square_norm   include/UnifiedCvo/cvo/gpu_utils.cuh:24-41, 73-78, 97-104
  compute_range_ell                     src/cvo/CvoGPU.cu:86-90
  compute_geometric_type_ip             src/cvo/CvoGPU.cu:203-215
  fill_in_A_mat_gpu  (K1)               src/cvo/CvoGPU.cu:477-593
  A_sparsity_indicator_ell_update       src/cvo/CvoGPU.cu:1167-1285   (host code: std::queue only; host build)
Tier 2 (reference text verbatim, but its Eigen fixed-size 3-vector / 3x3 primitives are supplied
by oracle/ref_mini_eigen.h, OUR stand-in — sum order c0+(c1+c2), documented there):
  skew_gpu                              include/UnifiedCvo/cvo/gpu_utils.cuh:8-15
  mahananobis_distance                  src/cvo/CvoGPU.cu:151-169
  fill_in_A_mat_gpu_dense_mat_kernel    src/cvo/CvoGPU.cu:217-327   (K1b)
  compute_flow_gpu_no_eigen  (K2)       src/cvo/CvoGPU.cu:729-790
  compute_step_size_xi       (K3)       src/cvo/CvoGPU.cu:953-998
  compute_step_size_poly_coeff (K4)     src/cvo/CvoGPU.cu:1001-1082
  update_tf (the 4-argument overload)   src/cvo/CvoGPU.cu:94-112      (host build only; cudaMemcpy -> memcpy)
  transform_point_R_T                   src/cvo/CvoGPU_impl.cu:31-82  (host build only)
  transform_point_pose_vec              src/cvo/CvoGPU_impl.cu:84-150 (host build only; the frames of the
                                        multi-frame edge update)
  skew<T, RC_MAJOR>                     src/cvo/LieGroup.cpp:11-19    (host build only)
  Exp_SEK3 (float)                      src/cvo/LieGroup.cpp:245-274  (host build only: the pose
                                        increment of align_impl, CvoGPU.cu:1462)

Three builds (outputs only under oracle/_ref/):
  libcvo_ref_host.so         g++ -O2 -ffp-contract=off; __global__ -> plain function, the grid is a
                             host loop.  Runs in the CPU test-suite.
  libcvo_ref_cuda.so         nvcc -O3 -gencode arch=compute_100a,code=sm_100a with the reference's
                             own CUDA flags otherwise (CMakeLists.txt:29,79: Release, default
                             --fmad=true, no fast-math) -> the reference's real device arithmetic
                             (logf / exp(double) overloads, FMA contraction) on a B200.
  libcvo_ref_cuda_nofma.so   the same with --fmad=false: isolates what contraction changes.
"""
from __future__ import annotations

import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CVO_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")

# (file, anchor regex of the first line of the definition, cited first line, tier, output name)
WANTED = [
    ("include/UnifiedCvo/cvo/CvoParams.hpp", r"^\s*struct CvoParams \{", 12, 1, "CvoParams"),
    ("include/UnifiedCvo/cvo/SparseKernelMat.hpp", r"^\s*struct\s*$", 5, 1, "SparseKernelMat"),
    ("include/UnifiedCvo/utils/PointSegmentedDistribution.hpp",
     r"^\s*template <unsigned int FEATURE_DIM, unsigned int NUM_CLASS>\s*$", 17, 1, "PointSegmentedDistribution"),
    ("include/UnifiedCvo/cvo/gpu_utils.cuh", r"^\s*__device__ T dot\(const T \* a, const T\* b, int dim\)", 24, 1, "dot"),
    ("include/UnifiedCvo/cvo/gpu_utils.cuh", r"^\s*__device__ T squared_dist\(const T \* a, const T\* b, int dim\)", 34, 1, "squared_dist_arr"),
    ("include/UnifiedCvo/cvo/gpu_utils.cuh", r"^\s*__device__ float squared_dist\(const T & a, const T & b\)", 73, 1, "squared_dist_pt"),
    ("include/UnifiedCvo/cvo/gpu_utils.cuh", r"^\s*__device__ T square_norm\(const T \*a, int dim\)", 98, 1, "square_norm"),
    ("src/cvo/CvoGPU.cu", r"^\s*float compute_range_ell\(", 86, 1, "compute_range_ell"),
    ("src/cvo/CvoGPU.cu", r"^\s*float compute_geometric_type_ip\(", 204, 1, "compute_geometric_type_ip"),
    ("src/cvo/CvoGPU.cu", r"^\s*void fill_in_A_mat_gpu\(const CvoParams \* cvo_params,", 478, 1, "fill_in_A_mat_gpu"),
    ("src/cvo/CvoGPU.cu", r"^\s*static bool A_sparsity_indicator_ell_update\(std::queue<float> & indicator_start_queue,", 1167, 1,
     "A_sparsity_indicator_ell_update"),
    ("include/UnifiedCvo/cvo/gpu_utils.cuh", r"^\s*void skew_gpu\(", 10, 2, "skew_gpu"),
    ("src/cvo/CvoGPU.cu", r"^\s*float mahananobis_distance\(", 152, 2, "mahananobis_distance"),
    ("src/cvo/CvoGPU.cu", r"^\s*void fill_in_A_mat_gpu_dense_mat_kernel\(", 218, 2, "fill_in_A_mat_gpu_dense_mat_kernel"),
    ("src/cvo/CvoGPU.cu", r"^\s*__global__ void compute_flow_gpu_no_eigen\(", 729, 2, "compute_flow_gpu_no_eigen"),
    ("src/cvo/CvoGPU.cu", r"^\s*__global__ void compute_step_size_xi\(", 953, 2, "compute_step_size_xi"),
    ("src/cvo/CvoGPU.cu", r"^\s*__global__ void compute_step_size_poly_coeff\(", 1001, 2, "compute_step_size_poly_coeff"),
    ("src/cvo/CvoGPU.cu", r"^\s*void update_tf\(const Mat33f & R, const Vec3f & T,\s*$", 94, 2, "update_tf"),
    ("src/cvo/CvoGPU_impl.cu", r"^\s*struct transform_point_R_T : public thrust::unary_function<CvoPoint,CvoPoint>", 31, 2,
     "transform_point_R_T"),
    ("src/cvo/CvoGPU_impl.cu", r"^\s*struct transform_point_pose_vec : public thrust::unary_function<CvoPoint,CvoPoint>", 84, 2,
     "transform_point_pose_vec"),
    ("src/cvo/LieGroup.cpp", r"^\s*Eigen::Matrix<T, 3, 3, RC_MAJOR> skew\(const Eigen::Matrix<T, 3, 1>& v\) \{", 12, 2, "skew"),
    ("src/cvo/LieGroup.cpp", r"^\s*Eigen::Matrix<float, 3, 4> Exp_SEK3\(const Eigen::Matrix<float, 6,1>& v, float dt\) \{", 245, 2, "Exp_SEK3"),
]


def strip_comments_for_matching(line: str, in_block: bool) -> tuple[str, bool]:
    """Returns the line with comments and string literals blanked (brace counting only)."""
    out = []
    i = 0
    n = len(line)
    while i < n:
        if in_block:
            j = line.find("*/", i)
            if j < 0:
                return "".join(out), True
            i = j + 2
            in_block = False
            continue
        c = line[i]
        if line.startswith("//", i):
            break
        if line.startswith("/*", i):
            in_block = True
            i += 2
            continue
        if c == '"':
            j = i + 1
            while j < n and line[j] != '"':
                j += 2 if line[j] == "\\" else 1
            i = j + 1
            continue
        if c == "'":
            j = i + 1
            while j < n and line[j] != "'":
                j += 2 if line[j] == "\\" else 1
            i = j + 1
            continue
        out.append(c)
        i += 1
    return "".join(out), in_block


def extract(path: str, anchor: str, cited: int) -> tuple[str, int, int]:
    """The definition whose first line matches `anchor`, through its closing brace (and the `;`
    of a struct).  Preceding qualifier lines (`template<...>`, `__device__`, `inline`,
    `__global__`) belonging to it are included."""
    with open(os.path.join(REF, path)) as f:
        lines = f.read().split("\n")
    rx = re.compile(anchor)
    hits = [i for i, l in enumerate(lines) if rx.search(l)]
    if not hits:
        raise SystemExit(f"make_ref: anchor {anchor!r} not found in {path}")
    # several overloads can share a prefix; take the hit closest to the cited line
    start = min(hits, key=lambda i: abs(i + 1 - cited))
    if abs(start + 1 - cited) > 3:
        raise SystemExit(f"make_ref: {path}: anchor found at line {start + 1}, cited {cited}: the "
                         "reference moved; re-check DESIGN.md's citations")
    first = start
    quals = re.compile(r"^\s*(template\s*<.*>|__device__|__host__|__global__|inline|static|__host__ __device__ __forceinline__.*)\s*$")
    while first > 0 and quals.match(lines[first - 1]):
        first -= 1
    depth = 0
    seen = False
    in_block = False
    end = None
    pp_stack = []  # True while inside the #else branch of a conditional (its braces duplicate the #if's)
    for i in range(start, len(lines)):
        code, in_block = strip_comments_for_matching(lines[i], in_block)
        pp = code.strip()
        if pp.startswith("#"):
            d = pp[1:].strip()
            if d.startswith("if"):
                pp_stack.append(False)
            elif d.startswith("else") or d.startswith("elif"):
                if pp_stack:
                    pp_stack[-1] = True
            elif d.startswith("endif"):
                if pp_stack:
                    pp_stack.pop()
            continue
        if any(pp_stack):
            continue
        for ch in code:
            if ch == "{":
                depth += 1
                seen = True
            elif ch == "}":
                depth -= 1
        if seen and depth == 0:
            end = i
            break
    if end is None:
        raise SystemExit(f"make_ref: unbalanced braces after {path}:{start + 1}")
    return "\n".join(lines[first:end + 1]), first + 1, end + 1


def run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise SystemExit(f"make_ref: build step failed ({cmd[0]})")


def main() -> int:
    if not os.path.isdir(os.path.join(REF, "src", "cvo")):
        # the GPU box: only the prebuilt files under oracle/_ref travel
        print(f"make_ref: {REF} not present; keeping whatever oracle/_ref already holds")
        return 0
    os.makedirs(OUT, exist_ok=True)
    manifest = []
    for path, anchor, cited, tier, name in WANTED:
        text, a, b = extract(path, anchor, cited)
        with open(os.path.join(OUT, name + ".inc"), "w") as f:
            f.write(f"// GENERATED by oracle/make_ref.py: {path}:{a}-{b} of the reference, verbatim.\n"
                    "// Reference text, NOT part of this repository (oracle/_ref/ is git-ignored).\n")
            f.write(text + "\n")
        manifest.append(f"{name}: {path}:{a}-{b} tier{tier}")
    with open(os.path.join(OUT, "MANIFEST.txt"), "w") as f:
        f.write("\n".join(manifest) + "\n")

    harness = os.path.join(HERE, "ref_harness.cu")
    inc = ["-I", OUT, "-I", HERE]
    defs = ["-DNUM_CLASSES=19", "-DFEATURE_DIMENSIONS=5", "-DCUDA_BLOCK_SIZE=512",
            "-DCVO_POINT_NEIGHBORS=256"]  # CMakeLists.txt:498
    # (1) host build: system g++ (the image's /opt/gcc wrapper lacks libgomp.spec)
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    run([gxx, "-x", "c++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp",
         "-fPIC", "-shared", "-DCVO_REF_HOST_BUILD=1", *defs, *inc, harness,
         "-o", os.path.join(OUT, "libcvo_ref_host.so")])
    # (2) device builds
    arch = ["-gencode", "arch=compute_100a,code=sm_100a"]
    common = ["nvcc", "-std=c++17", "-O3", "-DNDEBUG", "--expt-extended-lambda", "-lineinfo", *arch,
              "-Xcompiler", "-fPIC", "-shared", *defs, *inc, harness, "-lcudart"]
    run([*common, "-o", os.path.join(OUT, "libcvo_ref_cuda.so")])
    run([*common, "--fmad=false", "-DCVO_REF_NOFMA=1", "-o", os.path.join(OUT, "libcvo_ref_cuda_nofma.so")])
    # PTX of the reference-flag build: documents which multiply-adds nvcc contracts
    run(["nvcc", "-std=c++17", "-O3", "-DNDEBUG", "--expt-extended-lambda", *arch, *defs, *inc,
         "-ptx", harness, "-o", os.path.join(OUT, "ref_harness.ptx")])
    print("make_ref: built oracle/_ref/{libcvo_ref_host,libcvo_ref_cuda,libcvo_ref_cuda_nofma}.so from",
          len(manifest), "reference definitions")
    return 0


if __name__ == "__main__":
    sys.exit(main())
