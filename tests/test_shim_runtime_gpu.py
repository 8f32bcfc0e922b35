"""The reference-facing C++ classes, EXECUTED on a B200.

shim/CvoGPU_b200.cpp + shim/IRLS_State_GPU_b200.cpp are what a maintainer compiles into
cvo_gpu_img_lib in place of the reference's CUDA sources.  Eigen / PCL are not in this image, so
they are built here against shim/stubs/reference_api_stub.hpp (containers with real storage,
column-major like Eigen's) together with tests/shim_runtime/driver.cpp, which makes the calls a
reference driver makes: CvoGPU(yaml), get_params/write_params, align(CvoPointCloud...), the three
calls that reuse the pairwise pass, and the multi-frame align(frames, consts, edges) of
CvoGPU.cu:1637-1686 -> BinaryStateGPU::update_inner_product.  Every number it prints must equal
what the Python mirror gets from the same libcvo_b200.so: the marshalling (column-major Eigen
features/labels -> row-major arrays, Matrix4f <-> float[16], double pose -> float[12], CSR ->
Association / the row-strided SparseKernelMat) is what is under test."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

import unified_cvo_b200 as u
from helpers import DATA, ROOT, synthetic_pair
from test_multiframe import pose_rt

pytestmark = pytest.mark.gpu
GXX = shutil.which("g++") or "/usr/bin/g++"
CSRC = os.path.join(ROOT, "unified_cvo_b200", "csrc")


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    exe = tmp_path_factory.mktemp("shim") / "shim_driver"
    cmd = [GXX, "-std=c++17", "-O1", "-fPIC", "-Wall", "-DCVO_SHIM_SYNTAX_CHECK", "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "shim"), "-include", os.path.join(ROOT, "shim", "stubs", "reference_api_stub.hpp"),
           os.path.join(ROOT, "tests", "shim_runtime", "driver.cpp"), os.path.join(ROOT, "shim", "CvoGPU_b200.cpp"),
           os.path.join(ROOT, "shim", "IRLS_State_GPU_b200.cpp"), "-o", str(exe), "-L" + CSRC, "-lcvo_b200",
           "-Wl,-rpath," + CSRC]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]
    return str(exe)


def write_cloud(path, pc):
    n, F, C = pc.num_points(), pc.feature_dimensions(), pc.num_classes()
    geo = pc.geometric_types_
    with open(path, "wb") as fh:
        fh.write(struct.pack("<4i", n, F, C, 1 if geo is not None else 0))
        fh.write(np.ascontiguousarray(pc.positions_, np.float32).tobytes())
        if F:
            fh.write(np.ascontiguousarray(pc.features_, np.float32).tobytes())
        if C:
            fh.write(np.ascontiguousarray(pc.labels_, np.float32).tobytes())
        if geo is not None:
            fh.write(np.ascontiguousarray(geo, np.float32).tobytes())


def as_reserved(pc):
    """What the C++ driver's reserve/add_point cloud holds: geometric types are ALWAYS present
    (zeros when the source had none, CvoPointCloud.cpp:1393)."""
    geo = pc.geometric_types_ if pc.geometric_types_ is not None else np.zeros((pc.num_points(), 2), np.float32)
    return u.CvoPointCloud(pc.positions_, pc.features_, pc.labels_, geo)


def run(driver, *args, env=None):
    out = subprocess.run([driver, *args], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, **(env or {})))
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-2000:])
    rows = {}
    for line in out.stdout.splitlines():
        if not line.strip():
            continue
        k, *v = line.split()
        rows.setdefault(k, []).append(v)
    return rows


def checksum(assoc):
    rows = np.repeat(np.arange(len(assoc.row_ptr) - 1), np.diff(assoc.row_ptr))
    return float(np.sum((rows + 1).astype(np.float64) * (assoc.cols.astype(np.float64) + 1) * assoc.vals.astype(np.float64)))


@pytest.mark.parametrize("flavour", ["geometric", "colour+semantics"])
def test_two_cloud_calls_through_the_cpp_class_equal_the_python_mirror(driver, tmp_path, flavour):
    if flavour == "geometric":
        src, tgt, _ = synthetic_pair(2500, 2000, 2200, 20002)
        yaml = os.path.join(DATA, "cvo_geometric_params_img_gpu0.yaml")
    else:
        src, tgt, _ = synthetic_pair(2500, 2000, 2200, 31, F=5, C=19, geotype=True)
        # random colours and labels: loosen the kernels so that pairs survive them (the first
        # occurrence of a key wins, in the reference's reader and in ours)
        yaml = str(tmp_path / "params.yaml")
        with open(os.path.join(DATA, "cvo_semantic_params_img_gpu0.yaml")) as fh:
            text = fh.read()
        with open(yaml, "w") as fh:
            fh.write("c_ell: 1.0\nsp_thres: 0.001\nis_using_geometric_type: 1\nMAX_ITER: 40\n" + text)
    write_cloud(tmp_path / "s.bin", src)
    write_cloud(tmp_path / "t.bin", tgt)
    got = run(driver, "align", yaml, str(tmp_path / "s.bin"), str(tmp_path / "t.bin"))
    assert got["points"][0][:2] == ["2000", "2200"]
    # the same calls through the Python mirror of the class, on the same library
    p = u.read_params_yaml(yaml)
    p.is_exporting_association = 1
    g = u.CvoGPU(p)
    s2, t2 = as_reserved(src), as_reserved(tgt)
    assoc = u.Association()
    ret, T, info = g.align(s2, t2, np.eye(4, dtype=np.float32), association=assoc)
    assert int(got["align_ret"][0][0]) == ret and got["align_ret"][0][2] == "1"
    T_cpp = np.array([float.fromhex(x) for x in got["transform"][0]], np.float32).reshape(4, 4).T  # column-major
    assert np.array_equal(T_cpp, np.asarray(T, np.float32)), np.abs(T_cpp - T).max()
    a = got["association"][0]
    assert int(a[1]) == len(assoc.vals) > 0 and int(a[3]) == len(assoc.source_inliers) and int(a[5]) == len(assoc.vals)
    assert float(a[7]) == pytest.approx(checksum(assoc), rel=1e-12)
    I = np.eye(4, dtype=np.float32)
    assert float(got["inner_product"][0][0]) == pytest.approx(g.inner_product_gpu(s2, t2, I, 0.8), rel=1e-6)
    assert float(got["function_angle"][0][0]) == pytest.approx(g.function_angle(s2, t2, I, 0.8, True, True), rel=1e-6)
    a2 = g.compute_association_gpu(s2, t2, I, 0.8)
    assert int(got["association2"][0][1]) == len(a2.vals) > 20
    assert float(got["association2"][0][5]) == pytest.approx(checksum(a2), rel=1e-12)
    K = np.array([[0.30, 0.02, 0.0], [0.02, 0.20, 0.01], [0.0, 0.01, 0.50]], np.float32)
    a3 = g.compute_association_gpu(s2, t2, I, K)
    assert int(got["association3"][0][1]) == len(a3.vals)
    assert float(got["association3"][0][5]) == pytest.approx(checksum(a3), rel=1e-12)
    g.close()


def test_multiframe_align_through_the_cpp_classes_equals_the_python_mirror(driver, tmp_path):
    """CvoGPU::align(frames, consts, edges) -> BinaryStateGPU per edge -> (stand-in) CvoBatchIRLS loop:
    two outer iterations of update_inner_product + update_ell over a ring of four frames; nnz, the
    row-strided A_result_cpu_ (walked like IRLS_State_GPU.cpp:14-45) and the cap schedule equal the
    Python BinaryStateGPU's."""
    # multiframe_using_cpu defaults to 1 (CvoParams.hpp): then the overload builds BinaryStateCPU
    # edges (kept reference code); the GPU edges are what is under test here
    yaml = str(tmp_path / "params.yaml")
    with open(os.path.join(DATA, "cvo_geometric_params_img_gpu0.yaml")) as fh:
        text = fh.read()
    with open(yaml, "w") as fh:
        fh.write("multiframe_using_cpu: 0\nmultiframe_ell_init: 0.5\nmultiframe_num_neighbors: 24\n"
                 "multiframe_ell_min: 0.1\nmultiframe_ell_decay_rate: 0.8\n" + text)
    clouds, poses, paths = [], [], []
    for k in range(4):
        a, _, _ = synthetic_pair(1600, 900 + 64 * k, 900, 40)
        clouds.append(a)
        poses.append(pose_rt(0.0, 0.3 * k, 0.0, [0.01 * k, 0.0, 0.02 * k]))
        write_cloud(tmp_path / f"f{k}.bin", a)
        paths.append(str(tmp_path / f"f{k}.bin"))
    with open(tmp_path / "poses.bin", "wb") as fh:
        fh.write(np.ascontiguousarray(np.stack(poses), np.float64).tobytes())
    got = run(driver, "edges", yaml, str(tmp_path / "poses.bin"), *paths)
    assert got["multiframe_ret"][0][0] == "0" and got["multiframe_ret"][0][2] == "1"
    p = u.read_params_yaml(yaml)
    g = u.CvoGPU(p)
    frames = [u.CvoFrameGPU(g, as_reserved(c), P) for c, P in zip(clouds, poses)]
    states = [u.BinaryStateGPU(frames[k], frames[(k + 1) % 4]) for k in range(4)]  # multiframe_* defaults
    lines = got["edge"]
    assert len(lines) == 8
    i = 0
    for outer in range(2):
        for e, st in enumerate(states):
            nnz = st.update_inner_product()
            row = lines[i]
            i += 1
            assert (int(row[0]), int(row[1])) == (outer, e)
            assert int(row[3]) == nnz == int(row[5]) and nnz > 100
            assert int(row[7]) == len(st.A_result_cpu_.source_inliers)
            assert float(row[9]) == pytest.approx(checksum(st.A_result_cpu_), rel=1e-12)
            st.update_ell()
    g.close()
    # the same loop with IRLS.cpp:111-121 replaced by cvo::update_inner_product_batch (cvo_b200_batch.hpp):
    # one device call per outer iteration, the same matrices in every edge
    got_b = run(driver, "edges", yaml, str(tmp_path / "poses.bin"), *paths, env={"SHIM_DRIVER_BATCH": "1"})
    assert len(got_b["batch"]) == 2
    for outer in range(2):
        assert int(got_b["batch"][outer][2]) == sum(int(r[5]) for r in lines[4 * outer:4 * outer + 4])
    for a, b in zip(lines, got_b["edge"]):
        assert a[:2] == b[:2] and a[4:] == b[4:]  # entries, rows, checksum (nnz is not returned per edge there)
