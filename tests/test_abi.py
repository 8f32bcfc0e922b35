"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every
symbol include/cvo_b200.h declares, its structs match the ctypes mirror, the parameter
reader behaves like read_CvoParams_yaml, and GPU entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import unified_cvo_b200 as u
from unified_cvo_b200 import _abi
from helpers import DATA, ROOT


def test_header_symbols_all_exported_and_bound():
    hdr = open(os.path.join(ROOT, "include", "cvo_b200.h")).read()
    declared = set(re.findall(r"\b(cvo_b200_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"cvo_b200_handle"}
    assert declared == set(_abi.SYMBOLS), declared ^ set(_abi.SYMBOLS)
    lib = _abi.load_library()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.cvo_b200_abi_version() == 1


def test_struct_sizes_match_the_compiled_library():
    lib = _abi.load_library()
    assert lib.cvo_b200_sizeof(0) == C.sizeof(_abi.Params)
    assert lib.cvo_b200_sizeof(1) == C.sizeof(_abi.IterTrace) == 232
    assert lib.cvo_b200_sizeof(2) == C.sizeof(_abi.AlignInfo)
    assert lib.cvo_b200_sizeof(99) == -1


def test_param_defaults_follow_the_reference_constructor():
    p = u.default_params()  # CvoParams.hpp:75-126
    assert p.ell_init == pytest.approx(0.5) and p.ell_min == pytest.approx(0.05)
    assert p.sigma == pytest.approx(0.1) and p.sp_thres == pytest.approx(0.0006)
    assert p.c == 7.0 and p.d == 7.0 and p.MAX_ITER == 10000
    assert p.nearest_neighbors_max == 512 and p.indicator_window_size == 15
    assert p.is_using_geometry == 1 and p.is_using_intensity == 0 and p.is_using_kdtree == 0
    assert p.multiframe_least_squares_num_threads == 24


def test_yaml_reader_matches_shipped_files_and_first_duplicate_wins():
    p = u.read_params_yaml(os.path.join(DATA, "cvo_outdoor_params.yaml"))
    assert p.ell_init == pytest.approx(0.2) and p.sp_thres == pytest.approx(0.007)
    assert p.MAX_ITER == 100000 and p.max_step == pytest.approx(0.01)
    assert p.nearest_neighbors_max == 256 and p.is_using_geometric_type == 1
    assert p.ell_decay_rate == pytest.approx(0.8) and p.indicator_window_size == 10
    assert p.multiframe_ell_init == pytest.approx(4.0)  # "4  #2.5" -> comment stripped
    q = u.read_params_yaml(os.path.join(DATA, "cvo_intensity_params_img_gpu0.yaml"))
    assert q.nearest_neighbors_max == 256  # duplicate key (256 then 512): first occurrence wins
    assert q.c_ell == pytest.approx(0.05) and q.is_using_intensity == 1
    r = u.read_params_yaml(os.path.join(DATA, "cvo_rgbd_params.yaml"))
    assert r.max_step == pytest.approx(0.8) and r.MAX_ITER == 2000 and r.c_sigma == pytest.approx(0.6)
    with pytest.raises(u.CvoError):
        u.read_params_yaml(os.path.join(DATA, "does_not_exist.yaml"))


def test_pcd_reader_and_xyzrgb_constructor():
    pc = u.CvoPointCloud.from_pcd(os.path.join(DATA, "source.pcd"))
    assert pc.num_points() == 523 and pc.feature_dimensions() == 5 and pc.num_classes() == 0
    # first point: rgb 4290691523 = 0xFFBEC1C3 -> (190,193,195)/255  (CvoPointCloud.cpp:583-587)
    np.testing.assert_allclose(pc.features_[0], [190 / 255, 193 / 255, 195 / 255, 0, 0], rtol=1e-6)
    assert np.all(pc.geometric_types_ == [0, 1])
    xyz_only = u.CvoPointCloud.from_pcd(os.path.join(DATA, "target.pcd"), use_color=False)
    assert xyz_only.num_points() == 1080 and xyz_only.feature_dimensions() == 0
    assert np.all(xyz_only.geometric_types_ == [1, 0])
    moved = u.CvoPointCloud.transform(np.eye(4), pc)
    assert np.array_equal(moved.positions_, pc.positions_)
    assert (pc + xyz_only).num_points() == 1603


def test_no_cpu_fallback_without_a_gpu():
    lib = _abi.load_library()
    if lib.cvo_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(u.CvoError):
        u.CvoGPU(u.default_params())


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "unified_cvo_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(base, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "cvo_oracle" not in txt, f


REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "cvo_params")), reason="reference tree not present")
def test_yaml_reader_agrees_with_the_reference_reader_on_every_shipped_parameter_file():
    """Every cvo_params/*.yaml of the reference through cvo_b200_params_read_yaml against an
    independent parse: the set of keys is taken from the reference reader itself
    (fs["key"] in CvoParams.hpp:193-303), a key the reference does not read must leave the default
    untouched, duplicates resolve to the first occurrence (yaml-cpp), comments are stripped."""
    hdr = open(os.path.join(REF, "include", "UnifiedCvo", "cvo", "CvoParams.hpp")).read()
    keys = set(re.findall(r'fs\["([A-Za-z0-9_]+)"\]\.as<', hdr))
    fields = {f[0] for f in _abi.Params._fields_}
    assert len(keys) >= 45 and keys <= fields
    files = sorted(f for f in os.listdir(os.path.join(REF, "cvo_params")) if f.endswith(".yaml"))
    assert len(files) >= 16
    checked = 0
    for name in files:
        path = os.path.join(REF, "cvo_params", name)
        first = {}
        for line in open(path):
            line = line.split("#", 1)[0].strip()
            if ":" not in line:
                continue
            k, v = (t.strip() for t in line.split(":", 1))
            if k and v and k not in first:
                first[k] = v
        p, d = u.read_params_yaml(path), u.default_params()
        for k in fields:
            got = getattr(p, k)
            if k in keys and k in first:
                assert got == pytest.approx(float(first[k]), rel=1e-6), (name, k, got, first[k])
                checked += 1
            else:  # not in the file, or a key the reference never reads (e.g. is_ell_adaptive)
                dv = getattr(d, k)
                assert got == dv or (got != got and dv != dv), (name, k, got, dv)
    assert checked > 500


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "include")), reason="reference tree not present")
def test_defaults_and_field_order_agree_with_the_reference_header():
    """cvo_b200_params against cvo::CvoParams as the reference header declares it: the same fields
    in the same order (the shim reinterpret_casts one to the other) and every default of the
    constructor's initialiser list (CvoParams.hpp:12-126)."""
    hdr = open(os.path.join(REF, "include", "UnifiedCvo", "cvo", "CvoParams.hpp")).read()
    body = hdr[hdr.index("struct CvoParams"):hdr.index("CvoParams() :")]
    body = re.sub(r"//[^\n]*", "", body)
    decl = re.findall(r"\b(float|int|double|unsigned int|bool)\s+([A-Za-z_][A-Za-z0-9_]*)\s*;", body)
    ours = [(f[0], f[1]) for f in _abi.Params._fields_]
    assert [n for _, n in decl] == [n for n, _ in ours]
    ctype = {"float": C.c_float, "int": C.c_int, "double": C.c_double, "unsigned int": C.c_uint, "bool": C.c_bool}
    for (t, n), (_, ct) in zip(decl, ours):
        assert C.sizeof(ctype[t]) == C.sizeof(ct), (n, t, ct)
    init = hdr[hdr.index("CvoParams() :"):hdr.index("{}", hdr.index("CvoParams() :"))]
    defaults = dict(re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\(([-+0-9.eE]+)\)", init))
    assert len(defaults) >= 45
    p = u.default_params()
    for name, val in defaults.items():
        assert getattr(p, name) == pytest.approx(float(val), rel=1e-6), name


def test_oracle_owns_identical_abi_mirrors_and_does_not_import_the_product():
    """oracle/abi_types.py and unified_cvo_b200/_abi.py describe the same three structs field for
    field (the oracle takes the product's structs through void pointers), and importing the
    oracle pulls in nothing of the product (bench.py's reference arm must not map libcvo_b200.so)."""
    import subprocess
    import sys

    from oracle import abi_types as o

    for name in ("Params", "IterTrace", "AlignInfo"):
        a, b = getattr(o, name), getattr(_abi, name)
        assert [(n, t) if not hasattr(t, "_length_") else (n, t._type_, t._length_) for n, t in a._fields_] == \
               [(n, t) if not hasattr(t, "_length_") else (n, t._type_, t._length_) for n, t in b._fields_], name
        assert C.sizeof(a) == C.sizeof(b)
    code = ("import sys, oracle; oracle.lib(); "
            "assert not [m for m in sys.modules if m.startswith('unified_cvo_b200')], 'product imported'; "
            "assert 'libcvo_b200' not in open('/proc/self/maps').read(), 'product library mapped'; print('clean')")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0 and "clean" in out.stdout, out.stderr[-2000:]


def test_oracle_yaml_reader_equals_the_products_reader(data_dir):
    """oracle/params.py (pure Python, used by the reference arm) against cvo_b200_params_read_yaml
    on every yaml shipped in tests/data, field by field; and the two sets of constructor defaults."""
    import glob

    import oracle
    import unified_cvo_b200 as u

    files = sorted(glob.glob(os.path.join(data_dir, "*.yaml")))
    assert len(files) >= 5
    d0, d1 = oracle.default_params(), u.default_params()
    for n, _ in _abi.Params._fields_:
        assert getattr(d0, n) == getattr(d1, n), n
    for f in files:
        a, b = oracle.read_params_yaml(f), u.read_params_yaml(f)
        for n, _ in _abi.Params._fields_:
            assert getattr(a, n) == getattr(b, n), (os.path.basename(f), n)
