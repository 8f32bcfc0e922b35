"""bench.py's contract objects, as far as they can be checked without a GPU: the roofline / pipes
objects of the product arm (built from mock timings) and the reference arm's JSON line (the
restated cvo::cvo::align on a bounded sample).  The product arm itself needs a B200."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _P:
    ell_init = 1.5
    nearest_neighbors_max = 256


@pytest.mark.parametrize("world", [1, 2, 8])
def test_roofline_object_of_a_persistent_launch(world):
    import bench

    steps, iters_per_step = 5, 50
    n = 200_000
    rows = n // world
    m = dict(N=n, M=n, F=5, C=0, dev_s=0.06, pairs=1, iters=steps * iters_per_step, persist_frac=1.0)
    roof, pipes = bench.roofline_of(None, None, "C4", _P(), m, steps, rows, world)
    # (N_local + M)(16 + 4F + 4C) + 256 bytes per iteration
    per_iter = (rows + n) * (16 + 4 * 5) + 256
    assert roof["algorithmic_bytes_per_iteration"] == per_iter
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s"
    assert roof["iterations_per_launch"] == iters_per_step
    t_launch = 0.06 / steps
    assert roof["achieved"] == pytest.approx(iters_per_step * per_iter / t_launch / 1e9)
    assert roof["frac"] == pytest.approx(roof["achieved"] / roof["peak"])
    assert 0.0 < roof["frac"] < 1.0
    assert roof["ranks"] == world
    if world == 1:  # the committed 1-GPU capture of this workload
        assert roof["traffic"] and roof["traffic"] > per_iter
    else:  # no per-rank capture exists: null, and the line says why
        assert roof["traffic"] is None and "no per-rank capture" in roof["traffic_note"]
    for key in ("fma_pipe_pct", "issue_active_pct", "top_stalls", "source"):
        assert pipes.get(key) is not None
    assert os.path.exists(os.path.join(ROOT, pipes["source"]))
    json.dumps([roof, pipes])  # serialisable as they are


def test_committed_profile_indices_point_at_committed_files():
    import bench

    pipes, traffic = bench.load_json("pipes.json"), bench.load_json("traffic.json")
    assert pipes and traffic
    for key, rec in pipes.items():
        assert os.path.exists(os.path.join(ROOT, rec["source"])), (key, rec["source"])
    for key in ("C2:align_grid_kernel", "C4:align_grid_kernel"):
        assert key in pipes and key in traffic


def test_reference_arm_line_on_a_bounded_sample():
    """`bench.py --impl reference`: one JSON line, the contract keys, nothing of the product mapped."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "point_pairs_per_s" and d["unit"] == "pairs/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["N"] == 10_000 and d["config"]["M"] == 10_000  # BASELINE configs[1]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
    assert d["product_library_mapped"] is False and d["product_package_imported"] is False


def test_reference_arm_under_torchrun_prints_one_line_from_rank_0():
    """N > 1: launched like the product arm; rank 0 alone runs the CPU path (on the N > 1 workload,
    C4), the other ranks exit 0 without work."""
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "bench.py"),
                          "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
    assert d["config"]["N"] == 200_000 and d["config"]["M"] == 200_000  # BASELINE configs[3]
    assert d["cpu_baseline"]["sample"] and d["gpu_launches"] == 0
