#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native CVO hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Metric (BASELINE.json): point-pairs/s per CVO iteration = N_src * M_tgt * iterations / time (all
pairs counted, tested or skipped).  A "step" is one full registration
(CvoGPU::align) of one synthetic frame pair.

Workloads (documented choice):
  N=1  -> C2  (BASELINE configs[1], the configuration `metric` is quoted on: N=M=10 000, geometric
          kernel) is the headline line; the SAME line carries `scale_anchor` = C4 measured on this one
          GPU (value, ms_per_step), so that the scaling curve has a same-workload 1-GPU point.
  N>1  -> C4  (BASELINE configs[3]: N=M=200 000, geometry + 5-dim colour, first-frame parameters,
          MAX_ITER capped at 50), SOURCE rows sharded across ranks (strong scaling: the job is
          fixed).  Efficiency at N = value_N / (N * scale_anchor.value of the N=1 line).
          `parity_vs_single`: the first iterations of the sharded run against an unsharded run of
          the same job on rank 0's GPU, and the final pose bit-identical on every rank.
`value`  times the loop with the clouds already resident in HBM (CUDA events on the launching
         stream, inside libcvo_b200), L2 flushed between steps.
`e2e`    times the reference-facing call with HOST buffers (cvo_b200_align_host: upload + device-side
         build of the cloud, loop, pose read-back) by wall clock.
`roofline`  the dominant kernel, timed live with CUDA events on the handle's stream: when the whole
         loop is ONE launch of the persistent kernel its duration is the timed region itself;
         otherwise the dominant per-phase kernel (tile_kernel / pair_kernel) at the initial state.
         bound "hbm" because the contract asks for it: algorithmic bytes = (N+M)(16+4F+4C)+256 per
         iteration — the clouds are L2-resident and the path is latency / fp32-issue
         bound, so `frac` is tiny by construction; `traffic` = dram bytes per launch from the
         committed ncu capture (profiles/traffic.json); null in N>1 lines (captures are 1-GPU runs).
`pipes`  what actually bounds the dominant kernel: FMA-pipe and issue utilisation from the committed
         ncu --set full captures (profiles/pipes.json), not a derived "fraction of peak".
`frame_pairs`  frame-pairs/s on KITTI-05-sized clouds: a tracking frame and a first frame.
`edge_updates`  pose-graph edge updates/s (multi-frame IRLS edge loop); `graph16` = a 16-edge graph whose edges
         are sharded over the ranks (same graph at every N: edge-parallel strong scaling).
`cpu_baseline` / `--impl reference`  the restated reference CPU path cvo::cvo::align (oracle/
         cvo_cpu_baseline.c = Cvo.cpp:885-1089: kd-tree rebuilt per iteration, no row cap, no
         normalisation; OpenMP on all host cores), whole registrations or a bounded number of leading
         iterations; beside it the oracle's port of the GPU-path semantics.  The reference arm never
         imports the product (`product_library_mapped` is checked and reported).
"""
import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "point_pairs_per_s"
UNIT = "pairs/s"
DATA = os.path.join(ROOT, "tests", "data")

WORKLOADS = {
    # name: (synthetic config, yaml, overrides, description)
    "C2": ("C2", "cvo_outdoor_params.yaml",
           dict(is_using_intensity=0, is_using_geometric_type=0, ell_init=0.95),
           "synthetic N=M=10000 geometric kernel (BASELINE configs[1]); cvo_outdoor_params.yaml "
           "with intensity/geometric-type off, ell_init=0.95; one step = one full align()"),
    # FIRST_FRAME = what main_cvo_gpu_align_raw_image.cpp:43-45 does before the first pair of a
    # sequence: ell_init / ell_decay_rate / ell_decay_start <- their *_first_frame values
    "KITTI05": ("KITTI05", "cvo_intensity_params_img_gpu0.yaml", dict(FIRST_FRAME=1),
                "synthetic KITTI-05-sized N=M=16384, geometry+5-dim colour, first-frame parameters "
                "(ell_init=1.5); one step = one align()"),
    # a frame of a running sequence: the yaml's regular parameters (ell_init = 0.15) and the
    # constant-velocity initial guess
    "KITTI05_TRACK": ("KITTI05", "cvo_intensity_params_img_gpu0.yaml", dict(TRACK=1),
                      "synthetic KITTI-05-sized N=M=16384, geometry+5-dim colour, regular (tracking) "
                      "parameters, constant-velocity initial guess with 5 % error; one step = one align()"),
    "C4": ("C4", "cvo_intensity_params_img_gpu0.yaml", dict(FIRST_FRAME=1, MAX_ITER=50),
           "synthetic N=M=200000 geometry+5-dim colour (BASELINE configs[3]), first-frame "
           "parameters (ell_init=1.5), MAX_ITER=50; one step = one align() of 50 iterations"),
}


def tracking_init():
    """Initial guess of a tracking frame: the constant-velocity prediction the sequence drivers
    feed to align (main_cvo_gpu_align_raw_image.cpp:158-160), modelled as the true inter-frame
    motion with a 5 % error.  Returned as T_target_to_source (= inverse of the predicted result)."""
    a = np.deg2rad(2.0 * 0.95)
    R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    Tr = np.eye(4)
    Tr[:3, 3] = np.array([0.05, 0.02, 0.50]) * 0.95
    Rm = np.eye(4)
    Rm[:3, :3] = R
    return np.linalg.inv(Rm @ Tr).astype(np.float32)


def _synthetic_module(with_product):
    """unified_cvo_b200/synthetic.py is pure numpy; the reference arm loads it BY PATH so that
    nothing of the product package is imported there."""
    if with_product:
        from unified_cvo_b200 import synthetic
        return synthetic
    spec = importlib.util.spec_from_file_location("cvo_synthetic_standalone",
                                                  os.path.join(ROOT, "unified_cvo_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def apply_overrides(p, over):
    for k, v in over.items():
        if k == "TRACK":
            continue
        if k == "FIRST_FRAME":
            p.ell_init = p.ell_init_first_frame
            p.ell_decay_rate = p.ell_decay_rate_first_frame
            p.ell_decay_start = p.ell_decay_start_first_frame
        else:
            setattr(p, k, v)
    return p


def load_workload(name):
    """(source, target, params, description) through the product's host mirror."""
    import unified_cvo_b200 as u

    cfg, yaml, over, desc = WORKLOADS[name]
    d = _synthetic_module(True).make_config(cfg)
    p = apply_overrides(u.read_params_yaml(os.path.join(DATA, yaml)), over)

    def cloud(c):
        return u.CvoPointCloud(c["xyz"], c["features"], c["labels"], c["geotype"])

    return cloud(d["source"]), cloud(d["target"]), p, desc


def load_workload_oracle(name):
    """The same workload for the CPU arm: oracle clouds, oracle parameter reader, no product."""
    import oracle

    cfg, yaml, over, desc = WORKLOADS[name]
    d = _synthetic_module(False).make_config(cfg)
    p = apply_overrides(oracle.read_params_yaml(os.path.join(DATA, yaml)), over)

    def cloud(c):
        return oracle.Cloud(c["xyz"], c["features"], c["labels"], c["geotype"])

    return cloud(d["source"]), cloud(d["target"]), p, desc


def copy_params(p):
    import ctypes as C
    q = type(p)()
    C.memmove(C.byref(q), C.byref(p), C.sizeof(q))
    return q


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu, self.t_mark = [], None, gpu_index, None

    def mark_timed_region(self):
        """Samples from here on lie inside the timed region (earlier ones: the warm-up steps)."""
        self.t_mark = time.perf_counter()

    def start(self):
        # NVML in-process (no start-up delay: nvidia-smi needs ~1 s on an 8-GPU box, longer than a
        # short run); nvidia-smi -lms as the fallback.  Same counters as the recipe's clocks line.
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
                self.nv = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.nv = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.nvml, self.stop_flag = pynvml, False
            self.nv_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.nv, pynvml.NVML_CLOCK_SM))
            self.proc = "nvml"
            self.t = threading.Thread(target=self._poll_nvml, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        nv, h = self.nvml, self.nv
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = int(get_reasons(h))
                # same row layout as the nvidia-smi reader: [idx, sm, max, power, active, 4 reasons..., t]
                self.rows.append([str(self.gpu), f"{sm:.0f}", f"{self.nv_max:.0f}", "", hex(r)] +
                                 ["Active" if r & b else "Not Active" for _, b in bits] + [time.perf_counter()])
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")] + [time.perf_counter()])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if self.proc == "nvml":
            self.stop_flag = True
            self.t.join(timeout=1)
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        rows = [r for r in self.rows if len(r) > 9 and r[1].replace(".", "").isdigit()]
        timed = [r for r in rows if self.t_mark is not None and r[-1] >= self.t_mark]
        # a timed region of a few tens of ms holds few 20 ms samples: then the GPU-busy warm-up steps
        # right before it (same kernels, back to back) are reported with it, and the line says so
        window = "timed region"
        if len(timed) < 3:
            timed, window = rows, "warm-up steps + timed region (timed region shorter than 3 samples)"
        sm = [float(r[1]) for r in timed]
        mx = [float(r[2]) for r in timed if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in timed for n, v in zip(names, r[5:9]) if v == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": reasons, "window": window,
                "source": "NVML, 5 ms" if self.proc == "nvml" else "nvidia-smi -lms 20"}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "measured (MEASURED_PEAKS.json, burst copy bandwidth)"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def load_json(name):
    try:
        with open(os.path.join(ROOT, "profiles", name)) as fh:
            return json.load(fh)
    except Exception:
        return {}


def algorithmic_bytes_per_iteration(N, M, F, C):
    """compulsory traffic of one iteration = (N+M)(16+4F+4C) + 256 bytes."""
    return (N + M) * (16 + 4 * F + 4 * C) + 256


# ------------------------------------------------------------------------------- the CPU arm
def cpu_port_cvo_cpp(name, max_iter=None, threads=None):
    """The restated reference CPU path (oracle/cvo_cpu_baseline.c = cvo::cvo::align) on `name`."""
    import oracle

    src, tgt, p, desc = load_workload_oracle(name)
    if max_iter is not None:
        p = copy_params(p)
        p.MAX_ITER = int(max_iter)
    T_init = tracking_init() if WORKLOADS[name][2].get("TRACK") else None
    ret, T, info = oracle.cpu_baseline_align(p, src, tgt, T_init, threads=threads)
    return {"pairs": int(info["pairs"]), "seconds": float(info["seconds"]), "iterations": int(info["executed"]),
            "threads": int(info["threads"]), "ret": int(ret), "N": src.n, "M": tgt.n,
            "split_s": {k: round(float(info[k]), 4) for k in ("t_transform", "t_kdtree_build", "t_se_kernel",
                                                              "t_flow", "t_step")}}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    """CPU arm: the restated reference CPU path on the arm's own workload, all host cores."""
    if rank != 0:
        return
    import oracle

    n_thr = host_threads()  # torchrun exports OMP_NUM_THREADS=1 to its workers
    name = args.workload or ("C2" if world == 1 else "C4")
    N, M = (WORKLOADS[name][0] and _synthetic_module(False).CONFIGS[WORKLOADS[name][0]][1:3])
    # bounded sample: whole registrations where they take seconds (C2: ~450 iterations, ~3 s on 8
    # cores), a few leading iterations where one iteration takes seconds (C4)
    max_iter = None if N * M <= 4e8 else max(2, int(2e11 // (N * M)))
    for _ in range(min(args.warmup, 1)):
        cpu_port_cvo_cpp(name, max_iter=2, threads=n_thr)
    runs = [cpu_port_cvo_cpp(name, max_iter=max_iter, threads=n_thr) for _ in range(max(1, min(args.steps, 3)))]
    pairs = sum(r["pairs"] for r in runs)
    total = sum(r["seconds"] for r in runs)
    value = pairs / total
    sample = (f"{len(runs)} x the {name} registration by the restated cvo::cvo::align "
              f"({'whole registration' if max_iter is None else f'first {max_iter} iterations'}, "
              f"{runs[0]['iterations']} iterations each, {total / len(runs):.2f} s each)")
    mapped = "libcvo_b200" in open("/proc/self/maps").read()
    imported = any(m.startswith("unified_cvo_b200") for m in sys.modules)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": len(runs), "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * total / len(runs),
        "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{name}: {WORKLOADS[name][3]}", "N": int(N), "M": int(M),
                   "iterations_per_step": runs[0]["iterations"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": runs[0]["threads"], "kind": "port",
                         "port_of": "src/cvo/Cvo.cpp:885-1089 (cvo::cvo::align), oracle/cvo_cpu_baseline.c",
                         "sample": sample, "split_s_last_run": runs[-1]["split_s"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "product_library_mapped": bool(mapped), "product_package_imported": bool(imported),
    }
    assert not mapped and not imported, "the reference arm must not load the product"
    assert oracle is not None
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- our arm
def frame_pairs_leg(u, name, steps=5):
    """frame-pairs/s = 1 / (device time of one full align) on a KITTI-05-sized synthetic pair."""
    src, tgt, p, desc = load_workload(name)
    T_init = tracking_init() if WORKLOADS[name][2].get("TRACK") else None
    g = u.CvoGPU(p)
    g.set_cloud(0, src)
    g.set_cloud(1, tgt)
    for _ in range(2):
        g.align(src, tgt, T_init, resident=True)
    dev, iters, wall0 = 0.0, 0, time.perf_counter()
    for _ in range(steps):
        ret, T, info = g.align(src, tgt, T_init, resident=True)
        dev += info.registration_seconds
        iters += info.iterations + (0 if info.stop_reason == 8 else 1)
    wall = time.perf_counter() - wall0
    t0 = time.perf_counter()
    g.align_host(src, tgt, T_init)
    e2e = time.perf_counter() - t0
    from unified_cvo_b200 import synthetic
    err = float(np.abs(T - synthetic.gt_transform()).max())
    g.close()
    return {"workload": f"{name}: {desc}", "frame_pairs_per_s": steps / dev, "ms_per_frame_pair": 1e3 * dev / steps,
            "iterations_per_frame_pair": iters / steps, "e2e_frame_pairs_per_s": 1.0 / e2e,
            "wall_ms_per_frame_pair": 1e3 * wall / steps, "max_abs_pose_error_vs_truth": err, "ret": int(ret)}


def edge_updates_leg(u, rounds=5, cpu=True):
    """Pose-graph edge updates/s: four KITTI-05-sized frames resident on the
    device, a ring of four edges, every round = one outer IRLS iteration's edge loop
    (BinaryStateGPU::update_inner_product per edge: both frames moved by their poses, capped
    kernel matrix filled and copied to the host).  Wall clock, matrix read-back included."""
    from unified_cvo_b200 import synthetic
    src, tgt, p, _ = load_workload("KITTI05_TRACK")
    g = u.CvoGPU(p)
    I = np.eye(4)[:3]
    G = np.asarray(synthetic.gt_transform(), np.float64)[:3]  # maps target points into the source frame
    frames = [u.CvoFrameGPU(g, c, P) for c, P in ((src, I), (tgt, G), (src, I), (tgt, G))]
    ell, cap = 0.25, int(p.multiframe_num_neighbors)
    states = [u.BinaryStateGPU(frames[i], frames[(i + 1) % 4], cap, ell) for i in range(4)]
    u.update_edges(states)  # warm-up: buffers grow, the caps settle
    u.update_edges(states)
    per_call = {}
    for batched in (False, True):  # the per-edge C-ABI call, then the whole loop as one batch call (the headline)
        u.update_edges(states, batched=batched)
        launches0 = g.launch_count()
        t0 = time.perf_counter()
        total = 0
        for _ in range(rounds):
            total, _ = u.update_edges(states, batched=batched)
        dt = time.perf_counter() - t0
        per_call[batched] = rounds * len(states) / dt
    n_edges = rounds * len(states)
    out = {"workload": "4 KITTI-05-sized frames (N=16384, 5-dim colour), ring of 4 edges, ell=0.25, "
                       f"cap={cap}; one update = posed cloud build(s) on the device (a frame shared with the previous "
                       "edge is reused) + capped kernel matrix + CSR to host",
           "edge_updates_per_s": n_edges / dt, "ms_per_edge_update": 1e3 * dt / n_edges,
           "api": "cvo_b200_edge_update_batch (one call per round, one host wait)",
           "edge_updates_per_s_per_edge_calls": per_call[False],
           "nonzeros_per_round": int(total), "gpu_launches_per_edge_update": (g.launch_count() - launches0) / n_edges}
    if cpu:  # the CPU restatement of the same edge loop on the host cores (one round)
        import oracle
        oracle.use_all_host_threads()
        oc = {id(f): oracle.Cloud(f.points.positions_, f.points.features_, f.points.labels_,
                                  f.points.geometric_types_) for f in frames}
        t0 = time.perf_counter()
        cpu_total = 0
        for st in states:
            t, _ = oracle.edge_update(p, oc[id(st.frame1)], st.frame1.pose_float(), oc[id(st.frame2)],
                                      st.frame2.pose_float(), st.ell_, st.num_neighbors_)
            cpu_total += t
        dt_cpu = time.perf_counter() - t0
        out["cpu_baseline"] = {"edge_updates_per_s": len(states) / dt_cpu, "cores": oracle.num_threads(), "kind": "port",
                               "nonzeros_per_round": int(cpu_total),
                               "sample": "one round of the same four edges, oracle/cvo_oracle.c (grid-accelerated, OpenMP)"}
    g.close()
    return out


def edge_graph_leg(u, local_rank, rank, world, dist, rounds=8):
    """Edge-parallel pose graph: 8 KITTI-05-sized frames, 16 edges (ring + every
    second-neighbour chord), the same graph at every N.  Rank r owns edges r, r + world, ... and
    refills them with one batch call per round on its GPU; the other ranks' CSR matrices travel to
    rank 0 (where a solver would run) over the host control plane inside the timed region.  Wall
    clock between barriers, max over ranks; strong scaling."""
    from unified_cvo_b200 import synthetic
    src, tgt, p, _ = load_workload("KITTI05_TRACK")
    g = u.CvoGPU(p, device=local_rank)
    I = np.eye(4)[:3]
    G = np.asarray(synthetic.gt_transform(), np.float64)[:3]
    n_frames = 8
    edges = [(i, (i + 1) % n_frames) for i in range(n_frames)] + [(i, (i + 2) % n_frames) for i in range(n_frames)]
    from unified_cvo_b200.dist import shard_edges
    mine = shard_edges(len(edges), world, rank)
    need = sorted({f for k in mine for f in edges[k]})
    frames = {f: u.CvoFrameGPU(g, src if f % 2 == 0 else tgt, I if f % 2 == 0 else G) for f in need}
    cap = int(p.multiframe_num_neighbors)

    class _Remote:  # an edge owned by another rank: only its host-side result slot exists here
        def __init__(self, m):
            self.A_result_cpu_, self.last_max_row_nnz = u.Association(), 0
            self.frame2 = type("F", (), {"points": type("P", (), {"num_points": staticmethod(lambda: m)})()})()

    states = [u.BinaryStateGPU(frames[a], frames[b], cap, 0.25) if k in mine else _Remote(tgt.num_points() if b % 2 else src.num_points())
              for k, (a, b) in enumerate(edges)]

    def barrier():
        if dist is not None:
            dist.barrier()

    for _ in range(2):
        u.update_edges_sharded(states, rank, world, dist)
    barrier()
    t0 = time.perf_counter()
    for _ in range(rounds):
        u.update_edges_sharded(states, rank, world, dist)
    barrier()
    dt = time.perf_counter() - t0
    if dist is not None:
        ts = [None] * world
        dist.all_gather_object(ts, dt)
        dt = max(ts)
    total = int(sum(len(st.A_result_cpu_.vals) for st in states)) if rank == 0 else 0
    # the same rounds with the matrices left on the ranks that computed them (no gather): what the
    # GPUs themselves scale like
    barrier()
    t0 = time.perf_counter()
    for _ in range(rounds):
        u.update_edges_sharded(states, rank, world, dist, gather_to=None)
    barrier()
    dt_local = time.perf_counter() - t0
    if dist is not None:
        ts = [None] * world
        dist.all_gather_object(ts, dt_local)
        dt_local = max(ts)
    g.close()
    return {"workload": "8 KITTI-05-sized frames, 16 edges (ring + second-neighbour chords), ell=0.25, cap=%d; edges sharded "
                        "round-robin over the ranks, one cvo_b200_edge_update_batch per rank and round, CSR matrices gathered "
                        "to rank 0 over the host control plane inside the timed region" % cap,
            "n_gpus": world, "edges": len(edges), "rounds": rounds, "edge_updates_per_s": rounds * len(edges) / dt,
            "ms_per_round": 1e3 * dt / rounds, "nonzeros_per_round_on_rank0": total, "scaling": "strong",
            "edge_updates_per_s_without_gather": rounds * len(edges) / dt_local}


def measure(u, torch, g, name, src, tgt, p, steps, warmup, local_rank, rank, world, dist):
    """Timed region of one workload on the handle g (clouds already set).  Returns a dict."""
    T_init = tracking_init() if WORKLOADS[name][2].get("TRACK") else None
    N, M, F, C = src.num_points(), tgt.num_points(), src.feature_dimensions(), src.num_classes()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")  # > 126 MB L2
    # ---- warm-up (also ramps the clocks: at least ~0.5 s of work)
    # The NUMBER of warm-up steps must be the same on every rank (each align is a collective of the
    # job): it is derived from the slowest rank's first step, never from a rank's own clock.
    t_warm = time.perf_counter()
    g.align(src, tgt, T_init, resident=True)
    first = time.perf_counter() - t_warm
    if dist is not None:
        firsts = [None] * world
        dist.all_gather_object(firsts, first)
        first = max(firsts)
    n_warm = max(warmup, min(warmup + 50, int(0.5 / max(first, 1e-4)) + 1))
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(n_warm - 1):
        g.align(src, tgt, T_init, resident=True)
    # ---- timed region: K resident steps, device time per step, L2 flushed between steps
    barrier()
    sampler.mark_timed_region()
    launches0 = g.launch_count()
    dev_s, pairs, iters, persist_frac = 0.0, 0, 0, 0.0
    wall0 = time.perf_counter()
    last_T = None
    for _ in range(steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        _, last_T, info = g.align(src, tgt, T_init, resident=True)
        dev_s += info.registration_seconds
        pairs += info.pairs_tested
        iters += info.iterations + (0 if info.stop_reason == 8 else 1)
        persist_frac += info.cell_query_fraction / steps
    builds = g.last_candidate_builds()  # of the last step (persistent tile mode; 0 otherwise)
    barrier()
    wall = time.perf_counter() - wall0
    launches = g.launch_count() - launches0
    clocks = sampler.stop()
    # ---- end to end: host buffers in, pose out, through the reference-facing call
    e2e_s, e2e_pairs = 0.0, 0
    e2e_steps = max(1, min(steps, 5 if world == 1 else 3))
    for _ in range(e2e_steps):
        if world == 1:
            t0 = time.perf_counter()
            _, _, info = g.align_host(src, tgt, T_init)
        else:
            barrier()
            t0 = time.perf_counter()
            g.set_cloud(0, src)
            g.set_cloud(1, tgt)
            _, _, info = g.align(src, tgt, T_init, resident=True)
        e2e_s += time.perf_counter() - t0
        e2e_pairs += info.pairs_tested
    if dist is not None:  # max over ranks of the device time
        t = torch.tensor([dev_s, e2e_s], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s, e2e_s = float(t[0]), float(t[1])
    del flush
    return dict(N=N, M=M, F=F, C=C, dev_s=dev_s, pairs=pairs, iters=iters, persist_frac=persist_frac, wall=wall,
                launches=launches, clocks=clocks, builds=builds, e2e_s=e2e_s, e2e_pairs=e2e_pairs, e2e_steps=e2e_steps,
                last_T=last_T, T_init=T_init)


def roofline_of(u, g, name, p, m, steps, rows_local, world):
    """Contract object for the dominant kernel + the measured pipe figures."""
    peaks, peak_src = measured_peaks()
    alg_bytes_iter = algorithmic_bytes_per_iteration(rows_local, m["M"], m["F"], m["C"])
    traffic, pipes = load_json("traffic.json"), load_json("pipes.json")
    persistent = m["persist_frac"] >= 0.5 and os.environ.get("CVO_B200_PERSIST", "1") != "0"
    if persistent:
        # the whole loop is ONE launch of the persistent kernel per GPU: its duration is the timed
        # region itself (registration_seconds = CUDA events around that launch) and one launch
        # processes `iterations` iterations
        kernel = "align_grid_kernel"
        t_kernel = m["dev_s"] / steps
        launch_units = m["iters"] / steps
        share = 1.0
    else:
        # one launch per phase: time the candidate generator of an iteration at the initial state
        ms_tot, ms_k = g.time_iterations(np.eye(3), np.zeros(3), p.ell_init, p.nearest_neighbors_max, 20)
        kernel = "tile_kernel|pair_kernel (the candidate generator of the dense regimes)"
        t_kernel = max(ms_k, 1e-6) / 20 * 1e-3
        launch_units = 1.0
        share = ms_k / ms_tot if ms_tot > 0 else None
    achieved = launch_units * alg_bytes_iter / t_kernel / 1e9
    key = f"{name}:{kernel.split('|')[0]}"
    roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"], "traffic": traffic.get(key) if world == 1 else None,
            "peak_source": peak_src, "kernel": kernel, "kernel_us": t_kernel * 1e6,
            "iterations_per_launch": launch_units, "algorithmic_bytes_per_iteration": alg_bytes_iter,
            "kernel_share_of_step": share, "ranks": world,
            "note": "algorithmic bytes = (N_local+M)(16+4F+4C)+256 per iteration (SURVEY.md 8d). The clouds are "
                    "L2-resident; the path is latency / fp32-issue bound, not HBM bound (see pipes)"}
    if world > 1:  # ncu captures are one-GPU runs: there is no per-rank DRAM figure of a sharded job
        roof["traffic_note"] = ("null: no per-rank capture (ncu is never run on a multi-rank command); the 1-GPU launch of "
                                f"the same workload moves {traffic.get(key)} bytes (profiles/traffic.json)")
    pipe = pipes.get(key, {})
    pipe_obj = {"kernel": kernel, "fma_pipe_pct": pipe.get("fma_pipe_pct"), "issue_active_pct": pipe.get("issue_active_pct"),
                "warps_active_pct": pipe.get("warps_active_pct"), "top_stalls": pipe.get("top_stalls"),
                "source": pipe.get("source", "no ncu capture committed for this workload/kernel"),
                "dense_equivalent_pairs_per_s": rows_local * m["M"] * launch_units / t_kernel,
                "note": ("1-GPU capture of the same workload; " if world > 1 else "") +
                        "ncu --set full figures of the committed capture (profiles/), not derived numbers; "
                        "dense_equivalent = N_local*M pairs per iteration / kernel time (culled pairs counted)"}
    return roof, pipe_obj


def parity_vs_single(u, g, name, src, tgt, p, T_init, last_T, rank, world, local_rank, dist, k=4):
    """The first k iterations of the sharded run against an unsharded run of the same job on
    this rank's GPU; the final pose of the sharded run bit-identical on every rank."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import compare_traces
    q = p.copy()
    q.MAX_ITER = k
    g.write_params(q)
    _, _, _, tr_sh = g.align(src, tgt, T_init, trace_cap=k, resident=True)
    g.write_params(p)
    out = {"iterations_compared": k}
    if rank == 0:
        s = u.CvoGPU(q, device=local_rank)
        s.set_cloud(0, src)
        s.set_cloud(1, tgt)
        _, _, _, tr_1 = s.align(src, tgt, T_init, trace_cap=k, resident=True)
        s.close()
        bad = [(i, compare_traces(tr_sh[i], tr_1[i])) for i in range(min(len(tr_sh), len(tr_1)))]
        bad = [(i, b) for i, b in bad if b]
        tw = max(float(np.abs(np.array(list(tr_sh[i].omega) + list(tr_sh[i].v)) -
                              np.array(list(tr_1[i].omega) + list(tr_1[i].v))).max()) for i in range(len(tr_1)))
        out.update({"trace_ok": not bad, "mismatches": [f"iter {i}: {b}" for i, b in bad][:4],
                    "max_abs_twist_diff": tw, "nnz_equal": all(tr_sh[i].nnz == tr_1[i].nnz for i in range(len(tr_1)))})
    poses = [None] * world
    dist.all_gather_object(poses, np.asarray(last_T, np.float32).tobytes())
    out["pose_bit_identical_across_ranks"] = all(b == poses[0] for b in poses)
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    import unified_cvo_b200 as u

    name = args.workload or ("C2" if world == 1 else "C4")
    src, tgt, p, desc = load_workload(name)
    N = src.num_points()
    torch.cuda.set_device(local_rank)
    g = u.CvoGPU(p, device=local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=world)  # control plane only
        from unified_cvo_b200.dist import attach
        attach(g, N, rank, world, dist, fused=os.environ.get("CVO_B200_FUSED", "1") != "0")
    g.set_cloud(0, src)
    g.set_cloud(1, tgt)
    m = measure(u, torch, g, name, src, tgt, p, args.steps, args.warmup, local_rank, rank, world, dist)
    if world == 1:
        rows_local = N
    else:
        from unified_cvo_b200.dist import shard_rows
        rb, re_ = shard_rows(N, world, rank)
        rows_local = re_ - rb
    roof, pipes = roofline_of(u, g, name, p, m, args.steps, rows_local, world)
    par = None
    if world > 1:
        par = parity_vs_single(u, g, name, src, tgt, p, m["T_init"], m["last_T"], rank, world, local_rank, dist)
    graph = None
    if world > 1 and args.workload is None and not args.no_frames:
        graph = edge_graph_leg(u, local_rank, rank, world, dist)  # every rank takes part
    if rank != 0:
        return
    value = m["pairs"] / m["dev_s"]
    # what cvo_b200_align_host copies: the caller's arrays as they are (xyz 12 B, features 4F,
    # labels 4C, geometric type 8 B per point when present), the parameters and the 9 KB state
    geo = 8 if src.geometric_types_ is not None else 0
    h2d = (m["N"] + m["M"]) * (12 + 4 * m["F"] + 4 * m["C"] + geo) + 512 + 9000
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * m["dev_s"] / args.steps, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{name}: {desc}", "N": m["N"], "M": m["M"], "F": m["F"], "C": m["C"],
                   "iterations_per_step": m["iters"] / args.steps, "l2_flush_between_steps": True,
                   "persistent_kernel_fraction": m["persist_frac"],
                   "candidate_cell_builds_per_step": m["builds"],
                   "parallelism": "single GPU" if world == 1 else
                   f"source rows sharded x{world}; the whole loop is one persistent kernel per GPU whose two "
                   f"per-iteration exchanges are NVLink stores into the peers' mailboxes (no NCCL call, no launch "
                   f"per iteration); NCCL all-gathers only for batches that fall back to one launch per phase",
                   "timing": "CUDA events on the launching stream inside cvo_b200_align, summed over steps; "
                             "max over ranks"},
        "e2e": {"value": m["e2e_pairs"] / m["e2e_s"], "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 8800, "steps": m["e2e_steps"],
                "note": "wall clock around cvo_b200_align_host: upload, loop, pose read-back"},
        "gpu_launches": int(m["launches"]),
        "gpu_launches_per_iteration": m["launches"] / max(m["iters"], 1),
        "clocks": m["clocks"],
        "roofline": roof,
        "pipes": pipes,
        "wall_ms_per_step": 1e3 * m["wall"] / args.steps,
        "frame_pairs_per_s": args.steps / m["dev_s"],
    }
    if par is not None:
        line["parity_vs_single"] = par
    if graph is not None:
        line["edge_updates"] = {"graph16": graph}
    # the same-workload anchor of the scaling curve: C4 on THIS one GPU
    if world == 1 and args.workload is None and not args.no_anchor:
        g.close()
        s4, t4, p4, d4 = load_workload("C4")
        g4 = u.CvoGPU(p4, device=local_rank)
        g4.set_cloud(0, s4)
        g4.set_cloud(1, t4)
        m4 = measure(u, torch, g4, "C4", s4, t4, p4, 2, 2, local_rank, 0, 1, None)
        roof4, pipes4 = roofline_of(u, g4, "C4", p4, m4, 2, s4.num_points(), 1)
        line["scale_anchor"] = {
            "workload": f"C4: {d4}", "n_gpus": 1, "value": m4["pairs"] / m4["dev_s"], "unit": UNIT, "steps": 2,
            "ms_per_step": 1e3 * m4["dev_s"] / 2, "iterations_per_step": m4["iters"] / 2,
            "e2e_value": m4["e2e_pairs"] / m4["e2e_s"], "gpu_launches": int(m4["launches"]), "clocks": m4["clocks"],
            "roofline": roof4, "pipes": pipes4,
            "note": "the N>1 lines run this workload sharded: efficiency(N) = value_N / (N * this value)"}
        g4.close()
        del s4, t4
        g = None
    # frame-pairs/s on KITTI-05-sized clouds (north_star): a tracking frame and a first frame
    if world == 1 and args.workload is None and not args.no_frames:
        line["frame_pairs"] = [frame_pairs_leg(u, wl) for wl in ("KITTI05_TRACK", "KITTI05")]
        line["edge_updates"] = edge_updates_leg(u, cpu=not args.no_cpu_baseline)
        line["edge_updates"]["graph16"] = edge_graph_leg(u, local_rank, 0, 1, None)
        try:
            line["demo_registration"] = gpu_demo_registration(u, local_rank)
        except Exception as e:  # never let an auxiliary leg take the line down
            line["demo_registration"] = {"error": str(e)}
    # CPU baselines beside it (rank 0, N=1 only)
    if world == 1 and not args.no_cpu_baseline:
        import oracle
        n_thr = oracle.use_all_host_threads()
        base = cpu_port_cvo_cpp(name, max_iter=None if m["N"] * m["M"] <= 4e8 else max(2, int(2e11 // (m["N"] * m["M"]))),
                                threads=n_thr)
        cs = oracle.Cloud(src.positions_, src.features_, src.labels_, src.geometric_types_)
        ct = oracle.Cloud(tgt.positions_, tgt.features_, tgt.labels_, tgt.geometric_types_)
        q = p.copy()
        q.MAX_ITER = max(2, int(min(p.MAX_ITER, 2.5e11 // (m["N"] * m["M"]), 4000)))
        t0 = time.perf_counter()
        _, _, info, _ = oracle.align(q, cs, ct)
        dt = time.perf_counter() - t0
        executed = info.iterations + (0 if info.stop_reason == 8 else 1)
        line["cpu_baseline"] = {
            "value": base["pairs"] / base["seconds"], "unit": UNIT, "cores": base["threads"], "kind": "port",
            "port_of": "src/cvo/Cvo.cpp:885-1089 (cvo::cvo::align), oracle/cvo_cpu_baseline.c",
            "sample": f"one whole {name} registration by the restated reference CPU path: {base['iterations']} "
                      f"iterations, {base['seconds']:.2f} s (kd-tree rebuilt per iteration, no row cap, no "
                      f"normalisation)",
            "frame_pairs_per_s": 1.0 / base["seconds"], "split_s": base["split_s"],
            "oracle_port": {"value": info.pairs_tested / dt, "unit": UNIT, "cores": oracle.num_threads(),
                            "what": "oracle/cvo_oracle.c: CPU port of the GPU-path semantics (row cap, normalised "
                                    "twist), grid-accelerated candidate enumeration, OpenMP",
                            "sample": f"first {q.MAX_ITER} iterations at most ({executed} executed), {dt:.1f} s"}}
        if args.workload is None:  # BASELINE config 1: the demo pair through the same CPU path
            try:
                demo = cpu_demo_registration(n_thr)
                line["cpu_baseline"]["demo_registration"] = demo
            except Exception as e:  # never let an auxiliary leg take the line down
                line["cpu_baseline"]["demo_registration"] = {"error": str(e)}
    print(json.dumps(line), flush=True)
    if g is not None:
        g.close()


def gpu_demo_registration(u, device):
    """BASELINE configs[0] on the GPU: the README demo pair through CvoGPU::align, both flavours
    (the drivers' ell_init = distance of the cloud means: rows are cut at their cap, so the loop is
    523 latency-bound rows of 1 080 targets each - the worst case for this design)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import demo_clouds, demo_params
    out = {}
    for flavour, color in (("two_color_pcd (colour)", True), ("two_pcd (geometric only)", False)):
        src, tgt = demo_clouds(color=color)
        p = demo_params(src, tgt, color=color)
        g = u.CvoGPU(p, device=device)
        g.align_host(src, tgt)  # warm-up
        t0 = time.perf_counter()
        ret, T, info = g.align_host(src, tgt)
        wall = time.perf_counter() - t0
        it = info.iterations + (0 if info.stop_reason == 8 else 1)
        out[flavour] = {"seconds": float(info.registration_seconds), "wall_seconds_host_buffers": wall, "iterations": int(it),
                        "ret": int(ret), "pairs_per_s": float(info.pairs_tested) / float(info.registration_seconds)}
        g.close()
    return out


def cpu_demo_registration(n_thr):
    """BASELINE configs[0]: demo_data/source.pcd vs target.pcd, cvo_outdoor_params.yaml, the
    reference CPU align() (restated) on the host cores - whole-registration time."""
    import oracle
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import demo_clouds, demo_params, to_oracle_cloud
    out = {}
    for flavour, color in (("two_color_pcd (colour)", True), ("two_pcd (geometric only)", False)):
        src, tgt = demo_clouds(color=color)
        p = demo_params(src, tgt, color=color)
        ret, T, info = oracle.cpu_baseline_align(p, to_oracle_cloud(src), to_oracle_cloud(tgt), threads=n_thr)
        out[flavour] = {"seconds": float(info["seconds"]), "iterations": int(info["executed"]), "ret": int(ret),
                        "pairs_per_s": float(info["pairs"]) / float(info["seconds"]), "cores": int(info["threads"])}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-frames", action="store_true", help="skip the KITTI-05-sized frame-pairs/s legs")
    ap.add_argument("--no-anchor", action="store_true", help="skip the C4 scale anchor of the N=1 line")
    args = ap.parse_args()
    if os.environ.get("BENCH_WATCHDOG"):  # debugging aid: dump every thread's stack and exit after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["BENCH_WATCHDOG"]), exit=True)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29531")
    if world != args.gpus and world == 1 and args.gpus > 1:
        print(json.dumps({"error": f"--gpus {args.gpus} needs torchrun (WORLD_SIZE={world})"}))
        sys.exit(2)
    if rank != 0:
        # only rank 0 writes to stdout (the JSON line); whatever libraries print on the other
        # ranks goes to stderr
        os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
