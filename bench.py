#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native CVO hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Metric (BASELINE.json): point-pairs/s per CVO iteration = N_src * M_tgt * iterations / time (all
pairs counted, tested or skipped).  A "step" is one full registration
(CvoGPU::align) of one synthetic frame pair:
  N=1  -> workload C2  (BASELINE configs[1]: N=M=10 000, geometric kernel)
  N>1  -> workload C4  (BASELINE configs[3]: N=M=200 000, geometry + 5-dim colour, first-frame
          parameters, MAX_ITER capped at 50), SOURCE rows sharded across ranks, two 72/32-byte
          NCCL all-gathers per iteration (strong scaling: the job is fixed, per-GPU work shrinks).
`value`  times the loop with the clouds already resident in HBM (CUDA events on the launching
         stream, inside libcvo_b200), L2 flushed between steps.
`e2e`    times the reference-facing call with HOST buffers (cvo_b200_align_host: upload + device-side
         build of the cloud, loop, pose read-back) by wall clock.
`roofline`  the dominant kernel, timed live: in cell-query mode on one GPU the whole loop is ONE
         launch of align_grid_kernel (its duration = the timed region, one launch = all the
         iterations); otherwise the dominant per-phase kernel (pair_kernel) at the initial state.
         `traffic` = dram bytes per launch from the committed ncu capture (profiles/traffic.json).
`fp32`   pair tests/s against the measured packed-FMA peak (the meaningful bound of a dense scan).
`frame_pairs`  frame-pairs/s on KITTI-05-sized clouds: a tracking frame (regular parameters,
         constant-velocity initial guess) and a first frame (ell_init = 1.5).
`edge_updates`  pose-graph edge updates/s (multi-frame IRLS edge loop) on four KITTI-05-sized frames.
`cpu_baseline` / `--impl reference`  the CPU restatement of the reference's algorithm (oracle/,
         OpenMP on all host cores) on a bounded number of leading iterations of the same job.
"""
import argparse
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


def load_workload(name):
    import unified_cvo_b200 as u
    from unified_cvo_b200 import synthetic

    cfg, yaml, over, desc = WORKLOADS[name]
    d = synthetic.make_config(cfg)
    p = u.read_params_yaml(os.path.join(DATA, yaml))
    for k, v in over.items():
        if k == "TRACK":
            continue
        if k == "FIRST_FRAME":
            p.ell_init = p.ell_init_first_frame
            p.ell_decay_rate = p.ell_decay_rate_first_frame
            p.ell_decay_start = p.ell_decay_start_first_frame
        else:
            setattr(p, k, v)

    def cloud(c):
        return u.CvoPointCloud(c["xyz"], c["features"], c["labels"], c["geotype"])

    return cloud(d["source"]), cloud(d["target"]), p, desc


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) > 8 for n, v in zip(names, r[5:9]) if v == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": reasons}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full
    captures (profiles/traffic.json, written by tools/ncu_traffic.py): {"workload:kernel": bytes}."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh)
    except Exception:
        return {}


def algorithmic_bytes_per_iteration(N, M, F, C):
    """compulsory traffic of one iteration = (N+M)(16+4F+4C) + 256 bytes."""
    return (N + M) * (16 + 4 * F + 4 * C) + 256


def run_reference(args, rank, world):
    """CPU arm: the oracle's align (restated reference algorithm, OpenMP) on a bounded sample."""
    if rank != 0:
        return
    import oracle

    oracle.use_all_host_threads()  # torchrun exports OMP_NUM_THREADS=1 to its workers
    name = args.workload or ("C2" if world == 1 else "C4")
    src, tgt, p, desc = load_workload(name)
    N, M = src.num_points(), tgt.num_points()
    # bounded sample: a fixed number of leading iterations of the same registration
    per_iter_pairs = N * M
    # the oracle enumerates candidates through a uniform grid (bit-identical to its dense loop,
    # oracle/cvo_oracle.c), like the reference's CPU path searches a kd-tree: ~10 ms per C2 iteration
    sample_iters = max(2, int(min(p.MAX_ITER, 4e10 // per_iter_pairs, 400)))
    p = p.copy()
    p.MAX_ITER = sample_iters
    cs = oracle.Cloud(src.positions_, src.features_, src.labels_, src.geometric_types_)
    ct = oracle.Cloud(tgt.positions_, tgt.features_, tgt.labels_, tgt.geometric_types_)
    for _ in range(min(args.warmup, 1)):
        q = p.copy()
        q.MAX_ITER = 2
        oracle.align(q, cs, ct)
    times, pairs = [], 0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        _, _, info, _ = oracle.align(p, cs, ct)
        times.append(time.perf_counter() - t0)
        pairs += info.pairs_tested
    total = sum(times)
    value = pairs / total
    cores = oracle.num_threads()
    sample = f"first {sample_iters} iterations of the {name} registration per step, {args.steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{name}: {desc}", "N": N, "M": M},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "oracle/cvo_oracle.c with grid-accelerated candidate enumeration "
                                 "(bit-identical to its dense N x M loop), OpenMP"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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
    launches0 = g.launch_count()
    t0 = time.perf_counter()
    total = 0
    for _ in range(rounds):
        total, _ = u.update_edges(states)
    dt = time.perf_counter() - t0
    n_edges = rounds * len(states)
    out = {"workload": "4 KITTI-05-sized frames (N=16384, 5-dim colour), ring of 4 edges, ell=0.25, "
                       f"cap={cap}; one update = posed cloud build(s) on the device (a frame shared with the previous "
                       "edge is reused) + capped kernel matrix + CSR to host",
           "edge_updates_per_s": n_edges / dt, "ms_per_edge_update": 1e3 * dt / n_edges,
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


def run_ours(args, rank, world, local_rank):
    import torch
    import unified_cvo_b200 as u

    name = args.workload or ("C2" if world == 1 else "C4")
    src, tgt, p, desc = load_workload(name)
    T_init = tracking_init() if WORKLOADS[name][2].get("TRACK") else None
    N, M, F, C = src.num_points(), tgt.num_points(), src.feature_dimensions(), src.num_classes()
    torch.cuda.set_device(local_rank)
    g = u.CvoGPU(p, device=local_rank)

    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=world)  # control plane only
        uid = [u.CvoGPU.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        g.comm_init(rank, world, uid[0])
        from unified_cvo_b200.dist import shard_rows
        g.set_row_range(*shard_rows(N, world, rank))
        if os.environ.get("CVO_B200_FUSED", "1") != "0":  # NVLink mailboxes for the persistent kernel
            handles = [None] * world
            dist.all_gather_object(handles, g.comm_mailbox_handle())
            g.comm_open_peers(handles)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    g.set_cloud(0, src)
    g.set_cloud(1, tgt)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")  # > 126 MB L2

    # ---- warm-up (also ramps the clocks: at least ~0.5 s of work)
    t_warm = time.perf_counter()
    w = 0
    while w < args.warmup or time.perf_counter() - t_warm < 0.5:
        g.align(src, tgt, T_init, resident=True)
        w += 1
        if w > args.warmup + 50:
            break

    # ---- timed region: K resident steps, device time per step, L2 flushed between steps
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    launches0 = g.launch_count()
    dev_s, pairs, iters, grid_frac = 0.0, 0, 0, 0.0
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        _, _, info = g.align(src, tgt, T_init, resident=True)
        dev_s += info.registration_seconds
        pairs += info.pairs_tested
        iters += info.iterations + (0 if info.stop_reason == 8 else 1)
        grid_frac += info.cell_query_fraction / args.steps
    barrier()
    wall = time.perf_counter() - wall0
    launches = g.launch_count() - launches0
    clocks = sampler.stop()

    # ---- end to end: host buffers in, pose out, through the reference-facing call
    e2e_s, e2e_pairs = 0.0, 0
    if world == 1:
        for _ in range(max(1, min(args.steps, 5))):
            t0 = time.perf_counter()
            _, _, info = g.align_host(src, tgt, T_init)
            e2e_s += time.perf_counter() - t0
            e2e_pairs += info.pairs_tested
        e2e_steps = max(1, min(args.steps, 5))
    else:
        e2e_steps = max(1, min(args.steps, 3))
        for _ in range(e2e_steps):
            barrier()
            t0 = time.perf_counter()
            g.set_cloud(0, src)
            g.set_cloud(1, tgt)
            _, _, info = g.align(src, tgt, T_init, resident=True)
            e2e_s += time.perf_counter() - t0
            e2e_pairs += info.pairs_tested

    # max over ranks of the device time
    if dist is not None:
        t = torch.tensor([dev_s, e2e_s], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s, e2e_s = float(t[0]), float(t[1])

    # ---- roofline of the dominant kernel, measured live with CUDA events on the handle's stream
    peaks, peak_src = measured_peaks()
    if world == 1:
        rows_local = N
    else:
        from unified_cvo_b200.dist import shard_rows
        rb, re_ = shard_rows(N, world, rank)
        rows_local = re_ - rb
    alg_bytes_iter = algorithmic_bytes_per_iteration(rows_local, M, F, C)
    traffic = load_traffic()
    fused = world > 1 and os.environ.get("CVO_B200_FUSED", "1") != "0"
    if grid_frac >= 0.5 and (world == 1 or fused) and os.environ.get("CVO_B200_PERSIST", "1") != "0":
        # cell-query mode (one GPU, or sharded with the NVLink mailbox exchange): the whole loop is
        # ONE launch of align_grid_kernel per GPU, so the
        # kernel's duration is the timed region itself (registration_seconds = CUDA events around
        # that launch) and one launch processes `iterations` iterations
        kernel = "align_grid_kernel"
        t_kernel = dev_s / args.steps
        launch_units = iters / args.steps
        share = 1.0
    else:
        # one launch per phase: time the dominant kernel of an iteration at the initial state
        kernel = "pair_kernel" if grid_frac < 0.5 else "flow_kernel_t<true>"
        ms_tot, ms_k = g.time_iterations(np.eye(3), np.zeros(3), p.ell_init, p.nearest_neighbors_max, 40)
        t_kernel = ms_k / 40 * 1e-3
        launch_units = 1.0
        share = ms_k / ms_tot
    achieved = launch_units * alg_bytes_iter / t_kernel / 1e9
    roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"], "traffic": traffic.get(f"{name}:{kernel}"),
            "peak_source": peak_src, "kernel": kernel, "kernel_us": t_kernel * 1e6,
            "iterations_per_launch": launch_units, "algorithmic_bytes_per_iteration": alg_bytes_iter,
            "kernel_share_of_step": share,
            "note": "algorithmic bytes = (N+M)(16+4F+4C)+256 per iteration (SURVEY.md 8d): the clouds "
                    "are L2-resident and the path is latency/fp32 bound, not HBM bound; see fp32"}
    fma = g.fma_peak(1, 8192)
    pair_rate = rows_local * M * launch_units / t_kernel
    fp32 = {"pair_tests_per_s": pair_rate, "fma_peak_per_s": fma, "fma_per_pair": 3,
            "peak_pair_tests_per_s": fma / 3.0, "frac": pair_rate / (fma / 3.0),
            "note": "N*M pairs per iteration / kernel time against a dense scan at the measured "
                    "packed-FMA peak (3 FMA per pair); cell queries skip most pairs, so this "
                    "dense-equivalent fraction can exceed 1"}

    if rank != 0:
        return
    value = pairs / dev_s
    # what cvo_b200_align_host copies: the caller's arrays as they are (xyz 12 B, features 4F,
    # labels 4C, geometric type 8 B per point when present), the parameters and the 9 KB state
    geo = 8 if src.geometric_types_ is not None else 0
    h2d = (N + M) * (12 + 4 * F + 4 * C + geo) + 512 + 9000
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{name}: {desc}", "N": N, "M": M, "F": F, "C": C,
                   "iterations_per_step": iters / args.steps, "l2_flush_between_steps": True,
                   "cell_query_fraction": grid_frac,
                   "parallelism": "single GPU" if world == 1 else
                   f"source rows sharded x{world}; dense-scan batches: NCCL all-gather, cell-query "
                   f"batches: persistent kernel with NVLink mailbox exchange",
                   "timing": "CUDA events on the launching stream inside cvo_b200_align, summed over steps"},
        "e2e": {"value": e2e_pairs / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 8800, "steps": e2e_steps,
                "note": "wall clock around cvo_b200_align_host: upload, loop, pose read-back"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "fp32": fp32,
        "wall_ms_per_step": 1e3 * wall / args.steps,
        "frame_pairs_per_s": args.steps / dev_s,
    }
    # frame-pairs/s on KITTI-05-sized clouds (north_star): a tracking frame and a first frame
    if world == 1 and args.workload is None and not args.no_frames:
        line["frame_pairs"] = [frame_pairs_leg(u, wl) for wl in ("KITTI05_TRACK", "KITTI05")]
        line["edge_updates"] = edge_updates_leg(u, cpu=not args.no_cpu_baseline)
    # CPU baseline beside it (rank 0, N=1 only): bounded sample of the same registration
    if world == 1 and not args.no_cpu_baseline:
        import oracle
        oracle.use_all_host_threads()
        q = p.copy()
        q.MAX_ITER = max(2, int(min(p.MAX_ITER, 2.5e11 // (N * M), 4000)))
        cs = oracle.Cloud(src.positions_, src.features_, src.labels_, src.geometric_types_)
        ct = oracle.Cloud(tgt.positions_, tgt.features_, tgt.labels_, tgt.geometric_types_)
        t0 = time.perf_counter()
        _, _, info, _ = oracle.align(q, cs, ct)
        dt = time.perf_counter() - t0
        executed = info.iterations + (0 if info.stop_reason == 8 else 1)
        # the literal dense N x M loop of the same oracle on a short sample, for context
        oracle.set_accel(False)
        qd = p.copy()
        qd.MAX_ITER = max(2, int(min(p.MAX_ITER, 4e9 // (N * M), 40)))
        t0 = time.perf_counter()
        _, _, info_d, _ = oracle.align(qd, cs, ct)
        dt_d = time.perf_counter() - t0
        oracle.set_accel(True)
        line["cpu_baseline"] = {"value": info.pairs_tested / dt, "unit": UNIT, "cores": oracle.num_threads(),
                                "kind": "port",
                                "sample": f"the same {name} registration, first {q.MAX_ITER} iterations at most "
                                          f"({executed} executed; oracle/cvo_oracle.c with grid-accelerated "
                                          f"candidate enumeration, OpenMP), {dt:.1f} s",
                                "dense_loop_value": info_d.pairs_tested / dt_d,
                                "dense_loop_sample": f"first {qd.MAX_ITER} iterations, literal N x M loop, {dt_d:.1f} s"}
    print(json.dumps(line), flush=True)
    g.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-frames", action="store_true", help="skip the KITTI-05-sized frame-pairs/s legs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29531")
    if world != args.gpus and world == 1 and args.gpus > 1:
        print(json.dumps({"error": f"--gpus {args.gpus} needs torchrun (WORLD_SIZE={world})"}))
        sys.exit(2)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
