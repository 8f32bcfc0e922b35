"""Generates the committed golden fixtures from the ORACLE (oracle/cvo_oracle.c).

The reference ships no golden vectors for this path, so these fixtures
freeze the oracle's outputs (in its default arithmetic = the reference's GPU build, see
oracle/cvo_oracle.c above mul_add(); the oracle's K1-K4 are pinned against the reference's own
kernel text by tests/test_ref_pin*.py): they guard the oracle against regressions and give the
GPU parity tests inputs/outputs that do not need the oracle at run time.  Regenerate with:  python tests/golden/make_golden.py
Each fixture holds teacher-forced single iterations: the state (R, T, ell, cap) taken
from the oracle's own align() trajectory and the iteration's outputs at that state.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import oracle  # noqa: E402
from helpers import (demo_clouds, demo_params, geometric_params, synthetic_pair,  # noqa: E402
                     to_oracle_cloud)

CASES = {
    "demo_color": dict(kind="demo", color=True, sample=[0, 1, 2, 10, 100, 500, 1000, 3000, 6000]),
    "demo_geometric": dict(kind="demo", color=False, sample=[0, 1, 2, 10, 100, 500, 1000, 3000]),
    "synthetic_2k": dict(kind="synthetic", P=2500, N=2000, M=2000, seed=20002,
                         sample=[0, 1, 2, 5, 10, 50, 100, 200, 400]),
}


def case_inputs(c):
    if c["kind"] == "demo":
        src, tgt = demo_clouds(c["color"])
        return src, tgt, demo_params(src, tgt, c["color"])
    src, tgt, _ = synthetic_pair(c["P"], c["N"], c["M"], c["seed"])
    return src, tgt, geometric_params()


def run_case(c):
    src, tgt, p = case_inputs(c)
    cs, ct = to_oracle_cloud(src), to_oracle_cloud(tgt)
    n = max(c["sample"]) + 1
    ret, T, info, tr = oracle.align(p, cs, ct, None, trace_cap=n)
    out = []
    for k in c["sample"]:
        if k >= len(tr):
            continue
        if k == 0:
            R, Tt = np.eye(3, dtype=np.float32).reshape(9), np.zeros(3, np.float32)
        else:
            R, Tt = np.array(list(tr[k - 1].R), np.float32), np.array(list(tr[k - 1].T), np.float32)
        ell, cap = float(tr[k].ell), int(tr[k].num_neighbors)
        r = oracle.iterate(p, cs, ct, R, Tt, ell, cap)
        out.append(dict(k=k, R=[float(x) for x in R], T=[float(x) for x in Tt], ell=ell, cap=cap,
                        nnz=int(r.nnz), max_row_nnz=int(r.max_row_nnz),
                        twist=[float(x) for x in list(r.omega) + list(r.v)],
                        omega_sum=list(r.omega_sum), v_sum=list(r.v_sum),
                        BCDE=[r.B, r.C, r.D, r.E], step=float(r.step), a_sum=r.a_sum, dist=r.dist,
                        R_next=[float(x) for x in r.R], T_next=[float(x) for x in r.T]))
    return dict(iters=out, ret=ret, iterations=info.iterations, transform=[[float(x) for x in row] for row in T])


# pose-graph edge updates (oracle.edge_update): two frames at their own poses, fixed ell and cap
EDGE_CASES = [
    dict(name="geometric", P=1600, N=1000, M=1200, seed=77, F=0, C=0, geotype=False, ell=0.8, cap=24,
         pose1=[0.3, 0.5, -0.2, 0.05, 0.0, -0.1], pose2=[0.1, 2.2, 0.0, 0.1, 0.02, 0.35]),
    dict(name="colour_semantics", P=1200, N=700, M=900, seed=5, F=5, C=20, geotype=True, ell=1.2, cap=40,
         pose1=[0.0, 0.4, 0.0, 0.0, 0.0, 0.05], pose2=[0.0, 2.4, 0.0, 0.05, 0.02, 0.55]),
]


def edge_pose(v):
    """[rx, ry, rz (deg), tx, ty, tz] -> row-major 3x4 float32 (CvoFrame::pose_vec narrowed)."""
    ax, ay, az = np.deg2rad(v[:3])
    Rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
    Ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
    Rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
    P = np.zeros((3, 4))
    P[:, :3] = Rz @ Ry @ Rx
    P[:, 3] = v[3:]
    return P.reshape(12).astype(np.float32)


def edge_inputs(c):
    import unified_cvo_b200 as u
    from helpers import DATA
    f1, f2, _ = synthetic_pair(c["P"], c["N"], c["M"], c["seed"], F=c["F"], C=c["C"], geotype=c["geotype"])
    if c["F"]:
        p = u.read_params_yaml(os.path.join(DATA, "cvo_semantic_params_img_gpu0.yaml"))
        p.is_using_geometric_type = 1
        p.c_ell = 0.5
    else:
        p = geometric_params()
    return f1, f2, p


def run_edge_case(c):
    f1, f2, p = edge_inputs(c)
    total, sp = oracle.edge_update(p, to_oracle_cloud(f1), edge_pose(c["pose1"]), to_oracle_cloud(f2),
                                   edge_pose(c["pose2"]), c["ell"], c["cap"])
    row_ptr, cols, vals = oracle.sparse_to_csr(sp)
    return dict(name=c["name"], nnz=total, max_row_nnz=int(sp["nonzeros"].max()),
                row_ptr=[int(x) for x in row_ptr], cols=[int(x) for x in cols],
                vals=[float(x) for x in vals])


def demo_final_pose_spread(T0):
    """How far the ORACLE's own final pose of the demo registration moves when its initial pose moves
    by 1e-7 .. 3e-7 m (six starts, ~7 000 chaotic iterations each): the bar a GPU run can be held to."""
    import oracle
    from helpers import demo_clouds, demo_params, to_oracle_cloud
    src, tgt = demo_clouds(True)
    p = demo_params(src, tgt, True)
    cs, ct = to_oracle_cloud(src), to_oracle_cloud(tgt)
    out = []
    for ax, eps in ((0, 1e-7), (1, 1e-7), (2, 1e-7), (0, -1e-7), (1, 3e-7), (2, -2e-7)):
        Ti = np.eye(4, dtype=np.float32)
        Ti[ax, 3] = eps
        _, T, info, _ = oracle.align(p, cs, ct, Ti)
        out.append(dict(axis=ax, eps=eps, iterations=int(info.iterations),
                        max_abs_pose_diff=float(np.abs(np.asarray(T) - np.asarray(T0)).max())))
    return dict(starts=out, spread=max(o["max_abs_pose_diff"] for o in out))


def main():
    with open(os.path.join(HERE, "edge_updates.json"), "w") as fh:
        json.dump([run_edge_case(c) for c in EDGE_CASES], fh, indent=0)
    print("edge_updates", len(EDGE_CASES), "records")
    for name, c in CASES.items():
        res = run_case(c)
        with open(os.path.join(HERE, f"{name}_iters.json"), "w") as fh:
            json.dump(dict(case=name, iters=res["iters"]), fh, indent=0)
        if name == "demo_color":
            with open(os.path.join(HERE, "demo_color_align.json"), "w") as fh:
                json.dump(dict(case=name, ret=res["ret"], iterations=res["iterations"],
                               transform=res["transform"], oracle_spread=demo_final_pose_spread(res["transform"])),
                          fh, indent=0)
        print(name, len(res["iters"]), "records; align iterations", res["iterations"])


if __name__ == "__main__":
    main()
