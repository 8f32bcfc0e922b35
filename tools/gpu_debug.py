"""First-contact GPU script: small iterate() cases vs the oracle, verbose diffs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle
import unified_cvo_b200 as u
from helpers import *

def show(tr):
    return dict(nnz=tr.nnz, mx=tr.max_row_nnz, om=np.round(list(tr.omega), 6).tolist(), v=np.round(list(tr.v), 6).tolist(),
                B=tr.B, C=tr.C, D=tr.D, E=tr.E, step=tr.step, a_sum=tr.a_sum, dist=tr.dist)

def case(name, src, tgt, p, ell, cap, R=np.eye(3), T=np.zeros(3)):
    g = u.CvoGPU(p)
    g.set_cloud(0, src); g.set_cloud(1, tgt)
    t0 = time.time(); got = g.iterate(R, T, ell, cap); t1 = time.time()
    ref = oracle.iterate(p, to_oracle_cloud(src), to_oracle_cloud(tgt), np.asarray(R, np.float32).T.reshape(9), T, ell, cap)
    bad = compare_traces(got, ref)
    print(f"[{name}] N={src.num_points()} M={tgt.num_points()} ell={ell} cap={cap} gpu_ms={1e3*(t1-t0):.2f} ->", "OK" if not bad else bad)
    if bad:
        print("  got", show(got)); print("  ref", show(ref))
    g.close()
    return not bad

ok = True
src, tgt, _ = synthetic_pair(700, 500, 600, 7)
ok &= case("tiny-geo", src, tgt, geometric_params(), 0.95, 256)
ok &= case("tiny-geo-cap3", src, tgt, geometric_params(), 2.5, 3)
src, tgt, _ = synthetic_pair(2500, 2000, 2000, 20002)
ok &= case("2k-geo", src, tgt, geometric_params(), 0.95, 256)
ok &= case("2k-geo-ell.3", src, tgt, geometric_params(), 0.3, 12)
ds, dt = demo_clouds(True)
ok &= case("demo-color", ds, dt, demo_params(ds, dt, True), 5.76, 256)
ok &= case("demo-color-ell1", ds, dt, demo_params(ds, dt, True), 1.0, 20)
ds2, dt2 = demo_clouds(False)
ok &= case("demo-geo", ds2, dt2, demo_params(ds2, dt2, False), 5.76, 256)
src, tgt, _ = synthetic_pair(12500, 10000, 10000, 20002)
ok &= case("C2", src, tgt, geometric_params(), 0.95, 256)
print("ALL OK" if ok else "SOME FAILED")

# full align on a small case
src, tgt, Tgt = synthetic_pair(2500, 2000, 2000, 20002)
p = geometric_params()
g = u.CvoGPU(p)
t0 = time.time(); ret, Tm, info, tr = g.align(src, tgt, None, trace_cap=4096); t1 = time.time()
print("align gpu: ret", ret, "iters", info.iterations, "stop", info.stop_reason, "ell", info.final_ell, "reg_s", info.registration_seconds, "wall", t1 - t0)
r2, T2, i2, tr2 = oracle.align(p, to_oracle_cloud(src), to_oracle_cloud(tgt), None, trace_cap=4096)
print("align ref: ret", r2, "iters", i2.iterations, "stop", i2.stop_reason, "ell", i2.final_ell)
print("pose err", np.abs(Tm - T2).max())
nbad = 0
for k in range(min(len(tr), len(tr2))):
    bad = compare_traces(tr[k], tr2[k])
    if bad:
        nbad += 1
        if nbad < 5: print("iter", k, bad)
print("trace mismatches", nbad, "of", min(len(tr), len(tr2)))
print("fma peak scalar %.3e  packed %.3e lane-FMA/s" % (g.fma_peak(0, 8192), g.fma_peak(1, 8192)))
src, tgt, _ = synthetic_pair(12500, 10000, 10000, 20002)
g.set_cloud(0, src); g.set_cloud(1, tgt)
for ell in (0.95, 0.3, 0.1):
    ms, msp = g.time_iterations(np.eye(3), np.zeros(3), ell, 64, 50)
    print(f"C2 ell={ell}: {ms/50*1e3:.1f} us/iter, pair kernel {msp/50*1e3:.1f} us -> {1e8/(msp/50*1e-3):.3e} pairs/s (kernel) {1e8/(ms/50*1e-3):.3e} (iteration)")
