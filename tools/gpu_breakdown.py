"""Marginal cost of each kernel inside the real pipeline (CVO_B200_DEBUG_SKIP), graph-free."""
import os, sys, subprocess, json
HERE = os.path.dirname(os.path.abspath(__file__))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
    import numpy as np, unified_cvo_b200 as u
    from helpers import *
    name, ell = sys.argv[2], float(sys.argv[3])
    P, N, M, seed, F, C = u.synthetic.CONFIGS[name]
    src, tgt, _ = synthetic_pair(P, N, M, seed, F=F, C=C)
    p = geometric_params() if F == 0 else u.read_params_yaml(os.path.join(DATA, "cvo_intensity_params_img_gpu0.yaml"))
    g = u.CvoGPU(p); g.set_cloud(0, src); g.set_cloud(1, tgt)
    g.time_iterations(np.eye(3), np.zeros(3), ell, 64, 200, pair_kernel=False)
    nit = 40 if N > 50000 else 400
    best = min(g.time_iterations(np.eye(3), np.zeros(3), ell, 256, nit, pair_kernel=False)[0] for _ in range(3))
    print(json.dumps({"us_per_iter": best / nit * 1e3}))
else:
    name = sys.argv[1] if len(sys.argv) > 1 else "C2"
    ell = sys.argv[2] if len(sys.argv) > 2 else "0.95"
    res = {}
    for label, mask in [("all", 0), ("no_prep", 1), ("no_pair", 2), ("no_flow", 4), ("only_prep_step", 6), ("only_step", 7)]:
        env = dict(os.environ, CVO_B200_DEBUG_SKIP=str(mask))
        out = subprocess.run([sys.executable, __file__, "child", name, ell], capture_output=True, text=True, env=env)
        try: res[label] = json.loads(out.stdout.strip().splitlines()[-1])["us_per_iter"]
        except Exception: res[label] = out.stderr[-300:]
    print(name, "ell", ell, {k: (round(v, 1) if isinstance(v, float) else v) for k, v in res.items()})
    try:
        a = res["all"]; print("marginal: prep %.1f pair %.1f flow %.1f ; prep+step alone %.1f ; step alone %.1f" % (a - res["no_prep"], a - res["no_pair"], a - res["no_flow"], res["only_prep_step"], res["only_step"]))
    except Exception as e: print(e)
