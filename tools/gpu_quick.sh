#!/bin/bash
# quick correctness + speed check: usage gpu_quick.sh TAG
TAG=${1:-q}; O=gpurun_out; mkdir -p $O
timeout 700 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; tail -3 $O/${TAG}_pytest.log
timeout 300 python bench.py --steps 5 --no-cpu-baseline > $O/${TAG}_bench.log 2>&1
timeout 300 python bench.py --workload C4 --steps 2 --warmup 2 --no-cpu-baseline > $O/${TAG}_c4.log 2>&1
python - <<PY
import json
for f in ["$O/${TAG}_bench.log", "$O/${TAG}_c4.log"]:
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l)
            print(f, "value %.3g e2e %.3g ms/step %.2f iters %.0f us/iter %.1f"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"]["iterations_per_step"], 1e3*d["ms_per_step"]/d["config"]["iterations_per_step"]))
            for fp in d.get("frame_pairs",[]): print("   ", fp["workload"][:14], "fps %.1f e2e %.1f iters %.0f us/iter %.1f err %.2g"%(fp["frame_pairs_per_s"], fp["e2e_frame_pairs_per_s"], fp["iterations_per_frame_pair"], 1e3*fp["ms_per_frame_pair"]/fp["iterations_per_frame_pair"], fp["max_abs_pose_error_vs_truth"]))
PY
tail -2 $O/${TAG}_bench.log | cut -c1-300
