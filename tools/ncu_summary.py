"""Summarise an ncu --set full report: one block of key metrics per profiled launch, and (with
--traffic WORKLOAD) the dram bytes per launch merged into profiles/traffic.json and the pipe /
issue utilisation merged into profiles/pipes.json (what bench.py reports as `roofline.traffic`
and `pipes`).
usage: ncu_summary.py report.ncu-rep [--traffic WORKLOAD] [--source profiles/<name>_summary.txt] > profiles/<name>_summary.txt"""
import csv, io, json, os, subprocess, sys

rep = sys.argv[1]
wl = sys.argv[sys.argv.index("--traffic") + 1] if "--traffic" in sys.argv else None
src_name = sys.argv[sys.argv.index("--source") + 1] if "--source" in sys.argv else os.path.basename(rep)
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio"]
stall = [h for h in hdr if "pcsamp_warps_issue_stalled" in h and not h.endswith("_not_issued")]
traffic = {}
pipes = {}
K = hdr.index("Kernel Name")
for r in rows[2:]:
    name = r[K].split("(")[0].replace("void ", "")
    if name.startswith("align_grid_kernel"):
        name = "align_grid_kernel"  # all block-size / fused instantiations are the same kernel
    name = name.replace("<(bool)1>", "<true>").replace("<(bool)0>", "<false>").replace("<(int)", "<").replace(")>", ">") if "(int)" in name or "(bool)" in name else name
    print(f"== {r[K]}")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"   {w:70s} {r[i]:>16s} {units[i]}")
    tot = sum(float(r[hdr.index(h)] or 0) for h in stall) or 1.0
    top = sorted(((float(r[hdr.index(h)] or 0), h) for h in stall), reverse=True)[:6]
    print("   stall samples: " + ", ".join(f"{h.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * v / tot:.0f}%" for v, h in top))
    def to_bytes(col):
        i = hdr.index(col)
        v = float(r[i].replace(",", "") or 0)
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[i], 1)
    traffic[name] = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
    def num(col):
        return float(r[hdr.index(col)].replace(",", "") or 0) if col in hdr else None
    pipes[name] = {"fma_pipe_pct": num("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                   "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
                   "kernel_ms": num("gpu__time_duration.sum"),
                   "top_stalls": [f"{h.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * v / tot:.0f}%" for v, h in top[:4]],
                   "source": src_name}
if wl:
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    try:
        cur = json.load(open(path))
    except Exception:
        cur = {}
    for k, v in traffic.items():
        cur[f"{wl}:{k}"] = v
    json.dump(cur, open(path, "w"), indent=1, sort_keys=True)
    path = os.path.join(os.path.dirname(path), "pipes.json")
    try:
        cur = json.load(open(path))
    except Exception:
        cur = {}
    for k, v in pipes.items():
        cur[f"{wl}:{k}"] = v
    json.dump(cur, open(path, "w"), indent=1, sort_keys=True)
