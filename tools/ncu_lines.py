"""Aggregate an ncu report per CUDA source line for one kernel.
usage: ncu_lines.py report.ncu-rep kernel_regex [topn]"""
import csv, subprocess, sys, io
rep, kre = sys.argv[1], sys.argv[2]; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == 'Line No']
hdr = rows[hi[0]]
IE = hdr.index('Instructions Executed'); WS = hdr.index('Warp Stall Sampling (All Samples)')
end = hi[1] if len(hi) > 1 else len(rows)
data = []
for r in rows[hi[0] + 1:end]:
    if len(r) > IE and r[0].strip().isdigit():
        try: data.append((int(r[0]), float(r[IE] or 0), float(r[WS] or 0), r[1]))
        except ValueError: pass
tot = sum(d[1] for d in data); ts = sum(d[2] for d in data)
print('total warp-instructions', tot, 'stall samples', ts)
for d in sorted(data, key=lambda d: -d[1])[:topn]:
    print(f"{d[0]:5d} {d[1]:>10.0f} {100*d[1]/tot:5.1f}% stall {100*d[2]/max(ts,1):5.1f}%  {d[3].strip()[:100]}")
