"""SASS evidence for profiles/: per kernel of libcvo_b200.so the histogram of the instructions that
matter on sm_100a (UBLKCP = 1-D TMA bulk copy, SYNCS = mbarrier, FFMA2 = packed fp32x2 FMA, FMNMX3 =
3-input min/max, IDP4A, LDS/LDG widths) and an excerpt around the first TMA copy and the densest
FFMA2 run.  usage: sass_excerpt.py [kernel-name-substring ...] > profiles/sass_<name>.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "unified_cvo_b200", "csrc", "libcvo_b200.so")
want = sys.argv[1:] or ["tile_kernel"]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)}   (arch: {', '.join(sorted(set(re.findall(r'arch = (sm_\w+)', txt))))})")
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    if not any(w in dem for w in want):
        continue
    lines = [l for l in f.split("\n") if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l)]
    ops = [re.sub(r"^\s*/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?", "", l).split()[0].rstrip(";") for l in lines]
    hist = collections.Counter(o.split(".")[0] for o in ops)
    print(f"\n== {dem}\n   {len(lines)} instructions")
    keys = ["UBLKCP", "SYNCS", "FFMA2", "FMNMX3", "FMNMX", "FFMA", "FADD", "FMUL", "FSETP", "IDP4A", "DFMA", "DMUL", "MUFU",
            "LDS", "STS", "LDG", "STG", "ATOMG", "REDG", "SHFL", "VOTE", "BAR", "WARPSYNC", "MEMBAR", "ELECT"]
    print("   " + "  ".join(f"{k} {hist[k]}" for k in keys if hist[k]))
    wide = collections.Counter(o for o in ops if o.startswith(("LDS.", "LDG.", "STG.", "STS.")))
    print("   memory widths: " + "  ".join(f"{k} {v}" for k, v in sorted(wide.items())))
    def excerpt(idx, title, before=4, after=14):
        print(f"   -- {title}")
        for l in lines[max(0, idx - before): idx + after]:
            print("   " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip())
    tma = [i for i, o in enumerate(ops) if o.startswith("UBLKCP")]
    if tma:
        excerpt(tma[0], "first TMA bulk copy (cp.async.bulk -> UBLKCP) and its mbarrier (SYNCS)")
    best, run, start = (0, 0), 0, 0
    for i, o in enumerate(ops):  # densest window of 24 instructions in FFMA2
        pass
    dens = [sum(1 for o in ops[i:i + 24] if o.startswith("FFMA2")) for i in range(max(1, len(ops) - 24))]
    if dens and max(dens) > 0:
        i = dens.index(max(dens))
        excerpt(i, f"densest packed-FMA window ({max(dens)} FFMA2 in 24 instructions)", 0, 24)
