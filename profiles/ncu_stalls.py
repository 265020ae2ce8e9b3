"""Stall reasons and shared-memory wavefronts per source-line RANGE of one profiled kernel (companion of
ncu_lines.py; needs a report captured with --import-source on: line numbers are those of the file at capture time).
usage: python profiles/ncu_stalls.py REP.ncu-rep "name:lo-hi,name:lo-hi,..." [FILE_SUBSTR]"""
import collections, csv, io, subprocess, sys
rep, spec = sys.argv[1], sys.argv[2]
want_file = sys.argv[3] if len(sys.argv) > 3 else "wsolve_kernels.cu"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
ranges = []
for s in spec.split(","):
    name, rg = s.split(":"); lo, hi = map(int, rg.split("-")); ranges.append((name, lo, hi))
cur_file, H = None, None
agg = collections.defaultdict(lambda: collections.defaultdict(float))
tot = collections.defaultdict(float)
for r in csv.reader(io.StringIO(out)):
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1]; continue
    if r and r[0] == "Line No":
        H = r; continue
    if H is None or len(r) != len(H) or not r[0]:
        continue                                     # SASS rows have an empty line number
    try:
        line = int(r[0])
    except ValueError:
        continue
    vals = {}
    for i, h in enumerate(H):
        if h.startswith("stall_") and "Not Issued" not in h or h in ("# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Excessive"):
            try:
                vals[h] = float(r[i])
            except ValueError:
                pass
    for k, v in vals.items():
        tot[k] += v
    name = "other"
    if cur_file and want_file in cur_file:
        for (n, lo, hi) in ranges:
            if lo <= line <= hi:
                name = n; break
    else:
        name = "other-files"
    for k, v in vals.items():
        agg[name][k] += v
S = tot["# Samples"]
print("total samples %d, warp instructions %d, shared wavefronts %d (excessive %d)" % (S, tot["Instructions Executed"], tot["L1 Wavefronts Shared"], tot["L1 Wavefronts Shared Excessive"]))
for name in [n for n, _, _ in ranges] + ["other", "other-files"]:
    a = agg[name]
    if not a:
        continue
    st = sorted(((k, v) for k, v in a.items() if k.startswith("stall_")), key=lambda kv: -kv[1])[:5]
    print("%-12s %5.1f%% samples %5.1f%% inst %5.1f%% smem wavefronts (%4.1f%% of them excessive) | %s" % (
        name, 100 * a["# Samples"] / S, 100 * a["Instructions Executed"] / tot["Instructions Executed"],
        100 * a["L1 Wavefronts Shared"] / max(tot["L1 Wavefronts Shared"], 1),
        100 * a["L1 Wavefronts Shared Excessive"] / max(a["L1 Wavefronts Shared"], 1),
        "  ".join("%s %.0f%%" % (k[6:], 100 * v / max(a["# Samples"], 1)) for k, v in st)))
