"""Per-source-line breakdown of one profiled kernel: joins the SASS view of an ncu report
(`ncu -i REP --page source --csv`: address, executed instructions, stall samples) with the line table
of the shipped cubin (`cuobjdump -xelf` + `nvdisasm -g`).

usage: python profiles/ncu_lines.py REP.ncu-rep CUBIN_NAME FUNC_SUBSTR SOURCE_FILE [TOP]
  e.g. python profiles/ncu_lines.py gpurun_out/prof_vdp_ws2.ncu-rep wsolve_kernels WsCfgILi21ELi2ELi1 \
           nlp_filter_b200/csrc/wsolve_kernels.cu 40
"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, cubin_name, func_sub, srcfile = sys.argv[1:5]
top_n = int(sys.argv[5]) if len(sys.argv) > 5 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "nlp_filter_b200", "libnlpfilter.so")], cwd=tmp,
               check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.startswith(cubin_name) and f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# instructions of the chosen function: offset -> (file, line)
order, cur, infunc = {}, None, False
for l in sass.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        infunc = func_sub in m.group(1)
        continue
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
    if m:
        order[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Address" in r)
H = rows[hi]
ai, ii, ss = H.index("Address"), H.index("Instructions Executed"), H.index("# Samples")
base, agg, tot = None, collections.defaultdict(lambda: [0, 0]), [0, 0]
for r in rows[hi + 1:]:
    if len(r) <= max(ai, ii, ss):
        continue
    a = int(r[ai], 16)
    if base is None:
        base = a
    c = order.get(a - base) or ("?", 0)
    ins, smp = int(r[ii] or 0), int(r[ss] or 0)
    agg[c][0] += ins; agg[c][1] += smp; tot[0] += ins; tot[1] += smp
src = {}
for i, l in enumerate(open(os.path.join(root, srcfile)), 1):
    src[(os.path.basename(srcfile), i)] = l.rstrip()
print("total warp-inst", tot[0], "samples", tot[1])
for k, (ins, smp) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top_n]:
    print("%5.1f%% samp %5.1f%% inst  %s:%d  %s" % (100 * smp / max(tot[1], 1), 100 * ins / max(tot[0], 1), k[0], k[1], src.get(k, "")[:110].strip()))
if len(sys.argv) > 6:      # ranges "name:lo-hi,..."
    for spec in sys.argv[6].split(","):
        name, rg = spec.split(":"); lo, hi2 = map(int, rg.split("-"))
        s = [v for k, v in agg.items() if k[0] == os.path.basename(srcfile) and lo <= k[1] <= hi2]
        print("%-14s %5.1f%% samp %5.1f%% inst" % (name, 100 * sum(x[1] for x in s) / tot[1], 100 * sum(x[0] for x in s) / tot[0]))
