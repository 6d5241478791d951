#!/usr/bin/env python3
"""Warp-stall breakdown of ONE kernel of an `ncu --set full --import-source on` report, overall and per source region
(markdown).  The report's source page lists SASS instructions; nvdisasm of the object the kernel was built from gives the
source line of every instruction (same order), so that the samples can be summed per file:line bucket.

    python tools/ncu_stalls.py report.ncu-rep object.o 'kernel substring' [bucket_lines=20] > profiles/x.md
"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import Counter, defaultdict

rep, obj, pat = sys.argv[1], sys.argv[2], sys.argv[3]
bucket = int(sys.argv[4]) if len(sys.argv) > 4 else 20
raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
hdr, vals = raw[0], raw[2]
name = vals[hdr.index("Kernel Name")]
print(f"# Warp stalls: `{name}`\n")
print(f"report `{os.path.basename(rep)}`, duration {vals[hdr.index('gpu__time_duration.sum')]} us, "
      f"{vals[hdr.index('launch__registers_per_thread')]} registers/thread, issue slots busy "
      f"{vals[hdr.index('smsp__issue_active.avg.pct_of_peak_sustained_active')]} %, FP64 pipe "
      f"{vals[hdr.index('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active')]} %\n")
st = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(vals[i]))
      for i, h in enumerate(hdr) if "smsp__average_warps_issue_stalled" in h and "_per_issue_active" in h]
tot = sum(v for _, v in st)
print("## Whole kernel (warps per issued instruction, by state)\n\n| state | warps / issue | share |\n|---|---|---|")
for h, v in sorted(st, key=lambda x: -x[1]):
    if v > 0.001:
        print(f"| {h} | {v:.3f} | {100 * v / tot:.1f} % |")
# ---- per source region
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, capture_output=True)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout.split("\n")
mangled = re.sub(r"[^A-Za-z0-9_]", "", pat)
start = [i for i, l in enumerate(sass) if ".text." in l and pat in l][0]
end = next(i for i in range(start + 3, len(sass)) if ".section" in sass[i])
cur, seq = None, []
for l in sass[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]+\*/\s+\S", l):
        seq.append(cur)
src = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
h2 = src[1]
data = src[2:]
if len(data) != len(seq):
    print(f"\n(source page has {len(data)} instructions, the object {len(seq)}: per-region table skipped)")
    sys.exit(0)
isamp, iex = h2.index("# Samples"), h2.index("Instructions Executed")
cols = {h: i for i, h in enumerate(h2) if h.startswith("stall_") and "(Not Issued)" not in h}
agg = defaultdict(Counter)
for cur, row in zip(seq, data):
    key = "?" if cur is None else f"{cur[0]}:{cur[1] // bucket * bucket}-{cur[1] // bucket * bucket + bucket - 1}"
    agg[key]["samples"] += int(row[isamp]); agg[key]["exec"] += int(row[iex]); agg[key]["static"] += 1
    for hname, i in cols.items():
        agg[key][hname] += int(row[i] or 0)
total = sum(a["samples"] for a in agg.values())
print(f"\n## Per source region ({bucket}-line buckets; {len(seq)} static instructions, {total} samples)\n")
print("| source lines | static instr | executed (warp-level) | samples | top stall reasons |\n|---|---|---|---|---|")
for key, a in sorted(agg.items(), key=lambda x: -x[1]["samples"]):
    if a["samples"] < 0.005 * total:
        continue
    top = sorted(((k.replace("stall_", ""), v) for k, v in a.items() if k.startswith("stall_")), key=lambda x: -x[1])[:4]
    print(f"| {key} | {a['static']} | {a['exec']} | {100 * a['samples'] / total:.1f} % | " +
          ", ".join(f"{k} {100 * v / max(1, a['samples']):.0f} %" for k, v in top) + " |")
