#!/usr/bin/env python3
"""profiles/traffic.json from the committed `ncu --set full` metric summaries (profiles/<tag>_metrics_<workload>_<kernel>.md):
DRAM bytes per point of every hot kernel, keyed the way bench.py looks them up ("<workload>/<path>/<profiler kind>").

    python tools/make_traffic_json.py r2
"""
import json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
# capture name -> (workload, path, profiler kind of bench.py's roofline.kernels)
KIND = {"dimer20_kd": ("dimer20", "direct", "apply"), "dimer20_ke": ("dimer20", "direct", "ql_scalar"),
        "trimer10_hh": ("trimer10", "direct", "generic_or_householder"), "trimer10_eig": ("trimer10", "direct", "ql_scalar"),
        "trimer10_replay": ("trimer10", "direct", "apply"), "dimer100_replay": ("dimer100", "direct", "apply"),
        "dimer100_eig": ("dimer100", "direct", "ql_scalar")}
out = {}
for name, (wl, path, kind) in KIND.items():
    fn = os.path.join(ROOT, "profiles", f"{tag}_metrics_{name}.md")
    if not os.path.exists(fn):
        continue
    txt = open(fn).read()
    m = re.search(r"DRAM traffic per launch = ([0-9.]+) MB for (\d+) points = (\d+) B/point", txt)
    k = re.search(r"^## (.*)$", txt, re.M)
    if not m:
        continue
    out[f"{wl}/{path}/{kind}"] = {
        "dram_bytes_per_launch": float(m.group(1)) * 1e6, "points_per_launch": int(m.group(2)),
        "dram_bytes_per_point": float(m.group(1)) * 1e6 / int(m.group(2)),
        "kernel": k.group(1).strip() if k else name,
        "source": f"profiles/{tag}_metrics_{name}.md (ncu --set full --clock-control none, one launch of tools/ncu_target.py {wl}: "
                  + ("rows 256.. of the config-2 grid" if wl == "dimer20" else "the bench workload's seeded uniform points") + ")"}
# the Householder stage of the dense large-d path is three launches (stages A, B, C): one entry, summed
parts = []
for st in ("hhA", "hhB", "hhC"):
    fn = os.path.join(ROOT, "profiles", f"{tag}_metrics_trimer10_{st}.md")
    if os.path.exists(fn):
        m = re.search(r"DRAM traffic per launch = ([0-9.]+) MB for (\d+) points", open(fn).read())
        if m:
            parts.append((float(m.group(1)) * 1e6, int(m.group(2))))
if parts:
    out["trimer10/direct/generic_or_householder"] = {
        "dram_bytes_per_launch": sum(p[0] for p in parts), "points_per_launch": parts[0][1],
        "dram_bytes_per_point": sum(p[0] / p[1] for p in parts), "kernel": "hh_frag_kernel stages A + B + C",
        "source": f"profiles/{tag}_metrics_trimer10_hhA/hhB/hhC.md (ncu --set full --clock-control none, tools/ncu_target.py trimer10 65536)"}
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
