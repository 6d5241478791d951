#!/usr/bin/env python3
"""Key metrics of every kernel in an `ncu --set full` report as markdown (profiles/*_ncu_full_metrics.md).

    python tools/ncu_metrics_md.py gpurun_out/x.ncu-rep "title line" [points_per_launch] >> profiles/r1_ncu_full_metrics.md
"""
import csv, io, subprocess, sys
rep, title = sys.argv[1], sys.argv[2]
pts = float(sys.argv[3]) if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
print(f"# {title}\n")
for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    print(f"## {name}\n")
    vals = {}
    for w in want:
        if w in h:
            i = h.index(w)
            vals[w] = (r[i], units[i])
            print(f"- {w} = {r[i]} {units[i]}")
    if pts and "dram__bytes_read.sum" in vals:
        def tobytes(v, u):
            x = float(v.replace(",", ""))
            return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        tot = tobytes(*vals["dram__bytes_read.sum"]) + tobytes(*vals["dram__bytes_write.sum"])
        print(f"- DRAM traffic per launch = {tot/1e6:.1f} MB for {int(pts)} points = {tot/pts:.0f} B/point")
    print()
