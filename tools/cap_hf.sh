set -u
O=gpurun_out
ncu --clock-control none --set full --import-source on -k regex:hh_frag_kernel --launch-skip 1 --launch-count 1 -f -o /tmp/hf python tools/ncu_target.py trimer10 > $O/r2b_ncu_hf.log 2>&1
python tools/ncu_metrics_md.py /tmp/hf.ncu-rep "trimer10: hh_frag_kernel (ncu --set full, one launch, 16384 points)" 16384 > $O/r2b_metrics_trimer10_hf.md 2>> $O/r2b_ncu_hf.log
python tools/ncu_stalls.py /tmp/hf.ncu-rep quantum-targeted-energy-transfer_b200/build/qtet_hhfrag.o hh_frag_kernelILi66 ${BUCKET:-10} > $O/r2b_stalls_trimer10_hf.md 2>> $O/r2b_ncu_hf.log
python - <<EOF
import re
rows=[l for l in open("gpurun_out/r2b_stalls_trimer10_hf.md") if l.startswith("| qtet_hhfrag") or l.startswith("| sm_30")]
rows.sort(key=lambda l:(l.split("|")[1].split(":")[0], int(re.search(r":(\\d+)-", l).group(1))))
print("".join(rows))
EOF
grep -E "duration|issue_active|fp64_cycles" $O/r2b_metrics_trimer10_hf.md
