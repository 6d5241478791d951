#!/bin/bash
# Round-2 profiling pass (run under gpurun, one GPU): launch lists of the bench commands and one `ncu --set full` capture
# per hot kernel, post-processed ON the box (raw metrics + stall tables as markdown); the large .ncu-rep files stay there.
# Every capture comes after the same command has exited 0 without ncu.
set -u
O=gpurun_out
TAG=${1:-r2}
WHAT=${2:-all}
NCU="ncu --clock-control none"
B=quantum-targeted-energy-transfer_b200/build
if [ "$WHAT" = all ] || [ "$WHAT" = launches ]; then
for w in dimer20 trimer10 dimer100; do
  python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-solver > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err || { echo "bench $w failed"; continue; }
  $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/${TAG}_launches_$w.csv \
      python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-solver > $O/${TAG}_ncu_launch_$w.log 2>&1
  python tools/ncu_summary.py $O/${TAG}_launches_$w.csv > $O/${TAG}_launches_${w}_summary.txt
done
fi
cap() {  # cap <workload> <kernel regex> <skip> <object> <name> <points per launch>
  $NCU --set full --import-source on -k regex:$2 --launch-skip $3 --launch-count 1 -f -o /tmp/${TAG}_$5 python tools/ncu_target.py $1 > $O/${TAG}_ncu_$5.log 2>&1
  python tools/ncu_metrics_md.py /tmp/${TAG}_$5.ncu-rep "$1: $2 (ncu --set full, one launch, $6 points)" $6 > $O/${TAG}_metrics_$5.md 2>> $O/${TAG}_ncu_$5.log
  python tools/ncu_stalls.py /tmp/${TAG}_$5.ncu-rep $B/$4 $2 > $O/${TAG}_stalls_$5.md 2>> $O/${TAG}_ncu_$5.log
}
if [ "$WHAT" = all ] || [ "$WHAT" = full ]; then
python tools/ncu_target.py dimer20 > $O/${TAG}_target_dimer20.log 2>&1 && {
  cap dimer20 direct_kernel 4 qtet_direct.o dimer20_kd 65536
  cap dimer20 eig_scalar_kernel 4 qtet_direct.o dimer20_ke 65536
}
python tools/ncu_target.py trimer10 > $O/${TAG}_target_trimer10.log 2>&1 && {
  cap trimer10 hh_reg_kernel 1 qtet_bigreg.o trimer10_hh 16384
  cap trimer10 eig_scalar_big 1 qtet_big.o trimer10_eig 16384
  cap trimer10 replay_reg_kernel 1 qtet_bigreg.o trimer10_replay 16384
}
python tools/ncu_target.py dimer100 > $O/${TAG}_target_dimer100.log 2>&1 && {
  cap dimer100 replay_reg_kernel 1 qtet_bigreg.o dimer100_replay 16384
  cap dimer100 eig_scalar_big 1 qtet_big.o dimer100_eig 16384
}
fi
ls -la $O | tail -40
