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
cap() {  # cap <workload> <points> <ncu kernel regex> <launch skip> <object> <mangled-name fragment> <capture name>
  $NCU --set full --import-source on -k regex:$3 --launch-skip $4 --launch-count 1 -f -o /tmp/${TAG}_$7 python tools/ncu_target.py $1 $2 > $O/${TAG}_ncu_$7.log 2>&1
  python tools/ncu_metrics_md.py /tmp/${TAG}_$7.ncu-rep "$1: $3 (ncu --set full --clock-control none, one launch of tools/ncu_target.py $1 $2, $2 points)" $2 > $O/${TAG}_metrics_$7.md 2>> $O/${TAG}_ncu_$7.log
  python tools/ncu_stalls.py /tmp/${TAG}_$7.ncu-rep $B/$5 $6 > $O/${TAG}_stalls_$7.md 2>> $O/${TAG}_ncu_$7.log
}
if [ "$WHAT" = all ] || [ "$WHAT" = full ]; then
python tools/ncu_target.py dimer20 65536 > $O/${TAG}_target_dimer20.log 2>&1 && {
  # 65536 points = one chunk per pass: K_E, K_D, fallback list; the second pass is captured
  cap dimer20 65536 direct_kernel 1 qtet_direct.o direct_kernelILi21ELi8ELi2ELi4ELb1 dimer20_kd
  cap dimer20 65536 eig_scalar_kernel 1 qtet_direct.o eig_scalar_kernel dimer20_ke
}
python tools/ncu_target.py trimer10 65536 > $O/${TAG}_target_trimer10.log 2>&1 && {
  # one chunk per pass: Householder stages A, B, C, eigenvalues, eigenvectors + evolution + gradient
  cap trimer10 65536 hh_frag_kernel 3 qtet_hhfrag.o hh_frag_kernelILi66ELi3ELi4ELb0ELi1 trimer10_hhA
  cap trimer10 65536 hh_frag_kernel 4 qtet_hhfrag.o hh_frag_kernelILi48ELi1ELi10ELb0ELi2 trimer10_hhB
  cap trimer10 65536 hh_frag_kernel 5 qtet_hhfrag.o hh_frag_kernelILi66ELi3ELi4ELb0ELi3 trimer10_hhC
  cap trimer10 65536 eig_scalar_big 1 qtet_big.o eig_scalar_big_kernelILb1 trimer10_eig
  cap trimer10 65536 replay_reg_kernel 1 qtet_bigreg.o replay_reg_kernelILi66ELi3ELi4ELi4ELb1 trimer10_replay
}
python tools/ncu_target.py dimer100 65536 > $O/${TAG}_target_dimer100.log 2>&1 && {
  cap dimer100 65536 replay_reg_kernel 1 qtet_bigreg.o replay_reg_kernelILi101ELi4ELi2ELi2ELb1 dimer100_replay
  cap dimer100 65536 eig_scalar_big 1 qtet_big.o eig_scalar_big_kernelILb1 dimer100_eig
}
fi
ls -la $O | tail -50
