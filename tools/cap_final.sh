set -u
O=gpurun_out; TAG=r2; NCU="ncu --clock-control none"; B=quantum-targeted-energy-transfer_b200/build
cap() {
  $NCU --set full --import-source on -k regex:$3 --launch-skip $4 --launch-count 1 -f -o /tmp/${TAG}_$7 python tools/ncu_target.py $1 $2 > $O/${TAG}_ncu_$7.log 2>&1
  python tools/ncu_metrics_md.py /tmp/${TAG}_$7.ncu-rep "$1: $3 (ncu --set full --clock-control none, one launch of tools/ncu_target.py $1 $2, $2 points)" $2 > $O/${TAG}_metrics_$7.md 2>> $O/${TAG}_ncu_$7.log
  python tools/ncu_stalls.py /tmp/${TAG}_$7.ncu-rep $B/$5 $6 > $O/${TAG}_stalls_$7.md 2>> $O/${TAG}_ncu_$7.log
}
cap trimer10 65536 replay_reg_kernel 1 qtet_bigreg.o replay_reg_kernelILi66ELi3ELi4ELi4ELb1 trimer10_replay
cap trimer10 65536 hh_frag_kernel 5 qtet_hhfrag.o hh_frag_kernelILi66ELi3ELi4ELb0ELi3 trimer10_hhC
cap dimer100 65536 replay_reg_kernel 1 qtet_bigreg.o replay_reg_kernelILi101ELi4ELi2ELi2ELb1 dimer100_replay
grep -E "duration|fp64_cycles|pipe_tensor|issue_active|DRAM traffic" $O/r2_metrics_trimer10_replay.md $O/r2_metrics_trimer10_hhC.md $O/r2_metrics_dimer100_replay.md
