# ncu --set full captures of the small kernels on the step's critical path (2 launches each, 4th eager step)
B="python bench.py --no-graph --no-cpu-baseline --no-extras --steps 1 --warmup 3"
cap() {  # name regex skip count
  ncu --set full --clock-control none --import-source on -k "regex:$2" --launch-skip $3 --launch-count $4 -o gpurun_out/r02t_$1 -f $B > gpurun_out/r02t_$1.log 2>&1
  echo "$1 rc=$? profiled=$(grep -c 'Profiling' gpurun_out/r02t_$1.log)"
}
cap coupling 'rowgemm_tc_kernel.*EpiCoupling' 24 2
cap startdgrad 'rowgemm_tc_kernel.*EpiStartDgrad' 24 2
cap wnbwd 'wn_bwd_kernel' 24 2
cap colsum 'colsum_kernel' 39 3
cap invconv 'invconv_rows_kernel' 48 3
cap gemms 'rowgemm_tc_kernel.*(EpiBiasAct|EpiDgradAct)' 380 10
cap wgrad 'wgrad_tc_kernel' 39 3
ls -la gpurun_out/r02t_*.ncu-rep
