# end-of-round artefacts: launch list + ncu --set full of the forward row GEMMs of one flow (incl. the fixed `end` GEMM)
B="python bench.py --no-graph --no-cpu-baseline --no-extras --steps 1 --warmup 3"
ncu --set full --clock-control none --import-source on -k regex:rowgemm_tc_kernel --launch-skip 440 --launch-count 11 -o gpurun_out/r02f_rowgemm_fwd -f $B > gpurun_out/r02f_fwd.log 2>&1
echo fwd rc=$? profiled=$(grep -c Profiling gpurun_out/r02f_fwd.log)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02f_launches.csv $B > gpurun_out/r02f_ncu_launches.log 2>&1
echo launches rc=$?
