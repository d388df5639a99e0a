# ncu --set full of the row-GEMM launches of one flow, forward (start, 4 x (in_layer, res_skip), end) and backward
# (end dgrad, 4 layer dgrads, g_x0 dgrad, start dgrad), 4th eager step (145 rowgemm_tc launches per step)
B="python bench.py --no-graph --no-cpu-baseline --no-extras --steps 1 --warmup 3"
ncu --set full --clock-control none --import-source on -k regex:rowgemm_tc_kernel --launch-skip 440 --launch-count 11 -o gpurun_out/r02z_rowgemm_fwd -f $B > gpurun_out/r02z_fwd.log 2>&1
echo fwd rc=$? profiled=$(grep -c Profiling gpurun_out/r02z_fwd.log)
ncu --set full --clock-control none --import-source on -k regex:rowgemm_tc_kernel --launch-skip 520 --launch-count 9 -o gpurun_out/r02z_rowgemm_bwd -f $B > gpurun_out/r02z_bwd.log 2>&1
echo bwd rc=$? profiled=$(grep -c Profiling gpurun_out/r02z_bwd.log)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02z_launches.csv $B > gpurun_out/r02z_ncu_launches.log 2>&1
echo launches rc=$?
ls -la gpurun_out/r02z*
