set -x
python tools/profile_step.py > gpurun_out/r02p_profile_step.log 2>&1
echo profile rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02p_launches.csv python bench.py --no-graph --no-cpu-baseline --no-extras --steps 1 --warmup 3 > gpurun_out/r02p_ncu_launches.log 2>&1
echo launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:rowgemm_tc_kernel --launch-skip 600 --launch-count 40 -o gpurun_out/r02p_rowgemm -f python bench.py --no-graph --no-cpu-baseline --no-extras --steps 1 --warmup 3 > gpurun_out/r02p_ncu_rowgemm.log 2>&1
echo rowgemm rc=$?
tail -3 gpurun_out/r02p_ncu_rowgemm.log
