python tools/lstm_bench.py > gpurun_out/r02q_lstm_bench.log 2>&1
RADTTS_LSTM_CLUSTER=1 python tools/lstm_bench.py >> gpurun_out/r02q_lstm_bench.log 2>&1
RADTTS_LSTM_CLUSTER=0 python tools/lstm_bench.py >> gpurun_out/r02q_lstm_bench.log 2>&1
cat gpurun_out/r02q_lstm_bench.log
ncu --set full --clock-control none --import-source on -k regex:rowgemm_tc_kernel --launch-skip 420 --launch-count 70 -o gpurun_out/r02q_rowgemm -f python bench.py --no-graph --no-cpu-baseline --no-extras --steps 1 --warmup 3 > gpurun_out/r02q_ncu_rowgemm.log 2>&1
echo rowgemm rc=$?
grep -c "Profiling" gpurun_out/r02q_ncu_rowgemm.log
