ncu --set full --clock-control none --import-source on -k regex:${LSTM_K:-lstm_fwd_cluster_reg_kernel} --launch-skip 3 --launch-count 1 -o gpurun_out/r02ad_${LSTM_TAG:-lstm_fwd} -f python tools/lstm_bench.py > gpurun_out/r02ad_lstm_ncu.log 2>&1
echo rc=$? $(grep -c Profiling gpurun_out/r02ad_lstm_ncu.log)
