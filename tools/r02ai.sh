python -m pytest tests/test_lstm_gpu.py tests/test_model_gpu.py tests/test_trainer_gpu.py tests/test_cfg1_gpu.py tests/test_infer_gpu.py -m gpu -q -x > gpurun_out/r02ai_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r02ai_tests.log
python bench.py --no-cpu-baseline --no-extras --steps 16 > gpurun_out/r02ai_bench.json 2>gpurun_out/r02ai_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02ai_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"], d["roofline"]["step_frac_of_sustained_peak"], d["loss_last"], d["gpu_launches"])
PY
