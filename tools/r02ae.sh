python -m pytest tests -m gpu -q > gpurun_out/r02ae_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r02ae_tests.log
python bench.py --no-cpu-baseline --no-extras --steps 16 > gpurun_out/r02ae_bench.json 2>gpurun_out/r02ae_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02ae_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"], d["roofline"]["step_frac_of_sustained_peak"], d["loss_last"], d["gpu_launches"])
PY
python tools/graph_timeline.py gpurun_out/r02ae_timeline.json 2>&1 | tail -1
