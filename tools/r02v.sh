python -m pytest tests -m gpu -q > gpurun_out/r02v_tests.log 2>&1; echo tests rc=$?; tail -6 gpurun_out/r02v_tests.log
python bench.py --no-cpu-baseline --no-extras > gpurun_out/r02v_bench_def.json 2>gpurun_out/r02v_bench_def.err; python bench.py --no-cpu-baseline --no-extras --no-deferred-update > gpurun_out/r02v_bench_plain.json 2>/dev/null
python - <<'PY'
import json
for t in ("def","plain"):
    d=json.loads(open("gpurun_out/r02v_bench_%s.json"%t).read().strip().splitlines()[-1])
    print(t, d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"], d["loss_last"], d["gpu_launches"])
PY
