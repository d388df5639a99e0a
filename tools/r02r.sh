tools/_bin/gemm_micro 10 > gpurun_out/r02r_gemm_micro.txt 2>&1; echo micro rc=$?
python -m pytest tests/test_gemm_tiles_gpu.py tests/test_cfg2_gpu.py tests/test_flow_gpu.py -m gpu -q -x > gpurun_out/r02r_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r02r_tests.log
python tools/flow_bench.py > gpurun_out/r02r_flow_on.json 2>&1; RADTTS_GEMM_TILE_SELECT=0 python tools/flow_bench.py > gpurun_out/r02r_flow_off.json 2>&1
tail -1 gpurun_out/r02r_flow_on.json; tail -1 gpurun_out/r02r_flow_off.json
python bench.py --no-cpu-baseline --no-extras > gpurun_out/r02r_bench_on.json 2>gpurun_out/r02r_bench_on.err; RADTTS_GEMM_TILE_SELECT=0 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r02r_bench_off.json 2>/dev/null
python - <<'PY'
import json
for t in ("on","off"):
    d=json.loads(open("gpurun_out/r02r_bench_%s.json"%t).read().strip().splitlines()[-1])
    print(t, d["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"], d["loss_last"])
PY
