for c in 1 2 3; do RADTTS_RADAM_PIPELINED_CTAS=$c python bench.py --no-cpu-baseline --no-extras --steps 16 > gpurun_out/r02x_bench_def$c.json 2>/dev/null; done
python bench.py --no-cpu-baseline --no-extras --no-deferred-update --steps 16 > gpurun_out/r02x_bench_plain.json 2>/dev/null
python - <<'PY'
import json
for t in ("def1","def2","def3","plain"):
    d=json.loads(open("gpurun_out/r02x_bench_%s.json"%t).read().strip().splitlines()[-1])
    print(t, d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["ms_per_launch"], d["loss_last"])
PY
