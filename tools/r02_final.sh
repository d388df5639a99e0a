python -m pytest tests -m gpu -q > gpurun_out/r02final_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r02final_tests.log
python __graft_entry__.py smoke > gpurun_out/r02final_smoke.log 2>&1; echo smoke rc=$?; tail -1 gpurun_out/r02final_smoke.log
python bench.py > gpurun_out/r02final_bench.json 2> gpurun_out/r02final_bench.err; echo bench rc=$?
python tools/graph_timeline.py gpurun_out/r02final_timeline.json 2>&1 | tail -1
