"""N=2 data-parallel step == single-process step (needs 2 GPUs; skipped otherwise).  Exercises the path bench.py runs
at N>1: two CUDA graphs, external per-flow events, NCCL all-reduces on the comm stream overlapping the backward graph."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("graph", ["1", "0"])
def test_two_gpu_dp_step_equals_single_process(graph, cuda_lib):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    env = dict(os.environ, DP_TEST_GRAPH=graph)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("DPRESULT ")][-1]
    res = json.loads(line[len("DPRESULT "):])
    assert res["world"] == 2 and res["n_regions"] == 8
    assert all(res["flow_final"]), res            # every flow's region took the early (overlapped) route
    assert res["grad_rel_err"] < 1e-4, res
    assert res["param_rel_err"] < 1e-6, res
    assert res["replica_checksums_equal"], res
