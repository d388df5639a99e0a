"""Times the in_layer row GEMM with K = nseg x 1024 and epilogue parts disabled (RADTTS_TC_DEBUG) to separate
MMA time from epilogue time.  Usage: RADTTS_DEBUG_NSEG=1 RADTTS_TC_DEBUG=0|1|3 python tools/epi_experiment.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
dev = torch.device("cuda", 0)
model = bench.make_model(dev)
b = bench.to_device(bench.pinned_batch(32, 800, 150, seed=1000), dev)
ms, groups = bench.in_layer_kernel_probe(model, b, iters=20)
print("NSEG=%s DEBUG=%s  %.1f us per launch" % (os.environ.get("RADTTS_DEBUG_NSEG", "5"), os.environ.get("RADTTS_TC_DEBUG", "0"), ms * 1e3))
