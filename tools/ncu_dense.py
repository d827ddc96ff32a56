#!/usr/bin/env python3
"""Minimal target for ncu: Context::generate + N Dense-4M proofs (default 2), no timing, no oracle.
   ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python tools/ncu_dense.py"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "deep-prove_b200"))
import bench  # noqa: E402
import dpb200 as dp  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
wl = bench.DenseWorkload()
dp.init(0)
wl.ctx = dp.ZkmlContext(wl.NL, wl.W, wl.weights, wl.bias, wl.rq)
l0 = dp.lib().dp_kernel_launches()
wl.ctx.run_inference(wl.x)
for i in range(n):
    wl.ctx.prove_trace()
    print("launches so far:", dp.lib().dp_kernel_launches(), "(context:", l0, ")")
