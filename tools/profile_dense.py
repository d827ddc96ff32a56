#!/usr/bin/env python3
"""Developer profile of the Dense-4M proof: wall time per proof, host-side time by C-ABI entry point
(DP_HOST_PROF=1) and device time by kernel (CUDA events).  Usage: DP_HOST_PROF=1 python tools/profile_dense.py [n]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "deep-prove_b200"))
import bench  # noqa: E402
import dpb200 as dp  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
wl = bench.DenseWorkload() if (len(sys.argv) < 3 or sys.argv[2] == "dense4m") else bench.SumcheckWorkload()
dp.init(0)
wl.setup_device(dp)
for i in range(2):
    wl.step_resident(i)
dp.lib().dp_synchronize()
dp.lib().dp_hostprof_dump()          # discard warm-up accumulators
l0 = dp.lib().dp_kernel_launches()
t0 = time.perf_counter()
for i in range(n):
    wl.step_resident(i)
dp.lib().dp_synchronize()
dt = (time.perf_counter() - t0) / n
print("wall per proof: %.3f ms, kernel launches per proof: %d" % (dt * 1e3, (dp.lib().dp_kernel_launches() - l0) / n))
sys.stdout.flush()
dp.lib().dp_hostprof_dump()
dp.profile_reset(); dp.profile_enable(True)
for i in range(n):
    wl.step_resident(i)
prof = dp.profile_read()
dp.profile_enable(False)
tot = sum(v[1] for v in prof.values())
print("device kernel time per proof: %.3f ms" % (tot / n))
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print("  %-36s launches/proof %7.1f  ms/proof %8.3f  avg %8.2f us  alg GB/s %9.1f" % (k, v[0] / n, v[1] / n, 1e3 * v[1] / v[0], (v[2] / (v[1] * 1e-3) / 1e9) if v[1] > 0 else 0))
