#!/usr/bin/env python3
"""Throughput of concurrent Dense-4M proofs on one GPU vs number of host threads/streams."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-prove_b200"))
import bench, dpb200 as dp
wl = bench.DenseWorkload(); dp.init(0); wl.setup_device(dp)
for nw in [1, 2, 4, 8, 12, 16, 24, 32]:
    n = max(8, 4 * nw)
    wl.ctx.prove_concurrent(nw, nw)            # warm the threads' pools
    sec = wl.ctx.prove_concurrent(nw, n)
    print("workers %2d: %3d proofs in %.3f s -> %.1f proofs/s (%.1f ms/proof/stream)" % (nw, n, sec, n / sec, 1e3 * sec * nw / n))
    sys.stdout.flush()
