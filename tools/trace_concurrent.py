#!/usr/bin/env python3
"""CUPTI timeline of W concurrent Dense-4M (or CNN) proofs on one GPU: every kernel's start/end without serialising launches.
usage: trace_concurrent.py [workers=16] [proofs=2*workers] [dense4m|cnn264k]   -> gpurun_out/trace_<which>_w<W>.csv + a summary"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-prove_b200")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench, dpb200 as dp
W = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2 * W
which = sys.argv[3] if len(sys.argv) > 3 else "dense4m"
T = C.CDLL(os.path.join(ROOT, "deep-prove_b200", "libdp_trace.so"))
T.dp_trace_stop.restype = C.c_long; T.dp_trace_stop.argtypes = [C.c_char_p]
wl = bench.DenseWorkload() if which == "dense4m" else bench.CnnWorkload()
dp.init(0); wl.setup_device(dp)
wl.ctx.prove_concurrent(W, W)        # warm pools
assert T.dp_trace_start() == 0, "cupti start failed"
sec = wl.ctx.prove_concurrent(W, N)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
path = os.path.join(ROOT, "gpurun_out", "trace_%s_w%d.csv" % (which, W))
n = T.dp_trace_stop(path.encode())
print("workers %d: %d proofs in %.3f s -> %.1f proofs/s (with CUPTI tracing on); %d kernel records -> %s" % (W, N, sec, N / sec, n, path))
import trace_summary
trace_summary.main(path)
