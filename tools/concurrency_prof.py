import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-prove_b200"))
import bench, dpb200 as dp
wl = bench.DenseWorkload(); dp.init(0); wl.setup_device(dp)
nw = int(sys.argv[1])
wl.ctx.prove_concurrent(nw, nw)
sec = wl.ctx.prove_concurrent(nw, 4 * nw)
print("workers", nw, "proofs/s", 4 * nw / sec)
