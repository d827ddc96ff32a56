"""One sumcheck proof sharded over ranks with the CUDA engine (multigpu.prove_sharded_device): must equal the unsplit
proof of the CPU checker (identity asserted by the reference at zkml/src/model/mod.rs:987-993)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_py as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_world1_equals_prove_parallel(gpu):
    sys.path.insert(0, os.path.join(ROOT, "deep-prove_b200"))
    import multigpu as mg
    nv = 11
    full = [(O.splitmix_f(21, 1 << nv), False), (O.splitmix_e(22, 1 << nv), True)]
    products = [((3, 0), [0, 1]), ((1, 2), [1, 1, 0])]
    got = mg.prove_sharded_device([gpu.Mle.upload(a, e) for a, e in full], products, nv, 0, 1, None)
    exp = O.sumcheck_prove(full, products, nv)
    for g, e in zip(got, exp):
        assert (np.asarray(g) == e).all()


WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests")); sys.path.insert(0, os.path.join(%(root)r, "deep-prove_b200"))
import torch, torch.distributed as dist
import oracle_py as O
import dpb200 as dp
import multigpu as mg
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
dp.init(0)                                     # both ranks share cuda:0 in this test; the exchange goes over gloo
nv = 13
full = [(O.splitmix_f(31, 1 << nv), False), (O.splitmix_f(32, 1 << nv), False), (O.splitmix_e(33, 1 << nv), True)]
products = [((1, 0), [0, 1, 2]), ((9, 4), [2, 0])]
lo, hi = mg.shard_range(1 << nv, rank, world)
local = [dp.Mle.upload((a.reshape(-1, 2)[lo:hi] if e else a[lo:hi]).copy(), e) for a, e in full]
point, msgs, fin = mg.prove_sharded_device(local, products, nv, rank, world, mg.TorchAllGather(dist))
ep, em, ef = O.sumcheck_prove(full, products, nv)
ok = bool((point == ep).all() and (msgs == em).all() and (fin == ef).all())
# the same proof through the C++ host mirror: shared-memory mailbox, then a torch.distributed callback exchange
mb = mg.ShmMailbox("dpb200_test_%%d" %% os.getppid(), rank, world, dist.barrier)
for kw in ({"mailbox": mb}, {"allgather": mg.TorchAllGather(dist)}, {"mailbox": mb}):
    local = [dp.Mle.upload((a.reshape(-1, 2)[lo:hi] if e else a[lo:hi]).copy(), e) for a, e in full]
    p2, m2, f2 = mg.prove_sharded_native(local, products, nv, rank, world, **kw)
    ok = ok and bool((p2 == ep).all() and (m2 == em).all() and (f2 == ef).all())
mb.close(dist.barrier)
flags = [None] * world
dist.all_gather_object(flags, ok)
if rank == 0:
    print(json.dumps({"ok": all(flags), "world": world, "rounds": int(msgs.shape[0])}))
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_ranks_equal_unsplit_proof(gpu, tmp_path, world):
    w = tmp_path / "sharded_gpu.py"
    w.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29620 + world), str(w)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert out == {"ok": True, "world": world, "rounds": 13}
