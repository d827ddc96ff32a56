"""The identities the sharded Basefold (csrc/basefold.cu dp_pcs_commit_shard, host/mpcs.hpp commit_sharded / open_sharded) rests on,
checked on the CPU against the checker's unsharded commitment -- independent of any GPU code:
  (1) rank g's contiguous slice of the BIT-REVERSED codeword is a size-N/G decimation-in-frequency NTT of
      y_r[t] = w_N^(t r) sum_k x[t + k S] w_G^(k r),  r = bitrev(g)       (k_shard_expand; SURVEY.md 8e "Basefold sharding");
  (2) the Merkle tree over the whole codeword is the G subtrees of the slices plus log G levels over their roots;
  (3) the slice of the eq table is the table over the low variables times the eq factor of the rank's top bits;
and, over gloo with world_size 2 and 4, the exchange that assembles the commitment root from the ranks' subtree roots."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_py as O

P = 0xFFFFFFFF00000001
ROOT32 = 1753635133440165772
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def brev(i, bits):
    return int(format(i, "0%db" % bits)[::-1], 2) if bits else 0


def dif_ntt_bitrev(a, w):
    """in-place radix-2 DIF: natural-order input, bit-reversed output; w = primitive len(a)-th root"""
    a = list(a); n = len(a); half = n // 2; step = 1
    while half >= 1:
        for start in range(0, n, 2 * half):
            for j in range(half):
                u, v = a[start + j], a[start + j + half]
                a[start + j] = (u + v) % P
                a[start + j + half] = (u - v) * pow(w, j * step, P) % P
        half //= 2; step *= 2
    return a


@pytest.mark.parametrize("nv,logG", [(8, 1), (8, 2), (9, 3)])
def test_codeword_slice_is_a_local_ntt(nv, logG):
    ev = O.splitmix_f(40 + nv, 1 << nv)
    full_log = nv + 1                                   # a larger parameter set: the coset shift is 7^(2^(full_log - nv))
    root, cw, bh = O.pcs_commit(ev, False, full_log)
    coef = [int(v) for v in O.interpolate_hc(bh, False)]     # the vector the reference encodes is interp(bit-reversed evaluations)
    m, N, G = 1 << nv, 2 << nv, 1 << logG
    S = N // G
    shift = pow(7, 1 << (full_log - nv), P)
    wN = pow(ROOT32, 1 << (32 - (nv + 1)), P)
    wG = pow(wN, S, P)
    x = [coef[j] * pow(shift, j, P) % P for j in range(m)] + [0] * m
    for g in range(G):
        r = brev(g, logG)
        y = [pow(wN, t * r, P) * sum(x[t + k * S] * pow(wG, (k * r) % G, P) for k in range(G)) % P for t in range(S)]
        local = dif_ntt_bitrev(y, pow(wN, G, P))
        assert local == [int(v) for v in cw[g * S:(g + 1) * S]], "slice %d of %d" % (g, G)


@pytest.mark.parametrize("is_ext", [False, True])
def test_tree_is_subtrees_plus_top(is_ext):
    nv, logG = 9, 2
    ev = O.splitmix_e(3, 1 << nv) if is_ext else O.splitmix_f(3, 1 << nv)
    root, cw, _ = O.pcs_commit(ev, is_ext, nv)
    G, S = 1 << logG, (2 << nv) >> logG
    level = [O.merkle_root(cw[g * S:(g + 1) * S], is_ext) for g in range(G)]
    while len(level) > 1:
        level = [O.compress(level[2 * i], level[2 * i + 1]) for i in range(len(level) // 2)]
    assert (level[0] == root).all()


def test_eq_slice_is_low_table_times_top_factor():
    nv, logG = 7, 2
    rp = O.splitmix_e(9, nv)
    full = O.build_eq(rp)
    low = O.build_eq(rp[:nv - logG])
    Ml = 1 << (nv - logG)
    one = np.array([1, 0], dtype=np.uint64)
    for g in range(1 << logG):
        scal = one
        for j in range(logG):
            x = rp[nv - logG + j]
            scal = O.e_binop(2, scal[None, :], (x if (g >> j) & 1 else O.e_binop(1, one[None, :], x[None, :])[0])[None, :])[0]      # 1 = sub, 2 = mul
        want = O.e_binop(2, low, np.repeat(scal[None, :], Ml, axis=0))
        assert (want == full[g * Ml:(g + 1) * Ml]).all()


WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests")); sys.path.insert(0, os.path.join(%(root)r, "deep-prove_b200"))
import torch, torch.distributed as dist
import oracle_py as O
import multigpu as mg
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
nv = 10
ev = O.splitmix_f(17, 1 << nv)
root, cw, _ = O.pcs_commit(ev, False, nv)                 # every rank holds the polynomial (as the sharded commit does)
lo, hi = mg.shard_range(cw.shape[0], rank, world)
mine = O.merkle_root(cw[lo:hi], False)                    # this rank's subtree root
roots = mg.TorchAllGather(dist, device="cpu")(mine)       # the one exchange of the commit: 32 bytes per rank
level = [roots[g] for g in range(world)]
while len(level) > 1:
    level = [O.compress(level[2 * i], level[2 * i + 1]) for i in range(len(level) // 2)]
ok = bool((level[0] == root).all())
flags = [None] * world
dist.all_gather_object(flags, ok)
if rank == 0:
    print(json.dumps({"ok": all(flags), "world": world}))
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 4])
def test_commit_root_from_subtree_roots_over_gloo(tmp_path, world):
    w = tmp_path / "w.py"
    w.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1", "--master-port", str(29640 + world), str(w)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert out == {"ok": True, "world": world}
