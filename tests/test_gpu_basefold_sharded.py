"""Basefold commit + open of ONE polynomial sharded over ranks (SURVEY.md 8e / BASELINE configs[3]): the root and every word of the
proof must equal the unsharded commit + open (which is itself checked against the oracle in test_gpu_basefold.py), and the restated
verifier must accept it.  The ranks are threads of this process sharing cuda:0 (each with its own library context, stream and
device arena); the exchange is the same mailbox protocol the multi-process runs use over POSIX shared memory."""
import os
import sys
import threading

import numpy as np
import pytest

import oracle_py as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-prove_b200"))


def run_sharded(dp, ev, is_ext, nv, pt, world):
    import multigpu as mg
    box = mg.LocalMailbox()
    res, err = [None] * world, []

    def rank_main(r):
        try:
            dp.init(0)
            m = dp.Mle.upload(ev, is_ext)
            res[r] = mg.basefold_commit_open_sharded(m, nv, pt, r, world, mailbox=box.for_rank())
            del m
        except Exception as e:      # surfaced below
            err.append((r, repr(e)))
    th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not err, err
    return res


@pytest.mark.parametrize("nv,world,is_ext", [(14, 2, False), (15, 4, False), (16, 8, False), (14, 2, True), (15, 4, True)])
def test_sharded_equals_unsharded(gpu, nv, world, is_ext):
    ev = O.splitmix_e(31 + nv, 1 << nv) if is_ext else O.splitmix_f(31 + nv, 1 << nv)
    pt = O.splitmix_e(77 + world, nv)
    root, flat = gpu.pcs_open(gpu.Mle.upload(ev, is_ext), nv, pt)
    res = run_sharded(gpu, ev, is_ext, nv, pt, world)
    for r in range(world):
        sroot, sflat, _ = res[r]
        assert (sroot == root).all(), "rank %d: root differs" % r
        assert sflat.shape == flat.shape and (sflat == flat).all(), "rank %d: proof differs" % r
    assert O.pcs_verify(flat, root, nv, not is_ext, nv, pt, O.evaluate(ev, is_ext, pt)) is None


def test_sharded_baseline_size_world8(gpu):
    """nu = 20 over 8 ranks: slices of 2^17 evaluations per rank (multi-block kernels, wide Merkle levels, 13 rounds)"""
    nv, world = 20, 8
    ev = O.splitmix_f(5, 1 << nv)
    pt = O.splitmix_e(6, nv)
    root, flat = gpu.pcs_open(gpu.Mle.upload(ev, False), nv, pt)
    res = run_sharded(gpu, ev, False, nv, pt, world)
    for r in range(world):
        assert (res[r][0] == root).all() and res[r][1].shape == flat.shape and (res[r][1] == flat).all(), "rank %d differs" % r
