"""commit -> open -> verify accepts (the reference's own test, mpcs/src/basefold.rs:1239-1331), with the verifier restated
from basefold.rs:863-1098 and query_phase.rs:141-283,915-975,1116-1236; tampered proofs / wrong claims are rejected."""
import numpy as np
import pytest

import oracle_py as O


@pytest.mark.parametrize("nv,full_log,ext", [(8, 8, False), (9, 11, False), (10, 10, True), (12, 13, False), (5, 8, False)])
def test_open_verifies(nv, full_log, ext):
    ev = O.splitmix_e(500 + nv, 1 << nv) if ext else O.splitmix_f(500 + nv, 1 << nv)
    pt = O.splitmix_e(600 + nv, nv)
    root, _, _ = O.pcs_commit(ev, ext, full_log)
    value = O.evaluate(ev, ext, pt)
    if nv > 7:
        flat = O.pcs_open(ev, ext, full_log, pt)
        assert O.pcs_verify(flat, root, nv, not ext, full_log, pt, value) is None
        bad = value.copy(); bad[0] ^= np.uint64(1)
        assert "p0(0) + p0(1)" in O.pcs_verify(flat, root, nv, not ext, full_log, pt, bad)
        tam = flat.copy(); tam[flat.size // 2] ^= np.uint64(1)          # somewhere in the query openings
        assert O.pcs_verify(tam, root, nv, not ext, full_log, pt, value) is not None
        other_root = root.copy(); other_root[0] ^= np.uint64(1)
        assert "merkle" in O.pcs_verify(flat, other_root, nv, not ext, full_log, pt, value)
        assert O.pcs_verify(flat, root, nv, not ext, full_log, pt, value, label=b"other") is not None   # different transcript


@pytest.mark.parametrize("shape,full_log", [([(10, False)], 10), ([(10, False), (8, False), (9, True)], 10), ([(9, False), (12, False), (12, True), (8, False)], 13)])
def test_batch_open_verifies(shape, full_log):
    polys = [(O.splitmix_e(700 + i, 1 << nv) if ext else O.splitmix_f(700 + i, 1 << nv), ext) for i, (nv, ext) in enumerate(shape)]
    points = [O.splitmix_e(800 + i, nv) for i, (nv, _) in enumerate(shape)]
    flat = O.pcs_batch_open(polys, full_log, points)
    roots = [O.pcs_commit(p, e, full_log, want_codeword=False)[0] for p, e in polys]
    evals = np.array([O.evaluate(p, e, pt) for (p, e), pt in zip(polys, points)])
    nvs = [nv for nv, _ in shape]; isb = [not e for _, e in shape]
    assert O.pcs_batch_verify(flat, np.array(roots), nvs, isb, full_log, points, evals) is None
    bad = evals.copy(); bad[-1, 1] ^= np.uint64(1)
    assert "classic sumcheck" in O.pcs_batch_verify(flat, np.array(roots), nvs, isb, full_log, points, bad)
    tam = flat.copy(); tam[-3] ^= np.uint64(1)
    assert O.pcs_batch_verify(tam, np.array(roots), nvs, isb, full_log, points, evals) is not None


@pytest.mark.parametrize("nl,w", [(2, 64), (1, 256), (3, 128)])
def test_model_proof_is_accepted_by_the_restated_verifier(nl, w):
    """Verifier::verify (zkml/src/iop/verifier.rs:72-296) restated in oracle/zk_verify.hpp re-derives every challenge on its
    own transcript: acceptance pins the prover restatement's Fiat-Shamir order and claim chaining (dense, requant, relu,
    table proofs, commitment openings), which device-vs-checker equality alone cannot; forged proofs are rejected."""
    assert O.zkml_prove_verify(nl, w, 5, 6) is None
    assert O.zkml_prove_verify(nl, w, 5, 6, tamper=1) is not None      # wrong public output
    assert "dense" in O.zkml_prove_verify(nl, w, 5, 6, tamper=2)        # forged claim
    assert O.zkml_prove_verify(nl, w, 5, 6, tamper=3) is not None      # forged lookup fraction


@pytest.mark.parametrize("which", ["small", "cnn264k"])
def test_cnn_proof_is_accepted_by_the_restated_verifier(which):
    """the CNN path through the restated verifier: hadamard::verify, ConvCtx::verify_convolution + verify_fft_delegation
    (convolution.rs:1090-1375), PoolingCtx::verify_pooling (pooling.rs:525-650) on top of the MLP pieces"""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "deep-prove_b200"))
    import models
    desc, data, x, _ = models.cnn_small(seed=3) if which == "small" else models.cnn(seed=1)
    assert O.model_prove_verify(desc, data, x) is None
    if which == "small":
        assert O.model_prove_verify(desc, data, x, tamper=1) is not None
        assert "padded_fft" in O.model_prove_verify(desc, data, x, tamper=2)
        assert "pooling" in O.model_prove_verify(desc, data, x, tamper=3)


@pytest.mark.parametrize("n_polys,nv,full_log,ext", [(3, 9, 10, False), (5, 10, 10, False), (2, 9, 9, True), (1, 9, 9, False), (4, 8, 12, False)])
def test_batch_commit_simple_batch_open_verifies(n_polys, nv, full_log, ext):
    """batch_commit -> simple_batch_open -> simple_batch_verify accepts (the reference's batch_commit_open_verify,
    mpcs/src/basefold.rs:1300-1331); one polynomial under batch_commit has the plain commitment's root"""
    polys = [O.splitmix_e(900 + i, 1 << nv) if ext else O.splitmix_f(900 + i, 1 << nv) for i in range(n_polys)]
    pt = O.splitmix_e(950, nv)
    root, evals, flat = O.pcs_simple_batch(polys, ext, full_log, pt)
    assert all((evals[i] == O.evaluate(polys[i], ext, pt)).all() for i in range(n_polys))
    if n_polys == 1:
        assert (root == O.pcs_commit(polys[0], ext, full_log, want_codeword=False)[0]).all()
    assert O.pcs_simple_batch_verify(flat, root, nv, not ext, n_polys, full_log, pt, evals) is None
    bad = evals.copy(); bad[0, 0] ^= np.uint64(1)
    assert O.pcs_simple_batch_verify(flat, root, nv, not ext, n_polys, full_log, pt, bad) is not None
    tam = flat.copy(); tam[flat.size // 2] ^= np.uint64(1)
    assert O.pcs_simple_batch_verify(tam, root, nv, not ext, n_polys, full_log, pt, evals) is not None


def test_batch_open_with_shared_points_verifies():
    """several polynomials opened at the same point (basefold.rs:617-640 merges them) and one polynomial at two points: the
    restated batch_verify accepts"""
    shape = [(10, False), (10, False), (9, True), (10, False)]
    polys = [(O.splitmix_e(1000 + i, 1 << nv) if ext else O.splitmix_f(1000 + i, 1 << nv), ext) for i, (nv, ext) in enumerate(shape)]
    points = [O.splitmix_e(1100, 10), O.splitmix_e(1101, 9), O.splitmix_e(1102, 10)]
    eval_poly, eval_point = [0, 1, 2, 3, 0], [0, 0, 1, 2, 2]
    flat, roots, vals = O.pcs_batch_open_evals(polys, 11, points, eval_poly, eval_point)   # raises unless batch_verify accepts
    assert flat.size > 1000
