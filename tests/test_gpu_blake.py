"""GPU parity (-m gpu) under the reference's `blake` feature: BlakeHasher Merkle trees on the device + BlakeTranscript on the host
(the PINNED variant, see tests/test_blake.py): roots, opening proofs, batch openings, batch commitments and whole model proofs equal
the CPU checker's word for word, and the restated verifiers accept them."""
import numpy as np
import pytest
import oracle_py as O

pytestmark = pytest.mark.gpu


@pytest.fixture()
def blake(gpu):
    O.set_hash_mode(1); gpu.set_hasher(1)
    yield gpu
    O.set_hash_mode(0); gpu.set_hasher(0)


def rnd_poly(seed, nv, ext):
    return O.splitmix_e(seed, 1 << nv) if ext else O.splitmix_f(seed, 1 << nv)


@pytest.mark.parametrize("nv,full_log,ext", [(1, 8, False), (3, 8, True), (7, 10, False), (8, 8, False), (10, 10, True), (12, 14, False), (16, 16, False), (18, 20, False)])
def test_commit_roots(blake, nv, full_log, ext):
    ev = rnd_poly(100 + nv, nv, ext)
    root, cw, bh = O.pcs_commit(ev, ext, full_log)
    c = blake.Commitment(blake.Mle.upload(ev, ext), full_log)
    assert (c.codeword() == cw).all() and (c.root == root).all()


@pytest.mark.parametrize("nv,full_log,ext", [(8, 8, False), (10, 10, True), (12, 12, False), (14, 16, False), (15, 15, True), (20, 20, False)])
def test_open(blake, nv, full_log, ext):
    ev = rnd_poly(200 + nv, nv, ext)
    pt = O.splitmix_e(300 + nv, nv)
    exp = O.pcs_open(ev, ext, full_log, pt)
    root, got = blake.pcs_open(blake.Mle.upload(ev, ext), full_log, pt)
    assert got.shape == exp.shape and (got == exp).all()
    assert O.pcs_verify(got, root, nv, not ext, full_log, pt, O.evaluate(ev, ext, pt)) is None


def test_batch_open_and_simple_batch(blake):
    shape, full_log = [(9, False), (12, False), (12, True), (8, False)], 13
    polys = [(rnd_poly(400 + 7 * i + nv, nv, ext), ext) for i, (nv, ext) in enumerate(shape)]
    pts = [O.splitmix_e(500 + i, nv) for i, (nv, _) in enumerate(shape)]
    exp = O.pcs_batch_open(polys, full_log, pts)
    got = blake.pcs_batch_open([blake.Mle.upload(a, e) for a, e in polys], full_log, pts)
    assert got.shape == exp.shape and (got == exp).all()
    n_polys, nv, ext = 5, 10, False
    ps = [rnd_poly(900 + i, nv, ext) for i in range(n_polys)]
    pt = O.splitmix_e(950, nv)
    eroot, evals, e2 = O.pcs_simple_batch(ps, ext, 10, pt)
    root, g2 = blake.pcs_simple_batch([blake.Mle.upload(p, ext) for p in ps], 10, pt, evals)
    assert (root == eroot).all() and g2.shape == e2.shape and (g2 == e2).all()
    assert O.pcs_simple_batch_verify(g2, root, nv, not ext, n_polys, 10, pt, evals) is None


def test_sumcheck_with_blake_transcript(blake):
    nv = 12
    mles = [(O.splitmix_f(1, 1 << nv), False), (O.splitmix_e(2, 1 << nv), True), (O.splitmix_f(3, 1 << nv), False)]
    products = [((3, 5), [0, 1, 2]), ((7, 0), [1, 2])]
    got = blake.sumcheck_prove_parallel([blake.Mle.upload(a, e) for a, e in mles], products, nv)
    exp = O.sumcheck_prove(mles, products, nv)
    for g, e in zip(got, exp):
        assert (np.asarray(g) == e).all()


def test_model_proofs(blake):
    w, b, rq = O.synthetic_mlp(2, 64, 5)
    x = O.synthetic_input(64, 6)
    exp, _ = O.zkml_prove(2, 64, 5, 6)
    got = blake.ZkmlContext(2, 64, w, b, rq).prove(x)
    assert got.shape == exp.shape and (got == exp).all()
    assert O.zkml_prove_verify(2, 64, 5, 6) is None
    import models
    desc, data, xin, _ = models.cnn_small(seed=3)
    got = blake.ModelContext(desc, data, xin.size).prove(xin)
    exp, _ = O.model_prove(desc, data, xin)
    assert got.shape == exp.shape and (got == exp).all()


def test_hasher_switch_back_to_poseidon(gpu):
    """the default pair is untouched after a blake session"""
    assert gpu.lib().dp_get_merkle_hasher() == 0
    ev = rnd_poly(1, 9, False)
    assert (gpu.Commitment(gpu.Mle.upload(ev, False), 10).root == O.pcs_commit(ev, False, 10, want_codeword=False)[0]).all()
