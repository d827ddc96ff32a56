"""Device outputs against the committed regression vectors (tests/golden/checker_regression.json): the same seeded inputs
as tests/golden/make_golden.py, through the C ABI / host mirror."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

import oracle_py as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "checker_regression.json")))


def h(*arrays):
    m = hashlib.sha256()
    for a in arrays:
        m.update(np.ascontiguousarray(a, dtype=np.uint64).tobytes())
    return m.hexdigest()


def test_device_matches_committed_vectors(gpu):
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "deep-prove_b200"))
    import models
    nv = 8
    mles = [(O.splitmix_f(1, 1 << nv), False), (O.splitmix_e(2, 1 << nv), True), (O.splitmix_f(3, 1 << nv), False)]
    got = gpu.sumcheck_prove_parallel([gpu.Mle.upload(a, e) for a, e in mles], [((1, 0), [0, 1, 2]), ((5, 7), [1, 2])], nv)
    assert h(*got) == GOLD["sumcheck_nv8_deg3"]
    ev = O.splitmix_f(7, 1 << 10)
    root, flat = gpu.pcs_open(gpu.Mle.upload(ev, False), 10, O.splitmix_e(8, 10))
    assert [int(x) for x in root] == GOLD["basefold_commit_root_nv10"]
    assert h(flat) == GOLD["basefold_open_nv10"]
    polys = [(O.splitmix_f(400 + i, 1 << nvp), False) for i, nvp in enumerate((10, 8, 9))]
    flat = gpu.pcs_batch_open([gpu.Mle.upload(a, e) for a, e in polys], 10, [O.splitmix_e(500 + i, nvp) for i, nvp in enumerate((10, 8, 9))])
    assert h(flat) == GOLD["basefold_batch_open_10_8_9"]
    sb = [O.splitmix_f(900 + i, 1 << 9) for i in range(3)]
    pt = O.splitmix_e(950, 9)
    evals = np.array([O.evaluate(p, False, pt) for p in sb])
    r, flat = gpu.pcs_simple_batch([gpu.Mle.upload(p, False) for p in sb], 10, pt, evals)
    assert [int(x) for x in r] == GOLD["simple_batch_root_3x_nv9"] and h(flat) == GOLD["simple_batch_open_3x_nv9"]
    w, b, rq = O.synthetic_mlp(2, 64, 5)
    assert h(gpu.ZkmlContext(2, 64, w, b, rq).prove(O.synthetic_input(64, 6))) == GOLD["zkml_mlp_2x64"]
    filt, bias, x, uo = O.synthetic_conv(2, 2, 8, 4, 2, 3, 2, 7, 15, 16)
    assert h(gpu.conv_prove(filt, bias, uo, x)[2]) == GOLD["conv_layer_proof_2x2x8"]
    d, wts, xin, _ = models.cnn_small(seed=3)
    assert h(gpu.ModelContext(d, wts, xin.size).prove(xin)) == GOLD["cnn_small_proof"]
