"""Full zkml proof of a CNN (conv -> requant -> relu -> maxpool, twice, then three dense layers) on the device against
the CPU checker: identical flat proof (every layer proof, table proofs and the batched Basefold opening).  The large
case is SURVEY.md 8(d) Cfg 3, CNN-264k in its padded form."""
import numpy as np
import pytest

import oracle_py as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("small", [1, 0])
def test_cnn_full_proof(gpu, small):
    desc, data, x = O.synthetic_cnn(small, 3, 4)
    ctx = gpu.ModelContext(desc, data, x.size)
    got = ctx.prove(x)
    exp, _ = O.cnn_prove(small, 3, 4)
    assert got.shape == exp.shape
    if not (got == exp).all():
        first = int(np.argmax(got != exp))
        raise AssertionError("first difference at word %d of %d" % (first, exp.size))
    # a stored trace can be proved again (resident tensors are borrowed, never modified)
    ctx.run_inference(x)
    assert (ctx.prove_trace(want_proof=True) == exp).all()
    assert (ctx.prove_trace(want_proof=True) == exp).all()
    ctx.free()


def test_cnn264k_from_descriptor(gpu):
    """the bench's CNN-264k arrays (deep-prove_b200/models.py) through both arms"""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "deep-prove_b200"))
    import models
    desc, data, x, n_params = models.cnn(seed=7)
    assert 250_000 < n_params < 270_000
    ctx = gpu.ModelContext(desc, data, x.size)
    got = ctx.prove(x)
    exp, _ = O.model_prove(desc, data, x)
    assert got.shape == exp.shape and (got == exp).all()
    ctx.free()
