"""MatMul layer on the device (host/zkml.hpp prove_matmul + csrc/witness.cu k_wit_matmul; zkml/src/layers/matrix_mul.rs:701-874): the
full proof of a token-wise MLP -- MatMul (+ bias) -> requant -> relu -> MatMul, with the constant matrices stored plain or transposed
(Config::TransposeB) -- must be word for word the CPU checker's, from the host input (device inference + witness generation) and
from the stored trace."""
import os
import sys

import numpy as np
import pytest

import oracle_py as O

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "deep-prove_b200"))
import models  # noqa: E402


@pytest.mark.parametrize("tokens,d_model,d_hidden,transposed,bias", [
    (8, 32, 64, (False, True), (True, False)),
    (8, 32, 64, (True, False), (False, True)),
    (4, 64, 64, (True, True), (True, True)),
    (64, 256, 512, (False, False), (True, True)),        # 2^14 / 2^15-entry activations, 2^17-entry matrices: multi-block kernels
])
def test_matmul_mlp_full_proof(gpu, tokens, d_model, d_hidden, transposed, bias):
    desc, data, x = models.token_mlp(seed=11, tokens=tokens, d_model=d_model, d_hidden=d_hidden, transposed=transposed, bias=bias)
    ctx = gpu.ModelContext(desc, data, x.size)
    got = ctx.prove(x)
    exp, _ = O.model_prove(desc, data, x)
    assert got.shape == exp.shape
    if not (got == exp).all():
        raise AssertionError("first difference at word %d of %d" % (int(np.argmax(got != exp)), exp.size))
    ctx.run_inference(x)
    assert (ctx.prove_trace(want_proof=True) == exp).all()
    ctx.free()
