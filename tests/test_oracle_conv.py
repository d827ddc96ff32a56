"""CPU checker for the FFT-convolution layer (oracle/conv.hpp), pinned the way the reference pins it:
fft == naive DFT and ifft(fft) = id (tensor.rs fft is exercised by test_conv_fft_vs_naive, convolution.rs:1811),
FFT convolution == direct convolution on the unpadded region (same test), clearing (convolution.rs:1634-1647), and the
layer proof's own debug_assert invariants (convolution.rs:734-747,792-806,827-852,932-975,1035-1060), which the restatement
keeps as hard checks and which fire inside dpo_conv_prove."""
import numpy as np
import pytest

import oracle_py as O

P = 0xFFFFFFFF00000001
ROOT32 = 1753635133440165772


def test_fft_is_the_dft_with_the_two_adic_root():
    n, logn = 16, 4
    v = O.splitmix_e(7, 2 * n).reshape(2, n, 2)
    got = O.fft_ext(v)
    w = pow(ROOT32, 1 << (32 - logn), P)
    assert pow(w, n, P) == 1 and pow(w, n // 2, P) == P - 1
    for r in range(2):
        for k in range(n):
            for c in range(2):
                acc = sum(int(v[r, j, c]) * pow(w, j * k, P) for j in range(n)) % P
                assert acc == int(got[r, k, c])
    assert (O.fft_ext(got, inverse=True) == v).all()


@pytest.mark.parametrize("kw,kx,n_x,rn,kw_u,k_u,kx_u,n_x_u", [(2, 2, 8, 4, 2, 3, 2, 7), (4, 2, 8, 2, 3, 2, 1, 8), (4, 4, 16, 8, 3, 5, 3, 14)])
def test_fft_conv_equals_direct_convolution(kw, kx, n_x, rn, kw_u, k_u, kx_u, n_x_u):
    filt, bias, x, uo = O.synthetic_conv(kw, kx, n_x, rn, kw_u, k_u, kx_u, n_x_u, 5, 6)
    after, cleared = O.conv_op(filt, bias, uo, x)
    h = n_x_u - k_u + 1
    for i in range(kw_u):
        for y in range(h):
            for xx in range(h):
                acc = int(bias[i]) + sum(int(filt[i, j, a, b]) * int(x[j, y + a, xx + b]) for j in range(kx) for a in range(k_u) for b in range(k_u))
                assert acc == int(after[i, y, xx]) == int(cleared[i, y, xx])
    mask = np.zeros_like(cleared, dtype=bool)
    mask[:kw_u, :h, :h] = True
    assert (cleared[~mask] == 0).all() and (cleared[mask] == after[mask]).all()


@pytest.mark.parametrize("kw,kx,n_x,rn,kw_u,k_u,kx_u,n_x_u", [(2, 2, 8, 4, 2, 3, 2, 7), (4, 2, 8, 2, 3, 2, 2, 8), (2, 4, 16, 8, 2, 5, 3, 14), (1, 2, 8, 4, 1, 3, 2, 8)])
def test_conv_layer_proof_invariants_hold(kw, kx, n_x, rn, kw_u, k_u, kx_u, n_x_u):
    filt, bias, x, uo = O.synthetic_conv(kw, kx, n_x, rn, kw_u, k_u, kx_u, n_x_u, 15, 16)
    flat = O.conv_prove(filt, bias, uo, x)      # raises if any of the reference's invariants fails
    assert flat.size > 100
    assert (flat == O.conv_prove(filt, bias, uo, x)).all()
    assert not (flat == O.conv_prove(filt, bias, uo, x, label=b"other"))[: flat.size // 2].all()


def test_small_cnn_full_proof_passes_every_invariant():
    """conv -> requant -> relu -> maxpool (x2) -> 3 dense layers, full Prover::prove on the CPU checker: the conv
    invariants, the pooling input-claim identity (pooling.rs:453-490), requant recombination, the LogUp fractional-sum
    cancellation across lookups and tables, the model-input claim and Basefold's batch sanity checks all hold."""
    flat, _ = O.cnn_prove(1, 3, 4)
    assert flat.size > 10000
    again, _ = O.cnn_prove(1, 3, 4)
    assert (flat == again).all()
    desc, data, inp = O.synthetic_cnn(1, 3, 4)
    kinds = list(desc[:, 0])
    assert kinds == [3, 1, 2, 4, 3, 1, 2, 4, 0, 1, 2, 0, 1, 2, 0]


def test_numpy_model_descriptor_proves_on_the_checker():
    """deep-prove_b200/models.py (the arrays bench.py feeds to both arms): the small CNN proves and passes every invariant;
    the CNN-264k descriptor has the reference script's layer sizes (cifar-cnn.py:175-241 with --num-params 264000)"""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "deep-prove_b200"))
    import models
    desc, data, x, n = models.cnn_small(seed=3)
    flat, _ = O.model_prove(desc, data, x)
    assert flat.size > 10000 and (flat == O.model_prove(desc, data, x)[0]).all()
    d, w, xin, n_params = models.cnn(seed=1)
    assert [int(k) for k in d[:, 0]] == [3, 1, 2, 4, 3, 1, 2, 4, 0, 1, 2, 0, 1, 2, 0]
    assert tuple(d[0, 1:8]) == (16, 4, 32, 8, 12, 28, 28) and tuple(d[4, 1:8]) == (64, 16, 16, 8, 33, 10, 10)
    assert tuple(d[8, 1:3]) == (256, 4096) and tuple(d[11, 1:3]) == (256, 256) and tuple(d[14, 1:3]) == (16, 256)
    assert n_params == 12 * (75 + 1) + 33 * (12 * 25 + 1) + 247 * (825 + 1) + 173 * (247 + 1) + 10 * (173 + 1)
