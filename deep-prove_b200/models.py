"""Synthetic quantised models in the padded (power-of-two) layout the prover works in, as plain arrays: the layer
descriptor consumed by dph_model_context_new (host/capi.cpp) -- 9 int64 per node {kind, 8 shape words} plus the weights
in node order.  Pure numpy; both bench arms are fed the same arrays.

CNN-264k (SURVEY.md 8(d) Cfg 3): zkml/assets/scripts/CNN/cifar-cnn.py:175-241 with --num-params 264000 gives
c1 = 12, c2 = 33, fc1 = 247, fc2 = 173:  [3,32,32] -> conv5x5(12) -> requant -> relu -> maxpool -> conv5x5(33) -> requant
-> relu -> maxpool -> flatten(33*5*5) -> fc 247 -> requant -> relu -> fc 173 -> requant -> relu -> fc 10."""
import numpy as np

DENSE, REQUANT, RELU, CONV, POOL, MATMUL = 0, 1, 2, 3, 4, 5
BIT_LEN = 8


def _p2(x):
    p = 1
    while p < x:
        p <<= 1
    return p


def _clog2(x):
    return int(x - 1).bit_length() if x > 1 else 0


def _requant(int_part_log, intermediate_bits):
    """layers/requant.rs:395-410 shape: shift = fp_scale + right_shift is a multiple of BIT_LEN, multiplier = 0.75 * 2^fp_scale"""
    fp = ((int_part_log + 24 + 7) // 8) * 8 - int_part_log
    return [REQUANT, int_part_log, fp, 3 << (fp - 2), intermediate_bits, 0, 0, 0, 0]


def cnn(seed=1, img=32, c0=3, c1=12, c2=33, f1=247, f2=173, f3=10, k=5):
    """returns (desc [n_nodes, 9] int64, data int64, input int64 [padded c0 * img^2], n_real_params)"""
    rng = np.random.Generator(np.random.PCG64(seed))
    desc, data, n_params = [], [], 0
    kx_u, n_u, n_p = c0, img, _p2(img)
    x = np.zeros((_p2(c0), n_p, n_p), dtype=np.int64)
    x[:c0, :img, :img] = rng.integers(-127, 128, size=(c0, img, img))
    for kw_u in (c1, c2):
        kx, kw, rn = _p2(kx_u), _p2(kw_u), _p2(k)
        filt = np.zeros((kw, kx, rn, rn), dtype=np.int64)
        filt[:kw_u, :kx_u, :k, :k] = rng.integers(-127, 128, size=(kw_u, kx_u, k, k))
        bias = np.zeros(kw, dtype=np.int64)
        bias[:kw_u] = rng.integers(-1000, 1001, size=kw_u)
        n_params += kw_u * kx_u * k * k + kw_u
        h_out = n_u - k + 1
        desc.append([CONV, kw, kx, n_p, rn, kw_u, h_out, h_out, 0])
        data += [filt.reshape(-1), bias]
        terms = k * k * kx_u
        desc.append(_requant(_clog2(terms) + 2, 2 * (BIT_LEN - 1) + _clog2(terms + 1)))   # convolution.rs:362-366 output_bitsize
        desc.append([RELU] + [0] * 8)
        desc.append([POOL, kw, n_p, n_p, 0, 0, 0, 0, 0])
        kx_u, n_u, n_p = kw_u, h_out // 2, n_p // 2
    ncols, prev_real = _p2(c2) * n_p * n_p, None
    for li, out_u in enumerate((f1, f2, f3)):
        nrows = _p2(out_u)
        w = np.zeros((nrows, ncols), dtype=np.int64)
        if li == 0:   # flatten of [C][n_p][n_p]: real entries are c < c2, row < n_u, col < n_u
            mask = np.zeros((_p2(c2), n_p, n_p), dtype=bool)
            mask[:c2, :n_u, :n_u] = True
            real = mask.reshape(-1)
        else:
            real = np.arange(ncols) < prev_real
        w[:out_u, real] = rng.integers(-127, 128, size=(out_u, int(real.sum())))
        b = np.zeros(nrows, dtype=np.int64)
        b[:out_u] = rng.integers(-127, 128, size=out_u)
        n_params += out_u * int(real.sum()) + out_u
        desc.append([DENSE, nrows, ncols, 0, 0, 0, 0, 0, 0])
        data += [w.reshape(-1), b]
        if li < 2:
            desc.append(_requant(_clog2(ncols), 2 * (BIT_LEN - 1) + _clog2(ncols) + 1))   # dense.rs:416-421
            desc.append([RELU] + [0] * 8)
        ncols, prev_real = nrows, out_u
    return np.asarray(desc, dtype=np.int64), np.concatenate(data).astype(np.int64), x.reshape(-1), n_params


def cnn_small(seed=1):
    return cnn(seed, img=16, c0=3, c1=4, c2=6, f1=24, f2=16, f3=10, k=3)


def token_mlp(seed=1, tokens=8, d_model=64, d_hidden=128, transposed=(False, True), bias=(True, False)):
    """A token-wise MLP as the transformer feed-forward block uses it (zkml/src/layers/matrix_mul.rs: OperandMatrix::Input x
    OperandMatrix::Weight): X [tokens, d_model] -> MatMul(W1 [d_model, d_hidden] or its transpose, + bias) -> requant -> relu
    -> MatMul(W2 [d_hidden, d_model]) .  All dimensions powers of two (the padded layout).  Returns (desc, data, input)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    desc, data = [], []
    x = rng.integers(-127, 128, size=(tokens, d_model)).astype(np.int64)
    dims = [(d_model, d_hidden), (d_hidden, d_model)]
    for li, (k, c) in enumerate(dims):
        w = rng.integers(-127, 128, size=(k, c)).astype(np.int64)
        stored = w.T.copy() if transposed[li] else w          # Config::TransposeB keeps the matrix as [C][K]
        desc.append([MATMUL, tokens, k, c, int(transposed[li]), int(bias[li]), 0, 0, 0])
        data.append(stored.reshape(-1))
        if bias[li]:
            data.append(rng.integers(-127, 128, size=c).astype(np.int64))
        if li == 0:
            desc.append(_requant(_clog2(k), 2 * (BIT_LEN - 1) + _clog2(k) + 1))
            desc.append([RELU] + [0] * 8)
    return np.asarray(desc, dtype=np.int64), np.concatenate(data).astype(np.int64), x.reshape(-1)
