"""FFT-convolution layer on the device (csrc/conv.cu + host/conv.hpp) against the CPU checker: inference tensors and the
whole layer proof (clearing hadamard, batch iFFT/FFT with matrix delegation, hadamard product, FFT of the weights), same
transcript -> identical flat proof (convolution.rs:697-1077)."""
import numpy as np
import pytest

import oracle_py as O

pytestmark = pytest.mark.gpu

SHAPES = [(2, 2, 8, 4, 2, 3, 2, 7), (4, 2, 8, 2, 3, 2, 2, 8), (2, 4, 16, 8, 2, 5, 3, 14), (1, 2, 8, 4, 1, 3, 2, 8), (1, 1, 8, 4, 1, 3, 1, 8),
          (16, 4, 32, 8, 12, 5, 3, 32), (64, 16, 16, 8, 33, 5, 12, 14)]   # the last two are CNN-264k's conv1 / conv2 after padding


@pytest.mark.parametrize("n,log_n", [(8, 3), (64, 6), (2048, 11), (8192, 13)])
def test_fft_rows(gpu, n, log_n):
    rows = 3 if n < 8192 else 2
    rows_p2 = 4 if rows == 3 else 2
    v = O.splitmix_e(40 + log_n, rows_p2 * n).reshape(rows_p2, n, 2)
    m = gpu.Mle.upload(v.reshape(-1, 2), True)
    gpu.check(gpu.lib().dp_fft_rows(m.h, log_n, 0))
    got = m.download().reshape(rows_p2, n, 2)
    assert (got == O.fft_ext(v)).all()
    gpu.check(gpu.lib().dp_fft_rows(m.h, log_n, 1))
    assert (m.download().reshape(rows_p2, n, 2) == v).all()


@pytest.mark.parametrize("kw,kx,n_x,rn,kw_u,k_u,kx_u,n_x_u", SHAPES)
def test_conv_inference(gpu, kw, kx, n_x, rn, kw_u, k_u, kx_u, n_x_u):
    filt, bias, x, uo = O.synthetic_conv(kw, kx, n_x, rn, kw_u, k_u, kx_u, n_x_u, 25, 26)
    after, cleared, _ = gpu.conv_prove(filt, bias, uo, x, prove=False)
    ea, ec = O.conv_op(filt, bias, uo, x)
    assert (after == ea).all() and (cleared == ec).all()


@pytest.mark.parametrize("kw,kx,n_x,rn,kw_u,k_u,kx_u,n_x_u", SHAPES)
def test_conv_layer_proof(gpu, kw, kx, n_x, rn, kw_u, k_u, kx_u, n_x_u):
    filt, bias, x, uo = O.synthetic_conv(kw, kx, n_x, rn, kw_u, k_u, kx_u, n_x_u, 35, 36)
    _, _, got = gpu.conv_prove(filt, bias, uo, x)
    exp = O.conv_prove(filt, bias, uo, x)
    assert got.shape == exp.shape
    assert (got == exp).all()
