// FFT-convolution layer: witness tensors and the helper tables of its proof, on the device.
// Reference: zkml/src/tensor.rs:220-323 (get_root_of_unity, index_w, index_u, fft), :458-523 (fft_conv);
// zkml/src/layers/convolution.rs:152-162 (add_bias), :870 (beta_acc), :1535-1550 (index_wf);
// zkml/src/iop/prover.rs:164-289 (delegate_matrix_evaluation's phi vectors, phi_pow_init, phi_g_init).
// Everything here is small (rows of 2*n_x^2 <= 2^13 Ext values); the kernels exist so that the tensors a conv proof
// needs (input_fft, prod, the FFT-matrix tables) are produced where the sumchecks consume them -- no host round trips.
#include "common.cuh"
#include "powtab.cuh"
#include "../../include/deepprove_b200.h"

static inline int new_mle(u64 len, bool ext, dp_mle **out) {
    dp_mle *m = new dp_mle(); m->len = len; m->is_ext = ext; m->owned = true;
    if (int e = dp_dev_alloc(&m->data, m->bytes())) { delete m; return e; }
    *out = m; return DP_OK;
}

// ---- fft(v, flag) on rows (tensor.rs:261-323): natural order in/out, w = get_root_of_unity(log n) (inverse when flag) ----
__global__ void __launch_bounds__(256) k_fft_rows(gle *data, u32 lg, int inverse, PowTab tab) {
    extern __shared__ unsigned char smem_raw[];
    gle *sm = reinterpret_cast<gle *>(smem_raw);
    const u32 n = 1u << lg;
    gle *row = data + (u64)blockIdx.x * n;
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) sm[lg ? (__brev(i) >> (32 - lg)) : 0] = row[i];
    __syncthreads();
    for (u32 s = 1; s <= lg; s++) {
        const u32 half = 1u << (s - 1);
        for (u32 q = threadIdx.x; q < (n >> 1); q += blockDim.x) {
            u32 k = q & (half - 1), c = (q >> (s - 1)) << s;
            u32 e = (n >> s) * k;                                       // w[n / i * k], i = 2^s
            if (inverse && e) e = n - e;
            u64 w = tab_pow(tab, (u64)e << (32 - lg));
            gle u = sm[c + k], l = e_mul_base(sm[c + k + half], w);
            sm[c + k] = e_add(u, l); sm[c + k + half] = e_sub(u, l);
        }
        __syncthreads();
    }
    u64 ilen = inverse ? gl_inv((u64)n) : 1ULL;
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) row[i] = inverse ? e_mul_base(sm[i], ilen) : sm[i];
}
extern "C" int dp_fft_rows(dp_mle *m, uint32_t log_n, int inverse) {
    DP_REQUIRE_CTX();
    DP_CHECK(m && m->is_ext, DP_ERR_INVALID, "dp_fft_rows: need an Ext MLE");
    DP_CHECK(log_n >= 1 && log_n <= 13 && (m->len >> log_n) >= 1 && (m->len & ((1ULL << log_n) - 1)) == 0, DP_ERR_INVALID, "dp_fft_rows: row length must be 2^1..2^13 and divide the MLE length");
    PowTab tab; if (int e = dp_root_powtab(&tab)) return e;
    size_t smem = sizeof(gle) << log_n;
    if (smem > 48 * 1024) DP_CUDA(cudaFuncSetAttribute(k_fft_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DpProfScope prof("k_fft_rows", m->bytes() * 2);
    k_fft_rows<<<(unsigned)(m->len >> log_n), 256, smem, dp_ctx().stream>>>((gle *)m->data, log_n, inverse, tab); DP_LAUNCHED();
    DP_CUDA(cudaGetLastError());
    return DP_OK;
}

// ---- index_w / index_wf: an n_real x n_real block placed top-left in an n x n grid, row padded with zeros to out_len ----
template <bool EXT>
__global__ void k_pad_rows(const void *src, u64 rows, u32 n_real, u32 n, u64 out_len, gle *out) {
    u64 total = rows * out_len, t = (u64)blockIdx.x * blockDim.x + threadIdx.x, stride = (u64)gridDim.x * blockDim.x;
    for (; t < total; t += stride) {
        u64 r = t / out_len, idx = t % out_len, i = idx / n, j = idx % n;
        gle v = e_zero();
        if (i < n_real && j < n_real) { u64 s = r * n_real * n_real + i * n_real + j; v = EXT ? ((const gle *)src)[s] : e_from_base(((const u64 *)src)[s]); }
        out[t] = v;
    }
}
extern "C" int dp_pad_rows(const dp_mle *src, uint64_t rows, uint32_t n_real, uint32_t n, uint64_t out_len, dp_mle **out) {
    DP_REQUIRE_CTX();
    DP_CHECK(src && out && rows && n_real && n_real <= n, DP_ERR_INVALID, "dp_pad_rows: bad arguments");
    DP_CHECK(src->len == rows * n_real * n_real, DP_ERR_INVALID, "dp_pad_rows: source length != rows * n_real^2");
    u64 total = rows * out_len;
    DP_CHECK(total && (total & (total - 1)) == 0, DP_ERR_INVALID, "dp_pad_rows: rows * out_len must be a power of two");
    dp_mle *m; if (int e = new_mle(total, true, &m)) return e;
    int g = dp_grid_for(total, 256, 8);
    if (src->is_ext) k_pad_rows<true><<<g, 256, 0, dp_ctx().stream>>>(src->data, rows, n_real, n, out_len, (gle *)m->data);
    else k_pad_rows<false><<<g, 256, 0, dp_ctx().stream>>>(src->data, rows, n_real, n, out_len, (gle *)m->data);
    DP_LAUNCHED(); DP_CUDA(cudaGetLastError());
    *out = m; return DP_OK;
}

// ---- fft_conv accumulation (tensor.rs:489-509): out[i][k] = sum_j x_fft[j][k] * w_fft[i][j][k] ----
__global__ void k_conv_prod(const gle *__restrict__ x, const gle *__restrict__ w, u32 kw, u32 kx, u64 n, gle *__restrict__ out) {
    u64 total = (u64)kw * n, t = (u64)blockIdx.x * blockDim.x + threadIdx.x, stride = (u64)gridDim.x * blockDim.x;
    for (; t < total; t += stride) {
        u64 i = t / n, k = t % n;
        gle acc = e_zero();
        for (u32 j = 0; j < kx; j++) acc = e_add(acc, e_mul(x[(u64)j * n + k], w[(i * kx + j) * n + k]));
        out[t] = acc;
    }
}
extern "C" int dp_conv_prod(const dp_mle *x_fft, const dp_mle *w_fft, uint32_t kw, uint32_t kx, uint64_t row_len, dp_mle **out) {
    DP_REQUIRE_CTX();
    DP_CHECK(x_fft && w_fft && out && x_fft->is_ext && w_fft->is_ext, DP_ERR_INVALID, "dp_conv_prod: need Ext operands");
    DP_CHECK(x_fft->len == (u64)kx * row_len && w_fft->len == (u64)kw * kx * row_len, DP_ERR_INVALID, "dp_conv_prod: operand shapes do not match [kx, n] / [kw, kx, n]");
    dp_mle *m; if (int e = new_mle((u64)kw * row_len, true, &m)) return e;
    DpProfScope prof("k_conv_prod", (w_fft->bytes() + x_fft->bytes() + m->bytes()));
    k_conv_prod<<<dp_grid_for((u64)kw * row_len, 128, 8), 128, 0, dp_ctx().stream>>>((const gle *)x_fft->data, (const gle *)w_fft->data, kw, kx, row_len, (gle *)m->data); DP_LAUNCHED();
    DP_CUDA(cudaGetLastError());
    *out = m; return DP_OK;
}

// ---- index_u + to_element + add_bias: out[i][t] = to_element(rows[i][n_x^2 - 1 - t]) + bias[i] ----
__global__ void k_conv_out_elems(const gle *rows, u32 kw, u32 nx2, const long long *bias, long long *out) {
    u64 total = (u64)kw * nx2, t = (u64)blockIdx.x * blockDim.x + threadIdx.x, stride = (u64)gridDim.x * blockDim.x;
    for (; t < total; t += stride) {
        u64 i = t / nx2, k = t % nx2;
        u64 c = rows[i * 2 * nx2 + (nx2 - 1 - k)].c0;
        long long v = c <= (GL_P >> 1) ? (long long)c : -(long long)(GL_P - c);     // quantization/mod.rs:225-242
        out[t] = v + bias[i];
    }
}
extern "C" int dp_conv_output_elements(const dp_mle *out_rows, uint32_t kw, uint32_t n_x, const int64_t *bias, int64_t *out_host) {
    DP_REQUIRE_CTX();
    u64 nx2 = (u64)n_x * n_x;
    DP_CHECK(out_rows && out_rows->is_ext && bias && out_host && out_rows->len == 2 * nx2 * kw, DP_ERR_INVALID, "dp_conv_output_elements: need [kw, 2 n_x^2] Ext rows");
    long long *d = nullptr, *db = nullptr;
    if (int e = dp_dev_alloc((void **)&d, 8 * nx2 * kw)) return e;
    if (int e = dp_dev_alloc((void **)&db, 8 * (size_t)kw)) { dp_dev_free(d); return e; }
    cudaStream_t st = dp_ctx().stream;
    DP_CUDA(cudaMemcpyAsync(db, bias, 8 * (size_t)kw, cudaMemcpyHostToDevice, st));
    k_conv_out_elems<<<dp_grid_for(nx2 * kw, 256, 8), 256, 0, st>>>((const gle *)out_rows->data, kw, (u32)nx2, db, d); DP_LAUNCHED();
    DP_CUDA(cudaGetLastError());
    const int rc = dp_d2h(out_host, d, 8 * nx2 * kw, st);
    dp_dev_free(d); dp_dev_free(db);
    return rc;
}

// ---- phi_g_init (iop/prover.rs:231-289): the FFT / iFFT matrix row W(rx, .) and its per-level prefixes, one launch ----
struct PhiArgs { gle rx[16]; gle scale; gle *mid[16]; };
__global__ void __launch_bounds__(1024) k_phi_g_init(gle *phi_g, PhiArgs a, u32 n, int is_fft, PowTab tab) {
    const u32 tid = threadIdx.x;
    if (tid == 0) { phi_g[0] = a.scale; if (is_fft) phi_g[1] = a.scale; }
    __syncthreads();
    const u32 last = is_fft ? n : n - 1;
    for (u32 i = 1; i <= last; i++) {
        const u32 m = n - i, cnt = 1u << (i - 1);
        for (u32 b = tid; b < cnt; b += blockDim.x) {
            u32 e = b << m;                                               // phi_mul[b << m]: root (inverse when is_fft) of order 2^n
            if (is_fft && e) e = (1u << n) - e;
            u64 w = tab_pow(tab, (u64)e << (32 - n));
            gle rxm = a.rx[m], tmp1 = e_sub(e_one(), rxm), tmp2 = e_mul_base(rxm, w), pl = phi_g[b];
            phi_g[b ^ cnt] = e_mul(pl, e_sub(tmp1, tmp2));
            phi_g[b] = e_mul(pl, e_add(tmp1, tmp2));
        }
        __syncthreads();
        if (i < n) { gle *dst = a.mid[i - 1]; for (u32 k = tid; k < (1u << i); k += blockDim.x) dst[k] = phi_g[k]; }
        __syncthreads();
    }
    if (!is_fft) {
        gle rx0 = a.rx[0], one_m = e_sub(e_one(), rx0);
        for (u32 b = tid; b < (1u << (n - 1)); b += blockDim.x) {
            u64 w = tab_pow(tab, (u64)b << (32 - n));
            phi_g[b] = e_mul(phi_g[b], e_add(one_m, e_mul_base(rx0, w)));
        }
    }
}
extern "C" int dp_phi_g_init(const uint64_t *rx, uint32_t n, const uint64_t *scale, int is_fft, dp_mle **w_red, dp_mle **mid) {
    DP_REQUIRE_CTX();
    DP_CHECK(rx && scale && w_red && mid && n >= 2 && n <= 14, DP_ERR_INVALID, "dp_phi_g_init: need 2 <= n <= 14");
    PowTab tab; if (int e = dp_root_powtab(&tab)) return e;
    PhiArgs a; memset(&a, 0, sizeof a);
    for (u32 i = 0; i < n; i++) a.rx[i] = e_make(gl_canon(rx[2 * i]), gl_canon(rx[2 * i + 1]));
    a.scale = e_make(gl_canon(scale[0]), gl_canon(scale[1]));
    dp_mle *w; if (int e = new_mle(1ULL << n, true, &w)) return e;
    cudaStream_t st = dp_ctx().stream;
    DP_CUDA(cudaMemsetAsync(w->data, 0, w->bytes(), st));                // the reference's vec![E::ZERO; ..]: the FFT case leaves the top half zero
    for (u32 i = 0; i + 1 < n; i++) { if (int e = new_mle(2ULL << i, true, &mid[i])) return e; a.mid[i] = (gle *)mid[i]->data; }
    k_phi_g_init<<<1, 1024, 0, st>>>((gle *)w->data, a, n, is_fft, tab); DP_LAUNCHED();
    DP_CUDA(cudaGetLastError());
    *w_red = w; return DP_OK;
}

// ---- one level of delegate_matrix_evaluation (prover.rs:182-195): phi[i] = A + B * omega^(i << shift), omega of order 2^n_total ----
__global__ void k_phi_level(gle *out, u32 len, u32 n_total, u32 shift, gle A, gle B, int inverse, PowTab tab) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    u32 e = (i << shift) & ((1u << n_total) - 1);
    if (inverse && e) e = (1u << n_total) - e;
    out[i] = e_add(A, e_mul_base(B, tab_pow(tab, (u64)e << (32 - n_total))));
}
extern "C" int dp_phi_level(uint32_t len_log, uint32_t n_total, uint32_t shift, const uint64_t *A, const uint64_t *B, int inverse, dp_mle **out) {
    DP_REQUIRE_CTX();
    DP_CHECK(A && B && out && n_total >= 1 && n_total <= 32 && len_log + shift <= n_total, DP_ERR_INVALID, "dp_phi_level: bad arguments");
    PowTab tab; if (int e = dp_root_powtab(&tab)) return e;
    dp_mle *m; if (int e = new_mle(1ULL << len_log, true, &m)) return e;
    u32 len = 1u << len_log;
    k_phi_level<<<(len + 255) / 256, 256, 0, dp_ctx().stream>>>((gle *)m->data, len, n_total, shift, e_make(gl_canon(A[0]), gl_canon(A[1])), e_make(gl_canon(B[0]), gl_canon(B[1])), inverse, tab); DP_LAUNCHED();
    DP_CUDA(cudaGetLastError());
    *out = m; return DP_OK;
}

// ---- out = src repeated `times` times (beta_acc = vec![beta2; kx].concat(), convolution.rs:870) ----
__global__ void k_repeat(const gle *src, u64 len, u64 total, gle *out) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x, stride = (u64)gridDim.x * blockDim.x;
    for (; t < total; t += stride) out[t] = src[t & (len - 1)];
}
extern "C" int dp_mle_repeat(const dp_mle *src, uint32_t times, dp_mle **out) {
    DP_REQUIRE_CTX();
    DP_CHECK(src && out && src->is_ext && times && (times & (times - 1)) == 0, DP_ERR_INVALID, "dp_mle_repeat: need an Ext MLE and a power-of-two count");
    dp_mle *m; if (int e = new_mle(src->len * times, true, &m)) return e;
    k_repeat<<<dp_grid_for(m->len, 256, 8), 256, 0, dp_ctx().stream>>>((const gle *)src->data, src->len, m->len, (gle *)m->data); DP_LAUNCHED();
    DP_CUDA(cudaGetLastError());
    *out = m; return DP_OK;
}
