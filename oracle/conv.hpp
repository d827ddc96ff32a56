// TEST INFRASTRUCTURE ONLY (see oracle/README): CPU restatement of the reference's FFT-convolution layer and its proof.
//   tensor.rs:220-323        get_root_of_unity, index_w, index_u, fft
//   tensor.rs:326-372,458-523 ConvData, fft_conv
//   layers/convolution.rs:144-176,320-350,368-458,697-1077,1484-1550  add_bias, op, prove_batch_fft_weights,
//                            prove_convolution_step, clear_garbage, new_clearing_tensor, index_wf
//   layers/hadamard.rs:83-126 hadamard::prove
//   iop/prover.rs:164-399    delegate_matrix_evaluation, phi_pow_init, phi_g_init, prove_batch_fft, prove_batch_ifft
// The reference's debug_assert! invariants are kept as hard checks (they are what pins this restatement).
#pragma once
// NOTE: included from the middle of zkml.hpp (needs Claim, Element, ext_mle, sumcheck_prove; zk_prove needs this file)

namespace dpo {

static const u64 TWO_ADIC_GENERATOR_32 = 1753635133440165772ULL;   // Goldilocks::two_adic_generator(32)

// tensor.rs:220-231
static inline E get_root_of_unity(size_t n) {
    E rou = E::from_base(TWO_ADIC_GENERATOR_32);
    for (size_t i = 0; i < 32 - n; i++) rou = e_mul(rou, rou);
    return rou;
}
// tensor.rs:261-323  (flag false: FFT, true: iFFT incl. 1/n scaling); natural order in and out
static inline void fft_ext(std::vector<E> &v, bool flag) {
    size_t n = v.size(), logn = ceil_log2(n);
    std::vector<size_t> rev(n, 0);
    for (size_t i = 1; i < n; i++) rev[i] = (rev[i >> 1] >> 1) | ((i & 1) << (logn - 1));
    std::vector<E> w(n, E::zero());
    w[0] = E::one();
    if (n > 1) { w[1] = get_root_of_unity(logn); if (flag) w[1] = e_inv(w[1]); }
    for (size_t i = 2; i < n; i++) w[i] = e_mul(w[i - 1], w[1]);
    for (size_t i = 0; i < n; i++) if (rev[i] < i) std::swap(v[i], v[rev[i]]);
    for (size_t i = 2; i <= n; i <<= 1) {
        size_t half = i >> 1;
        for (size_t c = 0; c < n; c += i)
            for (size_t k = 0; k < half; k++) {
                E u = v[c + k], l = e_mul(v[c + k + half], w[n / i * k]);
                v[c + k] = e_add(u, l); v[c + k + half] = e_sub(u, l);
            }
    }
    if (flag) { E ilen = e_inv(E::from_base(f_from_u64(n))); for (auto &x : v) x = e_mul(x, ilen); }
}
static inline Element to_element(E e) {                      // quantization/mod.rs:225-242
    u64 c = e.c0; if (c == 0 && e.c1 == 0) return 0;
    return c <= (GL_P >> 1) ? (Element)c : -(Element)(GL_P - c);
}
// tensor.rs:236-253 / convolution.rs:1535-1550: place an n_real x n_real filter in the top-left of an n x n grid
template <class T, class F> static inline std::vector<E> index_wf_gen(const T *w, size_t n_real, size_t n, size_t out_len, F conv) {
    std::vector<E> o(out_len, E::zero());
    for (size_t idx = 0; idx < out_len; idx++) { size_t i = idx / n, j = idx % n; if (i < n_real && j < n_real) o[idx] = conv(w[i * n_real + j]); }
    return o;
}
static inline std::vector<E> index_w(const Element *w, size_t n_real, size_t n, size_t out_len) { return index_wf_gen(w, n_real, n, out_len, [](Element e) { return E::from_base(f_from_i64(e)); }); }
static inline std::vector<E> index_wf(const E *w, size_t n_real, size_t n, size_t out_len) { return index_wf_gen(w, n_real, n, out_len, [](E e) { return e; }); }

// Convolution<Element> after into_padded_and_ffted: filter shape [kw, kx, nw, nw], data laid out [kw][kx][real_nw][real_nw]
struct ConvLayer {
    size_t kw = 0, kx = 0, nw = 0, real_nw = 0;
    std::vector<Element> filter, bias;                       // bias: kw entries
    size_t unpadded_out[3] = {0, 0, 0};                      // conv2d_shape(unpadded input, unpadded filter): [k_w, h, h]
    size_t filter_size() const { return nw * nw; }
};
struct ConvData {                                            // tensor.rs:326-337
    std::vector<E> real_input;
    std::vector<std::vector<E>> input, input_fft, prod, output;
    std::vector<Element> output_as_element;                  // after op(): conv output AFTER bias addition, before clearing
};
// tensor.rs:458-523 + ConvData::new (:343-367)
static inline std::vector<Element> fft_conv(const ConvLayer &f, const std::vector<Element> &x, size_t n_x, ConvData &cd) {
    size_t new_n = 2 * n_x * n_x, chunk = n_x * n_x;
    if (x.size() != f.kx * chunk) throw std::runtime_error("fft_conv: input is not [kx, n_x, n_x]");
    if (f.nw != n_x) throw std::runtime_error("fft_conv: filter nw must equal the padded input width");
    cd = ConvData();
    for (Element e : x) cd.real_input.push_back(E::from_base(f_from_i64(e)));
    for (size_t c = 0; c < f.kx; c++) {
        std::vector<E> in(cd.real_input.begin() + c * chunk, cd.real_input.begin() + (c + 1) * chunk);
        std::reverse(in.begin(), in.end());
        std::vector<E> ff = in; ff.resize(new_n, E::zero());
        fft_ext(ff, false);
        cd.input.push_back(in); cd.input_fft.push_back(ff);
    }
    std::vector<std::vector<E>> out(f.kw, std::vector<E>(2 * f.nw * f.nw, E::zero()));
    size_t fs = f.real_nw * f.real_nw;
    for (size_t i = 0; i < f.kw; i++)
        for (size_t j = 0; j < f.kx; j++) {
            std::vector<E> wf = index_w(f.filter.data() + i * f.kx * fs + j * fs, f.real_nw, f.nw, 2 * f.nw * f.nw);
            fft_ext(wf, false);
            for (size_t k = 0; k < out[i].size(); k++) out[i][k] = e_add(out[i][k], e_mul(cd.input_fft[j][k], wf[k]));
        }
    cd.prod = out;
    for (auto &row : out) fft_ext(row, true);
    cd.output = out;
    for (auto &row : cd.output) for (size_t i = 0; i < row.size() / 2; i++) cd.output_as_element.push_back(to_element(row[n_x * n_x - 1 - i]));   // index_u
    return cd.output_as_element;                             // tensor [kw, n_x, n_x]
}
// convolution.rs:1508-1530: 1 where (i, j, k) is inside the unpadded output, 0 on the padding garbage
static inline std::vector<Element> new_clearing_tensor(const size_t og[3], const size_t padded[3]) {
    std::vector<Element> d(padded[0] * padded[1] * padded[2], 0);
    for (size_t i = 0; i < padded[0]; i++) for (size_t j = 0; j < padded[1]; j++) for (size_t k = 0; k < padded[2]; k++)
        if (i < og[0] && j < og[1] && k < og[2]) d[i * padded[1] * padded[2] + j * padded[2] + k] = 1;
    return d;
}
// Convolution::op (convolution.rs:320-350): fft conv, + bias, garbage cleared; proving data keeps the after-bias tensor
static inline std::vector<Element> conv_op(const ConvLayer &f, const std::vector<Element> &x, size_t n_x, ConvData &cd) {
    std::vector<Element> out = fft_conv(f, x, n_x, cd);
    size_t fsz = n_x * n_x;
    for (size_t i = 0; i < f.kw; i++) for (size_t j = 0; j < fsz; j++) out[i * fsz + j] += f.bias[i];   // add_bias (:152-162)
    cd.output_as_element = out;
    size_t padded[3] = {f.kw, n_x, n_x};
    std::vector<Element> clr = new_clearing_tensor(f.unpadded_out, padded), cleared(out.size());
    for (size_t i = 0; i < out.size(); i++) cleared[i] = out[i] * clr[i];                               // == clear_garbage (:1484-1506)
    return cleared;
}

// ---- hadamard::prove (hadamard.rs:83-126) ----
struct HadamardProof { IOPProof sumcheck; std::vector<E> individual_claim; };
static inline std::vector<E> elems_to_ext(const std::vector<Element> &v) { std::vector<E> o; o.reserve(v.size()); for (Element e : v) o.push_back(E::from_base(f_from_i64(e))); return o; }
static inline HadamardProof hadamard_prove(Transcript &t, const Claim &out_claim, const std::vector<Element> &v1, const std::vector<Element> &v2) {
    if (v1.size() != v2.size() || (v1.size() & (v1.size() - 1)) || out_claim.point.size() != ceil_log2(v1.size())) throw std::runtime_error("hadamard: shape mismatch");
    std::vector<E> beta = build_eq_x_r_vec(out_claim.point), a = elems_to_ext(v1), b = elems_to_ext(v2);
    VirtualPolynomial vp(out_claim.point.size());
    vp.add_mle_list({ext_mle(a.data(), a.size()), ext_mle(b.data(), b.size()), ext_mle(beta.data(), beta.size())}, E::one());
    auto res = sumcheck_prove(vp, t);
    if (!(res.first.extract_sum() == out_claim.eval)) throw std::runtime_error("hadamard: sumcheck sum != output claim");
    return {res.first, {res.second[0], res.second[1]}};
}

// ---- FFT matrix delegation (iop/prover.rs:164-289) ----
static inline std::vector<E> phi_pow_init(size_t n, bool is_fft) {                  // prover.rs:215-227
    size_t len = (size_t)1 << n; E phi = get_root_of_unity(n); if (is_fft) phi = e_inv(phi);
    std::vector<E> o(len); o[0] = E::one(); for (size_t i = 1; i < len; i++) o[i] = e_mul(o[i - 1], phi);
    return o;
}
static inline void phi_g_init(std::vector<E> &phi_g, std::vector<std::vector<E>> &mid, const std::vector<E> &rx, E scale, size_t n, bool is_fft) {   // prover.rs:231-289
    std::vector<E> phi_mul = phi_pow_init(n, is_fft);
    E one = E::one();
    if (is_fft) {
        phi_g[0] = scale; phi_g[1] = scale;
        for (size_t i = 1; i <= n; i++) {
            for (size_t b = 0; b < ((size_t)1 << (i - 1)); b++) {
                size_t l = b, r = b ^ ((size_t)1 << (i - 1)), m = n - i;
                E tmp1 = e_sub(one, rx[m]), tmp2 = e_mul(rx[m], phi_mul[b << m]);
                phi_g[r] = e_mul(phi_g[l], e_sub(tmp1, tmp2));
                phi_g[l] = e_mul(phi_g[l], e_add(tmp1, tmp2));
            }
            if (i < n) mid[i - 1].assign(phi_g.begin(), phi_g.begin() + ((size_t)1 << i));
        }
    } else {
        phi_g[0] = scale;
        for (size_t i = 1; i < n; i++) {
            for (size_t b = 0; b < ((size_t)1 << (i - 1)); b++) {
                size_t l = b, r = b ^ ((size_t)1 << (i - 1)), m = n - i;
                E tmp1 = e_sub(one, rx[m]), tmp2 = e_mul(rx[m], phi_mul[b << m]);
                phi_g[r] = e_mul(phi_g[l], e_sub(tmp1, tmp2));
                phi_g[l] = e_mul(phi_g[l], e_add(tmp1, tmp2));
            }
            mid[i - 1].assign(phi_g.begin(), phi_g.begin() + ((size_t)1 << i));
        }
        for (size_t b = 0; b < ((size_t)1 << (n - 1)); b++) phi_g[b] = e_mul(phi_g[b], e_add(e_sub(one, rx[0]), e_mul(rx[0], phi_mul[b])));
    }
}
struct MatrixEvalProof { std::vector<IOPProof> proofs; std::vector<std::vector<E>> claims; };
static inline MatrixEvalProof delegate_matrix_evaluation(Transcript &t, std::vector<std::vector<E>> &f_middle, const std::vector<E> &r1, std::vector<E> r2, bool is_fft) {   // prover.rs:164-212
    std::vector<E> omegas = phi_pow_init(r1.size(), is_fft);
    MatrixEvalProof out; E one = E::one(), two = E::from_base(2);
    size_t fm = f_middle.size();
    for (size_t l = r1.size() - 1; l-- > 0;) {
        std::vector<E> phi(f_middle[l].size());
        std::vector<E> beta = build_eq_x_r_vec(std::vector<E>(r2.begin(), r2.end() - 1));
        E rl = r1[(fm - 1) - l], last = r2.back();
        for (size_t i = 0; i < phi.size(); i++) {
            E om = omegas[i << ((fm - 1) - l)];
            if (!is_fft && l == fm - 1) phi[i] = e_mul(e_sub(one, last), e_add(e_sub(one, rl), e_mul(rl, om)));
            else phi[i] = e_add(e_sub(one, rl), e_mul(e_mul(e_sub(one, e_mul(two, last)), rl), om));
        }
        if (beta.size() != phi.size()) throw std::runtime_error("delegate_matrix_evaluation: size mismatch");
        VirtualPolynomial vp(ceil_log2(phi.size()));
        vp.add_mle_list({ext_mle(beta.data(), beta.size()), ext_mle(phi.data(), phi.size()), ext_mle(f_middle[l].data(), f_middle[l].size())}, E::one());
        auto res = sumcheck_prove(vp, t);
        r2 = res.first.point;
        out.proofs.push_back(res.first); out.claims.push_back(res.second);
    }
    return out;
}
struct BatchFFTProof { IOPProof proof; std::vector<E> claims; MatrixEvalProof matrix_eval; };
static inline std::shared_ptr<MLE> flat_rows_mle(const std::vector<std::vector<E>> &rows) { std::vector<E> f; for (auto &r : rows) f.insert(f.end(), r.begin(), r.end()); return ext_mle(f.data(), f.size()); }
// prover.rs:295-349
static inline BatchFFTProof prove_batch_fft(Transcript &t, const std::vector<E> &r, std::vector<std::vector<E>> x) {
    size_t padded_rows = 2 * x[0].size();
    for (auto &row : x) row.resize(padded_rows, E::zero());
    size_t l1 = ceil_log2(x[0].size()), l2 = ceil_log2(x.size());
    std::vector<E> r1(r.begin(), r.begin() + l1), r2(r.begin() + l1, r.begin() + l1 + l2);
    std::vector<E> w_red(x[0].size(), E::zero()); std::vector<std::vector<E>> f_middle(r1.size() - 1);
    phi_g_init(w_red, f_middle, r1, E::one(), l1, false);
    MLE fm = mle_fix_high_variables(*flat_rows_mle(x), r2);
    VirtualPolynomial vp(fm.num_vars);
    vp.add_mle_list({std::make_shared<MLE>(fm), ext_mle(w_red.data(), w_red.size())}, E::one());
    auto res = sumcheck_prove(vp, t);
    BatchFFTProof o; o.proof = res.first; o.claims = res.second;
    o.matrix_eval = delegate_matrix_evaluation(t, f_middle, r1, res.first.point, false);
    return o;
}
// prover.rs:351-399
static inline BatchFFTProof prove_batch_ifft(Transcript &t, const std::vector<E> &r, const std::vector<std::vector<E>> &prod) {
    E scale = e_inv(E::from_base(f_from_u64(prod[0].size())));
    size_t l1 = ceil_log2(prod[0].size()), l2 = ceil_log2(prod.size());
    std::vector<E> r1(r.begin(), r.begin() + l1), r2(r.begin() + l1, r.begin() + l1 + l2);
    if (!(r1.back() == E::zero())) throw std::runtime_error("Error in randomness init batch ifft");
    std::vector<E> w_red(prod[0].size(), E::zero()); std::vector<std::vector<E>> f_middle(r1.size() - 1);
    phi_g_init(w_red, f_middle, r1, scale, l1, true);
    MLE fm = mle_fix_high_variables(*flat_rows_mle(prod), r2);
    VirtualPolynomial vp(fm.num_vars);
    vp.add_mle_list({std::make_shared<MLE>(fm), ext_mle(w_red.data(), w_red.size())}, E::one());
    auto res = sumcheck_prove(vp, t);
    BatchFFTProof o; o.proof = res.first; o.claims = res.second;
    o.matrix_eval = delegate_matrix_evaluation(t, f_middle, r1, res.first.point, true);
    return o;
}
struct BatchFFTWeightsProof { IOPProof proof; std::vector<E> claims, partial_evals; MatrixEvalProof matrix_evaluation; };
// convolution.rs:368-458
static inline BatchFFTWeightsProof prove_batch_fft_weights(const ConvLayer &f, Transcript &t, const std::vector<E> &r) {
    size_t padded_rows = 2 * f.nw * f.nw, fs = f.real_nw * f.real_nw, l1 = ceil_log2(padded_rows);
    std::vector<E> r1(r.begin(), r.begin() + l1), r2(r.begin() + l1, r.end());
    std::vector<E> w_red(padded_rows, E::zero()); std::vector<std::vector<E>> f_middle(r1.size() - 1);
    std::vector<E> beta = build_eq_x_r_vec(r2);
    phi_g_init(w_red, f_middle, r1, E::one(), l1, false);
    std::vector<E> w1(fs, E::zero());
    for (size_t i = 0; i < f.kw; i++) for (size_t j = 0; j < f.kx; j++) for (size_t k = 0; k < fs; k++)
        w1[k] = e_add(w1[k], e_mul_base(beta[i * f.kx + j], f_from_i64(f.filter[i * fs * f.kx + j * fs + k])));
    BatchFFTWeightsProof o; o.partial_evals = w1;
    std::vector<E> padded = index_wf(w1.data(), f.real_nw, f.nw, padded_rows);
    VirtualPolynomial vp(l1);
    vp.add_mle_list({ext_mle(padded.data(), padded.size()), ext_mle(w_red.data(), w_red.size())}, E::one());
    auto res = sumcheck_prove(vp, t);
    o.proof = res.first; o.claims = res.second;
    o.matrix_evaluation = delegate_matrix_evaluation(t, f_middle, r1, res.first.point, false);
    return o;
}

struct ConvProof {                                           // convolution.rs:97-121
    IOPProof fft_proof, fft_proof_weights, ifft_proof, hadamard_proof;
    MatrixEvalProof fft_delegation, fft_delegation_weights, ifft_delegation;
    std::vector<E> fft_claims, ifft_claims, fft_weight_claims, hadamard_claims, partial_evals;
    E bias_claim;
    HadamardProof clearing_proof;
    Claim filter_claim, bias_poly_claim;                     // handed to the commitment prover (add_common_claims, :1003-1010)
};
// Convolution::prove_convolution_step (convolution.rs:697-1077).  `last_claim` is on the layer's (cleared) output tensor
// [kw, nw, nw]; returns the claim on the layer input tensor [kx, nw, nw].
static inline Claim prove_convolution_step(const ConvLayer &f, Transcript &t, const Claim &last_claim_in, const ConvData &pd, ConvProof &out) {
    size_t n_x = f.nw, lfs = ceil_log2(f.filter_size()), lkw = ceil_log2(f.kw), lrow = ceil_log2(2 * f.filter_size());
    size_t padded[3] = {f.kw, n_x, n_x};
    std::vector<Element> clearing = new_clearing_tensor(f.unpadded_out, padded);
    out.clearing_proof = hadamard_prove(t, last_claim_in, pd.output_as_element, clearing);
    Claim last_claim{out.clearing_proof.sumcheck.point, out.clearing_proof.individual_claim[0]};
    if (f.filter_size() * f.kw * 2 != pd.output.size() * pd.output[0].size()) throw std::runtime_error("Inconsistent output size");
    if (lfs + lkw != last_claim.point.size()) throw std::runtime_error("Inconsistent random point size");
    std::vector<E> r(last_claim.point.size() + 1, E::zero()), bias_point(lkw, E::zero());
    for (size_t i = 0; i < lfs; i++) r[i] = e_sub(E::one(), last_claim.point[i]);
    for (size_t i = 0; i < lkw; i++) { r[i + lfs + 1] = last_claim.point[i + lfs]; bias_point[i] = last_claim.point[i + lfs]; }
    E bias_eval = E::zero();
    std::vector<E> bias_e = elems_to_ext(f.bias);
    if (!bias_point.empty()) bias_eval = mle_evaluate(*ext_mle(bias_e.data(), bias_e.size()), bias_point);
    else if (f.bias.size() == 1) bias_eval = bias_e[0];
    if (!(mle_evaluate(*flat_rows_mle(pd.output), r) == e_sub(last_claim.eval, bias_eval))) throw std::runtime_error("Error in Conv 1");

    Transcript temp_t = t;
    BatchFFTProof ifft = prove_batch_ifft(t, r, pd.prod);
    if (ifft.proof.point.size() != lfs + 1) throw std::runtime_error("Error in ifft sumceck");
    sumcheck_verify(e_sub(last_claim.eval, bias_eval), ifft.proof, lfs + 1, 2, temp_t);   // throws on failure

    std::vector<E> r_ifft = ifft.proof.point;
    for (size_t i = lrow; i < r.size(); i++) r_ifft.push_back(r[i]);
    if (!(mle_evaluate(*flat_rows_mle(pd.prod), r_ifft) == ifft.claims[0])) throw std::runtime_error("Error in Conv 1 (prod)");
    std::vector<E> r1(r_ifft.begin() + lrow, r_ifft.end()), r2(r_ifft.begin(), r_ifft.begin() + lrow);
    std::vector<E> beta1 = build_eq_x_r_vec(r1), beta2 = build_eq_x_r_vec(r2);
    std::vector<E> beta_acc; for (size_t i = 0; i < f.kx; i++) beta_acc.insert(beta_acc.end(), beta2.begin(), beta2.end());
    size_t fs = f.real_nw * f.real_nw;
    std::vector<E> f1;
    for (size_t i = 0; i < f.kx; i++) {
        std::vector<E> agg(fs, E::zero());
        for (size_t j = 0; j < f.kw; j++) for (size_t k = 0; k < fs; k++) agg[k] = e_add(agg[k], e_mul_base(beta1[j], f_from_i64(f.filter[j * f.kx * fs + i * fs + k])));
        std::vector<E> p = index_wf(agg.data(), f.real_nw, f.nw, 2 * f.nw * f.nw);
        fft_ext(p, false);
        f1.insert(f1.end(), p.begin(), p.end());
    }
    auto m1 = ext_mle(f1.data(), f1.size()), m2 = flat_rows_mle(pd.input_fft), m3 = ext_mle(beta_acc.data(), beta_acc.size());
    VirtualPolynomial vp(m1->num_vars);
    vp.add_mle_list({m1, m2, m3}, E::one());
    auto had = sumcheck_prove(vp, t);
    out.hadamard_proof = had.first; out.hadamard_claims = had.second;
    std::vector<E> point = had.first.point; point.insert(point.end(), r1.begin(), r1.end());

    BatchFFTProof fftp = prove_batch_fft(t, had.first.point, pd.input);
    BatchFFTWeightsProof fw = prove_batch_fft_weights(f, t, point);
    size_t lw = ceil_log2(fs);
    std::vector<E> weights_rand = t.sample_vec(lw);                                   // read_challenges (:929-931)
    {   // the reference's debug block (:932-975): padded-weights evaluation is consistent with the FFT claim
        std::vector<E> wp = fw.proof.point; E vw = e_inv(e_sub(E::one(), wp.back())); wp.pop_back();
        std::vector<E> rr(wp.begin(), wp.begin() + ceil_log2(f.nw * f.nw));
        std::vector<E> eqt = build_eq_x_r_vec(rr);
        E y = E::zero();
        for (size_t i = 0; i < f.real_nw; i++) for (size_t j = 0; j < f.real_nw; j++) y = e_add(y, e_mul(eqt[i * f.nw + j], fw.partial_evals[i * f.real_nw + j]));
        if (!(y == e_mul(fw.claims[0], vw))) throw std::runtime_error("Error in padded weights eval");
        std::vector<E> fp = weights_rand; fp.insert(fp.end(), point.begin() + lrow, point.end());
        std::vector<E> wts = elems_to_ext(f.filter);
        if (!(mle_evaluate(*ext_mle(wts.data(), wts.size()), fp) == mle_evaluate(*ext_mle(fw.partial_evals.data(), fw.partial_evals.size()), weights_rand))) throw std::runtime_error("Error in fft_weights eval");
    }
    out.bias_poly_claim = Claim{bias_point, bias_eval};
    std::vector<E> fpnt = weights_rand; fpnt.insert(fpnt.end(), point.begin() + lrow, point.end());
    out.filter_claim = Claim{fpnt, mle_evaluate(*ext_mle(fw.partial_evals.data(), fw.partial_evals.size()), weights_rand)};
    out.fft_proof = fftp.proof; out.fft_claims = fftp.claims; out.fft_delegation = fftp.matrix_eval;
    out.fft_proof_weights = fw.proof; out.fft_weight_claims = fw.claims; out.fft_delegation_weights = fw.matrix_evaluation; out.partial_evals = fw.partial_evals;
    out.ifft_proof = ifft.proof; out.ifft_claims = ifft.claims; out.ifft_delegation = ifft.matrix_eval;
    out.bias_claim = bias_eval;

    std::vector<E> input_point = fftp.proof.point;
    E v = e_inv(e_sub(E::one(), input_point.back())); input_point.pop_back();
    {   // :1035-1060: the returned claim really is an evaluation of the (padded) input tensor
        std::vector<E> p = input_point; p.insert(p.end(), had.first.point.begin() + lrow, had.first.point.end());
        if (!(mle_evaluate(*flat_rows_mle(pd.input), p) == e_mul(fftp.claims[0], v))) throw std::runtime_error("Error in input eval CONV PROVER");
        for (size_t i = 0; i < lfs; i++) p[i] = e_sub(E::one(), p[i]);
        if (!(mle_evaluate(*ext_mle(pd.real_input.data(), pd.real_input.size()), p) == e_mul(fftp.claims[0], v))) throw std::runtime_error("Error in real input eval CONV PROVER");
    }
    for (auto &ip : input_point) ip = e_sub(E::one(), ip);
    Claim fin; fin.point = input_point; fin.point.insert(fin.point.end(), had.first.point.begin() + lrow, had.first.point.end());
    fin.eval = e_mul(fftp.claims[0], v);
    return fin;
}

// deterministic synthetic layer + input for tests and golden vectors (NOT from the reference)
static inline ConvLayer synthetic_conv(size_t kw, size_t kx, size_t n_x, size_t real_nw, size_t kw_u, size_t k_u, size_t n_x_u, u64 seed) {
    ConvLayer f; f.kw = kw; f.kx = kx; f.nw = n_x; f.real_nw = real_nw;
    SplitMix64 g(seed);
    f.filter.assign(kw * kx * real_nw * real_nw, 0);
    // real (unpadded) weights live in the top-left k_u x k_u corner of the first kw_u output channels; the rest is padding
    for (size_t i = 0; i < kw_u; i++) for (size_t j = 0; j < kx; j++) for (size_t a = 0; a < k_u; a++) for (size_t b = 0; b < k_u; b++)
        f.filter[((i * kx + j) * real_nw + a) * real_nw + b] = (Element)(g.next() % 255) - 127;
    f.bias.assign(kw, 0); for (size_t i = 0; i < kw_u; i++) f.bias[i] = (Element)(g.next() % 2001) - 1000;
    f.unpadded_out[0] = kw_u; f.unpadded_out[1] = f.unpadded_out[2] = n_x_u - k_u + 1;
    return f;
}
static inline std::vector<Element> synthetic_conv_input(size_t kx, size_t n_x, size_t kx_u, size_t n_x_u, u64 seed) {
    SplitMix64 g(seed); std::vector<Element> x(kx * n_x * n_x, 0);
    for (size_t c = 0; c < kx_u; c++) for (size_t i = 0; i < n_x_u; i++) for (size_t j = 0; j < n_x_u; j++) x[(c * n_x + i) * n_x + j] = (Element)(g.next() % 255) - 127;
    return x;
}

}  // namespace dpo
