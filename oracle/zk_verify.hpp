// TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's model VERIFIER, so that whole proofs (from this oracle and from
// the device) are accepted by the logic deep-prove checks them with.  The verifier re-derives every Fiat-Shamir
// challenge on its own transcript, so it also pins the prover restatement's transcript ORDER, which GPU-vs-oracle
// equality alone cannot.
//   Verifier::verify                       zkml/src/iop/verifier.rs:72-296, verify_table :320-383
//   ChallengeStorage::initialise           zkml/src/iop/mod.rs:70-90
//   verify_logup_proof                     zkml/src/lookup/logup_gkr/verifier.rs:16-211
//   TableType::evaluate_table_columns      zkml/src/lookup/context.rs:323-407
//   verify_dense / verify_requant / verify_activation   layers/dense.rs:576-640, requant.rs:689-816, activation.rs:459-512
//   same_poly::Verifier::verify            zkml/src/commit/same_poly.rs:157-185
//   CommitmentVerifier                     zkml/src/commit/context.rs:420-600
#pragma once
#include "zkml.hpp"
#include "verify.hpp"

namespace dpo {

struct ZkVerifyError : std::runtime_error { using std::runtime_error::runtime_error; };
static inline void zk_ensure(bool ok, const std::string &msg) { if (!ok) throw ZkVerifyError(msg); }

struct LogUpVerifierClaim { std::vector<Claim> claims; std::vector<E> numerators, denominators; const std::vector<E> &point() const { return claims[0].point; } };

// verify_logup_proof (lookup/logup_gkr/verifier.rs:16-160)
static inline LogUpVerifierClaim verify_logup_proof(const LogUpProof &proof, size_t num_instances, E constant_challenge, E column_separation_challenge, Transcript &t) {
    t.append_field_element(f_from_u64(num_instances));
    for (auto &o : proof.circuit_outputs) t.append_field_element_exts(o);
    LogUpVerifierClaim out;
    for (auto &e : proof.circuit_outputs) { out.numerators.push_back(e_add(e_mul(e[0], e[3]), e_mul(e[1], e[2]))); out.denominators.push_back(e_mul(e[2], e[3])); }
    E batching = t.get_and_append_challenge("initial_batching"), alpha = t.get_and_append_challenge("initial_alpha"), lambda = t.get_and_append_challenge("initial_lambda");
    E current = E::zero(), ac = E::one();
    for (auto &e : proof.circuit_outputs) {
        current = e_add(current, e_mul(ac, e_add(e_add(e_mul(batching, e_sub(e[1], e[0])), e[0]), e_mul(lambda, e_add(e_mul(batching, e_sub(e[3], e[2])), e[2])))));
        ac = e_mul(ac, alpha);
    }
    std::vector<E> sc_point = {batching};
    zk_ensure(proof.sumcheck_proofs.size() == proof.round_evaluations.size(), "logup: proofs / evaluations length mismatch");
    for (size_t i = 0; i < proof.sumcheck_proofs.size(); i++) {
        const IOPProof &sp = proof.sumcheck_proofs[i]; const std::vector<E> &re = proof.round_evaluations[i];
        t.append_field_element_ext(current);
        E eq_ev = eq_eval(sc_point, sp.point);                                   // identity_eval
        SumCheckSubClaim sub = sumcheck_verify(current, sp, i + 1, 3, t);
        E nb = t.get_and_append_challenge("logup_batching"), na = t.get_and_append_challenge("logup_alpha"), nl = t.get_and_append_challenge("logup_lambda");
        size_t per = re.size() / num_instances;
        E next = E::zero(), nac = E::one(), sc = E::zero(), pa = E::one();
        if (per == 4) {
            for (size_t k = 0; k + 3 < re.size(); k += 4) {
                const E *e = &re[k];
                next = e_add(next, e_mul(nac, e_add(e_add(e_mul(nb, e_sub(e[2], e[0])), e[0]), e_mul(nl, e_add(e_mul(nb, e_sub(e[1], e[3])), e[3])))));
                sc = e_add(sc, e_mul(pa, e_mul(eq_ev, e_add(e_add(e_mul(e[0], e[1]), e_mul(e[2], e[3])), e_mul(lambda, e_mul(e[3], e[1]))))));
                nac = e_mul(nac, na); pa = e_mul(pa, alpha);
            }
        } else {
            for (size_t k = 0; k + 1 < re.size(); k += 2) {
                const E *e = &re[k];
                next = e_add(next, e_mul(nac, e_add(e_mul(nb, e_sub(e[0], e[1])), e[1])));
                sc = e_add(sc, e_mul(e_mul(pa, eq_ev), e_add(e_sub(e_neg(e[1]), e[0]), e_mul(lambda, e_mul(e[0], e[1])))));
                nac = e_mul(nac, na); pa = e_mul(pa, alpha);
            }
        }
        zk_ensure(sc == sub.expected_evaluation, "logup: calculated sumcheck claim != sumcheck output claim at round " + std::to_string(i));
        current = next; alpha = na; lambda = nl;
        sc_point = sub.point; sc_point.push_back(nb);
    }
    // calculate_final_eval (:163-211)
    E fin;
    if (!proof.table) {
        size_t per = proof.output_claims.size() / num_instances; E acc = E::zero(), acm = E::one();
        for (size_t c = 0; c < proof.output_claims.size(); c += per) {
            E ch = constant_challenge, cs = E::one();
            for (size_t k = 0; k < per; k++) { ch = e_add(ch, e_mul(proof.output_claims[c + k].eval, cs)); cs = e_mul(cs, column_separation_challenge); }
            acc = e_add(acc, e_mul(ch, acm)); acm = e_mul(acm, alpha);
        }
        fin = acc;
    } else {
        E cols = constant_challenge, cs = E::one();
        for (size_t k = 1; k < proof.output_claims.size(); k++) { cols = e_add(cols, e_mul(proof.output_claims[k].eval, cs)); cs = e_mul(cs, column_separation_challenge); }
        fin = e_add(proof.output_claims[0].eval, e_mul(lambda, cols));
    }
    zk_ensure(fin == current, "logup: calculated final value does not match the final sumcheck output");
    for (auto &c : proof.output_claims) zk_ensure(c.point == sc_point, "logup: output claim at an unexpected point");
    out.claims = proof.output_claims;
    return out;
}
// TableType::evaluate_table_columns (lookup/context.rs:323-407)
static inline std::vector<E> evaluate_table_columns(const TableType &tt, const std::vector<E> &point) {
    auto bits = [&](size_t n) { E a = E::zero(); for (size_t i = 0; i < n; i++) a = e_add(a, e_mul_base(point[i], f_from_u64((u64)1 << i))); return a; };
    if (tt.kind == TT_RANGE) { zk_ensure(point.size() == Q_BIT_LEN, "range table point size"); return {bits(point.size())}; }
    if (tt.kind == TT_RELU) {
        zk_ensure(point.size() == Q_BIT_LEN, "relu table point size");
        E first = e_sub(bits(point.size()), E::from_base(f_from_u64((u64)1 << (Q_BIT_LEN - 1))));
        E second = e_mul(bits(point.size() - 1), point.back());
        return {first, second};
    }
    zk_ensure(point.size() == tt.size, "clamping table point size");
    E first = e_sub(bits(point.size()), E::from_base(f_from_u64((u64)1 << (tt.size - 1))));
    Element mx = (Element)1 << (tt.size - 1); std::vector<u64> col;
    for (Element i = -mx; i < mx; i++) col.push_back(f_from_i64(i < Q_MIN ? Q_MIN : (i > Q_MAX ? Q_MAX : i)));
    return {first, mle_evaluate(*base_mle(col), point)};
}


// pow_two_omegas / phi_eval (layers/convolution.rs:1450-1482)
static inline std::vector<E> pow_two_omegas(size_t n, bool is_fft) {
    std::vector<E> pows(n - 1); E rou = get_root_of_unity(n); if (is_fft) rou = e_inv(rou);
    pows[0] = rou; for (size_t i = 1; i + 1 < n; i++) pows[i] = e_mul(pows[i - 1], pows[i - 1]);
    return pows;
}
static inline E phi_eval(const std::vector<E> &r, E rand1, E rand2, const std::vector<E> &exponents, bool first_iter) {
    E ev = E::one(), one = E::one();
    for (size_t i = 0; i < r.size(); i++) ev = e_mul(ev, e_add(e_sub(one, r[i]), e_mul(r[i], exponents[exponents.size() - r.size() + i])));
    if (first_iter) return e_mul(e_sub(one, rand2), e_add(e_sub(one, rand1), e_mul(rand1, ev)));
    return e_add(e_sub(one, rand1), e_mul(e_mul(e_sub(one, e_mul(E::from_u64(2), rand2)), rand1), ev));
}
// ConvCtx::verify_fft_delegation (convolution.rs:1090-1141)
static inline void verify_fft_delegation(Transcript &t, E claim, const ConvProof &proof, const MatrixEvalProof &del, std::vector<E> prev_r, size_t lfs) {
    size_t iter = del.proofs.size();
    std::vector<E> exponents = pow_two_omegas(iter + 1, false);
    for (size_t i = 0; i < iter; i++) {
        sumcheck_verify(claim, del.proofs[i], lfs - i, 3, t);
        zk_ensure(eq_eval(del.proofs[i].point, std::vector<E>(prev_r.begin(), prev_r.begin() + del.proofs[i].point.size())) == del.claims[i][0], "Error in identity evaluation fft delegation");
        zk_ensure(phi_eval(del.proofs[i].point, proof.hadamard_proof.point[i], prev_r.back(), exponents, i == 0) == del.claims[i][1], "Error in phi computation fft delegation");
        claim = del.claims[i][2]; prev_r = del.proofs[i].point;
    }
    E one = E::one();
    zk_ensure(claim == e_add(e_mul(e_sub(one, e_mul(E::from_u64(2), proof.hadamard_proof.point[iter])), prev_r[0]), e_sub(one, prev_r[0])), "Error in final FFT delegation step");
}
// hadamard::verify (layers/hadamard.rs:128-159)
static inline Claim hadamard_verify(Transcript &t, const HadamardProof &proof, const Claim &out_claim, E expected_v2_eval) {
    SumCheckSubClaim sub = sumcheck_verify(out_claim.eval, proof.sumcheck, out_claim.point.size(), 3, t);
    E beta_eval = eq_eval(out_claim.point, proof.sumcheck.point);
    zk_ensure(expected_v2_eval == proof.individual_claim[1], "Hadamard verification failed for v2 eval");
    zk_ensure(e_mul(e_mul(beta_eval, proof.individual_claim[0]), proof.individual_claim[1]) == sub.expected_evaluation, "Hadamard verification failed for product eval");
    return {proof.sumcheck.point, proof.individual_claim[0]};
}
// ConvCtx::verify_convolution (convolution.rs:1143-1375); returns the claim on the layer input, fills the two commitment claims
static inline Claim verify_convolution(const ConvLayer &f, Transcript &t, const Claim &last_claim_in, const ConvProof &proof, Claim &filter_claim, Claim &bias_claim) {
    size_t lfs = ceil_log2(f.filter_size()), lrow = lfs + 1, lkx = ceil_log2(f.kx);
    size_t padded[3] = {f.kw, f.nw, f.nw};
    std::vector<Element> clearing = new_clearing_tensor(f.unpadded_out, padded);
    E expected_v2 = mle_evaluate(*base_mle(to_base_vec(clearing)), proof.clearing_proof.sumcheck.point);
    Claim last_claim = hadamard_verify(t, proof.clearing_proof, last_claim_in, expected_v2);
    E conv_claim = e_sub(last_claim.eval, proof.bias_claim);
    sumcheck_verify(conv_claim, proof.ifft_proof, lrow, 2, t);
    zk_ensure(proof.ifft_delegation.proofs.size() == lfs, "Inconsistency in iFFT delegation proofs/aux size");
    size_t iter = lfs; E claim = proof.ifft_claims[1], one = E::one();
    std::vector<E> exponents = pow_two_omegas(iter + 1, true), prev_r = proof.ifft_proof.point;
    for (size_t i = 0; i < iter; i++) {
        const IOPProof &dp = proof.ifft_delegation.proofs[i];
        sumcheck_verify(claim, dp, lfs - i, 3, t);
        zk_ensure(eq_eval(dp.point, std::vector<E>(prev_r.begin(), prev_r.begin() + dp.point.size())) == proof.ifft_delegation.claims[i][0], "Error in identity evaluation ifft delegation");
        zk_ensure(phi_eval(dp.point, e_sub(one, last_claim.point[i]), prev_r.back(), exponents, false) == proof.ifft_delegation.claims[i][1], "Error in phi computation ifft delegation");
        prev_r = dp.point; claim = proof.ifft_delegation.claims[i][2];
    }
    E scale = e_inv(E::from_u64((u64)1 << (iter + 1)));
    zk_ensure(claim == e_add(e_mul(scale, prev_r[0]), e_mul(scale, e_sub(one, prev_r[0]))), "Error in final iFFT delegation step");
    sumcheck_verify(proof.ifft_claims[0], proof.hadamard_proof, lrow + lkx, 3, t);
    zk_ensure(proof.hadamard_claims[2] == eq_eval(proof.ifft_proof.point, std::vector<E>(proof.hadamard_proof.point.begin(), proof.hadamard_proof.point.begin() + proof.ifft_proof.point.size())), "Error in Beta evaluation");
    sumcheck_verify(proof.hadamard_claims[1], proof.fft_proof, lrow, 2, t);
    zk_ensure(proof.fft_delegation.proofs.size() == lfs, "Inconsistency in FFT delegation proofs/aux size");
    verify_fft_delegation(t, proof.fft_claims[1], proof, proof.fft_delegation, proof.fft_proof.point, lfs);
    sumcheck_verify(proof.hadamard_claims[0], proof.fft_proof_weights, lrow, 2, t);
    verify_fft_delegation(t, proof.fft_weight_claims[1], proof, proof.fft_delegation_weights, proof.fft_proof_weights.point, lfs);
    std::vector<E> wp = proof.fft_proof_weights.point; E v = e_inv(e_sub(one, wp.back())); wp.pop_back();
    std::vector<E> eqt = build_eq_x_r_vec(wp);                                  // identity_eval(to_bits(i nw + j), weights_point) for all (i, j)
    E yw = E::zero();
    for (size_t i = 0; i < f.real_nw; i++) for (size_t j = 0; j < f.real_nw; j++) yw = e_add(yw, e_mul(proof.partial_evals[i * f.real_nw + j], eqt[i * f.nw + j]));
    zk_ensure(e_mul(proof.fft_weight_claims[0], v) == yw, "Error in padded_fft evaluation claim");
    std::vector<E> weights_rand = t.sample_vec(ceil_log2(f.real_nw * f.real_nw));
    std::vector<E> point = proof.hadamard_proof.point; point.insert(point.end(), last_claim.point.begin() + lfs, last_claim.point.end());
    bias_claim = Claim{std::vector<E>(last_claim.point.begin() + proof.ifft_delegation.proofs.size(), last_claim.point.end()), proof.bias_claim};
    std::vector<E> fp = weights_rand; fp.insert(fp.end(), point.begin() + lrow, point.end());
    filter_claim = Claim{fp, mle_evaluate(*ext_mle(proof.partial_evals.data(), proof.partial_evals.size()), weights_rand)};
    std::vector<E> ip = proof.fft_proof.point; E vv = e_inv(e_sub(one, ip.back())); ip.pop_back();
    for (auto &x : ip) x = e_sub(one, x);
    Claim out; out.point = ip; out.point.insert(out.point.end(), proof.hadamard_proof.point.begin() + lrow, proof.hadamard_proof.point.end());
    out.eval = e_mul(proof.fft_claims[0], vv);
    return out;
}
// PoolingCtx::verify_pooling (layers/pooling.rs:525-650); the commitment claims are appended to `cv`
struct VerifierCommitClaim { PureCommitment comm; Claim claim; };
struct CommitmentVerifier {
    std::vector<VerifierCommitClaim> claims, trivial_claims;
    void add_witness_claim(const PureCommitment &c, const Claim &cl) { (cl.point.size() <= RS_BASECODE_MSG_SIZE_LOG ? trivial_claims : claims).push_back({c, cl}); }   // context.rs:470-481
};
static inline PureCommitment pure_of(const Digest &root, size_t nv) { PureCommitment p; p.root = root; p.num_vars = nv; p.is_base = true; return p; }

// Verifier::verify for Dense / Requant / ReLU chains.  `model_roots[id][poly]` are the setup commitments (the verifier's
// Context); input / output are the public IO.
static inline void zk_verify(const ZkContext &ctx, const std::vector<Element> &input, const std::vector<Element> &output, const ModelProof &proof, Transcript &t) {
    const Model &m = *ctx.model;
    for (auto &nk : ctx.model_comms) for (auto &pk : nk.second) digest_to_transcript(pk.second->comm.root(), t);         // ctx.write_to_transcript
    // ChallengeStorage::initialise (iop/mod.rs:70-90): one constant challenge, then one per table type in context order
    E constant_challenge = t.get_and_append_challenge("table_constant");
    std::map<TableType, E> challenge_map;
    for (auto &tt : ctx.tables) challenge_map[tt] = tt.kind == TT_RELU ? t.get_and_append_challenge("Relu") : (tt.kind == TT_CLAMPING ? t.get_and_append_challenge("Clamping") : E::one());
    // lookup numerators / denominators of every node proof and table proof (verifier.rs:88-111)
    std::vector<E> nums, dens;
    auto take = [&](const LogUpProof &p) { for (auto &e : p.circuit_outputs) { nums.push_back(e_add(e_mul(e[0], e[3]), e_mul(e[1], e[2]))); dens.push_back(e_mul(e[2], e[3])); } };
    for (size_t id = 0; id < m.nodes.size(); id++) {
        if (proof.requant.count(id)) { take(proof.requant.at(id).clamping_lookup); take(proof.requant.at(id).shifted_lookup); }
        if (proof.activation.count(id)) take(proof.activation.at(id).lookup);
        if (proof.pooling.count(id)) take(proof.pooling.at(id).lookup);
    }
    for (auto &tp : proof.table_proofs) take(tp.lookup);
    // output claim (compute_model_output_claims): point from the transcript, eval of the public output
    Claim last; for (size_t i = 0, nv = ceil_log2(output.size()); i < nv; i++) last.point.push_back(t.read_challenge());
    last.eval = mle_evaluate(*base_mle(to_base_vec(output)), last.point);
    CommitmentVerifier cv;
    std::map<size_t, bool> used_model;
    for (size_t id = m.nodes.size(); id-- > 0;) {
        const Node &n = m.nodes[id];
        if (n.kind == OP_DENSE) {                                                    // verify_dense (dense.rs:576-640)
            zk_ensure(proof.dense.count(id), "no dense proof for node " + std::to_string(id));
            const DenseProof &p = proof.dense.at(id);
            SumCheckSubClaim sub = sumcheck_verify(e_sub(last.eval, p.bias_eval), p.sumcheck, ceil_log2(n.ncols), 2, t);
            std::vector<E> wp = sub.point; wp.insert(wp.end(), last.point.begin(), last.point.end());
            const auto &comms = ctx.model_comms.at(id);                               // add_common_claims: BTreeMap order
            cv.add_witness_claim(pure_of(comms.at("DenseBias")->comm.root(), comms.at("DenseBias")->comm.num_vars), {last.point, p.bias_eval});
            cv.add_witness_claim(pure_of(comms.at("DenseWeight")->comm.root(), comms.at("DenseWeight")->comm.num_vars), {wp, p.individual_claims[0]});
            used_model[id] = true;
            E prod = E::one(); for (E e : p.individual_claims) prod = e_mul(prod, e);
            zk_ensure(prod == sub.expected_evaluation, "dense: sumcheck claim failed");
            last = {sub.point, p.individual_claims[1]};
        } else if (n.kind == OP_MATMUL) {                                            // verify_matmul (matrix_mul.rs:1048-1139)
            zk_ensure(proof.dense.count(id), "no matmul proof for node " + std::to_string(id));
            const DenseProof &p = proof.dense.at(id);
            const size_t vr = ceil_log2(n.mm_r), vc = ceil_log2(n.mm_c), vk = ceil_log2(n.mm_k);
            zk_ensure(last.point.size() == vr + vc, "matmul: wrong claim point length");
            const std::vector<E> p_right(last.point.begin(), last.point.begin() + vc), p_left(last.point.begin() + vc, last.point.end());
            const auto &comms = ctx.model_comms.at(id);
            E claim_eval = last.eval;
            if (n.mm_bias) { cv.add_witness_claim(pure_of(comms.at("MatMulBias")->comm.root(), comms.at("MatMulBias")->comm.num_vars), {p_right, p.bias_eval}); claim_eval = e_sub(claim_eval, p.bias_eval); }
            SumCheckSubClaim sub = sumcheck_verify(claim_eval, p.sumcheck, vk, 2, t);
            zk_ensure(p.individual_claims.size() == 2, "matmul: two individual claims expected");
            std::vector<E> pl = sub.point; pl.insert(pl.end(), p_left.begin(), p_left.end());
            std::vector<E> pr;
            if (n.mm_t) { pr = sub.point; pr.insert(pr.end(), p_right.begin(), p_right.end()); } else { pr = p_right; pr.insert(pr.end(), sub.point.begin(), sub.point.end()); }
            cv.add_witness_claim(pure_of(comms.at("MatMulWeight")->comm.root(), comms.at("MatMulWeight")->comm.num_vars), {pr, p.individual_claims[1]});
            used_model[id] = true;
            zk_ensure(e_mul(p.individual_claims[0], p.individual_claims[1]) == sub.expected_evaluation, "matmul: sumcheck claim failed");
            last = {pl, p.individual_claims[0]};
        } else if (n.kind == OP_REQUANT) {                                           // verify_requant (requant.rs:689-816)
            zk_ensure(proof.requant.count(id), "no requant proof for node " + std::to_string(id));
            const RequantProof &p = proof.requant.at(id);
            TableType tc{TT_CLAMPING, n.rq.clamping_size()};
            size_t shifted_instances = n.rq.shift() / Q_BIT_LEN;
            LogUpVerifierClaim cc = verify_logup_proof(p.clamping_lookup, 1, constant_challenge, challenge_map.at(tc), t);
            LogUpVerifierClaim sc = verify_logup_proof(p.shifted_lookup, shifted_instances, constant_challenge, E::one(), t);
            E bc = t.get_and_append_challenge("requant_batching");
            E init = E::zero(), ch = E::one();
            std::vector<E> vals = {last.eval, cc.claims[1].eval, cc.claims[0].eval}; for (auto &c : sc.claims) vals.push_back(c.eval);
            for (E v : vals) { init = e_add(init, e_mul(ch, v)); ch = e_mul(ch, bc); }
            SumCheckSubClaim sub = sumcheck_verify(init, p.io_accumulation, cc.point().size(), 2, t);
            const std::vector<E> &ap = sub.point; const std::vector<E> &ae = p.accumulation_evals;
            E lb = eq_eval(last.point, ap), cb = eq_eval(cc.point(), ap), sb = eq_eval(sc.point(), ap);
            E calc = e_mul(e_add(lb, e_mul(bc, cb)), ae[1]);
            E comb = e_mul(bc, bc);
            calc = e_add(calc, e_mul(e_mul(comb, cb), ae[0]));
            comb = e_mul(comb, bc);
            for (size_t k = 2; k < ae.size(); k++) { calc = e_add(calc, e_mul(e_mul(ae[k], sb), comb)); comb = e_mul(comb, bc); }
            zk_ensure(calc == sub.expected_evaluation, "requant: calculated claim does not line up with the expected claim");
            // recombine_claims (requant.rs:483-515)
            std::vector<E> sh(ae.begin() + 2, ae.end());
            E full = e_mul(E::from_u64((u64)1 << n.rq.shift()), ae[0]), p2 = E::one();
            for (E v : sh) { full = e_add(full, e_mul(v, p2)); p2 = e_mul(p2, E::from_u64((u64)1 << Q_BIT_LEN)); }
            E next_eval = e_mul(e_sub(full, E::from_u64((u64)1 << (n.rq.shift() - 1))), e_inv(E::from_base(f_from_i64(n.rq.fixed_point_multiplier))));
            zk_ensure(ae.size() == p.commitments.size(), "requant: commitments / evaluations mismatch");
            for (size_t k = 0; k < ae.size(); k++) cv.add_witness_claim(pure_of(p.commitments[k], ap.size()), {ap, ae[k]});
            last = {ap, next_eval};
        } else if (n.kind == OP_RELU) {                                              // verify_activation (activation.rs:459-512)
            zk_ensure(proof.activation.count(id), "no activation proof for node " + std::to_string(id));
            const ActivationProof &p = proof.activation.at(id);
            LogUpVerifierClaim vc = verify_logup_proof(p.lookup, 1, constant_challenge, challenge_map.at({TT_RELU, 0}), t);
            // same_poly::Verifier::verify (same_poly.rs:157-185) on claims [last_claim, lookup output column]
            std::vector<Claim> cl = {last}; for (size_t k = 1; k < vc.claims.size(); k++) cl.push_back(vc.claims[k]);
            std::vector<E> fs; for (size_t k = 0; k < cl.size(); k++) fs.push_back(t.read_challenge());
            E y = E::zero(); for (size_t k = 0; k < cl.size(); k++) y = e_add(y, e_mul(cl[k].eval, fs[k]));   // aggregated_rlc
            SumCheckSubClaim sub = sumcheck_verify(y, p.io_accumulation.sumcheck, cl[0].point.size(), 2, t);
            E cy = E::zero(); for (size_t k = 0; k < cl.size(); k++) cy = e_add(cy, e_mul(fs[k], eq_eval(cl[k].point, p.io_accumulation.sumcheck.point)));
            zk_ensure(cy == p.io_accumulation.evals[0], "same_poly: beta evaluation do not match");
            zk_ensure(e_mul(p.io_accumulation.evals[0], p.io_accumulation.evals[1]) == sub.expected_evaluation, "same_poly: final evals of sumcheck is not valid");
            Claim new_out = p.io_accumulation.extract_claim();
            zk_ensure(p.commits.size() == 2, "activation: expected two commitments");
            cv.add_witness_claim(pure_of(p.commits[0], vc.claims[0].point.size()), vc.claims[0]);
            cv.add_witness_claim(pure_of(p.commits[1], new_out.point.size()), new_out);
            last = vc.claims[0];
        } else if (n.kind == OP_CONV) {                                              // ConvCtx::verify -> verify_convolution
            zk_ensure(proof.conv.count(id), "no convolution proof for node " + std::to_string(id));
            Claim fc, bcl; Claim in_claim = verify_convolution(*n.conv, t, last, proof.conv.at(id), fc, bcl);
            const auto &comms = ctx.model_comms.at(id);                               // add_common_claims: ConvBias, ConvFilter (BTreeMap order)
            cv.add_witness_claim(pure_of(comms.at("ConvBias")->comm.root(), comms.at("ConvBias")->comm.num_vars), bcl);
            cv.add_witness_claim(pure_of(comms.at("ConvFilter")->comm.root(), comms.at("ConvFilter")->comm.num_vars), fc);
            used_model[id] = true;
            last = in_claim;
        } else if (n.kind == OP_POOL) {                                              // verify_pooling (pooling.rs:525-650)
            zk_ensure(proof.pooling.count(id), "no pooling proof for node " + std::to_string(id));
            const PoolingProof &p = proof.pooling.at(id);
            LogUpVerifierClaim vc = verify_logup_proof(p.lookup, 4, constant_challenge, E::one(), t);
            E bc = t.get_and_append_challenge("batch_pooling");
            E init = E::zero(), comb = bc;
            for (auto &c : vc.claims) { init = e_add(init, e_mul(c.eval, comb)); comb = e_mul(comb, bc); }
            init = e_add(init, e_mul(comb, last.eval));
            size_t nv = vc.point().size();
            SumCheckSubClaim sub = sumcheck_verify(init, p.sumcheck, nv, 5, t);
            const std::vector<E> &zc = sub.point;
            E beta_eval = eq_eval(vc.point(), zc), lcb = eq_eval(last.point, zc);
            size_t ks = p.zerocheck_evals.size() - 1;
            E prod = beta_eval, sum = E::zero(), ch = bc;
            for (size_t k = 0; k < ks; k++) { prod = e_mul(prod, p.zerocheck_evals[k]); sum = e_add(sum, e_mul(ch, p.zerocheck_evals[k])); ch = e_mul(ch, bc); }
            E output_eval = p.zerocheck_evals[ks];
            E expected = e_add(e_add(prod, e_mul(sum, beta_eval)), e_mul(e_mul(output_eval, lcb), ch));
            zk_ensure(expected == sub.expected_evaluation, "Expected pooling zerocheck claim did not equal the verifier claim");
            zk_ensure(p.commitments.size() == p.zerocheck_evals.size(), "pooling: commitments / evaluations mismatch");
            for (size_t k = 0; k < p.zerocheck_evals.size(); k++) cv.add_witness_claim(pure_of(p.commitments[k], zc.size()), {zc, p.zerocheck_evals[k]});
            E r1 = t.get_and_append_challenge("input_batching"), r2 = r1;             // `[challenge; 2]`
            E m1 = e_sub(E::one(), r1), m2 = e_sub(E::one(), r2);
            E mult[4] = {e_mul(m1, m2), e_mul(m1, r2), e_mul(r1, m2), e_mul(r1, r2)};
            Claim next; next.point.push_back(r1); next.point.insert(next.point.end(), zc.begin(), zc.begin() + p.variable_gap);
            next.point.push_back(r2); next.point.insert(next.point.end(), zc.begin() + p.variable_gap, zc.end());
            next.eval = E::zero(); for (size_t k = 0; k < ks; k++) next.eval = e_add(next.eval, e_mul(e_sub(output_eval, p.zerocheck_evals[k]), mult[k]));
            last = next;
        } else throw ZkVerifyError("zk_verify: unknown node kind");
    }
    // table proofs, zipped with the context's table order (verifier.rs:215-240, verify_table :320-383)
    zk_ensure(proof.table_proofs.size() == ctx.tables.size(), "number of table proofs != number of tables");
    for (size_t k = 0; k < ctx.tables.size(); k++) {
        const TableType &tt = ctx.tables[k]; const TableProof &tp = proof.table_proofs[k];
        LogUpVerifierClaim vc = verify_logup_proof(tp.lookup, 1, constant_challenge, challenge_map.at(tt), t);
        cv.add_witness_claim(pure_of(tp.multiplicity_commit, vc.claims[0].point.size()), vc.claims[0]);
        std::vector<E> exp = evaluate_table_columns(tt, vc.claims[0].point);
        zk_ensure(exp.size() == vc.claims.size() - 1, "table: wrong number of column evaluation claims");
        for (size_t c = 0; c < exp.size(); c++) zk_ensure(vc.claims[1 + c].eval == exp[c], "table: claimed table eval was wrong");
    }
    // input claim (verify_input_claim)
    zk_ensure(mle_evaluate(*base_mle(to_base_vec(input)), last.point) == last.eval, "input claim does not match the public input");
    // CommitmentVerifier::verify (commit/context.rs:522-600)
    for (auto &nk : ctx.model_comms) zk_ensure(used_model.count(nk.first), "Not all model commits have been used");
    zk_ensure(cv.trivial_claims.size() == proof.trivial_proofs.size(), "number of trivial proofs != number of trivial claims");
    for (size_t k = 0; k < cv.trivial_claims.size(); k++) { Transcript tt("default"); basefold_verify(ctx.full_log, cv.trivial_claims[k].comm, cv.trivial_claims[k].claim.point, cv.trivial_claims[k].claim.eval, proof.trivial_proofs[k], tt); }
    std::vector<PureCommitment> comms; std::vector<std::vector<E>> points; std::vector<Evaluation> evals;
    for (size_t i = 0; i < cv.claims.size(); i++) { comms.push_back(cv.claims[i].comm); points.push_back(cv.claims[i].claim.point); evals.push_back({i, i, cv.claims[i].claim.eval}); }
    basefold_batch_verify(ctx.full_log, comms, points, evals, proof.batch_proof, t);
    // accumulated lookup fractions: numerator zero, denominator non-zero (verifier.rs:268-287)
    E fn = E::zero(), fd = E::one();
    for (size_t i = 0; i < nums.size(); i++) { fn = e_add(e_mul(fn, dens[i]), e_mul(nums[i], fd)); fd = e_mul(fd, dens[i]); }
    zk_ensure(fn.is_zero(), "Final numerator was non-zero");
    zk_ensure(!fd.is_zero(), "Final denominator was zero, lookup arguments are invalid");
}

}  // namespace dpo
