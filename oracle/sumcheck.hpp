// ORACLE -- TEST INFRASTRUCTURE ONLY (see field.hpp header).
// Restates sumcheck/src/{prover.rs,util.rs,structs.rs,verifier.rs} + sumcheck_macro/src/lib.rs.
#pragma once
#include "mle.hpp"
#include "poseidon.hpp"
#include <functional>
#include <atomic>

namespace dpo {

// VirtualPolynomial = sum_i c_i * prod_j f_ij  (virtual_poly.rs:50-60).  MLE identity is by index
// here (the reference dedupes by Arc pointer, virtual_poly.rs:168-177): shared MLEs are folded once.
struct VirtualPolynomial {
    size_t max_num_variables = 0;
    size_t max_degree = 0;
    std::vector<std::pair<E, std::vector<size_t>>> products;
    std::vector<std::shared_ptr<MLE>> mles;
    explicit VirtualPolynomial(size_t nv = 0) : max_num_variables(nv) {}
    size_t add_mle(std::shared_ptr<MLE> m) {
        for (size_t i = 0; i < mles.size(); i++) if (mles[i].get() == m.get()) return i;
        mles.push_back(m); return mles.size() - 1;
    }
    // add_mle_list (virtual_poly.rs:139-180)
    void add_mle_list(const std::vector<std::shared_ptr<MLE>> &list, E coef) {
        if (list.empty()) throw std::runtime_error("input mle_list is empty");
        std::vector<size_t> idx;
        for (auto &m : list) {
            if (m->num_vars > max_num_variables) throw std::runtime_error("invalid max num vars");
            if (m->num_vars != list[0]->num_vars) throw std::runtime_error("product mles must share num_vars");
            idx.push_back(add_mle(m));
        }
        if (list.size() > max_degree) max_degree = list.size();
        products.push_back({coef, idx});
    }
    E evaluate(const std::vector<E> &point) const {
        E acc = E::zero();
        for (auto &pr : products) {
            E t = pr.first;
            for (size_t i : pr.second) {
                std::vector<E> p(point.begin(), point.begin() + mles[i]->num_vars);
                t = e_mul(t, mle_evaluate(*mles[i], p));
            }
            acc = e_add(acc, t);
        }
        return acc;
    }
};

static inline size_t ceil_log2(size_t x) { size_t l = 0; while (((size_t)1 << l) < x) l++; return l; }

// sumcheck_code_gen!(degree, _, accessor) (sumcheck_macro/src/lib.rs:46-326), one product, one round.
// `round` is the 1-based round counter AFTER the increment at prover.rs:686 (the macro reads self.round).
// Per pair b (even): operand j evaluated at t = 0,1,2,..,d is v[b], v[b+1], v[b+1]+c, v[b+1]+2c, ...
// with c = v[b+1]-v[b]; out[t] += prod_j v_j(t).  A length-1 operand list contributes its constant
// at every point (:236-241); the result is scaled by 2^(max_nv - (max(ceil_log2(len),1) + round - 1))
// (:242-247).  The macro iterates over the FIRST operand's length after its Ext-first stable sort.
static inline std::vector<E> sumcheck_product_round(const std::vector<const MLE *> &ops_in, size_t max_nv, size_t round) {
    size_t d = ops_in.size();
    std::vector<const MLE *> ops;
    for (auto m : ops_in) if (m->is_ext) ops.push_back(m);
    for (auto m : ops_in) if (!m->is_ext) ops.push_back(m);
    size_t len = ops[0]->len();
    std::vector<E> acc(d + 1, E::zero());
    if (len == 1) {
        E p = E::one();
        for (auto m : ops) p = e_mul(p, m->get(0));
        for (size_t t = 0; t <= d; t++) acc[t] = p;
    } else {
        size_t npairs = len >> 1;
        unsigned T = dpo_threads();
        std::vector<std::vector<E>> part(T, std::vector<E>(d + 1, E::zero()));
        std::atomic<unsigned> slot{0};
        par_for(npairs, 2048, [&](size_t pb, size_t pe) {
            unsigned me = slot.fetch_add(1); std::vector<E> &a = part[me];
            E cur[5], step[5];
            for (size_t pi = pb; pi < pe; pi++) {
                size_t b = 2 * pi;
                for (size_t j = 0; j < d; j++) { cur[j] = ops[j]->get(b); step[j] = e_sub(ops[j]->get(b + 1), cur[j]); }
                for (size_t t = 0; t <= d; t++) {
                    E p = cur[0];
                    for (size_t j = 1; j < d; j++) p = e_mul(p, cur[j]);
                    a[t] = e_add(a[t], p);
                    for (size_t j = 0; j < d; j++) cur[j] = e_add(cur[j], step[j]);
                }
            }
        });
        for (auto &a : part) for (size_t t = 0; t <= d; t++) acc[t] = e_add(acc[t], a[t]);
    }
    size_t l2 = ceil_log2(len); if (l2 < 1) l2 = 1;
    size_t mult = max_nv - (l2 + round - 1);
    if (mult > 0) { u64 s = f_from_u64((u64)1 << mult); for (auto &a : acc) a = e_mul_base(a, s); }
    return acc;
}

// barycentric_weights / extrapolate (sumcheck/src/util.rs:19-136), points 0..k
static inline E extrapolate_uni(const std::vector<E> &evals, u64 at) {
    size_t n = evals.size();
    std::vector<E> w(n);
    for (size_t j = 0; j < n; j++) {
        E p = E::one();
        for (size_t i = 0; i < n; i++) if (i != j) p = e_mul(p, e_sub(E::from_u64(j), E::from_u64(i)));
        w[j] = e_inv(p);
    }
    E sum = E::zero(); std::vector<E> co(n);
    for (size_t j = 0; j < n; j++) { co[j] = e_mul(e_inv(e_sub(E::from_u64(at), E::from_u64(j))), w[j]); sum = e_add(sum, co[j]); }
    E sum_inv = sum.is_zero() ? E::zero() : e_inv(sum);
    E r = E::zero();
    for (size_t j = 0; j < n; j++) r = e_add(r, e_mul(co[j], evals[j]));
    return e_mul(r, sum_inv);
}

struct IOPProof {
    std::vector<E> point;
    std::vector<std::vector<E>> proofs;  // per round: max_degree+1 evaluations at 0..max_degree
    E extract_sum() const { return e_add(proofs[0][0], proofs[0][1]); }  // structs.rs:19
};

struct IOPProverState {
    VirtualPolynomial poly;               // MLEs are private copies once folded
    std::vector<std::shared_ptr<MLE>> work;  // current (folded) MLEs
    std::vector<E> challenges;
    size_t round = 0;

    explicit IOPProverState(const VirtualPolynomial &vp) : poly(vp), work(vp.mles) {
        if (vp.max_num_variables == 0) throw std::runtime_error("Attempt to prove a constant.");
    }
    // prove_round_and_update_state(_parallel) (prover.rs:351-470 / :625-741)
    std::vector<E> prove_round(const E *challenge) {
        if (round >= poly.max_num_variables) throw std::runtime_error("Prover is not active");
        if (round == 0) { if (challenge) throw std::runtime_error("first round should be prover first."); }
        else {
            if (!challenge) throw std::runtime_error("verifier message is empty");
            challenges.push_back(*challenge);
            E r = challenges[round - 1];
            for (auto &m : work) {
                if (challenges.size() == 1) {
                    if (m->num_vars == 0) throw std::runtime_error("calling sumcheck on constant");
                    auto c = std::make_shared<MLE>(*m); mle_fix_low_one(*c, r); m = c;   // fix_variables: new instance
                } else if (m->num_vars > 0) mle_fix_low_one(*m, r);                        // in place
            }
        }
        round++;
        std::vector<E> msg(poly.max_degree + 1, E::zero());
        for (auto &pr : poly.products) {
            size_t d = pr.second.size();
            if (d > 5) throw std::runtime_error("do not support degree > 5");
            std::vector<const MLE *> ops;
            for (size_t i : pr.second) ops.push_back(work[i].get());
            std::vector<E> sum = sumcheck_product_round(ops, poly.max_num_variables, round);
            for (auto &s : sum) s = e_mul(s, pr.first);
            std::vector<E> base = sum;
            for (size_t i = 0; i < poly.max_degree - d; i++) sum.push_back(extrapolate_uni(base, d + 1 + i));
            for (size_t t = 0; t <= poly.max_degree; t++) msg[t] = e_add(msg[t], sum[t]);
        }
        return msg;
    }
    // tail of prove_parallel (prover.rs:544-568): push last challenge, fix every MLE once more
    void finish(E last) {
        challenges.push_back(last);
        bool first = challenges.size() == 1;
        for (auto &m : work) {
            if (m->num_vars > 0) { if (first) { auto c = std::make_shared<MLE>(*m); mle_fix_low_one(*c, last); m = c; } else mle_fix_low_one(*m, last); }
        }
    }
    // get_mle_final_evaluations (prover.rs:474-490)
    std::vector<E> final_evaluations() const {
        std::vector<E> v;
        for (auto &m : work) { if (m->len() != 1) throw std::runtime_error("mle.evaluations.len() != 1"); v.push_back(m->get(0)); }
        return v;
    }
};

// IOPProverState::prove_parallel (prover.rs:498-585)
static inline std::pair<IOPProof, std::vector<E>> sumcheck_prove(const VirtualPolynomial &vp, Transcript &t) {
    IOPProof proof;
    if (vp.max_num_variables == 0) return {proof, {}};
    t.append_usize(vp.max_num_variables);
    t.append_usize(vp.max_degree);
    IOPProverState st(vp);
    E chal; bool have = false;
    for (size_t i = 0; i < vp.max_num_variables; i++) {
        std::vector<E> msg = st.prove_round(have ? &chal : nullptr);
        t.append_field_element_exts(msg);
        proof.proofs.push_back(msg);
        chal = t.get_and_append_challenge("Internal round"); have = true;
    }
    st.finish(chal);
    proof.point = st.challenges;
    return {proof, st.final_evaluations()};
}

// IOPProverState::prove_batch_polys (prover.rs:37-321) + merge_sumcheck_polys (util.rs:215-243): the "devirgo"
// split.  polys[t] is thread t's slice (the top log T variables fixed to t); every round the T messages are summed
// before Fiat-Shamir; after num_variables rounds the T residual values per MLE form a log T-variable polynomial that
// is finished with log T ordinary rounds.  Produces the same proof as sumcheck_prove on the un-split polynomial.
static inline std::pair<IOPProof, std::vector<E>> sumcheck_prove_batch_polys(const std::vector<VirtualPolynomial> &polys, Transcript &t) {
    size_t T = polys.size();
    if (T == 0 || (T & (T - 1))) throw std::runtime_error("prove_batch_polys: number of polys must be a power of two");
    size_t logT = ceil_log2(T), nv = polys[0].max_num_variables, deg = polys[0].max_degree;
    for (auto &p : polys) if (p.max_num_variables != nv || p.max_degree != deg) throw std::runtime_error("prove_batch_polys: polys differ in (num_variables, degree)");
    IOPProof proof;
    if (nv == 0) return {proof, {}};
    t.append_usize(nv + logT);
    t.append_usize(deg);
    std::vector<IOPProverState> st; for (auto &p : polys) st.emplace_back(p);
    E chal; bool have = false;
    for (size_t i = 0; i < nv; i++) {
        std::vector<E> msg(deg + 1, E::zero());
        for (auto &s : st) { auto m = s.prove_round(have ? &chal : nullptr); for (size_t k = 0; k <= deg; k++) msg[k] = e_add(msg[k], m[k]); }
        t.append_field_element_exts(msg);
        proof.proofs.push_back(msg);
        chal = t.get_and_append_challenge("Internal round"); have = true;
    }
    for (auto &s : st) s.finish(chal);
    std::vector<E> point = st[0].challenges;
    if (logT == 0) { proof.point = point; return {proof, st[0].final_evaluations()}; }
    // merge_sumcheck_polys
    VirtualPolynomial merged(logT); merged.max_degree = deg; merged.products = polys[0].products;
    size_t n_mles = polys[0].mles.size();
    for (size_t i = 0; i < n_mles; i++) { std::vector<E> v; for (auto &s : st) v.push_back(s.work[i]->get(0)); merged.mles.push_back(std::make_shared<MLE>(MLE::from_ext(logT, v))); }
    IOPProverState s2(merged);
    have = false;
    for (size_t i = 0; i < logT; i++) {
        auto msg = s2.prove_round(have ? &chal : nullptr);
        t.append_field_element_exts(msg);
        proof.proofs.push_back(msg);
        chal = t.get_and_append_challenge("Internal round"); have = true;
    }
    s2.finish(chal);
    point.insert(point.end(), s2.challenges.begin(), s2.challenges.end());
    proof.point = point;
    return {proof, s2.final_evaluations()};
}

// interpolate_uni_poly (util.rs:148-199): evaluate the degree-(n-1) poly through (i, p_i) at x
static inline E interpolate_uni_poly(const std::vector<E> &p, E x) {
    size_t n = p.size();
    E res = E::zero();
    for (size_t i = 0; i < n; i++) {
        E num = E::one(), den = E::one();
        for (size_t j = 0; j < n; j++) if (j != i) { num = e_mul(num, e_sub(x, E::from_u64(j))); den = e_mul(den, e_sub(E::from_u64(i), E::from_u64(j))); }
        res = e_add(res, e_mul(p[i], e_mul(num, e_inv(den))));
    }
    return res;
}

struct SumCheckSubClaim { std::vector<E> point; E expected_evaluation; };
// IOPVerifierState::verify (verifier.rs:12-169)
static inline SumCheckSubClaim sumcheck_verify(E claimed_sum, const IOPProof &proof, size_t num_vars, size_t max_degree, Transcript &t) {
    SumCheckSubClaim sc;
    if (num_vars == 0) { sc.expected_evaluation = claimed_sum; return sc; }
    t.append_usize(num_vars);
    t.append_usize(max_degree);
    if (proof.proofs.size() != num_vars) throw std::runtime_error("sumcheck: wrong number of rounds");
    E expected = claimed_sum;
    for (size_t i = 0; i < num_vars; i++) {
        const std::vector<E> &msg = proof.proofs[i];
        if (msg.size() != max_degree + 1) throw std::runtime_error("sumcheck: wrong message length");
        t.append_field_element_exts(msg);
        E r = t.get_and_append_challenge("Internal round");
        if (e_add(msg[0], msg[1]) != expected) throw std::runtime_error("sumcheck: p(0)+p(1) != claim");
        expected = interpolate_uni_poly(msg, r);
        sc.point.push_back(r);
    }
    sc.expected_evaluation = expected;
    return sc;
}

}  // namespace dpo
