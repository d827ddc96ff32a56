// TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's Basefold VERIFIER, so that opening proofs produced by the
// device (and by this oracle) are checked the way deep-prove checks them (commit -> open -> verify accepts,
// mpcs/src/basefold.rs:1239-1331 is the reference's own test of that).
//   Basefold::verify / batch_verify          mpcs/src/basefold.rs:863-1098
//   verifier_query_phase / batch_...         mpcs/src/basefold/query_phase.rs:141-283
//   SingleQueryResultWithMerklePath::check    query_phase.rs:915-975, Batched...::check :1116-1236, check_merkle_path :669
//   ClassicSumCheck::verify                   mpcs/src/sum_check/classic.rs:199-218,287-314
//   degree_2_zero_plus_one / degree_2_eval / interpolate2_weights   mpcs/src/util/arithmetic.rs:120-150
#pragma once
#include "basefold.hpp"

namespace dpo {

struct VerifyError : std::runtime_error { using std::runtime_error::runtime_error; };
struct PureCommitment { Digest root; size_t num_vars = 0; bool is_base = true; };   // BasefoldCommitment (structure.rs:140-190)

static inline E degree_2_zero_plus_one(const std::vector<E> &p) { return e_add(e_add(p[0], p[0]), e_add(p[1], p[2])); }
static inline E degree_2_eval(const std::vector<E> &p, E x) { return e_add(p[0], e_add(e_mul(x, p[1]), e_mul(e_mul(x, x), p[2]))); }
static inline E interpolate2_weights(E a0, E a1, E b0, E b1, E w, E x) { (void)b0; return e_add(a1, e_mul(e_mul(e_sub(x, a0), e_sub(b1, a1)), w)); }
static inline E eq_xy_eval(const std::vector<E> &x, const std::vector<E> &y) { return eq_eval(x, y); }

// check_merkle_path (query_phase.rs:669-700): leaf pair -> hash_or_noop -> compress up the path
static inline void check_merkle_path(const QueryOpening &q, const Digest &root, const char *what) {
    Digest cur;
    if (q.is_base) { u64 in[2] = {q.p0.c0, q.p1.c0}; cur = hash_or_noop(in, 2); }
    else { u64 in[4] = {q.p0.c0, q.p0.c1, q.p1.c0, q.p1.c1}; cur = hash_or_noop(in, 4); }
    size_t idx = q.index >> 1;
    for (const Digest &sib : q.path) { cur = (idx & 1) ? compress(sib, cur) : compress(cur, sib); idx >>= 1; }
    if (!(cur == root)) throw VerifyError(std::string("merkle path does not authenticate: ") + what);
}
static inline std::vector<E> final_codeword_of(std::vector<E> message, size_t full_log, bool batch) {
    // verifier_query_phase interpolates then bit-reverses (query_phase.rs:160-164); the batch variant bit-reverses first (:237-241)
    if (batch) { reverse_index_bits_in_place(message); interpolate_hc<E>(message, e_sub); }
    else { interpolate_hc<E>(message, e_sub); reverse_index_bits_in_place(message); }
    FVec cw = rs_encode(ext_fvec(message), full_log);
    reverse_index_bits_in_place(cw.e);
    return cw.e;
}
// replay of the commit phase on the verifier's transcript (basefold.rs:904-935)
static inline void replay_commit_phase(const CommitPhaseProof &cp, size_t num_rounds, size_t codeword_size, Transcript &t, std::vector<E> &fold_challenges, std::vector<size_t> &queries) {
    if (cp.sumcheck_messages.size() != num_rounds || cp.roots.size() + 1 != num_rounds) throw VerifyError("commit phase: wrong number of messages / roots");
    for (size_t i = 0; i < num_rounds; i++) {
        t.append_field_element_exts(cp.sumcheck_messages[i]);
        fold_challenges.push_back(t.get_and_append_challenge("commit round"));
        if (i + 1 < num_rounds) digest_to_transcript(cp.roots[i], t);
    }
    t.append_field_element_exts(cp.final_message);
    queries = query_indices(t, RS_NUM_QUERIES, codeword_size);
}
static inline void final_sumcheck_checks(const CommitPhaseProof &cp, const std::vector<E> &fold_challenges, const std::vector<E> &partial_eq, E eval) {
    if (!(eval == degree_2_zero_plus_one(cp.sumcheck_messages[0]))) throw VerifyError("claimed evaluation != p0(0) + p0(1)");
    for (size_t i = 0; i + 1 < fold_challenges.size(); i++)
        if (!(degree_2_eval(cp.sumcheck_messages[i], fold_challenges[i]) == degree_2_zero_plus_one(cp.sumcheck_messages[i + 1]))) throw VerifyError("basefold sumcheck consistency");
    E ip = E::zero(); for (size_t i = 0; i < cp.final_message.size(); i++) ip = e_add(ip, e_mul(cp.final_message[i], partial_eq[i]));
    if (!(degree_2_eval(cp.sumcheck_messages.back(), fold_challenges.back()) == ip)) throw VerifyError("last sumcheck message != <final_message, eq>");
}

// Basefold::verify (basefold.rs:863-962)
static inline void basefold_verify(size_t full_log, const PureCommitment &comm, const std::vector<E> &point, E eval, const BasefoldProof &proof, Transcript &t) {
    if (proof.trivial) {
        MerkleTree mt; mt.leaves = proof.trivial_evals; mt.inner = merkelize(proof.trivial_evals);
        if (!(mt.root() == comm.root)) throw VerifyError("MerkleRootMismatch");
        MLE m; m.is_ext = proof.trivial_evals.is_ext; m.num_vars = ceil_log2(proof.trivial_evals.len()); m.base = proof.trivial_evals.b; m.ext = proof.trivial_evals.e;
        if (!(mle_evaluate(m, point) == eval)) throw VerifyError("Trivial proof did not evaluate to the correct value");
        return;
    }
    size_t num_vars = point.size();
    if (num_vars != comm.num_vars || num_vars < RS_BASECODE_MSG_SIZE_LOG) throw VerifyError("point length != commitment num_vars");
    size_t num_rounds = num_vars - RS_BASECODE_MSG_SIZE_LOG;
    std::vector<E> fc; std::vector<size_t> queries;
    replay_commit_phase(proof.commit_phase, num_rounds, (size_t)1 << (num_vars + RS_RATE_LOG), t, fc, queries);
    std::vector<E> rev(fc.rbegin(), fc.rend());
    E coeff = eq_xy_eval(std::vector<E>(point.end() - fc.size(), point.end()), rev);
    std::vector<E> eq = build_eq_x_r_vec(std::vector<E>(point.begin(), point.end() - fc.size()));
    for (auto &e : eq) e = e_mul(e, coeff);
    // verifier_query_phase (query_phase.rs:141-211)
    std::vector<E> final_codeword = final_codeword_of(proof.commit_phase.final_message, full_log, false);
    if (proof.queries.size() != queries.size()) throw VerifyError("wrong number of query results");
    for (size_t qi = 0; qi < queries.size(); qi++) {
        const QueryResult &q = proof.queries[qi]; size_t index = queries[qi];
        if (q.x_index != index) throw VerifyError("query index does not match the transcript");
        if (q.oracle.size() + 1 != num_rounds) throw VerifyError("wrong number of oracle openings");
        for (size_t i = 0; i < q.oracle.size(); i++) check_merkle_path(q.oracle[i], proof.commit_phase.roots[i], "oracle");
        check_merkle_path(q.commitment, comm.root, "commitment");
        if (q.commitment.is_base != comm.is_base) throw VerifyError("commitment opening has the wrong field type");
        E left = q.commitment.p0, right = q.commitment.p1;
        size_t right_index = index | 1, left_index = right_index - 1;
        if (q.commitment.index != left_index) throw VerifyError("commitment opening at the wrong index");
        for (size_t i = 0; i < num_rounds; i++) {
            u64 x0, w; folding_coeffs(full_log, num_vars + RS_RATE_LOG - i - 1, left_index >> 1, x0, w);
            E res = interpolate2_weights(E::from_base(x0), left, E::from_base(f_neg(x0)), right, E::from_base(w), fc[i]);
            size_t next_index = right_index >> 1; E next;
            if (i + 1 < num_rounds) {
                right_index = next_index | 1; left_index = right_index - 1;
                if (q.oracle[i].index != left_index) throw VerifyError("oracle opening at the wrong index");
                left = q.oracle[i].p0; right = q.oracle[i].p1;
                next = (next_index & 1) == 0 ? left : right;
            } else next = final_codeword[next_index];
            if (!(res == next)) throw VerifyError("fold consistency failed at round " + std::to_string(i));
        }
    }
    final_sumcheck_checks(proof.commit_phase, fc, eq, eval);
}

// Basefold::batch_verify (basefold.rs:964-1098)
static inline void basefold_batch_verify(size_t full_log, const std::vector<PureCommitment> &comms, const std::vector<std::vector<E>> &points,
                                         const std::vector<Evaluation> &evals, const BasefoldProof &proof, Transcript &t) {
    if (comms.empty() && points.empty() && evals.empty()) return;
    size_t num_vars = 0; for (auto &p : points) num_vars = std::max(num_vars, p.size());
    size_t num_rounds = num_vars - RS_BASECODE_MSG_SIZE_LOG;
    for (auto &e : evals) if (points[e.point].size() != comms[e.poly].num_vars) throw VerifyError("evaluation point length != polynomial num_vars");
    if (proof.trivial) throw VerifyError("batch proof cannot be trivial");
    size_t bsl = ceil_log2(evals.size());
    std::vector<E> tt; for (size_t i = 0; i < bsl; i++) tt.push_back(t.get_and_append_challenge("batch coeffs"));
    std::vector<E> eq_xt = build_eq_x_r_vec(tt);
    E target = E::zero();
    for (size_t i = 0; i < evals.size(); i++) target = e_add(target, e_mul(e_mul(evals[i].value, E::from_base(f_from_u64((u64)1 << (num_vars - points[evals[i].point].size())))), eq_xt[i]));
    // ClassicSumCheck::verify (classic.rs:287-314 + verify_consistency :199-218), coefficient-form messages
    if (proof.sumcheck_proof.size() != num_vars) throw VerifyError("classic sumcheck: wrong number of rounds");
    std::vector<E> verify_point; E sum = target;
    for (size_t i = 0; i < num_vars; i++) { t.append_field_element_exts(proof.sumcheck_proof[i]); verify_point.push_back(t.get_and_append_challenge("sumcheck round")); }
    for (size_t i = 0; i < num_vars; i++) {
        if (!(sum == degree_2_zero_plus_one(proof.sumcheck_proof[i]))) throw VerifyError(i == 0 ? "classic sumcheck: wrong initial sum" : "classic sumcheck: consistency failure");
        sum = degree_2_eval(proof.sumcheck_proof[i], verify_point[i]);
    }
    E new_target = sum;
    std::vector<E> eq_xy; for (auto &p : points) eq_xy.push_back(eq_xy_eval(std::vector<E>(verify_point.begin(), verify_point.begin() + p.size()), p));
    std::vector<E> coeffs(comms.size(), E::zero());
    for (size_t i = 0; i < evals.size(); i++) coeffs[evals[i].poly] = e_add(coeffs[evals[i].poly], e_mul(eq_xy[evals[i].point], eq_xt[i]));
    std::vector<E> fc; std::vector<size_t> queries;
    replay_commit_phase(proof.commit_phase, num_rounds, (size_t)1 << (num_vars + RS_RATE_LOG), t, fc, queries);
    std::vector<E> rev(fc.rbegin(), fc.rend());
    E coeff = eq_xy_eval(std::vector<E>(verify_point.end() - fc.size(), verify_point.end()), rev);
    std::vector<E> eq = build_eq_x_r_vec(std::vector<E>(verify_point.begin(), verify_point.end() - fc.size()));
    for (auto &e : eq) e = e_mul(e, coeff);
    // batch_verifier_query_phase (query_phase.rs:213-283) + BatchedSingleQueryResultWithMerklePath::check (:1116-1236)
    std::vector<E> final_codeword = final_codeword_of(proof.commit_phase.final_message, full_log, true);
    if (proof.batched_queries.size() != queries.size()) throw VerifyError("wrong number of query results");
    for (size_t qi = 0; qi < queries.size(); qi++) {
        const BatchedQueryResult &q = proof.batched_queries[qi]; size_t index = queries[qi];
        if (q.x_index != index) throw VerifyError("query index does not match the transcript");
        if (q.oracle.size() + 1 != num_rounds || q.commitments.size() != comms.size()) throw VerifyError("wrong number of openings");
        for (size_t i = 0; i < q.oracle.size(); i++) check_merkle_path(q.oracle[i], proof.commit_phase.roots[i], "oracle");
        for (size_t i = 0; i < comms.size(); i++) { check_merkle_path(q.commitments[i], comms[i].root, "commitment"); if (q.commitments[i].is_base != comms[i].is_base) throw VerifyError("commitment opening has the wrong field type"); }
        E left = E::zero(), right = E::zero();
        size_t right_index = index | 1, left_index = right_index - 1;
        for (size_t i = 0; i < num_rounds; i++) {
            for (size_t c = 0; c < comms.size(); c++) if (comms[c].num_vars == num_vars - i) {
                if ((q.commitments[c].index >> 1) != (left_index >> 1)) throw VerifyError("commitment opening at the wrong index");
                left = e_add(left, e_mul(q.commitments[c].p0, coeffs[c])); right = e_add(right, e_mul(q.commitments[c].p1, coeffs[c]));
            }
            u64 x0, w; folding_coeffs(full_log, num_vars + RS_RATE_LOG - i - 1, left_index >> 1, x0, w);
            E res = interpolate2_weights(E::from_base(x0), left, E::from_base(f_neg(x0)), right, E::from_base(w), fc[i]);
            size_t next_index = right_index >> 1; E next;
            if (i + 1 < num_rounds) {
                right_index = next_index | 1; left_index = right_index - 1;
                if (q.oracle[i].index != left_index) throw VerifyError("oracle opening at the wrong index");
                left = q.oracle[i].p0; right = q.oracle[i].p1;
                next = (next_index & 1) == 0 ? left : right;
            } else {
                for (size_t c = 0; c < comms.size(); c++) if (comms[c].num_vars == num_vars - i - 1) {
                    if ((q.commitments[c].index >> 1) != (next_index >> 1)) throw VerifyError("commitment opening at the wrong index (last round)");
                    res = e_add(res, e_mul((next_index & 1) == 0 ? q.commitments[c].p0 : q.commitments[c].p1, coeffs[c]));
                }
                next = final_codeword[next_index];
            }
            if (!(res == next)) throw VerifyError("batched fold consistency failed at round " + std::to_string(i));
        }
    }
    final_sumcheck_checks(proof.commit_phase, fc, eq, new_target);
}

// inverse of flatten_proof (basefold.hpp), so that the DEVICE's flat proof image is what gets verified
static inline BasefoldProof unflatten_proof(const u64 *p, size_t n) {
    size_t k = 0;
    auto u = [&]() -> u64 { if (k >= n) throw VerifyError("flat proof truncated"); return p[k++]; };
    auto e = [&]() { u64 a = u(), b = u(); return E(a, b); };
    auto d = [&]() { Digest x; for (int i = 0; i < 4; i++) x.v[i] = u(); return x; };
    auto q = [&]() { QueryOpening o; o.index = u(); o.is_base = u() != 0; if (o.is_base) { o.p0 = E::from_base(u()); o.p1 = E::from_base(u()); } else { o.p0 = e(); o.p1 = e(); } size_t np = u(); for (size_t i = 0; i < np; i++) o.path.push_back(d()); return o; };
    BasefoldProof pr;
    size_t ns = u(); for (size_t i = 0; i < ns; i++) { std::vector<E> m; for (int j = 0; j < 3; j++) m.push_back(e()); pr.sumcheck_proof.push_back(m); }
    size_t nm = u(); for (size_t i = 0; i < nm; i++) { std::vector<E> m; for (int j = 0; j < 3; j++) m.push_back(e()); pr.commit_phase.sumcheck_messages.push_back(m); }
    size_t nr = u(); for (size_t i = 0; i < nr; i++) pr.commit_phase.roots.push_back(d());
    size_t nf = u(); for (size_t i = 0; i < nf; i++) pr.commit_phase.final_message.push_back(e());
    size_t nq = u(); for (size_t i = 0; i < nq; i++) { QueryResult r; r.x_index = u(); r.commitment = q(); size_t no = u(); for (size_t j = 0; j < no; j++) r.oracle.push_back(q()); pr.queries.push_back(r); }
    size_t nb = u(); for (size_t i = 0; i < nb; i++) { BatchedQueryResult r; r.x_index = u(); size_t no = u(); for (size_t j = 0; j < no; j++) r.oracle.push_back(q()); size_t nc = u(); for (size_t j = 0; j < nc; j++) r.commitments.push_back(q()); pr.batched_queries.push_back(r); }
    if (k != n) throw VerifyError("flat proof has trailing words");
    return pr;
}


// Basefold::simple_batch_verify (basefold.rs:1100-1202) + simple_batch_verifier_query_phase (query_phase.rs:285-372) +
// SimpleBatchSingleQueryResultWithMerklePath::check (:1468-1535) + authenticate_batch_leaves_root (merkle path of batch leaves)
static inline SimpleBatchProof unflatten_simple_batch_proof(const u64 *p, size_t n) {
    size_t k = 0;
    auto u = [&]() -> u64 { if (k >= n) throw VerifyError("flat proof truncated"); return p[k++]; };
    auto e = [&]() { u64 a = u(), b = u(); return E(a, b); };
    auto d = [&]() { Digest x; for (int i = 0; i < 4; i++) x.v[i] = u(); return x; };
    auto q = [&]() { QueryOpening o; o.index = u(); o.is_base = u() != 0; if (o.is_base) { o.p0 = E::from_base(u()); o.p1 = E::from_base(u()); } else { o.p0 = e(); o.p1 = e(); } size_t np = u(); for (size_t i = 0; i < np; i++) o.path.push_back(d()); return o; };
    SimpleBatchProof pr;
    size_t nm = u(); for (size_t i = 0; i < nm; i++) { std::vector<E> m; for (int j = 0; j < 3; j++) m.push_back(e()); pr.commit_phase.sumcheck_messages.push_back(m); }
    size_t nr = u(); for (size_t i = 0; i < nr; i++) pr.commit_phase.roots.push_back(d());
    size_t nf = u(); for (size_t i = 0; i < nf; i++) pr.commit_phase.final_message.push_back(e());
    size_t nq = u();
    for (size_t i = 0; i < nq; i++) {
        SimpleBatchQueryResult r; r.x_index = u(); r.index = u(); r.is_base = u() != 0; size_t m = u();
        for (size_t j = 0; j < m; j++) { if (r.is_base) { r.left.push_back(E::from_base(u())); r.right.push_back(E::from_base(u())); } else { r.left.push_back(e()); r.right.push_back(e()); } }
        size_t np = u(); for (size_t j = 0; j < np; j++) r.path.push_back(d());
        size_t no = u(); for (size_t j = 0; j < no; j++) r.oracle.push_back(q());
        pr.queries.push_back(r);
    }
    if (k != n) throw VerifyError("flat proof has trailing words");
    return pr;
}
static inline void basefold_simple_batch_verify(size_t full_log, const PureCommitment &comm, size_t num_polys, const std::vector<E> &point, const std::vector<E> &evals, const SimpleBatchProof &proof, Transcript &t) {
    if (evals.size() != num_polys) throw VerifyError("number of evaluations != number of committed polynomials");
    if (proof.trivial) { if (!(merkelize_batch(proof.trivial_evals).back()[0] == comm.root)) throw VerifyError("MerkleRootMismatch"); return; }
    size_t num_vars = point.size();
    if (num_vars != comm.num_vars || num_vars < RS_BASECODE_MSG_SIZE_LOG) throw VerifyError("point length != commitment num_vars");
    size_t num_rounds = num_vars - RS_BASECODE_MSG_SIZE_LOG;
    size_t bsl = ceil_log2(evals.size());
    std::vector<E> tt; for (size_t i = 0; i < bsl; i++) tt.push_back(t.get_and_append_challenge("batch coeffs"));
    std::vector<E> eq_xt = build_eq_x_r_vec(tt); eq_xt.resize(evals.size());
    std::vector<E> fc; std::vector<size_t> queries;
    replay_commit_phase(proof.commit_phase, num_rounds, (size_t)1 << (num_vars + RS_RATE_LOG), t, fc, queries);
    std::vector<E> rev(fc.rbegin(), fc.rend());
    E coeff = eq_xy_eval(std::vector<E>(point.end() - fc.size(), point.end()), rev);
    std::vector<E> eq = build_eq_x_r_vec(std::vector<E>(point.begin(), point.end() - fc.size()));
    for (auto &e : eq) e = e_mul(e, coeff);
    std::vector<E> final_codeword = final_codeword_of(proof.commit_phase.final_message, full_log, true);
    if (proof.queries.size() != queries.size()) throw VerifyError("wrong number of query results");
    for (size_t qi = 0; qi < queries.size(); qi++) {
        const SimpleBatchQueryResult &q = proof.queries[qi]; size_t index = queries[qi];
        if (q.x_index != index || q.left.size() != num_polys || q.right.size() != num_polys || q.is_base != comm.is_base) throw VerifyError("query result does not match the commitment / transcript");
        if (q.oracle.size() + 1 != num_rounds) throw VerifyError("wrong number of oracle openings");
        for (size_t i = 0; i < q.oracle.size(); i++) check_merkle_path(q.oracle[i], proof.commit_phase.roots[i], "oracle");
        {   // authenticate_batch_leaves_root: compress(hash(left values), hash(right values)) then up the path
            std::vector<u64> a, b;
            for (size_t k = 0; k < num_polys; k++) { if (q.is_base) { a.push_back(q.left[k].c0); b.push_back(q.right[k].c0); } else { a.push_back(q.left[k].c0); a.push_back(q.left[k].c1); b.push_back(q.right[k].c0); b.push_back(q.right[k].c1); } }
            Digest cur = num_polys == 1 ? (q.is_base ? hash_or_noop(std::vector<u64>{a[0], b[0]}.data(), 2) : hash_or_noop(std::vector<u64>{a[0], a[1], b[0], b[1]}.data(), 4))
                                        : compress(hash_or_noop(a.data(), a.size()), hash_or_noop(b.data(), b.size()));
            size_t idx = q.index >> 1;
            for (const Digest &sib : q.path) { cur = (idx & 1) ? compress(sib, cur) : compress(cur, sib); idx >>= 1; }
            if (!(cur == comm.root)) throw VerifyError("merkle path does not authenticate: batch commitment");
        }
        E left = E::zero(), right = E::zero();                   // leaves.batch(batch_coeffs)
        for (size_t k = 0; k < num_polys; k++) { left = e_add(left, e_mul(q.left[k], eq_xt[k])); right = e_add(right, e_mul(q.right[k], eq_xt[k])); }
        size_t right_index = index | 1, left_index = right_index - 1;
        if (q.index != left_index) throw VerifyError("commitment opening at the wrong index");
        for (size_t i = 0; i < num_rounds; i++) {
            u64 x0, w; folding_coeffs(full_log, num_vars + RS_RATE_LOG - i - 1, left_index >> 1, x0, w);
            E res = interpolate2_weights(E::from_base(x0), left, E::from_base(f_neg(x0)), right, E::from_base(w), fc[i]);
            size_t next_index = right_index >> 1; E next;
            if (i + 1 < num_rounds) {
                right_index = next_index | 1; left_index = right_index - 1;
                if (q.oracle[i].index != left_index) throw VerifyError("oracle opening at the wrong index");
                left = q.oracle[i].p0; right = q.oracle[i].p1;
                next = (next_index & 1) == 0 ? left : right;
            } else next = final_codeword[next_index];
            if (!(res == next)) throw VerifyError("simple-batch fold consistency failed at round " + std::to_string(i));
        }
    }
    E claimed = E::zero(); for (size_t k = 0; k < evals.size(); k++) claimed = e_add(claimed, e_mul(eq_xt[k], evals[k]));   // inner_product(batch_coeffs, evals)
    final_sumcheck_checks(proof.commit_phase, fc, eq, claimed);
}

}  // namespace dpo
