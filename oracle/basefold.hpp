// ORACLE -- TEST INFRASTRUCTURE ONLY (see field.hpp header).
// Restates the prover side of mpcs' Basefold over RS code + Poseidon Merkle trees:
//   mpcs/src/basefold.rs:86-154,304-354,466-770; basefold/{commit_phase.rs,sumcheck.rs,encoding/rs.rs,
//   query_phase.rs:31-138,373-534}; util/{merkle_tree.rs,hash.rs,arithmetic.rs:120-132,
//   arithmetic/hypercube.rs,plonky2_util}; sum_check/classic{.rs,/coeff.rs}.
// Field constants (p3-goldilocks, not vendored): GENERATOR = 7, two_adic_generator(bits) =
// 1753635133440165772^(2^(32-bits)) (= 7^((p-1)/2^32), order 2^32 -- checked in tests/test_oracle_core.py).
#pragma once
#include "sumcheck.hpp"
#include <algorithm>

namespace dpo {

static const u64 GL_GENERATOR = 7;
static const u64 GL_TWO_ADIC_ROOT = 1753635133440165772ULL;  // order 2^32
static inline u64 two_adic_generator(size_t bits) { u64 g = GL_TWO_ADIC_ROOT; for (size_t i = bits; i < 32; i++) g = f_mul(g, g); return g; }
static inline u64 f_exp_pow2(u64 a, size_t k) { for (size_t i = 0; i < k; i++) a = f_mul(a, a); return a; }

static inline size_t reverse_bits(size_t x, size_t bits) { size_t r = 0; for (size_t i = 0; i < bits; i++) if (x >> i & 1) r |= (size_t)1 << (bits - 1 - i); return r; }
template <class T> static inline void reverse_index_bits_in_place(std::vector<T> &v) {
    size_t n = v.size(), lg = ceil_log2(n);
    for (size_t i = 0; i < n; i++) { size_t j = reverse_bits(i, lg); if (i < j) std::swap(v[i], v[j]); }
}

// RS code spec (encoding/rs.rs:192-214)
static const size_t RS_NUM_QUERIES = 200, RS_RATE_LOG = 1, RS_BASECODE_MSG_SIZE_LOG = 7;

// A vector that is Base or Ext (FieldType<E>)
struct FVec {
    bool is_ext = false;
    std::vector<u64> b; std::vector<E> e;
    size_t len() const { return is_ext ? e.size() : b.size(); }
    E get(size_t i) const { return is_ext ? e[i] : E::from_base(b[i]); }
};

// interpolate_over_boolean_hypercube (hypercube.rs:16-37): evals -> multilinear coefficients
template <class T, class Sub> static inline void interpolate_hc(std::vector<T> &v, Sub sub) {
    size_t n = v.size();
    for (size_t half = 1; half < n; half <<= 1)
        for (size_t k = 0; k < n; k += 2 * half)
            for (size_t j = k + half; j < k + 2 * half; j++) v[j] = sub(v[j], v[j - half]);
}

// RSCode::encode_internal + coset_fft + fft (rs.rs:129-189,458-501): codeword[i] = poly(shift * w^i),
// w = two_adic_generator(lg(2m)), shift = 7^(2^(full_log - lg m)).  (rs.rs:540-556 naive_fft is the spec.)
template <class T, class MulB, class Add, class Sub>
static inline std::vector<T> rs_encode_t(const std::vector<T> &coeffs, size_t full_log, T zero, MulB mulb, Add add, Sub sub) {
    size_t m = coeffs.size(), lg_m = ceil_log2(m), N = m << RS_RATE_LOG, lg_n = lg_m + RS_RATE_LOG;
    std::vector<T> a(N, zero);
    u64 shift = f_exp_pow2(GL_GENERATOR, full_log - lg_m), sp = 1;
    for (size_t i = 0; i < m; i++) { a[i] = mulb(coeffs[i], sp); sp = f_mul(sp, shift); }
    // radix-2 DIT NTT: bit-reverse, then butterflies with w_len = two_adic_generator(log len)
    reverse_index_bits_in_place(a);
    for (size_t lg = 1; lg <= lg_n; lg++) {
        size_t len = (size_t)1 << lg, half = len >> 1;
        u64 wlen = two_adic_generator(lg);
        std::vector<u64> tw(half); tw[0] = 1; for (size_t j = 1; j < half; j++) tw[j] = f_mul(tw[j - 1], wlen);
        par_for(N / 2, 8192, [&](size_t qb, size_t qe) {
            for (size_t q = qb; q < qe; q++) { size_t k = (q / half) * len, j = q % half; T t = mulb(a[k + half + j], tw[j]); T u = a[k + j]; a[k + j] = add(u, t); a[k + half + j] = sub(u, t); }
        });
    }
    return a;
}
static inline FVec rs_encode(const FVec &c, size_t full_log) {
    FVec o; o.is_ext = c.is_ext;
    if (c.is_ext) o.e = rs_encode_t<E>(c.e, full_log, E::zero(), e_mul_base, e_add, e_sub);
    else o.b = rs_encode_t<u64>(c.b, full_log, 0, f_mul, f_add, f_sub);
    return o;
}

// ---- Merkle (merkle_tree.rs:261-420, hash.rs:11-64) ----
struct MerkleTree {
    std::vector<std::vector<Digest>> inner;  // inner[0] = hashes of leaf pairs ... inner.back() = {root}
    FVec leaves;
    Digest root() const { return inner.back()[0]; }
    size_t height() const { return inner.size(); }
    // merkle_path_without_leaf_sibling_or_root (merkle_tree.rs:139-152)
    std::vector<Digest> path(size_t leaf_index) const {
        std::vector<Digest> p;
        for (size_t l = 0; l + 1 < inner.size(); l++) p.push_back(inner[l][(leaf_index >> (l + 1)) ^ 1]);
        return p;
    }
};
static inline std::vector<std::vector<Digest>> merkelize(const FVec &v) {
    size_t n = v.len(), lg = ceil_log2(n);
    std::vector<std::vector<Digest>> tree;
    std::vector<Digest> h(n >> 1);
    par_for(n >> 1, 8192, [&](size_t ib, size_t ie) { for (size_t i = ib; i < ie; i++) {
        if (v.is_ext) { u64 in[4] = {v.e[2 * i].c0, v.e[2 * i].c1, v.e[2 * i + 1].c0, v.e[2 * i + 1].c1}; h[i] = hash_or_noop(in, 4); }
        else { u64 in[2] = {v.b[2 * i], v.b[2 * i + 1]}; h[i] = hash_or_noop(in, 2); }
    } });
    tree.push_back(h);
    for (size_t l = 1; l < lg; l++) {
        const auto &prev = tree[l - 1];
        std::vector<Digest> nx(prev.size() >> 1);
        par_for(nx.size(), 64, [&](size_t ib, size_t ie) { for (size_t i = ib; i < ie; i++) nx[i] = compress(prev[2 * i], prev[2 * i + 1]); });
        tree.push_back(nx);
    }
    return tree;
}

// BasefoldCommitmentWithWitness (structure.rs:59-138), single polynomial
struct Commitment {
    MerkleTree codeword_tree;   // leaves = bit-reversed codeword (or raw evals when trivial)
    FVec bh_evals;              // bit-reversed evaluations (raw evals when trivial)
    size_t num_vars = 0; bool is_base = true; bool trivial = false;
    size_t codeword_size() const { return codeword_tree.leaves.len(); }
    Digest root() const { return codeword_tree.root(); }
};

// Basefold::commit (basefold.rs:304-354) via get_poly_bh_evals_and_codeword (:86-154)
static inline Commitment basefold_commit(const FVec &evals, size_t full_log) {
    Commitment c; c.num_vars = ceil_log2(evals.len()); c.is_base = !evals.is_ext;
    if (c.num_vars > full_log) throw std::runtime_error("PolynomialTooLarge");
    if (c.num_vars <= RS_BASECODE_MSG_SIZE_LOG) {  // TooSmall: Merkle over the raw evaluations
        c.trivial = true; c.bh_evals = evals; c.codeword_tree.leaves = evals; c.codeword_tree.inner = merkelize(evals);
        return c;
    }
    FVec coeffs = evals;
    if (coeffs.is_ext) { interpolate_hc<E>(coeffs.e, e_sub); reverse_index_bits_in_place(coeffs.e); }
    else { interpolate_hc<u64>(coeffs.b, f_sub); reverse_index_bits_in_place(coeffs.b); }
    FVec cw = rs_encode(coeffs, full_log);
    c.bh_evals = evals;
    if (evals.is_ext) { reverse_index_bits_in_place(c.bh_evals.e); reverse_index_bits_in_place(cw.e); }
    else { reverse_index_bits_in_place(c.bh_evals.b); reverse_index_bits_in_place(cw.b); }
    c.codeword_tree.leaves = cw; c.codeword_tree.inner = merkelize(cw);
    return c;
}

// RSCode::prover_folding_coeffs (rs.rs:377-410) == folding_coeffs_naive (:503-521): for the bit-reversed
// codeword, x0 = w_{2^(level+1)}^{rev(index, level)} * 7^(2^(full_log + rate - level - 1)), w = 1/(x1-x0) = -1/(2 x0)
static inline void folding_coeffs(size_t full_log, size_t level, size_t index, u64 &x0, u64 &w) {
    size_t idx = reverse_bits(index, level);
    x0 = f_mul(f_pow(two_adic_generator(level + 1), idx), f_exp_pow2(GL_GENERATOR, full_log + RS_RATE_LOG - level - 1));
    w = f_inv(f_sub(f_neg(x0), x0));
}
// basefold_one_round_by_interpolation_weights (commit_phase.rs:511-526) + interpolate2_weights (arithmetic.rs:120-132)
static inline std::vector<E> fri_fold(const std::vector<E> &v, size_t full_log, E r) {
    size_t level = ceil_log2(v.size()) - 1;
    std::vector<E> out(v.size() >> 1);
    // x0 for consecutive indices: precompute the root powers once (any method gives the same elements)
    u64 g = two_adic_generator(level + 1), shift = f_exp_pow2(GL_GENERATOR, full_log + RS_RATE_LOG - level - 1);
    std::vector<u64> pw((size_t)1 << level); pw[0] = 1; for (size_t i = 1; i < pw.size(); i++) pw[i] = f_mul(pw[i - 1], g);
    par_for(out.size(), 256, [&](size_t ib, size_t ie) { for (size_t i = ib; i < ie; i++) {
        u64 x0 = f_mul(pw[reverse_bits(i, level)], shift);
        u64 w = f_inv(f_sub(f_neg(x0), x0));
        E a1 = v[2 * i], b1 = v[2 * i + 1];
        out[i] = e_add(a1, e_mul(e_mul(e_sub(r, E::from_base(x0)), e_sub(b1, a1)), E::from_base(w)));
    } });
    return out;
}

// basefold/sumcheck.rs: coefficient-form messages over adjacent pairs of the bit-reversed arrays
static inline void one_level_interp_hc(std::vector<E> &v) { if (v.size() == 1) return; for (size_t i = 0; i + 1 < v.size(); i += 2) v[i + 1] = e_sub(v[i + 1], v[i]); }
static inline void one_level_eval_hc(std::vector<E> &v, E r) { std::vector<E> o(v.size() >> 1); for (size_t i = 0; i < o.size(); i++) o[i] = e_add(v[2 * i], e_mul(r, v[2 * i + 1])); v.swap(o); }
static inline std::vector<E> parallel_pi(const std::vector<E> &evals, const std::vector<E> &eq) {
    if (evals.size() == 1) return {evals[0], evals[0], evals[0]};
    E c0 = E::zero(), c1 = E::zero(), c2 = E::zero();
    for (size_t i = 0; i + 1 < evals.size(); i += 2) {
        c0 = e_add(c0, e_mul(evals[i], eq[i]));
        c1 = e_add(c1, e_add(e_mul(evals[i + 1], eq[i]), e_mul(evals[i], eq[i + 1])));
        c2 = e_add(c2, e_mul(evals[i + 1], eq[i + 1]));
    }
    return {c0, c1, c2};
}

struct CommitPhaseProof { std::vector<std::vector<E>> sumcheck_messages; std::vector<Digest> roots; std::vector<E> final_message; };

// PoseidonHasher: the four digest elements; BlakeHasher: append_message(32 digest bytes) (mpcs/src/util/hash.rs:57-62, 96-98) -- the trait's
// append_message turns them into four field elements too (each 8-byte word taken modulo p), so one line serves both
static inline void digest_to_transcript(const Digest &d, Transcript &t) { for (int i = 0; i < 4; i++) t.append_field_element(d.v[i]); }
static inline FVec ext_fvec(const std::vector<E> &v) { FVec f; f.is_ext = true; f.e = v; return f; }

// commit_phase (commit_phase.rs:30-183), single polynomial
static inline CommitPhaseProof commit_phase(size_t full_log, const std::vector<E> &point, const Commitment &comm, Transcript &t,
                                            size_t num_vars, size_t num_rounds, std::vector<MerkleTree> &trees) {
    std::vector<E> running_oracle(comm.codeword_size());
    for (size_t i = 0; i < running_oracle.size(); i++) running_oracle[i] = comm.codeword_tree.leaves.get(i);
    std::vector<E> running_evals(comm.bh_evals.len());
    for (size_t i = 0; i < running_evals.size(); i++) running_evals[i] = comm.bh_evals.get(i);
    std::vector<E> eq = build_eq_x_r_vec(point);
    reverse_index_bits_in_place(eq);
    // sum_check_first_round_field_type
    one_level_interp_hc(eq); one_level_interp_hc(running_evals);
    std::vector<E> last = parallel_pi(running_evals, eq);
    CommitPhaseProof pr;
    std::vector<std::vector<Digest>> running_tree_inner;
    for (size_t i = 0; i < num_rounds; i++) {
        t.append_field_element_exts(last);
        pr.sumcheck_messages.push_back(last);
        E ch = t.get_and_append_challenge("commit round");
        std::vector<E> new_oracle = fri_fold(running_oracle, full_log, ch);
        if (i > 0) { MerkleTree mt; mt.inner = running_tree_inner; mt.leaves = ext_fvec(running_oracle); trees.push_back(mt); }
        if (i + 1 < num_rounds) {
            // sum_check_challenge_round
            one_level_eval_hc(running_evals, ch); one_level_eval_hc(eq, ch);
            one_level_interp_hc(eq); one_level_interp_hc(running_evals);
            last = parallel_pi(running_evals, eq);
            running_tree_inner = merkelize(ext_fvec(new_oracle));
            Digest root = running_tree_inner.back()[0];
            digest_to_transcript(root, t);
            pr.roots.push_back(root);
            running_oracle = new_oracle;
        } else {
            one_level_eval_hc(running_evals, ch); one_level_eval_hc(eq, ch);  // sum_check_last_round
            reverse_index_bits_in_place(running_evals);
            t.append_field_element_exts(running_evals);
            pr.final_message = running_evals;
            // sanity-check of the reference (commit_phase.rs:148-169): final oracle == encode(final_message)
            std::vector<E> coeffs = pr.final_message;
            interpolate_hc<E>(coeffs, e_sub); reverse_index_bits_in_place(coeffs);
            FVec bc = rs_encode(ext_fvec(coeffs), full_log);
            std::vector<E> no = new_oracle; reverse_index_bits_in_place(no);
            if (!(bc.e == no)) throw std::runtime_error("basefold sanity: final oracle != encode(final_message)");
        }
    }
    (void)num_vars;
    return pr;
}

struct QueryOpening { E p0, p1; size_t index; std::vector<Digest> path; bool is_base = false; };
struct QueryResult { size_t x_index; QueryOpening commitment; std::vector<QueryOpening> oracle; };
struct BatchedQueryResult { size_t x_index; std::vector<QueryOpening> commitments; std::vector<QueryOpening> oracle; };

static inline QueryOpening open_pair(const MerkleTree &tree, size_t index) {
    QueryOpening q; size_t p1 = index | 1, p0 = p1 - 1;
    q.p0 = tree.leaves.get(p0); q.p1 = tree.leaves.get(p1); q.index = p0; q.is_base = !tree.leaves.is_ext; q.path = tree.path(p0);
    return q;
}
static inline std::vector<size_t> query_indices(Transcript &t, size_t n, size_t codeword_size) {
    std::vector<size_t> idx;
    for (size_t i = 0; i < n; i++) { E c = t.get_and_append_challenge("query indices"); idx.push_back((size_t)(c.c0 % codeword_size)); }
    return idx;
}

struct BasefoldProof {
    CommitPhaseProof commit_phase;
    std::vector<QueryResult> queries;                 // Single
    std::vector<BatchedQueryResult> batched_queries;  // Batched
    std::vector<std::vector<E>> sumcheck_proof;       // classic sumcheck rounds (batch_open only)
    bool trivial = false; FVec trivial_evals;
};

// Basefold::open (basefold.rs:466-539)
static inline BasefoldProof basefold_open(size_t full_log, const Commitment &comm, const std::vector<E> &point, Transcript &t) {
    BasefoldProof pr;
    if (comm.trivial) { pr.trivial = true; pr.trivial_evals = comm.bh_evals; return pr; }
    std::vector<MerkleTree> trees;
    pr.commit_phase = commit_phase(full_log, point, comm, t, comm.num_vars, comm.num_vars - RS_BASECODE_MSG_SIZE_LOG, trees);
    for (size_t x : query_indices(t, RS_NUM_QUERIES, comm.codeword_size())) {
        QueryResult q; q.x_index = x;
        q.commitment = open_pair(comm.codeword_tree, x);
        size_t index = x >> 1;
        for (auto &tr : trees) { q.oracle.push_back(open_pair(tr, index)); index >>= 1; }
        pr.queries.push_back(q);
    }
    return pr;
}

// ---- batch_open (basefold.rs:546-770) with the classic coefficient-form sumcheck ----
struct Evaluation { size_t poly, point; E value; };

static inline BasefoldProof basefold_batch_open(size_t full_log, const std::vector<FVec> &polys, const std::vector<const Commitment *> &comms,
                                                const std::vector<std::vector<E>> &points, const std::vector<Evaluation> &evals, Transcript &t) {
    BasefoldProof pr;
    if (polys.empty() && comms.empty() && points.empty() && evals.empty()) { pr.trivial = true; return pr; }
    size_t num_vars = 0, min_nv = 1000;
    std::vector<size_t> pnv;
    for (auto &p : polys) { size_t nv = ceil_log2(p.len()); pnv.push_back(nv); num_vars = std::max(num_vars, nv); min_nv = std::min(min_nv, nv); }
    if (min_nv <= RS_BASECODE_MSG_SIZE_LOG) throw std::runtime_error("minimum number of variables must be greater than basecode_msg_size_log");
    size_t bsl = 0; while (((size_t)1 << bsl) < evals.size()) bsl++;
    std::vector<E> tt; for (size_t i = 0; i < bsl; i++) tt.push_back(t.get_and_append_challenge("batch coeffs"));
    std::vector<E> eq_xt = build_eq_x_r_vec(tt);
    E target = E::zero();
    for (size_t i = 0; i < evals.size(); i++)
        target = e_add(target, e_mul(e_mul(evals[i].value, E::from_u64((u64)1 << (num_vars - points[evals[i].point].size()))), eq_xt[i]));
    // merged polynomial per point: sum_i eq_xt[i] * poly_i  (scalar kept aside while a point has one poly)
    struct Merged { E scalar; std::vector<E> poly; bool empty = true; };
    std::vector<Merged> merged(points.size());
    for (size_t i = 0; i < evals.size(); i++) {
        Merged &m = merged[evals[i].point]; const FVec &p = polys[evals[i].poly];
        if (m.empty) { m.scalar = eq_xt[i]; m.poly.resize(p.len()); for (size_t k = 0; k < p.len(); k++) m.poly[k] = p.get(k); m.empty = false; }
        else {
            if (m.scalar != E::one()) { for (auto &x : m.poly) x = e_mul(x, m.scalar); m.scalar = E::one(); }
            // add_polynomial_with_coeff (mpcs/src/util.rs:248-...): the smaller one is repeated
            if (p.len() > m.poly.size()) { std::vector<E> big(p.len()); for (size_t k = 0; k < big.size(); k++) big[k] = m.poly[k % m.poly.size()]; m.poly.swap(big); }
            for (size_t k = 0; k < m.poly.size(); k++) m.poly[k] = e_add(m.poly[k], e_mul(p.get(k % p.len()), eq_xt[i]));
        }
    }
    // ClassicSumCheck<CoefficientsProver>::prove (classic.rs:230-285, coeff.rs:196-345): LSB-first pairs
    std::vector<std::vector<E>> eqs, ps;
    for (size_t k = 0; k < points.size(); k++) { eqs.push_back(build_eq_x_r_vec(points[k])); ps.push_back(merged[k].poly); }
    E sum = target; std::vector<E> challenges;
    for (size_t round = 0; round < num_vars; round++) {
        size_t size = (size_t)1 << (num_vars - round - 1);
        E c0 = E::zero(), c2 = E::zero();
        for (size_t k = 0; k < points.size(); k++) {
            const auto &l = eqs[k]; const auto &r = ps[k]; size_t plen = l.size();
            E a0 = E::zero(), a2 = E::zero();
            if (plen == 1) a0 = e_mul(e_mul(l[0], r[0]), E::from_u64(size));
            else {
                size_t pairs = plen >> 1, mult = size / pairs;
                for (size_t i = 0; i < pairs; i++) { a0 = e_add(a0, e_mul(l[2 * i], r[2 * i])); a2 = e_add(a2, e_mul(e_sub(l[2 * i + 1], l[2 * i]), e_sub(r[2 * i + 1], r[2 * i]))); }
                if (mult != 1) { a0 = e_mul(a0, E::from_u64(mult)); a2 = e_mul(a2, E::from_u64(mult)); }
            }
            c0 = e_add(c0, e_mul(merged[k].scalar, a0)); c2 = e_add(c2, e_mul(merged[k].scalar, a2));
        }
        E c1 = e_sub(e_sub(sum, e_dbl(c0)), c2);
        std::vector<E> msg = {c0, c1, c2};
        t.append_field_element_exts(msg);
        pr.sumcheck_proof.push_back(msg);
        E ch = t.get_and_append_challenge("sumcheck round");
        challenges.push_back(ch);
        sum = e_add(c0, e_mul(ch, e_add(c1, e_mul(ch, c2))));  // horner
        for (size_t k = 0; k < points.size(); k++) {
            if (eqs[k].size() > 1) { std::vector<E> o(eqs[k].size() >> 1); for (size_t i = 0; i < o.size(); i++) o[i] = e_add(eqs[k][2 * i], e_mul(e_sub(eqs[k][2 * i + 1], eqs[k][2 * i]), ch)); eqs[k].swap(o); }
            if (ps[k].size() > 1) { std::vector<E> o(ps[k].size() >> 1); for (size_t i = 0; i < o.size(); i++) o[i] = e_add(ps[k][2 * i], e_mul(e_sub(ps[k][2 * i + 1], ps[k][2 * i]), ch)); ps[k].swap(o); }
        }
    }
    // coeffs[poly] = sum_i eq_xy(point_i, challenges[..len]) * eq_xt[i]   (basefold.rs:690-701)
    std::vector<E> coeffs(comms.size(), E::zero());
    for (size_t i = 0; i < evals.size(); i++) {
        const auto &pt = points[evals[i].point];
        std::vector<E> ch(challenges.begin(), challenges.begin() + pt.size());
        coeffs[evals[i].poly] = e_add(coeffs[evals[i].poly], e_mul(eq_eval(ch, pt), eq_xt[i]));
    }
    // batch_commit_phase (commit_phase.rs:187-358)
    const std::vector<E> &point = challenges;
    size_t num_rounds = num_vars - RS_BASECODE_MSG_SIZE_LOG;
    std::vector<E> running_oracle((size_t)1 << (num_vars + RS_RATE_LOG), E::zero());
    for (size_t c = 0; c < comms.size(); c++) if (comms[c]->codeword_size() == running_oracle.size())
        for (size_t i = 0; i < running_oracle.size(); i++) running_oracle[i] = e_add(running_oracle[i], e_mul(comms[c]->codeword_tree.leaves.get(i), coeffs[c]));
    std::vector<E> sum_evals((size_t)1 << num_vars, E::zero());
    for (size_t c = 0; c < comms.size(); c++) {
        size_t rep = sum_evals.size() / comms[c]->bh_evals.len();
        for (size_t i = 0; i < comms[c]->bh_evals.len(); i++) { E mul = e_mul(comms[c]->bh_evals.get(i), coeffs[c]); for (size_t k = 0; k < rep; k++) sum_evals[i * rep + k] = e_add(sum_evals[i * rep + k], mul); }
    }
    std::vector<E> eq = build_eq_x_r_vec(point); reverse_index_bits_in_place(eq);
    one_level_interp_hc(eq); one_level_interp_hc(sum_evals);
    std::vector<E> last = parallel_pi(sum_evals, eq);
    pr.commit_phase.sumcheck_messages.push_back(last);
    std::vector<MerkleTree> trees; std::vector<std::vector<Digest>> running_tree_inner; std::vector<E> new_oracle;
    for (size_t i = 0; i < num_rounds; i++) {
        t.append_field_element_exts(last);
        E ch = t.get_and_append_challenge("commit round");
        if (i > 0) {
            MerkleTree mt; mt.inner = running_tree_inner; mt.leaves = ext_fvec(new_oracle); trees.push_back(mt);
            for (size_t c = 0; c < comms.size(); c++) if (comms[c]->codeword_size() == new_oracle.size())
                for (size_t k = 0; k < new_oracle.size(); k++) new_oracle[k] = e_add(new_oracle[k], e_mul(comms[c]->codeword_tree.leaves.get(k), coeffs[c]));
            running_oracle = new_oracle;
        }
        new_oracle = fri_fold(running_oracle, full_log, ch);
        if (i + 1 < num_rounds) {
            one_level_eval_hc(sum_evals, ch); one_level_eval_hc(eq, ch); one_level_interp_hc(eq); one_level_interp_hc(sum_evals);
            last = parallel_pi(sum_evals, eq);
            pr.commit_phase.sumcheck_messages.push_back(last);
            running_tree_inner = merkelize(ext_fvec(new_oracle));
            Digest root = running_tree_inner.back()[0];
            digest_to_transcript(root, t); pr.commit_phase.roots.push_back(root);
        } else {
            one_level_eval_hc(sum_evals, ch); one_level_eval_hc(eq, ch);
            reverse_index_bits_in_place(sum_evals);
            t.append_field_element_exts(sum_evals);
            pr.commit_phase.final_message = sum_evals;
            std::vector<E> cf = pr.commit_phase.final_message;   // reference sanity-check :326-345
            reverse_index_bits_in_place(cf); interpolate_hc<E>(cf, e_sub);
            FVec bc = rs_encode(ext_fvec(cf), full_log);
            std::vector<E> no = new_oracle; reverse_index_bits_in_place(no);
            if (!(bc.e == no)) throw std::runtime_error("batch basefold sanity: final oracle != encode(final_message)");
        }
    }
    // batch_prover_query_phase (query_phase.rs:67-101,419-474)
    size_t codeword_size = (size_t)1 << (num_vars + RS_RATE_LOG);
    for (size_t x : query_indices(t, RS_NUM_QUERIES, codeword_size)) {
        BatchedQueryResult q; q.x_index = x;
        size_t index = x >> 1;
        for (auto &tr : trees) { q.oracle.push_back(open_pair(tr, index)); index >>= 1; }
        for (auto c : comms) { size_t xi = x >> (ceil_log2(codeword_size) - ceil_log2(c->codeword_size())); q.commitments.push_back(open_pair(c->codeword_tree, xi)); }
        pr.batched_queries.push_back(q);
    }
    return pr;
}

// Flat u64 serialisation used ONLY to compare the device path with the oracle (not the reference's rmp-serde)
static inline void flat_e(std::vector<u64> &o, E e) { o.push_back(e.c0); o.push_back(e.c1); }
static inline void flat_d(std::vector<u64> &o, const Digest &d) { for (int i = 0; i < 4; i++) o.push_back(d.v[i]); }
static inline void flat_q(std::vector<u64> &o, const QueryOpening &q) {
    o.push_back(q.index); o.push_back(q.is_base ? 1 : 0);
    if (q.is_base) { o.push_back(q.p0.c0); o.push_back(q.p1.c0); } else { flat_e(o, q.p0); flat_e(o, q.p1); }
    o.push_back(q.path.size()); for (auto &d : q.path) flat_d(o, d);
}
static inline std::vector<u64> flatten_proof(const BasefoldProof &p) {
    std::vector<u64> o;
    o.push_back(p.sumcheck_proof.size()); for (auto &m : p.sumcheck_proof) for (E e : m) flat_e(o, e);
    o.push_back(p.commit_phase.sumcheck_messages.size()); for (auto &m : p.commit_phase.sumcheck_messages) for (E e : m) flat_e(o, e);
    o.push_back(p.commit_phase.roots.size()); for (auto &d : p.commit_phase.roots) flat_d(o, d);
    o.push_back(p.commit_phase.final_message.size()); for (E e : p.commit_phase.final_message) flat_e(o, e);
    o.push_back(p.queries.size());
    for (auto &q : p.queries) { o.push_back(q.x_index); flat_q(o, q.commitment); o.push_back(q.oracle.size()); for (auto &x : q.oracle) flat_q(o, x); }
    o.push_back(p.batched_queries.size());
    for (auto &q : p.batched_queries) { o.push_back(q.x_index); o.push_back(q.oracle.size()); for (auto &x : q.oracle) flat_q(o, x); o.push_back(q.commitments.size()); for (auto &x : q.commitments) flat_q(o, x); }
    return o;
}


// ---- batch_commit / simple_batch_open: several same-size polynomials under ONE Merkle tree, opened at ONE point ----
// (Basefold::batch_commit basefold.rs:356-452, merkelize with batch leaves merkle_tree.rs:261-330 + hash_two_leaves_batch_*
//  util/hash.rs:30-41, simple_batch_open basefold.rs:777-861, simple_batch_commit_phase commit_phase.rs:363-510,
//  simple_batch_prover_query_phase query_phase.rs:104-138,474-534)
static inline Digest hash_batch_leaf(const std::vector<FVec> &vals, size_t idx) {
    std::vector<u64> in;
    for (auto &v : vals) { if (v.is_ext) { in.push_back(v.e[idx].c0); in.push_back(v.e[idx].c1); } else in.push_back(v.b[idx]); }
    return hash_or_noop(in.data(), in.size());                 // hash_bases / hash_elems: m-to-1 (no-op pad when <= 4 base elements)
}
static inline std::vector<std::vector<Digest>> merkelize_batch(const std::vector<FVec> &vals) {
    if (vals.size() == 1) return merkelize(vals[0]);
    size_t n = vals[0].len(), lg = ceil_log2(n);
    std::vector<std::vector<Digest>> tree;
    std::vector<Digest> h(n >> 1);
    par_for(n >> 1, 2048, [&](size_t ib, size_t ie) { for (size_t i = ib; i < ie; i++) h[i] = compress(hash_batch_leaf(vals, 2 * i), hash_batch_leaf(vals, 2 * i + 1)); });
    tree.push_back(h);
    for (size_t l = 1; l < lg; l++) {
        const auto &prev = tree[l - 1]; std::vector<Digest> nx(prev.size() >> 1);
        par_for(nx.size(), 64, [&](size_t ib, size_t ie) { for (size_t i = ib; i < ie; i++) nx[i] = compress(prev[2 * i], prev[2 * i + 1]); });
        tree.push_back(nx);
    }
    return tree;
}
struct BatchCommitment {                                        // BasefoldCommitmentWithWitness with num_polys > 1
    std::vector<std::vector<Digest>> inner; std::vector<FVec> leaves, bh_evals;   // leaves = bit-reversed codewords (raw evals when trivial)
    size_t num_vars = 0, num_polys = 0; bool is_base = true, trivial = false;
    Digest root() const { return inner.back()[0]; }
    size_t codeword_size() const { return leaves[0].len(); }
    std::vector<Digest> path(size_t leaf_index) const { std::vector<Digest> p; for (size_t l = 0; l + 1 < inner.size(); l++) p.push_back(inner[l][(leaf_index >> (l + 1)) ^ 1]); return p; }
};
static inline BatchCommitment basefold_batch_commit(const std::vector<FVec> &polys, size_t full_log) {
    if (polys.empty()) throw std::runtime_error("cannot batch commit to zero polynomials");
    BatchCommitment c; c.num_polys = polys.size(); c.num_vars = ceil_log2(polys[0].len()); c.is_base = !polys[0].is_ext;
    for (auto &p : polys) if (ceil_log2(p.len()) != c.num_vars || p.is_ext != polys[0].is_ext) throw std::runtime_error("cannot batch commit to polynomials with different number of variables");
    for (auto &p : polys) { Commitment one = basefold_commit(p, full_log); c.trivial = one.trivial; c.leaves.push_back(one.codeword_tree.leaves); c.bh_evals.push_back(one.bh_evals); }
    c.inner = merkelize_batch(c.leaves);
    return c;
}
struct SimpleBatchQueryResult { size_t x_index, index; bool is_base; std::vector<E> left, right; std::vector<Digest> path; std::vector<QueryOpening> oracle; };
struct SimpleBatchProof { CommitPhaseProof commit_phase; std::vector<SimpleBatchQueryResult> queries; bool trivial = false; std::vector<FVec> trivial_evals; };
// simple_batch_commit_phase: the single-polynomial commit phase on  sum_i coeff_i codeword_i  /  sum_i coeff_i bh_evals_i
static inline SimpleBatchProof basefold_simple_batch_open(size_t full_log, const BatchCommitment &comm, const std::vector<E> &point, const std::vector<E> &evals, Transcript &t) {
    SimpleBatchProof pr;
    if (comm.trivial) { pr.trivial = true; pr.trivial_evals = comm.bh_evals; return pr; }
    if (comm.num_polys != evals.size() || point.size() != comm.num_vars) throw std::runtime_error("simple_batch_open: shape mismatch");
    size_t bsl = ceil_log2(evals.size());
    std::vector<E> tt; for (size_t i = 0; i < bsl; i++) tt.push_back(t.get_and_append_challenge("batch coeffs"));
    std::vector<E> eq_xt = build_eq_x_r_vec(tt); eq_xt.resize(evals.size());
    // a synthetic single commitment carrying the batched codeword / evaluations drives the unchanged commit phase
    Commitment batched; batched.num_vars = comm.num_vars; batched.is_base = false;
    std::vector<E> cw(comm.codeword_size(), E::zero()), ev((size_t)1 << comm.num_vars, E::zero());
    for (size_t k = 0; k < comm.num_polys; k++) {
        for (size_t i = 0; i < cw.size(); i++) cw[i] = e_add(cw[i], e_mul(comm.leaves[k].get(i), eq_xt[k]));
        for (size_t i = 0; i < ev.size(); i++) ev[i] = e_add(ev[i], e_mul(comm.bh_evals[k].get(i), eq_xt[k]));
    }
    batched.codeword_tree.leaves = ext_fvec(cw); batched.bh_evals = ext_fvec(ev);
    std::vector<MerkleTree> trees;
    pr.commit_phase = commit_phase(full_log, point, batched, t, comm.num_vars, comm.num_vars - RS_BASECODE_MSG_SIZE_LOG, trees);
    for (size_t x : query_indices(t, RS_NUM_QUERIES, comm.codeword_size())) {
        SimpleBatchQueryResult q; q.x_index = x; size_t p1 = x | 1, p0 = p1 - 1; q.index = p0; q.is_base = comm.is_base;
        for (auto &l : comm.leaves) { q.left.push_back(l.get(p0)); q.right.push_back(l.get(p1)); }
        q.path = comm.path(p0);
        size_t index = x >> 1;
        for (auto &tr : trees) { q.oracle.push_back(open_pair(tr, index)); index >>= 1; }
        pr.queries.push_back(q);
    }
    return pr;
}
static inline std::vector<u64> flatten_simple_batch_proof(const SimpleBatchProof &p) {
    std::vector<u64> o;
    o.push_back(p.commit_phase.sumcheck_messages.size()); for (auto &m : p.commit_phase.sumcheck_messages) for (E e : m) flat_e(o, e);
    o.push_back(p.commit_phase.roots.size()); for (auto &d : p.commit_phase.roots) flat_d(o, d);
    o.push_back(p.commit_phase.final_message.size()); for (E e : p.commit_phase.final_message) flat_e(o, e);
    o.push_back(p.queries.size());
    for (auto &q : p.queries) {
        o.push_back(q.x_index); o.push_back(q.index); o.push_back(q.is_base ? 1 : 0); o.push_back(q.left.size());
        for (size_t k = 0; k < q.left.size(); k++) { if (q.is_base) { o.push_back(q.left[k].c0); o.push_back(q.right[k].c0); } else { flat_e(o, q.left[k]); flat_e(o, q.right[k]); } }
        o.push_back(q.path.size()); for (auto &d : q.path) flat_d(o, d);
        o.push_back(q.oracle.size()); for (auto &x : q.oracle) flat_q(o, x);
    }
    return o;
}

}  // namespace dpo
