// ORACLE -- TEST INFRASTRUCTURE ONLY (see field.hpp header).
// CPU restatement of the zkml prover for the MLP path (Dense -> Requant -> ReLU chains):
//   zkml/src/iop/prover.rs:110-157,401-505 (Prover::prove, prove_tables), iop/context.rs:110-215,
//   layers/dense.rs:423-551, layers/requant.rs:208-330,531-680, layers/activation.rs:238-460,
//   lookup/context.rs:158-296,464-480,631-781, lookup/witness.rs, lookup/logup_gkr/{circuit,prover,structs}.rs,
//   commit/{mod.rs:10-28,context.rs:59-190,290-418,same_poly.rs:88-126}.
// The model is the synthetic quantised MLP of SURVEY.md 8(d) Cfg 2 (the reference does not pin "Dense-4M").
// The reference's debug asserts / sanity-check feature are kept as runtime checks (they throw).
#pragma once
#include "basefold.hpp"
#include <map>
#include <unordered_map>

namespace dpo {

typedef int64_t Element;                                   // zkml/src/lib.rs:40
static const size_t Q_BIT_LEN = 8;                          // quantization/mod.rs:20-26 (ZKML_BIT_LEN default)
static const Element Q_MIN = -127, Q_MAX = 127;             // quantization/mod.rs:28-29
static const Element COLUMN_SEPARATOR = (Element)1 << 32;   // lookup/context.rs:622

struct Claim { std::vector<E> point; E eval; };

struct RequantParams {                                      // layers/requant.rs:52-70
    size_t right_shift = 0, fp_scale = 0, intermediate_bit_size = 0; Element fixed_point_multiplier = 0;
    size_t shift() const { return fp_scale + right_shift; }
    size_t clamping_size() const { return intermediate_bit_size + ceil_log2((size_t)fixed_point_multiplier) - shift(); }   // :470-473
    Element apply(Element e) const {                         // :441-455
        Element rounding = (Element)1 << (shift() - 1);
        Element unclamped = (rounding + e * fixed_point_multiplier) >> shift();
        Element sign = unclamped >= 0 ? 1 : -1;
        Element a = unclamped < 0 ? -unclamped : unclamped;
        return a >= Q_MAX ? Q_MAX * sign : unclamped;
    }
};

enum OpKind { OP_DENSE = 0, OP_REQUANT = 1, OP_RELU = 2, OP_CONV = 3, OP_POOL = 4, OP_MATMUL = 5 };
struct Node {
    OpKind kind; size_t nrows = 0, ncols = 0;
    std::vector<Element> weights, bias;   // Dense, row-major nrows x ncols
    RequantParams rq;
    std::shared_ptr<struct ConvLayer> conv;                  // OP_CONV (padded layer, conv.hpp)
    size_t pool_c = 0, pool_h = 0, pool_w = 0;               // OP_POOL: padded input shape [C][H][W], Maxpool2D kernel = stride = 2
    // OP_MATMUL (layers/matrix_mul.rs, OperandMatrix::Input x OperandMatrix::Weight): the node's input is the LEFT matrix [mm_r][mm_k]
    // (row-major), `weights` the constant RIGHT matrix stored [mm_k][mm_c] -- or [mm_c][mm_k] with Config::TransposeB (mm_t) --, `bias`
    // (optional, mm_bias) has mm_c entries and is added to every output row (add_dim2); output [mm_r][mm_c]
    size_t mm_r = 0, mm_k = 0, mm_c = 0; bool mm_t = false, mm_bias = false;
};
struct Model { std::vector<Node> nodes; size_t input_len = 0; };

// TableType ordering = derive(Ord) on the enum (lookup/context.rs:53-63): Relu < Range < Clamping(size)
struct TableType { int kind; size_t size; bool operator<(const TableType &o) const { return kind != o.kind ? kind < o.kind : size < o.size; } };
static const int TT_RELU = 0, TT_RANGE = 2, TT_CLAMPING = 3;
static inline Element relu(Element e) { return e < 0 ? 0 : e; }

// get_merged_table_column (lookup/context.rs:158-296)
static inline void table_columns(const TableType &t, std::vector<Element> &merged, std::vector<std::vector<u64>> &cols) {
    merged.clear(); cols.clear();
    if (t.kind == TT_RELU) {
        cols.resize(2);
        for (Element i = Q_MIN - 1; i <= Q_MAX; i++) { Element o = relu(i); merged.push_back(i + o * COLUMN_SEPARATOR); cols[0].push_back(f_from_i64(i)); cols[1].push_back(f_from_i64(o)); }
    } else if (t.kind == TT_RANGE) {
        cols.resize(1);
        for (Element i = 0; i < ((Element)1 << Q_BIT_LEN); i++) { merged.push_back(i); cols[0].push_back(f_from_i64(i)); }
    } else {
        cols.resize(2);
        Element max = (Element)1 << (t.size - 1), min = -max;
        for (Element i = min; i < max; i++) { Element o = i < Q_MIN ? Q_MIN : (i > Q_MAX ? Q_MAX : i); merged.push_back(i + o * COLUMN_SEPARATOR); cols[0].push_back(f_from_i64(i)); cols[1].push_back(f_from_i64(o)); }
    }
}

// ---- LogUp GKR ----
struct LogUpLayer { bool has_num = true; std::vector<E> num, den; size_t num_vars() const { return ceil_log2(den.size() >> 1); } };
struct LogUpCircuit {
    std::vector<LogUpLayer> layers;
    std::vector<E> outputs() const { const LogUpLayer &l = layers.back(); std::vector<E> o; if (l.has_num) o = l.num; o.insert(o.end(), l.den.begin(), l.den.end()); return o; }
    size_t num_vars() const { return layers[0].num_vars(); }
};
static inline LogUpCircuit logup_circuit(const std::vector<const std::vector<u64> *> &cols, const std::vector<u64> *mult, E c, E gamma) {
    LogUpCircuit C; LogUpLayer l0; l0.has_num = mult != nullptr;
    size_t len = cols[0]->size();
    std::vector<E> pw; E p = E::one(); for (size_t k = 0; k < cols.size(); k++) { pw.push_back(p); p = e_mul(p, gamma); }
    l0.den.resize(len);
    for (size_t i = 0; i < len; i++) { E a = c; for (size_t k = 0; k < cols.size(); k++) a = e_add(a, e_mul_base(pw[k], (*cols[k])[i])); l0.den[i] = a; }
    if (mult) { l0.num.resize(len); for (size_t i = 0; i < len; i++) l0.num[i] = E::from_base((*mult)[i]); }
    C.layers.push_back(l0);
    while (C.layers.back().num_vars() != 0) {   // LogUpLayer::next_layer (circuit.rs:49-100)
        const LogUpLayer &l = C.layers.back(); size_t half = (size_t)1 << l.num_vars();
        LogUpLayer n; n.num.resize(half); n.den.resize(half);
        for (size_t i = 0; i < half; i++) {
            E n1 = l.has_num ? l.num[i] : e_neg(E::one()), n2 = l.has_num ? l.num[i + half] : e_neg(E::one());
            n.num[i] = e_add(e_mul(n1, l.den[i + half]), e_mul(l.den[i], n2));
            n.den[i] = e_mul(l.den[i], l.den[i + half]);
        }
        C.layers.push_back(n);
    }
    return C;
}

struct LogUpInput {
    bool table = false;
    std::vector<std::vector<u64>> column_evals; std::vector<u64> multiplicities;
    E constant_challenge, column_separation_challenge; size_t columns_per_instance = 1;
};
struct LogUpProof {
    std::vector<IOPProof> sumcheck_proofs; std::vector<std::vector<E>> round_evaluations; std::vector<Claim> output_claims;
    std::vector<std::vector<E>> circuit_outputs; bool table = false;
};
static std::shared_ptr<MLE> ext_mle(const E *p, size_t n) { auto m = std::make_shared<MLE>(); m->is_ext = true; m->num_vars = ceil_log2(n); m->ext.assign(p, p + n); return m; }
static std::shared_ptr<MLE> base_mle(const std::vector<u64> &v) { auto m = std::make_shared<MLE>(); m->is_ext = false; m->num_vars = ceil_log2(v.size()); m->base = v; return m; }

}  // namespace dpo
#include "conv.hpp"   // FFT-convolution layer (uses the definitions above; zk_prove below uses it)
namespace dpo {

// logup_gkr::prover::batch_prove (prover.rs:24-237)
static inline LogUpProof logup_batch_prove(const LogUpInput &in, Transcript &t) {
    std::vector<LogUpCircuit> circuits;
    if (in.table) {
        std::vector<const std::vector<u64> *> cs; for (auto &c : in.column_evals) cs.push_back(&c);
        circuits.push_back(logup_circuit(cs, &in.multiplicities, in.constant_challenge, in.column_separation_challenge));
    } else {
        for (size_t i = 0; i < in.column_evals.size(); i += in.columns_per_instance) {
            std::vector<const std::vector<u64> *> cs; for (size_t k = 0; k < in.columns_per_instance; k++) cs.push_back(&in.column_evals[i + k]);
            circuits.push_back(logup_circuit(cs, nullptr, in.constant_challenge, in.column_separation_challenge));
        }
    }
    LogUpProof pr; pr.table = in.table;
    size_t total_layers = 0;
    for (auto &c : circuits) { total_layers = std::max(total_layers, c.num_vars()); pr.circuit_outputs.push_back(c.outputs()); }
    t.append_field_element(f_from_u64(circuits.size()));
    for (auto &o : pr.circuit_outputs) t.append_field_element_exts(o);
    E batching = t.get_and_append_challenge("initial_batching"), alpha = t.get_and_append_challenge("initial_alpha"), lambda = t.get_and_append_challenge("initial_lambda");
    E claim = E::zero(), ac = E::one();
    for (auto &e : pr.circuit_outputs) {
        claim = e_add(claim, e_mul(ac, e_add(e_add(e_mul(batching, e_sub(e[1], e[0])), e[0]), e_mul(lambda, e_add(e_mul(batching, e_sub(e[3], e[2])), e[2])))));
        ac = e_mul(ac, alpha);
    }
    std::vector<E> point = {batching};
    for (size_t v = 1; v <= total_layers; v++) {
        t.append_field_element_ext(claim);
        auto eq = ext_mle(build_eq_x_r_vec(point).data(), (size_t)1 << v);   // compute_betas_eval
        VirtualPolynomial vp(v);
        E cur = E::one();
        for (auto &c : circuits) {
            size_t nl = c.layers.size();
            if (nl < v + 1) throw std::runtime_error("One of the circuits was not the same size as the others");
            const LogUpLayer &l = c.layers[nl - 1 - v];
            size_t half = (size_t)1 << v;
            if (l.has_num) {
                auto m0 = ext_mle(l.num.data(), half), m1 = ext_mle(l.num.data() + half, half), m2 = ext_mle(l.den.data(), half), m3 = ext_mle(l.den.data() + half, half);
                vp.add_mle_list({eq, m0, m3}, cur); vp.add_mle_list({eq, m1, m2}, cur); vp.add_mle_list({eq, m2, m3}, e_mul(cur, lambda));
            } else {
                auto m0 = ext_mle(l.den.data(), half), m1 = ext_mle(l.den.data() + half, half);
                vp.add_mle_list({eq, m1}, e_neg(cur)); vp.add_mle_list({eq, m0}, e_neg(cur)); vp.add_mle_list({eq, m0, m1}, e_mul(cur, lambda));
            }
            cur = e_mul(cur, alpha);
        }
        auto res = sumcheck_prove(vp, t);
        point = res.first.point;
        std::vector<E> evals(res.second.begin() + 1, res.second.end());
        batching = t.get_and_append_challenge("logup_batching"); alpha = t.get_and_append_challenge("logup_alpha"); lambda = t.get_and_append_challenge("logup_lambda");
        point.push_back(batching);
        pr.sumcheck_proofs.push_back(res.first);
        E acc = E::zero(), al = E::one();
        if (v != total_layers || in.table) {
            for (size_t i = 0; i + 3 < evals.size(); i += 4) {
                const E *e = &evals[i];
                acc = e_add(acc, e_mul(al, e_add(e_add(e_mul(batching, e_sub(e[2], e[0])), e[0]), e_mul(lambda, e_add(e_mul(batching, e_sub(e[1], e[3])), e[3])))));
                al = e_mul(al, alpha);
            }
        } else {
            for (size_t i = 0; i < evals.size(); i += 2) { const E *e = &evals[i]; acc = e_add(acc, e_mul(al, e_add(e_mul(batching, e_sub(e[0], e[1])), e[1]))); al = e_mul(al, alpha); }
        }
        claim = acc;
        pr.round_evaluations.push_back(evals);
    }
    // output claims on the base columns (prover.rs:172-183); the final claim must match them (logup_gkr/mod.rs:60-92)
    std::vector<const std::vector<u64> *> base;
    if (in.table) base.push_back(&in.multiplicities);
    for (auto &c : in.column_evals) base.push_back(&c);
    for (auto b : base) { Claim c; c.point = point; c.eval = mle_evaluate(*base_mle(*b), point); pr.output_claims.push_back(c); }
    return pr;
}

// same_poly::Prover::prove (commit/same_poly.rs:88-126)
struct SamePolyProof { IOPProof sumcheck; std::vector<E> evals; Claim extract_claim() const { return {sumcheck.point, evals[1]}; } };
static inline SamePolyProof same_poly_prove(const std::vector<E> &poly, const std::vector<Claim> &claims, Transcript &t) {
    std::vector<E> ch; for (size_t i = 0; i < claims.size(); i++) ch.push_back(t.read_challenge());   // read_challenges (lib.rs:145-150)
    std::vector<E> fb(poly.size(), E::zero());
    for (size_t k = 0; k < claims.size(); k++) { auto b = build_eq_x_r_vec(claims[k].point); for (size_t i = 0; i < fb.size(); i++) fb[i] = e_add(fb[i], e_mul(ch[k], b[i])); }
    VirtualPolynomial vp(ceil_log2(poly.size()));
    vp.add_mle_list({ext_mle(fb.data(), fb.size()), ext_mle(poly.data(), poly.size())}, E::one());
    auto res = sumcheck_prove(vp, t);
    return {res.first, res.second};
}

// ---- commitments ----
struct WitnessPoly { Commitment comm; std::vector<u64> evals; };   // (CommitmentWithWitness, DenseMultilinearExtension)
struct CommitmentClaim { std::shared_ptr<WitnessPoly> w; Claim claim; };

struct DenseProof { IOPProof sumcheck; E bias_eval; std::vector<E> individual_claims; };
struct RequantProof { IOPProof io_accumulation; std::vector<E> accumulation_evals; LogUpProof clamping_lookup, shifted_lookup; std::vector<Digest> commitments; };
struct ActivationProof { SamePolyProof io_accumulation; LogUpProof lookup; std::vector<Digest> commits; };
struct TableProof { Digest multiplicity_commit; LogUpProof lookup; };
struct PoolingProof { IOPProof sumcheck; LogUpProof lookup; std::vector<E> zerocheck_evals; size_t variable_gap = 0; std::vector<Digest> commitments; };   // layers/pooling.rs:60-75
struct ModelProof {
    std::map<size_t, DenseProof> dense; std::map<size_t, RequantProof> requant; std::map<size_t, ActivationProof> activation;
    std::vector<TableProof> table_proofs;
    BasefoldProof batch_proof; std::vector<BasefoldProof> trivial_proofs;
    std::map<size_t, ConvProof> conv; std::map<size_t, PoolingProof> pooling;
};

// Context::generate (iop/context.rs:110-215) for this model family: commits DenseBias / DenseWeight per dense node
struct ZkContext {
    const Model *model = nullptr; size_t full_log = 0;
    std::map<size_t, std::map<std::string, std::shared_ptr<WitnessPoly>>> model_comms;
    std::vector<TableType> tables;
};
static inline std::vector<u64> to_base_vec(const std::vector<Element> &v) { std::vector<u64> o(v.size()); for (size_t i = 0; i < v.size(); i++) o[i] = f_from_i64(v[i]); return o; }
static inline std::shared_ptr<WitnessPoly> commit_base(const std::vector<u64> &ev, size_t full_log) {
    auto w = std::make_shared<WitnessPoly>(); w->evals = ev; FVec f; f.is_ext = false; f.b = ev; w->comm = basefold_commit(f, full_log); return w;
}
static inline ZkContext zk_context(const Model &m) {
    ZkContext c; c.model = &m;
    size_t max_len = m.input_len; std::map<TableType, int> tabs;
    for (auto &n : m.nodes) {
        if (n.kind == OP_DENSE) max_len = std::max(max_len, std::max(n.nrows * n.ncols, n.nrows));
        if (n.kind == OP_MATMUL) max_len = std::max(max_len, std::max(n.mm_k * n.mm_c, n.mm_c));
        if (n.kind == OP_REQUANT) { tabs[{TT_RANGE, 0}] = 1; tabs[{TT_CLAMPING, n.rq.clamping_size()}] = 1; }
        if (n.kind == OP_RELU) tabs[{TT_RELU, 0}] = 1;
        if (n.kind == OP_CONV) max_len = std::max(max_len, std::max(n.conv->filter.size(), n.conv->kw * n.conv->nw * n.conv->nw));
        if (n.kind == OP_POOL) { tabs[{TT_RANGE, 0}] = 1; max_len = std::max(max_len, n.pool_c * n.pool_h * n.pool_w); }   // pooling.rs:130-170
    }
    for (auto &kv : tabs) { c.tables.push_back(kv.first); size_t mv = kv.first.kind == TT_CLAMPING ? kv.first.size : Q_BIT_LEN; max_len = std::max(max_len, (size_t)1 << mv); }   // context.rs:171-175
    c.full_log = ceil_log2(max_len);
    for (size_t id = 0; id < m.nodes.size(); id++) if (m.nodes[id].kind == OP_DENSE) {
        c.model_comms[id]["DenseBias"] = commit_base(to_base_vec(m.nodes[id].bias), c.full_log);
        c.model_comms[id]["DenseWeight"] = commit_base(to_base_vec(m.nodes[id].weights), c.full_log);
    } else if (m.nodes[id].kind == OP_MATMUL) {               // matrix_mul.rs:951-963: MatMulWeight (+ MatMulBias)
        if (m.nodes[id].mm_bias) c.model_comms[id]["MatMulBias"] = commit_base(to_base_vec(m.nodes[id].bias), c.full_log);
        c.model_comms[id]["MatMulWeight"] = commit_base(to_base_vec(m.nodes[id].weights), c.full_log);
    } else if (m.nodes[id].kind == OP_CONV) {                 // convolution.rs:532-540 (ConvBias < ConvFilter in the BTreeMap)
        c.model_comms[id]["ConvBias"] = commit_base(to_base_vec(m.nodes[id].conv->bias), c.full_log);
        c.model_comms[id]["ConvFilter"] = commit_base(to_base_vec(m.nodes[id].conv->filter), c.full_log);
    }
    return c;
}

// quantised inference trace (model/mod.rs run): outputs[i] = output of node i
// Maxpool2D::op (pooling.rs:667-677, tensor.rs:1335-1384) on a padded [C][H][W] tensor, kernel = stride = 2
static inline std::vector<Element> maxpool2d(const std::vector<Element> &x, size_t C, size_t H, size_t W) {
    std::vector<Element> o(C * (H / 2) * (W / 2));
    for (size_t c = 0; c < C; c++) for (size_t r = 0; r < H / 2; r++) for (size_t cc = 0; cc < W / 2; cc++) {
        Element mx = x[(c * H + 2 * r) * W + 2 * cc];
        for (size_t a = 0; a < 2; a++) for (size_t b = 0; b < 2; b++) mx = std::max(mx, x[(c * H + 2 * r + a) * W + 2 * cc + b]);
        o[(c * (H / 2) + r) * (W / 2) + cc] = mx;
    }
    return o;
}
// Maxpool2D::compute_polys (pooling.rs:686-771): out - in(2r+dr, 2c+dc) laid out like the output, in the order (dr,dc) = (0,0),(1,0),(0,1),(1,1)
static inline std::vector<std::vector<Element>> maxpool_diff_polys(const std::vector<Element> &x, const std::vector<Element> &out, size_t C, size_t H, size_t W) {
    std::vector<std::vector<Element>> d(4, std::vector<Element>(out.size()));
    static const size_t DR[4] = {0, 1, 0, 1}, DC[4] = {0, 0, 1, 1};
    for (size_t k = 0; k < 4; k++) for (size_t c = 0; c < C; c++) for (size_t r = 0; r < H / 2; r++) for (size_t cc = 0; cc < W / 2; cc++) {
        size_t oi = (c * (H / 2) + r) * (W / 2) + cc;
        d[k][oi] = out[oi] - x[(c * H + 2 * r + DR[k]) * W + 2 * cc + DC[k]];
    }
    return d;
}
static inline std::vector<std::vector<Element>> zk_run(const Model &m, const std::vector<Element> &input, std::map<size_t, ConvData> *conv_data = nullptr) {
    std::vector<std::vector<Element>> outs; std::vector<Element> cur = input;
    for (auto &n : m.nodes) {
        std::vector<Element> o;
        if (n.kind == OP_DENSE) { o.resize(n.nrows); for (size_t r = 0; r < n.nrows; r++) { Element a = n.bias[r]; for (size_t c = 0; c < n.ncols; c++) a += n.weights[r * n.ncols + c] * cur[c]; o[r] = a; } }
        else if (n.kind == OP_REQUANT) { for (Element e : cur) { Element lim = (Element)1 << n.rq.intermediate_bit_size; if (e > lim || e < -lim) throw std::runtime_error("Could not apply requantisation, tensor element had absolute value too large"); o.push_back(n.rq.apply(e)); } }
        else if (n.kind == OP_CONV) { ConvData cd; o = conv_op(*n.conv, cur, n.conv->nw, cd); if (conv_data) (*conv_data)[outs.size()] = std::move(cd); }
        else if (n.kind == OP_POOL) o = maxpool2d(cur, n.pool_c, n.pool_h, n.pool_w);
        else if (n.kind == OP_MATMUL) {   // MatMul::op (matrix_mul.rs:230-311): left x right (+ bias on every row)
            if (cur.size() != n.mm_r * n.mm_k) throw std::runtime_error("Incompatible shape found for input matrix");
            o.assign(n.mm_r * n.mm_c, 0);
            for (size_t r = 0; r < n.mm_r; r++) for (size_t c = 0; c < n.mm_c; c++) {
                Element a = n.mm_bias ? n.bias[c] : 0;
                for (size_t k = 0; k < n.mm_k; k++) a += cur[r * n.mm_k + k] * (n.mm_t ? n.weights[c * n.mm_k + k] : n.weights[k * n.mm_c + c]);
                o[r * n.mm_c + c] = a;
            }
        }
        else for (Element e : cur) o.push_back(relu(e));
        outs.push_back(o); cur = o;
    }
    return outs;
}

struct LogUpWitness { bool table = false; std::vector<std::shared_ptr<WitnessPoly>> commits; std::vector<std::vector<u64>> column_evals; std::vector<u64> multiplicity_evals; size_t columns_per_instance = 1; TableType tt; };

// Prover::prove (iop/prover.rs:401-486)
static inline ModelProof zk_prove(const ZkContext &ctx, const std::vector<Element> &input, Transcript &t) {
    const Model &m = *ctx.model; ModelProof proof;
    std::map<size_t, ConvData> conv_data;
    std::vector<std::vector<Element>> outs = zk_run(m, input, &conv_data);
    auto node_input = [&](size_t id) -> const std::vector<Element> & { return id == 0 ? input : outs[id - 1]; };
    // ctx.write_to_transcript (commit/context.rs:181-190)
    for (auto &nk : ctx.model_comms) for (auto &pk : nk.second) digest_to_transcript(pk.second->comm.root(), t);
    // generate_lookup_witnesses (lookup/context.rs:631-756)
    std::map<size_t, std::vector<LogUpWitness>> lookup_witness; std::map<TableType, std::unordered_map<Element, u64>> element_count;
    for (size_t id = 0; id < m.nodes.size(); id++) {
        const Node &n = m.nodes[id];
        if (n.kind == OP_REQUANT) {   // requant.rs:208-330
            size_t shift = n.rq.shift(); Element rc = (Element)1 << (shift - 1), mask = ((Element)1 << shift) - 1;
            std::vector<Element> cin, cout, shifted;
            for (Element v : node_input(id)) { Element tmp = v * n.rq.fixed_point_multiplier + rc; Element cl = tmp >> shift; cin.push_back(cl); cout.push_back(cl < Q_MIN ? Q_MIN : (cl > Q_MAX ? Q_MAX : cl)); shifted.push_back(tmp & mask); }
            size_t no_chunks = shift / Q_BIT_LEN; Element rmask = ((Element)1 << Q_BIT_LEN) - 1;
            std::vector<std::vector<Element>> chunks(no_chunks);
            for (size_t j = 0; j < no_chunks; j++) for (Element e : shifted) chunks[j].push_back((e >> (j * Q_BIT_LEN)) & rmask);
            TableType tc{TT_CLAMPING, n.rq.clamping_size()}, tr{TT_RANGE, 0};
            for (auto &ch : chunks) for (Element e : ch) element_count[tr][e]++;
            for (size_t i = 0; i < cin.size(); i++) element_count[tc][cin[i] + cout[i] * COLUMN_SEPARATOR]++;
            LogUpWitness wc; wc.tt = tc; wc.columns_per_instance = 2;
            for (auto *v : {&cin, &cout}) { auto ev = to_base_vec(*v); wc.commits.push_back(commit_base(ev, ctx.full_log)); wc.column_evals.push_back(ev); }
            LogUpWitness ws; ws.tt = tr; ws.columns_per_instance = 1;
            for (auto &ch : chunks) { auto ev = to_base_vec(ch); ws.commits.push_back(commit_base(ev, ctx.full_log)); ws.column_evals.push_back(ev); }
            lookup_witness[id] = {wc, ws};
        } else if (n.kind == OP_RELU) {   // activation.rs:238-323
            TableType tt{TT_RELU, 0}; LogUpWitness w; w.tt = tt; w.columns_per_instance = 2;
            const auto &a = node_input(id); const auto &b = outs[id];
            for (size_t i = 0; i < a.size(); i++) element_count[tt][a[i] + COLUMN_SEPARATOR * b[i]]++;
            for (auto *v : {&a, &b}) { auto ev = to_base_vec(*v); w.commits.push_back(commit_base(ev, ctx.full_log)); w.column_evals.push_back(ev); }
            lookup_witness[id] = {w};
        } else if (n.kind == OP_POOL) {   // pooling.rs:210-271: 4 difference columns looked up in Range; commits = columns + output
            TableType tr{TT_RANGE, 0}; LogUpWitness w; w.tt = tr; w.columns_per_instance = 1;
            auto diffs = maxpool_diff_polys(node_input(id), outs[id], n.pool_c, n.pool_h, n.pool_w);
            for (auto &d : diffs) { for (Element e : d) element_count[tr][e]++; auto ev = to_base_vec(d); w.commits.push_back(commit_base(ev, ctx.full_log)); w.column_evals.push_back(ev); }
            w.commits.push_back(commit_base(to_base_vec(outs[id]), ctx.full_log));
            lookup_witness[id] = {w};
        }
    }
    std::vector<LogUpWitness> table_witness;
    for (auto &kv : element_count) {   // BTreeMap order; multiplicities (lookup/context.rs:675-737)
        std::vector<Element> merged; std::vector<std::vector<u64>> cols; table_columns(kv.first, merged, cols);
        std::map<Element, u64> tcount; for (Element e : merged) tcount[e]++;
        std::vector<u64> mult(merged.size());
        for (size_t i = 0; i < merged.size(); i++) { auto it = kv.second.find(merged[i]); if (it == kv.second.end()) mult[i] = 0; else { u64 tc = tcount[merged[i]]; mult[i] = f_mul(f_from_u64(it->second), tc != 1 ? f_inv(f_from_u64(tc)) : 1); } }
        LogUpWitness w; w.table = true; w.tt = kv.first; w.multiplicity_evals = mult; w.column_evals = cols; w.commits.push_back(commit_base(mult, ctx.full_log));
        table_witness.push_back(w);
    }
    // initialise_from_table_set (lookup/context.rs:758-781)
    E constant_challenge = t.get_and_append_challenge("table_constant");
    std::map<TableType, E> challenge_map;
    for (auto &kv : element_count) challenge_map[kv.first] = kv.first.kind == TT_RELU ? t.get_and_append_challenge("Relu") : (kv.first.kind == TT_CLAMPING ? t.get_and_append_challenge("Clamping") : E::one());
    auto logup_input = [&](const LogUpWitness &w) { LogUpInput in; in.table = w.table; in.column_evals = w.column_evals; in.multiplicities = w.multiplicity_evals; in.constant_challenge = constant_challenge; in.column_separation_challenge = challenge_map.at(w.tt); in.columns_per_instance = w.columns_per_instance; return in; };
    // output claim (prover.rs:423-436)
    std::vector<CommitmentClaim> claims, trivial_claims;
    auto add_witness_claim = [&](std::shared_ptr<WitnessPoly> w, const Claim &c) { (w->comm.num_vars <= RS_BASECODE_MSG_SIZE_LOG ? trivial_claims : claims).push_back({w, c}); };
    const std::vector<Element> &final_out = outs.back();
    Claim last; { size_t nvo = ceil_log2(final_out.size()); for (size_t i = 0; i < nvo; i++) last.point.push_back(t.read_challenge()); last.eval = mle_evaluate(*base_mle(to_base_vec(final_out)), last.point); }
    for (size_t id = m.nodes.size(); id-- > 0;) {
        const Node &n = m.nodes[id];
        if (n.kind == OP_DENSE) {   // dense.rs:423-551
            E bias_eval = mle_evaluate(*base_mle(to_base_vec(n.bias)), last.point);
            MLE mat = mle_fix_high_variables(*base_mle(to_base_vec(n.weights)), last.point);
            auto mat_p = std::make_shared<MLE>(mat); auto in_p = base_mle(to_base_vec(node_input(id)));
            VirtualPolynomial vp(in_p->num_vars); vp.add_mle_list({mat_p, in_p}, E::one());
            auto res = sumcheck_prove(vp, t);
            if (res.first.extract_sum() != e_sub(last.eval, bias_eval)) throw std::runtime_error("dense: sumcheck output weird");   // dense.rs:493-503
            std::vector<E> wp = res.first.point; wp.insert(wp.end(), last.point.begin(), last.point.end());
            const auto &comms = ctx.model_comms.at(id);
            add_witness_claim(comms.at("DenseBias"), {last.point, bias_eval});        // BTreeMap order: DenseBias, DenseWeight
            add_witness_claim(comms.at("DenseWeight"), {wp, res.second[0]});
            proof.dense[id] = {res.first, bias_eval, res.second};
            last = {res.first.point, res.second[1]};
        } else if (n.kind == OP_MATMUL) {   // MatMul::prove_step (matrix_mul.rs:701-874), left = the node's input, right = the constant matrix
            const size_t vr = ceil_log2(n.mm_r), vc = ceil_log2(n.mm_c), vk = ceil_log2(n.mm_k);
            if (last.point.size() != vr + vc) throw std::runtime_error("Wrong length of last claim point");
            const std::vector<E> p_right(last.point.begin(), last.point.begin() + vc), p_left(last.point.begin() + vc, last.point.end());   // split_claim (:339-358)
            const auto &comms = ctx.model_comms.at(id);
            E claim_eval = last.eval, bias_eval = E::zero();
            if (n.mm_bias) { bias_eval = mle_evaluate(*base_mle(to_base_vec(n.bias)), p_right); claim_eval = e_sub(claim_eval, bias_eval); }
            MLE left = mle_fix_high_variables(*base_mle(to_base_vec(node_input(id))), p_left);            // rows of the left matrix are its HIGH variables
            MLE right = n.mm_t ? mle_fix_high_variables(*base_mle(to_base_vec(n.weights)), p_right)       // transposed: the output column is a ROW of the stored matrix
                               : mle_fix_variables(*base_mle(to_base_vec(n.weights)), p_right);           // else a column: the LOW variables
            if (left.num_vars != vk || right.num_vars != vk) throw std::runtime_error("matmul: free variables differ");
            auto lp = std::make_shared<MLE>(left); auto rp = std::make_shared<MLE>(right);
            VirtualPolynomial vp(vk); vp.add_mle_list({lp, rp}, E::one());
            auto res = sumcheck_prove(vp, t);
            if (res.first.extract_sum() != claim_eval) throw std::runtime_error("matmul: sumcheck output weird");
            // full_points (:364-383)
            std::vector<E> pl = res.first.point; pl.insert(pl.end(), p_left.begin(), p_left.end());
            std::vector<E> pr;
            if (n.mm_t) { pr = res.first.point; pr.insert(pr.end(), p_right.begin(), p_right.end()); } else { pr = p_right; pr.insert(pr.end(), res.first.point.begin(), res.first.point.end()); }
            if (n.mm_bias) add_witness_claim(comms.at("MatMulBias"), {p_right, bias_eval});               // add_common_claims walks the BTreeMap: MatMulBias < MatMulWeight
            add_witness_claim(comms.at("MatMulWeight"), {pr, res.second[1]});
            proof.dense[id] = {res.first, bias_eval, res.second};                                          // MatMulProof {sumcheck, individual_claims, bias_eval}: the same shape
            last = {pl, res.second[0]};
        } else if (n.kind == OP_REQUANT) {   // requant.rs:531-680
            auto ws = lookup_witness.at(id);
            LogUpInput cin = logup_input(ws[0]), sin = logup_input(ws[1]);
            LogUpProof cp = logup_batch_prove(cin, t), sp = logup_batch_prove(sin, t);
            size_t nv = ceil_log2(cin.column_evals[0].size());
            auto c0 = base_mle(cin.column_evals[0]), c1 = base_mle(cin.column_evals[1]);
            auto cbeta = ext_mle(build_eq_x_r_vec(cp.output_claims[0].point).data(), (size_t)1 << nv);
            auto lbeta = ext_mle(build_eq_x_r_vec(last.point).data(), (size_t)1 << nv);
            auto sbeta = ext_mle(build_eq_x_r_vec(sp.output_claims[0].point).data(), (size_t)1 << nv);
            E bc = t.get_and_append_challenge("requant_batching");
            VirtualPolynomial vp(nv);
            vp.add_mle_list({c1, lbeta}, E::one()); vp.add_mle_list({c1, cbeta}, bc);
            E comb = e_mul(bc, bc); vp.add_mle_list({c0, cbeta}, comb);
            comb = e_mul(comb, bc);
            for (auto &col : sin.column_evals) { vp.add_mle_list({sbeta, base_mle(col)}, comb); comb = e_mul(comb, bc); }
            auto res = sumcheck_prove(vp, t);
            const std::vector<E> &fe = res.second; std::vector<E> point = res.first.point;
            E cout_eval = fe[0], cin_eval = fe[3]; std::vector<E> sh(fe.begin() + 5, fe.end());
            // recombine_claims (requant.rs:483-515)
            E full = e_mul(E::from_u64((u64)1 << n.rq.shift()), cin_eval), p2 = E::one();
            for (E v : sh) { full = e_add(full, e_mul(v, p2)); p2 = e_mul(p2, E::from_u64((u64)1 << Q_BIT_LEN)); }
            E combined = e_mul(e_sub(full, E::from_u64((u64)1 << (n.rq.shift() - 1))), e_inv(E::from_base(f_from_i64(n.rq.fixed_point_multiplier))));
            RequantProof rp; rp.io_accumulation = res.first; rp.clamping_lookup = cp; rp.shifted_lookup = sp;
            std::vector<E> evs = {cin_eval, cout_eval}; evs.insert(evs.end(), sh.begin(), sh.end());
            std::vector<std::shared_ptr<WitnessPoly>> cw = ws[0].commits; cw.insert(cw.end(), ws[1].commits.begin(), ws[1].commits.end());
            for (size_t i = 0; i < evs.size(); i++) { add_witness_claim(cw[i], {point, evs[i]}); rp.accumulation_evals.push_back(evs[i]); rp.commitments.push_back(cw[i]->comm.root()); }
            proof.requant[id] = rp;
            last = {point, combined};
            // the claim handed on must be the requant INPUT evaluated at the point (verify_requant's recombination)
            if (mle_evaluate(*base_mle(to_base_vec(node_input(id))), point) != combined) throw std::runtime_error("requant: recombined claim mismatch");
        } else if (n.kind == OP_CONV) {   // convolution.rs:609-636 -> prove_convolution_step
            ConvProof cp; Claim in_claim = prove_convolution_step(*n.conv, t, last, conv_data.at(id), cp);
            const auto &comms = ctx.model_comms.at(id);                               // add_common_claims: BTreeMap order ConvBias, ConvFilter
            add_witness_claim(comms.at("ConvBias"), cp.bias_poly_claim);
            add_witness_claim(comms.at("ConvFilter"), cp.filter_claim);
            proof.conv[id] = cp;
            last = in_claim;
        } else if (n.kind == OP_POOL) {   // pooling.rs:342-519 prove_pooling
            auto ws = lookup_witness.at(id);
            LogUpInput in = logup_input(ws[0]);
            LogUpProof lp = logup_batch_prove(in, t);
            size_t nv = ceil_log2(outs[id].size());
            const std::vector<E> &lookup_point = lp.output_claims[0].point;
            E bc = t.get_and_append_challenge("batch_pooling");
            auto beta_poly = ext_mle(build_eq_x_r_vec(lookup_point).data(), (size_t)1 << nv);
            auto last_beta = ext_mle(build_eq_x_r_vec(last.point).data(), (size_t)1 << nv);
            std::vector<std::shared_ptr<MLE>> diffs; for (auto &col : in.column_evals) diffs.push_back(base_mle(col));
            VirtualPolynomial vp(nv);
            { auto all = diffs; all.push_back(beta_poly); vp.add_mle_list(all, E::one()); }
            E comb = bc; for (auto &d : diffs) { vp.add_mle_list({d, beta_poly}, comb); comb = e_mul(comb, bc); }
            auto out_mle = base_mle(to_base_vec(outs[id]));
            vp.add_mle_list({out_mle, last_beta}, comb);
            auto res = sumcheck_prove(vp, t);
            const std::vector<E> &fe = res.second; const std::vector<E> &zc_point = res.first.point;
            size_t ks = 4; E output_eval = fe[ks + 1];
            PoolingProof pp; pp.sumcheck = res.first; pp.lookup = lp;
            for (size_t i = 0; i <= ks; i++) { E ev = i < ks ? fe[i] : output_eval; add_witness_claim(ws[0].commits[i], {zc_point, ev}); pp.commitments.push_back(ws[0].commits[i]->comm.root()); pp.zerocheck_evals.push_back(ev); }
            size_t lw = ceil_log2(n.pool_w);
            E r1 = t.get_and_append_challenge("input_batching"), r2 = r1;             // `[challenge; 2]`: ONE challenge, copied (pooling.rs:453-456)
            E m1 = e_sub(E::one(), r1), m2 = e_sub(E::one(), r2);
            E mult[4] = {e_mul(m1, m2), e_mul(m1, r2), e_mul(r1, m2), e_mul(r1, r2)};
            E zc_in = E::zero(); for (size_t k = 0; k < ks; k++) zc_in = e_add(zc_in, e_mul(mult[k], e_sub(output_eval, fe[k])));
            Claim next; next.point.push_back(r1); next.point.insert(next.point.end(), zc_point.begin(), zc_point.begin() + (lw - 1));
            next.point.push_back(r2); next.point.insert(next.point.end(), zc_point.begin() + (lw - 1), zc_point.end());
            next.eval = zc_in; pp.variable_gap = lw - 1;
            if (mle_evaluate(*base_mle(to_base_vec(node_input(id))), next.point) != next.eval) throw std::runtime_error("pooling: input claim mismatch");
            proof.pooling[id] = pp;
            last = next;
        } else {   // activation.rs:385-460
            auto ws = lookup_witness.at(id);
            LogUpInput in = logup_input(ws[0]);
            LogUpProof lp = logup_batch_prove(in, t);
            std::vector<E> outp; for (Element e : outs[id]) outp.push_back(E::from_base(f_from_i64(e)));
            SamePolyProof acc = same_poly_prove(outp, {last, lp.output_claims[1]}, t);
            Claim input_claim = lp.output_claims[0];
            ActivationProof ap; ap.io_accumulation = acc; ap.lookup = lp;
            std::vector<Claim> cc = {input_claim, acc.extract_claim()};
            for (size_t i = 0; i < 2; i++) { add_witness_claim(ws[0].commits[i], cc[i]); ap.commits.push_back(ws[0].commits[i]->comm.root()); }
            proof.activation[id] = ap;
            last = input_claim;
        }
    }
    // the claim about the model input is checked by the verifier directly (iop/verifier.rs); here as a sanity check
    if (mle_evaluate(*base_mle(to_base_vec(input)), last.point) != last.eval) throw std::runtime_error("input claim mismatch");
    // prove_tables (prover.rs:110-157)
    for (auto &w : table_witness) {
        LogUpProof tp = logup_batch_prove(logup_input(w), t);
        add_witness_claim(w.commits[0], tp.output_claims[0]);
        proof.table_proofs.push_back({w.commits[0]->comm.root(), tp});
    }
    // LogUp soundness identity (lookup/logup_gkr/mod.rs, circuit.rs:298-322): sum of lookup fractions == sum of table fractions
    {
        E num = E::zero(), den = E::one();
        auto addf = [&](E n, E d) { num = e_add(e_mul(num, d), e_mul(den, n)); den = e_mul(den, d); };
        auto fold = [&](const LogUpProof &p, bool negate) { for (auto &o : p.circuit_outputs) { E n = e_add(e_mul(o[0], o[3]), e_mul(o[1], o[2])), d = e_mul(o[2], o[3]); addf(negate ? e_neg(n) : n, d); } };
        for (auto &kv : proof.requant) { fold(kv.second.clamping_lookup, false); fold(kv.second.shifted_lookup, false); }
        for (auto &kv : proof.activation) fold(kv.second.lookup, false);
        for (auto &kv : proof.pooling) fold(kv.second.lookup, false);
        for (auto &tp : proof.table_proofs) fold(tp.lookup, false);
        if (!num.is_zero()) throw std::runtime_error("logup: lookup and table fractional sums do not cancel");
    }
    // CommitmentProver::prove (commit/context.rs:355-418)
    for (auto &c : trivial_claims) { BasefoldProof p; p.trivial = true; p.trivial_evals = c.w->comm.bh_evals; proof.trivial_proofs.push_back(p); }
    std::vector<FVec> polys; std::vector<const Commitment *> comms; std::vector<std::vector<E>> points; std::vector<Evaluation> evals;
    for (size_t i = 0; i < claims.size(); i++) { FVec f; f.is_ext = false; f.b = claims[i].w->evals; polys.push_back(f); comms.push_back(&claims[i].w->comm); points.push_back(claims[i].claim.point); evals.push_back({i, i, claims[i].claim.eval}); }
    proof.batch_proof = basefold_batch_open(ctx.full_log, polys, comms, points, evals, t);
    return proof;
}

// ---- flat image for GPU-vs-oracle comparison ----
static inline void flat_iop(std::vector<u64> &o, const IOPProof &p) { o.push_back(p.point.size()); for (E e : p.point) flat_e(o, e); o.push_back(p.proofs.size()); for (auto &m : p.proofs) { o.push_back(m.size()); for (E e : m) flat_e(o, e); } }
static inline void flat_logup(std::vector<u64> &o, const LogUpProof &p) {
    o.push_back(p.sumcheck_proofs.size()); for (auto &s : p.sumcheck_proofs) flat_iop(o, s);
    o.push_back(p.round_evaluations.size()); for (auto &r : p.round_evaluations) { o.push_back(r.size()); for (E e : r) flat_e(o, e); }
    o.push_back(p.output_claims.size()); for (auto &c : p.output_claims) { o.push_back(c.point.size()); for (E e : c.point) flat_e(o, e); flat_e(o, c.eval); }
    o.push_back(p.circuit_outputs.size()); for (auto &r : p.circuit_outputs) { o.push_back(r.size()); for (E e : r) flat_e(o, e); }
    o.push_back(p.table ? 1 : 0);
}
static inline void flat_matrix_eval(std::vector<u64> &o, const MatrixEvalProof &m) {
    o.push_back(m.proofs.size()); for (auto &p : m.proofs) flat_iop(o, p);
    o.push_back(m.claims.size()); for (auto &c : m.claims) { o.push_back(c.size()); for (E e : c) flat_e(o, e); }
}
static inline void flat_evec(std::vector<u64> &o, const std::vector<E> &v) { o.push_back(v.size()); for (E e : v) flat_e(o, e); }
// field order of ConvProof (convolution.rs:97-121), then the two commitment claims and the returned input claim
static inline std::vector<u64> flatten_conv_proof(const ConvProof &p, const Claim &input_claim) {
    std::vector<u64> o;
    flat_iop(o, p.fft_proof); flat_evec(o, p.fft_claims); flat_iop(o, p.fft_proof_weights); flat_iop(o, p.ifft_proof);
    o.push_back(p.fft_delegation.proofs.size()); for (auto &q : p.fft_delegation.proofs) flat_iop(o, q);
    o.push_back(p.fft_delegation_weights.proofs.size()); for (auto &q : p.fft_delegation_weights.proofs) flat_iop(o, q);
    o.push_back(p.ifft_delegation.proofs.size()); for (auto &q : p.ifft_delegation.proofs) flat_iop(o, q);
    flat_iop(o, p.hadamard_proof); flat_evec(o, p.ifft_claims); flat_evec(o, p.fft_weight_claims);
    o.push_back(p.fft_delegation.claims.size()); for (auto &c : p.fft_delegation.claims) flat_evec(o, c);
    o.push_back(p.fft_delegation_weights.claims.size()); for (auto &c : p.fft_delegation_weights.claims) flat_evec(o, c);
    o.push_back(p.ifft_delegation.claims.size()); for (auto &c : p.ifft_delegation.claims) flat_evec(o, c);
    flat_evec(o, p.hadamard_claims); flat_e(o, p.bias_claim); flat_evec(o, p.partial_evals);
    flat_iop(o, p.clearing_proof.sumcheck); flat_evec(o, p.clearing_proof.individual_claim);
    flat_evec(o, p.filter_claim.point); flat_e(o, p.filter_claim.eval); flat_evec(o, p.bias_poly_claim.point); flat_e(o, p.bias_poly_claim.eval);
    flat_evec(o, input_claim.point); flat_e(o, input_claim.eval);
    return o;
}

static inline std::vector<u64> flatten_model_proof(const ModelProof &p, size_t n_nodes) {
    std::vector<u64> o;
    for (size_t id = 0; id < n_nodes; id++) {
        if (p.dense.count(id)) { const auto &d = p.dense.at(id); o.push_back(100 + id); flat_iop(o, d.sumcheck); flat_e(o, d.bias_eval); o.push_back(d.individual_claims.size()); for (E e : d.individual_claims) flat_e(o, e); }
        if (p.requant.count(id)) { const auto &r = p.requant.at(id); o.push_back(200 + id); flat_iop(o, r.io_accumulation); o.push_back(r.accumulation_evals.size()); for (E e : r.accumulation_evals) flat_e(o, e); flat_logup(o, r.clamping_lookup); flat_logup(o, r.shifted_lookup); o.push_back(r.commitments.size()); for (auto &d : r.commitments) flat_d(o, d); }
        if (p.activation.count(id)) { const auto &a = p.activation.at(id); o.push_back(300 + id); flat_iop(o, a.io_accumulation.sumcheck); o.push_back(a.io_accumulation.evals.size()); for (E e : a.io_accumulation.evals) flat_e(o, e); flat_logup(o, a.lookup); o.push_back(a.commits.size()); for (auto &d : a.commits) flat_d(o, d); }
        if (p.conv.count(id)) { o.push_back(400 + id); std::vector<u64> c = flatten_conv_proof(p.conv.at(id), Claim()); o.insert(o.end(), c.begin(), c.end()); }
        if (p.pooling.count(id)) { const auto &q = p.pooling.at(id); o.push_back(500 + id); flat_iop(o, q.sumcheck); flat_logup(o, q.lookup); flat_evec(o, q.zerocheck_evals); o.push_back(q.variable_gap); o.push_back(q.commitments.size()); for (auto &d : q.commitments) flat_d(o, d); }
    }
    o.push_back(p.table_proofs.size()); for (auto &t : p.table_proofs) { flat_d(o, t.multiplicity_commit); flat_logup(o, t.lookup); }
    o.push_back(p.trivial_proofs.size());
    std::vector<u64> b = flatten_proof(p.batch_proof); o.push_back(b.size()); o.insert(o.end(), b.begin(), b.end());
    return o;
}

// synthetic model of SURVEY.md 8(d) Cfg 2: n_layers x [Dense(width x width) + bias -> Requant -> ReLU]
static inline Model synthetic_mlp(size_t n_layers, size_t width, u64 seed) {
    Model m; m.input_len = width; SplitMix64 g(seed);
    for (size_t l = 0; l < n_layers; l++) {
        Node d; d.kind = OP_DENSE; d.nrows = d.ncols = width;
        d.weights.resize(width * width); for (auto &w : d.weights) w = (Element)(g.next() % 255) - 127;
        d.bias.resize(width); for (auto &b : d.bias) b = (Element)(g.next() % 255) - 127;
        m.nodes.push_back(d);
        Node r; r.kind = OP_REQUANT; size_t lw = ceil_log2(width);
        r.rq.intermediate_bit_size = 2 * (Q_BIT_LEN - 1) + lw + 1;                 // dense.rs:416-421 output_bitsize
        r.rq.right_shift = lw; r.rq.fp_scale = ((lw + 24 + 7) / 8) * 8 - lw;        // requant.rs:395-410 with int_part = log2(width)
        r.rq.fixed_point_multiplier = (Element)(3 * ((Element)1 << (r.rq.fp_scale - 2)));   // epsilon = 0.75
        m.nodes.push_back(r);
        Node a; a.kind = OP_RELU; m.nodes.push_back(a);
    }
    return m;
}
static inline RequantParams synthetic_requant(size_t int_part_log, size_t intermediate_bit_size) {
    RequantParams r; r.intermediate_bit_size = intermediate_bit_size; r.right_shift = int_part_log;
    r.fp_scale = ((int_part_log + 24 + 7) / 8) * 8 - int_part_log;
    r.fixed_point_multiplier = (Element)(3 * ((Element)1 << (r.fp_scale - 2)));
    return r;
}
// Synthetic CNN of SURVEY.md 8(d) Cfg 3, defined directly in the padded (power-of-two) domain the prover works in:
// cifar-cnn.py --num-params 264000 gives c1=12, c2=33, fc1=247, fc2=173 (zkml/assets/scripts/CNN/cifar-cnn.py:175-241):
//   in [3,32,32] -> conv 5x5 (12) -> requant -> relu -> maxpool -> conv 5x5 (33) -> requant -> relu -> maxpool -> flatten (33*5*5)
//   -> fc 247 -> requant -> relu -> fc 173 -> requant -> relu -> fc 10.   `scale` = 1 is that model; smaller test models
//   shrink the image (32 -> 16) and the channel counts.  Padded weights outside the real region are zero.
struct CnnShape { size_t img, c0, c1, c2, f1, f2, f3, k; };
static inline CnnShape cnn_shape(int small) { return small ? CnnShape{16, 3, 4, 6, 24, 16, 10, 3} : CnnShape{32, 3, 12, 33, 247, 173, 10, 5}; }
static inline size_t next_pow2(size_t x) { size_t p = 1; while (p < x) p <<= 1; return p; }
static inline Model synthetic_cnn(int small, u64 seed) {
    CnnShape s = cnn_shape(small); Model m; SplitMix64 g(seed);
    size_t kx_u = s.c0, n_u = s.img, n_p = next_pow2(s.img);
    m.input_len = next_pow2(s.c0) * n_p * n_p;
    size_t chans[2] = {s.c1, s.c2};
    for (int l = 0; l < 2; l++) {
        size_t kw_u = chans[l], kx = next_pow2(kx_u), kw = next_pow2(kw_u), rn = next_pow2(s.k);
        Node c; c.kind = OP_CONV; c.conv = std::make_shared<ConvLayer>(synthetic_conv(kw, kx, n_p, rn, kw_u, s.k, n_u, g.next()));
        for (size_t i = 0; i < kw; i++) for (size_t j = kx_u; j < kx; j++) for (size_t a = 0; a < rn * rn; a++) c.conv->filter[(i * kx + j) * rn * rn + a] = 0;   // padded input channels
        m.nodes.push_back(c);
        Node r; r.kind = OP_REQUANT; r.rq = synthetic_requant(ceil_log2(s.k * s.k * kx_u) + 2, 2 * (Q_BIT_LEN - 1) + ceil_log2(s.k * s.k * kx_u + 1)); m.nodes.push_back(r);   // convolution.rs:362-366 output_bitsize
        Node a; a.kind = OP_RELU; m.nodes.push_back(a);
        Node p; p.kind = OP_POOL; p.pool_c = kw; p.pool_h = p.pool_w = n_p; m.nodes.push_back(p);
        kx_u = kw_u; n_u = (n_u - s.k + 1) / 2; n_p /= 2;
    }
    size_t in_u_c = s.c2, in_p = next_pow2(s.c2) * n_p * n_p;       // flatten: [C][n_p][n_p] row-major, real region [c2][n_u][n_u]
    size_t outs_u[3] = {s.f1, s.f2, s.f3}; size_t ncols = in_p;
    for (int l = 0; l < 3; l++) {
        Node d; d.kind = OP_DENSE; d.nrows = next_pow2(outs_u[l]); d.ncols = ncols;
        d.weights.assign(d.nrows * d.ncols, 0); d.bias.assign(d.nrows, 0);
        for (size_t r = 0; r < outs_u[l]; r++) {
            for (size_t c = 0; c < d.ncols; c++) {
                bool real = l == 0 ? ((c / (n_p * n_p)) < in_u_c && ((c / n_p) % n_p) < n_u && (c % n_p) < n_u) : c < outs_u[l - 1];
                Element w = (Element)(g.next() % 255) - 127; if (real) d.weights[r * d.ncols + c] = w;
            }
            d.bias[r] = (Element)(g.next() % 255) - 127;
        }
        m.nodes.push_back(d);
        if (l < 2) {
            Node r; r.kind = OP_REQUANT; r.rq = synthetic_requant(ceil_log2(d.ncols), 2 * (Q_BIT_LEN - 1) + ceil_log2(d.ncols) + 1); m.nodes.push_back(r);
            Node a; a.kind = OP_RELU; m.nodes.push_back(a);
        }
        ncols = d.nrows;
    }
    return m;
}
static inline std::vector<Element> synthetic_cnn_input(int small, u64 seed) {
    CnnShape s = cnn_shape(small); size_t n_p = next_pow2(s.img);
    return synthetic_conv_input(next_pow2(s.c0), n_p, s.c0, s.img, seed);
}
static inline std::vector<Element> synthetic_input(size_t width, u64 seed) { SplitMix64 g(seed); std::vector<Element> v(width); for (auto &x : v) x = (Element)(g.next() % 128); return v; }

}  // namespace dpo
