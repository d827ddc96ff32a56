// ORACLE -- TEST INFRASTRUCTURE ONLY (see field.hpp header).
// Restates multilinear_extensions/src/{mle.rs,virtual_poly.rs} of the reference.
#pragma once
#include "field.hpp"
#include <memory>
#include <stdexcept>

namespace dpo {

// DenseMultilinearExtension{evaluations: FieldType<E>, num_vars}   (mle.rs:130-181)
// Little-endian index: bit 0 of the index is variable x_0 (mle.rs:226-229).
struct MLE {
    bool is_ext = false;
    size_t num_vars = 0;
    std::vector<u64> base;  // when !is_ext
    std::vector<E> ext;     // when is_ext
    size_t len() const { return is_ext ? ext.size() : base.size(); }
    static MLE from_base(size_t nv, std::vector<u64> v) { MLE m; m.is_ext = false; m.num_vars = nv; m.base = std::move(v); assert(m.base.size() == ((size_t)1 << nv)); return m; }
    static MLE from_ext(size_t nv, std::vector<E> v) { MLE m; m.is_ext = true; m.num_vars = nv; m.ext = std::move(v); assert(m.ext.size() == ((size_t)1 << nv)); return m; }
    E get(size_t i) const { return is_ext ? ext[i] : E::from_base(base[i]); }
};

// fix_variables / fix_variables_in_place(_parallel): LSB-first, adjacent pairs (mle.rs:454-525, 631-712)
//   g[i] = f[2i] + r * (f[2i+1] - f[2i]);  Base becomes Ext on the first fold.
static inline void mle_fix_low_one(MLE &m, E r) {
    assert(m.num_vars > 0);
    size_t half = m.len() >> 1;
    std::vector<E> out(half);
    if (m.is_ext) {
        par_for(half, 4096, [&](size_t b, size_t e) { for (size_t i = b; i < e; i++) out[i] = e_add(m.ext[2 * i], e_mul(e_sub(m.ext[2 * i + 1], m.ext[2 * i]), r)); });
    } else {
        par_for(half, 4096, [&](size_t b, size_t e) { for (size_t i = b; i < e; i++) out[i] = e_add(e_mul_base(r, f_sub(m.base[2 * i + 1], m.base[2 * i])), E::from_base(m.base[2 * i])); });
        m.base.clear(); m.base.shrink_to_fit();
    }
    m.ext.swap(out); m.is_ext = true; m.num_vars -= 1;
}
static inline MLE mle_fix_variables(const MLE &m, const std::vector<E> &point) {
    if (point.size() > m.num_vars) throw std::runtime_error("invalid size of partial point");
    MLE r = m;
    for (E p : point) mle_fix_low_one(r, p);
    return r;
}

// fix_high_variables(_in_place): MSB-first, lo/hi halves, point consumed in REVERSE (mle.rs:529-603)
//   for r in point.rev(): lo[i] += (hi[i] - lo[i]) * r
static inline void mle_fix_high_one(MLE &m, E r) {
    assert(m.num_vars > 0);
    size_t half = m.len() >> 1;
    std::vector<E> out(half);
    if (m.is_ext) {
        par_for(half, 4096, [&](size_t b, size_t e) { for (size_t i = b; i < e; i++) out[i] = e_add(m.ext[i], e_mul(e_sub(m.ext[i + half], m.ext[i]), r)); });
    } else {
        par_for(half, 4096, [&](size_t b, size_t e) { for (size_t i = b; i < e; i++) out[i] = e_add(e_mul_base(r, f_sub(m.base[i + half], m.base[i])), E::from_base(m.base[i])); });
        m.base.clear(); m.base.shrink_to_fit();
    }
    m.ext.swap(out); m.is_ext = true; m.num_vars -= 1;
}
static inline MLE mle_fix_high_variables(const MLE &m, const std::vector<E> &point) {
    if (point.size() > m.num_vars) throw std::runtime_error("invalid size of partial point");
    MLE r = m;
    for (size_t k = point.size(); k-- > 0;) mle_fix_high_one(r, point[k]);
    return r;
}

// evaluate (mle.rs:607-623): full LSB-first fold; a 0-variable Base MLE lifts its constant.
static inline E mle_evaluate(const MLE &m, const std::vector<E> &point) {
    if (point.size() != m.num_vars) throw std::runtime_error("MLE size does not match the point");
    MLE r = mle_fix_variables(m, point);
    return r.get(0);
}

// eq_eval (virtual_poly.rs:308-319)
static inline E eq_eval(const std::vector<E> &x, const std::vector<E> &y) {
    assert(x.size() == y.size());
    E res = E::one();
    for (size_t i = 0; i < x.size(); i++) {
        E xy = e_mul(x[i], y[i]);
        res = e_mul(res, e_add(e_sub(e_sub(e_add(xy, xy), x[i]), y[i]), E::one()));
    }
    return res;
}

// build_eq_x_r_vec(_sequential) (virtual_poly.rs:346-453): r processed from last to first,
//   buf[2j+1] = r*buf[j]; buf[2j] = buf[j] - r*buf[j].   Index bit i <-> r[i].
static inline std::vector<E> build_eq_x_r_vec(const std::vector<E> &r) {
    std::vector<E> buf((size_t)1 << r.size());
    buf[0] = E::one();
    size_t i = 0;
    for (size_t k = r.size(); k-- > 0; i++) {
        size_t next = (size_t)1 << (i + 1);
        for (size_t idx = next; idx >= 2; idx -= 2) {
            size_t index = idx - 2;
            E prev = buf[index >> 1];
            E tmp = e_mul(r[k], prev);
            buf[index + 1] = tmp;
            buf[index] = e_sub(prev, tmp);
        }
    }
    return buf;
}

// compute_betas_eval (zkml/src/commit/mod.rs:10-28) yields the same table in the same index
// order (checked in tests/test_oracle.py against the naive product formula).

}  // namespace dpo
