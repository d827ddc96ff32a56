// ORACLE -- TEST INFRASTRUCTURE ONLY (see field.hpp header).
// Poseidon2 width-8 permutation, duplex challenger, Digest/hash/compress and BasicTranscript.
// Restates: ff_ext/src/lib.rs:167-254 (NoAllocPoseidon wiring), poseidon/src/{challenger.rs,
// poseidon_hash.rs,digest.rs}, transcript/src/{lib.rs:22-93,basic.rs}.  The permutation internals and
// DuplexChallenger live in Plonky3 (p3-poseidon2 / p3-challenger @ f37dc2a), which is NOT vendored in
// /root/reference; they are restated from the published Poseidon2 design + memory of upstream:
//   PARITY PARTIALLY PINNED -- see oracle/gen_poseidon2_constants.py docstring for exactly what is and
//   is not confirmed.  GPU-vs-oracle equality is exact; oracle-vs-reference equality of digests and
//   challenges is probable but unproven.
#pragma once
#include "field.hpp"
#include "../include/dp_poseidon2_constants.h"
#include "../include/dp_blake3.h"
#include <atomic>
#include <array>
#include <cstring>

namespace dpo {

// p3 MDSMat4 = circ(2,3,1,1) applied to 4 lanes (ff_ext/src/lib.rs:193 passes `&MDSMat4`)
static inline void p2_mat4(u64 *x) {
    u64 t01 = f_add(x[0], x[1]), t23 = f_add(x[2], x[3]);
    u64 t0123 = f_add(t01, t23);
    u64 t01123 = f_add(t0123, x[1]), t01233 = f_add(t0123, x[3]);
    u64 n3 = f_add(t01233, f_dbl(x[0]));  // 3x0 + x1 + x2 + 2x3
    u64 n1 = f_add(t01123, f_dbl(x[2]));  // x0 + 2x1 + 3x2 + x3
    u64 n0 = f_add(t01123, t01);          // 2x0 + 3x1 + x2 + x3
    u64 n2 = f_add(t01233, t23);          // x0 + x1 + 2x2 + 3x3
    x[0] = n0; x[1] = n1; x[2] = n2; x[3] = n3;
}
// mds_light_permutation, WIDTH = 8: M4 on each 4-chunk, then state[i] += sums[i % 4]
static inline void p2_mds_light(u64 *s) {
    p2_mat4(s); p2_mat4(s + 4);
    u64 sums[4];
    for (int k = 0; k < 4; k++) sums[k] = f_add(s[k], s[4 + k]);
    for (int i = 0; i < 8; i++) s[i] = f_add(s[i], sums[i & 3]);
}
static inline u64 p2_sbox(u64 x) { u64 x2 = f_mul(x, x), x3 = f_mul(x2, x), x4 = f_mul(x2, x2); return f_mul(x3, x4); }

// NoAllocPoseidon::permute_mut (ff_ext/src/lib.rs:222-228):
//   external_initial (mds_light, then 4x [add rc, x^7, mds_light]); 22x internal [s0 += rc; s0 ^= 7;
//   s[i] = s[i]*diag[i] + sum]; external_terminal (4x [add rc, x^7, mds_light]).
static inline void poseidon2_permute_plain(u64 s[8]) {
    p2_mds_light(s);
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 8; i++) s[i] = p2_sbox(f_add(s[i], DP_P2_EXT_RC[0][r][i]));
        p2_mds_light(s);
    }
    for (int r = 0; r < 22; r++) {
        s[0] = p2_sbox(f_add(s[0], DP_P2_INT_RC[r]));
        u64 sum = 0;
        for (int i = 0; i < 8; i++) sum = f_add(sum, s[i]);
        for (int i = 0; i < 8; i++) s[i] = f_add(f_mul(s[i], DP_P2_DIAG[i]), sum);
    }
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 8; i++) s[i] = p2_sbox(f_add(s[i], DP_P2_EXT_RC[1][r][i]));
        p2_mds_light(s);
    }
}

// Same permutation, written for speed (the Merkle trees of the CPU baseline are ~all permutations): state words stay "weak"
// (any u64), products are reduced 128 -> 64 bits without canonicalising, linear layers accumulate in 128 bits, the internal
// layer's s[i] * diag[i] + sum is one reduction.  Exact arithmetic: canonical in, canonical out, identical to the plain form
// above (dpo_poseidon2_selfcheck compares them; the golden vectors pin both).
static inline u64 p2w_red(u128 x) {
    u64 lo = (u64)x, hi = (u64)(x >> 64), hh = hi >> 32, hl = hi & 0xFFFFFFFFULL;
    u64 t0 = lo - hh; t0 -= ((u64)0 - (u64)(lo < hh)) & 0xFFFFFFFFULL;
    u64 t1 = (hl << 32) - hl;
    u64 r = t0 + t1; r += ((u64)0 - (u64)(r < t1)) & 0xFFFFFFFFULL;
    return r;
}
static inline u64 p2w_mul(u64 a, u64 b) { return p2w_red((u128)a * b); }
static inline u64 p2w_add(u64 a, u64 c) { u64 r = a + c; r += ((u64)0 - (u64)(r < a)) & 0xFFFFFFFFULL; return r; }
static inline u64 p2w_sbox(u64 x) { u64 a = p2w_mul(x, x), b = p2w_mul(a, a); return p2w_mul(p2w_mul(a, x), b); }
static inline void p2w_mds_light(u64 *s) {
    u128 n[8];
    for (int h = 0; h < 8; h += 4) {
        const u64 *x = s + h;
        u128 a = (u128)x[0] + x[1], b = (u128)x[2] + x[3], all = a + b;
        n[h + 0] = all + x[1] + a; n[h + 1] = all + x[1] + x[2] + x[2]; n[h + 2] = all + x[3] + b; n[h + 3] = all + x[3] + x[0] + x[0];
    }
    for (int i = 0; i < 8; i++) s[i] = p2w_red(n[i] + n[i] + n[i ^ 4]);
}
static inline void poseidon2_permute(u64 s[8]) {
    p2w_mds_light(s);
    for (int r = 0; r < 4; r++) { for (int i = 0; i < 8; i++) s[i] = p2w_sbox(p2w_add(s[i], DP_P2_EXT_RC[0][r][i])); p2w_mds_light(s); }
    for (int r = 0; r < 22; r++) {
        s[0] = p2w_sbox(p2w_add(s[0], DP_P2_INT_RC[r]));
        u128 tot = 0; for (int i = 0; i < 8; i++) tot += s[i];
        for (int i = 0; i < 8; i++) s[i] = p2w_red((u128)s[i] * DP_P2_DIAG[i] + tot);
    }
    for (int r = 0; r < 4; r++) { for (int i = 0; i < 8; i++) s[i] = p2w_sbox(p2w_add(s[i], DP_P2_EXT_RC[1][r][i])); p2w_mds_light(s); }
    for (int i = 0; i < 8; i++) s[i] = f_canon(s[i]);
}

// DuplexChallenger<F, P, 8, 4>  (poseidon/src/challenger.rs:14-20; p3-challenger duplex_challenger.rs)
struct Challenger {
    u64 state[8];
    u64 in_buf[4]; int n_in = 0;
    u64 out_buf[4]; int n_out = 0;
    u64 n_perm = 0;  // statistics only
    Challenger() { memset(state, 0, sizeof state); }
    void duplexing() {
        for (int i = 0; i < n_in; i++) state[i] = in_buf[i];
        n_in = 0;
        poseidon2_permute(state); n_perm++;
        for (int i = 0; i < 4; i++) out_buf[i] = state[i];
        n_out = 4;
    }
    void observe(u64 v) {
        n_out = 0;
        in_buf[n_in++] = v;
        if (n_in == 4) duplexing();
    }
    u64 sample() {
        if (n_in != 0 || n_out == 0) duplexing();
        return out_buf[--n_out];  // pops from the END of the rate
    }
};

struct Digest {
    u64 v[4];
    bool operator==(const Digest &o) const { return memcmp(v, o.v, sizeof v) == 0; }
};

// Which MerkleHasher / Transcript pair the checker runs (the reference picks at compile time: feature `blake`, mpcs/src/lib.rs:339-342,
// zkml/src/bin/bench.rs:29-44).  0 = PoseidonHasher + BasicTranscript (default), 1 = BlakeHasher + BlakeTranscript.  The blake pair is
// fully determined by the reference sources + standard BLAKE3, i.e. it is the PINNED variant of every digest, challenge and proof byte.
inline std::atomic<int> &hash_mode_var() { static std::atomic<int> m{0}; return m; }
static inline bool blake_mode() { return hash_mode_var().load(std::memory_order_relaxed) == 1; }
// BlakeHasher::hash_bases (mpcs/src/util/hash.rs:83-89): BLAKE3 over the canonical little-endian u64 bytes of the elements
static inline Digest blake_hash_bases(const u64 *in, size_t n) {
    dpb3::Hasher h;
    for (size_t i = 0; i < n; i++) { u64 c = f_canon(in[i]); uint8_t b[8]; for (int k = 0; k < 8; k++) b[k] = (uint8_t)(c >> (8 * k)); h.update(b, 8); }
    uint8_t o[32]; h.finalize(o, 32);
    Digest d; memcpy(d.v, o, 32);          // the 32 digest bytes, carried as four little-endian words
    return d;
}
// PoseidonHash::hash_or_noop (poseidon_hash.rs:22-28): <= 4 elements -> zero-padded copy, no permutation
static inline Digest hash_or_noop(const u64 *in, size_t n) {
    if (blake_mode()) return blake_hash_bases(in, n);
    Digest d;
    if (n <= 4) {
        for (size_t i = 0; i < 4; i++) d.v[i] = i < n ? in[i] : 0;
        return d;
    }
    Challenger c;
    for (size_t i = 0; i < n; i++) c.observe(in[i]);
    for (int i = 0; i < 4; i++) d.v[i] = c.sample();
    return d;
}
// compress (poseidon_hash.rs:66-71): observe x(4) -> permute, observe y(4) -> permute, sample 4
static inline Digest compress(const Digest &x, const Digest &y) {
    if (blake_mode()) {   // BlakeHasher::hash_two_digests (hash.rs:90-95): BLAKE3(a.bytes || b.bytes)
        uint8_t in[64], o[32]; memcpy(in, x.v, 32); memcpy(in + 32, y.v, 32); dpb3::hash(in, 64, o);
        Digest d; memcpy(d.v, o, 32); return d;
    }
    Challenger c;
    for (int i = 0; i < 4; i++) c.observe(x.v[i]);
    for (int i = 0; i < 4; i++) c.observe(y.v[i]);
    Digest d;
    for (int i = 0; i < 4; i++) d.v[i] = c.sample();
    return d;
}

// bytes_to_field_elements (ff_ext/src/lib.rs:262-273): 8-byte LE chunks, last one zero-padded
static inline std::vector<u64> bytes_to_field_elements(const uint8_t *b, size_t n) {
    std::vector<u64> out;
    for (size_t i = 0; i < n; i += 8) {
        uint8_t a[8] = {0};
        memcpy(a, b + i, n - i < 8 ? n - i : 8);
        u64 v; memcpy(&v, a, 8);
        out.push_back(v);  // from_canonical_u64: caller guarantees < p for labels/usize
    }
    return out;
}

// BasicTranscript (transcript/src/basic.rs) over trait Transcript (transcript/src/lib.rs:22-93); in blake mode BlakeTranscript
// (transcript/src/blake.rs): one running BLAKE3 hasher; a base element is absorbed as update("field_element") + update(LE bytes of the
// canonical value WITHOUT trailing zero bytes -- BigUint::to_bytes_le, [0] for zero); an extension element as
// update("field_element_ext") + update(bytes(c0) || bytes(c1)); a message goes through the trait default (lib.rs:42-45: 8-byte chunks
// -> field elements -> append_field_elements); a challenge is update("challenge") + 16 bytes of finalize_xof, parsed as two LE u64,
// retried until both are canonical (ff_ext/src/lib.rs:29-41, 246-254).
struct Transcript {
    Challenger ch;
    dpb3::Hasher bh; bool blake = false;
    Transcript() : blake(blake_mode()) {}
    explicit Transcript(const char *label) : blake(blake_mode()) {
        if (blake) bh.update(label, strlen(label));                       // BlakeTranscript::new(label): raw label bytes
        else append_message((const uint8_t *)label, strlen(label));
    }
    static size_t biguint_le(u64 c, uint8_t b[8]) { size_t n = 0; for (int k = 0; k < 8; k++) { b[k] = (uint8_t)(c >> (8 * k)); if (b[k]) n = k + 1; } return n ? n : 1; }
    void append_field_element(u64 f) {
        if (!blake) { ch.observe(f); return; }
        uint8_t b[8]; size_t n = biguint_le(f_canon(f), b);
        bh.update("field_element", 13); bh.update(b, n);
    }
    void append_field_elements(const u64 *f, size_t n) { for (size_t i = 0; i < n; i++) append_field_element(f[i]); }
    void append_message(const uint8_t *m, size_t n) { for (u64 f : bytes_to_field_elements(m, n)) append_field_element(f); }
    void append_usize(u64 v) { append_message((const uint8_t *)&v, 8); }  // usize.to_le_bytes()
    void append_field_element_ext(E e) {
        if (!blake) { ch.observe(e.c0); ch.observe(e.c1); return; }
        uint8_t b[16]; size_t n0 = biguint_le(f_canon(e.c0), b), n1 = biguint_le(f_canon(e.c1), b + n0);
        bh.update("field_element_ext", 17); bh.update(b, n0 + n1);
    }
    void append_field_element_exts(const std::vector<E> &v) { for (E e : v) append_field_element_ext(e); }
    E read_challenge() {
        if (!blake) { u64 a = ch.sample(); u64 b = ch.sample(); return E(a, b); }
        for (;;) {
            bh.update("challenge", 9);
            uint8_t o[16]; bh.finalize(o, 16);
            u64 a, b; memcpy(&a, o, 8); memcpy(&b, o + 8, 8);
            if (a < GL_P && b < GL_P) return E(a, b);
        }
    }
    E get_and_append_challenge(const char *label) { append_message((const uint8_t *)label, strlen(label)); return read_challenge(); }
    std::vector<E> sample_vec(size_t n) { std::vector<E> v; for (size_t i = 0; i < n; i++) v.push_back(read_challenge()); return v; }
};

}  // namespace dpo
