// Host Fiat-Shamir: Poseidon2 width-8 permutation, duplex sponge, BasicTranscript.
// north_star keeps the transcript on the host; this is the host-language side of the boundary (the
// reference's `transcript` and `poseidon` crates), written in C++ because no Rust toolchain exists here.
// Reference: transcript/src/{lib.rs:22-93,basic.rs}, poseidon/src/challenger.rs:14-44,
// ff_ext/src/lib.rs:177-235.  Constants: include/dp_poseidon2_constants.h (provenance there).
#pragma once
#include "field.hpp"
#include "../../include/dp_poseidon2_constants.h"
#include <cstring>

namespace dp {

class Poseidon2 {
  public:
    static void permute(u64 s[8]) {
        linear_ext(s);
        for (int r = 0; r < 4; r++) { for (int i = 0; i < 8; i++) s[i] = pow7(fadd(s[i], DP_P2_EXT_RC[0][r][i])); linear_ext(s); }
        for (int r = 0; r < 22; r++) {
            s[0] = pow7(fadd(s[0], DP_P2_INT_RC[r]));
            u64 tot = 0;
            for (int i = 0; i < 8; i++) tot = fadd(tot, s[i]);
            for (int i = 0; i < 8; i++) s[i] = fadd(fmul(s[i], DP_P2_DIAG[i]), tot);
        }
        for (int r = 0; r < 4; r++) { for (int i = 0; i < 8; i++) s[i] = pow7(fadd(s[i], DP_P2_EXT_RC[1][r][i])); linear_ext(s); }
    }
  private:
    static u64 pow7(u64 x) { u64 a = fmul(x, x), b = fmul(a, a); return fmul(fmul(a, x), b); }
    // circ(2,3,1,1) on each half, then add the column sums (p3 MDSMat4 + mds_light_permutation)
    static void linear_ext(u64 *s) {
        for (int h = 0; h < 8; h += 4) {
            u64 *x = s + h;
            u64 a = fadd(x[0], x[1]), b = fadd(x[2], x[3]), all = fadd(a, b);
            u64 y0 = fadd(fadd(all, x[1]), a), y2 = fadd(fadd(all, x[3]), b);  // rows (2,3,1,1) and (1,1,2,3)
            u64 y1 = fadd(fadd(all, x[1]), fadd(x[2], x[2]));
            u64 y3 = fadd(fadd(all, x[3]), fadd(x[0], x[0]));
            x[0] = y0; x[1] = y1; x[2] = y2; x[3] = y3;
        }
        for (int k = 0; k < 4; k++) { u64 c = fadd(s[k], s[k + 4]); s[k] = fadd(s[k], c); s[k + 4] = fadd(s[k + 4], c); }
    }
};

// DuplexChallenger<F, Poseidon2, WIDTH = 8, RATE = 4>
class Sponge {
  public:
    Sponge() { memset(st_, 0, sizeof st_); }
    void observe(u64 v) {
        nout_ = 0;
        in_[nin_++] = v;
        if (nin_ == 4) duplex();
    }
    u64 sample() {
        if (nin_ || !nout_) duplex();
        return out_[--nout_];
    }
    u64 permutations() const { return nperm_; }
  private:
    void duplex() {
        for (int i = 0; i < nin_; i++) st_[i] = in_[i];
        nin_ = 0;
        Poseidon2::permute(st_); nperm_++;
        memcpy(out_, st_, 32); nout_ = 4;
    }
    u64 st_[8], in_[4], out_[4];
    int nin_ = 0, nout_ = 0;
    u64 nperm_ = 0;
};

// trait Transcript<E> as implemented by BasicTranscript<E>
class BasicTranscript {
  public:
    explicit BasicTranscript(const std::string &label) { append_message(label); }
    void append_field_element(u64 f) { sp_.observe(canon(f)); }
    void append_field_elements(const std::vector<u64> &f) { for (u64 x : f) append_field_element(x); }
    void append_message(const uint8_t *m, size_t n) {
        for (size_t i = 0; i < n; i += 8) { u64 v = 0; memcpy(&v, m + i, n - i < 8 ? n - i : 8); sp_.observe(v); }
    }
    void append_message(const std::string &s) { append_message((const uint8_t *)s.data(), s.size()); }
    void append_usize(u64 v) { sp_.observe(v); }  // usize::to_le_bytes() -> one field element
    void append_field_element_ext(const Ext &e) { sp_.observe(e.c0); sp_.observe(e.c1); }
    void append_field_element_exts(const ExtVec &v) { for (auto &e : v) append_field_element_ext(e); }
    Ext read_challenge() { u64 a = sp_.sample(); u64 b = sp_.sample(); return Ext(a, b); }
    Ext get_and_append_challenge(const std::string &label) { append_message(label); return read_challenge(); }
    u64 permutations() const { return sp_.permutations(); }
  private:
    Sponge sp_;
};

}  // namespace dp
