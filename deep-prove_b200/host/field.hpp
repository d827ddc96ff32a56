// Host-side Goldilocks / GoldilocksExt2 arithmetic for the orchestration layer (transcript, claims,
// O(log n)-sized glue).  The O(n) loops are NOT here: they go through include/deepprove_b200.h.
// Reference: ff_ext/src/lib.rs:7,13 (p3 Goldilocks, BinomialExtensionField<_,2>, W = 7).
#pragma once
#include <cstdint>
#include <vector>
#include <string>
#include <stdexcept>

namespace dp {

typedef uint64_t u64;
typedef unsigned __int128 u128;
constexpr u64 P = 0xFFFFFFFF00000001ULL;
constexpr u64 EPS = 0xFFFFFFFFULL;

// branch-free on purpose: the Fiat-Shamir sponge sits on the critical path between device rounds and its
// operands are random, so data-dependent branches mispredict constantly (6 us -> <1 us per permutation)
inline u64 canon(u64 a) { u64 t = a + EPS; return t < a ? t : a; }            // a >= p  <=>  a + (2^32-1) wraps
inline u64 fadd(u64 a, u64 b) { u64 s = a + b; u64 c = (u64)0 - (u64)(s < a); s += c & EPS; return canon(s); }   // a,b < p
inline u64 fsub(u64 a, u64 b) { u64 d = a - b; u64 br = (u64)0 - (u64)(a < b); return d - (br & EPS); }           // + p == - EPS (mod 2^64)
inline u64 fneg(u64 a) { return fsub(0, a); }
inline u64 fmul(u64 a, u64 b) {
    u128 t = (u128)a * b;
    u64 lo = (u64)t, hi = (u64)(t >> 64), hh = hi >> 32, hl = hi & EPS;
    u64 t0 = lo - hh; t0 -= ((u64)0 - (u64)(lo < hh)) & EPS;
    u64 t1 = (hl << 32) - hl;                                                   // hl * (2^32 - 1)
    u64 r = t0 + t1; r += ((u64)0 - (u64)(r < t1)) & EPS;
    return canon(r);
}
inline u64 fpow(u64 a, u64 e) { u64 r = 1; while (e) { if (e & 1) r = fmul(r, a); a = fmul(a, a); e >>= 1; } return r; }
inline u64 finv(u64 a) { if (!a) throw std::runtime_error("inverse of zero"); return fpow(a, P - 2); }
inline u64 from_i64(int64_t v) { return v >= 0 ? canon((u64)v) : fneg(canon((u64)(-v))); }

struct Ext {
    u64 c0 = 0, c1 = 0;
    Ext() {}
    Ext(u64 a, u64 b) : c0(a), c1(b) {}
    static Ext one() { return Ext(1, 0); }
    static Ext zero() { return Ext(0, 0); }
    static Ext from_base(u64 a) { return Ext(a, 0); }
    bool operator==(const Ext &o) const { return c0 == o.c0 && c1 == o.c1; }
    bool operator!=(const Ext &o) const { return !(*this == o); }
    Ext operator+(const Ext &o) const { return Ext(fadd(c0, o.c0), fadd(c1, o.c1)); }
    Ext operator-(const Ext &o) const { return Ext(fsub(c0, o.c0), fsub(c1, o.c1)); }
    Ext operator-() const { return Ext(fneg(c0), fneg(c1)); }
    Ext operator*(const Ext &o) const {
        return Ext(fadd(fmul(c0, o.c0), fmul(7, fmul(c1, o.c1))), fadd(fmul(c0, o.c1), fmul(c1, o.c0)));
    }
    Ext operator*(u64 b) const { return Ext(fmul(c0, b), fmul(c1, b)); }
    Ext &operator+=(const Ext &o) { return *this = *this + o; }
    Ext &operator-=(const Ext &o) { return *this = *this - o; }
    Ext &operator*=(const Ext &o) { return *this = *this * o; }
    Ext inverse() const {
        u64 n = fsub(fmul(c0, c0), fmul(7, fmul(c1, c1)));
        u64 ni = finv(n);
        return Ext(fmul(c0, ni), fmul(fneg(c1), ni));
    }
    bool is_zero() const { return !c0 && !c1; }
};
typedef std::vector<Ext> ExtVec;

inline std::vector<u64> flatten(const ExtVec &v) { std::vector<u64> o; o.reserve(2 * v.size()); for (auto &e : v) { o.push_back(e.c0); o.push_back(e.c1); } return o; }

}  // namespace dp
