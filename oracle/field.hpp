// ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's field tower.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
// compile, link or call anything under oracle/.  The product (deep-prove_b200/) never does.
//
// F = Goldilocks, p = 2^64 - 2^32 + 1            (reference: ff_ext/src/lib.rs:7,285-310 -> p3-goldilocks)
// E = F[X]/(X^2 - 7) = BinomialExtensionField<F,2> (reference: ff_ext/src/lib.rs:13; W = 7 in p3-goldilocks)
// Values are kept CANONICAL (< p) at all times; E is AoS [c0, c1] like the reference's Vec<E>.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstddef>
#include <vector>
#include <cassert>
#include <thread>
#include <atomic>
#include <functional>
#include <algorithm>

namespace dpo {

typedef uint64_t u64;
typedef unsigned __int128 u128;
typedef uint32_t u32;

static const u64 GL_P = 0xFFFFFFFF00000001ULL;

// 2^64 = 2^32 - 1 and 2^96 = -1 (mod p): branch-free reduction of a 128-bit product (same value as x % p)
static inline u64 f_canon(u64 a) { u64 t = a + 0xFFFFFFFFULL; return t < a ? t : a; }
static inline u64 f_reduce128(u128 x) {
    u64 lo = (u64)x, hi = (u64)(x >> 64), hh = hi >> 32, hl = hi & 0xFFFFFFFFULL;
    u64 t0 = lo - hh; t0 -= ((u64)0 - (u64)(lo < hh)) & 0xFFFFFFFFULL;
    u64 t1 = (hl << 32) - hl, r = t0 + t1; r += ((u64)0 - (u64)(r < t1)) & 0xFFFFFFFFULL;
    return f_canon(r);
}
static inline u64 f_from_u64(u64 x) { return f_canon(x); }
static inline u64 f_add(u64 a, u64 b) { u64 s = a + b; s += ((u64)0 - (u64)(s < a)) & 0xFFFFFFFFULL; return f_canon(s); }
static inline u64 f_sub(u64 a, u64 b) { u64 d = a - b; return d - (((u64)0 - (u64)(a < b)) & 0xFFFFFFFFULL); }
static inline u64 f_neg(u64 a) { return a ? GL_P - a : 0; }
static inline u64 f_mul(u64 a, u64 b) { return f_reduce128((u128)a * b); }
static inline u64 f_dbl(u64 a) { return f_add(a, a); }
static inline u64 f_pow(u64 a, u64 e) {
    u64 r = 1;
    while (e) { if (e & 1) r = f_mul(r, a); a = f_mul(a, a); e >>= 1; }
    return r;
}
static inline u64 f_inv(u64 a) { assert(a != 0); return f_pow(a, GL_P - 2); }
// i64 -> F as the reference's quantised tensors do (negative values wrap to p - |v|)
static inline u64 f_from_i64(int64_t v) { return v >= 0 ? f_from_u64((u64)v) : f_neg(f_from_u64((u64)(-v))); }

struct E {
    u64 c0, c1;
    E() : c0(0), c1(0) {}
    E(u64 a, u64 b) : c0(a), c1(b) {}
    static E zero() { return E(0, 0); }
    static E one() { return E(1, 0); }
    static E from_base(u64 a) { return E(a, 0); }
    static E from_u64(u64 a) { return E(f_from_u64(a), 0); }
    bool operator==(const E &o) const { return c0 == o.c0 && c1 == o.c1; }
    bool operator!=(const E &o) const { return !(*this == o); }
    bool is_zero() const { return c0 == 0 && c1 == 0; }
};
static inline E e_add(E a, E b) { return E(f_add(a.c0, b.c0), f_add(a.c1, b.c1)); }
static inline E e_sub(E a, E b) { return E(f_sub(a.c0, b.c0), f_sub(a.c1, b.c1)); }
static inline E e_neg(E a) { return E(f_neg(a.c0), f_neg(a.c1)); }
static inline E e_dbl(E a) { return e_add(a, a); }
// (a0 + a1 X)(b0 + b1 X) = (a0 b0 + 7 a1 b1) + (a0 b1 + a1 b0) X
static inline E e_mul(E a, E b) {
    u64 t = f_mul(a.c1, b.c1);
    return E(f_add(f_mul(a.c0, b.c0), f_mul(t, 7)), f_add(f_mul(a.c0, b.c1), f_mul(a.c1, b.c0)));
}
static inline E e_mul_base(E a, u64 b) { return E(f_mul(a.c0, b), f_mul(a.c1, b)); }
// 1/(a0 + a1 X) = (a0 - a1 X) / (a0^2 - 7 a1^2)
static inline E e_inv(E a) {
    u64 n = f_sub(f_mul(a.c0, a.c0), f_mul(7, f_mul(a.c1, a.c1)));
    u64 ni = f_inv(n);
    return E(f_mul(a.c0, ni), f_mul(f_neg(a.c1), ni));
}
static inline E e_pow(E a, u64 e) {
    E r = E::one();
    while (e) { if (e & 1) r = e_mul(r, a); a = e_mul(a, a); e >>= 1; }
    return r;
}

// par_for: plain std::thread fork-join over [0, n) (no OpenMP in this image).  Threads = DPO_THREADS env or all hardware
// threads; dpo_set_threads() changes it at run time (bench.py's throughput mode: k concurrent proofs x T/k threads each).
inline std::atomic<unsigned> &dpo_threads_var() {
    static std::atomic<unsigned> n{[] { const char *e = getenv("DPO_THREADS"); unsigned v = e ? (unsigned)atoi(e) : std::thread::hardware_concurrency(); return v ? v : 1u; }()};
    return n;
}
static inline unsigned dpo_threads() { return dpo_threads_var().load(std::memory_order_relaxed); }
template <class F> static inline void par_for(size_t n, size_t min_per_thread, F f) {
    unsigned T = dpo_threads();
    if (T <= 1 || n < 2 * min_per_thread) { f((size_t)0, n); return; }
    size_t chunks = std::min<size_t>(T, n / min_per_thread); if (chunks < 2) { f((size_t)0, n); return; }
    std::vector<std::thread> th; size_t per = (n + chunks - 1) / chunks;
    for (size_t c = 1; c < chunks; c++) { size_t b = c * per, e = std::min(n, b + per); if (b < e) th.emplace_back([=] { f(b, e); }); }
    f((size_t)0, std::min(n, per));
    for (auto &t : th) t.join();
}

// splitmix64 -- the synthetic-input generator named in SURVEY.md 8(d) (seeds 1,2,3,...).
struct SplitMix64 {
    u64 s;
    explicit SplitMix64(u64 seed) : s(seed) {}
    u64 next() {
        u64 z = (s += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    u64 next_f() { return next() % GL_P; }
    E next_e() { u64 a = next_f(); u64 b = next_f(); return E(a, b); }
};

}  // namespace dpo
