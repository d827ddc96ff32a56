// BLAKE3 compression on the device for the `blake` Merkle hasher (mpcs/src/util/hash.rs:79-95 BlakeHasher): every hash this path
// needs is a single chunk -- a leaf pair (16 or 32 bytes), two digests (64 bytes), or the values of all polynomials of a batch
// commitment at one index (<= 1024 bytes).  State in registers, message schedule resolved at compile time.
// The algorithm is the published BLAKE3 (host twin: include/dp_blake3.h, pinned against the Python `blake3` package).
#pragma once
#include "gl.cuh"

namespace b3 {
enum : u32 { CHUNK_START = 1, CHUNK_END = 2, PARENT = 4, ROOT = 8 };
__device__ __forceinline__ u32 rotr(u32 x, int n) { return __funnelshift_r(x, x, n); }
__device__ __forceinline__ void g(u32 &a, u32 &b, u32 &c, u32 &d, u32 mx, u32 my) {
    a = a + b + mx; d = rotr(d ^ a, 16); c = c + d; b = rotr(b ^ c, 12);
    a = a + b + my; d = rotr(d ^ a, 8);  c = c + d; b = rotr(b ^ c, 7);
}
// message word order of round r = the permutation (2 6 3 10 7 0 4 13 1 11 12 5 9 14 15 8) applied r times
__device__ __forceinline__ constexpr int sched(int r, int i) {
    constexpr int P[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
    int k = i; for (int t = 0; t < r; t++) k = P[k]; return k;
}
template <int R> __device__ __forceinline__ void round_fn(u32 (&s)[16], const u32 (&m)[16]) {
    g(s[0], s[4], s[8], s[12], m[sched(R, 0)], m[sched(R, 1)]);   g(s[1], s[5], s[9], s[13], m[sched(R, 2)], m[sched(R, 3)]);
    g(s[2], s[6], s[10], s[14], m[sched(R, 4)], m[sched(R, 5)]);  g(s[3], s[7], s[11], s[15], m[sched(R, 6)], m[sched(R, 7)]);
    g(s[0], s[5], s[10], s[15], m[sched(R, 8)], m[sched(R, 9)]);  g(s[1], s[6], s[11], s[12], m[sched(R, 10)], m[sched(R, 11)]);
    g(s[2], s[7], s[8], s[13], m[sched(R, 12)], m[sched(R, 13)]); g(s[3], s[4], s[9], s[14], m[sched(R, 14)], m[sched(R, 15)]);
}
// chaining value of one block (counter 0: every hash here is a single chunk)
__device__ __forceinline__ void compress(u32 (&cv)[8], const u32 (&m)[16], u32 block_len, u32 flags) {
    u32 s[16] = {cv[0], cv[1], cv[2], cv[3], cv[4], cv[5], cv[6], cv[7], 0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0u, 0u, block_len, flags};
    round_fn<0>(s, m); round_fn<1>(s, m); round_fn<2>(s, m); round_fn<3>(s, m); round_fn<4>(s, m); round_fn<5>(s, m); round_fn<6>(s, m);
#pragma unroll
    for (int i = 0; i < 8; i++) cv[i] = s[i] ^ s[i + 8];
}
__device__ __forceinline__ void iv(u32 (&cv)[8]) {
    cv[0] = 0x6A09E667u; cv[1] = 0xBB67AE85u; cv[2] = 0x3C6EF372u; cv[3] = 0xA54FF53Au; cv[4] = 0x510E527Fu; cv[5] = 0x9B05688Cu; cv[6] = 0x1F83D9ABu; cv[7] = 0x5BE0CD19u;
}
__device__ __forceinline__ void put(u32 (&m)[16], int k, u64 v) { m[2 * k] = (u32)v; m[2 * k + 1] = (u32)(v >> 32); }
__device__ __forceinline__ void digest_words(const u32 (&cv)[8], u64 out[4]) {
#pragma unroll
    for (int k = 0; k < 4; k++) out[k] = (u64)cv[2 * k] | ((u64)cv[2 * k + 1] << 32);
}
// BLAKE3 of nw u64 words (nw <= 8: one block), little-endian
__device__ __forceinline__ void hash_words(const u64 *w, int nw, u64 out[4]) {
    u32 m[16];
#pragma unroll
    for (int k = 0; k < 8; k++) put(m, k, k < nw ? w[k] : 0ULL);
    u32 cv[8]; iv(cv);
    compress(cv, m, 8u * nw, CHUNK_START | CHUNK_END | ROOT);
    digest_words(cv, out);
}
// BLAKE3 of n u64 words fetched through get(i), n <= 128 (one chunk of up to 16 blocks)
template <class F> __device__ __forceinline__ void hash_stream(F get, int n, u64 out[4]) {
    u32 cv[8]; iv(cv);
    const int nblocks = n == 0 ? 1 : (n + 7) / 8;
    for (int b = 0; b < nblocks; b++) {
        u32 m[16];
#pragma unroll
        for (int k = 0; k < 8; k++) put(m, k, 8 * b + k < n ? get(8 * b + k) : 0ULL);
        const int words = n - 8 * b < 8 ? n - 8 * b : 8;
        compress(cv, m, 8u * (u32)words, (b == 0 ? CHUNK_START : 0u) | (b == nblocks - 1 ? (CHUNK_END | ROOT) : 0u));
    }
    digest_words(cv, out);
}
}  // namespace b3
