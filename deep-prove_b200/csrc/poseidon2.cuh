// Poseidon2 (Goldilocks, width 8, x^7, 4+22+4 rounds) with the state held in registers, and the two
// Merkle hash shapes built on it.  Reference: ff_ext/src/lib.rs:177-235 (NoAllocPoseidon),
// poseidon/src/poseidon_hash.rs:17-71 over p3 DuplexChallenger<_,_,8,4> (overwrite-mode absorb,
// outputs popped from the end of the rate).  Constants: include/dp_poseidon2_constants.h (provenance and
// pinning status are documented in the generator script that emits that header).
#pragma once
#include "gl.cuh"
#include "../../include/dp_poseidon2_constants.h"

__constant__ u64 c_p2_ext[2][4][8];
__constant__ u64 c_p2_int[22];
__constant__ u64 c_p2_diag[8];

static inline cudaError_t p2_upload_constants(const u64 *ext /*64*/, const u64 *internal /*22*/, const u64 *diag /*8*/) {
    cudaError_t e;
    if ((e = cudaMemcpyToSymbol(c_p2_ext, ext, sizeof(u64) * 64)) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_p2_int, internal, sizeof(u64) * 22)) != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_p2_diag, diag, sizeof(u64) * 8);
}

__device__ __forceinline__ u64 p2_pow7(u64 x) { u64 x2 = gl_sqr(x), x4 = gl_sqr(x2); return gl_mul(gl_mul(x2, x), x4); }
// p3 MDSMat4 = circ(2,3,1,1) on four lanes
__device__ __forceinline__ void p2_mat4(u64 &x0, u64 &x1, u64 &x2, u64 &x3) {
    u64 t01 = gl_add(x0, x1), t23 = gl_add(x2, x3), t = gl_add(t01, t23);
    u64 a = gl_add(t, x1), b = gl_add(t, x3);
    u64 n3 = gl_add(b, gl_dbl(x0)), n1 = gl_add(a, gl_dbl(x2));
    u64 n0 = gl_add(a, t01), n2 = gl_add(b, t23);
    x0 = n0; x1 = n1; x2 = n2; x3 = n3;
}
__device__ __forceinline__ void p2_mds_light(u64 (&s)[8]) {
    p2_mat4(s[0], s[1], s[2], s[3]); p2_mat4(s[4], s[5], s[6], s[7]);
#pragma unroll
    for (int k = 0; k < 4; k++) { u64 c = gl_add(s[k], s[k + 4]); s[k] = gl_add(s[k], c); s[k + 4] = gl_add(s[k + 4], c); }
}
__device__ __forceinline__ void p2_permute(u64 (&s)[8]) {
    p2_mds_light(s);
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = p2_pow7(gl_add(s[i], c_p2_ext[0][r][i]));
        p2_mds_light(s);
    }
#pragma unroll 1
    for (int r = 0; r < 22; r++) {
        s[0] = p2_pow7(gl_add(s[0], c_p2_int[r]));
        u64 sum = gl_add(gl_add(gl_add(s[0], s[1]), gl_add(s[2], s[3])), gl_add(gl_add(s[4], s[5]), gl_add(s[6], s[7])));
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = gl_add(gl_mul(s[i], c_p2_diag[i]), sum);
    }
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = p2_pow7(gl_add(s[i], c_p2_ext[1][r][i]));
        p2_mds_light(s);
    }
}
// ---- lane-parallel variant: 8 consecutive lanes hold the 8 state words of ONE permutation ----------------
// Used where a level has too few hashes to fill the GPU with one thread per hash (tree tops, witness-sized
// trees): the dependent-instruction chain per permutation drops ~3x (every lane does one S-box in the full
// rounds; the linear layers are warp shuffles inside the 8-lane group).
__device__ __forceinline__ u64 shfl8(u64 v, int src) { return __shfl_sync(0xffffffffu, v, src, 8); }
__device__ __forceinline__ u64 shfl8_xor(u64 v, int m) { return __shfl_xor_sync(0xffffffffu, v, m, 8); }
__device__ __forceinline__ u64 p2x8_mds_light(u64 s, int lane8) {
    int g = lane8 & 4, r = lane8 & 3;
    u64 a = shfl8(s, g + ((r + 1) & 3)), b = shfl8(s, g + ((r + 2) & 3)), c = shfl8(s, g + ((r + 3) & 3));
    // row r of circ(2,3,1,1): 2 x_r + 3 x_{r+1} + x_{r+2} + x_{r+3}
    u64 t = gl_add(gl_add(s, a), gl_add(b, c));
    u64 o = gl_add(gl_add(t, s), gl_dbl(a));
    return gl_add(gl_dbl(o), shfl8_xor(o, 4));     // state[i] += out[i] + out[i ^ 4]
}
__device__ __forceinline__ u64 p2x8_permute(u64 s, int lane8) {
    s = p2x8_mds_light(s, lane8);
#pragma unroll 1
    for (int r = 0; r < 4; r++) { s = p2_pow7(gl_add(s, c_p2_ext[0][r][lane8])); s = p2x8_mds_light(s, lane8); }
    const u64 dg = c_p2_diag[lane8];
#pragma unroll 1
    for (int r = 0; r < 22; r++) {
        if (lane8 == 0) s = p2_pow7(gl_add(s, c_p2_int[r]));
        u64 sum = gl_add(s, shfl8_xor(s, 1)); sum = gl_add(sum, shfl8_xor(sum, 2)); sum = gl_add(sum, shfl8_xor(sum, 4));
        s = gl_add(gl_mul(s, dg), sum);
    }
#pragma unroll 1
    for (int r = 0; r < 4; r++) { s = p2_pow7(gl_add(s, c_p2_ext[1][r][lane8])); s = p2x8_mds_light(s, lane8); }
    return s;
}
// lanes 0..3 pass x[lane] / y[lane] (lanes 4..7 pass anything); returns on lane k < 4 the digest word 3 - k
__device__ __forceinline__ u64 p2x8_compress(u64 xw, u64 yw, int lane8) {
    u64 s = lane8 < 4 ? xw : 0ULL;
    s = p2x8_permute(s, lane8);
    if (lane8 < 4) s = yw;
    s = p2x8_permute(s, lane8);
    return s;    // lane k holds state[k]; digest = [s3, s2, s1, s0]
}

// compress(x, y): absorb x -> permute -> overwrite rate with y -> permute -> [s3, s2, s1, s0]
__device__ __forceinline__ void p2_compress(const u64 x[4], const u64 y[4], u64 out[4]) {
    u64 s[8] = {x[0], x[1], x[2], x[3], 0, 0, 0, 0};
    p2_permute(s);
    s[0] = y[0]; s[1] = y[1]; s[2] = y[2]; s[3] = y[3];
    p2_permute(s);
    out[0] = s[3]; out[1] = s[2]; out[2] = s[1]; out[3] = s[0];
}
