// Powers of the 2^32-th root of unity g (Goldilocks two_adic_generator(32)): g^e = T0[e & 2047] * T1[(e >> 11) & 2047] * T2[e >> 22].
// The table lives in HBM once per process (built by basefold.cu); get_root_of_unity(n)^j == g^(j << (32 - n)) (tensor.rs:220-231).
#pragma once
#include "gl.cuh"
struct PowTab { const u64 *t0, *t1, *t2; };
__device__ __forceinline__ u64 tab_pow(const PowTab &t, u64 e) {
    return gl_mul(gl_mul(t.t0[e & 2047], t.t1[(e >> 11) & 2047]), t.t2[(e >> 22) & 2047]);
}
int dp_root_powtab(PowTab *out);   // basefold.cu: builds the table on first use (thread-safe)
