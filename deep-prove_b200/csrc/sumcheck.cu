// Sumcheck prover rounds on device (K1 + K2 fused).
// Reference: sumcheck/src/prover.rs:498-741 (prove_parallel / prove_round_and_update_state_parallel),
// sumcheck_macro/src/lib.rs:46-326 (the per-pair arithmetic), multilinear_extensions/src/mle.rs:631-712
// (fix_variables_parallel).
//
// B200 design: ONE launch per round for the whole VirtualPolynomial.  grid.y = product index, grid.x
// tiles the pairs; each operand is streamed once with 16-byte loads.  From round 2 on the fold by the
// previous challenge is fused into the same pass: a thread reads 4 consecutive elements of the old
// table, folds them to one adjacent pair, the owning product writes the pair to the (half-size)
// ping-pong table, and every product using the operand feeds the pair straight into the round
// polynomial accumulators.  Accumulators live in registers, are combined with warp shuffles, then per
// block through shared memory, and the last block of each product (ticket counter) reduces the block
// partials -- field addition is exact, so the summation order is free and the result is bit-identical
// to the reference's rayon fold.  The (deg+1) sums per product go back through a pinned buffer; the
// O(deg^2) glue (2^k multiplicity, coefficient, barycentric extrapolation -- prover.rs:713-724,
// util.rs:101-136) runs on the host inside dp_sc_round, exactly where the reference has it.
#include "common.cuh"
#include <mutex>
#include <algorithm>

enum : u32 { OPM_B = 0, OPM_E = 1, OPM_BF = 2, OPM_EF = 3 };
struct ScOp { const void *src; gle *dst; u32 mode; u32 pad; };
struct ScProd { ScOp op[5]; u64 npairs; u32 d; u32 allbase; u32 konst; u32 pad; };
static constexpr int SC_NACC = 6;
static constexpr int SC_THREADS = 256;

__device__ __forceinline__ gle fold_b(u64 a, u64 b, gle r) { return e_add(e_mul_base(r, gl_sub(b, a)), e_from_base(a)); }
__device__ __forceinline__ gle fold_e(gle a, gle b, gle r) { return e_add(a, e_mul(e_sub(b, a), r)); }
// CG: L1-bypassing loads (ld.global.cg) -- the resident cluster kernel reads tables other CTAs of the cluster wrote a round earlier
template <bool CG> __device__ __forceinline__ gle ldx_e(const gle *p) {
    if (CG) { ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2 *>(p)); return e_make(v.x, v.y); }
    return ld_e(p);
}
template <bool CG> __device__ __forceinline__ ulonglong2 ldx_b2(const u64 *p) { return CG ? __ldcg(reinterpret_cast<const ulonglong2 *>(p)) : ld_b2(p); }
template <bool CG> __device__ __forceinline__ u64 ldx_b(const u64 *p) { return CG ? __ldcg(p) : *p; }

template <bool CG = false>
__device__ __forceinline__ void sc_load_pair(const ScOp &op, u64 i, gle r, gle &lo, gle &hi) {
    switch (op.mode) {
    case OPM_B: {
        ulonglong2 v = ldx_b2<CG>((const u64 *)op.src + 2 * i);
        lo = e_from_base(v.x); hi = e_from_base(v.y);
    } break;
    case OPM_E: {
        const gle *s = (const gle *)op.src + 2 * i;
        lo = ldx_e<CG>(s); hi = ldx_e<CG>(s + 1);
    } break;
    case OPM_BF: {
        const u64 *s = (const u64 *)op.src + 4 * i;
        ulonglong2 v0 = ldx_b2<CG>(s), v1 = ldx_b2<CG>(s + 2);
        lo = fold_b(v0.x, v0.y, r); hi = fold_b(v1.x, v1.y, r);
        if (op.dst) { st_e(op.dst + 2 * i, lo); st_e(op.dst + 2 * i + 1, hi); }
    } break;
    default: {
        const gle *s = (const gle *)op.src + 4 * i;
        gle f0 = ldx_e<CG>(s), f1 = ldx_e<CG>(s + 1), f2 = ldx_e<CG>(s + 2), f3 = ldx_e<CG>(s + 3);
        lo = fold_e(f0, f1, r); hi = fold_e(f2, f3, r);
        if (op.dst) { st_e(op.dst + 2 * i, lo); st_e(op.dst + 2 * i + 1, hi); }
    } break;
    }
}
// length-1 operands (sumcheck_macro/src/lib.rs:236-241): value at every evaluation point
template <bool CG = false>
__device__ __forceinline__ gle sc_load_const(const ScOp &op, gle r) {
    switch (op.mode) {
    case OPM_B: return e_from_base(ldx_b<CG>((const u64 *)op.src));
    case OPM_E: return ldx_e<CG>((const gle *)op.src);
    case OPM_BF: { ulonglong2 v = ldx_b2<CG>((const u64 *)op.src); gle x = fold_b(v.x, v.y, r); if (op.dst) st_e(op.dst, x); return x; }
    default: { const gle *s = (const gle *)op.src; gle x = fold_e(ldx_e<CG>(s), ldx_e<CG>(s + 1), r); if (op.dst) st_e(op.dst, x); return x; }
    }
}

// Out-of-line copies for the latency-bound kernels (general rounds, resident rounds).  Their cost is instruction FETCH, not issue:
// with every E x E (122 instructions) and every operand's 4-way mode switch inlined, k_sc_res<3> was 6.9 k instructions (110 KB) and
// k_sc_round<3> 5.5 k -- more than the SM's instruction cache holds -- so a round of a few hundred pairs spent most of its ~15 us
// refetching its own code from L2.  One shared copy of each costs a call (~20 cycles) per use and keeps the round loop resident.
__device__ __noinline__ gle e_mul_ni(gle a, gle b) { return e_mul(a, b); }
struct gle2 { gle lo, hi; };
template <bool CG>
__device__ __noinline__ gle2 sc_load_pair_ni(const ScOp *op, u64 i, gle r, bool store = true) {
    gle2 o;
    switch (op->mode) {
    case OPM_B: { ulonglong2 v = ldx_b2<CG>((const u64 *)op->src + 2 * i); o.lo = e_from_base(v.x); o.hi = e_from_base(v.y); } break;
    case OPM_E: { const gle *s = (const gle *)op->src + 2 * i; o.lo = ldx_e<CG>(s); o.hi = ldx_e<CG>(s + 1); } break;
    case OPM_BF: {
        const u64 *s = (const u64 *)op->src + 4 * i;
        ulonglong2 v0 = ldx_b2<CG>(s), v1 = ldx_b2<CG>(s + 2);
        o.lo = fold_b(v0.x, v0.y, r); o.hi = fold_b(v1.x, v1.y, r);
        if (store && op->dst) { st_e(op->dst + 2 * i, o.lo); st_e(op->dst + 2 * i + 1, o.hi); }
    } break;
    default: {
        const gle *s = (const gle *)op->src + 4 * i;
        gle f0 = ldx_e<CG>(s), f1 = ldx_e<CG>(s + 1), f2 = ldx_e<CG>(s + 2), f3 = ldx_e<CG>(s + 3);
        o.lo = e_add(f0, e_mul_ni(e_sub(f1, f0), r)); o.hi = e_add(f2, e_mul_ni(e_sub(f3, f2), r));
        if (store && op->dst) { st_e(op->dst + 2 * i, o.lo); st_e(op->dst + 2 * i + 1, o.hi); }
    } break;
    }
    return o;
}

// a[t] += p with a runtime t but register-resident a[] (predicated select, no local memory)
template <int N> __device__ __forceinline__ void acc_add_dyn(u64 (&a)[N], int t, u64 p) {
#pragma unroll
    for (int k = 0; k < N; k++) if (k == t) a[k] = gl_add(a[k], p);
}
template <int N> __device__ __forceinline__ void acc_add_dyn_e(gle (&a)[N], int t, gle p) {
#pragma unroll
    for (int k = 0; k < N; k++) if (k == t) a[k] = e_add(a[k], p);
}

// General body: any mix of operand modes, decided at run time per operand (mixed Base/Ext products, shared MLEs, constants).
// The uniform-mode rounds that carry the bandwidth (all operands Base / all folding) take the lean kernels below instead.
template <int D, bool CG = false>
__device__ __forceinline__ void sc_body(const ScProd &pd, gle r, gle acc[SC_NACC], const u64 tid, const u64 stride) {
    if (pd.konst) {
        if (tid == 0) {
            gle p = sc_load_const<CG>(pd.op[0], r);
#pragma unroll
            for (int j = 1; j < D; j++) p = e_mul_ni(p, sc_load_const<CG>(pd.op[j], r));
#pragma unroll
            for (int t = 0; t <= D; t++) acc[t] = p;
        }
        return;
    }
    if (pd.allbase) {
        // all operands Base and unfolded: stay in F (sumcheck_macro/src/lib.rs:294-299 lifts at the end)
        u64 a[D + 1];
#pragma unroll
        for (int t = 0; t <= D; t++) a[t] = 0;
        for (u64 i = tid; i < pd.npairs; i += stride) {
            u64 cur[D], st[D];
#pragma unroll
            for (int j = 0; j < D; j++) {
                ulonglong2 v = ldx_b2<CG>((const u64 *)pd.op[j].src + 2 * i);
                cur[j] = v.x; st[j] = gl_sub(v.y, v.x);
            }
#pragma unroll 1
            for (int t = 0; t <= D; t++) {
                u64 p = cur[0];
#pragma unroll
                for (int j = 1; j < D; j++) p = gl_mul(p, cur[j]);
                acc_add_dyn<D + 1>(a, t, p);
                if (t < D) {
#pragma unroll
                    for (int j = 0; j < D; j++) cur[j] = gl_add(cur[j], st[j]);
                }
            }
        }
#pragma unroll
        for (int t = 0; t <= D; t++) acc[t] = e_from_base(a[t]);
        return;
    }
    gle a[D + 1];
#pragma unroll
    for (int t = 0; t <= D; t++) a[t] = e_zero();
    for (u64 i = tid; i < pd.npairs; i += stride) {
        gle cur[D], st[D];
#pragma unroll
        for (int j = 0; j < D; j++) {
            const gle2 v = sc_load_pair_ni<CG>(&pd.op[j], i, r);
            cur[j] = v.lo; st[j] = e_sub(v.hi, v.lo);
        }
#pragma unroll 1
        for (int t = 0; t <= D; t++) {
            gle p = cur[0];
#pragma unroll
            for (int j = 1; j < D; j++) p = (pd.op[j].mode == OPM_B) ? e_mul_base(p, cur[j].c0) : e_mul_ni(p, cur[j]);
            acc_add_dyn_e<D + 1>(a, t, p);
            if (t < D) {
#pragma unroll
                for (int j = 0; j < D; j++) cur[j] = e_add(cur[j], st[j]);
            }
        }
    }
#pragma unroll
    for (int t = 0; t <= D; t++) acc[t] = a[t];
}

// Small resident rounds (fewer pairs than a quarter of the lanes a product owns): one (pair, evaluation point) per LANE instead of one
// pair per lane.  The 2^tpl lanes of a pair load and fold it redundantly (same instructions, no extra time), then each computes the
// product at ITS point only -- the dependent chain per lane drops from D folds x 2 + (D+1)(D-1) extension multiplications to
// D x 2 + (D-1) (8 instead of 14 for D = 3), which is most of a small round's device time (phase clocks: 6.2 us of "work" for <= 64
// pairs).  Lane t == 0 of a pair writes the folded table; the other lanes' accumulators stay zero, so the ordinary reduction applies.
template <int D, bool CG>
__device__ __forceinline__ void sc_body_split(const ScProd &pd, gle r, gle acc[SC_NACC], const u64 L, const u32 tpl) {
    const u64 i = L >> tpl; const u32 t = (u32)L & ((1u << tpl) - 1);
    if (i >= pd.npairs || t > (u32)D) return;
    gle cur[D];
#pragma unroll
    for (int j = 0; j < D; j++) {
        const gle2 v = sc_load_pair_ni<CG>(&pd.op[j], i, r, t == 0);
        const gle st = e_sub(v.hi, v.lo);
        gle c = v.lo;
        for (u32 q = 0; q < t; q++) c = e_add(c, st);
        cur[j] = c;
    }
    gle p = cur[0];
#pragma unroll
    for (int j = 1; j < D; j++) p = (pd.op[j].mode == OPM_B) ? e_mul_base(p, cur[j].c0) : e_mul_ni(p, cur[j]);
#pragma unroll
    for (int tt = 0; tt <= D; tt++) if (tt == (int)t) acc[tt] = p;
}

// completion signal: `out` and `flag` live in mapped pinned host memory, so the round message reaches the host
// with the kernel's own stores -- no D2H copy node, and the host spins on the flag instead of a stream sync.
// Called by ONE thread after a block barrier that follows the message stores of its block mates: its system-scope fence
// orders those stores (observed through the barrier) before the flag.
__device__ __forceinline__ void sc_signal(u64 seq, u64 *flag, u32 *done) {
    if (gridDim.y == 1) { __threadfence_system(); *(volatile u64 *)flag = seq; return; }
    __threadfence_system();
    u32 t = atomicAdd(done, 1u);
    if (t == gridDim.y - 1) { *done = 0; __threadfence_system(); *(volatile u64 *)flag = seq; }
}

// Block + grid reduction of the per-thread accumulators of one product (blockIdx.y) and hand-over to the host:
// registers -> warp shuffles -> shared memory -> one partial per block; the last block to finish (ticket) adds the block
// partials.  Field addition is exact, so the summation order is free and the message is bit-identical to the reference's fold.
// The sums are LAZY: canonical limbs are added as plain 96-bit integers (add.cc / addc: 2 instructions and one carry of latency per
// step instead of a 9-instruction modular add) and reduced once per stage -- the reduction tree is the serial part of every round.
struct wsum96 { u64 lo; u32 hi; };
__device__ __forceinline__ void ws_add(wsum96 &a, u64 lo, u32 hi) { asm("{\n\tadd.cc.u64 %0, %0, %2;\n\taddc.u32 %1, %1, %3;\n\t}" : "+l"(a.lo), "+r"(a.hi) : "l"(lo), "r"(hi)); }
__device__ __forceinline__ u64 ws_reduce(const wsum96 &a) { return gl_reduce128(a.lo, (u64)a.hi); }
__device__ __forceinline__ void ws_warp_reduce(wsum96 &a) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { u64 l = __shfl_down_sync(0xffffffffu, a.lo, d); u32 h = __shfl_down_sync(0xffffffffu, a.hi, d); ws_add(a, l, h); }
}
template <int BLOCK>
__device__ __forceinline__ void sc_epilogue(const gle (&acc)[SC_NACC], const int nacc, gle *__restrict__ partials, u32 *__restrict__ counters,
                                            gle *__restrict__ out, u64 seq, u64 *flag, u32 *done) {
    __shared__ u64 wlo[BLOCK / 32][SC_NACC][2];
    __shared__ u32 whi[BLOCK / 32][SC_NACC][2];
    __shared__ bool is_last;
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t < SC_NACC; t++) if (t < nacc) {
        wsum96 v0 = {acc[t].c0, 0}, v1 = {acc[t].c1, 0};
        ws_warp_reduce(v0); ws_warp_reduce(v1);
        if (lane == 0) { wlo[warp][t][0] = v0.lo; whi[warp][t][0] = v0.hi; wlo[warp][t][1] = v1.lo; whi[warp][t][1] = v1.hi; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * nacc) {                     // thread (t, limb) adds the warps' sums and reduces once
        const u32 t = threadIdx.x >> 1, limb = threadIdx.x & 1;
        wsum96 v = {0, 0};
#pragma unroll
        for (int w = 0; w < BLOCK / 32; w++) ws_add(v, wlo[w][t][limb], whi[w][t][limb]);
        const u64 r = ws_reduce(v);
        u64 *dst = gridDim.x == 1 ? (u64 *)(out + (u64)blockIdx.y * SC_NACC + t) : (u64 *)(partials + ((u64)blockIdx.y * gridDim.x + blockIdx.x) * SC_NACC + t);
        dst[limb] = r;
    }
    if (gridDim.x == 1) { __syncthreads(); if (threadIdx.x == 0) sc_signal(seq, flag, done); return; }   // no cross-block stage
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 ticket = atomicAdd(counters + blockIdx.y, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // last block of this product: all threads add the block partials (independent L2 loads in flight), then the same tree
    wsum96 v[SC_NACC][2];
#pragma unroll
    for (int t = 0; t < SC_NACC; t++) { v[t][0] = {0, 0}; v[t][1] = {0, 0}; }
    for (u32 x = threadIdx.x; x < gridDim.x; x += BLOCK) {
        const gle *pp = partials + ((u64)blockIdx.y * gridDim.x + x) * SC_NACC;
#pragma unroll
        for (int t = 0; t < SC_NACC; t++) if (t < nacc) {
            ulonglong2 q = __ldcg(reinterpret_cast<const ulonglong2 *>(pp + t));
            ws_add(v[t][0], q.x, 0); ws_add(v[t][1], q.y, 0);
        }
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < SC_NACC; t++) if (t < nacc) {
        ws_warp_reduce(v[t][0]); ws_warp_reduce(v[t][1]);
        if (lane == 0) { wlo[warp][t][0] = v[t][0].lo; whi[warp][t][0] = v[t][0].hi; wlo[warp][t][1] = v[t][1].lo; whi[warp][t][1] = v[t][1].hi; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * nacc) {
        const u32 t = threadIdx.x >> 1, limb = threadIdx.x & 1;
        wsum96 w = {0, 0};
#pragma unroll
        for (int k = 0; k < BLOCK / 32; k++) ws_add(w, wlo[k][t][limb], whi[k][t][limb]);
        ((u64 *)(out + (u64)blockIdx.y * SC_NACC + t))[limb] = ws_reduce(w);
    }
    __syncthreads();
    if (threadIdx.x == 0) { counters[blockIdx.y] = 0; sc_signal(seq, flag, done); }
}

template <int DSEL>   // DSEL = 0: any degree (mixed-degree polynomials); 1..5: every product has this degree
__global__ void __launch_bounds__(SC_THREADS)
k_sc_round(const ScProd *__restrict__ descs, const __grid_constant__ ScProd desc0, gle r, gle *__restrict__ partials, u32 *__restrict__ counters,
           gle *__restrict__ out, u64 seq, u64 *flag, u32 *done) {
    __shared__ ScProd pd;
    {
        // single-product polynomials carry their descriptor in the kernel parameters; otherwise `descs` is mapped host
        // memory for small grids (no copy node) and a device copy for large ones (hundreds of CTAs each fetching 216 B
        // over PCIe was the dominant cost of large rounds: ncu r01b showed 44 % of stall cycles at this barrier)
        const u64 *s = descs ? (const u64 *)(descs + blockIdx.y) : (const u64 *)&desc0;
        u64 *d = (u64 *)&pd;
        for (int k = threadIdx.x; k < (int)(sizeof(ScProd) / 8); k += blockDim.x) d[k] = s[k];
    }
    __syncthreads();
    gle acc[SC_NACC];
#pragma unroll
    for (int t = 0; t < SC_NACC; t++) acc[t] = e_zero();
    const u64 gtid = (u64)blockIdx.x * blockDim.x + threadIdx.x, gstride = (u64)gridDim.x * blockDim.x;
    if (DSEL != 0) sc_body<DSEL == 0 ? 1 : DSEL>(pd, r, acc, gtid, gstride);
    else switch (pd.d) {
    case 1: sc_body<1>(pd, r, acc, gtid, gstride); break;
    case 2: sc_body<2>(pd, r, acc, gtid, gstride); break;
    case 3: sc_body<3>(pd, r, acc, gtid, gstride); break;
    case 4: sc_body<4>(pd, r, acc, gtid, gstride); break;
    default: sc_body<5>(pd, r, acc, gtid, gstride); break;
    }
    sc_epilogue<SC_THREADS>(acc, DSEL != 0 ? DSEL + 1 : (int)pd.d + 1, partials, counters, out, seq, flag, done);
}

// ---- lean kernels for the rounds that move the bytes ---------------------------------------------------------------------
// Every operand of every product is in the SAME storage mode this round (all Base: the first round of a Base polynomial; all
// folding Base->Ext: its second round; all folding Ext->Ext: every later round; all Ext: the first round of an Ext polynomial)
// and every product has the same degree D in {2, 3}.  The mode and the degree are compile-time, so the body is a straight line
// of loads and multiply chains (no per-operand switch, 64-118 registers instead of 122 for everything), the products of the
// last multiplication are accumulated UNREDUCED in 192-bit (Base) / 2 x 160-bit (Ext) sums and reduced once per thread and
// evaluation point, and the block shape follows the measurements in tools/kbench.cu: Base rounds 256 threads x 2 pairs in
// flight (HBM-bound: 1.9 TB/s at nu = 20), folding rounds 128 threads x 1 pair (issue-bound: 122 instructions per E x E).
struct acc192 { u64 lo, hi; u32 top; };                     // lo + hi 2^64 + top 2^128
__device__ __forceinline__ void acc_mac(acc192 &a, u64 x, u64 y) {
    u64 pl = x * y, ph = __umul64hi(x, y);
    asm("{\n\tadd.cc.u64 %0, %0, %3;\n\taddc.cc.u64 %1, %1, %4;\n\taddc.u32 %2, %2, 0;\n\t}" : "+l"(a.lo), "+l"(a.hi), "+r"(a.top) : "l"(pl), "l"(ph));
}
__device__ __forceinline__ u64 acc_reduce(const acc192 &a) { return gl_reduce160(a.lo, a.hi, a.top); }
struct eacc { acc192 c0, c1; };
__device__ __forceinline__ void eacc_mac(eacc &a, gle x, gle y) {   // a += x * y   (c1 = x0 y1 + x1 y0; c0 = x0 y0 + 7 w, w = weak(x1 y1))
    acc_mac(a.c1, x.c0, y.c1); acc_mac(a.c1, x.c1, y.c0);
    u64 w = gl_reduce128_weak(x.c1 * y.c1, __umul64hi(x.c1, y.c1));
    acc_mac(a.c0, x.c0, y.c0); acc_mac(a.c0, w, 7ULL);
}
template <int MODE> struct ScLeanShape { static constexpr int BLOCK = MODE == OPM_B ? 256 : 128, MINB = 4, PPT = MODE == OPM_B ? 2 : 1; };

template <int D, int MODE>
__global__ void __launch_bounds__(ScLeanShape<MODE>::BLOCK, ScLeanShape<MODE>::MINB)
k_sc_lean(const ScProd *__restrict__ descs, const __grid_constant__ ScProd desc0, gle r, gle *__restrict__ partials, u32 *__restrict__ counters,
          gle *__restrict__ out, u64 seq, u64 *flag, u32 *done) {
    constexpr int BLOCK = ScLeanShape<MODE>::BLOCK, PPT = ScLeanShape<MODE>::PPT;
    __shared__ ScProd pd;
    {
        const u64 *s = descs ? (const u64 *)(descs + blockIdx.y) : (const u64 *)&desc0;
        u64 *d = (u64 *)&pd;
        for (int k = threadIdx.x; k < (int)(sizeof(ScProd) / 8); k += BLOCK) d[k] = s[k];
    }
    __syncthreads();
    const u64 npairs = pd.npairs, T = (u64)gridDim.x * BLOCK;
    gle acc[SC_NACC];
#pragma unroll
    for (int t = 0; t < SC_NACC; t++) acc[t] = e_zero();
    if (MODE == OPM_B) {
        acc192 a[D + 1];
#pragma unroll
        for (int t = 0; t <= D; t++) { a[t].lo = 0; a[t].hi = 0; a[t].top = 0; }
        const u64 *src[D];
#pragma unroll
        for (int j = 0; j < D; j++) src[j] = (const u64 *)pd.op[j].src;
        for (u64 i0 = (u64)blockIdx.x * BLOCK + threadIdx.x; i0 < npairs; i0 += PPT * T) {
            ulonglong2 v[PPT][D];
#pragma unroll
            for (int p = 0; p < PPT; p++) {
                u64 i = i0 + p * T; const bool ok = i < npairs; if (!ok) i = i0;
#pragma unroll
                for (int j = 0; j < D; j++) v[p][j] = ld_b2(src[j] + 2 * i);
                if (!ok) v[p][0] = make_ulonglong2(0, 0);          // a zero factor contributes nothing at every point
            }
#pragma unroll
            for (int p = 0; p < PPT; p++) {
                u64 c[D], s[D];
#pragma unroll
                for (int j = 0; j < D; j++) { c[j] = v[p][j].x; s[j] = gl_sub(v[p][j].y, v[p][j].x); }
#pragma unroll
                for (int t = 0; t <= D; t++) {
                    u64 q = c[0];
#pragma unroll
                    for (int j = 1; j < D - 1; j++) q = gl_mul_weak(q, c[j]);
                    acc_mac(a[t], q, c[D - 1]);
                    if (t < D) {
#pragma unroll
                        for (int j = 0; j < D; j++) c[j] = gl_add(c[j], s[j]);
                    }
                }
            }
        }
#pragma unroll
        for (int t = 0; t <= D; t++) acc[t] = e_from_base(acc_reduce(a[t]));
    } else {
        eacc a[D + 1];
#pragma unroll
        for (int t = 0; t <= D; t++) { a[t].c0 = {0, 0, 0}; a[t].c1 = {0, 0, 0}; }
        for (u64 i = (u64)blockIdx.x * BLOCK + threadIdx.x; i < npairs; i += T) {
            gle c[D], s[D];
#pragma unroll
            for (int j = 0; j < D; j++) {
                gle lo, hi;
                if (MODE == OPM_BF) {
                    const u64 *q = (const u64 *)pd.op[j].src + 4 * i;
                    ulonglong2 v0 = ld_b2(q), v1 = ld_b2(q + 2);
                    lo = fold_b(v0.x, v0.y, r); hi = fold_b(v1.x, v1.y, r);
                    if (pd.op[j].dst) { st_e(pd.op[j].dst + 2 * i, lo); st_e(pd.op[j].dst + 2 * i + 1, hi); }
                } else if (MODE == OPM_EF) {
                    const gle *q = (const gle *)pd.op[j].src + 4 * i;
                    gle f0 = ld_e(q), f1 = ld_e(q + 1), f2 = ld_e(q + 2), f3 = ld_e(q + 3);
                    lo = fold_e(f0, f1, r); hi = fold_e(f2, f3, r);
                    if (pd.op[j].dst) { st_e(pd.op[j].dst + 2 * i, lo); st_e(pd.op[j].dst + 2 * i + 1, hi); }
                } else {
                    const gle *q = (const gle *)pd.op[j].src + 2 * i;
                    lo = ld_e(q); hi = ld_e(q + 1);
                }
                c[j] = lo; s[j] = e_sub(hi, lo);
            }
#pragma unroll
            for (int t = 0; t <= D; t++) {
                gle q = c[0];
#pragma unroll
                for (int j = 1; j < D - 1; j++) q = e_mul(q, c[j]);
                eacc_mac(a[t], q, c[D - 1]);
                if (t < D) {
#pragma unroll
                    for (int j = 0; j < D; j++) c[j] = e_add(c[j], s[j]);
                }
            }
        }
#pragma unroll
        for (int t = 0; t <= D; t++) acc[t] = e_make(acc_reduce(a[t].c0), acc_reduce(a[t].c1));
    }
    sc_epilogue<BLOCK>(acc, D + 1, partials, counters, out, seq, flag, done);
}


// ---- resident rounds: every remaining round of a small sumcheck in ONE launch of a thread-block CLUSTER --------------------
// Once the tables are small a round is pure latency: launch + two-stage reduction + the host's Fiat-Shamir.  This kernel stays
// resident instead: up to 8 CTAs (one cluster: co-scheduled by the hardware, hardware barrier between them) split the pairs of
// every product by warp, the warp partials meet in a small global scratch, CTA 0 adds them, writes the round message to mapped
// host memory, raises the flag and polls a mapped mailbox for the next challenge (one thread, one PCIe read in flight); the
// challenge reaches the other CTAs through the cluster barrier.  A round costs the compute of its widest product spread over
// <= 2048 threads + two cluster barriers (~0.2 us each) + two PCIe hops + the host sponge -- no launch, no cross-block ticket
// stage, no cold instruction cache.  The folded tables ping-pong in HBM/L2 (the cluster barrier orders the stores; readers use
// L1-bypassing loads).  The per-round bookkeeping the host does for k_sc_round (operand order, fold destinations, ping-pong
// buffers) is replayed identically by every CTA in shared memory.  The kernel cannot hang: the poll gives up after
// SC_RES_TIMEOUT cycles or when the host posts the abort value, and the other CTAs only ever wait on the cluster barrier.
// Opt-in per handle (dp_sc_set_resident_tail): the caller promises not to wait on other work in the same stream between rounds.
static constexpr u32 SC_RES_MAXM = 96, SC_RES_MAXP = 48, SC_RES_MAXCTA = 8;
static constexpr u64 SC_RES_WORK = 2048;                     // resident once sum over products of (pairs this round) <= this (larger rounds are faster in the multi-block kernels: profiles/r02x.log)
static constexpr long long SC_RES_TIMEOUT = 6000000000LL;   // ~3 s at 1.9 GHz
static constexpr u64 SC_TAIL_ABORT = ~0ULL, SC_TAIL_FAILED = ~0ULL - 1;
struct TMle { const void *cur; gle *work; u64 len, len0; u32 is_ext, where; };
struct TProd { u32 n_idx; u32 idx[5]; };
struct ScTail { u32 n_mles, n_products, n_rounds, first_has_challenge; long long *dbg; TMle m[SC_RES_MAXM]; TProd p[SC_RES_MAXP]; };   // dbg: optional per-round phase clocks (DP_SC_RES_DEBUG)

__device__ __forceinline__ u32 cl_rank() { u32 r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ u32 cl_size() { u32 r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
// all threads of all CTAs of the cluster; release/acquire at cluster scope (global stores before it are visible after it)
__device__ __forceinline__ void cl_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int DSEL>
__global__ void __launch_bounds__(SC_THREADS)
k_sc_res(const ScTail *__restrict__ cfg, gle r0, gle *out /* mapped */, gle *pairs /* mapped */, volatile u64 *flag /* mapped */,
         volatile u64 *chal /* mapped: [seq, c0, c1] */, u64 seq0, gle *xpart /* device: warp partials */, u64 *xctl /* device: [status, c0, c1] */) {
    __shared__ TMle sm[SC_RES_MAXM];
    __shared__ TProd sp[SC_RES_MAXP];
    __shared__ ScProd spd[SC_RES_MAXP];
    __shared__ gle *sdst[SC_RES_MAXM];
    __shared__ unsigned char sfold[SC_RES_MAXM], swriter[SC_RES_MAXM];
    const u32 nm = cfg->n_mles, np = cfg->n_products, nr = cfg->n_rounds;
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = SC_THREADS / 32;
    const u32 rank = cl_rank(), ncta = cl_size();
    const u32 W = ncta * nwarps, gw = rank * nwarps + warp;              // warps of the cluster, this warp's index
    // pairs of a product are split over G warps (G a power of two, np * G <= W); with more products than warps a warp takes several
    u32 G = 1; while (np * (G << 1) <= W) G <<= 1;
    for (u32 i = threadIdx.x; i < nm; i += blockDim.x) { sm[i] = cfg->m[i]; swriter[i] = 0xff; }
    for (u32 i = threadIdx.x; i < np; i += blockDim.x) sp[i] = cfg->p[i];
    __syncthreads();
    // the product that writes an MLE's folded table is the first one referencing it (the host's `written` rule); static
    if (threadIdx.x == 0) for (u32 p = 0; p < np; p++) for (u32 j = 0; j < sp[p].n_idx; j++) { u32 mi = sp[p].idx[j]; if (swriter[mi] == 0xff) swriter[mi] = (unsigned char)p; }
    __syncthreads();
    gle r = r0;
    long long *const dbg = cfg->dbg;
#define RES_MARK(ph) do { if (dbg && rank == 0 && threadIdx.x == 0) dbg[k * 8 + (ph)] = clock64(); } while (0)
    // The descriptors of a round depend on the fold bookkeeping only, never on the challenge VALUE: they are rebuilt right after a
    // round's message has been signalled, i.e. in the shadow of the host's Fiat-Shamir (phase clocks: ~2.2 us per round that used
    // to sit between the challenge's arrival and the first load).
    auto build_descs = [&](const bool fold) {
        for (u32 i = threadIdx.x; i < nm; i += blockDim.x) {
            bool f = fold && sm[i].len > 1;
            sfold[i] = f;
            sdst[i] = f ? ((sm[i].where == 1) ? sm[i].work + (sm[i].len0 >> 1) : sm[i].work) : nullptr;
        }
        __syncthreads();
        for (u32 p = threadIdx.x; p < np; p += blockDim.x) {   // the descriptor the host builds for k_sc_round, one thread per product
            const TProd &pr = sp[p]; ScProd &pd = spd[p];
            u32 order[5], c = 0;
            for (u32 j = 0; j < pr.n_idx; j++) { u32 mi = pr.idx[j]; if (sm[mi].is_ext || sfold[mi]) order[c++] = j; }
            for (u32 j = 0; j < pr.n_idx; j++) { u32 mi = pr.idx[j]; if (!(sm[mi].is_ext || sfold[mi])) order[c++] = j; }
            bool allbase = true, wrote[5] = {false, false, false, false, false}; u64 newlen = 0;
            for (u32 jj = 0; jj < pr.n_idx; jj++) {
                u32 mi = pr.idx[order[jj]]; const TMle &m = sm[mi]; ScOp &op = pd.op[jj];
                op.src = m.cur; op.dst = nullptr;
                if (sfold[mi]) {
                    op.mode = m.is_ext ? OPM_EF : OPM_BF;
                    bool dup = false; for (u32 q = 0; q < jj; q++) if (wrote[q] && pr.idx[order[q]] == mi) dup = true;
                    if (swriter[mi] == p && !dup) { op.dst = sdst[mi]; wrote[jj] = true; }
                    newlen = m.len >> 1; allbase = false;
                } else { op.mode = m.is_ext ? OPM_E : OPM_B; newlen = m.len; if (m.is_ext) allbase = false; }
            }
            pd.d = pr.n_idx; pd.allbase = allbase; pd.konst = newlen == 1; pd.npairs = newlen >> 1;
        }
        __syncthreads();
    };
    build_descs(cfg->first_has_challenge != 0);
    for (u32 k = 0; k < nr; k++) {
        const u64 seq = seq0 + k;
        RES_MARK(0);
        if (k > 0) {   // wait for the host's challenge for this round
            if (rank == 0 && threadIdx.x == 0) {
                long long t0 = clock64(); u64 v, status = 0;
                while ((v = chal[0]) != seq) {
                    if (v == SC_TAIL_ABORT) { status = 1; break; }
                    if (clock64() - t0 > SC_RES_TIMEOUT) { status = 2; break; }
                }
                u64 c0 = 0, c1 = 0;
                if (!status) { __threadfence_system(); c0 = chal[1]; c1 = chal[2]; }
                __stcg(xctl + 1, c0); __stcg(xctl + 2, c1); __stcg(xctl, status);
            }
            cl_sync();
            const u64 status = __ldcg(xctl);
            if (status) { if (rank == 0 && threadIdx.x == 0 && status == 2) { __threadfence_system(); *flag = SC_TAIL_FAILED; } return; }   // uniform over the cluster
            r = e_make(__ldcg(xctl + 1), __ldcg(xctl + 2));
        }
        RES_MARK(1);
        RES_MARK(2);
        // work: slot (p, s) = sub-slice s of product p's pairs; warp gw owns slots gw, gw + W, ...
        for (u32 slot = gw; slot < np * G; slot += W) {
            const u32 p = slot / G, s = slot % G;
            const ScProd &pd = spd[p];
            gle acc[SC_NACC];
#pragma unroll
            for (int t = 0; t < SC_NACC; t++) acc[t] = e_zero();
            const u64 first = (u64)s * 32 + lane, stride = (u64)G * 32;
            const u32 tpl = pd.d <= 3 ? 2 : 3;
            const bool split = !pd.konst && !pd.allbase && (pd.npairs << tpl) <= stride;     // uniform over the warps of this product
            if (split) {
                if (DSEL != 0) sc_body_split<DSEL == 0 ? 1 : DSEL, true>(pd, r, acc, first, tpl);
                else switch (pd.d) {
                case 1: sc_body_split<1, true>(pd, r, acc, first, tpl); break;
                case 2: sc_body_split<2, true>(pd, r, acc, first, tpl); break;
                case 3: sc_body_split<3, true>(pd, r, acc, first, tpl); break;
                case 4: sc_body_split<4, true>(pd, r, acc, first, tpl); break;
                default: sc_body_split<5, true>(pd, r, acc, first, tpl); break;
                }
            } else
            if (DSEL != 0) sc_body<DSEL == 0 ? 1 : DSEL, true>(pd, r, acc, first, stride);
            else switch (pd.d) {
            case 1: sc_body<1, true>(pd, r, acc, first, stride); break;
            case 2: sc_body<2, true>(pd, r, acc, first, stride); break;
            case 3: sc_body<3, true>(pd, r, acc, first, stride); break;
            case 4: sc_body<4, true>(pd, r, acc, first, stride); break;
            default: sc_body<5, true>(pd, r, acc, first, stride); break;
            }
            const int nacc = DSEL != 0 ? DSEL + 1 : (int)pd.d + 1;
#pragma unroll
            for (int t = 0; t < SC_NACC; t++) if (t < nacc) {     // lazy 96-bit warp sums, one reduction per limb (see sc_epilogue)
                wsum96 v0 = {acc[t].c0, 0}, v1 = {acc[t].c1, 0};
                ws_warp_reduce(v0); ws_warp_reduce(v1);
                if (lane == 0) __stcg(reinterpret_cast<ulonglong2 *>(xpart + (u64)slot * SC_NACC + t), make_ulonglong2(ws_reduce(v0), ws_reduce(v1)));
            }
        }
        RES_MARK(3);
        cl_sync();                                             // warp partials and folded tables of every CTA are visible
        RES_MARK(4);
        if (rank == 0) {                                       // the message first: it is what the host is waiting for
            // few partials per (product, point): one THREAD each (many products run side by side); many (a single product spread over
            // >= 32 warps): one WARP each, its lanes gather the partials with independent L2 loads -- e3 nu=12: 237 -> 215 us per proof,
            // while the warp form costs the 6-product LogUp shape 10 % (profiles/r03h.log)
            if (G < 32) {
                for (u32 x = threadIdx.x; x < np * SC_NACC; x += blockDim.x) {
                    const u32 p = x / SC_NACC, t = x % SC_NACC;
                    if (t > spd[p].d) continue;
                    gle v = e_zero();
                    for (u32 s = 0; s < G; s++) { ulonglong2 q = __ldcg(reinterpret_cast<const ulonglong2 *>(xpart + ((u64)p * G + s) * SC_NACC + t)); v = e_add(v, e_make(q.x, q.y)); }
                    st_e(out + (u64)p * SC_NACC + t, v);
                }
            } else
            for (u32 x = warp; x < np * SC_NACC; x += nwarps) {
                const u32 p = x / SC_NACC, t = x % SC_NACC;
                if (t > spd[p].d) continue;                    // warp-uniform
                wsum96 v0 = {0, 0}, v1 = {0, 0};
                for (u32 s = lane; s < G; s += 32) { ulonglong2 q = __ldcg(reinterpret_cast<const ulonglong2 *>(xpart + ((u64)p * G + s) * SC_NACC + t)); ws_add(v0, q.x, 0); ws_add(v1, q.y, 0); }
                ws_warp_reduce(v0); ws_warp_reduce(v1);
                if (lane == 0) st_e(out + (u64)p * SC_NACC + t, e_make(ws_reduce(v0), ws_reduce(v1)));
            }
        }
        for (u32 i = threadIdx.x; i < nm; i += blockDim.x) if (sfold[i]) {     // fold bookkeeping (every CTA keeps its own copy)
            TMle &m = sm[i];
            m.cur = sdst[i]; m.where = (sdst[i] == m.work) ? 1 : 2; m.len >>= 1; m.is_ext = 1;
        }
        __syncthreads();
        if (rank == 0) {
            if (k == nr - 1) {   // last round: hand the (<= 2)-entry tables over with the message (k_sc_gather's job)
                for (u32 i = threadIdx.x; i < nm; i += blockDim.x) {
                    const TMle &m = sm[i]; gle a, b;
                    if (m.is_ext) { a = ldx_e<true>((const gle *)m.cur); b = m.len > 1 ? ldx_e<true>((const gle *)m.cur + 1) : a; }
                    else { a = e_from_base(ldx_b<true>((const u64 *)m.cur)); b = m.len > 1 ? e_from_base(ldx_b<true>((const u64 *)m.cur + 1)) : a; }
                    st_e(pairs + 2 * i, a); st_e(pairs + 2 * i + 1, b);
                }
                __syncthreads();
            }
            RES_MARK(5);
            if (threadIdx.x == 0) { __threadfence_system(); *flag = seq; }
            RES_MARK(6);
        }
        if (k + 1 < nr) build_descs(true);                     // next round's descriptors while the host hashes
        // CTAs other than 0 run ahead to the next cluster barrier; xpart is not rewritten before it (the work phase follows it)
    }
}

struct ScFin { const void *src; u32 mode; u32 len; };
__global__ void k_sc_final(const ScFin *__restrict__ f, u32 n, gle r, gle *__restrict__ out) {
    u32 m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    ScFin d = f[m];
    gle v;
    if (d.len == 1) v = d.mode == OPM_E ? ld_e((const gle *)d.src) : e_from_base(*(const u64 *)d.src);
    else if (d.mode == OPM_E) { const gle *s = (const gle *)d.src; v = fold_e(ld_e(s), ld_e(s + 1), r); }
    else { ulonglong2 q = ld_b2((const u64 *)d.src); v = fold_b(q.x, q.y, r); }
    st_e(out + m, v);
}

// after the LAST round every MLE has <= 2 entries: copy them out with the round message so the closing fold
// (prover.rs:544-568) needs no further device round trip
__global__ void k_sc_gather(const ScFin *__restrict__ f, u32 n, gle *__restrict__ out /* 2 per MLE */) {
    u32 m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    ScFin d = f[m];
    gle a, b = e_zero();
    if (d.mode == OPM_E) { const gle *s = (const gle *)d.src; a = ld_e(s); if (d.len > 1) b = ld_e(s + 1); }
    else { const u64 *s = (const u64 *)d.src; a = e_from_base(s[0]); if (d.len > 1) b = e_from_base(s[1]); }
    st_e(out + 2 * m, a); st_e(out + 2 * m + 1, b);
}

// ---------------------------------------------------------------------------------------------------
struct ScMle {
    const void *cur = nullptr; u64 len = 0; bool is_ext = false;
    gle *work = nullptr;  // [len0/2 + len0/4] E: ping (offset 0) / pong (offset len0/2)
    u64 len0 = 0; int where = 0;  // 0 = caller's input, 1 = ping, 2 = pong
};
struct dp_sc {
    u32 n_mles = 0, n_products = 0, max_nv = 0, max_deg = 0, round = 0;
    bool finished = false;
    std::vector<ScMle> mles;
    std::vector<dp_sc_product> products;
    ScProd *d_descs = nullptr, *h_descs = nullptr;
    gle *d_partials = nullptr, *d_out = nullptr, *h_out = nullptr;
    u32 *d_counters = nullptr; bool counters_pooled = false, round_in_flight = false;
    ScFin *d_fin = nullptr, *h_fin = nullptr;
    gle *h_pairs = nullptr; bool have_pairs = false;
    u64 seq = 0; u64 *h_flag = nullptr; u32 *d_done = nullptr;
    int gx = 1;
    bool tail_enabled = false, tail_active = false; u32 tail_last_round = 0;
    ScTail *h_tail = nullptr; u64 *h_chal = nullptr;
    gle *d_xpart = nullptr; u64 *d_xctl = nullptr;        // resident cluster kernel: warp partials, [status, c0, c1]
    long long *h_dbg = nullptr; u32 dbg_rounds = 0;
    u64 last_ops = 0;
    std::vector<gle> challenges;
    u64 last_bytes = 0;
};

static u32 ceil_log2_u64(u64 x) { u32 l = 0; while ((1ULL << l) < x) l++; return l; }

// barycentric extrapolation of a degree-(n-1) polynomial given at 0..n-1 to the point `at`
// (sumcheck/src/util.rs:19-136; exact arithmetic, so plain Lagrange gives the same element)
static gle sc_extrapolate(const gle *evals, u32 n, u64 at) {
    gle res = e_zero();
    for (u32 j = 0; j < n; j++) {
        u64 num = 1, den = 1;
        for (u32 i = 0; i < n; i++) if (i != j) {
            num = gl_mul(num, gl_sub(gl_canon(at), (u64)i));
            den = gl_mul(den, gl_sub((u64)j, (u64)i));
        }
        res = e_add(res, e_mul_base(evals[j], gl_mul(num, gl_inv(den))));
    }
    return res;
}

static int sc_free_all(dp_sc *s) {
    for (auto &m : s->mles) if (m.work) { dp_dev_free(m.work); m.work = nullptr; }
    dp_dev_free(s->d_descs); dp_dev_free(s->d_partials); dp_dev_free(s->d_out); dp_dev_free(s->d_fin);
    if (s->counters_pooled && !s->round_in_flight) dp_zero_block_put(s->d_counters); else dp_dev_free(s->d_counters);   // a round that never signalled may have left tickets behind
    dp_dev_free(s->d_xpart); dp_dev_free(s->d_xctl);
    dp_pinned_free(s->h_flag);
    dp_pinned_free(s->h_descs); dp_pinned_free(s->h_out); dp_pinned_free(s->h_fin); dp_pinned_free(s->h_pairs);
    if (s->h_dbg) {   // DP_SC_RES_DEBUG: device clocks of CTA 0 / thread 0 per round: wait-for-challenge | broadcast | descriptors | work | cluster barrier | sum+store | fence+flag
        dp_stream_sync(dp_ctx().stream);
        for (u32 k = 0; k < s->dbg_rounds && k < 64; k++) {
            const long long *d = s->h_dbg + 8 * k;
            fprintf(stderr, "[sc_res] round %2u cycles: wait %6lld bcast %5lld desc %5lld work %6lld clsync %5lld sum %5lld signal %5lld | total-excl-wait %6lld\n", k,
                    d[1] - d[0], 0LL, d[2] - d[1], d[3] - d[2], d[4] - d[3], d[5] - d[4], d[6] - d[5], d[6] - d[1]);
        }
        dp_pinned_free(s->h_dbg);
    }
    dp_pinned_free(s->h_tail); dp_pinned_free(s->h_chal);
    return DP_OK;
}

// host glue: multiplicity, coefficient, extrapolation, sum over products (prover.rs:694-733)
static int sc_glue(dp_sc *s, uint64_t *out_evals) {
    gle msg[SC_NACC + 1];
    for (u32 t = 0; t <= s->max_deg; t++) msg[t] = e_zero();
    for (u32 p = 0; p < s->n_products; p++) {
        const dp_sc_product &pr = s->products[p];
        u32 d = pr.n_idx;
        gle sum[16];
        u64 len = s->mles[pr.idx[0]].len;
        u32 l2 = std::max<u32>(ceil_log2_u64(len), 1);
        int mult = (int)s->max_nv - (int)(l2 + s->round - 1);  // sumcheck_macro/src/lib.rs:242
        DP_CHECK(mult >= 0, DP_ERR_INVALID, "dp_sc_round: negative num_vars multiplicity");
        gle coef = e_make(pr.coef[0], pr.coef[1]);
        for (u32 t = 0; t <= d; t++) {
            gle v = s->h_out[p * SC_NACC + t];
            if (mult > 0) v = e_mul_base(v, gl_canon(1ULL << mult));
            sum[t] = e_mul(v, coef);
        }
        for (u32 i = 0; i < s->max_deg - d; i++) sum[d + 1 + i] = sc_extrapolate(sum, d + 1, d + 1 + i);
        for (u32 t = 0; t <= s->max_deg; t++) msg[t] = e_add(msg[t], sum[t]);
    }
    for (u32 t = 0; t <= s->max_deg; t++) { out_evals[2 * t] = msg[t].c0; out_evals[2 * t + 1] = msg[t].c1; }
    return DP_OK;
}



// Can a resident kernel talk to the host while it runs?  Not under tools that serialise launches (Nsight Compute blocks the
// launching thread until the profiled kernel has finished, so the kernel would wait for a challenge its host cannot post).
// Checked once per process: known injection environments are refused outright, otherwise a one-thread probe kernel waits
// (<= ~50 ms) for a word the host posts right after the launch call returns.
__global__ void k_tail_probe(volatile u64 *mailbox, volatile u64 *result, long long limit_cycles) {
    long long t0 = clock64(); u64 ok = 2;
    while (clock64() - t0 < limit_cycles) { if (mailbox[0] == 1) { ok = 1; break; } }
    __threadfence_system();
    result[0] = ok;
}
static bool sc_tail_supported() {
    static std::once_flag once; static bool supported = false;
    std::call_once(once, [] {
        if (getenv("DP_SC_NO_TAIL") || getenv("CUDA_INJECTION64_PATH") || getenv("CUDA_INJECTION32_PATH") || getenv("NV_NSIGHT_INJECTION_TRANSPORT_TYPE") ||
            getenv("NV_NSIGHT_INJECTION_PORT_BASE") || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR")) return;
        u64 *pin = nullptr;
        if (dp_pinned_alloc((void **)&pin, 128) != DP_OK) return;
        pin[0] = 0; pin[8] = 0;
        cudaStream_t st = dp_ctx().stream;
        k_tail_probe<<<1, 1, 0, st>>>(pin, pin + 8, 100000000LL); dp_count_launch();
        __atomic_store_n(&pin[0], (u64)1, __ATOMIC_RELEASE);          // reached at once unless the launch call itself waits for the kernel
        if (dp_stream_sync(st) == cudaSuccess) supported = (pin[8] == 1);
        dp_pinned_free(pin);
    });
    return supported;
}

// ---- resident rounds, host side ----
// total pair-items of the round that would start the resident kernel (every product's pairs after this round's fold)
static u64 sc_round_work(const dp_sc *s, bool fold) {
    u64 work = 0;
    for (auto &pr : s->products) { const ScMle &m = s->mles[pr.idx[0]]; u64 nl = (fold && m.len > 1) ? m.len >> 1 : m.len; work += std::max<u64>(nl >> 1, 1); }
    return work;
}
static bool sc_tail_eligible(const dp_sc *s, bool fold) {
    if (!s->tail_enabled || s->n_mles > SC_RES_MAXM || s->n_products > SC_RES_MAXP || s->max_nv - s->round < 2) return false;
    std::vector<char> used(s->n_mles, 0);
    for (auto &pr : s->products) for (u32 j = 0; j < pr.n_idx; j++) used[pr.idx[j]] = 1;
    for (u32 i = 0; i < s->n_mles; i++) if (!used[i] && s->mles[i].len > 1) return false;
    static const u64 limit = [] { const char *e = getenv("DP_SC_RES_WORK"); return e ? (u64)atoll(e) : SC_RES_WORK; }();
    return sc_round_work(s, fold) <= limit;
}
template <int DSEL>
static cudaError_t sc_res_launch(u32 ncta, cudaStream_t st, const ScTail *cfg, gle r, gle *out, gle *pairs, u64 *flag, u64 *chal, u64 seq, gle *xpart, u64 *xctl) {
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(ncta); lc.blockDim = dim3(SC_THREADS); lc.dynamicSmemBytes = 0; lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = ncta; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    return cudaLaunchKernelEx(&lc, k_sc_res<DSEL>, cfg, r, out, pairs, (volatile u64 *)flag, (volatile u64 *)chal, seq, xpart, xctl);
}
// one round served by the resident kernel (started here on its first round)
static int sc_tail_round(dp_sc *s, gle r, bool fold, uint64_t *out_evals) {
    cudaStream_t st = dp_ctx().stream;
    s->seq++;
    if (!s->tail_active) {
        int e = 0;
        if (!s->h_tail && (e = dp_pinned_alloc((void **)&s->h_tail, sizeof(ScTail)))) return e;
        if (!s->h_chal && (e = dp_pinned_alloc((void **)&s->h_chal, 64))) return e;
        ScTail &t = *s->h_tail;
        t.n_mles = s->n_mles; t.n_products = s->n_products; t.n_rounds = s->max_nv - s->round; t.first_has_challenge = fold ? 1 : 0;
        for (u32 i = 0; i < s->n_mles; i++) {
            ScMle &m = s->mles[i];
            if (m.len > 1 && !m.work) { if ((e = dp_dev_alloc((void **)&m.work, sizeof(gle) * ((m.len0 >> 1) + (m.len0 >> 2) + 1)))) return e; }
            t.m[i].cur = m.cur; t.m[i].work = m.work; t.m[i].len = m.len; t.m[i].len0 = m.len0; t.m[i].is_ext = m.is_ext; t.m[i].where = (u32)m.where;
        }
        for (u32 p = 0; p < s->n_products; p++) { t.p[p].n_idx = s->products[p].n_idx; for (u32 j = 0; j < 5; j++) t.p[p].idx[j] = s->products[p].idx[j]; }
        s->h_chal[0] = 0;
        t.dbg = nullptr;
        if (getenv("DP_SC_RES_DEBUG")) { if (!s->h_dbg && (e = dp_pinned_alloc((void **)&s->h_dbg, sizeof(long long) * 8 * 64))) return e; memset(s->h_dbg, 0, sizeof(long long) * 8 * 64); t.dbg = s->h_dbg; s->dbg_rounds = t.n_rounds; }
        // cluster size by the first resident round's work: ~512 pair-items per CTA, a power of two <= 8 (portable cluster size)
        const u64 work = sc_round_work(s, fold);
        u32 ncta = 1; while (ncta < SC_RES_MAXCTA && (u64)ncta * 512 < work) ncta <<= 1;
        const u32 slots = std::max<u32>(ncta * (SC_THREADS / 32), s->n_products);
        if (!s->d_xpart && (e = dp_dev_alloc((void **)&s->d_xpart, sizeof(gle) * SC_NACC * slots))) return e;
        if (!s->d_xctl && (e = dp_dev_alloc((void **)&s->d_xctl, 64))) return e;
        u32 dsel = s->products[0].n_idx;
        for (auto &pr : s->products) if (pr.n_idx != dsel) dsel = 0;
        DpProfScope prof("k_sc_res(resident cluster rounds; time includes the host's Fiat-Shamir between rounds)", 0);
        cudaError_t ce;
        switch (dsel) {
        case 1: ce = sc_res_launch<1>(ncta, st, s->h_tail, r, s->h_out, s->h_pairs, s->h_flag, s->h_chal, s->seq, s->d_xpart, s->d_xctl); break;
        case 2: ce = sc_res_launch<2>(ncta, st, s->h_tail, r, s->h_out, s->h_pairs, s->h_flag, s->h_chal, s->seq, s->d_xpart, s->d_xctl); break;
        case 3: ce = sc_res_launch<3>(ncta, st, s->h_tail, r, s->h_out, s->h_pairs, s->h_flag, s->h_chal, s->seq, s->d_xpart, s->d_xctl); break;
        case 4: ce = sc_res_launch<4>(ncta, st, s->h_tail, r, s->h_out, s->h_pairs, s->h_flag, s->h_chal, s->seq, s->d_xpart, s->d_xctl); break;
        case 5: ce = sc_res_launch<5>(ncta, st, s->h_tail, r, s->h_out, s->h_pairs, s->h_flag, s->h_chal, s->seq, s->d_xpart, s->d_xctl); break;
        default: ce = sc_res_launch<0>(ncta, st, s->h_tail, r, s->h_out, s->h_pairs, s->h_flag, s->h_chal, s->seq, s->d_xpart, s->d_xctl); break;
        }
        DP_LAUNCHED();
        DP_CUDA(ce);
        DP_CUDA(cudaGetLastError());
        s->tail_active = true;
    } else {   // post the challenge: value first, then the sequence number the kernel polls
        s->h_chal[1] = r.c0; s->h_chal[2] = r.c1;
        __atomic_store_n(&s->h_chal[0], s->seq, __ATOMIC_RELEASE);
    }
    if (fold) for (auto &m : s->mles) if (m.len > 1) { m.len >>= 1; m.is_ext = true; m.cur = nullptr; m.where = 3; }   // the tables now live in the kernel's bookkeeping
    s->round += 1;
    s->last_bytes = 0;
    {
        DP_HOST_TIMED("dp_sc_round(sync wait)");
        const bool ok = dp_wait_flag(s->h_flag, s->seq, true, SC_TAIL_FAILED, 8.0) == s->seq;
        if (!ok) { s->h_chal[0] = SC_TAIL_ABORT; dp_stream_sync(st); DP_CHECK(false, DP_ERR_CUDA, "dp_sc_round: resident kernel timed out waiting for a challenge (other work queued in the same stream?)"); }
    }
    if (s->round == s->max_nv) s->have_pairs = true;   // written to mapped memory before the flag
    return sc_glue(s, out_evals);
}

extern "C" {

int dp_sc_create(dp_mle *const *mles, uint32_t n_mles, const dp_sc_product *products, uint32_t n_products,
                 uint32_t max_nv, uint32_t max_deg, dp_sc **out) {
    DP_HOST_TIMED("dp_sc_create");
    DP_REQUIRE_CTX();
    DP_CHECK(mles && products && out && n_mles > 0 && n_products > 0, DP_ERR_INVALID, "dp_sc_create: null/empty argument");
    DP_CHECK(max_nv != 0, DP_ERR_INVALID, "Attempt to prove a constant.");  // prover.rs:590-593
    DP_CHECK(max_nv <= 40, DP_ERR_INVALID, "dp_sc_create: max_num_variables too large");
    for (u32 i = 0; i < n_mles; i++) DP_CHECK(mles[i] != nullptr, DP_ERR_INVALID, "dp_sc_create: null MLE");   // every entry is dereferenced below, referenced by a product or not
    u32 seen_deg = 0;
    for (u32 p = 0; p < n_products; p++) {
        const dp_sc_product &pr = products[p];
        DP_CHECK(pr.n_idx >= 1, DP_ERR_INVALID, "input mle_list is empty");  // virtual_poly.rs:143
        DP_CHECK(pr.n_idx <= 5, DP_ERR_UNSUPPORTED, "do not support degree > 5");  // prover.rs:710
        for (u32 j = 0; j < pr.n_idx; j++) {
            DP_CHECK(pr.idx[j] < n_mles, DP_ERR_INVALID, "dp_sc_create: product index out of range");
            const dp_mle *m = mles[pr.idx[j]];
            DP_CHECK(m != nullptr, DP_ERR_INVALID, "dp_sc_create: null MLE");
            DP_CHECK(m->num_vars() <= max_nv, DP_ERR_INVALID, "invalid max num vars");  // virtual_poly.rs:151-154
            DP_CHECK(m->len == mles[pr.idx[0]]->len, DP_ERR_INVALID, "mle in mle_list must be in same num_vars() in same product");
        }
        seen_deg = std::max(seen_deg, pr.n_idx);
    }
    DP_CHECK(seen_deg <= max_deg, DP_ERR_INVALID, "dp_sc_create: max_degree smaller than a product's degree");
    dp_sc *s = new dp_sc();
    s->n_mles = n_mles; s->n_products = n_products; s->max_nv = max_nv; s->max_deg = max_deg;
    s->mles.resize(n_mles);
    u64 max_pairs = 1;
    for (u32 i = 0; i < n_mles; i++) {
        ScMle &m = s->mles[i];
        m.cur = mles[i]->data; m.len = m.len0 = mles[i]->len; m.is_ext = mles[i]->is_ext; m.where = 0;
        max_pairs = std::max<u64>(max_pairs, m.len >> 1);
    }
    s->products.assign(products, products + n_products);
    for (auto &pr : s->products) { pr.coef[0] = gl_canon(pr.coef[0]); pr.coef[1] = gl_canon(pr.coef[1]); }
    s->gx = (int)std::min<u64>((max_pairs + 127) / 128 + 1, (u64)dp_ctx().sm_count * 16);   // upper bound of any round's grid (partials buffer): lean folding rounds use 128-thread blocks, one pair per thread
    // any failure below releases what was already allocated (dp_sc_destroy tolerates the partially built handle)
    auto build = [&]() -> int {
        int e = 0;
        if ((e = dp_dev_alloc((void **)&s->d_descs, sizeof(ScProd) * n_products))) return e;
        if ((e = dp_dev_alloc((void **)&s->d_partials, sizeof(gle) * SC_NACC * (size_t)s->gx * n_products))) return e;
        if ((e = dp_dev_alloc((void **)&s->d_out, sizeof(gle) * SC_NACC * n_products))) return e;
        if (sizeof(u32) * (n_products + 1) <= 256) { if ((e = dp_zero_block_get((void **)&s->d_counters))) return e; s->counters_pooled = true; }
        else { if ((e = dp_dev_alloc((void **)&s->d_counters, sizeof(u32) * (n_products + 1)))) return e; DP_CUDA(cudaMemsetAsync(s->d_counters, 0, sizeof(u32) * (n_products + 1), dp_ctx().stream)); }
        s->d_done = s->d_counters + n_products;
        if ((e = dp_pinned_alloc((void **)&s->h_flag, 64))) return e;
        *s->h_flag = 0;
        if ((e = dp_dev_alloc((void **)&s->d_fin, sizeof(ScFin) * n_mles + sizeof(gle) * 2 * n_mles))) return e;
        if ((e = dp_pinned_alloc((void **)&s->h_descs, sizeof(ScProd) * n_products))) return e;
        if ((e = dp_pinned_alloc((void **)&s->h_out, sizeof(gle) * std::max<size_t>(SC_NACC * n_products, n_mles)))) return e;
        if ((e = dp_pinned_alloc((void **)&s->h_fin, sizeof(ScFin) * n_mles))) return e;
        if ((e = dp_pinned_alloc((void **)&s->h_pairs, sizeof(gle) * 2 * n_mles))) return e;
        return DP_OK;
    };
    if (int e = build()) { sc_free_all(s); delete s; return e; }
    *out = s;
    return DP_OK;
}

int dp_sc_round(dp_sc *s, const uint64_t *challenge, uint64_t *out_evals) {
    DP_HOST_TIMED("dp_sc_round(total)");
    DP_REQUIRE_CTX();
    DP_CHECK(s && out_evals, DP_ERR_INVALID, "dp_sc_round: null argument");
    DP_CHECK(!s->finished && s->round < s->max_nv, DP_ERR_STATE, "Prover is not active");  // prover.rs:636-639
    gle r = e_zero();
    bool fold = false;
    if (s->round == 0) {
        DP_CHECK(challenge == nullptr, DP_ERR_STATE, "first round should be prover first.");  // prover.rs:655
    } else {
        DP_CHECK(challenge != nullptr, DP_ERR_STATE, "verifier message is empty");  // prover.rs:657
        r = e_make(gl_canon(challenge[0]), gl_canon(challenge[1]));
        s->challenges.push_back(r);
        fold = true;
        if (s->challenges.size() == 1)
            for (auto &m : s->mles) DP_CHECK(m.len > 1, DP_ERR_INVALID, "calling sumcheck on constant");  // prover.rs:667-669
    }
    if (s->tail_active || sc_tail_eligible(s, fold)) return sc_tail_round(s, r, fold, out_evals);
    // plan this round's folds: every MLE with >= 1 variable halves (prover.rs:659-684)
    std::vector<gle *> dst(s->n_mles, nullptr);
    std::vector<char> folds(s->n_mles, 0), written(s->n_mles, 0);
    if (fold) {
        for (u32 i = 0; i < s->n_mles; i++) {
            ScMle &m = s->mles[i];
            if (m.len <= 1) continue;
            if (!m.work) { if (int e = dp_dev_alloc((void **)&m.work, sizeof(gle) * ((m.len0 >> 1) + (m.len0 >> 2) + 1))) return e; }
            folds[i] = 1;
            dst[i] = (m.where == 1) ? m.work + (m.len0 >> 1) : m.work;
        }
    }
    u64 bytes = 0, round_pairs = 1;
    for (u32 p = 0; p < s->n_products; p++) {
        const dp_sc_product &pr = s->products[p];
        ScProd &d = s->h_descs[p];
        memset(&d, 0, sizeof d);
        d.d = pr.n_idx;
        // the macro sorts Ext operands first (sumcheck_macro/src/lib.rs:87-140); multiplication commutes,
        // so only operand 0 must be an Ext one when the product is mixed (the body multiplies INTO it)
        u32 order[5]; u32 k = 0;
        for (u32 j = 0; j < pr.n_idx; j++) { const ScMle &m = s->mles[pr.idx[j]]; if (m.is_ext || folds[pr.idx[j]]) order[k++] = j; }
        for (u32 j = 0; j < pr.n_idx; j++) { const ScMle &m = s->mles[pr.idx[j]]; if (!(m.is_ext || folds[pr.idx[j]])) order[k++] = j; }
        bool allbase = true;
        u64 newlen = 0;
        for (u32 jj = 0; jj < pr.n_idx; jj++) {
            u32 mi = pr.idx[order[jj]];
            const ScMle &m = s->mles[mi];
            ScOp &op = d.op[jj];
            op.src = m.cur;
            if (folds[mi]) {
                op.mode = m.is_ext ? OPM_EF : OPM_BF;
                if (!written[mi]) { op.dst = dst[mi]; written[mi] = 1; bytes += (m.len >> 1) * 16; }
                bytes += m.len * (m.is_ext ? 16 : 8);
                newlen = m.len >> 1;
                allbase = false;
            } else {
                op.mode = m.is_ext ? OPM_E : OPM_B;
                bytes += m.len * (m.is_ext ? 16 : 8);
                newlen = m.len;
                if (m.is_ext) allbase = false;
            }
        }
        d.allbase = allbase ? 1 : 0;
        d.konst = newlen == 1 ? 1 : 0;
        d.npairs = newlen >> 1;
        round_pairs = std::max<u64>(round_pairs, d.npairs);
    }
    // uniform rounds take the lean kernels: every operand of every product in the same mode, one degree in {2, 3}, no constants
    u32 dsel = s->products[0].n_idx;
    for (auto &pr : s->products) if (pr.n_idx != dsel) dsel = 0;
    int lean_mode = -1;
    if ((dsel == 2 || dsel == 3) && round_pairs > 2048) {
        lean_mode = (int)s->h_descs[0].op[0].mode;
        for (u32 p = 0; p < s->n_products && lean_mode >= 0; p++) {
            const ScProd &d = s->h_descs[p];
            if (d.konst) lean_mode = -1;
            for (u32 j = 0; j < d.d && lean_mode >= 0; j++) if ((int)d.op[j].mode != lean_mode) lean_mode = -1;
        }
    }
    // field operations of this round in the operand field: (d-1)(d+1) mul + (5d+4) add per pair, + (1 mul + 2 add) for
    // each of the 2 elements a folding operand contributes to a pair
    u64 ops = 0;
    for (u32 p = 0; p < s->n_products; p++) { const ScProd &d = s->h_descs[p]; u64 nf = 0; for (u32 j = 0; j < d.d; j++) if (d.op[j].mode >= OPM_BF) nf++; ops += d.npairs * ((u64)(d.d - 1) * (d.d + 1) + 5 * d.d + 4 + 6 * nf); }
    s->last_ops = ops;
    // grid sized for THIS round.  lean: Base rounds 256 threads x 2 pairs in flight, folding rounds 128 threads x 1 pair (tools/kbench.cu);
    // general kernel: 2 pairs per thread minimum so small rounds run in a single block
    int gx;
    if (lean_mode == (int)OPM_B) gx = (int)std::min<u64>((round_pairs + 511) / 512, (u64)dp_ctx().sm_count * 4);
    else if (lean_mode >= 0) gx = (int)std::min<u64>((round_pairs + 127) / 128, (u64)s->gx);
    else gx = round_pairs > 8192 ? std::min(s->gx, dp_grid_for(round_pairs, SC_THREADS, 6))
                                 : std::min(s->gx, dp_grid_for(round_pairs <= 256 ? round_pairs : (round_pairs + 1) / 2, SC_THREADS, 4));
    cudaStream_t st = dp_ctx().stream;
    // MLEs no product references still have to be folded (cannot happen through add_mle_list, kept for safety)
    for (u32 i = 0; i < s->n_mles; i++) if (folds[i] && !written[i]) {
        ScMle &m = s->mles[i];
        if (int e = dpk_fold_low(m.cur, m.is_ext, m.len, r, dst[i])) return e;
    }
    dim3 grid((unsigned)gx, s->n_products);
    {
        DpProfScope prof(lean_mode == (int)OPM_B ? "k_sc_lean(msg, Base)" : lean_mode == (int)OPM_E ? "k_sc_lean(msg, Ext)" : lean_mode == (int)OPM_BF ? "k_sc_lean(fold Base->Ext + msg)"
                         : lean_mode == (int)OPM_EF ? "k_sc_lean(fold Ext + msg)" : fold ? "k_sc_round(fold+msg)" : "k_sc_round(msg)", bytes, ops);
        s->seq++; s->round_in_flight = true;
        const ScProd *descs_arg = s->h_descs;                 // mapped pinned memory: no copy node on the latency path
        if (s->n_products == 1) descs_arg = nullptr;          // descriptor travels in the kernel parameters
        else if ((u64)gx * s->n_products > 32) { DP_CUDA(cudaMemcpyAsync(s->d_descs, s->h_descs, sizeof(ScProd) * s->n_products, cudaMemcpyHostToDevice, st)); descs_arg = s->d_descs; }
#define SC_ARGS descs_arg, s->h_descs[0], r, s->d_partials, s->d_counters, s->h_out, s->seq, s->h_flag, s->d_done
#define SC_LEAN(D, M) k_sc_lean<D, M><<<grid, ScLeanShape<M>::BLOCK, 0, st>>>(SC_ARGS)
        if (lean_mode >= 0) {
            switch (lean_mode * 4 + (int)dsel) {
            case OPM_B * 4 + 2: SC_LEAN(2, OPM_B); break;   case OPM_B * 4 + 3: SC_LEAN(3, OPM_B); break;
            case OPM_E * 4 + 2: SC_LEAN(2, OPM_E); break;   case OPM_E * 4 + 3: SC_LEAN(3, OPM_E); break;
            case OPM_BF * 4 + 2: SC_LEAN(2, OPM_BF); break; case OPM_BF * 4 + 3: SC_LEAN(3, OPM_BF); break;
            case OPM_EF * 4 + 2: SC_LEAN(2, OPM_EF); break; default: SC_LEAN(3, OPM_EF); break;
            }
        } else
        switch (dsel) {   // one small kernel per uniform degree keeps the instruction footprint low
        case 1: k_sc_round<1><<<grid, SC_THREADS, 0, st>>>(SC_ARGS); break;
        case 2: k_sc_round<2><<<grid, SC_THREADS, 0, st>>>(SC_ARGS); break;
        case 3: k_sc_round<3><<<grid, SC_THREADS, 0, st>>>(SC_ARGS); break;
        case 4: k_sc_round<4><<<grid, SC_THREADS, 0, st>>>(SC_ARGS); break;
        case 5: k_sc_round<5><<<grid, SC_THREADS, 0, st>>>(SC_ARGS); break;
        default: k_sc_round<0><<<grid, SC_THREADS, 0, st>>>(SC_ARGS); break;
        }
#undef SC_LEAN
#undef SC_ARGS
        DP_LAUNCHED();
    }
    DP_CUDA(cudaGetLastError());
    // commit the folds to the bookkeeping while the GPU works
    for (u32 i = 0; i < s->n_mles; i++) if (folds[i]) {
        ScMle &m = s->mles[i];
        m.cur = dst[i]; m.where = (dst[i] == m.work) ? 1 : 2; m.len >>= 1; m.is_ext = true;
    }
    s->round += 1;
    s->last_bytes = bytes;
    if (s->round == s->max_nv) {   // last round: bring the (<= 2)-entry tables back together with the message
        bool ok = true;
        for (u32 i = 0; i < s->n_mles; i++) { const ScMle &m = s->mles[i]; if (m.len > 2) ok = false; s->h_fin[i].src = m.cur; s->h_fin[i].mode = m.is_ext ? OPM_E : OPM_B; s->h_fin[i].len = (u32)m.len; }
        if (ok) {
            gle *d_pairs = (gle *)((char *)s->d_fin + sizeof(ScFin) * s->n_mles);
            DP_CUDA(cudaMemcpyAsync(s->d_fin, s->h_fin, sizeof(ScFin) * s->n_mles, cudaMemcpyHostToDevice, st));
            k_sc_gather<<<(s->n_mles + 127) / 128, 128, 0, st>>>(s->d_fin, s->n_mles, d_pairs); DP_LAUNCHED();
            DP_CUDA(cudaMemcpyAsync(s->h_pairs, d_pairs, sizeof(gle) * 2 * s->n_mles, cudaMemcpyDeviceToHost, st));
            s->have_pairs = true;
        }
    }
    {
        DP_HOST_TIMED("dp_sc_round(sync wait)");
        if (s->have_pairs) { DP_CUDA(dp_stream_sync(st)); s->round_in_flight = false; }      // last round: the gathered pairs follow the kernel
        else {
            volatile u64 *f = s->h_flag;
            const bool ok = dp_wait_flag(f, s->seq, false, 0, 2.0) == s->seq;
            if (ok) s->round_in_flight = false;
            if (!ok) { DP_CUDA(dp_stream_sync(st)); DP_CHECK(*f == s->seq, DP_ERR_CUDA, "dp_sc_round: kernel finished without signalling"); s->round_in_flight = false; }
        }
    }
    return sc_glue(s, out_evals);
}

int dp_sc_finish(dp_sc *s, const uint64_t *last_challenge, uint64_t *out_final) {
    DP_HOST_TIMED("dp_sc_finish");
    DP_REQUIRE_CTX();
    DP_CHECK(s && last_challenge && out_final, DP_ERR_INVALID, "dp_sc_finish: null argument");
    DP_CHECK(!s->finished && s->round == s->max_nv, DP_ERR_STATE, "dp_sc_finish: rounds not complete");
    gle r = e_make(gl_canon(last_challenge[0]), gl_canon(last_challenge[1]));
    s->challenges.push_back(r);
    if (s->have_pairs) {   // O(#MLEs) host glue on the pairs fetched with the last message
        for (u32 i = 0; i < s->n_mles; i++) {
            gle a = s->h_pairs[2 * i], b = s->h_pairs[2 * i + 1];
            gle v = s->mles[i].len > 1 ? e_add(a, e_mul(e_sub(b, a), r)) : a;
            out_final[2 * i] = v.c0; out_final[2 * i + 1] = v.c1;
        }
        s->finished = true;
        return DP_OK;
    }
    for (u32 i = 0; i < s->n_mles; i++) {
        const ScMle &m = s->mles[i];
        // get_mle_final_evaluations asserts len == 1 after the last fix (prover.rs:479-483)
        DP_CHECK(m.len <= 2, DP_ERR_STATE, "mle.evaluations.len() != 1, must be called after prove_round_and_update_state");
        s->h_fin[i].src = m.cur; s->h_fin[i].mode = m.is_ext ? OPM_E : OPM_B; s->h_fin[i].len = (u32)m.len;
    }
    cudaStream_t st = dp_ctx().stream;
    gle *d_vals = (gle *)((char *)s->d_fin + sizeof(ScFin) * s->n_mles);
    // d_fin region: [ScFin x n][gle x n]; ScFin is 16 bytes so the gle part stays 16-byte aligned
    DP_CUDA(cudaMemcpyAsync(s->d_fin, s->h_fin, sizeof(ScFin) * s->n_mles, cudaMemcpyHostToDevice, st));
    k_sc_final<<<(s->n_mles + 127) / 128, 128, 0, st>>>(s->d_fin, s->n_mles, r, d_vals); DP_LAUNCHED();
    DP_CUDA(cudaGetLastError());
    DP_CUDA(cudaMemcpyAsync(s->h_out, d_vals, sizeof(gle) * s->n_mles, cudaMemcpyDeviceToHost, st));
    DP_CUDA(dp_stream_sync(st));
    for (u32 i = 0; i < s->n_mles; i++) { out_final[2 * i] = s->h_out[i].c0; out_final[2 * i + 1] = s->h_out[i].c1; }
    s->finished = true;
    return DP_OK;
}

int dp_sc_destroy(dp_sc *s) {
    if (!s) return DP_OK;
    DP_HOST_TIMED("dp_sc_destroy");
    std::lock_guard<std::recursive_mutex> lk(dp_ctx().mu);
    if (dp_ctx().ready) {
        if (s->tail_active && s->round < s->max_nv) { s->h_chal[0] = SC_TAIL_ABORT; dp_stream_sync(dp_ctx().stream); }   // release the waiting kernel
        sc_free_all(s);
    }
    delete s;
    return DP_OK;
}

uint64_t dp_sc_last_round_bytes(const dp_sc *s) { return s ? s->last_bytes : 0; }

int dp_sc_set_resident_tail(dp_sc *s, int enable) {
    DP_REQUIRE_CTX();
    DP_CHECK(s && !s->tail_active, DP_ERR_STATE, "dp_sc_set_resident_tail: null handle or tail already running");
    s->tail_enabled = enable != 0 && sc_tail_supported();
    return DP_OK;
}

int dp_sc_current_mle(dp_sc *s, uint32_t idx, dp_mle **out) {
    DP_REQUIRE_CTX();
    DP_CHECK(s && out && idx < s->n_mles, DP_ERR_INVALID, "dp_sc_current_mle: bad argument");
    const ScMle &m = s->mles[idx];
    DP_CHECK(!s->tail_active, DP_ERR_STATE, "dp_sc_current_mle: tables are owned by the resident tail kernel");
    dp_mle *v = new dp_mle();
    v->data = const_cast<void *>(m.cur); v->len = m.len; v->is_ext = m.is_ext; v->owned = false;
    *out = v;
    return DP_OK;
}

}  // extern "C"
