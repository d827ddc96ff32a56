// Library-internal context: one process per GPU, one stream, stream-ordered memory pool.
#pragma once
#include "gl.cuh"
#include "../../include/deepprove_b200.h"
#include <mutex>
#include <string>
#include <vector>
#include <cstdio>
#include <cstring>

struct DpCtx {
    bool ready = false;
    int device = -1;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaMemPool_t pool = nullptr;       // this thread's stream-ordered pool (no cross-stream reuse dependencies)
    std::recursive_mutex mu;
    unsigned long long launches = 0;
};
DpCtx &dp_ctx();
void dp_set_error(const std::string &s);
int dp_fail(int code, const std::string &s);

#define DP_CUDA(x)                                                                                          \
    do {                                                                                                    \
        cudaError_t e_ = (x);                                                                               \
        if (e_ != cudaSuccess) return dp_fail(DP_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); \
    } while (0)
#define DP_REQUIRE_CTX()                                                                    \
    std::lock_guard<std::recursive_mutex> lk_(dp_ctx().mu);                                 \
    if (!dp_ctx().ready) return dp_fail(DP_ERR_NO_DEVICE, "dp_init() has not succeeded: no CUDA device context")
#define DP_CHECK(cond, code, msg) \
    do { if (!(cond)) return dp_fail((code), (msg)); } while (0)
void dp_count_launch();
#define DP_LAUNCHED() dp_count_launch()

// optional per-kernel timing with CUDA events on the launch stream (bench.py's roofline leg)
int dp_prof_begin(const char *name, u64 algorithmic_bytes, u64 units = 0);   // returns a token (or -1 when disabled); units: permutations / field ops
void dp_prof_end(int token);
struct DpProfScope {
    int tok;
    DpProfScope(const char *name, u64 bytes, u64 units = 0) : tok(dp_prof_begin(name, bytes, units)) {}
    ~DpProfScope() { dp_prof_end(tok); }
};

// host-side wall-clock accumulators (DP_HOST_PROF=1): where a round-granular call spends its CPU time
struct DpHostTimer {
    const char *name; double t0;
    explicit DpHostTimer(const char *n);
    ~DpHostTimer();
};
#define DP_HOST_TIMED(name) DpHostTimer dp_host_timer_##__LINE__(name)

// stream-ordered allocation helpers
int dp_dev_alloc(void **p, size_t bytes);
int dp_dev_free(void *p);
void dp_arena_defer(bool on);   // hold releases back while several streams are in flight (dp_pcs_commit_many)
// pinned host staging buffers, cached per context (cudaHostAlloc/cudaFreeHost cost ~ms and synchronise)
int dp_pinned_alloc(void **p, size_t bytes);
void dp_pinned_free(void *p);

struct dp_mle {
    void *data = nullptr;   // u64[len] or gle[len]
    u64 len = 0;
    bool is_ext = false;
    bool owned = true;
    u32 num_vars() const { u32 l = 0; while ((1ULL << l) < len) l++; return l; }
    size_t bytes() const { return (size_t)len * (is_ext ? 16 : 8); }
};

static inline int dp_grid_for(u64 work_items, int threads, int max_ctas_per_sm) {
    u64 g = (work_items + threads - 1) / threads;
    u64 cap = (u64)dp_ctx().sm_count * max_ctas_per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// ---- kernel launchers shared between translation units (mle.cu) ----
int dpk_eq_build(const gle *point_host, u32 nv, gle *out_dev);                       // K5
int dpk_fix_high(const void *src, bool src_ext, u64 len, const gle *point_host, u32 k, gle *out_dev);  // K3
int dpk_fold_low(const void *src, bool src_ext, u64 len, gle r, gle *out_dev);      // K2 (stand-alone)
