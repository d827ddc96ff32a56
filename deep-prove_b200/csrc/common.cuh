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
    u64 *sync_flag = nullptr; u64 sync_seq = 0;   // dp_stream_sync: this thread's pinned completion word
    std::vector<void *> zero_blocks;               // 256-byte device blocks known to hold zeros (dp_zero_block_get/put)
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
    const char *name; double t0, c0;
    explicit DpHostTimer(const char *n);
    ~DpHostTimer();
};
#define DP_HOST_TIMED(name) DpHostTimer dp_host_timer_##__LINE__(name)

// ---- waiting for the device -------------------------------------------------------------------------------------------
// Every round of the prover ends with the host waiting for a word the GPU writes into mapped pinned memory.  Two modes:
//   DP_WAIT_SPIN  (default): the calling thread spins on the word -- lowest latency, one busy core per proof in flight.
//   DP_WAIT_BLOCK: the thread registers (word, expected value) with ONE process-wide poller thread and sleeps on a futex; the
//     poller sweeps all registered words and wakes the owner.  ~5 us more per wait, but a waiting proof costs no CPU, so the
//     number of proofs in flight per GPU is no longer capped by the host's cores (the measurement boxes give a GPU 16 CPUs).
// dp_wait_flag returns the value seen: `want`, or `fail` (if has_fail), or ~0 - 2 on timeout.
enum : int { DP_WAIT_SPIN = 0, DP_WAIT_BLOCK = 1 };
static constexpr u64 DP_WAIT_TIMEOUT = ~0ULL - 2;
int dp_wait_mode();
u64 dp_wait_flag(volatile u64 *flag, u64 want, bool has_fail, u64 fail, double timeout_s);
// cudaStreamSynchronize for the hot path: in DP_WAIT_BLOCK mode a one-thread kernel raises this thread's pinned word at the
// end of the stream's queued work and the thread sleeps on it (cudaStreamSynchronize itself spins in the driver).
cudaError_t dp_stream_sync(cudaStream_t st);

// Device -> caller's (pageable) host memory, then wait: staged through pinned memory so that the copy call returns at once
// and the wait goes through the wait service.  (cudaMemcpyAsync into pageable memory blocks INSIDE the driver, spinning, until
// every kernel queued before it has finished: 182 such calls per Dense-4M proof were 7 ms of a 44 ms proof and, with 32
// proofs in flight, 6 of the 16 CPUs a GPU gets.)  Several pieces may be gathered before one wait: dp_d2h_begin / _add / _finish.
int dp_d2h(void *host_dst, const void *dev_src, size_t bytes, cudaStream_t st);
struct DpD2H {
    struct Piece { void *dst; size_t off, bytes; };
    void *pin = nullptr; size_t cap = 0, used = 0; std::vector<Piece> pieces; cudaStream_t st;
    explicit DpD2H(cudaStream_t s, size_t reserve_bytes);
    int add(void *host_dst, const void *dev_src, size_t bytes);
    int finish();          // wait for the stream, then scatter into the destinations
    ~DpD2H();
};

// 256-byte device blocks that hold zeros: ticket counters of the multi-block reductions.  The kernels leave them zero again, so a
// block is cleared once when it is first carved out of the arena and then recycled (a memset node per sumcheck -- 151 per
// Dense-4M proof -- was pure driver-call overhead).  Only put back blocks whose kernels have all completed normally.
int dp_zero_block_get(void **p);
void dp_zero_block_put(void *p);

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
