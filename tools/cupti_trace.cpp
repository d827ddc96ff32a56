// Concurrent-kernel timeline of the prover as it really runs (development tool): CUPTI activity records (start / end timestamps,
// grid, block, registers, shared memory, stream) of every kernel, WITHOUT serialising launches -- unlike Nsight Compute, which
// blocks the launching thread and so can never see the resident sumcheck kernel or 16 proofs in flight.
//   g++ -O2 -shared -fPIC -I/usr/local/cuda/include -o deep-prove_b200/libdp_trace.so tools/cupti_trace.cpp -L/usr/local/cuda/lib64 -lcupti -lcudart
//   python tools/trace_concurrent.py 16 32      (loads it with ctypes; writes gpurun_out/trace_*.csv)
#include <cupti.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

struct Rec { std::string name; unsigned long long start, end; unsigned gx, gy, gz, bx, by, bz, regs, smem, stream; };
static std::vector<Rec> g_recs;
static std::mutex g_mu;

static void CUPTIAPI buf_requested(uint8_t **buffer, size_t *size, size_t *max_records) {
    *size = 16u << 20; *buffer = (uint8_t *)aligned_alloc(8, *size); *max_records = 0;
}
static void CUPTIAPI buf_completed(CUcontext, uint32_t, uint8_t *buffer, size_t, size_t valid) {
    CUpti_Activity *r = nullptr;
    std::lock_guard<std::mutex> lk(g_mu);
    while (cuptiActivityGetNextRecord(buffer, valid, &r) == CUPTI_SUCCESS) {
        if (r->kind == CUPTI_ACTIVITY_KIND_CONCURRENT_KERNEL || r->kind == CUPTI_ACTIVITY_KIND_KERNEL) {
            auto *k = (CUpti_ActivityKernel9 *)r;
            Rec x; x.name = k->name ? k->name : "?"; x.start = k->start; x.end = k->end;
            x.gx = k->gridX; x.gy = k->gridY; x.gz = k->gridZ; x.bx = k->blockX; x.by = k->blockY; x.bz = k->blockZ;
            x.regs = k->registersPerThread; x.smem = (unsigned)(k->staticSharedMemory + k->dynamicSharedMemory); x.stream = k->streamId;
            g_recs.push_back(std::move(x));
        }
    }
    free(buffer);
}
extern "C" int dp_trace_start(void) {
    { std::lock_guard<std::mutex> lk(g_mu); g_recs.clear(); }
    static bool registered = false;
    if (!registered) { if (cuptiActivityRegisterCallbacks(buf_requested, buf_completed) != CUPTI_SUCCESS) return 1; registered = true; }
    return cuptiActivityEnable(CUPTI_ACTIVITY_KIND_CONCURRENT_KERNEL) == CUPTI_SUCCESS ? 0 : 2;
}
extern "C" long dp_trace_stop(const char *path) {
    cudaDeviceSynchronize();
    cuptiActivityFlushAll(1);
    cuptiActivityDisable(CUPTI_ACTIVITY_KIND_CONCURRENT_KERNEL);
    std::lock_guard<std::mutex> lk(g_mu);
    FILE *f = fopen(path, "w");
    if (!f) return -1;
    fprintf(f, "name,start_ns,end_ns,grid,block,regs,smem,stream\n");
    for (auto &r : g_recs) fprintf(f, "%s,%llu,%llu,%u,%u,%u,%u,%u\n", r.name.c_str(), r.start, r.end, r.gx * r.gy * r.gz, r.bx * r.by * r.bz, r.regs, r.smem, r.stream);
    fclose(f);
    return (long)g_recs.size();
}

// ---- CUDA runtime API calls: how many, and how long the calling threads sit inside them (driver-lock contention shows up here) ----
#include <atomic>
#include <chrono>
#include <map>
struct ApiStat { std::atomic<unsigned long long> n{0}, ns{0}; };
static ApiStat g_api[1024];
static CUpti_SubscriberHandle g_sub; static bool g_sub_on = false;
static inline unsigned long long now_ns() { return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static void CUPTIAPI api_cb(void *, CUpti_CallbackDomain domain, CUpti_CallbackId cbid, const void *cbdata) {
    if (domain != CUPTI_CB_DOMAIN_RUNTIME_API || cbid >= 1024) return;
    const CUpti_CallbackData *d = (const CUpti_CallbackData *)cbdata;
    if (d->callbackSite == CUPTI_API_ENTER) *d->correlationData = now_ns();
    else { g_api[cbid].n.fetch_add(1, std::memory_order_relaxed); g_api[cbid].ns.fetch_add(now_ns() - *d->correlationData, std::memory_order_relaxed); }
}
extern "C" int dp_trace_api_start(void) {
    for (auto &a : g_api) { a.n = 0; a.ns = 0; }
    if (!g_sub_on) { if (cuptiSubscribe(&g_sub, (CUpti_CallbackFunc)api_cb, nullptr) != CUPTI_SUCCESS) return 1; g_sub_on = true; }
    return cuptiEnableDomain(1, g_sub, CUPTI_CB_DOMAIN_RUNTIME_API) == CUPTI_SUCCESS ? 0 : 2;
}
extern "C" int dp_trace_api_stop(const char *path) {
    if (g_sub_on) cuptiEnableDomain(0, g_sub, CUPTI_CB_DOMAIN_RUNTIME_API);
    FILE *f = fopen(path, "w");
    if (!f) return -1;
    fprintf(f, "api,calls,total_ms,mean_us\n");
    for (int i = 0; i < 1024; i++) if (g_api[i].n) {
        const char *name = nullptr; cuptiGetCallbackName(CUPTI_CB_DOMAIN_RUNTIME_API, i, &name);
        fprintf(f, "%s,%llu,%.3f,%.2f\n", name ? name : "?", g_api[i].n.load(), g_api[i].ns.load() / 1e6, g_api[i].ns.load() / 1e3 / g_api[i].n.load());
    }
    fclose(f);
    return 0;
}
