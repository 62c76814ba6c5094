// common.cuh -- shared host/device helpers for libqcgraft (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include "../../include/qcgraft.h"

struct qc_attr_entry { const void* func; int smem; int occ; int threads; int device; };
// optional per-kernel timing (qc_profile_enable): one event after every launch and one at every API entry, all on the
// caller's stream; qc_profile_collect() turns consecutive events into per-kernel durations
struct qc_prof_rec { const char* name; cudaEvent_t ev; };
constexpr int QC_PROF_MAX = 4096;
struct qc_ctx {
    int device;
    int sm_count;
    int max_smem_optin;
    std::atomic<long long> launches;
    char err[512];
    bool prof = false;
    int prof_n = 0;
    qc_prof_rec* prof_recs = nullptr;
    // ring of zero-initialised work counters (one 128-byte line each) for kernels that hand out work items dynamically;
    // a call takes the next line and resets it on its own stream before the launch
    int* sched = nullptr;
    std::atomic<unsigned> sched_i{0};
};
constexpr int QC_SCHED_SLOTS = 64;
static inline void qc_prof_mark(qc_ctx* ctx, const char* name, cudaStream_t st) {
    if (!ctx || !ctx->prof || !ctx->prof_recs || ctx->prof_n >= QC_PROF_MAX) return;
    qc_prof_rec& r = ctx->prof_recs[ctx->prof_n];
    if (!r.ev && cudaEventCreate(&r.ev) != cudaSuccess) return;
    r.name = name;
    cudaEventRecord(r.ev, st);
    ctx->prof_n += 1;
}
// at the top of every kernel-launching entry point (name == nullptr marks a boundary, not a kernel)
#define QC_ENTER(ctx, stream) qc_prof_mark((ctx), nullptr, (cudaStream_t)(stream))

#define QC_FAIL(ctx, code, ...)                                   \
    do {                                                          \
        if (ctx) snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__); \
        return (code);                                            \
    } while (0)

#define QC_CUDA(ctx, expr)                                                              \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess)                                                          \
            QC_FAIL(ctx, QC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                   \
                    cudaGetErrorString(_e), __FILE__, __LINE__);                        \
    } while (0)

// every kernel launch goes through this so qc_launch_count() is honest
#define QC_LAUNCHED_AS(ctx, name, st)                                                   \
    do {                                                                                \
        (ctx)->launches.fetch_add(1, std::memory_order_relaxed);                        \
        QC_CUDA(ctx, cudaGetLastError());                                               \
        if ((ctx)->prof) qc_prof_mark((ctx), (name), (cudaStream_t)(st));               \
    } while (0)

// Every entry point that launches kernels runs under this guard: the kernels, the stream and the pointers belong to
// ctx->device whatever the caller's current device is (a pipeline on cuda:1 called while torch's current device is 0).
struct qc_device_guard {
    int prev = -1, dev = -1;
    explicit qc_device_guard(const qc_ctx* ctx) {
        if (!ctx) return;
        dev = ctx->device;
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
    }
    ~qc_device_guard() {
        if (prev >= 0 && prev != dev) cudaSetDevice(prev);
    }
    qc_device_guard(const qc_device_guard&) = delete;
    qc_device_guard& operator=(const qc_device_guard&) = delete;
};

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) + occupancy query, cached per (device, kernel, smem, threads):
// both are host-side driver calls worth several microseconds, too slow to repeat on every launch.  The
// attribute is per-function, per-DEVICE state (last value wins), so the table is process-global (api.cu), keyed by
// device, and the limit is only ever raised.
#include <mutex>
struct qc_attr_table {
    std::mutex mu;
    qc_attr_entry e[256];
    int n = 0;
};
qc_attr_table& qc_attrs();      // defined in api.cu

template <typename F>
static inline int qc_prepare(qc_ctx* ctx, F func, int smem, int threads, int* occ_out) {
    const void* key = reinterpret_cast<const void*>(func);
    qc_attr_table& t = qc_attrs();
    std::lock_guard<std::mutex> lock(t.mu);
    int prev_max = 0;
    for (int i = 0; i < t.n; ++i) {
        if (t.e[i].func != key || t.e[i].device != ctx->device) continue;
        if (t.e[i].smem == smem && t.e[i].threads == threads) {
            if (occ_out) *occ_out = t.e[i].occ;
            return QC_OK;
        }
        if (t.e[i].smem > prev_max) prev_max = t.e[i].smem;
    }
    // static + dynamic shared memory above 48 KB needs the opt-in; kernels here carry up to 8 KB of static tables
    if (smem > 32 * 1024 && smem > prev_max)
        QC_CUDA(ctx, cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int occ = 1;
    QC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, func, threads, (size_t)smem));
    if (occ < 1) occ = 1;
    if (t.n < 256) t.e[t.n++] = qc_attr_entry{key, smem, occ, threads, ctx->device};
    if (occ_out) *occ_out = occ;
    return QC_OK;
}

static inline size_t qc_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

#ifdef __CUDACC__
// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk -> SASS UBLKCP) ---------------------------
__device__ __forceinline__ uint32_t qc_smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void qc_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(qc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void qc_fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void qc_fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void qc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(qc_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void qc_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(qc_smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-B aligned
__device__ __forceinline__ void qc_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     qc_smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(qc_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ uint16_t qc_f2bf(float f) {
    return __bfloat16_as_ushort(__float2bfloat16_rn(f));
}
__device__ __forceinline__ float qc_bf2f(uint16_t b) { return __uint_as_float(((uint32_t)b) << 16); }
// streaming 128-bit store (outputs are written once, never re-read by this kernel)
__device__ __forceinline__ void qc_st_cs(void* p, uint4 v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
#endif
