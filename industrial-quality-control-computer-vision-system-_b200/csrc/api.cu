// api.cu -- context management for libqcgraft.
#include "common.cuh"
#include <new>

qc_attr_table& qc_attrs() {
    static qc_attr_table t;
    return t;
}

extern "C" {

int qc_create(int device, qc_ctx** out) {
    if (!out) return QC_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return QC_ERR_CUDA;
    qc_ctx* c = new (std::nothrow) qc_ctx();
    if (!c) return QC_ERR_INVALID;
    c->device = device;
    c->launches = 0;
    c->err[0] = 0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return QC_ERR_CUDA; }
    c->sm_count = prop.multiProcessorCount;
    c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (prop.major != 10) {
        // sm_100a cubins only load on compute capability 10.x
        delete c;
        return QC_ERR_UNSUPPORTED;
    }
    {
        qc_device_guard dev_guard(c);
        if (cudaMalloc(&c->sched, (size_t)QC_SCHED_SLOTS * 32 * sizeof(int)) != cudaSuccess ||
            cudaMemset(c->sched, 0, (size_t)QC_SCHED_SLOTS * 32 * sizeof(int)) != cudaSuccess) {
            cudaGetLastError();
            if (c->sched) cudaFree(c->sched);
            delete c;
            return QC_ERR_CUDA;
        }
    }
    *out = c;
    return QC_OK;
}

void qc_destroy(qc_ctx* ctx) {
    if (ctx && ctx->sched) {
        qc_device_guard dev_guard(ctx);
        cudaFree(ctx->sched);
    }
    if (ctx && ctx->prof_recs) {
        for (int i = 0; i < QC_PROF_MAX; ++i)
            if (ctx->prof_recs[i].ev) cudaEventDestroy(ctx->prof_recs[i].ev);
        delete[] ctx->prof_recs;
    }
    delete ctx;
}

const char* qc_last_error(const qc_ctx* ctx) { return ctx ? ctx->err : "null context"; }

const char* qc_version(void) { return "qcgraft 0.1 (sm_100a)"; }

int qc_sm_count(const qc_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

int64_t qc_launch_count(const qc_ctx* ctx) { return ctx ? (int64_t)ctx->launches.load() : 0; }

int qc_profile_enable(qc_ctx* ctx, int on) {
    if (!ctx) return QC_ERR_INVALID;
    if (on && !ctx->prof_recs) {
        ctx->prof_recs = new (std::nothrow) qc_prof_rec[QC_PROF_MAX]();
        if (!ctx->prof_recs) return QC_ERR_INVALID;
    }
    ctx->prof = on != 0;
    ctx->prof_n = 0;
    return QC_OK;
}

int qc_profile_collect(qc_ctx* ctx, const char** names, float* ms, int cap) {
    if (!ctx || !ctx->prof_recs) return 0;
    qc_device_guard dev_guard(ctx);
    int n = 0;
    for (int i = 0; i < ctx->prof_n; ++i) {
        const qc_prof_rec& r = ctx->prof_recs[i];
        if (!r.name || i == 0) continue;                        // boundary marks carry no kernel
        float t = 0.f;
        if (cudaEventSynchronize(r.ev) != cudaSuccess ||
            cudaEventElapsedTime(&t, ctx->prof_recs[i - 1].ev, r.ev) != cudaSuccess) { cudaGetLastError(); continue; }
        if (n < cap) {
            if (names) names[n] = r.name;
            if (ms) ms[n] = t;
        }
        n += 1;
    }
    ctx->prof_n = 0;
    return n;
}

size_t qc_sizeof(int which) {
    switch (which) {
        case 0: return sizeof(qc_pre_plan);
        case 1: return sizeof(qc_yolo_params);
        case 2: return sizeof(qc_seg_region);
        case 3: return sizeof(qc_seg_params);
        case 4: return sizeof(qc_jpeg_info);
        case 5: return sizeof(qc_rules_row);
        case 6: return sizeof(qc_rules_refined);
        case 7: return sizeof(qc_rules_out);
        default: return 0;
    }
}

}  // extern "C"
