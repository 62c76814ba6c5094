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
    *out = c;
    return QC_OK;
}

void qc_destroy(qc_ctx* ctx) { delete ctx; }

const char* qc_last_error(const qc_ctx* ctx) { return ctx ? ctx->err : "null context"; }

const char* qc_version(void) { return "qcgraft 0.1 (sm_100a)"; }

int qc_sm_count(const qc_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

int64_t qc_launch_count(const qc_ctx* ctx) { return ctx ? (int64_t)ctx->launches.load() : 0; }

size_t qc_sizeof(int which) {
    switch (which) {
        case 0: return sizeof(qc_pre_plan);
        case 1: return sizeof(qc_yolo_params);
        case 2: return sizeof(qc_seg_region);
        case 3: return sizeof(qc_seg_params);
        default: return 0;
    }
}

}  // extern "C"
