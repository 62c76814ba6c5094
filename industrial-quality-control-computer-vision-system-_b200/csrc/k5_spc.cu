// k5_spc.cu -- per-batch SPC statistics: the 9-element float64 vector that is the payload of
// the one NCCL all-reduce per step.  Derived from
// analytics/sec_analysis.py:105-147 (_extract_metrics) applied to the detections the
// reference would build from K2's rows:
//   [0] images  [1] sum defect_count  [2] sum defect_count^2  [3] sum ensemble_confidence
//   [4] detections  [5] critical  [6] major  [7] minor  [8] sum bbox width*height
// Severity: models/yolo_model.py:162-172 on the int()-truncated box of :138 (float32 conf vs
// 0.9/0.8 compared in float32 as numpy does; area/1024^2 in float64); final_severity ==
// yolo severity because the mocked region classifier copies it (models/ensemble.py:181-195).
// ensemble_confidence = 0.6*conf + 0.4*min(conf*1.1, 1.0) in float64 (models/ensemble.py:186-208).
#include "common.cuh"

namespace {

constexpr int K5_THREADS = 1024;

// one warp per image, lanes over that image's valid rows only (no per-slot division, no reads of empty rows)
__global__ void __launch_bounds__(K5_THREADS)
k5_spc(const float* __restrict__ det, const int* __restrict__ count, int n, int max_det, double* __restrict__ out9) {
    double acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = warp; i < n; i += K5_THREADS / 32) {
        int c = count[i];
        c = c < 0 ? 0 : (c > max_det ? max_det : c);
        if (lane == 0) {
            acc[0] += 1.0;
            acc[1] += (double)c;
            acc[2] += (double)c * (double)c;
        }
        for (int r = lane; r < c; r += 32) {
            const float* row = det + ((long long)i * max_det + r) * 6;
            const int x1 = (int)row[0], y1 = (int)row[1], x2 = (int)row[2], y2 = (int)row[3];
            const float conf = row[4];
            const long long area = (long long)(x2 - x1) * (long long)(y2 - y1);
            const double na = (double)area / (1024.0 * 1024.0);
            const double cd = (double)conf;
            acc[3] += cd * 0.6 + fmin(cd * 1.1, 1.0) * 0.4;
            acc[4] += 1.0;
            if (conf > 0.9f || na > 0.1) acc[5] += 1.0;
            else if (conf > 0.8f || na > 0.05) acc[6] += 1.0;
            else acc[7] += 1.0;
            acc[8] += (double)area;
        }
    }
    __shared__ double sh[K5_THREADS / 32][9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sh[warp][k] = v;
    }
    __syncthreads();
    if (tid < 9) {
        double v = 0.0;
        for (int w = 0; w < K5_THREADS / 32; ++w) v += sh[w][tid];
        out9[tid] = v;
    }
}

}  // namespace

extern "C" int qc_spc_stats(qc_ctx* ctx, const float* det, const int32_t* count, int n, int max_det, double* out9,
                            void* stream) {
    qc_device_guard dev_guard(ctx);
    if (!ctx) return QC_ERR_INVALID;
    QC_ENTER(ctx, stream);
    if (!det || !count || !out9 || n < 0 || max_det <= 0) QC_FAIL(ctx, QC_ERR_INVALID, "qc_spc_stats: bad argument");
    k5_spc<<<1, K5_THREADS, 0, (cudaStream_t)stream>>>(det, count, n, max_det, out9);
    QC_LAUNCHED_AS(ctx, "k5_spc", stream);
    return QC_OK;
}
