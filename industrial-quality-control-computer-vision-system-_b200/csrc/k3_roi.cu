// k3_roi.cu -- K3: ROI-align crop-resize of kept boxes into classifier inputs.
//
// Implements the hand-off that models/ensemble.py:129-142,181-195 only mocks ("would normally
// extract the region and run ResNet on it"): each box becomes a [3,out_h,out_w] tensor with
// torchvision.ops.roi_align semantics (spatial_scale=1; sampling_ratio / aligned as given)
// on the uint8 HWC frame, followed by /255 and Normalize(mean,std) of
// models/resent_model.py:59-66, written as bf16 (RNE) or f32 NCHW.
//
// One thread per output pixel (all 3 channels share the bilinear weights); a warp covers 32
// consecutive output columns so plane stores are contiguous.  Source taps come through the
// read-only path; a ROI's source window is small and stays in L1/L2.
// Algorithmic bytes per ROI: roi_w*roi_h*3 (read) + 3*out_h*out_w*sizeof(out) (write).
#include "common.cuh"
#include <math.h>

namespace {

template <typename OutT>
__global__ void __launch_bounds__(256)
k3_roi_align(const uint8_t* __restrict__ imgs, int n, int H, int W, long long img_stride, int row_stride,
             const float* __restrict__ rois, int K, int PH, int PW, int sampling_ratio, int aligned, int swap_rb,
             const float* __restrict__ norm, OutT* __restrict__ dst) {
    const int k = blockIdx.y;
    const float* r = rois + (long long)k * 5;
    const int img = (int)r[0];
    const float offset = aligned ? 0.5f : 0.f;
    const float x1 = __fsub_rn(r[1], offset), y1 = __fsub_rn(r[2], offset);
    const float x2 = __fsub_rn(r[3], offset), y2 = __fsub_rn(r[4], offset);
    float roi_w = __fsub_rn(x2, x1), roi_h = __fsub_rn(y2, y1);
    if (!aligned) { roi_w = fmaxf(roi_w, 1.f); roi_h = fmaxf(roi_h, 1.f); }
    const float bin_h = __fdiv_rn(roi_h, (float)PH), bin_w = __fdiv_rn(roi_w, (float)PW);
    const int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(roi_h, (float)PH));
    const int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(roi_w, (float)PW));
    const float count = (float)max(gh * gw, 1);
    const bool img_ok = img >= 0 && img < n;
    const uint8_t* base = imgs + (long long)(img_ok ? img : 0) * img_stride;

    float nd[3], nm[3], ns[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { nd[c] = norm[c * 3 + 0]; nm[c] = norm[c * 3 + 1]; ns[c] = norm[c * 3 + 2]; }

    const int total = PH * PW;
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < total; o += gridDim.x * blockDim.x) {
        const int ph = o / PW, pw = o - ph * PW;
        float acc[3] = {0.f, 0.f, 0.f};
        if (img_ok) {
            for (int iy = 0; iy < gh; ++iy) {
                const float yy = __fadd_rn(__fadd_rn(y1, __fmul_rn((float)ph, bin_h)),
                                           __fdiv_rn(__fmul_rn((float)iy + 0.5f, bin_h), (float)gh));
                for (int ix = 0; ix < gw; ++ix) {
                    const float xx = __fadd_rn(__fadd_rn(x1, __fmul_rn((float)pw, bin_w)),
                                               __fdiv_rn(__fmul_rn((float)ix + 0.5f, bin_w), (float)gw));
                    float y = yy, x = xx;
                    if (y < -1.f || y > (float)H || x < -1.f || x > (float)W) continue;
                    if (y <= 0.f) y = 0.f;
                    if (x <= 0.f) x = 0.f;
                    int yl = (int)y, xl = (int)x, yh, xh;
                    if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
                    if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
                    const float ly = __fsub_rn(y, (float)yl), lx = __fsub_rn(x, (float)xl);
                    const float hy = __fsub_rn(1.f, ly), hx = __fsub_rn(1.f, lx);
                    const float w1 = __fmul_rn(hy, hx), w2 = __fmul_rn(hy, lx);
                    const float w3 = __fmul_rn(ly, hx), w4 = __fmul_rn(ly, lx);
                    const uint8_t* p1 = base + (long long)yl * row_stride + xl * 3;
                    const uint8_t* p2 = base + (long long)yl * row_stride + xh * 3;
                    const uint8_t* p3 = base + (long long)yh * row_stride + xl * 3;
                    const uint8_t* p4 = base + (long long)yh * row_stride + xh * 3;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int sc = swap_rb ? 2 - c : c;
                        float v = __fmul_rn(w1, (float)__ldg(p1 + sc));
                        v = __fadd_rn(v, __fmul_rn(w2, (float)__ldg(p2 + sc)));
                        v = __fadd_rn(v, __fmul_rn(w3, (float)__ldg(p3 + sc)));
                        v = __fadd_rn(v, __fmul_rn(w4, (float)__ldg(p4 + sc)));
                        acc[c] = __fadd_rn(acc[c], v);
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = __fdiv_rn(acc[c], count);
            v = __fdiv_rn(__fsub_rn(__fdiv_rn(v, nd[c]), nm[c]), ns[c]);
            OutT* q = dst + ((long long)k * 3 + c) * total + o;
            if (sizeof(OutT) == 2) *reinterpret_cast<uint16_t*>(q) = qc_f2bf(v);
            else *reinterpret_cast<float*>(q) = v;
        }
    }
}

}  // namespace

extern "C" int qc_roi_align(qc_ctx* ctx, const uint8_t* imgs, int n, int h, int w, int64_t img_stride,
                            int32_t row_stride, const float* rois, int K, int out_h, int out_w, int sampling_ratio,
                            int aligned, int swap_rb, const float* norm, int out_dtype, void* dst, void* stream) {
    if (!ctx) return QC_ERR_INVALID;
    if (K == 0) return QC_OK;
    if (!imgs || !rois || !norm || !dst || n <= 0 || h <= 0 || w <= 0 || K < 0 || out_h <= 0 || out_w <= 0 ||
        row_stride < w * 3)
        QC_FAIL(ctx, QC_ERR_INVALID, "qc_roi_align: bad argument");
    if (K > 65535) QC_FAIL(ctx, QC_ERR_UNSUPPORTED, "qc_roi_align: K > 65535 (split the call)");
    cudaStream_t st = (cudaStream_t)stream;
    const int total = out_h * out_w;
    dim3 grid((unsigned)((total + 255) / 256), (unsigned)K);
    if (out_dtype == QC_DT_BF16)
        k3_roi_align<__nv_bfloat16><<<grid, 256, 0, st>>>(imgs, n, h, w, img_stride, row_stride, rois, K, out_h, out_w,
                                                          sampling_ratio, aligned, swap_rb, norm, (__nv_bfloat16*)dst);
    else if (out_dtype == QC_DT_F32)
        k3_roi_align<float><<<grid, 256, 0, st>>>(imgs, n, h, w, img_stride, row_stride, rois, K, out_h, out_w,
                                                  sampling_ratio, aligned, swap_rb, norm, (float*)dst);
    else
        QC_FAIL(ctx, QC_ERR_INVALID, "qc_roi_align: bad out_dtype");
    QC_LAUNCHED(ctx);
    return QC_OK;
}
