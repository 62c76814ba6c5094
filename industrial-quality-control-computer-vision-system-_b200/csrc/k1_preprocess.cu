// k1_preprocess.cu -- K1: uint8 HWC -> (stretch | letterbox) bilinear resize -> constant
// border -> channel swap -> LUT normalise -> NCHW bf16/f32 pack.
//
// Replaces (see include/qcgraft.h): cv2.cvtColor(BGR2RGB) inference/detector.py:233-234,
// models/yolo_model.py:79-80; cv2.resize inference/detector.py:240-242; ultralytics
// LetterBox + [...,::-1] + NCHW + float()/255 behind models/yolo_model.py:105-110.
//
// Two kernels:
//   k1_pack_stream : no resize, no border, packed rows (every BASELINE config): the batch
//                    is a flat byte stream.  Persistent CTAs pull 12 KB chunks into shared
//                    memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier, double
//                    buffered), de-interleave the 3-byte pixels from smem with conflict-free
//                    64-bit reads, normalise through a lane-private (bank == lane) LUT and
//                    write each plane with fully coalesced 128-bit streaming stores.
//   k1_generic     : any geometry/alignment.  Exact integer model of cv2's 8-bit INTER_LINEAR
//                    (11-bit coefficients) + border + LUT.
// HBM-bound, no tensor cores: algorithmic bytes = h0*w0*3 + 3*H*W*sizeof(out) per image.
#include "common.cuh"
#include <math.h>
#include <algorithm>

namespace {

constexpr int K1_THREADS = 256;
constexpr int K1_CHUNK_PX = 4096;                 // pixels per chunk (12288 B of input)
constexpr int K1_CHUNK_BYTES = K1_CHUNK_PX * 3;
constexpr int K1_LUT_WORDS = 256 * 32;            // one lane-private table

template <typename T> struct OutPack;
template <> struct OutPack<__nv_bfloat16> {
    static __device__ __forceinline__ uint32_t enc(float f) { return qc_f2bf(f); }
    // 8 pixels -> one uint4
    static __device__ __forceinline__ void store8(__nv_bfloat16* p, const uint32_t* v) {
        uint4 o;
        o.x = v[0] | (v[1] << 16); o.y = v[2] | (v[3] << 16);
        o.z = v[4] | (v[5] << 16); o.w = v[6] | (v[7] << 16);
        qc_st_cs(p, o);
    }
};
template <> struct OutPack<float> {
    static __device__ __forceinline__ uint32_t enc(float f) { return __float_as_uint(f); }
    static __device__ __forceinline__ void store8(float* p, const uint32_t* v) {
        qc_st_cs(p, make_uint4(v[0], v[1], v[2], v[3]));
        qc_st_cs(p + 4, make_uint4(v[4], v[5], v[6], v[7]));
    }
};

// ------------------------------------------------------------------------------------------
// fast path
// ------------------------------------------------------------------------------------------
template <typename OutT, int NT, bool SWAP>
__global__ void __launch_bounds__(K1_THREADS)
k1_pack_stream(const uint8_t* __restrict__ src, OutT* __restrict__ dst, const float* __restrict__ lut,
               long long P, long long src_img_stride, int chunks_per_img, long long total_items) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t* lut_s = reinterpret_cast<uint32_t*>(smem);
    uint8_t* stage = smem + (size_t)NT * K1_LUT_WORDS * 4;
    __shared__ __align__(8) uint64_t bar[2];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    if (tid == 0) {
        qc_mbar_init(&bar[0], 1);
        qc_mbar_init(&bar[1], 1);
        qc_fence_mbar_init();
    }
    // lane-private LUT: word (v*32 + lane) -> bank == lane, so 32 random lookups never conflict
    for (int i = tid; i < NT * K1_LUT_WORDS; i += K1_THREADS) {
        int c = i / K1_LUT_WORDS, v = (i >> 5) & 255;
        lut_s[i] = OutPack<OutT>::enc(lut[c * 256 + v]);
    }
    __syncthreads();

    auto issue = [&](long long item, int buf) {
        long long img = item / chunks_per_img;
        long long chunk = item - img * chunks_per_img;
        long long px0 = chunk * K1_CHUNK_PX;
        long long npx = P - px0 < K1_CHUNK_PX ? P - px0 : K1_CHUNK_PX;
        uint32_t bytes = (uint32_t)(npx * 3);
        qc_mbar_expect_tx(&bar[buf], bytes);
        qc_bulk_g2s(stage + buf * K1_CHUNK_BYTES, src + img * src_img_stride + px0 * 3, bytes, &bar[buf]);
    };

    long long item = blockIdx.x;
    if (tid == 0 && item < total_items) issue(item, 0);

    for (int it = 0; item < total_items; ++it, item += gridDim.x) {
        const int buf = it & 1;
        const long long next = item + gridDim.x;
        if (tid == 0 && next < total_items) issue(next, buf ^ 1);
        qc_mbar_wait(&bar[buf], (it >> 1) & 1);

        const long long img = item / chunks_per_img;
        const long long chunk = item - img * chunks_per_img;
        const long long px0 = chunk * K1_CHUNK_PX;
        const int npx = (int)(P - px0 < K1_CHUNK_PX ? P - px0 : K1_CHUNK_PX);
        const uint8_t* sbuf = stage + buf * K1_CHUNK_BYTES;

#pragma unroll
        for (int g = 0; g < K1_CHUNK_PX / (8 * K1_THREADS); ++g) {
            const int p = (g * K1_THREADS + tid) * 8;      // first pixel of this thread's group
            if (p < npx) {
                const uint2* s2 = reinterpret_cast<const uint2*>(sbuf + p * 3);
                uint2 a = s2[0], b = s2[1], c = s2[2];
                uint32_t w[6] = {a.x, a.y, b.x, b.y, c.x, c.y};
#pragma unroll
                for (int oc = 0; oc < 3; ++oc) {
                    const int sc = SWAP ? 2 - oc : oc;
                    const uint32_t* tab = lut_s + (NT == 1 ? 0 : oc) * K1_LUT_WORDS + lane;
                    uint32_t v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int k = 3 * j + sc;
                        uint32_t byte = (w[k >> 2] >> (8 * (k & 3))) & 255u;
                        v[j] = tab[byte << 5];
                    }
                    OutT* o = dst + (img * 3 + oc) * P + px0 + p;
                    OutPack<OutT>::store8(o, v);
                }
            }
        }
        __syncthreads();   // everyone is done with `buf` before it is refilled two iterations on
    }
}

// ------------------------------------------------------------------------------------------
// generic path (exact cv2 INTER_LINEAR integer model)
// ------------------------------------------------------------------------------------------
struct LinCoef { int s0, s1, a0, a1; };

__device__ __forceinline__ LinCoef lin_coef(int d, int dst, int src, bool clamp_f) {
    // cv2::resize: inv_scale = (double)dst/src; scale = 1./inv_scale;
    // f = (float)((d+0.5)*scale - 0.5); s = cvFloor(f); f -= s; (x only: clamp -> f = 0)
    const double inv_scale = __ddiv_rn((double)dst, (double)src);
    const double scale = __ddiv_rn(1.0, inv_scale);
    float f = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
    int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    if (clamp_f) {
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= src - 1) { s = src - 1; f = 0.f; }
    }
    LinCoef c;
    c.a1 = __float2int_rn(__fmul_rn(f, 2048.f));
    c.a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    c.s0 = min(max(s, 0), src - 1);
    c.s1 = min(max(s + 1, 0), src - 1);
    return c;
}

template <typename OutT>
__global__ void __launch_bounds__(256)
k1_generic(const uint8_t* __restrict__ src, OutT* __restrict__ dst, const float* __restrict__ lut,
           int n, long long src_img_stride, int src_row_stride, qc_pre_plan pl, int pad_value, int swap_rb) {
    __shared__ uint32_t lut_s[3 * 256];
    for (int i = threadIdx.x; i < 768; i += blockDim.x) lut_s[i] = OutPack<OutT>::enc(lut[i]);
    __syncthreads();

    const int groups_per_row = (pl.dst_w + 7) >> 3;
    const long long total = (long long)n * pl.dst_h * groups_per_row;
    const bool resize = (pl.new_h != pl.src_h) || (pl.new_w != pl.src_w);
    const long long plane = (long long)pl.dst_h * pl.dst_w;

    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int gx = (int)(t % groups_per_row);
        const long long r = t / groups_per_row;
        const int y = (int)(r % pl.dst_h);
        const int img = (int)(r / pl.dst_h);
        const uint8_t* simg = src + (long long)img * src_img_stride;
        const int ry = y - pl.top;
        const bool row_in = ry >= 0 && ry < pl.new_h;
        LinCoef cy = {0, 0, 2048, 0};
        if (row_in) {
            if (resize) cy = lin_coef(ry, pl.new_h, pl.src_h, false);
            else cy.s0 = cy.s1 = ry;
        }
        const uint8_t* row0 = simg + (long long)cy.s0 * src_row_stride;
        const uint8_t* row1 = simg + (long long)cy.s1 * src_row_stride;

        uint32_t vals[3][8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int x = gx * 8 + j;
            const int rx = x - pl.left;
            int px[3] = {pad_value, pad_value, pad_value};
            if (row_in && rx >= 0 && rx < pl.new_w && x < pl.dst_w) {
                if (resize) {
                    LinCoef cx = lin_coef(rx, pl.new_w, pl.src_w, true);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        int r0 = (int)__ldg(row0 + cx.s0 * 3 + c) * cx.a0 + (int)__ldg(row0 + cx.s1 * 3 + c) * cx.a1;
                        int r1 = (int)__ldg(row1 + cx.s0 * 3 + c) * cx.a0 + (int)__ldg(row1 + cx.s1 * 3 + c) * cx.a1;
                        px[c] = (((cy.a0 * (r0 >> 4)) >> 16) + ((cy.a1 * (r1 >> 4)) >> 16) + 2) >> 2;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 3; ++c) px[c] = __ldg(row0 + rx * 3 + c);
                }
            }
#pragma unroll
            for (int oc = 0; oc < 3; ++oc) {
                const int sc = swap_rb ? 2 - oc : oc;
                vals[oc][j] = lut_s[oc * 256 + (px[sc] & 255)];
            }
        }
        const int x0 = gx * 8;
#pragma unroll
        for (int oc = 0; oc < 3; ++oc) {
            OutT* o = dst + ((long long)img * 3 + oc) * plane + (long long)y * pl.dst_w + x0;
            const bool full = (x0 + 8 <= pl.dst_w) && ((reinterpret_cast<uintptr_t>(o) & 15) == 0);
            if (full) {
                OutPack<OutT>::store8(o, vals[oc]);
            } else {
                for (int j = 0; j < 8 && x0 + j < pl.dst_w; ++j) {
                    if (sizeof(OutT) == 2) reinterpret_cast<uint16_t*>(o)[j] = (uint16_t)vals[oc][j];
                    else reinterpret_cast<uint32_t*>(o)[j] = vals[oc][j];
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// k1_resize_tma: the real letterbox / stretch resize with staged source rows.  A CTA takes (image, band of 8 output
// rows): the <= 16 source rows the band taps are pulled into shared memory by 1-D TMA bulk copies (cp.async.bulk +
// mbarrier -> SASS UBLKCP; plain loads when base or pitch is not 16-byte aligned), the x coefficient table of the plan
// (cv2's float64 coordinate maths, evaluated once per CTA instead of twice per output pixel) sits beside them, and a
// thread produces 8 adjacent output pixels of one row -- exact cv2 INTER_LINEAR integer model, border, LUT -- stored
// as one 128-bit (bf16) or two 128-bit (f32) streaming stores per plane.
constexpr int K1R_MAX_THREADS = 512;
constexpr int K1R_RB = 4;                    // output rows per band (2 source rows each); two bands in flight per CTA

template <typename OutT>
__global__ void __launch_bounds__(K1R_MAX_THREADS)
k1_resize_tma(const uint8_t* __restrict__ src, OutT* __restrict__ dst, const float* __restrict__ lut, int n,
              long long src_img_stride, int src_row_stride, qc_pre_plan pl, int pad_value, int swap_rb, int row_bytes,
              int tma_ok, int bands_per_img) {
    extern __shared__ __align__(128) uint8_t k1r_sm[];
    const int groups = (pl.dst_w + 7) >> 3;
    uint32_t* lut_s = reinterpret_cast<uint32_t*>(k1r_sm);                       // [3][256]
    // x table, stored [x & 7][x >> 3] so that the lanes of a warp (adjacent 8-pixel groups) read adjacent entries
    int4* xtab = reinterpret_cast<int4*>(k1r_sm + 3 * 256 * 4);                  // (3*s0, 3*s1, a0, a1) per output column
    int4* ytab = xtab + groups * 8;                                              // (s0, s1, a0, a1) per output row; s0 < 0: border
    uint8_t* rows = reinterpret_cast<uint8_t*>(ytab) + (((size_t)pl.dst_h * 16 + 127) & ~(size_t)127);   // [2][2*K1R_RB][row_bytes]
    __shared__ __align__(8) uint64_t bar[2];
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) { qc_mbar_init(&bar[0], 1); qc_mbar_init(&bar[1], 1); qc_fence_mbar_init(); }
    for (int i = tid; i < 768; i += nt) lut_s[i] = OutPack<OutT>::enc(lut[i]);
    const bool resize_x = pl.new_w != pl.src_w, resize_y = pl.new_h != pl.src_h;
    for (int x = tid; x < pl.dst_w; x += nt) {
        const int rx = x - pl.left;
        int4 e = make_int4(-1, -1, 0, 0);                                         // border column
        if (rx >= 0 && rx < pl.new_w) {
            if (resize_x) { const LinCoef c = lin_coef(rx, pl.new_w, pl.src_w, true); e = make_int4(3 * c.s0, 3 * c.s1, c.a0, c.a1); }
            else e = make_int4(3 * rx, 3 * rx, 2048, 0);
        }
        xtab[(x & 7) * groups + (x >> 3)] = e;
    }
    for (int y = tid; y < pl.dst_h; y += nt) {
        const int ry = y - pl.top;
        int4 e = make_int4(-1, -1, 2048, 0);                                      // border row
        if (ry >= 0 && ry < pl.new_h) {
            if (resize_y) { const LinCoef c = lin_coef(ry, pl.new_h, pl.src_h, false); e = make_int4(c.s0, c.s1, c.a0, c.a1); }
            else e = make_int4(ry, ry, 2048, 0);
        }
        ytab[y] = e;
    }
    const long long plane = (long long)pl.dst_h * pl.dst_w;
    const long long total = (long long)n * bands_per_img;
    const size_t buf_bytes = (size_t)2 * K1R_RB * row_bytes;
    // the copies of a band's source rows (TMA: one thread issues them; plain loads happen at use)
    auto prefetch = [&](long long item, int buf) {
        if (!tma_ok || tid != 0) return;
        const int img = (int)(item / bands_per_img), band = (int)(item - (long long)img * bands_per_img);
        const int y0 = band * K1R_RB, nr = min(K1R_RB, pl.dst_h - y0);
        const uint8_t* simg = src + (long long)img * src_img_stride;
        uint8_t* rb = rows + (size_t)buf * buf_bytes;
        int nvalid = 0;
        for (int k = 0; k < nr; ++k) nvalid += ytab[y0 + k].x >= 0 ? 2 : 0;
        if (!nvalid) return;
        qc_mbar_expect_tx(&bar[buf], (uint32_t)nvalid * (uint32_t)row_bytes);
        for (int k = 0; k < nr; ++k) {
            const int4 e = ytab[y0 + k];
            if (e.x < 0) continue;
            qc_bulk_g2s(rb + (size_t)(2 * k) * row_bytes, simg + (long long)e.x * src_row_stride, (uint32_t)row_bytes, &bar[buf]);
            qc_bulk_g2s(rb + (size_t)(2 * k + 1) * row_bytes, simg + (long long)e.y * src_row_stride, (uint32_t)row_bytes, &bar[buf]);
        }
    };
    unsigned par0 = 0u, par1 = 0u;
    __syncthreads();
    long long item = blockIdx.x;
    if (item < total) prefetch(item, 0);
    for (int it = 0; item < total; ++it, item += gridDim.x) {
        const int buf = it & 1;
        const long long next = item + gridDim.x;
        if (next < total) prefetch(next, buf ^ 1);         // rows of the next band stream in during this band's arithmetic
        const int img = (int)(item / bands_per_img), band = (int)(item - (long long)img * bands_per_img);
        const int y0 = band * K1R_RB, y1 = min(y0 + K1R_RB, pl.dst_h);
        uint8_t* rb = rows + (size_t)buf * buf_bytes;
        bool any = false;
        for (int k = 0; k < y1 - y0; ++k) any |= ytab[y0 + k].x >= 0;
        if (tma_ok) {
            if (any) {
                if (buf) { qc_mbar_wait(&bar[1], par1); par1 ^= 1u; } else { qc_mbar_wait(&bar[0], par0); par0 ^= 1u; }
            }
        } else if (any) {
            const uint8_t* simg = src + (long long)img * src_img_stride;
            const int nb = pl.src_w * 3;
            for (int i = tid; i < (y1 - y0) * 2 * nb; i += nt) {
                const int slot = i / nb, j = i - slot * nb;
                const int4 e = ytab[y0 + (slot >> 1)];
                const int sr = (slot & 1) ? e.y : e.x;
                if (sr >= 0) rb[(size_t)slot * row_bytes + j] = __ldg(simg + (long long)sr * src_row_stride + j);
            }
            __syncthreads();
        }
        // compute: thread = (band row, 8 adjacent output columns)
        for (int w = tid; w < (y1 - y0) * groups; w += nt) {
            const int k = w / groups, gx = w - k * groups;
            const int y = y0 + k, x0 = gx * 8;
            const int4 ye = ytab[y];
            const int s0 = ye.x, b0 = ye.z, b1 = ye.w;
            const uint8_t* r0 = rb + (size_t)(2 * k) * row_bytes;
            const uint8_t* r1 = r0 + row_bytes;
            uint32_t vals[3][8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int x = x0 + j;
                int px[3] = {pad_value, pad_value, pad_value};
                if (x < pl.dst_w && s0 >= 0) {
                    const int4 e = xtab[j * groups + gx];
                    if (e.x >= 0) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const int h0 = (int)r0[e.x + c] * e.z + (int)r0[e.y + c] * e.w;
                            const int h1 = (int)r1[e.x + c] * e.z + (int)r1[e.y + c] * e.w;
                            px[c] = (resize_x || resize_y) ? ((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2)
                                                           : (int)r0[e.x + c];
                        }
                    }
                }
#pragma unroll
                for (int oc = 0; oc < 3; ++oc) vals[oc][j] = lut_s[oc * 256 + (px[swap_rb ? 2 - oc : oc] & 255)];
            }
#pragma unroll
            for (int oc = 0; oc < 3; ++oc) {
                OutT* o = dst + ((long long)img * 3 + oc) * plane + (long long)y * pl.dst_w + x0;
                const bool full = (x0 + 8 <= pl.dst_w) && ((reinterpret_cast<uintptr_t>(o) & 15) == 0);
                if (full) {
                    OutPack<OutT>::store8(o, vals[oc]);
                } else {
                    for (int j = 0; j < 8 && x0 + j < pl.dst_w; ++j) {
                        if (sizeof(OutT) == 2) reinterpret_cast<uint16_t*>(o)[j] = (uint16_t)vals[oc][j];
                        else reinterpret_cast<uint32_t*>(o)[j] = vals[oc][j];
                    }
                }
            }
        }
        __syncthreads();                                   // `buf` may be refilled by the next iteration's prefetch
    }
}

inline long long py_round(double v) { return (long long)nearbyint(v); }  // round-half-even == Python round()

}  // namespace

extern "C" int qc_pre_make_plan(int mode, int src_h, int src_w, int H, int W, int stride, qc_pre_plan* pl) {
    if (!pl || src_h <= 0 || src_w <= 0 || H <= 0 || W <= 0) return QC_ERR_INVALID;
    memset(pl, 0, sizeof(*pl));
    pl->src_h = src_h; pl->src_w = src_w;
    if (mode == QC_PRE_STRETCH) {            // inference/detector.py:240-242
        pl->new_h = pl->dst_h = H; pl->new_w = pl->dst_w = W;
        pl->gain = 1.f; pl->pad_x = pl->pad_y = 0;
        return QC_OK;
    }
    if (mode != QC_PRE_LETTERBOX && mode != QC_PRE_LETTERBOX_AUTO) return QC_ERR_INVALID;
    // ultralytics LetterBox (scaleup=True, center=True)
    double r = fmin((double)H / src_h, (double)W / src_w);
    int new_w = (int)py_round(src_w * r), new_h = (int)py_round(src_h * r);
    double dw = W - new_w, dh = H - new_h;
    if (mode == QC_PRE_LETTERBOX_AUTO) {
        if (stride <= 0) return QC_ERR_INVALID;
        dw = (double)(((int)dw % stride + stride) % stride);
        dh = (double)(((int)dh % stride + stride) % stride);
    }
    dw /= 2; dh /= 2;
    int top = (int)py_round(dh - 0.1), bottom = (int)py_round(dh + 0.1);
    int left = (int)py_round(dw - 0.1), right = (int)py_round(dw + 0.1);
    pl->new_h = new_h; pl->new_w = new_w; pl->top = top; pl->left = left;
    pl->dst_h = new_h + top + bottom; pl->dst_w = new_w + left + right;
    // ultralytics scale_boxes(img1_shape = network input, img0_shape = original)
    double gain = fmin((double)pl->dst_h / src_h, (double)pl->dst_w / src_w);
    pl->gain = (float)gain;
    pl->pad_x = (int)py_round((pl->dst_w - src_w * gain) / 2 - 0.1);
    pl->pad_y = (int)py_round((pl->dst_h - src_h * gain) / 2 - 0.1);
    return QC_OK;
}

template <typename OutT>
static int launch_k1(qc_ctx* ctx, const uint8_t* src, int n, int64_t src_img_stride, int32_t src_row_stride,
                     const qc_pre_plan* pl, const float* lut, bool lut_uniform, int pad_value, int swap_rb,
                     OutT* dst, cudaStream_t st) {
    const long long P = (long long)pl->src_h * pl->src_w;
    const bool identity = pl->new_h == pl->src_h && pl->new_w == pl->src_w && pl->top == 0 && pl->left == 0 &&
                          pl->dst_h == pl->src_h && pl->dst_w == pl->src_w;
    const bool packed = src_row_stride == pl->src_w * 3 && (src_img_stride % 16) == 0 &&
                        (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (P % 16) == 0;
    if (identity && packed && lut_uniform) {
        const int chunks = (int)((P + K1_CHUNK_PX - 1) / K1_CHUNK_PX);
        const long long total = (long long)n * chunks;
        const size_t smem = (size_t)K1_LUT_WORDS * 4 + 2 * K1_CHUNK_BYTES;
        auto kern = swap_rb ? k1_pack_stream<OutT, 1, true> : k1_pack_stream<OutT, 1, false>;
        int occ = 1;
        if (int rc = qc_prepare(ctx, kern, (int)smem, K1_THREADS, &occ)) return rc;
        long long grid = (long long)ctx->sm_count * occ;
        if (grid > total) grid = total;
        kern<<<(unsigned)grid, K1_THREADS, smem, st>>>(src, dst, lut, P, src_img_stride, chunks, total);
        QC_LAUNCHED_AS(ctx, "k1_pack_stream", st);
        return QC_OK;
    }
    if (identity && packed) {   // per-channel LUT (mean/std normalisation): three lane-private tables
        const int chunks = (int)((P + K1_CHUNK_PX - 1) / K1_CHUNK_PX);
        const long long total = (long long)n * chunks;
        const size_t smem = (size_t)3 * K1_LUT_WORDS * 4 + 2 * K1_CHUNK_BYTES;
        auto kern = swap_rb ? k1_pack_stream<OutT, 3, true> : k1_pack_stream<OutT, 3, false>;
        if (int rc = qc_prepare(ctx, kern, (int)smem, K1_THREADS, nullptr)) return rc;
        long long grid = ctx->sm_count;
        if (grid > total) grid = total;
        kern<<<(unsigned)grid, K1_THREADS, smem, st>>>(src, dst, lut, P, src_img_stride, chunks, total);
        QC_LAUNCHED_AS(ctx, "k1_pack_stream", st);
        return QC_OK;
    }
    {
        // staged-row kernel: shared memory = LUT + x table + two bands of 8 source rows
        const bool tma_ok = (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (src_img_stride & 15) == 0 && (src_row_stride & 15) == 0;
        const int row_bytes = tma_ok ? ((pl->src_w * 3 + 15) & ~15) : pl->src_w * 3;
        const int groups = (pl->dst_w + 7) / 8;
        const size_t smem = 3 * 256 * 4 + (size_t)groups * 8 * 16 + (((size_t)pl->dst_h * 16 + 127) & ~(size_t)127) +
                            (size_t)2 * 2 * K1R_RB * row_bytes + 16;
        if ((long long)smem + 2048 <= ctx->max_smem_optin && (!tma_ok || row_bytes <= src_row_stride)) {
            const int bands = (pl->dst_h + K1R_RB - 1) / K1R_RB;
            // threads: one pass over the band's (row, 8-pixel group) items when they fit, else equal passes
            const int items_band = K1R_RB * groups;
            const int passes = (items_band + 383) / 384;
            const int threads = std::min(K1R_MAX_THREADS, std::max(128, (((items_band + passes - 1) / passes) + 31) & ~31));
            int occ = 1;
            if (int rc = qc_prepare(ctx, k1_resize_tma<OutT>, (int)smem, threads, &occ)) return rc;
            const long long total = (long long)n * bands;
            long long g = (long long)ctx->sm_count * std::max(occ, 1);
            if (g > total) g = total;
            k1_resize_tma<OutT><<<(unsigned)g, threads, smem, st>>>(src, dst, lut, n, src_img_stride, src_row_stride, *pl,
                                                                   pad_value, swap_rb, row_bytes, tma_ok ? 1 : 0, bands);
            QC_LAUNCHED_AS(ctx, "k1_resize_tma", st);
            return QC_OK;
        }
    }
    const long long items = (long long)n * pl->dst_h * ((pl->dst_w + 7) / 8);
    long long grid = (items + 255) / 256;
    const long long cap = (long long)ctx->sm_count * 16;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    k1_generic<OutT><<<(unsigned)grid, 256, 0, st>>>(src, dst, lut, n, src_img_stride, src_row_stride, *pl,
                                                     pad_value, swap_rb);
    QC_LAUNCHED_AS(ctx, "k1_generic", st);
    return QC_OK;
}

extern "C" int qc_preprocess(qc_ctx* ctx, const uint8_t* src, int n, int64_t src_img_stride,
                             int32_t src_row_stride, const qc_pre_plan* plan, const float* lut, int pad_value,
                             int flags, int out_dtype, void* dst, void* stream) {
    qc_device_guard dev_guard(ctx);
    if (!ctx) return QC_ERR_INVALID;
    QC_ENTER(ctx, stream);
    if (!src || !plan || !lut || !dst || n < 0) QC_FAIL(ctx, QC_ERR_INVALID, "qc_preprocess: null/invalid argument");
    if (n == 0) return QC_OK;
    if (plan->src_h <= 0 || plan->src_w <= 0 || plan->dst_h <= 0 || plan->dst_w <= 0 || plan->new_h <= 0 ||
        plan->new_w <= 0 || plan->top < 0 || plan->left < 0 || plan->top + plan->new_h > plan->dst_h ||
        plan->left + plan->new_w > plan->dst_w || src_row_stride < plan->src_w * 3)
        QC_FAIL(ctx, QC_ERR_INVALID, "qc_preprocess: inconsistent plan");
    if (pad_value < 0 || pad_value > 255) QC_FAIL(ctx, QC_ERR_INVALID, "qc_preprocess: pad_value out of range");
    cudaStream_t st = (cudaStream_t)stream;
    const bool lut_uniform = (flags & QC_PRE_LUT_UNIFORM) != 0;
    const int swap = (flags & QC_PRE_SWAP_RB) ? 1 : 0;
    if (out_dtype == QC_DT_BF16)
        return launch_k1<__nv_bfloat16>(ctx, src, n, src_img_stride, src_row_stride, plan, lut, lut_uniform,
                                        pad_value, swap, (__nv_bfloat16*)dst, st);
    if (out_dtype == QC_DT_F32)
        return launch_k1<float>(ctx, src, n, src_img_stride, src_row_stride, plan, lut, lut_uniform, pad_value,
                                swap, (float*)dst, st);
    QC_FAIL(ctx, QC_ERR_INVALID, "qc_preprocess: bad out_dtype %d", out_dtype);
}
