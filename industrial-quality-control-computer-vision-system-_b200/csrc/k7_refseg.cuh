// k7_refseg.cuh -- K7: the reference's own per-class segmentation methods on device.
// Included by k4_segment.cu INSIDE its anonymous namespace: the kernels here share K4's region tables, band
// decomposition, bit rows and union-find stages.
//
//   QC_SEG_REF_THRESHOLD   inference/segmentation.py:197-246 + _clean_mask :449-462
//     k7_blur     skimage gaussian(sigma=1) = img_as_float + scipy correlate1d (float64, 9 taps, axis 0 then axis 1,
//                 mode='nearest'), operation order of scipy's ni_filters.c symmetric branch, no fma; min / max
//     k7_hist     np.histogram(256 bins over [min, max]) with numpy's edge-corrected bin index
//     k7_otsu     skimage threshold_otsu: float32 counts and weights, float64 means, first argmax, bin centre
//     k7_cmp      blurred < T*factor | blurred > T*factor -> bit rows
//     K4 union-find (k4_ccl_local / k4_border / k4_compress) + k7_area + k7_filter   components < min_area removed
//     k7_fill     scipy.ndimage.binary_fill_holes: 4-connected background flood from the ROI border (bit-parallel
//                 run filling by carry propagation, row sweeps inside strips, iterated to the fixpoint)
//     k4_mask     (bit-row input) closing then opening, 3x3 cross x iters (skimage disk(2) == 2 iterations)
//     k7_emit / k7_finish   mask, area, extent, the integer sums behind the confidence of :527-555
//   QC_SEG_REF_ADAPTIVE    inference/segmentation.py:248-298
//     k7_adaptive cv2.adaptiveThreshold(GAUSSIAN_C): float32 separable Gaussian with cv2's summation order (its AVX2
//                 path: fused multiply-adds, three column groups), cvRound -> uint8 mean, src - mean vs -C -> bit rows
//     then K4's morphology -> CCL -> stats -> min-area filter unchanged.
// All of it is HBM/L2-bound byte, bit and float64 work; nothing here is GEMM-shaped.
#pragma once

constexpr double K7_W0 = 0x1.9884a307594fbp-2;     // scipy.ndimage._filters._gaussian_kernel1d(1.0, 0, 4)[4..8]
constexpr double K7_W1 = 0x1.ef8eb9ad499bap-3;
constexpr double K7_W2 = 0x1.ba4b99d1799abp-5;
constexpr double K7_W3 = 0x1.22724cb7eb269p-8;
constexpr double K7_W4 = 0x1.18a9c4fd536c6p-13;
constexpr double K7_INV255 = 1.0 / 255;            // img_as_float multiplies by the rounded reciprocal

// scipy NI_Correlate1D, symmetric filter: o = x0*w0; o += (x[-4]+x[4])*w4; ... ; o += (x[-1]+x[1])*w1
__device__ __forceinline__ double k7_tap9(double c, double a4, double b4, double a3, double b3, double a2, double b2,
                                          double a1, double b1) {
    double o = __dmul_rn(c, K7_W0);
    o = __dadd_rn(o, __dmul_rn(__dadd_rn(a4, b4), K7_W4));
    o = __dadd_rn(o, __dmul_rn(__dadd_rn(a3, b3), K7_W3));
    o = __dadd_rn(o, __dmul_rn(__dadd_rn(a2, b2), K7_W2));
    o = __dadd_rn(o, __dmul_rn(__dadd_rn(a1, b1), K7_W1));
    return o;
}

struct K7Gray {
    const uint8_t* base;
    long long pitch;
};
__device__ __forceinline__ K7Gray k7_gray_of(const Img& im, const qc_seg_region& rg, const uint8_t* gray_ws) {
    K7Gray g;
    if (im.channels == 3 || gray_ws) { g.base = gray_ws + rg.out_offset; g.pitch = rg.w; }   // region-packed gray copy
    else { g.base = im.data + (long long)rg.img * im.img_stride + (long long)rg.y * im.row_stride + rg.x; g.pitch = im.row_stride; }
    return g;
}

// per-region scalars of the threshold method
struct K7Scalars {
    unsigned long long* minmax;   // [R][2] bit patterns of the (non-negative) float64 min / max
    double* thr_adj;              // [R] threshold x class factor
};

// gray input: region-packed copy of every ROI (the K7 kernels then stream contiguous, padded rows)
__global__ void __launch_bounds__(K4_THREADS)
k7_pack_gray(Img im, const qc_seg_region* __restrict__ regs, int R, int total_bands, const int* __restrict__ g_band_off,
             uint8_t* __restrict__ gray_ws) {
    __shared__ int s_off[K4_LOC_MAX + 1];
    const int* off = loc_load(s_off, g_band_off, R);
    for (int t = blockIdx.x; t < total_bands; t += gridDim.x) {
        const qc_seg_region rg = regs[loc_find(off, R, t)];
        const int BR = k4_band_rows(rg.w);
        const int y0 = (t - rg.band_offset) * BR, y1 = min(y0 + BR, rg.h);
        const uint8_t* src = im.data + (long long)rg.img * im.img_stride + (long long)rg.y * im.row_stride + rg.x;
        const int n = (y1 - y0) * rg.w;
        for (int i = threadIdx.x; i < n; i += K4_THREADS) {
            const int yy = i / rg.w, x = i - yy * rg.w;
            gray_ws[rg.out_offset + (long long)(y0 + yy) * rg.w + x] = __ldg(src + (long long)(y0 + yy) * im.row_stride + x);
        }
    }
}

__global__ void k7_init(int R, unsigned long long* __restrict__ minmax, long long* __restrict__ sums) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    if (minmax)
        for (int i = gid; i < R; i += gsz) {
            minmax[2 * i] = 0x7FF0000000000000ull;       // +inf
            minmax[2 * i + 1] = 0ull;                    // 0.0 (the blurred image is non-negative)
        }
    if (sums)
        for (int i = gid; i < R * 4; i += gsz) sums[i] = 0;
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(K4_THREADS)
k7_blur(Img im, const qc_seg_region* __restrict__ regs, int R, int total_bands, const int* __restrict__ g_band_off,
        const uint8_t* __restrict__ gray_ws, double* __restrict__ blurred, unsigned long long* __restrict__ minmax) {
    __shared__ double s_src[40][41];
    __shared__ double s_v[32][41];
    __shared__ int s_off[K4_LOC_MAX + 1];
    __shared__ double s_mn[K4_THREADS / 32], s_mx[K4_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int* off = loc_load(s_off, g_band_off, R);
    for (int t = blockIdx.x; t < total_bands; t += gridDim.x) {
        const int r = loc_find(off, R, t);
        const qc_seg_region rg = regs[r];
        const int BR = k4_band_rows(rg.w);
        const int y0 = (t - rg.band_offset) * BR, y1 = min(y0 + BR, rg.h);
        const K7Gray g = k7_gray_of(im, rg, gray_ws);
        double mn = __longlong_as_double(0x7FF0000000000000ll), mx = 0.0;
        for (int ty = y0; ty < y1; ty += 32) {
            for (int tx = 0; tx < rg.w; tx += 32) {
                __syncthreads();
                // source tile with 'nearest' borders: 4 columns per thread-step -- inside the row one (unaligned)
                // 32-bit read as two aligned words + a funnel shift, single clamped bytes only at the ROI's edges
                for (int i = tid; i < 40 * 10; i += K4_THREADS) {
                    const int rr = i / 10, cw = i - rr * 10;
                    const int yy = min(max(ty - 4 + rr, 0), rg.h - 1), x0 = tx - 4 + 4 * cw;
                    const uint8_t* rowp = g.base + (long long)yy * g.pitch;
                    unsigned wv;
                    if (x0 >= 0 && x0 + 3 < rg.w) {
                        const uintptr_t a = reinterpret_cast<uintptr_t>(rowp + x0);
                        const unsigned* q = reinterpret_cast<const unsigned*>(a & ~(uintptr_t)3);
                        wv = __funnelshift_r(__ldg(q), __ldg(q + 1), (unsigned)(a & 3) * 8u);    // q+1: inside the padded workspace
                    } else {
                        wv = 0u;
#pragma unroll
                        for (int j = 0; j < 4; ++j) wv |= (unsigned)rowp[min(max(x0 + j, 0), rg.w - 1)] << (8 * j);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        s_src[rr][4 * cw + j] = __dmul_rn((double)((wv >> (8 * j)) & 255u), K7_INV255);
                }
                __syncthreads();
                // Register sliding windows: a thread produces 4 adjacent outputs along the filter axis from 12 staged
                // values (3 loads per output instead of 9; the kernel is bound by 64-bit shared-memory loads).
                for (int i = tid; i < 8 * 40; i += K4_THREADS) {           // axis 0 first: lanes = columns
                    const int rg4 = i / 40, cc = i - rg4 * 40;
                    double v[12];
#pragma unroll
                    for (int k = 0; k < 12; ++k) v[k] = s_src[4 * rg4 + k][cc];
#pragma unroll
                    for (int o = 0; o < 4; ++o)
                        s_v[4 * rg4 + o][cc] = k7_tap9(v[o + 4], v[o], v[o + 8], v[o + 1], v[o + 7], v[o + 2], v[o + 6], v[o + 3], v[o + 5]);
                }
                __syncthreads();
                for (int i = tid; i < 32 * 8; i += K4_THREADS) {           // then axis 1: lanes = rows (pitch 41: no bank conflicts)
                    const int rr = i & 31, cg = i >> 5;
                    const int y = ty + rr, x = tx + 4 * cg;
                    if (y >= y1 || x >= rg.w) continue;
                    double v[12];
#pragma unroll
                    for (int k = 0; k < 12; ++k) v[k] = s_v[rr][4 * cg + k];
                    double* q = blurred + rg.out_offset + (long long)y * rg.w + x;
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        if (x + o < rg.w) {
                            const double r9 = k7_tap9(v[o + 4], v[o], v[o + 8], v[o + 1], v[o + 7], v[o + 2], v[o + 6], v[o + 3], v[o + 5]);
                            q[o] = r9;
                            mn = fmin(mn, r9);
                            mx = fmax(mx, r9);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        __syncthreads();
        if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
        __syncthreads();
        if (tid == 0) {
            for (int k = 1; k < K4_THREADS / 32; ++k) { mn = fmin(mn, s_mn[k]); mx = fmax(mx, s_mx[k]); }
            atomicMin(&minmax[2 * r], (unsigned long long)__double_as_longlong(mn));
            atomicMax(&minmax[2 * r + 1], (unsigned long long)__double_as_longlong(mx));
        }
    }
}

// np.linspace(first, last, 257)[i]
__device__ __forceinline__ double k7_edge(int i, double first, double last, double step) {
    return i == 256 ? last : __dadd_rn(__dmul_rn((double)i, step), first);
}

__global__ void __launch_bounds__(K4_THREADS)
k7_hist(const qc_seg_region* __restrict__ regs, int R, int total_bands, const int* __restrict__ g_band_off,
        const double* __restrict__ blurred, const unsigned long long* __restrict__ minmax,
        unsigned* __restrict__ hist) {
    __shared__ unsigned s_h[256];
    __shared__ double s_e[257];
    __shared__ int s_off[K4_LOC_MAX + 1];
    const int tid = threadIdx.x;
    const int* off = loc_load(s_off, g_band_off, R);
    int cur_r = -1;
    for (int t = blockIdx.x; t < total_bands; t += gridDim.x) {
        const int r = loc_find(off, R, t);
        const qc_seg_region rg = regs[r];
        const double first = __longlong_as_double((long long)minmax[2 * r]);
        const double last = __longlong_as_double((long long)minmax[2 * r + 1]);
        if (first == last) continue;                       // constant image: threshold_otsu returns the value
        const int BR = k4_band_rows(rg.w);
        const int y0 = (t - rg.band_offset) * BR, y1 = min(y0 + BR, rg.h);
        __syncthreads();
        s_h[tid] = 0;
        if (r != cur_r) {                                  // np.linspace(first, last, 257)
            const double step = __ddiv_rn(__dsub_rn(last, first), 256.0);
            for (int i = tid; i < 257; i += K4_THREADS) s_e[i] = k7_edge(i, first, last, step);
            cur_r = r;
        }
        __syncthreads();
        // numpy's bin of x is fixed by the comparisons against the edges (its multiplicative estimate is only ever
        // corrected by one step down / up); any estimate within a bin of the truth followed by the same corrective
        // comparisons lands in the same bin
        const double scale = 256.0 / (last - first);
        const double* p = blurred + rg.out_offset + (long long)y0 * rg.w;
        const int n = (y1 - y0) * rg.w;
        const int lane = tid & 31;
        for (int i0 = 0; i0 < n; i0 += 4 * K4_THREADS) {
            double xv[4];                                  // four loads in flight per thread: the pass is latency-bound
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * K4_THREADS + tid;
                xv[u] = i < n ? p[i] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i0 + u * K4_THREADS >= n) break;       // block-uniform
                const int i = i0 + u * K4_THREADS + tid;
                int idx = -1;
                if (i < n) {
                    const double x = xv[u];
                    idx = min(max((int)((x - first) * scale), 0), 255);
                    while (idx > 0 && x < s_e[idx]) idx -= 1;
                    while (idx < 255 && x >= s_e[idx + 1]) idx += 1;
                }
                // neighbouring pixels of a blurred image mostly share a bin: one shared-memory atomic per run of
                // equal bins inside the warp instead of one per pixel
                const int prev = __shfl_up_sync(0xffffffffu, idx, 1);
                const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || idx != prev);
                if (idx >= 0 && ((heads >> lane) & 1u)) {
                    const unsigned rest = lane == 31 ? 0u : (heads >> (lane + 1));
                    const int len = rest ? __ffs(rest) : 32 - lane;
                    atomicAdd(&s_h[idx], (unsigned)len);
                }
            }
        }
        __syncthreads();
        if (s_h[tid]) atomicAdd(&hist[(long long)r * 256 + tid], s_h[tid]);
    }
}

constexpr int K7_OTSU_WARPS = 4;
__global__ void __launch_bounds__(K7_OTSU_WARPS * 32)
k7_otsu(const qc_seg_region* __restrict__ regs, int R, const unsigned* __restrict__ hist,
        const unsigned long long* __restrict__ minmax, double* __restrict__ thr_adj, int* __restrict__ out_thr) {
    __shared__ float s_c32[K7_OTSU_WARPS][256], s_w1[K7_OTSU_WARPS][256], s_w2[K7_OTSU_WARPS][256];
    __shared__ double s_ctr[K7_OTSU_WARPS][256], s_s1[K7_OTSU_WARPS][256], s_s2[K7_OTSU_WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = blockIdx.x * K7_OTSU_WARPS + warp; r < R; r += gridDim.x * K7_OTSU_WARPS) {
        const double first = __longlong_as_double((long long)minmax[2 * r]);
        const double last = __longlong_as_double((long long)minmax[2 * r + 1]);
        double thr = first;
        int bidx = 0;
        __syncwarp();
        if (first != last) {
            const double step = __ddiv_rn(__dsub_rn(last, first), 256.0);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = lane * 8 + k;
                s_c32[warp][i] = __uint2float_rn(hist[(long long)r * 256 + i]);       // counts.astype('float32')
                s_ctr[warp][i] = __ddiv_rn(__dadd_rn(k7_edge(i, first, last, step), k7_edge(i + 1, first, last, step)), 2.0);
            }
            __syncwarp();
            if (lane == 0) {            // np.cumsum: sequential, float32 weights / float64 weighted sums
                float a32 = 0.0f;
                double a64 = 0.0;
                for (int i = 0; i < 256; ++i) {
                    a32 = __fadd_rn(a32, s_c32[warp][i]);
                    a64 = __dadd_rn(a64, __dmul_rn((double)s_c32[warp][i], s_ctr[warp][i]));
                    s_w1[warp][i] = a32;
                    s_s1[warp][i] = a64;
                }
            } else if (lane == 1) {
                float a32 = 0.0f;
                double a64 = 0.0;
                for (int i = 255; i >= 0; --i) {
                    a32 = __fadd_rn(a32, s_c32[warp][i]);
                    a64 = __dadd_rn(a64, __dmul_rn((double)s_c32[warp][i], s_ctr[warp][i]));
                    s_w2[warp][i] = a32;
                    s_s2[warp][i] = a64;
                }
            }
            __syncwarp();
            double best = -1.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = lane * 8 + k;
                if (i >= 255) continue;
                const double m1 = __ddiv_rn(s_s1[warp][i], (double)s_w1[warp][i]);
                const double m2 = __ddiv_rn(s_s2[warp][i + 1], (double)s_w2[warp][i + 1]);
                const double d = __dsub_rn(m1, m2);
                const double v = __dmul_rn((double)__fmul_rn(s_w1[warp][i], s_w2[warp][i + 1]), __dmul_rn(d, d));
                if (v > best) { best = v; bidx = i; }                 // first maximum inside the lane
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
                if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
            }
            thr = s_ctr[warp][bidx];
        }
        if (lane == 0) {
            thr_adj[r] = __dmul_rn(thr, regs[r].thr_scale);
            out_thr[r] = bidx;
        }
    }
}

__global__ void __launch_bounds__(K4_THREADS)
k7_cmp(const qc_seg_region* __restrict__ regs, int R, int total_bands, const int* __restrict__ g_band_off,
       const double* __restrict__ blurred, const double* __restrict__ thr_adj, unsigned* __restrict__ bits) {
    __shared__ int s_off[K4_LOC_MAX + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int* off = loc_load(s_off, g_band_off, R);
    for (int t = blockIdx.x; t < total_bands; t += gridDim.x) {
        const int r = loc_find(off, R, t);
        const qc_seg_region rg = regs[r];
        const int wpr = wpr_of(rg);
        const int BR = k4_band_rows(rg.w);
        const int y0 = (t - rg.band_offset) * BR, y1 = min(y0 + BR, rg.h);
        const double T = thr_adj[r];
        const int items = (y1 - y0) * wpr;
        // a warp turns four 32-pixel words per step (four independent loads in flight per lane)
        for (int it0 = warp * 4; it0 < items; it0 += (K4_THREADS / 32) * 4) {
            double v[4];
            bool ok[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int it = it0 + j;
                const int yy = it / wpr, wx = it - yy * wpr;
                const int x = wx * 32 + lane;
                ok[j] = it < items && x < rg.w;
                v[j] = ok[j] ? blurred[rg.out_offset + (long long)(y0 + yy) * rg.w + x] : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int it = it0 + j;
                if (it >= items) break;                    // warp-uniform
                const bool fg = ok[j] && (rg.polarity ? (v[j] > T) : (v[j] < T));
                const unsigned word = __ballot_sync(0xffffffffu, fg);
                const int yy = it / wpr, wx = it - yy * wpr;
                if (lane == 0) bits[rg.bits_offset + (long long)(y0 + yy) * wpr + wx] = word;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// cv2.adaptiveThreshold, ADAPTIVE_THRESH_GAUSSIAN_C.  Tile = 32 x 32 outputs; the tile's source bytes (replicate
// border materialised along x), the row pass of every contributing row, then the column pass.
__host__ __device__ inline int k7_block_size(int w, int h) {
    int b = (w < h ? w : h) / 10;
    if (b < 11) b = 11;
    return b | 1;
}

__global__ void k7_block_out(const qc_seg_region* __restrict__ regs, int R, int* __restrict__ out_thr,
                             long long* __restrict__ sums) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < R; r += gridDim.x * blockDim.x) {
        out_thr[r] = k7_block_size(regs[r].w, regs[r].h);          // the adaptive method reports its block size
        if (sums) { sums[4 * r] = sums[4 * r + 1] = sums[4 * r + 2] = 0; sums[4 * r + 3] = (long long)regs[r].w * regs[r].h; }
    }
}

__global__ void __launch_bounds__(K4_THREADS)
k7_adaptive(Img im, const qc_seg_region* __restrict__ regs, int R, int total_bands,
            const int* __restrict__ g_band_off, const uint8_t* __restrict__ gray_ws, const float* __restrict__ akern,
            int akern_pitch, int adaptive_c, unsigned* __restrict__ bits, int pmax, int p_skip) {
    extern __shared__ __align__(16) unsigned char k7sm[];
    // layout: float kern[2*pmax+1 rounded up to 4] | float rp[(32+2*pmax)][33] | uint8 src[(32+2*pmax)][(32+2*pmax)]
    const int kmax = (2 * pmax + 1 + 3) & ~3;
    float* s_k = reinterpret_cast<float*>(k7sm);
    float* s_rp = s_k + kmax;
    uint8_t* s_src = reinterpret_cast<uint8_t*>(s_rp + (32 + 2 * pmax) * 33);
    __shared__ int s_off[K4_LOC_MAX + 1];
    const int tid = threadIdx.x, lane = tid & 31;
    const int* off = loc_load(s_off, g_band_off, R);
    for (int t = blockIdx.x; t < total_bands; t += gridDim.x) {
        const int r = loc_find(off, R, t);
        const qc_seg_region rg = regs[r];
        const int wpr = wpr_of(rg);
        const int BR = k4_band_rows(rg.w);
        const int y0 = (t - rg.band_offset) * BR, y1 = min(y0 + BR, rg.h);
        const K7Gray g = k7_gray_of(im, rg, gray_ws);
        const int bs = k7_block_size(rg.w, rg.h), p = bs >> 1;
        if (p <= p_skip) continue;                      // small blocks: k7_adaptive_small
        const int sw = 32 + 2 * p;
        const int c0 = rg.w & ~7;                       // columns below: both passes fused (cv2's 8-lane vectors)
        const int c1 = c0 + (((rg.w - c0) >> 2) << 2);  // [c0, c1): row pass fused, column pass unfused; then unfused
        __syncthreads();
        for (int i = tid; i < bs; i += K4_THREADS) s_k[i] = akern[(long long)((bs - 11) >> 1) * akern_pitch + i];
        for (int ty = y0; ty < y1; ty += 32) {
            const int yend = min(ty + 32, y1);
            const int ry0 = max(ty - p, 0), ry1 = min(yend + p, rg.h), nr = ry1 - ry0;
            for (int tx = 0; tx < rg.w; tx += 32) {
                __syncthreads();
                for (int i = tid; i < nr * sw; i += K4_THREADS) {
                    const int rr = i / sw, cc = i - rr * sw;
                    const int xx = min(max(tx - p + cc, 0), rg.w - 1);
                    s_src[rr * sw + cc] = g.base[(long long)(ry0 + rr) * g.pitch + xx];
                }
                __syncthreads();
                for (int i = tid; i < nr * 32; i += K4_THREADS) {           // row pass
                    const int rr = i >> 5, c = i & 31;
                    const int x = tx + c;
                    if (x >= rg.w) continue;
                    const uint8_t* sp = s_src + rr * sw + c;
                    float s = __fmul_rn((float)sp[0], s_k[0]);
                    if (x < c1) {
                        for (int k = 1; k < bs; ++k) s = __fmaf_rn((float)sp[k], s_k[k], s);
                    } else {
                        for (int k = 1; k < bs; ++k) s = __fadd_rn(s, __fmul_rn((float)sp[k], s_k[k]));
                    }
                    s_rp[rr * 33 + c] = s;
                }
                __syncthreads();
                for (int i = tid; i < 32 * 32; i += K4_THREADS) {            // column pass + threshold
                    const int rr = i >> 5, c = i & 31;
                    const int y = ty + rr, x = tx + c;
                    const bool valid = y < yend && x < rg.w;
                    bool fg = false;
                    if (valid) {
                        auto row = [&](int yy) -> float { return s_rp[(min(max(yy, 0), rg.h - 1) - ry0) * 33 + c]; };
                        float s = __fmul_rn(row(y), s_k[p]);
                        if (x < c0) {
                            for (int k = 1; k <= p; ++k) s = __fmaf_rn(__fadd_rn(row(y + k), row(y - k)), s_k[p + k], s);
                        } else {
                            for (int k = 1; k <= p; ++k)
                                s = __fadd_rn(s, __fmul_rn(s_k[p + k], __fadd_rn(row(y + k), row(y - k))));
                        }
                        const int mean = min(max(__float2int_rn(s), 0), 255);          // cvRound + saturate_cast<uchar>
                        const int d = (int)s_src[(y - ry0) * sw + c + p] - mean;
                        fg = rg.polarity ? (d > -adaptive_c) : (d <= -adaptive_c);
                    }
                    const unsigned word = __ballot_sync(0xffffffffu, fg);
                    if (lane == 0 && y < yend) bits[rg.bits_offset + (long long)y * wpr + (tx >> 5)] = word;
                }
            }
        }
    }
}


// Blocks up to 33 (every ROI whose short side is below 340 px): 32 x 32 tiles with the source tile held as FLOAT
// (replicate border materialised on both axes), register sliding windows in both passes -- a thread produces 4
// adjacent outputs along the filter axis, so every staged value is loaded once per 4 outputs -- and the kernel padded
// with zero taps to a multiple of 4 (fma(x, 0, s) == s, so cv2's sequential sums are reproduced bit for bit).
constexpr int K7A_P = 16;                    // largest half-block of this path
constexpr int K7A_FP = 68;                   // float pitch of the source tile (32 + 2*16 + 4)
constexpr int K7A_ROWS = 32 + 2 * K7A_P;     // source / row-pass rows of a tile
__global__ void __launch_bounds__(K4_THREADS)
k7_adaptive_small(Img im, const qc_seg_region* __restrict__ regs, int R, int total_bands,
                  const int* __restrict__ g_band_off, const uint8_t* __restrict__ gray_ws,
                  const float* __restrict__ akern, int akern_pitch, int adaptive_c, unsigned* __restrict__ bits) {
    __shared__ __align__(16) float s_f[K7A_ROWS * K7A_FP + 16];        // source tile, virtual rows ty-p .. ty+31+p
    __shared__ __align__(16) float s_rp[(K7A_ROWS + 8) * 33];          // row-pass results, 4 zeroed pad rows on each side
    __shared__ __align__(16) float s_k[40], s_kc[24];                   // taps (zero padded), centre + lower half taps
    __shared__ int s_off[K4_LOC_MAX + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int* off = loc_load(s_off, g_band_off, R);
    for (int i = tid; i < (K7A_ROWS + 8) * 33; i += K4_THREADS) s_rp[i] = 0.f;
    for (int i = tid; i < K7A_ROWS * K7A_FP + 16; i += K4_THREADS) s_f[i] = 0.f;
    float* const rp = s_rp + 4 * 33;                                     // virtual row 0
    for (int t = blockIdx.x; t < total_bands; t += gridDim.x) {
        const int r = loc_find(off, R, t);
        const qc_seg_region rg = regs[r];
        const int bs = k7_block_size(rg.w, rg.h), p = bs >> 1;
        if (p > K7A_P) continue;                                         // large blocks: k7_adaptive
        const int wpr = wpr_of(rg);
        const int BR = k4_band_rows(rg.w);
        const int y0 = (t - rg.band_offset) * BR, y1 = min(y0 + BR, rg.h);
        const K7Gray g = k7_gray_of(im, rg, gray_ws);
        const int sw = 32 + 2 * p;
        const int c0 = rg.w & ~7;                       // columns below: both passes fused (cv2's 8-lane vectors)
        const int c1 = c0 + (((rg.w - c0) >> 2) << 2);  // [c0, c1): row pass fused, column pass unfused; then unfused
        const int nk4 = (bs + 3) >> 2, np4 = (p + 3) >> 2;
        __syncthreads();
        if (tid < 40) s_k[tid] = tid < bs ? akern[(long long)((bs - 11) >> 1) * akern_pitch + tid] : 0.f;
        if (tid < 24) s_kc[tid] = tid <= p ? akern[(long long)((bs - 11) >> 1) * akern_pitch + p + tid] : 0.f;
        for (int ty = y0; ty < y1; ty += 32) {
            const int yend = min(ty + 32, y1);
            const int nvr = (yend - ty) + 2 * p;                         // virtual rows ty-p .. yend-1+p
            for (int tx = 0; tx < rg.w; tx += 32) {
                __syncthreads();
                // ---- stage: float source tile, both borders replicated; 4 columns per thread-step
                const int q4 = (sw + 3) >> 2;
                for (int i = tid; i < nvr * q4; i += K4_THREADS) {
                    const int vr = i / q4, cq = i - vr * q4;
                    const int yy = min(max(ty - p + vr, 0), rg.h - 1);
                    const uint8_t* gr = g.base + (long long)yy * g.pitch;
                    const int xb = tx - p + 4 * cq;
                    float4 v;
                    v.x = (float)gr[min(max(xb, 0), rg.w - 1)];
                    v.y = (float)gr[min(max(xb + 1, 0), rg.w - 1)];
                    v.z = (float)gr[min(max(xb + 2, 0), rg.w - 1)];
                    v.w = (float)gr[min(max(xb + 3, 0), rg.w - 1)];
                    *reinterpret_cast<float4*>(s_f + vr * K7A_FP + 4 * cq) = v;
                }
                __syncthreads();
                // ---- row pass: thread = (virtual row, 4 adjacent columns); sequential sum over the taps (cv2's order)
                for (int i = tid; i < nvr * 8; i += K4_THREADS) {
                    const int vr = i >> 3, q = i & 7;
                    const float* fp = s_f + vr * K7A_FP + 4 * q;
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                    float4 cur = *reinterpret_cast<const float4*>(fp);
                    if (tx + 4 * q < c1) {
                        for (int t4 = 0; t4 < nk4; ++t4) {
                            const float4 nx = *reinterpret_cast<const float4*>(fp + 4 * t4 + 4);
                            const float4 k = *reinterpret_cast<const float4*>(s_k + 4 * t4);
                            a0 = __fmaf_rn(cur.x, k.x, a0); a1 = __fmaf_rn(cur.y, k.x, a1); a2 = __fmaf_rn(cur.z, k.x, a2); a3 = __fmaf_rn(cur.w, k.x, a3);
                            a0 = __fmaf_rn(cur.y, k.y, a0); a1 = __fmaf_rn(cur.z, k.y, a1); a2 = __fmaf_rn(cur.w, k.y, a2); a3 = __fmaf_rn(nx.x, k.y, a3);
                            a0 = __fmaf_rn(cur.z, k.z, a0); a1 = __fmaf_rn(cur.w, k.z, a1); a2 = __fmaf_rn(nx.x, k.z, a2); a3 = __fmaf_rn(nx.y, k.z, a3);
                            a0 = __fmaf_rn(cur.w, k.w, a0); a1 = __fmaf_rn(nx.x, k.w, a1); a2 = __fmaf_rn(nx.y, k.w, a2); a3 = __fmaf_rn(nx.z, k.w, a3);
                            cur = nx;
                        }
                    } else {
                        for (int t4 = 0; t4 < nk4; ++t4) {
                            const float4 nx = *reinterpret_cast<const float4*>(fp + 4 * t4 + 4);
                            const float4 k = *reinterpret_cast<const float4*>(s_k + 4 * t4);
#define K7_MA(acc, x, kk) acc = __fadd_rn(acc, __fmul_rn(x, kk))
                            K7_MA(a0, cur.x, k.x); K7_MA(a1, cur.y, k.x); K7_MA(a2, cur.z, k.x); K7_MA(a3, cur.w, k.x);
                            K7_MA(a0, cur.y, k.y); K7_MA(a1, cur.z, k.y); K7_MA(a2, cur.w, k.y); K7_MA(a3, nx.x, k.y);
                            K7_MA(a0, cur.z, k.z); K7_MA(a1, cur.w, k.z); K7_MA(a2, nx.x, k.z); K7_MA(a3, nx.y, k.z);
                            K7_MA(a0, cur.w, k.w); K7_MA(a1, nx.x, k.w); K7_MA(a2, nx.y, k.w); K7_MA(a3, nx.z, k.w);
                            cur = nx;
                        }
                    }
                    float* o = rp + vr * 33 + 4 * q;
                    o[0] = a0; o[1] = a1; o[2] = a2; o[3] = a3;
                }
                __syncthreads();
                // ---- column pass: thread = (column, 4 adjacent rows), symmetric pairs outwards (cv2's order)
                {
                    const int c = lane, vb = 4 * warp + p;                 // virtual row of the thread's first output
                    const float* cp = rp + c;
                    float u0 = cp[vb * 33], u1 = cp[(vb + 1) * 33], u2 = cp[(vb + 2) * 33], u3 = cp[(vb + 3) * 33];
                    float d0 = u0, d1 = u1, d2 = u2;                            // (row vb+3-tb is never a down tap)
                    const float kc = s_kc[0];
                    float a0 = __fmul_rn(u0, kc), a1 = __fmul_rn(u1, kc), a2 = __fmul_rn(u2, kc), a3 = __fmul_rn(u3, kc);
                    const bool fused = tx + c < c0;
                    for (int t4 = 0; t4 < np4; ++t4) {
                        const int tb = 4 * t4;
                        const float4 k = make_float4(s_kc[tb + 1], s_kc[tb + 2], s_kc[tb + 3], s_kc[tb + 4]);
                        // taps t = tb+1 .. tb+4: the up window slides down one row per tap, the down window up
                        const float n0 = cp[(vb + 4 + tb) * 33], n1 = cp[(vb + 5 + tb) * 33], n2 = cp[(vb + 6 + tb) * 33], n3 = cp[(vb + 7 + tb) * 33];
                        const float m0 = cp[(vb - 1 - tb) * 33], m1 = cp[(vb - 2 - tb) * 33], m2 = cp[(vb - 3 - tb) * 33], m3 = cp[(vb - 4 - tb) * 33];
                        if (fused) {
#define K7_CF(acc, a, b, kk) acc = __fmaf_rn(__fadd_rn(a, b), kk, acc)
                            K7_CF(a0, u1, m0, k.x); K7_CF(a1, u2, d0, k.x); K7_CF(a2, u3, d1, k.x); K7_CF(a3, n0, d2, k.x);
                            K7_CF(a0, u2, m1, k.y); K7_CF(a1, u3, m0, k.y); K7_CF(a2, n0, d0, k.y); K7_CF(a3, n1, d1, k.y);
                            K7_CF(a0, u3, m2, k.z); K7_CF(a1, n0, m1, k.z); K7_CF(a2, n1, m0, k.z); K7_CF(a3, n2, d0, k.z);
                            K7_CF(a0, n0, m3, k.w); K7_CF(a1, n1, m2, k.w); K7_CF(a2, n2, m1, k.w); K7_CF(a3, n3, m0, k.w);
                        } else {
#define K7_CU(acc, a, b, kk) acc = __fadd_rn(acc, __fmul_rn(kk, __fadd_rn(a, b)))
                            K7_CU(a0, u1, m0, k.x); K7_CU(a1, u2, d0, k.x); K7_CU(a2, u3, d1, k.x); K7_CU(a3, n0, d2, k.x);
                            K7_CU(a0, u2, m1, k.y); K7_CU(a1, u3, m0, k.y); K7_CU(a2, n0, d0, k.y); K7_CU(a3, n1, d1, k.y);
                            K7_CU(a0, u3, m2, k.z); K7_CU(a1, n0, m1, k.z); K7_CU(a2, n1, m0, k.z); K7_CU(a3, n2, d0, k.z);
                            K7_CU(a0, n0, m3, k.w); K7_CU(a1, n1, m2, k.w); K7_CU(a2, n2, m1, k.w); K7_CU(a3, n3, m0, k.w);
                        }
                        u0 = n0; u1 = n1; u2 = n2; u3 = n3;                // rows vb+4+tb .. vb+7+tb
                        d2 = m1; d1 = m2; d0 = m3;                         // rows vb-4-tb .. vb-2-tb
                    }
                    const float acc[4] = {a0, a1, a2, a3};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int y = ty + 4 * warp + j, x = tx + c;
                        bool fg = false;
                        if (y < yend && x < rg.w) {
                            const int mean = min(max(__float2int_rn(acc[j]), 0), 255);     // cvRound + saturate_cast<uchar>
                            const int d = (int)s_f[(4 * warp + j + p) * K7A_FP + c + p] - mean;
                            fg = rg.polarity ? (d > -adaptive_c) : (d <= -adaptive_c);
                        }
                        const unsigned word = __ballot_sync(0xffffffffu, fg);
                        if (lane == 0 && y < yend) bits[rg.bits_offset + (long long)y * wpr + (tx >> 5)] = word;
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// component areas on the union-find roots (k4_compress left parent[g] = root for every foreground block and zeroed
// area[g]), then the size filter on the bit rows
__global__ void __launch_bounds__(K4_THREADS)
k7_area(const qc_seg_region* __restrict__ regs, int R, int total_bands, const int* __restrict__ g_band_off,
        const unsigned* __restrict__ bits, const int* __restrict__ parent, int* __restrict__ area) {
    __shared__ int s_off[K4_LOC_MAX + 1];
    const int tid = threadIdx.x;
    const int* off = loc_load(s_off, g_band_off, R);
    for (int t = blockIdx.x; t < total_bands; t += gridDim.x) {
        const qc_seg_region rg = regs[loc_find(off, R, t)];
        const int wpr = wpr_of(rg), bw = bw_of(rg);
        const int BR = k4_band_rows(rg.w);
        const int y0 = (t - rg.band_offset) * BR, y1 = min(y0 + BR, rg.h);
        const int by0 = y0 >> 1, by1 = (y1 + 1) >> 1;
        const unsigned* B = bits + rg.bits_offset;
        for (int ci = tid; ci < (by1 - by0) * wpr; ci += K4_THREADS) {
            const int byy = ci / wpr, wx = ci - byy * wpr, by = by0 + byy;
            const unsigned r0 = __ldg(B + (long long)(2 * by) * wpr + wx);
            const unsigned r1 = (2 * by + 1 < rg.h) ? __ldg(B + (long long)(2 * by + 1) * wpr + wx) : 0u;
            const unsigned v = r0 | r1;
            if (!v) continue;
            const int* prow = parent + rg.blk_offset + (long long)by * bw + wx * 16;
            int cur = -1, acc = 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                if (!((v >> (2 * k)) & 3u)) continue;
                const int root = prow[k];
                if (root != cur) {
                    if (cur >= 0) atomicAdd(area + cur, acc);
                    cur = root;
                    acc = 0;
                }
                acc += __popc((r0 >> (2 * k)) & 3u) + __popc((r1 >> (2 * k)) & 3u);
            }
            if (cur >= 0) atomicAdd(area + cur, acc);
        }
    }
}

__global__ void __launch_bounds__(K4_THREADS)
k7_filter(const qc_seg_region* __restrict__ regs, int R, int total_bands, const int* __restrict__ g_band_off,
          const unsigned* __restrict__ bits, const int* __restrict__ parent, const int* __restrict__ area,
          unsigned* __restrict__ bits_out) {
    __shared__ int s_off[K4_LOC_MAX + 1];
    const int tid = threadIdx.x;
    const int* off = loc_load(s_off, g_band_off, R);
    for (int t = blockIdx.x; t < total_bands; t += gridDim.x) {
        const qc_seg_region rg = regs[loc_find(off, R, t)];
        const int wpr = wpr_of(rg), bw = bw_of(rg);
        const int BR = k4_band_rows(rg.w);
        const int y0 = (t - rg.band_offset) * BR, y1 = min(y0 + BR, rg.h);
        const int by0 = y0 >> 1, by1 = (y1 + 1) >> 1;
        const unsigned* B = bits + rg.bits_offset;
        unsigned* O = bits_out + rg.bits_offset;
        for (int ci = tid; ci < (by1 - by0) * wpr; ci += K4_THREADS) {
            const int byy = ci / wpr, wx = ci - byy * wpr, by = by0 + byy;
            const bool two = 2 * by + 1 < rg.h;
            const unsigned r0 = __ldg(B + (long long)(2 * by) * wpr + wx);
            const unsigned r1 = two ? __ldg(B + (long long)(2 * by + 1) * wpr + wx) : 0u;
            const unsigned v = r0 | r1;
            unsigned keep = 0;
            if (v) {
                const int* prow = parent + rg.blk_offset + (long long)by * bw + wx * 16;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    if (!((v >> (2 * k)) & 3u)) continue;
                    if (area[prow[k]] >= rg.min_area) keep |= 3u << (2 * k);
                }
            }
            O[(long long)(2 * by) * wpr + wx] = r0 & keep;
            if (two) O[(long long)(2 * by + 1) * wpr + wx] = r1 & keep;
        }
    }
}

// ------------------------------------------------------------------------------------------
// scipy.ndimage.binary_fill_holes: the background that is 4-connected to the ROI border stays, everything else is
// set.  `out` grows from the border pixels; a row step ORs in the neighbour row and closes horizontally over whole
// runs at once (adding the seeds to the run mask lets the carry run up each run; bit reversal does the other
// direction).  Rows are swept down then up inside strips, one warp per strip, until a full iteration changes nothing
// (bits only ever get set, so reads of a neighbour strip that is being swept at the same time are at worst stale).
__device__ __forceinline__ unsigned k7_hfill(unsigned m, unsigned s) {
    const unsigned up = (((m + s) ^ m) & m) | s;
    const unsigned mr = __brev(m), sr = __brev(s);
    return up | __brev((((mr + sr) ^ mr) & mr) | sr);
}

constexpr int K7_FILL_THREADS = 512;
constexpr int K7_FILL_SMEM_WORDS = 25 * 1024;          // 100 KB per CTA, two CTAs per SM

// one region; F = free (background) bit rows, O = outside bit rows, both [h][wpr] in shared or global memory.
// A strip of rows is swept by a GROUP of G lanes (one lane per 32-pixel word; G = 4, 8, 16 or 32 >= words per row), so a
// warp sweeps 32/G strips at once and the 16 warps of the CTA cover the region with strips of very few rows: an
// iteration is only a handful of dependent row steps.
__device__ __forceinline__ void k7_fill_region(const unsigned* F, volatile unsigned* O, int h, int wpr, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NW = K7_FILL_THREADS / 32;
    const int G = wpr <= 4 ? 4 : wpr <= 8 ? 8 : wpr <= 16 ? 16 : 32;
    const int spw = 32 / G;                                // strips per warp
    const int S = max(1, (h + NW * spw - 1) / (NW * spw)); // rows per strip
    const int nch = (wpr + 31) >> 5;                       // 32-word chunks per row (> 1 only for rows wider than 1024 px)
    const int gl = lane & (G - 1);                         // lane inside its group = word inside the chunk
    const int a = (warp * spw + lane / G) * S, b = min(a + S, h);
    for (;;) {
        int changed = 0;
        for (int dir = 0; dir < 2; ++dir) {
            for (int k = 0; k < S; ++k) {
                const int y = dir == 0 ? a + k : b - 1 - k;
                const bool row_ok = a + k < b;
                const int yn = dir == 0 ? y - 1 : y + 1;               // the row the flood comes from
                for (int c = 0; c < nch; ++c) {
                    const int wx = c * 32 + gl;
                    const bool valid = row_ok && wx < wpr;
                    const unsigned f = valid ? F[y * wpr + wx] : 0u;
                    const unsigned o = valid ? O[y * wpr + wx] : 0u;
                    unsigned s = o;
                    if (valid && yn >= 0 && yn < h) s |= O[yn * wpr + wx] & f;
                    for (;;) {
                        s = k7_hfill(f, s);
                        unsigned l = __shfl_up_sync(0xffffffffu, s, 1, G) >> 31;
                        unsigned rt = __shfl_down_sync(0xffffffffu, s, 1, G) & 1u;
                        if (gl == 0) l = (valid && wx > 0) ? (O[y * wpr + wx - 1] >> 31) : 0u;
                        if (gl == G - 1) rt = (valid && wx + 1 < wpr) ? (O[y * wpr + wx + 1] & 1u) : 0u;
                        const unsigned ns = s | (f & (l | (rt << 31)));
                        const bool more = ns != s;
                        s = ns;
                        if (!__any_sync(0xffffffffu, more)) break;
                    }
                    if (valid && s != o) { O[y * wpr + wx] = s; changed = 1; }
                    __syncwarp();
                }
            }
        }
        if (!__syncthreads_or(changed)) break;
    }
}

// The flood as a union-find over horizontal runs (the normal path; the sweeps above are the fallback for a region whose
// runs do not fit the CTA's shared memory).  Sweeps need one iteration per change of direction of the background's
// paths -- hundreds on a textured ROI -- while the union-find does a fixed amount of work:
//   nodes   = maximal horizontal runs of free pixels (a run may span words), numbered 1.. in raster order through a
//             CTA-wide prefix sum of the per-word run starts; node 0 = "outside"
//   unions  = one per maximal interval of F[y] & F[y-1] (two runs overlap in exactly one interval: 4-connectivity),
//             found as the start bits of that AND; runs touching the ROI's outer ring are united with node 0
//   output  = every run whose root is not 0 is a hole: result = ~outside inside the ROI.
__device__ __forceinline__ unsigned k7_run_starts(const unsigned* F, int i, int wx) {
    const unsigned f = F[i];
    return f & ~((f << 1) | (wx > 0 ? F[i - 1] >> 31 : 0u));
}

__global__ void __launch_bounds__(K7_FILL_THREADS, 2)
k7_fill(const qc_seg_region* __restrict__ regs, int R, const unsigned* __restrict__ bits_in,
        unsigned* __restrict__ bits_out, unsigned* __restrict__ free_ws, int* __restrict__ queue) {
    extern __shared__ __align__(16) unsigned k7f_sm[];
    __shared__ int s_warp[K7_FILL_THREADS / 32];
    __shared__ int s_total, s_next;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // regions differ 50x in size and a CTA only gets a handful of them: a shared counter hands them out one at a time
    // (a static round-robin left the slowest CTA with more than twice the mean work)
    for (;;) {
        __syncthreads();
        if (tid == 0) s_next = atomicAdd(queue, 1);
        __syncthreads();
        const int r = s_next;
        if (r >= R) break;
        const qc_seg_region rg = regs[r];
        const int wpr = wpr_of(rg), h = rg.h, nwords = h * wpr;
        const unsigned tail = (rg.w & 31) ? ((1u << (rg.w & 31)) - 1u) : 0xFFFFFFFFu;   // valid bits of the last word
        const unsigned* B = bits_in + rg.bits_offset;
        unsigned* G = bits_out + rg.bits_offset;
        const bool f_smem = 2 * nwords <= K7_FILL_SMEM_WORDS;
        unsigned* F = f_smem ? k7f_sm : free_ws + rg.bits_offset;
        __syncthreads();
        for (int i = tid; i < nwords; i += K7_FILL_THREADS) {
            const int wx = i % wpr;
            unsigned f = ~__ldg(B + i);
            if (wx == wpr - 1) f &= tail;
            F[i] = f;
        }
        __syncthreads();
        // ---- run numbering: exclusive prefix of the run starts over the words in raster order -----------------
        bool uf = f_smem;
        int* base = reinterpret_cast<int*>(k7f_sm + nwords);
        int* parent = base + nwords;
        const int cap = K7_FILL_SMEM_WORDS - 2 * nwords;                  // nodes the rest of the buffer holds
        if (uf) {
            const int per = (nwords + K7_FILL_THREADS - 1) / K7_FILL_THREADS;
            const int i0 = min(tid * per, nwords), i1 = min(i0 + per, nwords);
            int mine = 0;
            for (int i = i0; i < i1; ++i) mine += __popc(k7_run_starts(F, i, i % wpr));
            int inc = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += v;
            }
            if (lane == 31) s_warp[warp] = inc;
            __syncthreads();
            if (warp == 0) {
                int v = lane < K7_FILL_THREADS / 32 ? s_warp[lane] : 0, w = v;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int u = __shfl_up_sync(0xffffffffu, w, d);
                    if (lane >= d) w += u;
                }
                if (lane < K7_FILL_THREADS / 32) s_warp[lane] = w - v;
                if (lane == 31) s_total = w;
            }
            __syncthreads();
            int run = s_warp[warp] + inc - mine;
            for (int i = i0; i < i1; ++i) {
                base[i] = run;
                run += __popc(k7_run_starts(F, i, i % wpr));
            }
            uf = s_total + 1 <= cap;                                       // block-uniform
        }
        if (uf) {
            const int nodes = s_total + 1;
            for (int i = tid; i < nodes; i += K7_FILL_THREADS) parent[i] = i;
            __syncthreads();
            for (int i = tid; i < nwords; i += K7_FILL_THREADS) {
                const int y = i / wpr, wx = i - y * wpr;
                const unsigned f = F[i];
                if (!f) continue;
                const unsigned st = k7_run_starts(F, i, wx);
                const int b0 = base[i];
                // runs on the ROI's outer ring are outside
                unsigned ring = (y == 0 || y == h - 1) ? (st | (f & 1u)) : 0u;      // every run of the row (f&1: one continuing into this word)
                if (wx == 0) ring |= f & 1u;
                if (wx == wpr - 1) ring |= f & (1u << ((rg.w - 1) & 31));
                while (ring) {
                    const int c = __ffs(ring) - 1;
                    ring &= ring - 1;
                    suf_union(parent, b0 + __popc(st & ((2u << c) - 1u)), 0);
                }
                if (y == 0) continue;
                const unsigned fa = F[i - wpr];
                const unsigned both = f & fa;
                if (!both) continue;
                const unsigned prev = wx > 0 ? ((F[i - 1] & F[i - 1 - wpr]) >> 31) : 0u;
                unsigned sb = both & ~((both << 1) | prev);
                const unsigned sta = k7_run_starts(F, i - wpr, wx);
                const int ba = base[i - wpr];
                while (sb) {
                    const int c = __ffs(sb) - 1;
                    sb &= sb - 1;
                    const unsigned le = (2u << c) - 1u;
                    suf_union(parent, b0 + __popc(st & le), ba + __popc(sta & le));
                }
            }
            __syncthreads();
            for (int i = tid; i < nwords; i += K7_FILL_THREADS) {
                const int wx = i % wpr;
                const unsigned f = F[i];
                const unsigned st = k7_run_starts(F, i, wx);
                const int b0 = base[i];
                unsigned heads = st | (f & 1u);                           // first bit of every run piece inside this word
                unsigned out = 0;
                while (heads) {
                    const int c = __ffs(heads) - 1;
                    heads &= heads - 1;
                    if (suf_find(parent, b0 + __popc(st & ((2u << c) - 1u))) == 0) {
                        const unsigned t = f >> c;
                        out |= (t & ~(t + 1u)) << c;                      // the run's pixels in this word
                    }
                }
                G[i] = wx == wpr - 1 ? (~out & tail) : ~out;
            }
            continue;
        }
        // ---- fallback: sweeps, outside rows in the rest of the shared buffer or in the output itself ------------
        unsigned* O = f_smem ? k7f_sm + nwords : G;
        __syncthreads();
        for (int i = tid; i < nwords; i += K7_FILL_THREADS) {
            const int y = i / wpr, wx = i - y * wpr;
            unsigned ring = 0;
            if (y == 0 || y == h - 1) ring = 0xFFFFFFFFu;
            if (wx == 0) ring |= 1u;
            if (wx == wpr - 1) ring |= 1u << ((rg.w - 1) & 31);
            O[i] = F[i] & ring;
        }
        __syncthreads();
        k7_fill_region(F, O, h, wpr, tid);
        for (int i = tid; i < nwords; i += K7_FILL_THREADS) {
            const int wx = i % wpr;
            const unsigned o = O[i];
            G[i] = wx == wpr - 1 ? (~o & tail) : ~o;
        }
    }
}

// ------------------------------------------------------------------------------------------
// final pass of the threshold method: mask bytes / 0-1 labels, area, extent, confidence sums.  A thread takes 16
// pixels: their mask bits, their gray bytes (two aligned 128-bit loads of the region-packed copy) and three byte dot
// products (dp4a) for sum(gray), sum(gray^2) and sum(gray over the mask).
__global__ void __launch_bounds__(K4_THREADS)
k7_emit(Img im, const qc_seg_region* __restrict__ regs, int R, int total_bands, const int* __restrict__ g_band_off,
        const uint8_t* __restrict__ gray_ws, const unsigned* __restrict__ bits, uint8_t* __restrict__ out_mask,
        int* __restrict__ out_labels, int* __restrict__ stats, long long* __restrict__ out_area,
        long long* __restrict__ sums) {
    __shared__ int s_off[K4_LOC_MAX + 1];
    __shared__ long long s_acc[K4_THREADS / 32][4];
    __shared__ int s_ext[K4_THREADS / 32][4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int* off = loc_load(s_off, g_band_off, R);
    for (int t = blockIdx.x; t < total_bands; t += gridDim.x) {
        const int r = loc_find(off, R, t);
        const qc_seg_region rg = regs[r];
        const int wpr = wpr_of(rg);
        const int BR = k4_band_rows(rg.w);
        const int y0 = (t - rg.band_offset) * BR, y1 = min(y0 + BR, rg.h);
        const uint8_t* gbase = gray_ws + rg.out_offset;            // region-packed gray, pitch = w
        const unsigned* B = bits + rg.bits_offset;
        unsigned a_fg = 0, u_fg = 0, u_all = 0, u_sq = 0;
        int mnx = INT_MAX, mny = INT_MAX, mxx = -1, mxy = -1;
        const int segs = (rg.w + 15) >> 4;
        const int items = (y1 - y0) * segs;
        for (int it = tid; it < items; it += K4_THREADS) {
            const int yy = it / segs, sg = it - yy * segs, y = y0 + yy, x0 = sg * 16;
            const int nv = min(16, rg.w - x0);
            unsigned b16 = (__ldg(B + (long long)y * wpr + (x0 >> 5)) >> (x0 & 31)) & 0xFFFFu;
            if (nv < 16) b16 &= (1u << nv) - 1u;
            if (sums) {
                const uint4 gv = ld16_any(gbase + (long long)y * rg.w + x0);
                const unsigned gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int left = nv - 4 * q;                   // valid bytes of this word
                    if (left <= 0) break;
                    const unsigned wv = left >= 4 ? gw[q] : (gw[q] & ((1u << (8 * left)) - 1u));
                    const unsigned m4 = (((b16 >> (4 * q)) & 15u) * 0x00204081u) & 0x01010101u;   // mask bits -> 0/1 bytes
                    u_all = __dp4a(wv, 0x01010101u, u_all);
                    u_sq = __dp4a(wv, wv, u_sq);
                    u_fg = __dp4a(wv, m4, u_fg);
                }
            }
            if (b16) {
                a_fg += __popc(b16);
                mnx = min(mnx, x0 + __ffs(b16) - 1); mxx = max(mxx, x0 + 31 - __clz(b16));
                mny = min(mny, y); mxy = max(mxy, y);
            }
            if (out_mask || out_labels) {
                const long long o = rg.out_offset + (long long)y * rg.w + x0;
                for (int j = 0; j < nv; ++j) {
                    const unsigned bit = (b16 >> j) & 1u;
                    if (out_mask) out_mask[o + j] = bit ? 255 : 0;
                    if (out_labels) out_labels[o + j] = (int)bit;
                }
            }
        }
        long long l_fg = a_fg, s_fg = u_fg, s_all = u_all, s_sq = u_sq;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            l_fg += __shfl_xor_sync(0xffffffffu, l_fg, o);
            s_fg += __shfl_xor_sync(0xffffffffu, s_fg, o);
            s_all += __shfl_xor_sync(0xffffffffu, s_all, o);
            s_sq += __shfl_xor_sync(0xffffffffu, s_sq, o);
            mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
            mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
            mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
            mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
        }
        __syncthreads();
        if (lane == 0) {
            s_acc[warp][0] = l_fg; s_acc[warp][1] = s_fg; s_acc[warp][2] = s_all; s_acc[warp][3] = s_sq;
            s_ext[warp][0] = mnx; s_ext[warp][1] = mny; s_ext[warp][2] = mxx; s_ext[warp][3] = mxy;
        }
        __syncthreads();
        if (tid == 0) {
            for (int k = 1; k < K4_THREADS / 32; ++k) {
                l_fg += s_acc[k][0]; s_fg += s_acc[k][1]; s_all += s_acc[k][2]; s_sq += s_acc[k][3];
                mnx = min(mnx, s_ext[k][0]); mny = min(mny, s_ext[k][1]);
                mxx = max(mxx, s_ext[k][2]); mxy = max(mxy, s_ext[k][3]);
            }
            typedef unsigned long long ull;
            if (l_fg) {
                atomicAdd(reinterpret_cast<ull*>(out_area + r), (ull)l_fg);
                if (rg.max_comp > 1) {
                    int* s = stats + ((long long)rg.stats_offset + 1) * 5;
                    atomicMin(s + 0, mnx); atomicMin(s + 1, mny); atomicMax(s + 2, mxx); atomicMax(s + 3, mxy);
                    atomicAdd(s + 4, (int)l_fg);
                }
            }
            if (sums) {
                atomicAdd(reinterpret_cast<ull*>(sums + 4 * r + 0), (ull)s_fg);
                atomicAdd(reinterpret_cast<ull*>(sums + 4 * r + 1), (ull)s_all);
                atomicAdd(reinterpret_cast<ull*>(sums + 4 * r + 2), (ull)s_sq);
            }
        }
    }
}

__global__ void k7_finish(const qc_seg_region* __restrict__ regs, int R, int* __restrict__ stats,
                          const long long* __restrict__ out_area, int* __restrict__ out_ncomp,
                          long long* __restrict__ sums) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < R; r += gridDim.x * blockDim.x) {
        const qc_seg_region rg = regs[r];
        int* s0 = stats + (long long)rg.stats_offset * 5;
        s0[0] = s0[1] = s0[2] = s0[3] = s0[4] = 0;
        const bool any = out_area[r] > 0;
        out_ncomp[r] = any ? 2 : 1;
        if (rg.max_comp > 1) {
            int* s = s0 + 5;
            if (any) { s[2] = s[2] - s[0] + 1; s[3] = s[3] - s[1] + 1; }
            else { s[0] = s[1] = s[2] = s[3] = s[4] = 0; }
        }
        if (sums) sums[4 * r + 3] = (long long)rg.w * rg.h;
    }
}

// ------------------------------------------------------------------------------------------
// K8: contour statistics of the final mask without tracing contours (row a12; inference/segmentation.py:486-525:
// cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE), contours with contourArea < 10 dropped, perimeter = sum of
// arcLength).  The external contours are the outer borders of the 8-connected components of the HOLE-FILLED mask;
// Suzuki's border following moves between pixel centres and every move appears exactly once in one 2x2 pixel window
// around a lattice vertex (outside the ROI = 0):
//     two edge-adjacent foreground pixels -> one axis move;  three foreground pixels -> one diagonal move;
//     two diagonal foreground pixels -> two diagonal moves (a pinch);  one or four -> none
// so arcLength = n_axis + sqrt(2) * n_diag and contourArea = |sum of cross products p x p' over the moves| / 2 (moves
// oriented with the foreground on the right).  Pinned against cv2 on the CPU (oracle/contours.py): counts and areas
// exact, perimeter to the float32 rounding of cv2's per-segment lengths (< 1e-7 relative).
//   k8_kept_bits  (adaptive method) kept mask as bit rows from the final labels + per-label areas
//   k7_fill + K4 union-find on the filled mask (reused)
//   k8_moves      one thread per 32 lattice vertices: window patterns by bit arithmetic, one atomic per run of moves of
//                 the same component into (n_axis | n_diag << 32, cross) accumulators on the component's root block
//   k8_reduce     per region: components with contourArea >= 10 -> (n_axis, n_diag, n_contours)
__global__ void __launch_bounds__(K4_THREADS)
k8_kept_bits(const qc_seg_region* __restrict__ regs, int R, int total_bands, const int* __restrict__ g_band_off,
             const unsigned* __restrict__ bits, const int* __restrict__ label_of_block, const int* __restrict__ stats,
             unsigned* __restrict__ bits_out) {
    __shared__ int s_off[K4_LOC_MAX + 1];
    const int tid = threadIdx.x;
    const int* off = loc_load(s_off, g_band_off, R);
    for (int t = blockIdx.x; t < total_bands; t += gridDim.x) {
        const qc_seg_region rg = regs[loc_find(off, R, t)];
        const int wpr = wpr_of(rg), bw = bw_of(rg);
        const int BR = k4_band_rows(rg.w);
        const int y0 = (t - rg.band_offset) * BR, y1 = min(y0 + BR, rg.h);
        const int by0 = y0 >> 1, by1 = (y1 + 1) >> 1;
        const unsigned* B = bits + rg.bits_offset;
        unsigned* O = bits_out + rg.bits_offset;
        const int* st = stats + (long long)rg.stats_offset * 5;
        for (int ci = tid; ci < (by1 - by0) * wpr; ci += K4_THREADS) {
            const int byy = ci / wpr, wx = ci - byy * wpr, by = by0 + byy;
            const bool two = 2 * by + 1 < rg.h;
            const unsigned r0 = __ldg(B + (long long)(2 * by) * wpr + wx);
            const unsigned r1 = two ? __ldg(B + (long long)(2 * by + 1) * wpr + wx) : 0u;
            const unsigned v = r0 | r1;
            unsigned keep = 0;
            if (v) {
                const int* prow = label_of_block + rg.blk_offset + (long long)by * bw + wx * 16;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    if (!((v >> (2 * k)) & 3u)) continue;
                    const int lab = prow[k];            // components beyond max_comp have no recorded area: kept (as k4_final)
                    if (lab >= rg.max_comp || st[(long long)lab * 5 + 4] >= rg.min_area) keep |= 3u << (2 * k);
                }
            }
            O[(long long)(2 * by) * wpr + wx] = r0 & keep;
            if (two) O[(long long)(2 * by + 1) * wpr + wx] = r1 & keep;
        }
    }
}

__global__ void __launch_bounds__(K4_THREADS)
k8_moves(const qc_seg_region* __restrict__ regs, int R, int total_bands, const int* __restrict__ g_band_off,
         const unsigned* __restrict__ bits, const int* __restrict__ parent, unsigned long long* __restrict__ acc_n,
         long long* __restrict__ acc_x) {
    __shared__ int s_off[K4_LOC_MAX + 1];
    const int tid = threadIdx.x;
    const int* off = loc_load(s_off, g_band_off, R);
    for (int t = blockIdx.x; t < total_bands; t += gridDim.x) {
        const qc_seg_region rg = regs[loc_find(off, R, t)];
        const int wpr = wpr_of(rg), bw = bw_of(rg);
        const int BR = k4_band_rows(rg.w);
        const int y0 = (t - rg.band_offset) * BR, y1 = min(y0 + BR, rg.h);
        const unsigned* B = bits + rg.bits_offset;
        const int* P = parent + rg.blk_offset;
        // vertex rows vy = y0 .. y1-1 (+ row h for the last band), vertex words 0 .. w/32 (vertex column w included)
        const int vy1 = y1 == rg.h ? rg.h + 1 : y1;
        const int vwords = (rg.w >> 5) + 1;
        auto word = [&](int y, int wx) -> unsigned {
            return (y < 0 || y >= rg.h || wx < 0 || wx >= wpr) ? 0u : __ldg(B + (long long)y * wpr + wx);
        };
        for (int it = tid; it < (vy1 - y0) * vwords; it += K4_THREADS) {
            const int vyy = it / vwords, wx = it - vyy * vwords, vy = y0 + vyy;
            const unsigned up = word(vy - 1, wx), dn = word(vy, wx);
            if (!(up | dn) && !((word(vy - 1, wx - 1) | word(vy, wx - 1)) >> 31)) continue;
            // bit j <-> vertex column vx = 32*wx + j: a = (vy-1, vx-1), b = (vy-1, vx), c = (vy, vx-1), d = (vy, vx)
            const unsigned b = up, d = dn;
            const unsigned a = (up << 1) | (word(vy - 1, wx - 1) >> 31), c = (dn << 1) | (word(vy, wx - 1) >> 31);
            const unsigned e_cd = ~a & ~b & c & d, e_ab = a & b & ~c & ~d, e_bd = ~a & b & ~c & d, e_ac = a & ~b & c & ~d;
            const unsigned t_d = a & b & c & ~d, t_a = ~a & b & c & d, t_b = a & ~b & c & d, t_c = a & b & ~c & d;
            const unsigned p_ad = a & ~b & ~c & d, p_bc = ~a & b & c & ~d;
            unsigned ev = e_cd | e_ab | e_bd | e_ac | t_d | t_a | t_b | t_c | p_ad | p_bc;
            int cur = -1;
            unsigned long long n_acc = 0;
            long long x_acc = 0;
            auto flush = [&]() {
                if (cur >= 0) {
                    atomicAdd(acc_n + rg.blk_offset + cur, n_acc);
                    if (x_acc) atomicAdd(reinterpret_cast<unsigned long long*>(acc_x + rg.blk_offset + cur), (unsigned long long)x_acc);
                }
            };
            while (ev) {
                const int j = __ffs(ev) - 1;
                ev &= ev - 1;
                const unsigned m = 1u << j;
                const long long vx = 32 * wx + j;
                const long long ax = vx - 1, ay = vy - 1, bx = vx, by_ = vy - 1, cx = vx - 1, cy = vy, dx = vx, dy = vy;
                long long px, py;                       // a foreground pixel of the window (identifies the component)
                unsigned long long n;                   // n_axis | n_diag << 32
                long long cr = 0;
                if (e_cd & m) { n = 1ull; cr = cx * dy - dx * cy; px = cx; py = cy; }                  // c -> d
                else if (e_ab & m) { n = 1ull; cr = bx * ay - ax * by_; px = ax; py = ay; }            // b -> a
                else if (e_bd & m) { n = 1ull; cr = dx * by_ - bx * dy; px = bx; py = by_; }           // d -> b
                else if (e_ac & m) { n = 1ull; cr = ax * cy - cx * ay; px = ax; py = ay; }             // a -> c
                else if (t_d & m) { n = 1ull << 32; cr = bx * cy - cx * by_; px = ax; py = ay; }       // b -> c
                else if (t_a & m) { n = 1ull << 32; cr = cx * by_ - bx * cy; px = bx; py = by_; }      // c -> b
                else if (t_b & m) { n = 1ull << 32; cr = ax * dy - dx * ay; px = ax; py = ay; }        // a -> d
                else if (t_c & m) { n = 1ull << 32; cr = dx * ay - ax * dy; px = ax; py = ay; }        // d -> a
                else if (p_ad & m) { n = 2ull << 32; px = ax; py = ay; }                               // a <-> d
                else { n = 2ull << 32; px = bx; py = by_; }                                            // b <-> c
                const int root = P[(py >> 1) * bw + (px >> 1)] - rg.blk_offset;
                if (root != cur) { flush(); cur = root; n_acc = 0; x_acc = 0; }
                n_acc += n;
                x_acc += cr;
            }
            flush();
        }
    }
}

__global__ void __launch_bounds__(K4_THREADS)
k8_reduce(const qc_seg_region* __restrict__ regs, int R, int total_bands, const int* __restrict__ g_band_off,
          const unsigned* __restrict__ bits, const int* __restrict__ parent, const unsigned long long* __restrict__ acc_n,
          const long long* __restrict__ acc_x, long long* __restrict__ out_contours) {
    __shared__ int s_off[K4_LOC_MAX + 1];
    __shared__ long long s_acc[K4_THREADS / 32][3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int* off = loc_load(s_off, g_band_off, R);
    for (int t = blockIdx.x; t < total_bands; t += gridDim.x) {
        const int r = loc_find(off, R, t);
        const qc_seg_region rg = regs[r];
        const int wpr = wpr_of(rg), bw = bw_of(rg);
        const int BR = k4_band_rows(rg.w);
        const int y0 = (t - rg.band_offset) * BR, y1 = min(y0 + BR, rg.h);
        const int by0 = y0 >> 1, by1 = (y1 + 1) >> 1;
        const unsigned* B = bits + rg.bits_offset;
        long long n_axis = 0, n_diag = 0, n_cont = 0;
        for (int ci = tid; ci < (by1 - by0) * wpr; ci += K4_THREADS) {
            const int byy = ci / wpr, wx = ci - byy * wpr, by = by0 + byy;
            const unsigned r0 = __ldg(B + (long long)(2 * by) * wpr + wx);
            const unsigned r1 = (2 * by + 1 < rg.h) ? __ldg(B + (long long)(2 * by + 1) * wpr + wx) : 0u;
            unsigned v = r0 | r1;
            v = (v | (v >> 1)) & 0x55555555u;
            while (v) {
                const int k = (__ffs(v) - 1) >> 1;
                v &= v - 1;
                const int g = rg.blk_offset + by * bw + wx * 16 + k;
                if (parent[g] != g) continue;                       // component roots only
                const long long x2 = acc_x[g];
                if ((x2 < 0 ? -x2 : x2) >= 2 * 10) {                // contourArea = |cross| / 2 >= 10  (:498-500)
                    const unsigned long long nn = acc_n[g];
                    n_axis += (long long)(nn & 0xFFFFFFFFull);
                    n_diag += (long long)(nn >> 32);
                    n_cont += 1;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n_axis += __shfl_xor_sync(0xffffffffu, n_axis, o);
            n_diag += __shfl_xor_sync(0xffffffffu, n_diag, o);
            n_cont += __shfl_xor_sync(0xffffffffu, n_cont, o);
        }
        __syncthreads();
        if (lane == 0) { s_acc[warp][0] = n_axis; s_acc[warp][1] = n_diag; s_acc[warp][2] = n_cont; }
        __syncthreads();
        if (tid == 0) {
            for (int k = 1; k < K4_THREADS / 32; ++k) { n_axis += s_acc[k][0]; n_diag += s_acc[k][1]; n_cont += s_acc[k][2]; }
            typedef unsigned long long ull;
            if (n_cont) {
                atomicAdd(reinterpret_cast<ull*>(out_contours + 4 * r + 0), (ull)n_axis);
                atomicAdd(reinterpret_cast<ull*>(out_contours + 4 * r + 1), (ull)n_diag);
                atomicAdd(reinterpret_cast<ull*>(out_contours + 4 * r + 2), (ull)n_cont);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// K9: cv2.Canny(roi, low, high) (aperture 3, L1 gradient) on every ROI and its overlap with the ROI's final mask -- the
// integers behind _calculate_adaptive_confidence (inference/segmentation.py:557-578).  All integer arithmetic, pinned bit
// for bit against cv2 (oracle/canny.py):
//   k9_canny_nms  32x32 tiles: 3x3 Sobel (replicate border), mag = |dx| + |dy| (zero outside the ROI), non-maximum
//                 suppression along the quantised gradient direction (cv2's Q15 tangents) -> candidate / strong bit rows
//   K4 union-find on the candidate rows (8-connectivity); k9_mark flags the components that hold a strong pixel
//   k9_overlap    edges = candidates of flagged components (hysteresis); per region (edges AND mask, edges) counts
constexpr int K9_TG22 = 13573;               // tan(22.5 deg) in Q15

__global__ void __launch_bounds__(K4_THREADS)
k9_canny_nms(const qc_seg_region* __restrict__ regs, int R, int total_bands, const int* __restrict__ g_band_off,
             const uint8_t* __restrict__ gray_ws, int low, int high, unsigned* __restrict__ cand_bits,
             unsigned* __restrict__ strong_bits) {
    __shared__ uint8_t s_g[36][40];
    __shared__ short2 s_d[34][34];
    __shared__ int s_m[34][35];
    __shared__ int s_off[K4_LOC_MAX + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int* off = loc_load(s_off, g_band_off, R);
    for (int t = blockIdx.x; t < total_bands; t += gridDim.x) {
        const qc_seg_region rg = regs[loc_find(off, R, t)];
        const int wpr = wpr_of(rg);
        const int BR = k4_band_rows(rg.w);
        const int y0 = (t - rg.band_offset) * BR, y1 = min(y0 + BR, rg.h);
        const uint8_t* g = gray_ws + rg.out_offset;
        for (int ty = y0; ty < y1; ty += 32) {
            const int yend = min(ty + 32, y1);
            for (int tx = 0; tx < rg.w; tx += 32) {
                __syncthreads();
                for (int i = tid; i < 36 * 36; i += K4_THREADS) {           // gray tile, 2-pixel halo, replicate border
                    const int rr = i / 36, cc = i - rr * 36;
                    const int yy = min(max(ty - 2 + rr, 0), rg.h - 1), xx = min(max(tx - 2 + cc, 0), rg.w - 1);
                    s_g[rr][cc] = g[(long long)yy * rg.w + xx];
                }
                __syncthreads();
                for (int i = tid; i < 34 * 34; i += K4_THREADS) {           // Sobel + magnitude, 1-pixel halo
                    const int rr = i / 34, cc = i - rr * 34;
                    const int y = ty - 1 + rr, x = tx - 1 + cc;
                    int dx = 0, dy = 0, m = 0;
                    if (y >= 0 && y < rg.h && x >= 0 && x < rg.w) {
                        const int a = s_g[rr][cc], b = s_g[rr][cc + 1], c = s_g[rr][cc + 2];
                        const int d = s_g[rr + 1][cc], f = s_g[rr + 1][cc + 2];
                        const int gg = s_g[rr + 2][cc], hh = s_g[rr + 2][cc + 1], ii = s_g[rr + 2][cc + 2];
                        dx = (c + 2 * f + ii) - (a + 2 * d + gg);
                        dy = (gg + 2 * hh + ii) - (a + 2 * b + c);
                        m = abs(dx) + abs(dy);
                    }
                    s_d[rr][cc] = make_short2((short)dx, (short)dy);
                    s_m[rr][cc] = m;
                }
                __syncthreads();
                for (int rr = warp; rr < 32; rr += K4_THREADS / 32) {        // non-maximum suppression
                    const int y = ty + rr, x = tx + lane;
                    bool cand = false, strong = false;
                    if (y < yend && x < rg.w) {
                        const int m = s_m[rr + 1][lane + 1];
                        if (m > low) {
                            const short2 d = s_d[rr + 1][lane + 1];
                            const long long ax = abs((int)d.x), ay = (long long)abs((int)d.y) << 15;
                            const long long tg22x = ax * K9_TG22;
                            if (ay < tg22x) cand = m > s_m[rr + 1][lane] && m >= s_m[rr + 1][lane + 2];
                            else if (ay > tg22x + (ax << 16)) cand = m > s_m[rr][lane + 1] && m >= s_m[rr + 2][lane + 1];
                            else {
                                const int sgn = ((d.x ^ d.y) < 0) ? -1 : 1;
                                cand = m > s_m[rr][lane + 1 - sgn] && m > s_m[rr + 2][lane + 1 + sgn];
                            }
                            strong = cand && m > high;
                        }
                    }
                    const unsigned cw = __ballot_sync(0xffffffffu, cand), sw = __ballot_sync(0xffffffffu, strong);
                    if (lane == 0 && y < yend) {
                        cand_bits[rg.bits_offset + (long long)y * wpr + (tx >> 5)] = cw;
                        strong_bits[rg.bits_offset + (long long)y * wpr + (tx >> 5)] = sw;
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(K4_THREADS)
k9_mark(const qc_seg_region* __restrict__ regs, int R, int total_bands, const int* __restrict__ g_band_off,
        const unsigned* __restrict__ strong_bits, const int* __restrict__ parent, int* __restrict__ flag) {
    __shared__ int s_off[K4_LOC_MAX + 1];
    const int tid = threadIdx.x;
    const int* off = loc_load(s_off, g_band_off, R);
    for (int t = blockIdx.x; t < total_bands; t += gridDim.x) {
        const qc_seg_region rg = regs[loc_find(off, R, t)];
        const int wpr = wpr_of(rg), bw = bw_of(rg);
        const int BR = k4_band_rows(rg.w);
        const int y0 = (t - rg.band_offset) * BR, y1 = min(y0 + BR, rg.h);
        const int by0 = y0 >> 1, by1 = (y1 + 1) >> 1;
        const unsigned* S = strong_bits + rg.bits_offset;
        for (int ci = tid; ci < (by1 - by0) * wpr; ci += K4_THREADS) {
            const int byy = ci / wpr, wx = ci - byy * wpr, by = by0 + byy;
            unsigned v = __ldg(S + (long long)(2 * by) * wpr + wx);
            if (2 * by + 1 < rg.h) v |= __ldg(S + (long long)(2 * by + 1) * wpr + wx);
            v = (v | (v >> 1)) & 0x55555555u;
            while (v) {
                const int k = (__ffs(v) - 1) >> 1;
                v &= v - 1;
                flag[parent[rg.blk_offset + by * bw + wx * 16 + k]] = 1;      // (every writer stores the same value)
            }
        }
    }
}

__global__ void __launch_bounds__(K4_THREADS)
k9_overlap(const qc_seg_region* __restrict__ regs, int R, int total_bands, const int* __restrict__ g_band_off,
           const unsigned* __restrict__ cand_bits, const unsigned* __restrict__ mask_bits, const int* __restrict__ parent,
           const int* __restrict__ flag, long long* __restrict__ sums) {
    __shared__ int s_off[K4_LOC_MAX + 1];
    __shared__ int s_acc[K4_THREADS / 32][2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int* off = loc_load(s_off, g_band_off, R);
    for (int t = blockIdx.x; t < total_bands; t += gridDim.x) {
        const int r = loc_find(off, R, t);
        const qc_seg_region rg = regs[r];
        const int wpr = wpr_of(rg), bw = bw_of(rg);
        const int BR = k4_band_rows(rg.w);
        const int y0 = (t - rg.band_offset) * BR, y1 = min(y0 + BR, rg.h);
        const int by0 = y0 >> 1, by1 = (y1 + 1) >> 1;
        const unsigned* Cb = cand_bits + rg.bits_offset;
        const unsigned* Mb = mask_bits + rg.bits_offset;
        int n_edge = 0, n_both = 0;
        for (int ci = tid; ci < (by1 - by0) * wpr; ci += K4_THREADS) {
            const int byy = ci / wpr, wx = ci - byy * wpr, by = by0 + byy;
            const bool two = 2 * by + 1 < rg.h;
            const unsigned c0 = __ldg(Cb + (long long)(2 * by) * wpr + wx);
            const unsigned c1 = two ? __ldg(Cb + (long long)(2 * by + 1) * wpr + wx) : 0u;
            unsigned v = c0 | c1;
            if (!v) continue;
            v = (v | (v >> 1)) & 0x55555555u;
            unsigned keep = 0;
            while (v) {
                const int k = (__ffs(v) - 1) >> 1;
                v &= v - 1;
                if (flag[parent[rg.blk_offset + by * bw + wx * 16 + k]]) keep |= 3u << (2 * k);
            }
            const unsigned e0 = c0 & keep, e1 = c1 & keep;
            n_edge += __popc(e0) + __popc(e1);
            n_both += __popc(e0 & __ldg(Mb + (long long)(2 * by) * wpr + wx));
            if (two) n_both += __popc(e1 & __ldg(Mb + (long long)(2 * by + 1) * wpr + wx));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n_edge += __shfl_xor_sync(0xffffffffu, n_edge, o);
            n_both += __shfl_xor_sync(0xffffffffu, n_both, o);
        }
        __syncthreads();
        if (lane == 0) { s_acc[warp][0] = n_edge; s_acc[warp][1] = n_both; }
        __syncthreads();
        if (tid == 0) {
            for (int k = 1; k < K4_THREADS / 32; ++k) { n_edge += s_acc[k][0]; n_both += s_acc[k][1]; }
            typedef unsigned long long ull;
            if (n_edge) {
                atomicAdd(reinterpret_cast<ull*>(sums + 4 * r + 0), (ull)n_both);
                atomicAdd(reinterpret_cast<ull*>(sums + 4 * r + 1), (ull)n_edge);
            }
        }
    }
}
