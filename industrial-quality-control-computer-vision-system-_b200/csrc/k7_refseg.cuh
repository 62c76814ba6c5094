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
    if (im.channels == 3) { g.base = gray_ws + rg.out_offset; g.pitch = rg.w; }
    else { g.base = im.data + (long long)rg.img * im.img_stride + (long long)rg.y * im.row_stride + rg.x; g.pitch = im.row_stride; }
    return g;
}

// per-region scalars of the threshold method
struct K7Scalars {
    unsigned long long* minmax;   // [R][2] bit patterns of the (non-negative) float64 min / max
    double* thr_adj;              // [R] threshold x class factor
};

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
                for (int i = tid; i < 40 * 40; i += K4_THREADS) {
                    const int rr = i / 40, cc = i - rr * 40;
                    const int yy = min(max(ty - 4 + rr, 0), rg.h - 1), xx = min(max(tx - 4 + cc, 0), rg.w - 1);
                    s_src[rr][cc] = __dmul_rn((double)g.base[(long long)yy * g.pitch + xx], K7_INV255);
                }
                __syncthreads();
                for (int i = tid; i < 32 * 40; i += K4_THREADS) {          // axis 0 first
                    const int rr = i / 40, cc = i - rr * 40;
                    s_v[rr][cc] = k7_tap9(s_src[rr + 4][cc], s_src[rr][cc], s_src[rr + 8][cc], s_src[rr + 1][cc],
                                          s_src[rr + 7][cc], s_src[rr + 2][cc], s_src[rr + 6][cc], s_src[rr + 3][cc],
                                          s_src[rr + 5][cc]);
                }
                __syncthreads();
                for (int i = tid; i < 32 * 32; i += K4_THREADS) {          // then axis 1
                    const int rr = i >> 5, cc = i & 31;
                    const int y = ty + rr, x = tx + cc;
                    if (y < y1 && x < rg.w) {
                        const double o = k7_tap9(s_v[rr][cc + 4], s_v[rr][cc], s_v[rr][cc + 8], s_v[rr][cc + 1],
                                                 s_v[rr][cc + 7], s_v[rr][cc + 2], s_v[rr][cc + 6], s_v[rr][cc + 3],
                                                 s_v[rr][cc + 5]);
                        blurred[rg.out_offset + (long long)y * rg.w + x] = o;
                        mn = fmin(mn, o);
                        mx = fmax(mx, o);
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
    __shared__ int s_off[K4_LOC_MAX + 1];
    const int tid = threadIdx.x;
    const int* off = loc_load(s_off, g_band_off, R);
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
        __syncthreads();
        const double denom = __dsub_rn(last, first), step = __ddiv_rn(denom, 256.0);
        const double* p = blurred + rg.out_offset + (long long)y0 * rg.w;
        const int n = (y1 - y0) * rg.w;
        for (int i = tid; i < n; i += K4_THREADS) {
            const double x = p[i];
            // numpy: f_indices = ((a - first) / norm_denom) * n_bins; astype(intp); == n_bins -> -1;
            // then one corrective step down / up against the linspace edges
            int idx = (int)__dmul_rn(__ddiv_rn(__dsub_rn(x, first), denom), 256.0);
            if (idx == 256) idx = 255;
            if (x < k7_edge(idx, first, last, step)) idx -= 1;
            if (idx != 255 && x >= k7_edge(idx + 1, first, last, step)) idx += 1;
            atomicAdd(&s_h[idx], 1u);
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
        for (int it = warp; it < items; it += K4_THREADS / 32) {
            const int yy = it / wpr, wx = it - yy * wpr;
            const int y = y0 + yy, x = wx * 32 + lane;
            bool fg = false;
            if (x < rg.w) {
                const double v = blurred[rg.out_offset + (long long)y * rg.w + x];
                fg = rg.polarity ? (v > T) : (v < T);
            }
            const unsigned word = __ballot_sync(0xffffffffu, fg);
            if (lane == 0) bits[rg.bits_offset + (long long)y * wpr + wx] = word;
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
            int akern_pitch, int adaptive_c, unsigned* __restrict__ bits, int pmax) {
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
__global__ void __launch_bounds__(K7_FILL_THREADS)
k7_fill(const qc_seg_region* __restrict__ regs, int R, const unsigned* __restrict__ bits_in,
        unsigned* __restrict__ bits_out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = K7_FILL_THREADS / 32;
    for (int r = blockIdx.x; r < R; r += gridDim.x) {
        const qc_seg_region rg = regs[r];
        const int wpr = wpr_of(rg), h = rg.h;
        const unsigned tail = (rg.w & 31) ? ((1u << (rg.w & 31)) - 1u) : 0xFFFFFFFFu;   // valid bits of the last word
        const unsigned* B = bits_in + rg.bits_offset;
        volatile unsigned* O = bits_out + rg.bits_offset;
        auto freew = [&](int y, int wx) -> unsigned {
            const unsigned f = ~__ldg(B + (long long)y * wpr + wx);
            return wx == wpr - 1 ? (f & tail) : f;
        };
        __syncthreads();
        // seeds: the free pixels of the ROI's outer ring
        for (int i = tid; i < h * wpr; i += K7_FILL_THREADS) {
            const int y = i / wpr, wx = i - y * wpr;
            const unsigned f = freew(y, wx);
            unsigned ring = 0;
            if (y == 0 || y == h - 1) ring = 0xFFFFFFFFu;
            if (wx == 0) ring |= 1u;
            if (wx == wpr - 1) ring |= 1u << ((rg.w - 1) & 31);
            O[i] = f & ring;
        }
        __syncthreads();
        const int S = max(8, (h + NW - 1) / NW);              // rows per strip
        const int nstrips = (h + S - 1) / S;
        const int nch = (wpr + 31) >> 5;
        for (;;) {
            int changed = 0;
            for (int st = warp; st < nstrips; st += NW) {
                const int a = st * S, b = min(a + S, h);
                for (int dir = 0; dir < 2; ++dir) {
                    for (int k = 0; k < b - a; ++k) {
                        const int y = dir == 0 ? a + k : b - 1 - k;
                        const int yn = dir == 0 ? y - 1 : y + 1;              // the row the flood comes from
                        for (int c = 0; c < nch; ++c) {
                            const int wx = c * 32 + lane;
                            const bool valid = wx < wpr;
                            const unsigned f = valid ? freew(y, wx) : 0u;
                            const unsigned o = valid ? O[(long long)y * wpr + wx] : 0u;
                            unsigned s = o;
                            if (valid && yn >= 0 && yn < h) s |= O[(long long)yn * wpr + wx] & f;
                            for (;;) {
                                s = k7_hfill(f, s);
                                unsigned l = __shfl_up_sync(0xffffffffu, s, 1) >> 31;
                                unsigned rt = __shfl_down_sync(0xffffffffu, s, 1) & 1u;
                                if (lane == 0) l = (valid && wx > 0) ? (O[(long long)y * wpr + wx - 1] >> 31) : 0u;
                                if (lane == 31) rt = (valid && wx + 1 < wpr) ? (O[(long long)y * wpr + wx + 1] & 1u) : 0u;
                                const unsigned ns = s | (f & (l | (rt << 31)));
                                const bool more = ns != s;
                                s = ns;
                                if (!__any_sync(0xffffffffu, more)) break;
                            }
                            if (valid && s != o) { O[(long long)y * wpr + wx] = s; changed = 1; }
                            __syncwarp();
                        }
                    }
                }
            }
            if (!__syncthreads_or(changed)) break;
        }
        // holes = free pixels the flood never reached: result = ~outside inside the ROI
        for (int i = tid; i < h * wpr; i += K7_FILL_THREADS) {
            const int wx = i % wpr;
            const unsigned o = O[i];
            O[i] = wx == wpr - 1 ? (~o & tail) : ~o;
        }
    }
}

// ------------------------------------------------------------------------------------------
// final pass of the threshold method: mask bytes / 0-1 labels, area, extent, confidence sums
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
        const K7Gray g = k7_gray_of(im, rg, gray_ws);
        const unsigned* B = bits + rg.bits_offset;
        long long a_fg = 0, s_fg = 0, s_all = 0, s_sq = 0;
        int mnx = INT_MAX, mny = INT_MAX, mxx = -1, mxy = -1;
        const int n = (y1 - y0) * rg.w;
        for (int i = tid; i < n; i += K4_THREADS) {
            const int yy = i / rg.w, x = i - yy * rg.w, y = y0 + yy;
            const unsigned bit = (__ldg(B + (long long)y * wpr + (x >> 5)) >> (x & 31)) & 1u;
            const long long o = rg.out_offset + (long long)y * rg.w + x;
            if (out_mask) out_mask[o] = bit ? 255 : 0;
            if (out_labels) out_labels[o] = (int)bit;
            if (sums) {
                const int v = g.base[(long long)y * g.pitch + x];
                s_all += v;
                s_sq += v * v;
                if (bit) s_fg += v;
            }
            if (bit) {
                a_fg += 1;
                mnx = min(mnx, x); mxx = max(mxx, x); mny = min(mny, y); mxy = max(mxy, y);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a_fg += __shfl_xor_sync(0xffffffffu, a_fg, o);
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
            s_acc[warp][0] = a_fg; s_acc[warp][1] = s_fg; s_acc[warp][2] = s_all; s_acc[warp][3] = s_sq;
            s_ext[warp][0] = mnx; s_ext[warp][1] = mny; s_ext[warp][2] = mxx; s_ext[warp][3] = mxy;
        }
        __syncthreads();
        if (tid == 0) {
            for (int k = 1; k < K4_THREADS / 32; ++k) {
                a_fg += s_acc[k][0]; s_fg += s_acc[k][1]; s_all += s_acc[k][2]; s_sq += s_acc[k][3];
                mnx = min(mnx, s_ext[k][0]); mny = min(mny, s_ext[k][1]);
                mxx = max(mxx, s_ext[k][2]); mxy = max(mxy, s_ext[k][3]);
            }
            typedef unsigned long long ull;
            if (a_fg) {
                atomicAdd(reinterpret_cast<ull*>(out_area + r), (ull)a_fg);
                if (rg.max_comp > 1) {
                    int* s = stats + ((long long)rg.stats_offset + 1) * 5;
                    atomicMin(s + 0, mnx); atomicMin(s + 1, mny); atomicMax(s + 2, mxx); atomicMax(s + 3, mxy);
                    atomicAdd(s + 4, (int)a_fg);
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
