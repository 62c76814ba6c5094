// k4_segment.cu -- K4: defect masking on device.
//
// Per region (full frame or detection ROI), replaces the reference's
//   cv2.cvtColor(RGB2GRAY)                         inference/segmentation.py:66-69
//   Otsu threshold + polarity compare              :212-234 (u8 Otsu == cv2.threshold OTSU)
//   cv2.morphologyEx CLOSE, OPEN with 3x3 cross    :284-286 (iters=2 == skimage disk(2), :458-460)
//   cv2.connectedComponents + bincount + isin      :464-484
//   np.sum(mask > 0)                               :98
// Output labels carry cv2.connectedComponents' numbering (8-connectivity; components ordered by
// their first 2x2 block in block-raster order), which is exactly "rank of the minimum block
// index of the component" -- so the union-find works on 2x2 blocks with min-index roots and a
// raster-order prefix sum over root flags gives the label.
//
// Pipeline (all HBM/L2-bound integer work, no tensor cores):
//   k4_init      zero histograms / area, sentinel-initialise stats rows
//   k4_gray      RGB -> gray (cv2 fixed point) into a region-packed copy (RGB input only)
//   k4_hist      256-bin histogram: 128-bit loads, per-warp lane-private u16 counters, per-region flush
//   k4_otsu      parallel Otsu from exact integer prefix sums; replays cv2's sequential float64
//                recurrence only when the maximum is not decisive (ties / plateaus)
//   k4_mask      threshold (128-bit loads, SIMD byte compares) -> bit-packed rows in smem -> close/open
//                (4*iters cross steps with halo) -> global bit rows (1 bit/pixel)
//   k4_ccl_local band-local 2x2-block union-find in shared memory: bit-parallel run contacts per 32-px word,
//                first-contact links + pointer jumping, unions only for the remaining contacts
//   k4_border    unions across band borders (atomicMin in L2, roots = min block index)
//   k4_compress  path compression + per-1024-block root counts
//   k4_rank      root -> label (1-based rank in block-raster order; sums the earlier chunks' root counts itself), ncomp
//   k4_stats     parent -> label in place; per-label area / bbox via run-aggregated atomics
//   k4_final     int32 labels + u8 mask (min-area filter applied), kept area (from the stats rows alone when no
//                pixel output is wanted), bbox -> (x,y,w,h)
// Algorithmic bytes per region: w*h (gray read) + w*h (mask) + 4*w*h (labels) + 20*ncomp.
#include "common.cuh"
#include <math.h>
#include <float.h>
#include <limits.h>
#include <algorithm>

namespace {

constexpr int K4_THREADS = 256;
// Pixel rows per band, chosen per region so rows x words-per-row stays ~512 (bounded shared-memory footprint):
// 16 rows for a 2048-px-wide frame, 48 rows for a 300-px ROI, 96 rows for a 150-px ROI -- the per-band fixed costs
// (barriers, region lookup, flushes) are paid once per ~16K pixels whatever the region shape, and the band-local
// union-find trees stay shallow enough for 8 CTAs per SM to hide their shared-memory latency.
constexpr int K4_BAND_RW = 512;
constexpr int K4_MIN_ROWS = 16;      // (32 was tried for wide frames: mask and border gain what ccl_local loses to occupancy)
__host__ __device__ inline int k4_band_rows(int w) {
    const int wpr = (w + 31) >> 5;
    const int r = (K4_BAND_RW / wpr) & ~15;
    return r < K4_MIN_ROWS ? K4_MIN_ROWS : (r > 512 ? 512 : r);
}
constexpr int K4_CHUNK = 256;        // word cells (16 blocks each) per ranking chunk

struct SegWs {
    unsigned* hist;       // [R][256]
    int* cmp;             // [R][2]  inclusive foreground interval [lo, hi]
    uint8_t* gray;        // [total_pixels] (RGB input only)
    unsigned* bits;       // [total_bits_words]
    int* parent;          // [total_blocks]   union-find parent, later the label
    int* rank;            // [total_blocks]   label of root blocks
    int* chunk_count;     // [total_chunks]
    int* chunk_base;      // [total_chunks]
    int* band_off;        // [R+1] cumulative bands
    int* chunk_off;       // [R+1] cumulative ranking chunks
    // K7 (reference methods) only
    unsigned* bits2;      // [total_bits_words]
    unsigned* bits3;      // [total_bits_words]
    double* blurred;      // [total_pixels]   sigma=1 float64 image (threshold method)
    unsigned long long* minmax;   // [R][2]
    double* thr_adj;      // [R]
    unsigned long long* acc_n;    // [total_blocks]  K8 per-root (n_axis | n_diag << 32)
    long long* acc_x;             // [total_blocks]  K8 per-root sum of cross products
    int* queue;                   // [16] work counters of the kernels that hand regions out dynamically (zeroed per call)
    size_t total;
};

SegWs carve(void* base, const qc_seg_params* p, int method = QC_SEG_COMPOSITE) {
    SegWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = qc_align_up(off + bytes, 256); return o; };
    uint8_t* b = (uint8_t*)base;
    const size_t R = (size_t)p->num_regions;
    w.hist = (unsigned*)(b + take(R * 256 * 4));
    w.cmp = (int*)(b + take(R * 2 * 4));
    // (+64: ld16_any and the K7 tiles may read a few bytes past the last region)
    w.gray = (uint8_t*)(b + take((p->channels == 3 || method != QC_SEG_COMPOSITE) ? (size_t)p->total_pixels + 64 : 0));
    w.bits = (unsigned*)(b + take((size_t)p->total_bits_words * 4));
    w.parent = (int*)(b + take((size_t)p->total_blocks * 4));
    w.rank = (int*)(b + take((size_t)p->total_blocks * 4));
    w.chunk_count = (int*)(b + take((size_t)p->total_chunks * 4));
    w.chunk_base = (int*)(b + take((size_t)p->total_chunks * 4));
    w.band_off = (int*)(b + take((R + 1) * 4));
    w.chunk_off = (int*)(b + take((R + 1) * 4));
    const bool ref = method != QC_SEG_COMPOSITE, thr = method == QC_SEG_REF_THRESHOLD;
    w.bits2 = (unsigned*)(b + take(ref ? (size_t)p->total_bits_words * 4 : 0));
    w.bits3 = (unsigned*)(b + take(ref ? (size_t)p->total_bits_words * 4 : 0));
    w.blurred = (double*)(b + take(thr ? (size_t)p->total_pixels * 8 : 0));
    w.minmax = (unsigned long long*)(b + take(thr ? R * 16 : 0));
    w.thr_adj = (double*)(b + take(thr ? R * 8 : 0));
    w.acc_n = (unsigned long long*)(b + take(ref ? (size_t)p->total_blocks * 8 : 0));
    w.acc_x = (long long*)(b + take(ref ? (size_t)p->total_blocks * 8 : 0));
    w.queue = (int*)(b + take(64));
    w.total = off;
    return w;
}

struct Img {
    const uint8_t* data;
    long long img_stride;
    int row_stride, channels, swap_rb;
};

// Work item -> region.  The cumulative offset tables (R+1 ints, written by k4_init) are cached in
// shared memory when they fit, so every thread can binary-search them without a barrier.
constexpr int K4_LOC_MAX = 2047;
__device__ __forceinline__ const int* loc_load(int* s_off, const int* __restrict__ g_off, int R) {
    if (R > K4_LOC_MAX) return g_off;
    for (int i = threadIdx.x; i <= R; i += blockDim.x) s_off[i] = g_off[i];
    __syncthreads();
    return s_off;
}
__device__ __forceinline__ int loc_find(const int* off, int R, int t) {   // largest r with off[r] <= t
    int lo = 0, hi = R - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (off[mid] <= t) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ int wpr_of(const qc_seg_region& r) { return (r.w + 31) >> 5; }
__device__ __forceinline__ int bw_of(const qc_seg_region& r) { return (r.w + 1) >> 1; }
__device__ __forceinline__ int bh_of(const qc_seg_region& r) { return (r.h + 1) >> 1; }

// ------------------------------------------------------------------------------------------
__global__ void k4_init(const qc_seg_region* __restrict__ regs, int R, unsigned* __restrict__ hist,
                        int* __restrict__ stats, long long* __restrict__ out_area, int total_stats_rows,
                        int* __restrict__ band_off, int* __restrict__ chunk_off, int total_bands, int total_chunks) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    for (int i = gid; i <= R; i += gsz) {
        band_off[i] = i < R ? regs[i].band_offset : total_bands;
        chunk_off[i] = i < R ? regs[i].chunk_offset : total_chunks;
    }
    for (int i = gid; i < R * 256; i += gsz) hist[i] = 0;
    for (int i = gid; i < R; i += gsz) out_area[i] = 0;
    // rows of (minx, miny, maxx, maxy, area) = (INT_MAX, INT_MAX, -1, -1, 0), written as one coalesced int stream
    // (total_stats_rows == 0 when the labelling tail initialises the rows it hands out itself: k4_rank)
    const long long nints = (long long)total_stats_rows * 5;
    for (long long i = gid; i < nints; i += gsz) {
        const int c = (int)(i % 5);
        stats[i] = c < 2 ? INT_MAX : (c < 4 ? -1 : 0);
    }
}

// ------------------------------------------------------------------------------------------
// Histogram.  128-bit loads, per-warp lane-private u16 counters ([bin][lane]: no atomics, no intra-warp address
// collisions; counter updates are issued four at a time with duplicates merged in registers).
constexpr int K4_HIST_THREADS = 384;                             // 12 warps: the kernel is latency-bound, one CTA per SM
constexpr int K4_HIST_SMEM = (K4_HIST_THREADS / 32) * 256 * 32 * 2;   // 192 KB

// RGB -> gray (cv2 fixed point) into the region-packed gray workspace; plain high-occupancy kernel so the
// histogram and threshold passes can stream contiguous bytes whatever the ROI alignment in the frame
__global__ void __launch_bounds__(K4_THREADS)
k4_gray(Img im, const qc_seg_region* __restrict__ regs, int R, int total_bands, const int* __restrict__ g_band_off,
        uint8_t* __restrict__ gray_ws) {
    __shared__ int s_off[K4_LOC_MAX + 1];
    const int* off = loc_load(s_off, g_band_off, R);
    for (int t = blockIdx.x; t < total_bands; t += gridDim.x) {
        const qc_seg_region rg = regs[loc_find(off, R, t)];
        const int BR = k4_band_rows(rg.w);
        const int y0 = (t - rg.band_offset) * BR;
        const int y1 = min(y0 + BR, rg.h);
        const uint8_t* ibase = im.data + (long long)rg.img * im.img_stride;
        auto gray_of = [&](unsigned c0, unsigned c1, unsigned c2) -> unsigned {
            // cv2 RGB2GRAY fixed point: (R*9798 + G*19235 + B*3735 + 2^14) >> 15
            return (9798u * (im.swap_rb ? c2 : c0) + 19235u * c1 + 3735u * (im.swap_rb ? c0 : c2) + 16384u) >> 15;
        };
        // Work items = (row, group) flattened over the band, so a 150-px ROI row (39 groups) does not leave half of
        // a warp idle and the per-row set-up is paid per item, not per lane and row.  Per row: item 0 = the <= 3
        // pixels in front of the first pixel whose 3 bytes start 4-byte aligned (pixel index == address mod 4),
        // items 1..ngrp = groups of 4 pixels = three aligned 32-bit words, last item = the <= 3 tail pixels.
        const int G = (rg.w >> 2) + 2;
        const float inv_G = 1.0f / (float)G;
        const int nitems = (y1 - y0) * G;
        for (int it = threadIdx.x; it < nitems; it += K4_THREADS) {
            const int yy = (int)(((float)it + 0.5f) * inv_G);                // it / G (exact: it < 2^20)
            const int gi = it - yy * G, y = y0 + yy;
            const uint8_t* p = ibase + (long long)(rg.y + y) * im.row_stride + (long long)rg.x * 3;
            uint8_t* gq = gray_ws + rg.out_offset + (long long)y * rg.w;
            const int head = min(rg.w, (int)(reinterpret_cast<uintptr_t>(p) & 3));
            const int ngrp = (rg.w - head) >> 2;
            if (gi >= 1 && gi <= ngrp) {
                const int g = gi - 1;
                const unsigned* pw = reinterpret_cast<const unsigned*>(p + 3 * head) + 3 * g;
                const unsigned a = __ldg(pw), b2 = __ldg(pw + 1), c = __ldg(pw + 2);
                const unsigned o0 = gray_of(a & 255u, (a >> 8) & 255u, (a >> 16) & 255u);
                const unsigned o1 = gray_of(a >> 24, b2 & 255u, (b2 >> 8) & 255u);
                const unsigned o2 = gray_of((b2 >> 16) & 255u, b2 >> 24, c & 255u);
                const unsigned o3 = gray_of((c >> 8) & 255u, (c >> 16) & 255u, c >> 24);
                uint8_t* o = gq + head + 4 * g;
                if ((reinterpret_cast<uintptr_t>(o) & 3) == 0) {
                    *reinterpret_cast<unsigned*>(o) = o0 | (o1 << 8) | (o2 << 16) | (o3 << 24);
                } else {
                    o[0] = (uint8_t)o0; o[1] = (uint8_t)o1; o[2] = (uint8_t)o2; o[3] = (uint8_t)o3;
                }
            } else if (gi == 0 || gi == ngrp + 1) {
                const int xa = gi == 0 ? 0 : head + 4 * ngrp, xb = gi == 0 ? head : rg.w;
                for (int x = xa; x < xb; ++x) gq[x] = (uint8_t)gray_of(__ldg(p + 3 * x), __ldg(p + 3 * x + 1), __ldg(p + 3 * x + 2));
            }
        }
    }
}

// one warp counts a contiguous byte span into its lane-private table: scalar head/tail, 128-bit body in batches of
// four vectors per lane (2 KB per warp), the next batch loaded before the current one is counted
__device__ __forceinline__ void hist_span(const uint8_t* __restrict__ p, int len, unsigned short* tab, int lane) {
    const int head = min(len, (int)((16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15));
    const int nvec = (len - head) >> 4;
    for (int x = lane; x < head; x += 32) tab[(unsigned)__ldg(p + x) << 5]++;
    const uint4* pv = reinterpret_cast<const uint4*>(p + head);
    uint4 cur[4], nxt[4];
    auto load = [&](uint4* v, int s0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int sidx = s0 + j * 32 + lane;
            v[j] = sidx < nvec ? __ldg(pv + sidx) : make_uint4(0, 0, 0, 0);
        }
    };
    load(cur, 0);
    for (int s0 = 0; s0 < nvec; s0 += 128) {
        if (s0 + 128 < nvec) load(nxt, s0 + 128);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (s0 + j * 32 + lane < nvec) {
                const unsigned w4[4] = {cur[j].x, cur[j].y, cur[j].z, cur[j].w};
                // 4 pixels per round: the four counter loads are issued together and duplicates inside the group
                // are merged in registers (every duplicate stores the same total)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const unsigned g0 = w4[q] & 255u, g1 = (w4[q] >> 8) & 255u;
                    const unsigned g2 = (w4[q] >> 16) & 255u, g3 = w4[q] >> 24;
                    const unsigned c0 = tab[g0 << 5], c1 = tab[g1 << 5], c2 = tab[g2 << 5], c3 = tab[g3 << 5];
                    const unsigned e01 = g0 == g1, e02 = g0 == g2, e03 = g0 == g3;
                    const unsigned e12 = g1 == g2, e13 = g1 == g3, e23 = g2 == g3;
                    tab[g0 << 5] = (unsigned short)(c0 + 1 + e01 + e02 + e03);
                    tab[g1 << 5] = (unsigned short)(c1 + 1 + e01 + e12 + e13);
                    tab[g2 << 5] = (unsigned short)(c2 + 1 + e02 + e12 + e23);
                    tab[g3 << 5] = (unsigned short)(c3 + 1 + e03 + e13 + e23);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) cur[j] = nxt[j];
    }
    for (int x = head + nvec * 16 + lane; x < len; x += 32) tab[(unsigned)__ldg(p + x) << 5]++;
}

// Work items are quarter-bands (~4K pixels of a detection ROI, 8K of a 2048-px frame); every WARP takes a contiguous
// range of them, counts into its own lane-private table and flushes (reduce + atomicAdd of the non-zero bins + clear)
// only when the region changes.  No CTA barrier, a dozen independent streams per SM whatever the region sizes, an
// even split of the pixels, and only the few warps that share a region ever meet on the same global counters.
constexpr int K4_HIST_SPLIT = 4;
constexpr int K4_HIST_FLUSH_PX = 1 << 20;        // forced flush: a lane never counts more than 32K pixels in between

__global__ void __launch_bounds__(K4_HIST_THREADS)
k4_hist(Img im, const qc_seg_region* __restrict__ regs, int R, int total_bands, const int* __restrict__ g_band_off,
        unsigned* __restrict__ hist, const uint8_t* __restrict__ gray_ws) {
    extern __shared__ __align__(16) unsigned char hsm[];
    __shared__ int s_off[K4_LOC_MAX + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int* off = loc_load(s_off, g_band_off, R);
    unsigned short* tab = reinterpret_cast<unsigned short*>(hsm) + warp * 8192 + lane;   // [bin*32]
    uint4* z = reinterpret_cast<uint4*>(hsm) + warp * 1024;
    const unsigned* t32 = reinterpret_cast<const unsigned*>(hsm) + warp * 4096;
    const int nw = K4_HIST_THREADS / 32;
    const int nitems = total_bands * K4_HIST_SPLIT;
    const int per = (nitems + gridDim.x * nw - 1) / (gridDim.x * nw);
    const int i0 = (blockIdx.x * nw + warp) * per, i1 = min(i0 + per, nitems);
    if (i0 >= i1) return;
    auto clear = [&]() {
        for (int i = lane; i < 1024; i += 32) z[i] = make_uint4(0, 0, 0, 0);
        __syncwarp();
    };
    auto flush = [&](int r) {
        __syncwarp();
        // lane owns bins lane, lane+32, ...: sum the 32 lane-private u16 counters of each (16 words, rotated start
        // so the 32 bins of a round hit 32 banks)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int bin = lane + 32 * b;
            unsigned sum = 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const unsigned v = t32[bin * 16 + ((k + (bin >> 1)) & 15)];
                sum += (v & 0xFFFFu) + (v >> 16);
            }
            if (sum) atomicAdd(&hist[(long long)r * 256 + bin], sum);
        }
        __syncwarp();
    };
    clear();
    int cur_r = -1, cur_t = -1, acc = 0;
    qc_seg_region rg;
    for (int it = i0; it < i1; ++it) {
        const int t = it / K4_HIST_SPLIT, q = it - t * K4_HIST_SPLIT;
        if (t != cur_t) {
            cur_t = t;
            const int r = loc_find(off, R, t);
            if (r != cur_r || acc > K4_HIST_FLUSH_PX) {
                if (cur_r >= 0) { flush(cur_r); clear(); }
                acc = 0;
            }
            if (r != cur_r) { cur_r = r; rg = regs[r]; }
        }
        const int BR = k4_band_rows(rg.w);
        const int yb0 = (t - rg.band_offset) * BR, yb1 = min(yb0 + BR, rg.h);
        const int qr = (yb1 - yb0 + K4_HIST_SPLIT - 1) / K4_HIST_SPLIT;
        const int y0 = yb0 + q * qr, y1 = min(y0 + qr, yb1);
        if (y0 >= y1) continue;
        acc += (y1 - y0) * rg.w;
        if (im.channels == 1 && rg.w != im.row_stride) {
            const uint8_t* ibase = im.data + (long long)rg.img * im.img_stride + (long long)rg.y * im.row_stride + rg.x;
            for (int y = y0; y < y1; ++y) hist_span(ibase + (long long)y * im.row_stride, rg.w, tab, lane);
        } else {
            // full-width rows of the caller's gray frame, or the region-packed gray copy: one contiguous span
            const uint8_t* p = im.channels == 1
                ? im.data + (long long)rg.img * im.img_stride + (long long)(rg.y + y0) * im.row_stride
                : gray_ws + rg.out_offset + (long long)y0 * rg.w;
            hist_span(p, (y1 - y0) * rg.w, tab, lane);
        }
    }
    if (cur_r >= 0) flush(cur_r);
}

// ------------------------------------------------------------------------------------------
// Otsu threshold, bit-exact with cv2's getThreshVal_Otsu_8u.  cv2's recurrence is sequential in
// float64; 256 threads first evaluate the between-class variance of every split from exact integer
// prefix sums.  If the best split beats every other by a relative 1e-9 (rounding in either
// formulation is ~1e-13) its index IS cv2's answer; otherwise (plateaus of two-level images,
// ties, bins on the FLT_EPSILON skip boundary) one thread replays cv2's exact recurrence.
// Warp-cooperative replay of cv2's recurrence.  What is inherently sequential is split into two short chains run by
// lane 0: q1 (one addition per bin) and mu1.  cv2's mu1 step divides by q1; with the correctly rounded reciprocal
// y = RN(1/q1) prepared in parallel, Markstein's sequence q = RN(a*y), r = a - q*q1 (exact in one fma),
// RN(q + r*y) gives exactly the correctly rounded quotient RN(a/q1) that the division instruction would (checked
// against exact rational arithmetic on 3e5 operand pairs of this range; the cv2-pinned parity tests cover the rest),
// in five dependent double operations per bin instead of a ~30-instruction software division.  Everything else
// (p_i, i*p_i, the reciprocals, sigma of every bin, first strict maximum) is evaluated by all 32 lanes.  The total
// mean is a sum of exact integers below 2^53, so its summation order is irrelevant.
constexpr int K4_OTSU_WARPS = 4;
__device__ int otsu_replay_cv2(const unsigned* h, long long npix, double mu_sum, double* s_q, double* s_y, double* s_ip,
                               double* s_m, int lane) {
    const double scale = __ddiv_rn(1.0, (double)npix);
    const double mu = __dmul_rn(mu_sum, scale);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int i = lane * 8 + k;
        const double p_i = __dmul_rn((double)h[i], scale);
        s_m[i] = p_i;                                    // (p_i parked in s_m until the mu1 chain overwrites it)
        s_ip[i] = __dmul_rn((double)i, p_i);
    }
    __syncwarp();
    if (lane == 0) {
        double q1 = 0.0;
        for (int i = 0; i < 256; ++i) { q1 = __dadd_rn(q1, s_m[i]); s_q[i] = q1; }
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int i = lane * 8 + k;
        const double q1 = s_q[i], q2 = __dsub_rn(1.0, q1);
        const bool skip = fmin(q1, q2) < (double)FLT_EPSILON || fmax(q1, q2) > 1.0 - (double)FLT_EPSILON;
        s_y[i] = skip ? -1.0 : __drcp_rn(q1);
    }
    __syncwarp();
    if (lane == 0) {
        double mu1 = 0.0, qprev = 0.0;
        for (int i = 0; i < 256; ++i) {
            mu1 = __dmul_rn(mu1, qprev);
            qprev = s_q[i];
            const double y = s_y[i];
            if (y >= 0.0) {
                const double a = __dadd_rn(mu1, s_ip[i]);
                const double q = __dmul_rn(a, y);
                const double r = __fma_rn(-q, qprev, a);
                mu1 = __fma_rn(r, y, q);                  // == __ddiv_rn(a, q1)
            }
            s_m[i] = mu1;
        }
    }
    __syncwarp();
    double best = 0.0;
    int bidx = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int i = lane * 8 + k;
        if (s_y[i] < 0.0) continue;
        const double q1 = s_q[i];
        const double q2 = __dsub_rn(1.0, q1), mu1 = s_m[i];
        const double mu2 = __ddiv_rn(__dsub_rn(mu, __dmul_rn(q1, mu1)), q2);
        const double d = __dsub_rn(mu1, mu2);
        const double sigma = __dmul_rn(__dmul_rn(__dmul_rn(q1, q2), d), d);
        if (sigma > best) { best = sigma; bidx = i; }              // first maximum inside the lane
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    __syncwarp();
    return best > 0.0 ? bidx : 0;
}

// one WARP per region (4 regions per CTA): lane l owns bins 8l..8l+7.  Small ROIs have sparse histograms whose
// empty bins create exact ties, so the serial replay is common there; with a warp per region all replays of a
// batch run concurrently instead of queueing behind one another.
__global__ void __launch_bounds__(K4_OTSU_WARPS * 32)
k4_otsu(const qc_seg_region* __restrict__ regs, int R, const unsigned* __restrict__ hist,
        int* __restrict__ cmp, int* __restrict__ out_thr) {
    __shared__ unsigned s_h[K4_OTSU_WARPS][256];
    __shared__ double s_q[K4_OTSU_WARPS][256], s_y[K4_OTSU_WARPS][256], s_ip[K4_OTSU_WARPS][256], s_m[K4_OTSU_WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = blockIdx.x * K4_OTSU_WARPS + warp; r < R; r += gridDim.x * K4_OTSU_WARPS) {
        const qc_seg_region rg = regs[r];
        unsigned h[8];
        unsigned long long w1 = 0, s1 = 0;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            h[k] = hist[(long long)r * 256 + lane * 8 + k];
            s_h[warp][lane * 8 + k] = h[k];
            w1 += h[k];
            s1 += (unsigned long long)h[k] * (unsigned)(lane * 8 + k);
        }
        __syncwarp();
        unsigned long long pw = w1, ps = s1;                   // inclusive prefix over lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long a = __shfl_up_sync(0xffffffffu, pw, o), b = __shfl_up_sync(0xffffffffu, ps, o);
            if (lane >= o) { pw += a; ps += b; }
        }
        const unsigned long long N = __shfl_sync(0xffffffffu, pw, 31), S = __shfl_sync(0xffffffffu, ps, 31);
        unsigned long long cw = pw - w1, cs = ps - s1;        // totals of the bins before this lane's first bin
        double best = -1.0;
        int bidx = 0, flags = 0;       // bit0: some bin not skipped, bit1: a bin sits on the skip boundary
        double sig[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            cw += h[k];
            cs += (unsigned long long)h[k] * (unsigned)(lane * 8 + k);
            sig[k] = -1.0;
            if (N > 0) {
                const double q1 = (double)cw / (double)N, q2 = 1.0 - q1;
                const double mn = fmin(q1, q2), mx = fmax(q1, q2);
                const double eps = (double)FLT_EPSILON;
                if (fabs(mn - eps) <= 1e-3 * eps || fabs(mx - (1.0 - eps)) <= 1e-3 * eps) flags |= 2;
                if (!(mn < eps || mx > 1.0 - eps)) {
                    const double mu1 = (double)cs / (double)cw;
                    const double mu2 = (double)(S - cs) / (double)(N - cw);
                    const double d = mu1 - mu2;
                    sig[k] = q1 * q2 * d * d;
                    flags |= 1;
                    if (sig[k] > best) { best = sig[k]; bidx = lane * 8 + k; }   // first maximum inside the lane
                }
            }
        }
        double b = best;
        int bi = bidx;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, b, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > b || (ob == b && oi < bi)) { b = ob; bi = oi; }
            flags |= __shfl_xor_sync(0xffffffffu, flags, o);
        }
        int nclose = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            nclose += (sig[k] >= 0.0 && lane * 8 + k != bi && sig[k] >= b * (1.0 - 1e-9)) ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nclose += __shfl_xor_sync(0xffffffffu, nclose, o);
        int max_val;
        if (!(flags & 1) && !(flags & 2)) max_val = 0;                              // every split skipped
        else if (b > 0.0 && nclose == 0 && !(flags & 2)) max_val = bi;              // decisive maximum
        else max_val = otsu_replay_cv2(s_h[warp], (long long)rg.w * rg.h, (double)S, s_q[warp], s_y[warp], s_ip[warp],
                                       s_m[warp], lane);
        if (lane == 0) {
            out_thr[r] = max_val;
            int lo, hi;
            if (rg.thr_scale == 1.0) {
                if (rg.polarity) { lo = max_val + 1; hi = 255; }      // cv2 THRESH_BINARY: v > t
                else { lo = 0; hi = max_val; }                        // THRESH_BINARY_INV: v <= t
            } else {
                const double T = __dmul_rn((double)max_val, rg.thr_scale);   // python: threshold * factor
                if (rg.polarity) { lo = (int)floor(T) + 1; hi = 255; }       // v > T
                else { lo = 0; hi = (int)ceil(T) - 1; }                      // v < T
            }
            cmp[2 * r] = lo;
            cmp[2 * r + 1] = hi;
        }
    }
}

// ------------------------------------------------------------------------------------------
// union-find on shared-memory labels (roots = minimum local index)
__device__ __forceinline__ int suf_find(volatile int* L, int x) {
    // path halving: a concurrent reader only ever sees an ancestor, and parents only decrease
    int p = L[x];
    while (p != x) {
        const int gp = L[p];
        if (gp != p) L[x] = gp;
        x = p; p = gp;
    }
    return x;
}
__device__ __forceinline__ void suf_union(int* L, int a, int b) {
    bool done;
    do {
        a = suf_find(L, a);
        b = suf_find(L, b);
        if (a < b) { const int old = atomicMin(L + b, a); done = (old == b); b = old; }
        else if (b < a) { const int old = atomicMin(L + a, b); done = (old == a); a = old; }
        else done = true;
    } while (!done);
}

// 16 bytes from an arbitrary address inside the gray workspace: two aligned 128-bit loads, a 4-way word select and
// one funnel shift per word.  Reads up to 31 bytes past p -- always inside the workspace (rows are contiguous and the
// carve rounds the gray region up to 256 bytes in front of the next array).
__device__ __forceinline__ uint4 ld16_any(const uint8_t* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint4* q = reinterpret_cast<const uint4*>(a & ~(uintptr_t)15);
    const unsigned off = (unsigned)(a & 15);
    const uint4 v0 = q[0];
    if (off == 0) return v0;
    const uint4 v1 = q[1];
    const unsigned ws = off >> 2, bs = (off & 3u) * 8u;
    const unsigned s0 = ws == 0 ? v0.x : ws == 1 ? v0.y : ws == 2 ? v0.z : v0.w;
    const unsigned s1 = ws == 0 ? v0.y : ws == 1 ? v0.z : ws == 2 ? v0.w : v1.x;
    const unsigned s2 = ws == 0 ? v0.z : ws == 1 ? v0.w : ws == 2 ? v1.x : v1.y;
    const unsigned s3 = ws == 0 ? v0.w : ws == 1 ? v1.x : ws == 2 ? v1.y : v1.z;
    const unsigned s4 = ws == 0 ? v1.x : ws == 1 ? v1.y : ws == 2 ? v1.z : v1.w;
    return make_uint4(__funnelshift_r(s0, s1, bs), __funnelshift_r(s1, s2, bs), __funnelshift_r(s2, s3, bs),
                      __funnelshift_r(s3, s4, bs));
}

// threshold -> morphology -> global bit rows (small shared-memory footprint: many CTAs per SM)
__global__ void __launch_bounds__(K4_THREADS)
k4_mask(Img im, const qc_seg_region* __restrict__ regs, int R, int total_bands,
            const int* __restrict__ g_band_off, int iters, const int* __restrict__ cmp,
            const uint8_t* __restrict__ gray_ws, unsigned* __restrict__ bits, int buf_words,
            const unsigned* __restrict__ bits_in) {
    extern __shared__ __align__(16) unsigned sm[];
    unsigned* bufA = sm;
    unsigned* bufB = sm + buf_words;
    __shared__ int s_off[K4_LOC_MAX + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int halo = 4 * iters;
    const int* off = loc_load(s_off, g_band_off, R);
    for (int t = blockIdx.x; t < total_bands; t += gridDim.x) {
        __syncthreads();                 // previous band is done with bufA/bufB
        const int r = loc_find(off, R, t);
        const qc_seg_region rg = regs[r];
        const int wpr = wpr_of(rg);
        const int BR = k4_band_rows(rg.w);
        const int y0 = (t - rg.band_offset) * BR;
        const int y1 = min(y0 + BR, rg.h);
        const int row_lo = max(y0 - halo, 0), row_hi = min(y1 + halo, rg.h);
        const int nrows = row_hi - row_lo;
        if (bits_in) {
            // K7: the pre-morphology mask already exists as bit rows (adaptive threshold / filled mask)
            const unsigned* Bin = bits_in + rg.bits_offset + (long long)row_lo * wpr;
            for (int i = tid; i < nrows * wpr; i += K4_THREADS) bufA[i] = __ldg(Bin + i);
        }
        const int lo = bits_in ? 0 : cmp[2 * r], hi = bits_in ? 0 : cmp[2 * r + 1];
        // ---- threshold into bit rows: a lane turns 16 pixels into 16 bits, lane pairs form words ----
        const uint8_t* gbase;
        long long gpitch;
        if (im.channels == 3) { gbase = gray_ws + rg.out_offset; gpitch = rg.w; }
        else { gbase = im.data + (long long)rg.img * im.img_stride + (long long)rg.y * im.row_stride + rg.x; gpitch = im.row_stride; }
        const int segs = 2 * wpr;                     // 16-pixel segments per row (last ones may be partial/empty)
        const unsigned lo4 = (unsigned)min(lo, 255) * 0x01010101u, hi4 = (unsigned)max(hi, 0) * 0x01010101u;
        const bool none = lo > 255 || hi < 0 || lo > hi;     // empty foreground interval
        // work items = (row, 16-pixel segment), flattened so that narrow regions (a 150-px ROI row is 10 segments)
        // still fill the warp; a warp takes 128 consecutive items per round and issues its four 128-bit loads
        // before the first compare.  segs is even, so the two halves of a 32-pixel word sit in adjacent lanes.
        const int nitems = bits_in ? 0 : nrows * segs;
        const float inv_segs = 1.0f / (float)segs;
        for (int t0 = warp * 128; t0 < nitems; t0 += (K4_THREADS / 32) * 128) {
            uint4 v[4];
            bool fastp[4];
            int yy_[4], sg_[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int t = t0 + 32 * j + lane;
                int yy = (int)(((float)t + 0.5f) * inv_segs);                // t / segs (exact: t < 2^20)
                const int sg = t - yy * segs, x0 = sg * 16;
                yy_[j] = yy; sg_[j] = sg;
                const uint8_t* p = gbase + (long long)(row_lo + yy) * gpitch;
                const bool al = (reinterpret_cast<uintptr_t>(p) & 15) == 0;
                // gray workspace rows (RGB input): any alignment, partial tail segments included (the bits past
                // the row end are cleared below); caller's gray image: aligned full segments only
                fastp[j] = t < nitems && !none && x0 < rg.w && (im.channels == 3 || (al && x0 + 16 <= rg.w));
                if (fastp[j])
                    v[j] = im.channels == 3 ? ld16_any(p + x0) : __ldg(reinterpret_cast<const uint4*>(p + x0));
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (t0 + 32 * j >= nitems) break;            // warp-uniform
                const int t = t0 + 32 * j + lane;
                const int yy = yy_[j], sg = sg_[j], x0 = sg * 16;
                unsigned b16 = 0;
                if (fastp[j]) {
                    // per-byte compares (0xFF where lo <= g <= hi), then 4 mask bytes -> 4 bits with one
                    // multiply: bytes at bit 0/8/16/24 land on bits 24..27 of t * 0x01020408
                    const unsigned w4[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        // the interval is one-sided in practice ([lo,255] light, [0,hi] dark): one compare
                        unsigned mk = hi >= 255 ? __vsetgeu4(w4[q], lo4)
                                    : lo <= 0   ? __vsetleu4(w4[q], hi4)
                                                : (__vsetgeu4(w4[q], lo4) & __vsetleu4(w4[q], hi4));
                        b16 |= ((mk * 0x01020408u) >> 24) << (4 * q);
                    }
                    if (x0 + 16 > rg.w) b16 &= (1u << (rg.w - x0)) - 1u;
                } else if (t < nitems && x0 < rg.w && !none) {
                    const uint8_t* p = gbase + (long long)(row_lo + yy) * gpitch;
                    const int nv = min(16, rg.w - x0);
                    for (int k = 0; k < nv; ++k) {
                        const int g = (int)p[x0 + k];
                        b16 |= (unsigned)(g >= lo && g <= hi) << k;
                    }
                }
                const unsigned hi16 = __shfl_down_sync(0xffffffffu, b16, 1);
                if (!(lane & 1) && t < nitems) bufA[yy * wpr + (sg >> 1)] = b16 | (hi16 << 16);
            }
        }
        __syncthreads();
        // ---- close (dilate^k, erode^k) then open (erode^k, dilate^k), 3x3 cross ----------------
        unsigned* src = bufA;
        unsigned* dst = bufB;
        for (int step = 0; step < 4 * iters; ++step) {
            const int phase = step / iters;              // 0: dilate, 1: erode, 2: erode, 3: dilate
            const bool erode = (phase == 1 || phase == 2);
            const unsigned ob = erode ? 0xFFFFFFFFu : 0u;               // value of everything outside
            const unsigned tail = (rg.w & 31) ? ~((1u << (rg.w & 31)) - 1u) : 0u;   // bits >= w of the last word
            const int nstrip = (nrows + 3) >> 2;
            for (int task = tid; task < nstrip * wpr; task += K4_THREADS) {
                const int st = task / wpr, wx = task - st * wpr;
                const int r0 = st * 4, r1 = min(r0 + 4, nrows);
                const bool lastw = wx == wpr - 1;
                auto ld = [&](int rr, int wxx) -> unsigned {          // buffer row rr, word wxx, border-aware
                    if (rr < 0 || rr >= nrows || wxx < 0 || wxx >= wpr) return ob;
                    unsigned v = src[rr * wpr + wxx];
                    if (erode && wxx == wpr - 1) v |= tail;
                    return v;
                };
                unsigned up = ld(r0 - 1, wx), c = ld(r0, wx);
                for (int rr = r0; rr < r1; ++rr) {
                    const unsigned dn = ld(rr + 1, wx);
                    const unsigned l = ld(rr, wx - 1), rgt = ld(rr, wx + 1);
                    const unsigned xl = (c << 1) | (l >> 31);      // pixel x-1 aligned to x
                    const unsigned xr = (c >> 1) | (rgt << 31);    // pixel x+1 aligned to x
                    unsigned o = erode ? (c & xl & xr & up & dn) : (c | xl | xr | up | dn);
                    if (lastw) o &= ~tail;
                    dst[rr * wpr + wx] = o;
                    up = c; c = dn;
                }
            }
            __syncthreads();
            unsigned* tmp = src; src = dst; dst = tmp;
        }
        // ---- emit the band's own bit rows ---------------------------------------------------------
        for (int wi = tid; wi < (y1 - y0) * wpr; wi += K4_THREADS) {
            const int yy = wi / wpr, wx = wi - yy * wpr;
            bits[rg.bits_offset + (long long)(y0 + yy) * wpr + wx] = src[(y0 - row_lo + yy) * wpr + wx];
        }
    }
}

// band-local CCL on 2x2 blocks (8-connectivity): bit rows of the band -> shared memory (zero-padded by one
// row/word on every side so neighbour reads need no bounds checks), union-find on shared-memory labels.
// Non-empty 32-pixel word cells are compacted first; then every foreground block gets its own thread
// (16 threads per non-empty cell), which keeps warps busy whatever the sparsity.
__global__ void __launch_bounds__(K4_THREADS)
k4_ccl_local(const qc_seg_region* __restrict__ regs, int R, int total_bands, const int* __restrict__ g_band_off,
             const unsigned* __restrict__ bits, int* __restrict__ parent, int band_words, int lab_words) {
    extern __shared__ __align__(16) unsigned sm[];
    unsigned* band_s = sm;                                         // [rows+2][wpr+2], zero border
    int* lab = reinterpret_cast<int*>(sm + band_words);            // [rows/2][wpr*16]
    unsigned short* cell_list = reinterpret_cast<unsigned short*>(lab + lab_words);   // (block row << 8 | word)
    __shared__ int s_off[K4_LOC_MAX + 1];
    __shared__ int s_ncell;
    const int tid = threadIdx.x;
    const int* off = loc_load(s_off, g_band_off, R);
    for (int t = blockIdx.x; t < total_bands; t += gridDim.x) {
        __syncthreads();                 // previous band is done with band_s/lab/cell_list
        const qc_seg_region rg = regs[loc_find(off, R, t)];
        const int wpr = wpr_of(rg), pw = wpr + 2;
        const int BR = k4_band_rows(rg.w);
        const int y0 = (t - rg.band_offset) * BR;
        const int y1 = min(y0 + BR, rg.h);
        const int brows = y1 - y0;
        {
            const unsigned* B = bits + rg.bits_offset + (long long)y0 * wpr;
            for (int i = tid; i < (BR + 2) * pw; i += K4_THREADS) {
                const int rr = i / pw - 1, wx = i - (rr + 1) * pw - 1;      // band row, word (both may be -1 / past the end)
                band_s[i] = (rr >= 0 && rr < brows && wx >= 0 && wx < wpr) ? __ldg(B + rr * wpr + wx) : 0u;
            }
            if (tid == 0) s_ncell = 0;
        }
        __syncthreads();
        // padded accessor: ry in [-1, rows], wx in [-1, wpr]
        auto rowword = [&](int ry, int wx) -> unsigned { return band_s[(ry + 1) * pw + wx + 1]; };
        const int bw = bw_of(rg), bwp = wpr * 16;
        const int by0 = y0 >> 1, nbr = (brows + 1) >> 1;
        // (1) per cell: run-head labels, and compaction of the non-empty cells into cell_list
        const int ncell = nbr * wpr;
        for (int c0 = 0; c0 < ncell; c0 += K4_THREADS) {
            const int ci = c0 + tid;
            const int byy = ci / wpr, wx = ci - byy * wpr;
            unsigned v = 0;
            if (ci < ncell) v = rowword(2 * byy, wx) | rowword(2 * byy + 1, wx);
            if (v) {
                const int base = byy * bwp + wx * 16;
                int head = 0;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    if ((v >> (2 * k)) & 3u) {
                        const bool link = k > 0 && ((v >> (2 * k)) & 1u) && ((v >> (2 * k - 1)) & 1u);
                        if (!link) head = base + k;
                        lab[base + k] = head;
                    }
                }
            }
            const unsigned bm = __ballot_sync(0xffffffffu, v != 0);
            if (bm) {
                int basei = 0;
                const int leader = __ffs(bm) - 1;
                if ((tid & 31) == leader) basei = atomicAdd(&s_ncell, __popc(bm));
                basei = __shfl_sync(0xffffffffu, basei, leader);
                if (v) cell_list[basei + __popc(bm & ((1u << (tid & 31)) - 1u))] = (unsigned short)((byy << 8) | wx);
            }
        }
        __syncthreads();
        const int nne = s_ncell;
        // (2) links: one thread per non-empty cell, one contact per (run, upper run).  Everything is bit arithmetic
        // on the cell's words in the "even-bit" domain (bit 2k <-> block k): FB = foreground blocks, LK = block k
        // linked to block k-1, heads = FB & ~LK.  For a run, the top-row pixels m are dilated by one pixel and
        // intersected with the row above (u): every maximal group of contacted upper blocks that are themselves
        // linked belongs to one upper run and yields a single contact.  visit(it, fn) calls fn(head, target, first)
        // for every contact of every run of cell `it` (targets always have a smaller index than the head).
        auto visit = [&](int it, auto&& fn) {
            const int ce = cell_list[it];
            const int byy = ce >> 8, wx = ce & 255;
            const unsigned r0 = rowword(2 * byy, wx), r1 = rowword(2 * byy + 1, wx);
            const unsigned v = r0 | r1;
            const unsigned FB = (v | (v >> 1)) & 0x55555555u;
            const unsigned LK = v & (v << 1) & 0x55555555u;
            unsigned HB = FB & ~LK;
            const int base = byy * bwp + wx * 16;
            unsigned u = 0, ULK = 0, uL = 0, uR = 0, vuL = 0, vuR = 0, vu = 0;
            if (byy > 0) {                                               // the band above is stitched by k4_border
                u = rowword(2 * byy - 1, wx);
                vu = rowword(2 * byy - 2, wx) | u;
                ULK = vu & (vu << 1) & 0x55555555u;
                uL = rowword(2 * byy - 1, wx - 1) >> 31;
                uR = rowword(2 * byy - 1, wx + 1) & 1u;
                vuL = (rowword(2 * byy - 2, wx - 1) | rowword(2 * byy - 1, wx - 1)) >> 31;
                vuR = (rowword(2 * byy - 2, wx + 1) | rowword(2 * byy - 1, wx + 1)) & 1u;
            }
            while (HB) {
                const int b0 = __ffs(HB) - 1;                            // even bit of the run's first block
                HB &= HB - 1;
                const unsigned below_next = HB ? ((HB & (0u - HB)) - 1u) : 0xFFFFFFFFu;
                const unsigned RB = FB & below_next & ~((1u << b0) - 1u);    // blocks of this run
                const int head = base + (b0 >> 1);
                bool first = true;
                if (b0 == 0 && (v & 1u)) {                               // run continues from the previous word
                    const unsigned vl = (rowword(2 * byy, wx - 1) | rowword(2 * byy + 1, wx - 1)) >> 31;
                    if (vl) { fn(head, head - 1, first); first = false; }
                }
                if (byy == 0) continue;
                const unsigned m = r0 & (RB * 3u);                       // top-row pixels of the run
                if (!m) continue;
                const unsigned c = (m | (m << 1) | (m >> 1)) & u;
                const unsigned CB = (c | (c >> 1)) & 0x55555555u;       // contacted upper blocks
                unsigned CN = CB & ~((CB << 2) & ULK);                   // first block of every linked group
                const int upb = base - bwp;
                while (CN) {
                    const int kk = (__ffs(CN) - 1) >> 1;
                    CN &= CN - 1;
                    fn(head, upb + kk, first);
                    first = false;
                }
                // contacts in the neighbouring words (pixel -1 / pixel 32 of the row above), unless the upper block
                // next to them is contacted too and linked to them
                if ((m & 1u) && uL && !((CB & 1u) && (vu & 1u) && vuL)) { fn(head, upb - 1, first); first = false; }
                if ((m >> 31) && uR && !((CB >> 30) && (vu >> 31) && vuR)) { fn(head, upb + 16, first); first = false; }
            }
        };
        // (2a) every run adopts its first contact as parent with a plain store (targets have smaller indices: a
        // forest) -- in dense regions this already links almost everything, without a single atomic
        for (int it = tid; it < nne; it += K4_THREADS)
            visit(it, [&](int head, int target, bool first) { if (first) lab[head] = target; });
        __syncthreads();
        // (2b) pointer jumping over the run heads until every head points at a root (chains are as long as the band
        // is tall; each round halves them)
        for (;;) {
            int changed = 0;
            for (int it = tid; it < nne; it += K4_THREADS) {
                const int ce = cell_list[it];
                const int byy = ce >> 8, wx = ce & 255;
                const unsigned v = rowword(2 * byy, wx) | rowword(2 * byy + 1, wx);
                unsigned HB = ((v | (v >> 1)) & 0x55555555u) & ~(v & (v << 1) & 0x55555555u);
                const int base = byy * bwp + wx * 16;
                while (HB) {
                    const int h = base + ((__ffs(HB) - 1) >> 1);
                    HB &= HB - 1;
                    const int pp = lab[h];
                    const int g = lab[pp];
                    if (g != pp) { lab[h] = g; changed = 1; }
                }
            }
            if (!__syncthreads_or(changed)) break;
        }
        // (2c) the remaining contacts: unions on trees that are now one or two levels deep; most find equal roots
        for (int it = tid; it < nne; it += K4_THREADS)
            visit(it, [&](int head, int target, bool first) { if (!first) suf_union(lab, head, target); });
        __syncthreads();
        // (3a) compress the run heads (one thread per non-empty cell): every ancestor of a block is a run head, so
        // after this pass a block reaches its root in two hops and the per-block pass below needs no loop
        for (int it = tid; it < nne; it += K4_THREADS) {
            const int ce = cell_list[it];
            const int byy = ce >> 8, wx = ce & 255;
            const unsigned v = rowword(2 * byy, wx) | rowword(2 * byy + 1, wx);
            unsigned HB = ((v | (v >> 1)) & 0x55555555u) & ~(v & (v << 1) & 0x55555555u);
            const int base = byy * bwp + wx * 16;
            while (HB) {
                const int h = base + ((__ffs(HB) - 1) >> 1);
                HB &= HB - 1;
                const int rt = suf_find(lab, h);
                lab[h] = rt;
            }
        }
        __syncthreads();
        // (3b) flatten to global block indices (foreground blocks only)
        const float inv_bwp = 1.0f / (float)bwp;
        for (int it = tid; it < nne * 16; it += K4_THREADS) {
            const int ce = cell_list[it >> 4], k = it & 15;
            const int byy = ce >> 8, wx = ce & 255;
            const unsigned v = rowword(2 * byy, wx) | rowword(2 * byy + 1, wx);
            if (!((v >> (2 * k)) & 3u)) continue;
            const int rt = lab[lab[byy * bwp + wx * 16 + k]];
            const int rby = (int)(((float)rt + 0.5f) * inv_bwp);         // rt / bwp (exact: rt < 2^20)
            parent[rg.blk_offset + (long long)(by0 + byy) * bw + wx * 16 + k] =
                rg.blk_offset + (by0 + rby) * bw + (rt - rby * bwp);
        }
    }
}

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(int* L, int x) {
    int p = __ldcg(L + x);
    while (p != x) {
        const int gp = __ldcg(L + p);
        if (gp != p) __stcg(L + x, gp);          // path halving (any ancestor is a valid parent)
        x = p; p = gp;
    }
    return x;
}
__device__ __forceinline__ void uf_union(int* L, int a, int b) {
    bool done;
    do {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a < b) { const int old = atomicMin(L + b, a); done = (old == b); b = old; }
        else if (b < a) { const int old = atomicMin(L + a, b); done = (old == a); a = old; }
        else done = true;
    } while (!done);
}

// stitches bands together: the first block row of every band (except a region's first) is united
// with the block row above it, in global memory.  One thread per 32-pixel word of that row.
__global__ void __launch_bounds__(K4_THREADS)
k4_border(const qc_seg_region* __restrict__ regs, int R, int total_bands, const int* __restrict__ g_band_off,
          const unsigned* __restrict__ bits, int* __restrict__ parent) {
    __shared__ int s_off[K4_LOC_MAX + 1];
    const int tid = threadIdx.x;
    const int* off = loc_load(s_off, g_band_off, R);
    for (int t = blockIdx.x; t < total_bands; t += gridDim.x) {
        const qc_seg_region rg = regs[loc_find(off, R, t)];
        const int y0 = (t - rg.band_offset) * k4_band_rows(rg.w);
        if (y0 == 0) continue;
        const int wpr = wpr_of(rg), bw = bw_of(rg);
        const int by = y0 >> 1;
        const unsigned* B = bits + rg.bits_offset;
        auto word = [&](int y, int wx) -> unsigned {
            return (wx < 0 || wx >= wpr) ? 0u : __ldg(B + (long long)y * wpr + wx);
        };
        for (int it = tid; it < wpr * 16; it += K4_THREADS) {            // one thread per block of the band's top row
            const int wx = it >> 4, k = it & 15;
            const unsigned r0 = word(y0, wx);
            const unsigned a = (r0 >> (2 * k)) & 1u, b = (r0 >> (2 * k + 1)) & 1u;
            if (!(a | b)) continue;
            const unsigned u = word(y0 - 1, wx);
            const unsigned u0 = (u >> (2 * k)) & 1u, u1 = (u >> (2 * k + 1)) & 1u;
            const int me = rg.blk_offset + by * bw + wx * 16 + k, upb = me - bw;
            if (u0 | u1) {
                // same redundancy rule as the band-local pass: skip when the left neighbour is in my run, is itself
                // connected upwards and the two upper blocks are linked
                bool skip = false;
                if (k > 0 && a && ((r0 >> (2 * k - 1)) & 1u) && u0 && ((u >> (2 * k - 1)) & 1u)) skip = true;
                if (!skip) uf_union(parent, me, upb);
            }
            if (a && !u0) {
                const unsigned um = k > 0 ? (u >> (2 * k - 1)) & 1u : word(y0 - 1, wx - 1) >> 31;
                if (um) uf_union(parent, me, upb - 1);
            }
            if (b && !u1) {
                const unsigned up2 = k < 15 ? (u >> (2 * k + 2)) & 1u : word(y0 - 1, wx + 1) & 1u;
                if (up2) uf_union(parent, me, upb + 1);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Ranking chunk = K4_CHUNK consecutive word cells (block-row major) of one region; one thread per
// cell.  Block-raster order == (cell order, then block order inside the word).
struct Cell {
    unsigned r0, r1;      // the two pixel rows of the cell
    int base;             // global block index of the cell's first block
    bool valid;
};
__device__ __forceinline__ Cell load_cell(const qc_seg_region& rg, const unsigned* __restrict__ bits, int cell) {
    Cell c;
    const int wpr = wpr_of(rg), ncell = bh_of(rg) * wpr;
    c.valid = cell < ncell;
    c.r0 = c.r1 = 0u;
    c.base = 0;
    if (c.valid) {
        const int by = cell / wpr, wx = cell - by * wpr;
        const unsigned* B = bits + rg.bits_offset;
        c.r0 = __ldg(B + (long long)(2 * by) * wpr + wx);
        c.r1 = (2 * by + 1 < rg.h) ? __ldg(B + (long long)(2 * by + 1) * wpr + wx) : 0u;
        c.base = rg.blk_offset + by * bw_of(rg) + wx * 16;
    }
    return c;
}

__global__ void __launch_bounds__(K4_CHUNK)
k4_compress(const qc_seg_region* __restrict__ regs, int R, int total_chunks, const int* __restrict__ g_chunk_off,
            const unsigned* __restrict__ bits, int* __restrict__ parent, int* __restrict__ chunk_count,
            int* __restrict__ zero_area, unsigned long long* __restrict__ zero_n = nullptr,
            long long* __restrict__ zero_x = nullptr) {
    __shared__ int s_off[K4_LOC_MAX + 1];
    __shared__ unsigned s_v[K4_CHUNK];          // r0|r1 of every cell of the chunk
    __shared__ int s_base[K4_CHUNK];
    __shared__ int s_cnt;
    const int tid = threadIdx.x;
    const int* off = loc_load(s_off, g_chunk_off, R);
    for (int c = blockIdx.x; c < total_chunks; c += gridDim.x) {
        const int r = loc_find(off, R, c);
        const qc_seg_region rg = regs[r];
        __syncthreads();
        {
            const Cell ce = load_cell(rg, bits, (c - rg.chunk_offset) * K4_CHUNK + tid);
            s_v[tid] = ce.r0 | ce.r1;
            s_base[tid] = ce.base;
            if (tid == 0) s_cnt = 0;
        }
        __syncthreads();
        // one thread per block (16 per cell): the pointer chases of a cell run in parallel, not in sequence
        int cnt = 0;
        for (int it = tid; it < K4_CHUNK * 16; it += K4_CHUNK) {
            const int cell = it >> 4, k = it & 15;
            if (!((s_v[cell] >> (2 * k)) & 3u)) continue;
            const int g = s_base[cell] + k;
            const int rt = uf_find(parent, g);
            parent[g] = rt;
            if (zero_area) zero_area[g] = 0;
            if (zero_n) { zero_n[g] = 0ull; zero_x[g] = 0ll; }
            cnt += (rt == g);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if ((tid & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
        __syncthreads();
        if (tid == 0) chunk_count[c] = s_cnt;
    }
}

// root -> label.  The number of labels before a chunk = the sum of the root counts of the region's earlier chunks
// (k4_compress wrote them): every CTA adds those up itself -- at most a few hundred integers from L2 -- instead of
// a separate scan kernel; the CTA of a region's last chunk also writes ncomp.
__global__ void __launch_bounds__(K4_CHUNK)
k4_rank(const qc_seg_region* __restrict__ regs, int R, int total_chunks, const int* __restrict__ g_chunk_off,
        const unsigned* __restrict__ bits, const int* __restrict__ parent, const int* __restrict__ chunk_count,
        int* __restrict__ rank, int* __restrict__ out_ncomp, int* __restrict__ stats) {
    __shared__ int s_off[K4_LOC_MAX + 1];
    __shared__ int s_cnt[K4_CHUNK / 32];
    __shared__ int s_pre[K4_CHUNK / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int* off = loc_load(s_off, g_chunk_off, R);
    for (int c = blockIdx.x; c < total_chunks; c += gridDim.x) {
        const int r = loc_find(off, R, c);
        const qc_seg_region rg = regs[r];
        // labels of the region's earlier chunks
        int pre = 0;
        for (int i = rg.chunk_offset + tid; i < c; i += K4_CHUNK) pre += __ldg(chunk_count + i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(0xffffffffu, pre, o);
        const Cell ce = load_cell(rg, bits, (c - rg.chunk_offset) * K4_CHUNK + tid);
        unsigned rootmask = 0;                                   // bit k: block k of the cell is a root
        unsigned m = ((ce.r0 | ce.r1) | ((ce.r0 | ce.r1) >> 1)) & 0x55555555u;
        while (m) {
            const int k = (__ffs(m) - 1) >> 1;
            m &= m - 1;
            if (parent[ce.base + k] == ce.base + k) rootmask |= 1u << k;
        }
        const int cnt = __popc(rootmask);
        int x = cnt;                                             // inclusive scan over the 256 cells
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        __syncthreads();
        if (lane == 31) s_cnt[warp] = x;
        if (lane == 0) s_pre[warp] = pre;
        __syncthreads();
        int prev = 0, base = 0, tot = 0;
#pragma unroll
        for (int k = 0; k < K4_CHUNK / 32; ++k) { prev += (k < warp) ? s_cnt[k] : 0; base += s_pre[k]; tot += s_cnt[k]; }
        int next = base + prev + x - cnt;                        // labels before this cell
        while (rootmask) {
            const int k = __ffs(rootmask) - 1;
            rootmask &= rootmask - 1;
            rank[ce.base + k] = ++next;
            if (next < rg.max_comp) {                            // the label's stats row: (minx, miny, maxx, maxy, area)
                int* st = stats + ((long long)rg.stats_offset + next) * 5;
                st[0] = INT_MAX; st[1] = INT_MAX; st[2] = -1; st[3] = -1; st[4] = 0;
            }
        }
        const int ncell = bh_of(rg) * wpr_of(rg);
        if (tid == 0 && c == rg.chunk_offset + (ncell + K4_CHUNK - 1) / K4_CHUNK - 1)
            out_ncomp[r] = base + tot + 1;                       // labels incl. background, like cv2
    }
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(K4_THREADS)
k4_stats(const qc_seg_region* __restrict__ regs, int R, int total_bands, const int* __restrict__ g_band_off,
         const unsigned* __restrict__ bits, int* __restrict__ parent, const int* __restrict__ rank,
         int* __restrict__ stats) {
    __shared__ int s_off[K4_LOC_MAX + 1];
    const int tid = threadIdx.x;
    const int* off = loc_load(s_off, g_band_off, R);
    for (int t = blockIdx.x; t < total_bands; t += gridDim.x) {
        const qc_seg_region rg = regs[loc_find(off, R, t)];
        const int wpr = wpr_of(rg), bw = bw_of(rg);
        const int BR = k4_band_rows(rg.w);
        const int y0 = (t - rg.band_offset) * BR;
        const int y1 = min(y0 + BR, rg.h);
        const int by0 = y0 >> 1, by1 = (y1 + 1) >> 1;
        const unsigned* B = bits + rg.bits_offset;
        for (int ci = tid; ci < (by1 - by0) * wpr; ci += K4_THREADS) {
            const int byy = ci / wpr, wx = ci - byy * wpr;
            const int by = by0 + byy;
            const unsigned r0 = __ldg(B + (long long)(2 * by) * wpr + wx);
            const unsigned r1 = (2 * by + 1 < rg.h) ? __ldg(B + (long long)(2 * by + 1) * wpr + wx) : 0u;
            const unsigned v = r0 | r1;
            if (!v) continue;
            int* prow = parent + rg.blk_offset + (long long)by * bw + wx * 16;
            int cur = -1, area = 0, mnx = INT_MAX, mxx = -1, mny = INT_MAX, mxy = -1;
            auto flush = [&]() {
                if (cur > 0 && cur < rg.max_comp) {
                    int* s = stats + ((long long)rg.stats_offset + cur) * 5;
                    atomicMin(s + 0, mnx); atomicMin(s + 1, mny); atomicMax(s + 2, mxx); atomicMax(s + 3, mxy);
                    atomicAdd(s + 4, area);
                }
            };
            // the two dependent loads per block (parent, then rank of the root) are issued for all 16 blocks first
            int labs[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) labs[k] = ((v >> (2 * k)) & 3u) ? prow[k] : -1;
#pragma unroll
            for (int k = 0; k < 16; ++k) labs[k] = labs[k] >= 0 ? rank[labs[k]] : 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const unsigned a = (r0 >> (2 * k)) & 1u, b = (r0 >> (2 * k + 1)) & 1u;
                const unsigned c = (r1 >> (2 * k)) & 1u, d = (r1 >> (2 * k + 1)) & 1u;
                if (!(a | b | c | d)) continue;
                const int lab = labs[k];
                prow[k] = lab;                      // parent -> final label, in place
                if (lab != cur) { flush(); cur = lab; area = 0; mnx = INT_MAX; mxx = -1; mny = INT_MAX; mxy = -1; }
                const int bx = wx * 16 + k;
                area += (int)(a + b + c + d);
                mnx = min(mnx, 2 * bx + ((a | c) ? 0 : 1));
                mxx = max(mxx, 2 * bx + ((b | d) ? 1 : 0));
                mny = min(mny, 2 * by + ((a | b) ? 0 : 1));
                mxy = max(mxy, 2 * by + ((c | d) ? 1 : 0));
            }
            flush();
        }
    }
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(K4_THREADS)
k4_final(const qc_seg_region* __restrict__ regs, int R, int total_bands, const int* __restrict__ g_band_off,
         const unsigned* __restrict__ bits,
         const int* __restrict__ parent, int* __restrict__ stats, const int* __restrict__ ncomp,
         uint8_t* __restrict__ out_mask, int* __restrict__ out_labels, long long* __restrict__ out_area) {
    __shared__ int s_off[K4_LOC_MAX + 1];
    __shared__ int s_area[K4_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int* off = loc_load(s_off, g_band_off, R);
    for (int t = blockIdx.x; t < total_bands; t += gridDim.x) {
        __syncthreads();                 // s_area of the previous band has been consumed
        const int s_r = loc_find(off, R, t);
        const qc_seg_region rg = regs[s_r];
        const int wpr = wpr_of(rg), bw = bw_of(rg);
        const int BR = k4_band_rows(rg.w);
        const int y0 = (t - rg.band_offset) * BR;
        const int y1 = min(y0 + BR, rg.h);
        const unsigned* B = bits + rg.bits_offset;
        const int* st = stats + (long long)rg.stats_offset * 5;
        const int* P = parent + rg.blk_offset;
        int kept = 0;
        // No pixel outputs wanted and every component has a stats row: the kept area is the sum of the recorded
        // areas that pass the min-area filter -- one CTA per region sums <= max_comp rows instead of every band
        // re-reading its pixels (the detection-ROI mode of the pipeline).
        const bool light = !out_mask && !out_labels && ncomp[s_r] <= rg.max_comp;
        if (light) {
            if (y0 != 0) continue;
            const int rows = ncomp[s_r];
            for (int k = 1 + tid; k < rows; k += K4_THREADS) {
                const int a = st[(long long)k * 5 + 4];
                kept += a >= rg.min_area ? a : 0;
            }
        }
        // warp per pixel row, 4 consecutive pixels per lane; background quads (the vast majority) take a
        // branch-light path that only stores zeros
        const bool al = ((rg.out_offset & 3) == 0) && ((rg.w & 3) == 0);
        for (int y = y0 + warp; y < y1 && !light; y += K4_THREADS / 32) {
            const unsigned* Brow = B + (long long)y * wpr;
            const int* prow_y = P + (long long)(y >> 1) * bw;
            const long long orow = rg.out_offset + (long long)y * rg.w;
            for (int x0 = lane * 4; x0 < rg.w; x0 += 128) {
                const unsigned nib = (__ldg(Brow + (x0 >> 5)) >> (x0 & 31)) & 15u;
                int l0 = 0, l1 = 0, l2 = 0, l3 = 0;
                if (nib) {
                    int l01 = 0, l23 = 0;
                    if (nib & 3u) l01 = prow_y[x0 >> 1];
                    if (nib & 12u) l23 = prow_y[(x0 >> 1) + 1];
                    // min-area filter (components beyond max_comp have no recorded area: kept)
                    if (l01 > 0 && l01 < rg.max_comp && st[(long long)l01 * 5 + 4] < rg.min_area) l01 = 0;
                    if (l23 > 0 && l23 < rg.max_comp && st[(long long)l23 * 5 + 4] < rg.min_area) l23 = 0;
                    l0 = (nib & 1u) ? l01 : 0; l1 = (nib & 2u) ? l01 : 0;
                    l2 = (nib & 4u) ? l23 : 0; l3 = (nib & 8u) ? l23 : 0;
                    kept += (l0 > 0) + (l1 > 0) + (l2 > 0) + (l3 > 0);
                }
                const long long o = orow + x0;
                if (al) {
                    if (out_labels) qc_st_cs(out_labels + o, make_uint4((unsigned)l0, (unsigned)l1, (unsigned)l2, (unsigned)l3));
                    if (out_mask)
                        *reinterpret_cast<unsigned*>(out_mask + o) = (l0 > 0 ? 0xFFu : 0u) | (l1 > 0 ? 0xFF00u : 0u) |
                                                                      (l2 > 0 ? 0xFF0000u : 0u) | (l3 > 0 ? 0xFF000000u : 0u);
                } else {
                    const int lab[4] = {l0, l1, l2, l3};
                    const int nvalid = min(4, rg.w - x0);
                    for (int j = 0; j < nvalid; ++j) {
                        if (out_labels) out_labels[o + j] = lab[j];
                        if (out_mask) out_mask[o + j] = lab[j] > 0 ? 255 : 0;
                    }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, o);
        if (lane == 0) s_area[warp] = kept;
        __syncthreads();
        if (tid == 0) {
            long long s = 0;
            for (int k = 0; k < K4_THREADS / 32; ++k) s += s_area[k];
            if (s) atomicAdd(reinterpret_cast<unsigned long long*>(out_area + s_r), (unsigned long long)s);
        }
        // first band of each region converts (minx,miny,maxx,maxy) -> (x,y,w,h); only columns 2,3
        // change, nobody else reads them
        if (y0 == 0) {
            const int rows = min(ncomp[s_r], rg.max_comp);
            int* sw = stats + (long long)rg.stats_offset * 5;
            for (int k = tid; k < rows; k += K4_THREADS) {
                if (k == 0 || sw[k * 5 + 4] == 0) { sw[k * 5 + 0] = sw[k * 5 + 1] = sw[k * 5 + 2] = sw[k * 5 + 3] = 0; continue; }
                sw[k * 5 + 2] = sw[k * 5 + 2] - sw[k * 5 + 0] + 1;
                sw[k * 5 + 3] = sw[k * 5 + 3] - sw[k * 5 + 1] + 1;
            }
        }
    }
}

#include "k7_refseg.cuh"

}  // namespace

// ------------------------------------------------------------------------------------------
extern "C" int qc_segment_plan(qc_seg_region* regs, int num_regions, qc_seg_params* p) {
    if (!regs || !p || num_regions < 0) return QC_ERR_INVALID;
    long long px = 0, bits = 0, blocks = 0, stats = 0;
    long long bands = 0, chunks = 0;
    int mw = 0, mh = 0;
    for (int i = 0; i < num_regions; ++i) {
        qc_seg_region& r = regs[i];
        if (r.w <= 0 || r.h <= 0 || r.x < 0 || r.y < 0 || r.img < 0 || r.img >= p->n || r.x + r.w > p->w ||
            r.y + r.h > p->h || r.max_comp < 1)
            return QC_ERR_INVALID;
        const long long wpr = (r.w + 31) / 32, bw = (r.w + 1) / 2, bh = (r.h + 1) / 2;
        r.out_offset = px;
        r.stats_offset = (int32_t)stats;
        r.band_offset = (int32_t)bands;
        r.chunk_offset = (int32_t)chunks;
        r.bits_offset = (int32_t)bits;
        r.blk_offset = (int32_t)blocks;
        px += (long long)r.w * r.h;
        px = (px + 15) & ~15ll;                 // keep every region's outputs 16-byte aligned
        stats += r.max_comp;
        bands += (r.h + k4_band_rows(r.w) - 1) / k4_band_rows(r.w);
        chunks += (wpr * bh + K4_CHUNK - 1) / K4_CHUNK;
        bits += wpr * r.h;
        blocks += bw * bh;
        mw = std::max(mw, r.w);
        mh = std::max(mh, r.h);
        if (blocks > INT_MAX || bits > INT_MAX || stats > INT_MAX) return QC_ERR_UNSUPPORTED;
    }
    p->num_regions = num_regions;
    p->total_pixels = px;
    p->total_bits_words = (int32_t)bits;
    p->total_blocks = (int32_t)blocks;
    p->total_stats_rows = (int32_t)stats;
    p->total_bands = (int32_t)bands;
    p->total_chunks = (int32_t)chunks;
    p->max_region_w = mw;
    p->max_region_h = mh;
    return QC_OK;
}

extern "C" size_t qc_segment_workspace_bytes(const qc_seg_params* p) {
    if (!p || p->num_regions <= 0) return 0;
    return carve(nullptr, p).total;
}

extern "C" size_t qc_segment_ref_workspace_bytes(const qc_seg_params* p, int method) {
    if (!p || p->num_regions <= 0 || method < QC_SEG_COMPOSITE || method > QC_SEG_REF_ADAPTIVE) return 0;
    return carve(nullptr, p, method).total;
}

namespace {

struct SegOut {
    uint8_t* mask; int32_t* labels; int32_t* ncomp; int32_t* stats; int32_t* thr; int64_t* area; int64_t* sums;
    int64_t* contours;
};
struct SegAdaptive {
    const float* akern; int pitch, max_block, c;
};

int seg_run(qc_ctx* ctx, int method, const uint8_t* imgs, const qc_seg_params* p, const qc_seg_region* regions,
            const SegAdaptive& ad, void* workspace, size_t workspace_bytes, const SegOut& o, void* stream) {
    const char* fn = method == QC_SEG_COMPOSITE ? "qc_segment" : "qc_segment_ref";
    if (!ctx) return QC_ERR_INVALID;
    qc_device_guard dev_guard(ctx);
    QC_ENTER(ctx, stream);
    if (!p) QC_FAIL(ctx, QC_ERR_INVALID, "%s: null params", fn);
    if (method < QC_SEG_COMPOSITE || method > QC_SEG_REF_ADAPTIVE) QC_FAIL(ctx, QC_ERR_INVALID, "%s: bad method %d", fn, method);
    if (p->num_regions == 0) return QC_OK;
    if (!imgs || !regions || !o.ncomp || !o.stats || !o.thr || !o.area)
        QC_FAIL(ctx, QC_ERR_INVALID, "%s: null argument", fn);
    if ((p->channels != 1 && p->channels != 3) || p->morph_iters < 0 || p->morph_iters > 4 || p->total_bands <= 0 ||
        p->row_stride < p->w * p->channels)
        QC_FAIL(ctx, QC_ERR_INVALID, "%s: bad params (run qc_segment_plan first)", fn);
    SegWs w = carve(workspace, p, method);
    if (!workspace || workspace_bytes < w.total) QC_FAIL(ctx, QC_ERR_WORKSPACE, "workspace %zu < %zu", workspace_bytes, w.total);
    cudaStream_t st = (cudaStream_t)stream;
    const int R = p->num_regions;
    Img im{imgs, p->img_stride, p->row_stride, p->channels, p->swap_rb};
    const int sms = ctx->sm_count;
    const int band_grid = std::min(p->total_bands, sms * 8);
    const int chunk_grid = std::min(p->total_chunks, sms * 8);

    const int halo = 4 * p->morph_iters;
    const int wpr_max = (p->max_region_w + 31) / 32;
    if (wpr_max > 256)      // k4_ccl_local packs (block row, word) of a cell into 8+8 bits
        QC_FAIL(ctx, QC_ERR_UNSUPPORTED, "%s: region width %d > 8192", fn, p->max_region_w);
    // conservative bounds over every region: rows x words <= RW whatever the region width
    const int RW = std::max(K4_BAND_RW, K4_MIN_ROWS * wpr_max);
    const int buf_words = RW + 2 * halo * wpr_max;
    const size_t smem_mask = (size_t)buf_words * 4 * 2;
    const int band_words = (RW + 2 * 512 + 2 * wpr_max + 4 + 3) & ~3;
    const int lab_words = 8 * RW;
    const size_t smem_ccl = (size_t)band_words * 4 + (size_t)lab_words * 4 + (size_t)(RW / 2) * 2;
    const size_t smem_need = std::max(smem_mask, smem_ccl);
    if ((long long)smem_need + (K4_LOC_MAX + 1) * 4 + 1024 > ctx->max_smem_optin)
        QC_FAIL(ctx, QC_ERR_UNSUPPORTED, "%s: region width %d needs %zu B of shared memory (max %d)", fn,
                p->max_region_w, smem_need, ctx->max_smem_optin);
    int occ = 1;

    // stage helpers -----------------------------------------------------------------------------------------
    auto morph = [&](const unsigned* bits_in, unsigned* bits_out) -> int {          // (threshold |) close, open
        if (int rc = qc_prepare(ctx, k4_mask, (int)smem_mask, K4_THREADS, &occ)) return rc;
        k4_mask<<<std::min(p->total_bands, sms * std::max(occ, 1)), K4_THREADS, smem_mask, st>>>(
            im, regions, R, p->total_bands, w.band_off, p->morph_iters, w.cmp, w.gray, bits_out, buf_words, bits_in);
        QC_LAUNCHED_AS(ctx, "k4_mask", st);
        return QC_OK;
    };
    auto unionfind = [&](const unsigned* bits, int* zero_area, bool zero_k8 = false) -> int {   // parent[g] = root afterwards
        if (int rc = qc_prepare(ctx, k4_ccl_local, (int)smem_ccl, K4_THREADS, &occ)) return rc;
        k4_ccl_local<<<std::min(p->total_bands, sms * std::max(occ, 1)), K4_THREADS, smem_ccl, st>>>(
            regions, R, p->total_bands, w.band_off, bits, w.parent, band_words, lab_words);
        QC_LAUNCHED_AS(ctx, "k4_ccl_local", st);
        k4_border<<<band_grid, K4_THREADS, 0, st>>>(regions, R, p->total_bands, w.band_off, bits, w.parent);
        QC_LAUNCHED_AS(ctx, "k4_border", st);
        k4_compress<<<chunk_grid, K4_CHUNK, 0, st>>>(regions, R, p->total_chunks, w.chunk_off, bits, w.parent,
                                                     w.chunk_count, zero_area, zero_k8 ? w.acc_n : nullptr,
                                                     zero_k8 ? w.acc_x : nullptr);
        QC_LAUNCHED_AS(ctx, "k4_compress", st);
        return QC_OK;
    };
    int queue_used = 0;
    QC_CUDA(ctx, cudaMemsetAsync(w.queue, 0, 64, st));
    auto fill_holes = [&](const unsigned* bits_in, unsigned* bits_out, unsigned* free_ws) -> int {
        if (int rc = qc_prepare(ctx, k7_fill, K7_FILL_SMEM_WORDS * 4, K7_FILL_THREADS, &occ)) return rc;
        if (queue_used >= 16) QC_FAIL(ctx, QC_ERR_INVALID, "%s: out of work counters", fn);
        k7_fill<<<std::min(R, sms * std::max(occ, 1)), K7_FILL_THREADS, K7_FILL_SMEM_WORDS * 4, st>>>(regions, R, bits_in, bits_out, free_ws,
                                                                                                    w.queue + queue_used++);
        QC_LAUNCHED_AS(ctx, "k7_fill", st);
        return QC_OK;
    };
    // K8: contour statistics of the final mask `fbits` (bit rows); scratch: two bit buffers, the union-find arrays
    auto contour_stats = [&](const unsigned* fbits, unsigned* filled, unsigned* free_ws) -> int {
        cudaMemsetAsync(o.contours, 0, (size_t)R * 4 * sizeof(int64_t), st);
        if (int rc = fill_holes(fbits, filled, free_ws)) return rc;
        if (int rc = unionfind(filled, nullptr, true)) return rc;
        k8_moves<<<band_grid, K4_THREADS, 0, st>>>(regions, R, p->total_bands, w.band_off, filled, w.parent, w.acc_n, w.acc_x);
        QC_LAUNCHED_AS(ctx, "k8_moves", st);
        k8_reduce<<<band_grid, K4_THREADS, 0, st>>>(regions, R, p->total_bands, w.band_off, filled, w.parent, w.acc_n, w.acc_x,
                                                    (long long*)o.contours);
        QC_LAUNCHED_AS(ctx, "k8_reduce", st);
        return QC_OK;
    };
    auto label_tail = [&](const unsigned* bits) -> int {                            // cv2 numbering, stats, outputs
        k4_rank<<<chunk_grid, K4_CHUNK, 0, st>>>(regions, R, p->total_chunks, w.chunk_off, bits, w.parent, w.chunk_count,
                                                 w.rank, o.ncomp, o.stats);
        QC_LAUNCHED_AS(ctx, "k4_rank", st);
        k4_stats<<<band_grid, K4_THREADS, 0, st>>>(regions, R, p->total_bands, w.band_off, bits, w.parent, w.rank,
                                                   o.stats);
        QC_LAUNCHED_AS(ctx, "k4_stats", st);
        k4_final<<<band_grid, K4_THREADS, 0, st>>>(regions, R, p->total_bands, w.band_off, bits, w.parent, o.stats,
                                                   o.ncomp, o.mask, o.labels, (long long*)o.area);
        QC_LAUNCHED_AS(ctx, "k4_final", st);
        return QC_OK;
    };

    // stats rows: the labelling tail (composite, adaptive) initialises the rows of the labels it hands out; the threshold
    // method's two rows per region are written by k7_emit / k7_finish from initialised sentinels
    const int init_rows = method == QC_SEG_REF_THRESHOLD ? p->total_stats_rows : 0;
    k4_init<<<std::min(sms * 4, std::max(1, (std::max(init_rows, R * 64) + 255) / 256)), 256, 0, st>>>(
        regions, R, w.hist, o.stats, (long long*)o.area, init_rows, w.band_off, w.chunk_off,
        p->total_bands, p->total_chunks);
    QC_LAUNCHED_AS(ctx, "k4_init", st);
    if (p->channels == 3) {
        k4_gray<<<std::min(p->total_bands, sms * 8), K4_THREADS, 0, st>>>(im, regions, R, p->total_bands, w.band_off, w.gray);
        QC_LAUNCHED_AS(ctx, "k4_gray", st);
    } else if (method != QC_SEG_COMPOSITE) {
        k7_pack_gray<<<std::min(p->total_bands, sms * 8), K4_THREADS, 0, st>>>(im, regions, R, p->total_bands, w.band_off, w.gray);
        QC_LAUNCHED_AS(ctx, "k7_pack_gray", st);
    }
    const uint8_t* gray7 = w.gray;          // K7 reads the region-packed gray copy whatever the input layout

    if (method == QC_SEG_COMPOSITE) {
        if (int rc = qc_prepare(ctx, k4_hist, K4_HIST_SMEM, K4_HIST_THREADS, nullptr)) return rc;
        const int nw = K4_HIST_THREADS / 32;
        const int items = p->total_bands * K4_HIST_SPLIT;
        k4_hist<<<std::min((items + nw - 1) / nw, sms), K4_HIST_THREADS, K4_HIST_SMEM, st>>>(
            im, regions, R, p->total_bands, w.band_off, w.hist, w.gray);
        QC_LAUNCHED_AS(ctx, "k4_hist", st);
        k4_otsu<<<std::min((R + K4_OTSU_WARPS - 1) / K4_OTSU_WARPS, sms * 8), K4_OTSU_WARPS * 32, 0, st>>>(
            regions, R, w.hist, w.cmp, o.thr);
        QC_LAUNCHED_AS(ctx, "k4_otsu", st);
        if (int rc = morph(nullptr, w.bits)) return rc;
        if (int rc = unionfind(w.bits, nullptr)) return rc;
        return label_tail(w.bits);
    }

    if (method == QC_SEG_REF_ADAPTIVE) {
        if (!ad.akern || ad.pitch <= 0) QC_FAIL(ctx, QC_ERR_INVALID, "%s: adaptive method needs the Gaussian kernel table", fn);
        const int bmax = k7_block_size(p->max_region_w, p->max_region_h);       // >= every region's block size
        if (bmax > ad.max_block || bmax > ad.pitch)
            QC_FAIL(ctx, QC_ERR_INVALID, "%s: kernel table covers block <= %d, regions need %d", fn, ad.max_block, bmax);
        const int pmax = bmax / 2;
        // blocks up to 33: the register sliding-window kernel; larger blocks (ROIs with a short side >= 340 px): the
        // generic tile kernel, launched only when the region table can hold such a region
        k7_adaptive_small<<<band_grid, K4_THREADS, 0, st>>>(im, regions, R, p->total_bands, w.band_off, gray7, ad.akern,
                                                            ad.pitch, ad.c, w.bits2);
        QC_LAUNCHED_AS(ctx, "k7_adaptive_small", st);
        if (pmax > K7A_P) {
            const size_t smem_ad = (size_t)((2 * pmax + 1 + 3) & ~3) * 4 + (size_t)(32 + 2 * pmax) * 33 * 4 +
                                   (size_t)(32 + 2 * pmax) * (32 + 2 * pmax);
            if ((long long)smem_ad + (K4_LOC_MAX + 1) * 4 + 1024 > ctx->max_smem_optin)
                QC_FAIL(ctx, QC_ERR_UNSUPPORTED, "%s: adaptive block %d needs %zu B of shared memory", fn, bmax, smem_ad);
            if (int rc = qc_prepare(ctx, k7_adaptive, (int)smem_ad, K4_THREADS, &occ)) return rc;
            k7_adaptive<<<std::min(p->total_bands, sms * std::max(occ, 1)), K4_THREADS, smem_ad, st>>>(
                im, regions, R, p->total_bands, w.band_off, gray7, ad.akern, ad.pitch, ad.c, w.bits2, pmax, K7A_P);
            QC_LAUNCHED_AS(ctx, "k7_adaptive", st);
        }
        k7_block_out<<<(R + 255) / 256, 256, 0, st>>>(regions, R, o.thr, (long long*)o.sums);
        QC_LAUNCHED_AS(ctx, "k7_block_out", st);
        if (int rc = morph(w.bits2, w.bits)) return rc;
        if (int rc = unionfind(w.bits, nullptr)) return rc;
        if (int rc = label_tail(w.bits)) return rc;
        if (o.contours || o.sums) {
            // kept mask (labels + areas are final) -> bits2
            k8_kept_bits<<<band_grid, K4_THREADS, 0, st>>>(regions, R, p->total_bands, w.band_off, w.bits, w.parent, o.stats, w.bits2);
            QC_LAUNCHED_AS(ctx, "k8_kept_bits", st);
        }
        if (o.contours)        // filled -> bits3; free rows of oversized regions -> bits
            if (int rc = contour_stats(w.bits2, w.bits3, w.bits)) return rc;
        if (o.sums) {
            // K9: Canny(50, 150) edges and their overlap with the kept mask (:557-578); candidates -> bits, strong -> bits3
            k9_canny_nms<<<band_grid, K4_THREADS, 0, st>>>(regions, R, p->total_bands, w.band_off, gray7, 50, 150, w.bits, w.bits3);
            QC_LAUNCHED_AS(ctx, "k9_canny_nms", st);
            if (int rc = unionfind(w.bits, w.rank)) return rc;                 // w.rank = "component holds a strong pixel"
            k9_mark<<<band_grid, K4_THREADS, 0, st>>>(regions, R, p->total_bands, w.band_off, w.bits3, w.parent, w.rank);
            QC_LAUNCHED_AS(ctx, "k9_mark", st);
            k9_overlap<<<band_grid, K4_THREADS, 0, st>>>(regions, R, p->total_bands, w.band_off, w.bits, w.bits2, w.parent, w.rank,
                                                         (long long*)o.sums);
            QC_LAUNCHED_AS(ctx, "k9_overlap", st);
        }
        return QC_OK;
    }

    // QC_SEG_REF_THRESHOLD
    k7_init<<<std::max(1, std::min(sms, (R * 4 + 255) / 256)), 256, 0, st>>>(R, w.minmax, (long long*)o.sums);
    QC_LAUNCHED_AS(ctx, "k7_init", st);
    k7_blur<<<band_grid, K4_THREADS, 0, st>>>(im, regions, R, p->total_bands, w.band_off, gray7, w.blurred, w.minmax);
    QC_LAUNCHED_AS(ctx, "k7_blur", st);
    k7_hist<<<band_grid, K4_THREADS, 0, st>>>(regions, R, p->total_bands, w.band_off, w.blurred, w.minmax, w.hist);
    QC_LAUNCHED_AS(ctx, "k7_hist", st);
    k7_otsu<<<std::min((R + K7_OTSU_WARPS - 1) / K7_OTSU_WARPS, sms * 8), K7_OTSU_WARPS * 32, 0, st>>>(
        regions, R, w.hist, w.minmax, w.thr_adj, o.thr);
    QC_LAUNCHED_AS(ctx, "k7_otsu", st);
    k7_cmp<<<band_grid, K4_THREADS, 0, st>>>(regions, R, p->total_bands, w.band_off, w.blurred, w.thr_adj, w.bits);
    QC_LAUNCHED_AS(ctx, "k7_cmp", st);
    if (int rc = unionfind(w.bits, w.rank)) return rc;                             // w.rank doubles as the area table
    k7_area<<<band_grid, K4_THREADS, 0, st>>>(regions, R, p->total_bands, w.band_off, w.bits, w.parent, w.rank);
    QC_LAUNCHED_AS(ctx, "k7_area", st);
    k7_filter<<<band_grid, K4_THREADS, 0, st>>>(regions, R, p->total_bands, w.band_off, w.bits, w.parent, w.rank, w.bits2);
    QC_LAUNCHED_AS(ctx, "k7_filter", st);
    // (regions too large for shared memory flood in place in bits3, with their free rows parked in bits -- free until
    // the morphology pass rewrites it)
    if (int rc = fill_holes(w.bits2, w.bits3, w.bits)) return rc;
    if (int rc = morph(w.bits3, w.bits)) return rc;
    k7_emit<<<band_grid, K4_THREADS, 0, st>>>(im, regions, R, p->total_bands, w.band_off, gray7, w.bits, o.mask, o.labels,
                                              o.stats, (long long*)o.area, (long long*)o.sums);
    QC_LAUNCHED_AS(ctx, "k7_emit", st);
    k7_finish<<<(R + 255) / 256, 256, 0, st>>>(regions, R, o.stats, (const long long*)o.area, o.ncomp, (long long*)o.sums);
    QC_LAUNCHED_AS(ctx, "k7_finish", st);
    if (o.contours) return contour_stats(w.bits, w.bits3, w.bits2);      // final mask = bits; bits2 / bits3 are free now
    return QC_OK;
}

}  // namespace

extern "C" int qc_segment(qc_ctx* ctx, const uint8_t* imgs, const qc_seg_params* p, const qc_seg_region* regions,
                          void* workspace, size_t workspace_bytes, uint8_t* out_mask, int32_t* out_labels,
                          int32_t* out_ncomp, int32_t* out_stats, int32_t* out_thr, int64_t* out_area, void* stream) {
    return seg_run(ctx, QC_SEG_COMPOSITE, imgs, p, regions, SegAdaptive{nullptr, 0, 0, 0}, workspace, workspace_bytes,
                   SegOut{out_mask, out_labels, out_ncomp, out_stats, out_thr, out_area, nullptr, nullptr}, stream);
}

extern "C" int qc_segment_ref(qc_ctx* ctx, int method, const uint8_t* imgs, const qc_seg_params* p,
                              const qc_seg_region* regions, const float* akern, int akern_pitch, int akern_max_block,
                              int adaptive_c, void* workspace, size_t workspace_bytes, uint8_t* out_mask,
                              int32_t* out_labels, int32_t* out_ncomp, int32_t* out_stats, int32_t* out_thr,
                              int64_t* out_area, int64_t* out_sums, int64_t* out_contours, void* stream) {
    return seg_run(ctx, method, imgs, p, regions, SegAdaptive{akern, akern_pitch, akern_max_block, adaptive_c},
                   workspace, workspace_bytes,
                   SegOut{out_mask, out_labels, out_ncomp, out_stats, out_thr, out_area, out_sums, out_contours}, stream);
}
