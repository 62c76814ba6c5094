// k2_yolo.cu -- K2: YOLOv8 head post-processing on device.
//
// Replaces what runs inside `self.model(img, conf=, iou=)` at models/yolo_model.py:105-110
// after the conv net (ultralytics, restated in SURVEY.md App. A): Detect._inference (DFL
// softmax.arange, dist2bbox, *stride, sigmoid), ops.non_max_suppression (conf filter,
// class max, xywh->xyxy, class offset 7680, torchvision.ops.nms, [:max_det]) and
// ops.scale_boxes.  Output rows are what models/yolo_model.py:133-135 reads.
//
// Three kernels per call:
//   k2_filter   streams ONLY the nc class planes (5 of 69 channels): per anchor sigmoid, first
//               max over classes, `score > (float)conf`, warp-ballot aggregated compaction into
//               a per-image candidate list (key = score-desc | anchor | class).
//   k2_decode   4 threads per candidate (one per box side): gathers the 16 DFL logits of its
//               side, softmax, expectation, then the reference's exact float32 op order
//               anchor-/+dist -> (cx,cy,w,h)*stride -> xy-/+wh/2.
//   k2_nms      one CTA per image: bitonic sort of (key,slot) by (score desc, anchor asc) ==
//               torch's stable descending sort; class offset boxes; greedy NMS in 64-box
//               chunks with a shared-memory 64x64 IoU bitmask resolved by one thread and the
//               chunk's kept boxes applied to all later boxes by the whole CTA; stops at
//               max_det; scale_boxes epilogue.  IoU math is torchvision's CPU kernel verbatim
//               (float32, no FMA, compared against the double threshold).
// Streaming phase is HBM-bound; sort/NMS phase is latency-bound (reported separately).
#include "common.cuh"
#include <math.h>
#include <algorithm>

namespace {

constexpr int K2_NMS_THREADS = 1024;
constexpr int K2_SORT_SMEM_MAX = 4096;   // candidates sorted / held in shared memory
constexpr int K2_REMOVED_MAX = 32768;    // removed[] byte flags in shared memory

struct LevelTable {
    const void* ptr[QC_MAX_LEVELS];
    long long batch_stride[QC_MAX_LEVELS];
    long long chan_stride[QC_MAX_LEVELS];
    int h[QC_MAX_LEVELS], w[QC_MAX_LEVELS];
    int start[QC_MAX_LEVELS + 1];   // first anchor of each level
    float stride[QC_MAX_LEVELS];
    int num_levels;
    int A;
};

__device__ __forceinline__ uint32_t f2ord_desc(float f) {
    // monotone map float -> uint32 such that larger floats give SMALLER keys
    uint32_t u = __float_as_uint(f);
    u ^= (u & 0x80000000u) ? 0xFFFFFFFFu : 0x80000000u;   // ascending order-preserving
    return ~u;
}
__device__ __forceinline__ float ord_desc2f(uint32_t k) {
    uint32_t u = ~k;
    u ^= (u & 0x80000000u) ? 0x80000000u : 0xFFFFFFFFu;
    return __uint_as_float(u);
}

template <bool BF16>
__device__ __forceinline__ float load_elem(const void* base, long long idx) {
    if (BF16) return qc_bf2f(__ldg(reinterpret_cast<const uint16_t*>(base) + idx));
    return __ldg(reinterpret_cast<const float*>(base) + idx);
}

__device__ __forceinline__ float sigmoid_f32(float x) {
    // 1/(1+exp(-x)) in float32 (torch.sigmoid); IEEE division, accurate expf
    return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x)));
}

// ------------------------------------------------------------------------------------------
constexpr int K2_NC_MAX = 16;    // classes held in registers by the vector path

// Candidate test for one anchor given its nc class logits.  torch: s_c = sigmoid(x_c) (float32), first
// maximum over c, candidate iff s > (float)conf.  sigmoid is monotone, so only logits close to the maximum
// logit can hold the maximum sigmoid (they get the exact first-max treatment), and an anchor whose maximum
// logit is clearly below logit(conf) needs no sigmoid at all.  "Close": float32 sigmoid flattens as it
// saturates (every logit above ~16.7 gives exactly 1.0f; 13.95 and 14.0 round to the same float), so above 10
// the band is everything >= 10 -- torch.max then returns the FIRST class among the equal scores, which need
// not be the class of the largest logit.  Below 10 two logits 1e-3 apart differ by >= 7 ulps of the sigmoid.
__device__ __forceinline__ float k2_near_band(float m) { return fminf(m - 1e-3f * fmaxf(1.f, fabsf(m)), 10.f); }

template <int NCM>
__device__ __forceinline__ bool anchor_candidate(const float* x, int nc, float conf, float logit_lo, float& best,
                                                 int& best_c) {
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < NCM; ++c)
        if (c < nc) m = fmaxf(m, x[c]);
    if (!(m >= logit_lo)) return false;                 // sigmoid(m) < conf with a wide margin (NaN -> false)
    const float near = k2_near_band(m);
    best = -1.f; best_c = 0;
#pragma unroll
    for (int c = 0; c < NCM; ++c) {
        if (c < nc && x[c] >= near) {
            const float sg = sigmoid_f32(x[c]);
            if (sg > best) { best = sg; best_c = c; }   // first maximum wins (torch.max)
        }
    }
    return best > conf;                                 // float32 compare, like torch
}

template <bool BF16, int NCM>
__global__ void __launch_bounds__(256)
k2_filter(LevelTable lt, int nc, int reg_max, float conf, float logit_lo, int cap, int vec_ok,
          unsigned long long* __restrict__ cand_key, int* __restrict__ cand_count) {
    __shared__ int s_cnt, s_gbase;
    const int img = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int step = vec_ok ? 4 : 1;                    // anchors per thread per iteration
    for (int a0 = blockIdx.x * blockDim.x * step; a0 < lt.A; a0 += gridDim.x * blockDim.x * step) {
        const int a = a0 + threadIdx.x * step;
        bool is_cand[4] = {false, false, false, false};
        float best[4];
        int best_c[4];
        if (a < lt.A) {
            int l = 0;
#pragma unroll
            for (int k = 1; k < QC_MAX_LEVELS; ++k)
                if (k < lt.num_levels && a >= lt.start[k]) l = k;
            const long long base = (long long)img * lt.batch_stride[l] + (a - lt.start[l]);
            const long long cs = lt.chan_stride[l];
            const void* p = lt.ptr[l];
            if (vec_ok) {
                // 4 consecutive anchors x nc classes: one 128-bit (f32) / 64-bit (bf16) load per class plane
                float x[4][NCM];
#pragma unroll
                for (int c = 0; c < NCM; ++c) {
                    if (c < nc) {
                        const long long e = base + (long long)(4 * reg_max + c) * cs;
                        if (BF16) {
                            const uint2 v = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(p) + e));
                            x[0][c] = qc_bf2f((uint16_t)(v.x & 0xFFFF)); x[1][c] = qc_bf2f((uint16_t)(v.x >> 16));
                            x[2][c] = qc_bf2f((uint16_t)(v.y & 0xFFFF)); x[3][c] = qc_bf2f((uint16_t)(v.y >> 16));
                        } else {
                            const float4 v = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + e));
                            x[0][c] = v.x; x[1][c] = v.y; x[2][c] = v.z; x[3][c] = v.w;
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) is_cand[j] = anchor_candidate<NCM>(x[j], nc, conf, logit_lo, best[j], best_c[j]);
            } else {
                float m = -INFINITY;
                for (int c = 0; c < nc; ++c) m = fmaxf(m, load_elem<BF16>(p, base + (long long)(4 * reg_max + c) * cs));
                if (m >= logit_lo) {
                    const float near = k2_near_band(m);
                    best[0] = -1.f; best_c[0] = 0;
                    for (int c = 0; c < nc; ++c) {
                        const float xv = load_elem<BF16>(p, base + (long long)(4 * reg_max + c) * cs);
                        if (xv >= near) {
                            const float sg = sigmoid_f32(xv);
                            if (sg > best[0]) { best[0] = sg; best_c[0] = c; }
                        }
                    }
                    is_cand[0] = best[0] > conf;
                }
            }
        }
        // Slot allocation, aggregated per CTA: ballots give every candidate its offset inside the warp, one
        // shared-memory atomic per warp its offset inside the CTA, and ONE global atomic per CTA the base in the
        // image's list.  (A returning atomic per warp on the image's single counter serialises ~1000 L2 round trips.)
        unsigned bm[4];
        int woff[4], wtot = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            bm[j] = j < step ? __ballot_sync(0xffffffffu, is_cand[j]) : 0u;
            woff[j] = wtot;
            wtot += __popc(bm[j]);
        }
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        int wbase = 0;
        if (lane == 0 && wtot) wbase = atomicAdd(&s_cnt, wtot);
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        __syncthreads();
        if (threadIdx.x == 0) s_gbase = s_cnt ? atomicAdd(&cand_count[img], s_cnt) : 0;
        __syncthreads();
        const int gbase = s_gbase;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j < step && is_cand[j]) {
                const int slot = gbase + wbase + woff[j] + __popc(bm[j] & ((1u << lane) - 1));
                if (slot < cap)
                    cand_key[(long long)img * cap + slot] =
                        ((unsigned long long)f2ord_desc(best[j]) << 32) | ((unsigned long long)(a + j) << 8) |
                        (unsigned long long)best_c[j];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
template <bool BF16, int RM>   // RM == 16: registers only; RM == 0: runtime reg_max (<= 32)
__global__ void __launch_bounds__(256)
k2_decode(LevelTable lt, int reg_max_rt, int cap, const unsigned long long* __restrict__ cand_key,
          const int* __restrict__ cand_count, float4* __restrict__ cand_box) {
    const int img = blockIdx.y;
    int n = cand_count[img];
    if (n > cap) n = cap;
    const int side = threadIdx.x & 3;
    for (int ci = (blockIdx.x * blockDim.x + threadIdx.x) >> 2; ci < ((n + 63) & ~63);
         ci += (gridDim.x * blockDim.x) >> 2) {
        const bool valid = ci < n;
        float dist = 0.f, ax = 0.f, ay = 0.f, stride = 1.f;
        if (valid) {
            const unsigned long long key = cand_key[(long long)img * cap + ci];
            const int a = (int)((key >> 8) & 0xFFFFFFu);
            int l = 0;
#pragma unroll
            for (int k = 1; k < QC_MAX_LEVELS; ++k)
                if (k < lt.num_levels && a >= lt.start[k]) l = k;
            const int al = a - lt.start[l];
            const int y = al / lt.w[l], x = al - y * lt.w[l];
            ax = (float)x + 0.5f;
            ay = (float)y + 0.5f;
            stride = lt.stride[l];
            const long long base = (long long)img * lt.batch_stride[l] + al;
            const long long cs = lt.chan_stride[l];
            const void* p = lt.ptr[l];
            // softmax over reg_max bins of this side, expectation with arange weights
            const int reg_max = RM ? RM : reg_max_rt;
            float v[RM ? RM : 32];
            float mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < (RM ? RM : 32); ++k) {
                if (k < reg_max) {
                    v[k] = load_elem<BF16>(p, base + (long long)(side * reg_max + k) * cs);
                    mx = fmaxf(mx, v[k]);
                }
            }
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < (RM ? RM : 32); ++k) {
                if (k < reg_max) {
                    v[k] = expf(__fsub_rn(v[k], mx));
                    sum = __fadd_rn(sum, v[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < (RM ? RM : 32); ++k)
                if (k < reg_max) dist = __fadd_rn(dist, __fmul_rn((float)k, __fdiv_rn(v[k], sum)));
        }
        // gather the four sides into the side-0 lane of each quad
        const float d_t = __shfl_down_sync(0xffffffffu, dist, 1);
        const float d_r = __shfl_down_sync(0xffffffffu, dist, 2);
        const float d_b = __shfl_down_sync(0xffffffffu, dist, 3);
        if (valid && side == 0) {
            // dist2bbox(xywh=True) * stride, then ops.xywh2xyxy -- same float32 op order
            const float x1 = __fsub_rn(ax, dist), y1 = __fsub_rn(ay, d_t);
            const float x2 = __fadd_rn(ax, d_r), y2 = __fadd_rn(ay, d_b);
            const float cx = __fmul_rn(__fdiv_rn(__fadd_rn(x1, x2), 2.f), stride);
            const float cy = __fmul_rn(__fdiv_rn(__fadd_rn(y1, y2), 2.f), stride);
            const float w = __fmul_rn(__fsub_rn(x2, x1), stride);
            const float h = __fmul_rn(__fsub_rn(y2, y1), stride);
            const float hw = __fdiv_rn(w, 2.f), hh = __fdiv_rn(h, 2.f);
            cand_box[(long long)img * cap + ci] =
                make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
        }
    }
}

// ------------------------------------------------------------------------------------------
// keys for the NMS-only entry point
__global__ void k2_make_keys(const float* __restrict__ scores, const int* __restrict__ cls,
                             const int* __restrict__ counts, int cap, unsigned long long* __restrict__ cand_key,
                             int* __restrict__ cand_count) {
    const int img = blockIdx.y;
    int n = counts[img];
    if (n > cap) n = cap;
    if (n < 0) n = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) cand_count[img] = n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const long long o = (long long)img * cap + i;
        cand_key[o] = ((unsigned long long)f2ord_desc(scores[o]) << 32) | ((unsigned long long)i << 8) |
                      (unsigned long long)(cls[o] & 0xFF);
    }
}

// torchvision/csrc/ops/cpu/nms_kernel.cpp inner test, verbatim in float32 (no FMA):
//   ovr = inter / (iarea + jarea - inter);  suppress iff (double)ovr > thr
// Disjoint boxes (other class offsets, far apart) leave early with inter == 0; the IEEE division is
// only executed when inter/den is within 1e-4 of the threshold -- outside that band the comparison of
// inter against den*(thr -/+ 1e-4) already decides (float32 rounding is ~1e-7 relative).
__device__ __forceinline__ bool iou_gt(const float4 a, const float4 b, double thr, float thr_lo, float thr_hi,
                                       bool fast) {
    // Branch-light form: every lane computes inter and den (a dozen float ops, no early exits, so a warp whose lanes
    // meet different cases does not serialise them); only lanes the +-1e-4 band cannot decide run the division.
    // Disjoint boxes give inter == 0: decided "false" by inter < den*thr_lo when den > 0, and 0/den (<= 0 or NaN)
    // through the exact path otherwise -- the same answers as torchvision's expression.
    const float w = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
    const float h = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
    const float iarea = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float jarea = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    const float inter = __fmul_rn(fmaxf(0.f, w), fmaxf(0.f, h));
    const float den = __fsub_rn(__fadd_rn(iarea, jarea), inter);
    const bool lo = inter < __fmul_rn(den, thr_lo), hi = inter > __fmul_rn(den, thr_hi);
    bool res = hi;
    if (!(fast && den > 0.f && (lo || hi))) res = (double)__fdiv_rn(inter, den) > thr;
    return res;
}

struct NmsArgs {
    int cap, max_det, max_nms;
    float max_wh;
    double iou_thres;
    int scale_boxes;
    float gain, pad_x, pad_y, orig_w, orig_h;
    int mode;   // 0: full (write det rows + anchors), 1: nms-only (write kept candidate index)
    int class_prune;   // 1: boxes of different classes can never overlap after the offset (host-checked)
};

__device__ void bitonic_sort_pairs(unsigned long long* keys, int* vals, int npow2) {
    // one compare-exchange per thread per pass: thread t owns the pair (i, i | j), i = t with a zero bit spliced
    // in at position log2(j)
    const int half = npow2 >> 1;
    for (int k = 2; k <= npow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < half; t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int ixj = i | j;
                const unsigned long long a = keys[i], b = keys[ixj];
                const bool up = (i & k) == 0;
                if ((a > b) == up) {
                    keys[i] = b; keys[ixj] = a;
                    const int tv = vals[i]; vals[i] = vals[ixj]; vals[ixj] = tv;
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(K2_NMS_THREADS, 1)
k2_nms(NmsArgs args, const unsigned long long* __restrict__ cand_key, const float4* __restrict__ cand_box,
       const int* __restrict__ cand_count, unsigned long long* __restrict__ g_keys, int* __restrict__ g_vals,
       float4* __restrict__ g_sorted, float* __restrict__ out_det, int* __restrict__ out_idx,
       int* __restrict__ out_count, int* __restrict__ out_ncand) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ unsigned long long cmask2[2][256];  // per class: which boxes of the current chunk have it
                                                   // (double-buffered by chunk parity: reset races with nothing)
    __shared__ unsigned long long s_keptmask;
    __shared__ int s_kept_total;

    const int img = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cap = args.cap;
    int n = cand_count[img];
    if (n > cap) n = cap;
    if (out_ncand && tid == 0) out_ncand[img] = n;

    int npow2 = 1;
    while (npow2 < n) npow2 <<= 1;

    // carve shared memory: removed flags | kept list | (keys, vals, sorted boxes if they fit)
    uint8_t* removed = smem_raw;                                                 // K2_REMOVED_MAX
    int* kept_list = reinterpret_cast<int*>(smem_raw + K2_REMOVED_MAX);          // max_det ints
    uint8_t* dyn = smem_raw + K2_REMOVED_MAX + (size_t)((args.max_det + 3) & ~3) * 4;
    dyn = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dyn) + 15) & ~(uintptr_t)15);
    const bool in_smem = npow2 <= K2_SORT_SMEM_MAX;
    float4* sorted = in_smem ? reinterpret_cast<float4*>(dyn) : g_sorted + (long long)img * cap;
    unsigned long long* keys = in_smem ? reinterpret_cast<unsigned long long*>(dyn + (size_t)K2_SORT_SMEM_MAX * 16)
                                       : g_keys + (long long)img * 2 * cap;
    int* vals = in_smem ? reinterpret_cast<int*>(dyn + (size_t)K2_SORT_SMEM_MAX * 24) : g_vals + (long long)img * 2 * cap;
    // per-box intra-chunk suppression rows; the global path reuses the upper half of the sort buffer (free after the sort)
    unsigned long long* Mall = in_smem ? reinterpret_cast<unsigned long long*>(dyn + (size_t)K2_SORT_SMEM_MAX * 28)
                                       : g_keys + (long long)img * 2 * cap + cap;

    // ---- sort (score desc, anchor asc) ---------------------------------------------------
    for (int i = tid; i < npow2; i += blockDim.x) {
        keys[i] = i < n ? cand_key[(long long)img * cap + i] : ~0ull;
        vals[i] = i;
    }
    __syncthreads();
    bitonic_sort_pairs(keys, vals, npow2);
    if (n > args.max_nms) n = args.max_nms;      // top max_nms by score
    if (n > K2_REMOVED_MAX) n = K2_REMOVED_MAX;  // (rejected on the host; belt and braces)

    // ---- class-offset boxes in sorted order: x[:, :4] + cls * max_wh (float32) ------------
    for (int i = tid; i < n; i += blockDim.x) {
        const float4 b = cand_box[(long long)img * cap + vals[i]];
        const float off = __fmul_rn((float)(int)(keys[i] & 0xFF), args.max_wh);
        sorted[i] = make_float4(__fadd_rn(b.x, off), __fadd_rn(b.y, off), __fadd_rn(b.z, off), __fadd_rn(b.w, off));
        removed[i] = 0;
    }
    if (tid == 0) s_kept_total = 0;
    if (tid < 512) cmask2[tid >> 8][tid & 255] = 0ull;
    __syncthreads();

    // ---- greedy NMS in chunks of 64 sorted boxes -------------------------------------------
    const double thr = args.iou_thres;
    // fast paths are valid for thresholds in (1e-3, 1): widen the decision band by 1e-4 on both sides
    const bool fast = thr > 1e-3 && thr < 1.0;
    const float thr_lo = (float)(thr - 1e-4), thr_hi = (float)(thr + 1e-4);
    // The 64x64 upper-triangular suppression bitmask of every chunk depends only on the boxes, not on the
    // running state: all of them are computed up front by the whole CTA, so the sequential loop below is just
    // {resolve (warp 0) | barrier | apply kept boxes to later chunks | barrier}.
    for (int i = warp; i < n; i += K2_NMS_THREADS / 32) {
        const int cb = i & ~63, mm = min(64, n - cb), il = i - cb;
        const float4 bi = sorted[i];
        const int j0 = lane, j1 = lane + 32;
        const bool r0 = iou_gt(bi, sorted[cb + min(j0, mm - 1)], thr, thr_lo, thr_hi, fast);
        const bool r1 = iou_gt(bi, sorted[cb + min(j1, mm - 1)], thr, thr_lo, thr_hi, fast);
        const bool b0 = j0 > il && j0 < mm && r0, b1 = j1 > il && j1 < mm && r1;
        const unsigned lo = __ballot_sync(0xffffffffu, b0), hi = __ballot_sync(0xffffffffu, b1);
        if (lane == 0) Mall[i] = ((unsigned long long)hi << 32) | lo;
    }
    __syncthreads();
    for (int c0 = 0; c0 < n; c0 += 64) {
        const int m = min(64, n - c0);
        unsigned long long* cmask = cmask2[(c0 >> 6) & 1];
        if (args.class_prune && tid >= 64 && tid < 64 + m)
            atomicOr(&cmask[keys[c0 + tid - 64] & 0xFF], 1ull << (tid - 64));
        if (warp == 0) {
            // serial greedy pass over the chunk, executed redundantly by every lane of warp 0: lane l holds
            // mask rows l and l+32, row i is broadcast by shuffle (independent of the running state, so the
            // shuffles pipeline ahead of the dependent chain on `rem`)
            const bool r0 = lane < m && removed[c0 + lane];
            const bool r1 = lane + 32 < m && removed[c0 + lane + 32];
            const unsigned rlo = __ballot_sync(0xffffffffu, r0), rhi = __ballot_sync(0xffffffffu, r1);
            const unsigned long long m_lo = lane < m ? Mall[c0 + lane] : 0ull;
            const unsigned long long m_hi = lane + 32 < m ? Mall[c0 + lane + 32] : 0ull;
            unsigned long long rem = ((unsigned long long)rhi << 32) | rlo, keptmask = 0;
            int kt = s_kept_total;
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                const unsigned long long mi = __shfl_sync(0xffffffffu, i < 32 ? m_lo : m_hi, i & 31);
                const bool k = !((rem >> i) & 1ull) && i < m && kt < args.max_det;
                keptmask |= (unsigned long long)k << i;
                rem |= k ? mi : 0ull;
                if (k && lane == 0) kept_list[kt] = c0 + i;
                kt += k;
            }
            __syncwarp();
            if (lane == 0) { s_keptmask = keptmask; s_kept_total = kt; }
        }
        __syncthreads();
        const unsigned long long keptmask = s_keptmask;
        if (s_kept_total >= args.max_det) break;
        if (keptmask) {
            for (int j = c0 + 64 + tid; j < n; j += K2_NMS_THREADS) {
                if (removed[j]) continue;
                const float4 bj = sorted[j];
                // cross-class pairs cannot overlap: only kept boxes of this box's class are tested
                unsigned long long km = args.class_prune ? (keptmask & cmask[keys[j] & 0xFF]) : keptmask;
                while (km) {
                    const int i = __ffsll((long long)km) - 1;
                    km &= km - 1;
                    if (iou_gt(sorted[c0 + i], bj, thr, thr_lo, thr_hi, fast)) { removed[j] = 1; break; }
                }
            }
        }
        __syncthreads();
        if (args.class_prune && tid < m) cmask[keys[c0 + tid] & 0xFF] = 0ull;    // reset only the touched entries
    }
    __syncthreads();

    // ---- epilogue: x[keep] (un-offset boxes) -> scale_boxes -> rows -------------------------
    const int kt = s_kept_total;
    if (tid == 0) out_count[img] = kt;
    for (int r = tid; r < args.max_det; r += blockDim.x) {
        if (args.mode == 1) {
            out_idx[(long long)img * args.max_det + r] = r < kt ? vals[kept_list[r]] : -1;
            continue;
        }
        float* row = out_det + ((long long)img * args.max_det + r) * 6;
        if (r < kt) {
            const int pos = kept_list[r];
            const unsigned long long key = keys[pos];
            float4 b = cand_box[(long long)img * cap + vals[pos]];
            if (args.scale_boxes) {
                b.x = __fdiv_rn(__fsub_rn(b.x, args.pad_x), args.gain);
                b.y = __fdiv_rn(__fsub_rn(b.y, args.pad_y), args.gain);
                b.z = __fdiv_rn(__fsub_rn(b.z, args.pad_x), args.gain);
                b.w = __fdiv_rn(__fsub_rn(b.w, args.pad_y), args.gain);
                b.x = fminf(fmaxf(b.x, 0.f), args.orig_w);
                b.y = fminf(fmaxf(b.y, 0.f), args.orig_h);
                b.z = fminf(fmaxf(b.z, 0.f), args.orig_w);
                b.w = fminf(fmaxf(b.w, 0.f), args.orig_h);
            }
            row[0] = b.x; row[1] = b.y; row[2] = b.z; row[3] = b.w;
            row[4] = ord_desc2f((uint32_t)(key >> 32));
            row[5] = (float)(int)(key & 0xFF);
            out_idx[(long long)img * args.max_det + r] = (int)((key >> 8) & 0xFFFFFFu);
        } else {
            row[0] = row[1] = row[2] = row[3] = row[4] = row[5] = 0.f;
            out_idx[(long long)img * args.max_det + r] = -1;
        }
    }
}


// ------------------------------------------------------------------------------------------
// Class-parallel NMS (the decode path, when the host has verified that boxes of different classes cannot
// intersect after the +cls*max_wh offset): torchvision's greedy pass over the score-sorted list then factorises
// into nc independent greedy passes, so one CTA takes one (image, class) -- nc times more CTAs in flight and
// nc times shorter sequential chains than one CTA per image.  Each CTA: pick its class out of the image's
// candidate list, bitonic sort, offset boxes (same float32 rounding as the reference), 64x64 chunk bitmasks,
// greedy resolve, first max_det kept boxes -> per-class kept list in global memory.  The LAST class CTA of an
// image to finish (ticket counter) merges the nc sorted kept lists by rank (keys are unique) and writes the
// first max_det rows -- identical to x[keep][:max_det] of the single sorted pass.
constexpr int K2C_THREADS = 512;
constexpr int K2C_SMEM_N = 1024;         // candidates of one class sorted / held in shared memory

__device__ __forceinline__ int lower_bound_cg(const unsigned long long* __restrict__ a, int n, unsigned long long k) {
    int lo = 0, hi = n;                   // first index with a[idx] >= k  (== number of keys < k)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldcg(a + mid) < k) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// k-th smallest (1-based) of n unique 64-bit keys by MSB-first 8-bit radix select; all threads of the CTA call it
// and get the same value.  Only used when an image holds more than max_nms candidates (the reference keeps the
// max_nms best-scoring ones before NMS).
__device__ unsigned long long kth_smallest_key(const unsigned long long* __restrict__ ck, int n, int k, int* s_hist,
                                               unsigned long long* s_prefix, int* s_rem) {
    if (threadIdx.x == 0) { *s_prefix = 0ull; *s_rem = k; }
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
        const unsigned long long prefix = *s_prefix;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned long long key = ck[i];
            if (pass == 0 || (key >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&s_hist[(int)((key >> shift) & 255ull)], 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int rem = *s_rem, b = 0;
            while (b < 255 && rem > s_hist[b]) { rem -= s_hist[b]; ++b; }
            *s_rem = rem;
            *s_prefix = prefix | ((unsigned long long)b << shift);
        }
        __syncthreads();
    }
    return *s_prefix;
}

__global__ void __launch_bounds__(K2C_THREADS, 2)
k2_nms_class(NmsArgs args, int nc, const unsigned long long* __restrict__ cand_key,
             const float4* __restrict__ cand_box, const int* __restrict__ cand_count,
             unsigned long long* __restrict__ g_keys, int* __restrict__ g_vals, float4* __restrict__ g_sorted,
             unsigned long long* __restrict__ g_mall, unsigned long long* __restrict__ kept_key,
             float4* __restrict__ kept_box, int* __restrict__ kept_count, int* __restrict__ ticket,
             float* __restrict__ out_det, int* __restrict__ out_idx, int* __restrict__ out_count,
             int* __restrict__ out_ncand) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ int s_hist[256];              // candidates per class of this image
    __shared__ float4 s_kb[64];              // kept boxes of the current chunk
    __shared__ int s_fill;
    __shared__ unsigned long long s_keptmask;
    __shared__ int s_kept_total, s_ticket;

    const int cls = blockIdx.x, img = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cap = args.cap;
    int ntot = cand_count[img];
    if (ntot > cap) ntot = cap;
    if (out_ncand && tid == 0 && cls == 0) out_ncand[img] = ntot;
    const unsigned long long* ck = cand_key + (long long)img * cap;

    // more than max_nms candidates: only the max_nms best keys (score desc, anchor asc) take part
    unsigned long long key_max = ~0ull;
    if (ntot > args.max_nms) key_max = kth_smallest_key(ck, ntot, args.max_nms, s_hist, &s_keptmask, &s_ticket);
    // ---- class histogram of the image's candidates (every class CTA derives the same table) --------
    __syncthreads();
    for (int i = tid; i < 256; i += K2C_THREADS) s_hist[i] = 0;
    if (tid == 0) { s_fill = 0; s_kept_total = 0; }
    __syncthreads();
    for (int i = tid; i < ntot; i += K2C_THREADS) {
        const unsigned long long k = ck[i];
        if (k <= key_max) atomicAdd(&s_hist[(int)(k & 0xFF)], 1);
    }
    __syncthreads();
    int n = s_hist[cls];
    int npow2 = 1;
    while (npow2 < n) npow2 <<= 1;
    const bool in_smem = npow2 <= K2C_SMEM_N;
    // global fallback regions of this class: classes partition the candidates, so sum(npow2_c) <= 2*cap
    long long koff = 0, boff = 0;
    if (!in_smem) {
        for (int c = 0; c < cls; ++c) {
            const int m = s_hist[c];
            int p2 = m ? 1 : 0;
            while (p2 < m) p2 <<= 1;
            koff += p2; boff += m;
        }
    }
    // carve shared memory: removed bitmask | kept list | sorted boxes, keys, vals, chunk masks
    unsigned* removed = reinterpret_cast<unsigned*>(smem_raw);                           // K2_REMOVED_MAX bits
    int* kept_list = reinterpret_cast<int*>(smem_raw + K2_REMOVED_MAX / 8);              // max_det ints
    uint8_t* dyn = smem_raw + K2_REMOVED_MAX / 8 + (size_t)((args.max_det + 3) & ~3) * 4;
    dyn = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dyn) + 15) & ~(uintptr_t)15);
    const double thr = args.iou_thres;
    const bool fast = thr > 1e-3 && thr < 1.0;
    const float thr_lo = (float)(thr - 1e-4), thr_hi = (float)(thr + 1e-4);

    // The body is instantiated twice -- shared-memory buffers (the normal case) and the global-memory regions of an
    // oversized class -- so that the compiler knows the address space of every access (LDS/STS instead of generic).
    auto core = [&](float4* sorted, unsigned long long* keys, int* vals, unsigned long long* Mall) {
        // ---- gather this class (any order: the sort fixes it) -----------------------------------------
        for (int i0 = 0; i0 < ntot; i0 += K2C_THREADS) {
            const int i = i0 + tid;
            unsigned long long k = 0;
            const bool mine = i < ntot && (int)((k = ck[i]) & 0xFF) == cls && k <= key_max;
            const unsigned bm = __ballot_sync(0xffffffffu, mine);
            if (bm) {
                int base = 0;
                const int leader = __ffs(bm) - 1;
                if (lane == leader) base = atomicAdd(&s_fill, __popc(bm));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (mine) {
                    const int slot = base + __popc(bm & ((1u << lane) - 1u));
                    keys[slot] = k; vals[slot] = i;
                }
            }
        }
        for (int i = n + tid; i < npow2; i += K2C_THREADS) { keys[i] = ~0ull; vals[i] = 0; }
        __syncthreads();
        bitonic_sort_pairs(keys, vals, npow2);
        if (n > K2_REMOVED_MAX) n = K2_REMOVED_MAX;  // (max_nms <= K2_REMOVED_MAX is checked on the host)
        const float off = __fmul_rn((float)cls, args.max_wh);
        for (int i = tid; i < n; i += K2C_THREADS) {
            const float4 b = cand_box[(long long)img * cap + vals[i]];
            sorted[i] = make_float4(__fadd_rn(b.x, off), __fadd_rn(b.y, off), __fadd_rn(b.z, off), __fadd_rn(b.w, off));
        }
        for (int i = tid; i < min((n + 31) / 32 + 2, K2_REMOVED_MAX / 32); i += K2C_THREADS) removed[i] = 0u;
        __syncthreads();

        // 64x64 upper-triangular suppression bitmask of every chunk (depends only on the boxes)
        for (int i = warp; i < n; i += K2C_THREADS / 32) {
            const int cb = i & ~63, mm = min(64, n - cb), il = i - cb;
            const float4 bi = sorted[i];
            const int j0 = lane, j1 = lane + 32;
            const bool r0 = iou_gt(bi, sorted[cb + min(j0, mm - 1)], thr, thr_lo, thr_hi, fast);
            const bool r1 = iou_gt(bi, sorted[cb + min(j1, mm - 1)], thr, thr_lo, thr_hi, fast);
            const bool b0 = j0 > il && j0 < mm && r0, b1 = j1 > il && j1 < mm && r1;
            const unsigned lo = __ballot_sync(0xffffffffu, b0), hi = __ballot_sync(0xffffffffu, b1);
            if (lane == 0) Mall[i] = ((unsigned long long)hi << 32) | lo;
        }
        __syncthreads();
        for (int c0 = 0; c0 < n; c0 += 64) {
            const int m = min(64, n - c0);
            if (warp == 0) {
                // Greedy pass over the chunk, redundantly in every lane of warp 0.  Row i only holds bits > i, so bit i
                // of `rem` is final once row i-1 has been applied: the chain per step is {test bit i, OR row i} and the
                // kept set falls out as ~rem at the end.
                const unsigned long long m_lo = lane < m ? Mall[c0 + lane] : 0ull;
                const unsigned long long m_hi = lane + 32 < m ? Mall[c0 + lane + 32] : 0ull;
                unsigned long long rem = ((unsigned long long)removed[(c0 >> 5) + 1] << 32) | removed[c0 >> 5];
                if (m < 64) rem |= ~0ull << m;
#pragma unroll
                for (int i = 0; i < 64; ++i) {
                    const unsigned long long mi = __shfl_sync(0xffffffffu, i < 32 ? m_lo : m_hi, i & 31);
                    if (!((rem >> i) & 1ull)) rem |= mi;
                }
                unsigned long long keptmask = ~rem;
                const int kt0 = s_kept_total;
                int room = args.max_det - kt0;                 // only the first max_det kept boxes count
                while (__popcll(keptmask) > room) keptmask &= ~(1ull << (63 - __clzll((long long)keptmask)));
                if ((keptmask >> lane) & 1ull) kept_list[kt0 + __popcll(keptmask & ((1ull << lane) - 1ull))] = c0 + lane;
                if ((keptmask >> (lane + 32)) & 1ull)
                    kept_list[kt0 + __popcll(keptmask & ((1ull << (lane + 32)) - 1ull))] = c0 + lane + 32;
                __syncwarp();
                if (lane == 0) { s_keptmask = keptmask; s_kept_total = kt0 + __popcll(keptmask); }
            }
            __syncthreads();
            const unsigned long long keptmask = s_keptmask;
            if (s_kept_total >= args.max_det) break;
            const int nk = __popcll(keptmask);
            if (nk && c0 + 64 < n) {
                // the chunk's kept boxes, densely packed, then into registers: lanes hold <= 2 kept boxes each and
                // every warp walks the later boxes four at a time (loads first, eight branch-light IoU tests, four
                // votes) -- no dependent chain between boxes
                if (tid < 64 && ((keptmask >> tid) & 1ull))
                    s_kb[__popcll(keptmask & ((1ull << tid) - 1ull))] = sorted[c0 + tid];
                __syncthreads();
                const bool h0 = lane < nk, h1 = lane + 32 < nk;
                const float4 k0 = h0 ? s_kb[lane] : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 k1 = h1 ? s_kb[lane + 32] : make_float4(0.f, 0.f, 0.f, 0.f);
                constexpr int NW = K2C_THREADS / 32;
                for (int j0 = c0 + 64 + warp; j0 < n; j0 += 4 * NW) {
                    float4 bj[4];
                    bool live[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = j0 + u * NW;
                        live[u] = j < n && !((removed[j >> 5] >> (j & 31)) & 1u);      // warp-uniform
                        bj[u] = sorted[min(j, n - 1)];
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const bool s0 = iou_gt(k0, bj[u], thr, thr_lo, thr_hi, fast);
                        const bool s1 = iou_gt(k1, bj[u], thr, thr_lo, thr_hi, fast);
                        const bool sup = live[u] && ((h0 && s0) || (h1 && s1));
                        const int j = j0 + u * NW;
                        if (__any_sync(0xffffffffu, sup) && lane == 0) atomicOr(&removed[j >> 5], 1u << (j & 31));
                    }
                }
            }
            __syncthreads();
        }
        __syncthreads();
        // ---- publish this class's kept list ------------------------------------------------------------
        const int kt = s_kept_total;
        const long long kbase = ((long long)img * nc + cls) * args.max_det;
        for (int r = tid; r < kt; r += K2C_THREADS) {
            const int pos = kept_list[r];
            __stcg(kept_key + kbase + r, keys[pos]);
            __stcg(kept_box + kbase + r, cand_box[(long long)img * cap + vals[pos]]);      // un-offset box
        }
    };
    if (n > 0) {
        if (in_smem)
            core(reinterpret_cast<float4*>(dyn), reinterpret_cast<unsigned long long*>(dyn + (size_t)K2C_SMEM_N * 16),
                 reinterpret_cast<int*>(dyn + (size_t)K2C_SMEM_N * 24),
                 reinterpret_cast<unsigned long long*>(dyn + (size_t)K2C_SMEM_N * 28));
        else
            core(g_sorted + (long long)img * cap + boff, g_keys + (long long)img * 2 * cap + koff,
                 g_vals + (long long)img * 2 * cap + koff, g_mall + (long long)img * cap + boff);
    }
    const int kt = s_kept_total;
    // ---- take a ticket ----------------------------------------------------------------------------------
    if (tid == 0) __stcg(kept_count + img * nc + cls, kt);
    __threadfence();
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(&ticket[img], 1);
    __syncthreads();
    if (s_ticket != nc - 1) return;
    __threadfence();

    // ---- last CTA of the image: rank-merge the nc kept lists, scale_boxes epilogue ----------------------
    int* s_cnt = s_hist;                                  // reuse: kept count per class
    for (int c = tid; c < nc; c += K2C_THREADS) s_cnt[c] = __ldcg(kept_count + img * nc + c);
    __syncthreads();
    int total = 0;
    for (int c = 0; c < nc; ++c) total += s_cnt[c];
    const int nout = min(total, args.max_det);
    if (tid == 0) { out_count[img] = nout; ticket[img] = 0; }
    // the kept lists go to shared memory (the sort buffers are free now) so the binary searches chase
    // shared-memory, not L2, latency; oversized (nc * max_det) tables are searched in global memory
    unsigned long long* mk = reinterpret_cast<unsigned long long*>(dyn);
    const bool merge_smem = (size_t)nc * args.max_det * 8 <= (size_t)K2C_SMEM_N * 36;
    if (merge_smem) {
        for (int e = tid; e < nc * args.max_det; e += K2C_THREADS) {
            const int c = e / args.max_det, r = e - c * args.max_det;
            if (r < s_cnt[c]) mk[e] = __ldcg(kept_key + (long long)img * nc * args.max_det + e);
        }
        __syncthreads();
    }
    for (int e = tid; e < nc * args.max_det; e += K2C_THREADS) {
        const int c = e / args.max_det, r = e - c * args.max_det;
        if (r >= s_cnt[c]) continue;
        const long long eb = ((long long)img * nc + c) * args.max_det;
        const unsigned long long key = merge_smem ? mk[e] : __ldcg(kept_key + eb + r);
        int rank = r;
        for (int c2 = 0; c2 < nc && rank < args.max_det; ++c2) {
            if (c2 == c || s_cnt[c2] == 0) continue;
            if (merge_smem) {
                const unsigned long long* a = mk + c2 * args.max_det;
                int lo = 0, hi = s_cnt[c2];
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (a[mid] < key) lo = mid + 1; else hi = mid;
                }
                rank += lo;
            } else {
                rank += lower_bound_cg(kept_key + ((long long)img * nc + c2) * args.max_det, s_cnt[c2], key);
            }
        }
        if (rank >= args.max_det) continue;
        float4 b = __ldcg(kept_box + eb + r);
        if (args.scale_boxes) {
            b.x = __fdiv_rn(__fsub_rn(b.x, args.pad_x), args.gain);
            b.y = __fdiv_rn(__fsub_rn(b.y, args.pad_y), args.gain);
            b.z = __fdiv_rn(__fsub_rn(b.z, args.pad_x), args.gain);
            b.w = __fdiv_rn(__fsub_rn(b.w, args.pad_y), args.gain);
            b.x = fminf(fmaxf(b.x, 0.f), args.orig_w);
            b.y = fminf(fmaxf(b.y, 0.f), args.orig_h);
            b.z = fminf(fmaxf(b.z, 0.f), args.orig_w);
            b.w = fminf(fmaxf(b.w, 0.f), args.orig_h);
        }
        float* row = out_det + ((long long)img * args.max_det + rank) * 6;
        row[0] = b.x; row[1] = b.y; row[2] = b.z; row[3] = b.w;
        row[4] = ord_desc2f((uint32_t)(key >> 32));
        row[5] = (float)(int)(key & 0xFF);
        out_idx[(long long)img * args.max_det + rank] = (int)((key >> 8) & 0xFFFFFFu);
    }
    for (int r = nout + tid; r < args.max_det; r += K2C_THREADS) {
        float* row = out_det + ((long long)img * args.max_det + r) * 6;
        row[0] = row[1] = row[2] = row[3] = row[4] = row[5] = 0.f;
        out_idx[(long long)img * args.max_det + r] = -1;
    }
}

size_t nms_class_smem_bytes(int max_det) {
    return (size_t)K2_REMOVED_MAX / 8 + (size_t)((max_det + 3) & ~3) * 4 + 16 + (size_t)K2C_SMEM_N * (16 + 8 + 4 + 8);
}

// workspace layout (all 256-B aligned):
//   cand_count int[n] + ticket int[n] (zeroed together) | cand_key u64[n*cap] | cand_box float4[n*cap] |
//   g_keys u64[n*2cap] | g_vals int[n*2cap] | g_sorted float4[n*cap]      (global-sort fallback)
//   g_mall u64[n*cap] | kept_key u64[n*nc*max_det] | kept_box float4[n*nc*max_det] | kept_count int[n*nc]   (per-class NMS)
struct Ws {
    int* count; int* ticket; unsigned long long* key; float4* box; unsigned long long* gkeys; int* gvals; float4* gsorted;
    unsigned long long* gmall; unsigned long long* kept_key; float4* kept_box; int* kept_count;
    size_t total;
};
Ws carve(void* base, int n, int cap, int nc = 0, int max_det = 0) {
    Ws w;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = qc_align_up(off + bytes, 256); return o; };
    uint8_t* b = (uint8_t*)base;
    w.count = (int*)(b + take((size_t)n * 8));
    w.ticket = w.count + n;
    w.key = (unsigned long long*)(b + take((size_t)n * cap * 8));
    w.box = (float4*)(b + take((size_t)n * cap * 16));
    w.gkeys = (unsigned long long*)(b + take((size_t)n * 2 * cap * 8));
    w.gvals = (int*)(b + take((size_t)n * 2 * cap * 4));
    w.gsorted = (float4*)(b + take((size_t)n * cap * 16));
    w.gmall = (unsigned long long*)(b + take((size_t)n * cap * 8 * (nc > 0)));
    w.kept_key = (unsigned long long*)(b + take((size_t)n * nc * max_det * 8));
    w.kept_box = (float4*)(b + take((size_t)n * nc * max_det * 16));
    w.kept_count = (int*)(b + take((size_t)n * nc * 4));
    w.total = off;
    return w;
}

size_t nms_smem_bytes(int max_det) {
    return (size_t)K2_REMOVED_MAX + (size_t)((max_det + 3) & ~3) * 4 + 16 + (size_t)K2_SORT_SMEM_MAX * (16 + 8 + 4 + 8);
}

int total_anchors(const qc_yolo_params* p) {
    long long A = 0;
    for (int l = 0; l < p->num_levels; ++l) A += (long long)p->level_h[l] * p->level_w[l];
    return (int)A;
}

int launch_nms(qc_ctx* ctx, const NmsArgs& a, int n, const Ws& w, float* out_det, int* out_idx, int* out_count,
               int* out_ncand, cudaStream_t st) {
    const size_t smem = nms_smem_bytes(a.max_det);
    if ((int)smem > ctx->max_smem_optin) QC_FAIL(ctx, QC_ERR_UNSUPPORTED, "k2_nms: max_det %d needs too much shared memory", a.max_det);
    if (int rc = qc_prepare(ctx, k2_nms, (int)smem, K2_NMS_THREADS, nullptr)) return rc;
    k2_nms<<<n, K2_NMS_THREADS, smem, st>>>(a, w.key, w.box, w.count, w.gkeys, w.gvals, w.gsorted, out_det, out_idx,
                                            out_count, out_ncand);
    QC_LAUNCHED_AS(ctx, "k2_nms", st);
    return QC_OK;
}

}  // namespace

extern "C" size_t qc_yolo_workspace_bytes(const qc_yolo_params* p) {
    if (!p || p->n <= 0 || p->num_levels <= 0 || p->num_levels > QC_MAX_LEVELS) return 0;
    return carve(nullptr, p->n, total_anchors(p), p->nc, p->max_det).total;
}

extern "C" size_t qc_nms_workspace_bytes(int n, int cap) {
    if (n <= 0 || cap <= 0) return 0;
    return carve(nullptr, n, cap).total;
}

extern "C" int qc_yolo_decode_nms(qc_ctx* ctx, const void* const* levels, const qc_yolo_params* p, void* workspace,
                                  size_t workspace_bytes, float* out_det, int32_t* out_anchor, int32_t* out_count,
                                  int32_t* out_ncand, void* stream) {
    qc_device_guard dev_guard(ctx);
    if (!ctx) return QC_ERR_INVALID;
    QC_ENTER(ctx, stream);
    if (!levels || !p || !out_det || !out_anchor || !out_count)
        QC_FAIL(ctx, QC_ERR_INVALID, "qc_yolo_decode_nms: null argument");
    if (p->n == 0) return QC_OK;
    if (p->n < 0 || p->num_levels <= 0 || p->num_levels > QC_MAX_LEVELS || p->nc <= 0 || p->nc > 128 ||
        p->reg_max <= 0 || p->reg_max > 32 || p->max_det <= 0 || p->max_nms <= 0)
        QC_FAIL(ctx, QC_ERR_INVALID, "qc_yolo_decode_nms: bad params");
    if (p->max_nms > K2_REMOVED_MAX) QC_FAIL(ctx, QC_ERR_UNSUPPORTED, "max_nms > %d unsupported", K2_REMOVED_MAX);
    if (p->in_dtype != QC_DT_F32 && p->in_dtype != QC_DT_BF16) QC_FAIL(ctx, QC_ERR_INVALID, "bad in_dtype");
    LevelTable lt;
    memset(&lt, 0, sizeof(lt));
    lt.num_levels = p->num_levels;
    int A = 0;
    for (int l = 0; l < p->num_levels; ++l) {
        if (!levels[l] || p->level_h[l] <= 0 || p->level_w[l] <= 0) QC_FAIL(ctx, QC_ERR_INVALID, "bad level %d", l);
        lt.ptr[l] = levels[l];
        lt.batch_stride[l] = p->level_batch_stride[l];
        lt.chan_stride[l] = p->level_chan_stride[l];
        lt.h[l] = p->level_h[l]; lt.w[l] = p->level_w[l];
        lt.stride[l] = p->level_stride[l];
        lt.start[l] = A;
        A += p->level_h[l] * p->level_w[l];
    }
    for (int l = p->num_levels; l <= QC_MAX_LEVELS; ++l) lt.start[l] = A;
    lt.A = A;
    if (A >= (1 << 24)) QC_FAIL(ctx, QC_ERR_UNSUPPORTED, "too many anchors");
    const int cap = A;
    Ws w = carve(workspace, p->n, cap, p->nc, p->max_det);
    if (!workspace || workspace_bytes < w.total) QC_FAIL(ctx, QC_ERR_WORKSPACE, "workspace %zu < %zu", workspace_bytes, w.total);
    cudaStream_t st = (cudaStream_t)stream;
    QC_CUDA(ctx, cudaMemsetAsync(w.count, 0, (size_t)p->n * 8, st));      // candidate counts + merge tickets

    const float conf = (float)p->conf_thres;
    // logit(conf) minus a wide margin: anchors whose best logit is below it cannot reach conf
    float logit_lo = -INFINITY;
    if (p->conf_thres > 0.0 && p->conf_thres < 1.0) {
        const double lg = log(p->conf_thres / (1.0 - p->conf_thres));
        logit_lo = (float)(lg - 1e-2 - 1e-3 * fabs(lg));
    } else if (p->conf_thres >= 1.0) {
        logit_lo = 8.0f;          // sigmoid(8) = 0.99966 < 1: nothing below can exceed conf >= 1
    }
    const bool bf = p->in_dtype == QC_DT_BF16;
    // vector path: every level plane starts 16-byte (f32) / 8-byte (bf16) aligned and holds a multiple of 4 anchors
    int vec_ok = p->nc <= K2_NC_MAX;
    const size_t esz = bf ? 2 : 4;
    for (int l = 0; l < p->num_levels && vec_ok; ++l) {
        if ((p->level_h[l] * p->level_w[l]) % 4 || (lt.chan_stride[l] % 4) || (lt.batch_stride[l] % 4) ||
            (reinterpret_cast<uintptr_t>(levels[l]) % (4 * esz)))
            vec_ok = 0;
    }
    const int per_cta = 256 * (vec_ok ? 4 : 1);
    dim3 gf((unsigned)std::min((A + per_cta - 1) / per_cta, 64), (unsigned)p->n);
    if (p->nc <= 8) {
        if (bf) k2_filter<true, 8><<<gf, 256, 0, st>>>(lt, p->nc, p->reg_max, conf, logit_lo, cap, vec_ok, w.key, w.count);
        else k2_filter<false, 8><<<gf, 256, 0, st>>>(lt, p->nc, p->reg_max, conf, logit_lo, cap, vec_ok, w.key, w.count);
    } else {
        if (bf) k2_filter<true, K2_NC_MAX><<<gf, 256, 0, st>>>(lt, p->nc, p->reg_max, conf, logit_lo, cap, vec_ok, w.key, w.count);
        else k2_filter<false, K2_NC_MAX><<<gf, 256, 0, st>>>(lt, p->nc, p->reg_max, conf, logit_lo, cap, vec_ok, w.key, w.count);
    }
    QC_LAUNCHED_AS(ctx, "k2_filter", st);

    dim3 gd(32, (unsigned)p->n);
    if (p->reg_max == 16) {
        if (bf) k2_decode<true, 16><<<gd, 256, 0, st>>>(lt, 16, cap, w.key, w.count, w.box);
        else k2_decode<false, 16><<<gd, 256, 0, st>>>(lt, 16, cap, w.key, w.count, w.box);
    } else {
        if (bf) k2_decode<true, 0><<<gd, 256, 0, st>>>(lt, p->reg_max, cap, w.key, w.count, w.box);
        else k2_decode<false, 0><<<gd, 256, 0, st>>>(lt, p->reg_max, cap, w.key, w.count, w.box);
    }
    QC_LAUNCHED_AS(ctx, "k2_decode", st);

    NmsArgs a;
    a.cap = cap; a.max_det = p->max_det; a.max_nms = p->max_nms; a.max_wh = p->max_wh; a.iou_thres = p->iou_thres;
    a.scale_boxes = p->scale_boxes; a.gain = p->gain; a.pad_x = (float)p->pad_x; a.pad_y = (float)p->pad_y;
    a.orig_w = (float)p->orig_w; a.orig_h = (float)p->orig_h; a.mode = 0;
    // decoded coordinates lie within [-(reg_max-1)*stride, extent + (reg_max-1)*stride]: if max_wh exceeds that
    // span, boxes of different classes never intersect after the +cls*max_wh offset (IoU == 0 exactly)
    float span = 0.f, smax = 0.f;
    for (int l = 0; l < p->num_levels; ++l) {
        span = fmaxf(span, fmaxf((float)p->level_w[l], (float)p->level_h[l]) * p->level_stride[l]);
        smax = fmaxf(smax, p->level_stride[l]);
    }
    a.class_prune = (p->max_wh > span + 2.f * p->reg_max * smax + 2.f) ? 1 : 0;
    if (a.class_prune && p->nc <= 256) {
        // classes are independent: one CTA per (image, class), merged by the last CTA of each image
        const size_t smem = nms_class_smem_bytes(a.max_det);
        if ((int)smem + 4096 <= ctx->max_smem_optin) {
            if (int rc = qc_prepare(ctx, k2_nms_class, (int)smem, K2C_THREADS, nullptr)) return rc;
            k2_nms_class<<<dim3((unsigned)p->nc, (unsigned)p->n), K2C_THREADS, smem, st>>>(
                a, p->nc, w.key, w.box, w.count, w.gkeys, w.gvals, w.gsorted, w.gmall, w.kept_key, w.kept_box,
                w.kept_count, w.ticket, out_det, out_anchor, out_count, out_ncand);
            QC_LAUNCHED_AS(ctx, "k2_nms_class", st);
            return QC_OK;
        }
    }
    return launch_nms(ctx, a, p->n, w, out_det, out_anchor, out_count, out_ncand, st);
}

extern "C" int qc_nms(qc_ctx* ctx, const float* boxes, const float* scores, const int32_t* cls, const int32_t* counts,
                      int n, int cap, double iou_thres, float max_wh, int max_det, void* workspace,
                      size_t workspace_bytes, int32_t* out_keep, int32_t* out_count, void* stream) {
    qc_device_guard dev_guard(ctx);
    if (!ctx) return QC_ERR_INVALID;
    QC_ENTER(ctx, stream);
    if (!boxes || !scores || !cls || !counts || !out_keep || !out_count || cap <= 0 || max_det <= 0 || n < 0)
        QC_FAIL(ctx, QC_ERR_INVALID, "qc_nms: bad argument");
    if (n == 0) return QC_OK;
    if (cap > K2_REMOVED_MAX) QC_FAIL(ctx, QC_ERR_UNSUPPORTED, "qc_nms: cap > %d unsupported", K2_REMOVED_MAX);
    if (reinterpret_cast<uintptr_t>(boxes) & 15) QC_FAIL(ctx, QC_ERR_INVALID, "qc_nms: boxes must be 16-byte aligned");
    Ws w = carve(workspace, n, cap);
    if (!workspace || workspace_bytes < w.total) QC_FAIL(ctx, QC_ERR_WORKSPACE, "workspace %zu < %zu", workspace_bytes, w.total);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 g((unsigned)std::min((cap + 255) / 256, 32), (unsigned)n);
    k2_make_keys<<<g, 256, 0, st>>>(scores, cls, counts, cap, w.key, w.count);
    QC_LAUNCHED_AS(ctx, "k2_make_keys", st);
    w.box = const_cast<float4*>(reinterpret_cast<const float4*>(boxes));
    NmsArgs a;
    a.cap = cap; a.max_det = max_det; a.max_nms = cap; a.max_wh = max_wh; a.iou_thres = iou_thres;
    a.scale_boxes = 0; a.gain = 1.f; a.pad_x = a.pad_y = 0.f; a.orig_w = a.orig_h = 0.f; a.mode = 1;
    a.class_prune = 0;            // arbitrary caller boxes: no span guarantee
    return launch_nms(ctx, a, n, w, nullptr, out_keep, out_count, nullptr, st);
}
