// k6_rules.cu -- K6: the quadratic step of the reference's post-NMS rule engine on device.
//
// inference/postprocess.py:233-303 (_merge_overlapping_detections) groups the filtered detections by class, builds
// the pairwise IoU matrix with Python loops (:859-877, integer intersection / integer union in float64) and labels it
// with DBSCAN(metric='precomputed', eps = 1 - 0.3, min_samples = 1) on 1 - IoU.  The matrix diagonal is 0, so the
// distance of a box to itself is 1 > eps: a box is not its own neighbour, boxes without a partner come out as noise
// (label -1, all of them sharing that label), and the clusters are the connected components of the link graph
// `1.0 - iou <= 1.0 - 0.3`, numbered by their first member.  One CTA per image computes exactly those labels:
//   pairs (i < j) of the same class -> IoU in float64 (IEEE division, same comparison) -> union-find in shared memory
//   with minimum-index roots -> label = number of component roots of the class in front of mine.
// Rows that are not kept (filtered out before clustering) get -2.
//
// k6_chain runs the WHOLE detection chain of PostProcessor.process_results for a batch, one CTA per image:
//   refine (:111-184: bbox from the mask extent +-5 px :764-812, severity :814-840, area estimate for rows without a
//   mask) -> filter (:186-231) -> the clustering above -> cluster merge (:305-358: union box, np.mean of the
//   confidences in numpy's pairwise order, maximum severity, sum of the positive areas, base = first most confident
//   member) in the reference's output order (classes by first appearance, np.unique(labels) within a class) ->
//   validate (:360-464: clamp, area consistency, validation score, keep > 0.3).  All float64 with IEEE operations in
//   the reference's order; the host only turns the rows back into the reference's dictionaries (qcgraft/rules.py).
#include "common.cuh"

namespace {

constexpr int K6_THREADS = 256;
constexpr int K6_CAP_MAX = 2048;

__device__ __forceinline__ int k6_find(volatile int* L, int x) {
    int p = L[x];
    while (p != x) {
        const int gp = L[p];
        if (gp != p) L[x] = gp;
        x = p; p = gp;
    }
    return x;
}
__device__ __forceinline__ void k6_union(int* L, int a, int b) {
    bool done;
    do {
        a = k6_find(L, a);
        b = k6_find(L, b);
        if (a < b) { const int old = atomicMin(L + b, a); done = (old == b); b = old; }
        else if (b < a) { const int old = atomicMin(L + a, b); done = (old == a); a = old; }
        else done = true;
    } while (!done);
}

__global__ void __launch_bounds__(K6_THREADS)
k6_cluster(const int* __restrict__ box, const int* __restrict__ cls, const uint8_t* __restrict__ keep,
           const int* __restrict__ count, int cap, double max_dist, int* __restrict__ out_label) {
    extern __shared__ __align__(16) int k6sm[];
    int4* s_box = reinterpret_cast<int4*>(k6sm);            // [cap]
    int* s_cls = k6sm + 4 * cap;                            // [cap]  class, or -1 for rows that do not take part
    int* parent = s_cls + cap;                              // [cap]
    int* partner = parent + cap;                            // [cap]  1 when the row has at least one link
    const int img = blockIdx.x, tid = threadIdx.x;
    int k = count[img];
    k = k < 0 ? 0 : (k > cap ? cap : k);
    const long long base = (long long)img * cap;
    for (int i = tid; i < k; i += K6_THREADS) {
        s_box[i] = reinterpret_cast<const int4*>(box)[base + i];
        s_cls[i] = (!keep || keep[base + i]) ? cls[base + i] : -1;
        parent[i] = i;
        partner[i] = 0;
    }
    __syncthreads();
    // all pairs i < j: a warp takes row i, its lanes the partners j (uniform control flow, independent iterations);
    // only the few linked pairs diverge into the union
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NW = K6_THREADS / 32;
    for (int i = warp; i < k; i += NW) {
        const int ci = s_cls[i];
        if (ci < 0) continue;                                        // warp-uniform
        const int4 a = s_box[i];
        const long long area_a = (long long)(a.z - a.x) * (a.w - a.y);
        bool any_link = false;
        for (int j0 = i + 1; j0 < k; j0 += 32) {
            const int j = j0 + lane;
            bool link = false;
            if (j < k && s_cls[j] == ci) {
                const int4 b = s_box[j];
                const int x1 = max(a.x, b.x), y1 = max(a.y, b.y), x2 = min(a.z, b.z), y2 = min(a.w, b.w);
                if (x2 > x1 && y2 > y1) {
                    const long long inter = (long long)(x2 - x1) * (y2 - y1);
                    const long long uni = area_a + (long long)(b.z - b.x) * (b.w - b.y) - inter;
                    const double iou = uni > 0 ? __ddiv_rn((double)inter, (double)uni) : 0.0;
                    link = __dsub_rn(1.0, iou) <= max_dist;
                }
            }
            if (link) { partner[j] = 1; k6_union(parent, i, j); }
            any_link |= __any_sync(0xffffffffu, link);
        }
        if (any_link && lane == 0) partner[i] = 1;
    }
    __syncthreads();
    // label = number of component roots of my class in front of my root (roots are minimum indices and no longer
    // change; path halving only rewrites non-root entries)
    for (int i = warp; i < k; i += NW) {
        int lab = -2;
        if (s_cls[i] >= 0) {
            lab = -1;
            if (partner[i]) {
                const int r = k6_find(parent, i);
                const int cr = s_cls[r];
                int n = 0;
                for (int j0 = 0; j0 < r; j0 += 32) {
                    const int j = j0 + lane;
                    const bool root = j < r && parent[j] == j && partner[j] && s_cls[j] == cr;
                    n += __popc(__ballot_sync(0xffffffffu, root));
                }
                lab = n;
            }
        }
        if (lane == 0) out_label[base + i] = lab;
    }
    for (int i = k + tid; i < cap; i += K6_THREADS) out_label[base + i] = -2;
}


// ---- the whole chain ---------------------------------------------------------------------------------------------
struct K6Cursor { const int* rep; const double* conf; int me, k, pos; };
__device__ __forceinline__ double k6_next(K6Cursor& c) {          // next member of cluster `me` in row order
    while (c.rep[c.pos] != c.me) ++c.pos;
    return c.conf[c.pos++];
}
// numpy's pairwise_sum over the next n members (what np.mean(list_of_floats) adds up)
__device__ double k6_pairwise(K6Cursor& c, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, k6_next(c));
        return res;
    }
    if (n <= 128) {
        double r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = k6_next(c);
        int i = 8;
        for (; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], k6_next(c));
        }
        double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                               __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __dadd_rn(res, k6_next(c));
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    const double a = k6_pairwise(c, n2);
    return __dadd_rn(a, k6_pairwise(c, n - n2));
}

__global__ void __launch_bounds__(K6_THREADS)
k6_chain(const qc_rules_row* __restrict__ rows, const int* __restrict__ count, const int* __restrict__ shape, int cap,
         double min_conf, double max_area_pct, unsigned long long elongated_classes, double max_dist,
         qc_rules_refined* __restrict__ out_refined, qc_rules_out* __restrict__ out_rows, int* __restrict__ out_member,
         int* __restrict__ out_count) {
    extern __shared__ __align__(16) int k6sm[];
    int4* s_box = reinterpret_cast<int4*>(k6sm);             // [cap] refined boxes
    double* s_conf = reinterpret_cast<double*>(s_box + cap); // [cap]
    long long* s_area = reinterpret_cast<long long*>(s_conf + cap);   // [cap]
    int* s_cls = reinterpret_cast<int*>(s_area + cap);       // [cap] class, -1 = filtered out
    int* parent = s_cls + cap;
    int* partner = parent + cap;
    int* s_sev = partner + cap;
    int* s_rep = s_sev + cap;                                // cluster representative (its first member), -1 = filtered
    int* s_key = s_rep + cap;                                // first row of the class
    __shared__ int s_groups;
    const int img = blockIdx.x, tid = threadIdx.x;
    int k = count[img];
    k = k < 0 ? 0 : (k > cap ? cap : k);
    const int H = shape[2 * img], W = shape[2 * img + 1];
    const long long base = (long long)img * cap;
    if (tid == 0) s_groups = 0;
    // ---- refine + filter ------------------------------------------------------------------------------------------
    for (int i = tid; i < k; i += K6_THREADS) {
        const qc_rules_row r = rows[base + i];
        int x1 = r.x1, y1 = r.y1, x2 = r.x2, y2 = r.y2, sev = r.severity;
        long long area;
        double area_pct = 0.0;
        if (r.has_seg) {
            area = r.area_pixels;
            area_pct = r.area_pct;
            if (r.ext[0] >= 0) {                              // _refine_bbox_from_mask
                y1 = max(0, r.ext[0] - 5); x1 = max(0, r.ext[2] - 5);
                y2 = min(H, r.ext[1] + 5); x2 = min(W, r.ext[3] + 5);
            }
            bool done = false;                                // _refine_severity
            if (area_pct > 5.0) {
                if (sev == 1) { sev = 2; done = true; }
                else if (sev == 2) { sev = 3; done = true; }
            }
            if (!done && r.compactness < 0.3 && r.cls >= 0 && r.cls < 64 && ((elongated_classes >> r.cls) & 1ull) && sev == 1)
                sev = 2;
        } else {
            area = (long long)__dmul_rn((double)((long long)(r.x2 - r.x1) * (r.y2 - r.y1)), 0.6);   // int(w*h*0.6)
        }
        const int w = x2 - x1, h = y2 - y1;
        bool keep = !(r.conf < min_conf) && !(area_pct > max_area_pct);
        if (h > 0) {
            const double aspect = __ddiv_rn((double)w, (double)h);
            keep = keep && !(aspect > 10.0 || aspect < 0.1);
        } else {
            keep = false;                                     // aspect = inf
        }
        keep = keep && !(w < 5 || h < 5);
        s_box[i] = make_int4(x1, y1, x2, y2);
        s_conf[i] = r.conf;
        s_area[i] = area;
        s_sev[i] = sev;
        s_cls[i] = keep ? r.cls : -1;
        parent[i] = i;
        partner[i] = 0;
        qc_rules_refined o;
        o.x1 = x1; o.y1 = y1; o.x2 = x2; o.y2 = y2; o.severity = sev; o.kept = keep ? 1 : 0; o.area_pixels = area;
        out_refined[base + i] = o;
    }
    __syncthreads();
    // ---- same-class links (as k6_cluster) -----------------------------------------------------------------------
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NW = K6_THREADS / 32;
    for (int i = warp; i < k; i += NW) {
        const int ci = s_cls[i];
        if (ci < 0) continue;
        const int4 a = s_box[i];
        const long long area_a = (long long)(a.z - a.x) * (a.w - a.y);
        bool any_link = false;
        for (int j0 = i + 1; j0 < k; j0 += 32) {
            const int j = j0 + lane;
            bool link = false;
            if (j < k && s_cls[j] == ci) {
                const int4 b = s_box[j];
                const int x1 = max(a.x, b.x), y1 = max(a.y, b.y), x2 = min(a.z, b.z), y2 = min(a.w, b.w);
                if (x2 > x1 && y2 > y1) {
                    const long long inter = (long long)(x2 - x1) * (y2 - y1);
                    const long long uni = area_a + (long long)(b.z - b.x) * (b.w - b.y) - inter;
                    const double iou = uni > 0 ? __ddiv_rn((double)inter, (double)uni) : 0.0;
                    link = __dsub_rn(1.0, iou) <= max_dist;
                }
            }
            if (link) { partner[j] = 1; k6_union(parent, i, j); }
            any_link |= __any_sync(0xffffffffu, link);
        }
        if (any_link && lane == 0) partner[i] = 1;
    }
    __syncthreads();
    // ---- cluster representative = first member; order key = (first row of the class, noise first, representative) --
    for (int i = tid; i < k; i += K6_THREADS) {
        const int ci = s_cls[i];
        int rep = -1, firstc = -1;
        if (ci >= 0) {
            int cnt = 0, firstn = -1;
            for (int j = 0; j < k; ++j) {
                if (s_cls[j] != ci) continue;
                if (firstc < 0) firstc = j;
                if (firstn < 0 && !partner[j]) firstn = j;
                ++cnt;
            }
            rep = cnt == 1 ? i : (partner[i] ? k6_find(parent, i) : firstn);
        }
        s_rep[i] = rep;
        s_key[i] = firstc;
    }
    __syncthreads();
    for (int i = tid; i < k; i += K6_THREADS) {
        if (s_rep[i] != i) continue;
        // position in the reference's merged list: clusters with a smaller (class first row, noise-first, first member)
        const int kc = s_key[i], kn = partner[i] ? 1 : 0;
        int rank = 0, members = 0;
        for (int j = 0; j < k; ++j) {
            if (s_rep[j] == i) ++members;
            if (s_rep[j] != j || j == i) continue;
            const int jc = s_key[j], jn = partner[j] ? 1 : 0;
            if (jc < kc || (jc == kc && (jn < kn || (jn == kn && j < i)))) ++rank;
        }
        atomicAdd(&s_groups, 1);
        // ---- _merge_detection_cluster ---------------------------------------------------------------------------
        int4 b = s_box[i];
        double conf = s_conf[i];
        int sev = s_sev[i], basei = i;
        long long area = s_area[i];
        if (members > 1) {
            double best = s_conf[i];
            long long asum = 0;
            bool any_area = false;
            for (int j = i; j < k; ++j) {
                if (s_rep[j] != i) continue;
                const int4 q = s_box[j];
                b.x = min(b.x, q.x); b.y = min(b.y, q.y); b.z = max(b.z, q.z); b.w = max(b.w, q.w);
                if (s_conf[j] > best) { best = s_conf[j]; basei = j; }          // first maximum
                sev = max(sev, s_sev[j]);
                if (s_area[j] > 0) { asum += s_area[j]; any_area = true; }
            }
            K6Cursor c{s_rep, s_conf, i, k, i};
            conf = __ddiv_rn(k6_pairwise(c, members), (double)members);
            area = any_area ? asum : s_area[basei];
        }
        // ---- _validate_detections ---------------------------------------------------------------------------------
        int x1 = max(0, min(b.x, W - 1)), y1 = max(0, min(b.y, H - 1));
        int x2 = max(x1 + 1, min(b.z, W)), y2 = max(y1 + 1, min(b.w, H));
        const int w = x2 - x1, h = y2 - y1;
        const long long bbox_area = (long long)w * h;
        int flags = 0;
        if (area > bbox_area * 2) { area = bbox_area; flags |= 1; }
        const double cs = fmin(__ddiv_rn(conf, 0.8), 1.0);
        const double area_ratio = __ddiv_rn((double)bbox_area, (double)((long long)H * W));
        const double ss = (0.001 <= area_ratio && area_ratio <= 0.5) ? 1.0 : 0.5;
        const double aspect = __ddiv_rn((double)w, (double)max(h, 1));
        const double as = (0.2 <= aspect && aspect <= 5.0) ? 1.0 : 0.5;
        const qc_rules_row rb = rows[base + basei];
        const double gs = rb.has_seg ? fmin(__ddiv_rn(rb.seg_conf, 0.5), 1.0) : 0.7;
        const double nx = __ddiv_rn(__ddiv_rn((double)(x1 + x2), 2.0), (double)W);
        const double ny = __ddiv_rn(__ddiv_rn((double)(y1 + y2), 2.0), (double)H);
        const double ls = fmin(fmin(__ddiv_rn(fmin(nx, __dsub_rn(1.0, nx)), 0.05), __ddiv_rn(fmin(ny, __dsub_rn(1.0, ny)), 0.05)), 1.0);
        double score = __dmul_rn(0.3, cs);
        score = __dadd_rn(score, __dmul_rn(0.2, ss));
        score = __dadd_rn(score, __dmul_rn(0.2, as));
        score = __dadd_rn(score, __dmul_rn(0.2, gs));
        score = __dadd_rn(score, __dmul_rn(0.1, ls));
        qc_rules_out o;
        o.base = basei; o.merged_from = members; o.x1 = x1; o.y1 = y1; o.x2 = x2; o.y2 = y2; o.severity = sev;
        o.flags = flags | (score > 0.3 ? 2 : 0);
        o.area_pixels = area; o.conf = conf; o.validation_score = score;
        out_rows[base + rank] = o;
    }
    __syncthreads();
    for (int i = tid; i < cap; i += K6_THREADS) {
        int m = -1;
        if (i < k && s_rep[i] >= 0) {
            const int r = s_rep[i];
            // rank of my representative: recompute (cheap) instead of keeping another array
            const int kc = s_key[r], kn = partner[r] ? 1 : 0;
            int rank = 0;
            for (int j = 0; j < k; ++j) {
                if (s_rep[j] != j || j == r) continue;
                const int jc = s_key[j], jn = partner[j] ? 1 : 0;
                if (jc < kc || (jc == kc && (jn < kn || (jn == kn && j < r)))) ++rank;
            }
            m = rank;
        }
        out_member[base + i] = m;
    }
    if (tid == 0) out_count[img] = s_groups;
}

}  // namespace

extern "C" int qc_rules_cluster(qc_ctx* ctx, const int32_t* box, const int32_t* cls, const uint8_t* keep,
                                const int32_t* count, int n, int cap, double overlap_threshold, int32_t* out_label,
                                void* stream) {
    qc_device_guard dev_guard(ctx);
    if (!ctx) return QC_ERR_INVALID;
    QC_ENTER(ctx, stream);
    if (n == 0) return QC_OK;
    if (!box || !cls || !count || !out_label || n < 0 || cap <= 0)
        QC_FAIL(ctx, QC_ERR_INVALID, "qc_rules_cluster: bad argument");
    if (cap > K6_CAP_MAX) QC_FAIL(ctx, QC_ERR_UNSUPPORTED, "qc_rules_cluster: cap %d > %d", cap, K6_CAP_MAX);
    if (reinterpret_cast<uintptr_t>(box) & 15) QC_FAIL(ctx, QC_ERR_INVALID, "qc_rules_cluster: box must be 16-byte aligned");
    const size_t smem = (size_t)cap * 28;
    if (int rc = qc_prepare(ctx, k6_cluster, (int)smem, K6_THREADS, nullptr)) return rc;
    // the reference compares `1.0 - iou` against `eps = 1.0 - overlap_threshold` (both float64)
    k6_cluster<<<n, K6_THREADS, smem, (cudaStream_t)stream>>>(box, cls, keep, count, cap, 1.0 - overlap_threshold, out_label);
    QC_LAUNCHED_AS(ctx, "k6_cluster", stream);
    return QC_OK;
}

extern "C" int qc_rules_apply(qc_ctx* ctx, const qc_rules_row* rows, const int32_t* count, const int32_t* shapes, int n,
                              int cap, double min_conf, double max_area_pct, uint64_t elongated_classes,
                              double overlap_threshold, qc_rules_refined* out_refined, qc_rules_out* out_rows,
                              int32_t* out_member, int32_t* out_count, void* stream) {
    qc_device_guard dev_guard(ctx);
    if (!ctx) return QC_ERR_INVALID;
    QC_ENTER(ctx, stream);
    if (n == 0) return QC_OK;
    if (!rows || !count || !shapes || !out_refined || !out_rows || !out_member || !out_count || n < 0 || cap <= 0)
        QC_FAIL(ctx, QC_ERR_INVALID, "qc_rules_apply: bad argument");
    if (cap > K6_CAP_MAX) QC_FAIL(ctx, QC_ERR_UNSUPPORTED, "qc_rules_apply: cap %d > %d", cap, K6_CAP_MAX);
    const size_t smem = (size_t)cap * (16 + 8 + 8 + 6 * 4);
    if (int rc = qc_prepare(ctx, k6_chain, (int)smem, K6_THREADS, nullptr)) return rc;
    k6_chain<<<n, K6_THREADS, smem, (cudaStream_t)stream>>>(rows, count, shapes, cap, min_conf, max_area_pct,
                                                            (unsigned long long)elongated_classes, 1.0 - overlap_threshold,
                                                            out_refined, out_rows, out_member, out_count);
    QC_LAUNCHED_AS(ctx, "k6_chain", stream);
    return QC_OK;
}
