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
// Rows that are not kept (filtered out before clustering) get -2.  Everything else of the rule chain is O(k) per
// image and stays with the host tables of qcgraft/rules.py.
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
