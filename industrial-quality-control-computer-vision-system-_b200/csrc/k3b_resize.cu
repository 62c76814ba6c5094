// k3b_resize.cu -- K3b: the whole-frame classifier input of the reference, bit-exact.
//
// models/resent_model.py:57-66,114-130 feeds ResNet-50 with
//     PIL.Image.fromarray(rgb) -> transforms.Resize((224,224)) -> ToTensor() -> Normalize(mean,std)
// transforms.Resize on a PIL image is Pillow's antialiased separable resampler (Image.resize BILINEAR):
// float64 triangle coefficients normalised per output index, quantised to 22-bit integers, a horizontal pass
// then a vertical pass, each accumulated in int32 from 2^21 and clipped to uint8.
//
//   qc_pil_resize_make_tables   host: Pillow's precompute_coeffs + normalize_coeffs_8bpc for both axes
//   k3b_resize                  one CTA per (image, band of output rows): horizontal pass of the source rows the
//                               band taps into a shared-memory uint8 tile, vertical pass from that tile, then
//                               ToTensor + Normalize through a 256x3 table built with the reference's float32
//                               ops (x/255, -mean, /std: IEEE divisions), bf16 (RNE) or f32 NCHW stores.
// Integer work, HBM-bound: algorithmic bytes per image = h*w*3 (read once) + 3*out_h*out_w*sizeof(out).
#include "common.cuh"
#include <math.h>
#include <algorithm>
#include <vector>

namespace {

constexpr int K3B_THREADS = 256;
constexpr int K3B_PREC = 32 - 8 - 2;          // Pillow PRECISION_BITS
constexpr int K3B_HDR = 8;                    // table header ints

// table layout (int32): [0]=src_h [1]=src_w [2]=dst_h [3]=dst_w [4]=ksize_x [5]=ksize_y [6]=band_rows [7]=max_src_rows
//   bounds_x[dst_w][2] | kk_x[dst_w][ksize_x] | bounds_y[dst_h][2] | kk_y[dst_h][ksize_y]
struct Tables {
    int src_h, src_w, dst_h, dst_w, kx, ky, band_rows, max_src_rows;
    size_t off_bx, off_kx, off_by, off_ky, total_ints;
};

int ksize_of(int in_size, int out_size) {
    const double scale = (double)in_size / (double)out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    return (int)ceil(filterscale) * 2 + 1;       // support = 1.0 * filterscale (bilinear)
}

Tables layout(int src_h, int src_w, int dst_h, int dst_w) {
    Tables t;
    t.src_h = src_h; t.src_w = src_w; t.dst_h = dst_h; t.dst_w = dst_w;
    t.kx = ksize_of(src_w, dst_w);
    t.ky = ksize_of(src_h, dst_h);
    t.band_rows = 0; t.max_src_rows = 0;
    size_t o = K3B_HDR;
    t.off_bx = o; o += (size_t)dst_w * 2;
    t.off_kx = o; o += (size_t)dst_w * t.kx;
    t.off_by = o; o += (size_t)dst_h * 2;
    t.off_ky = o; o += (size_t)dst_h * t.ky;
    t.total_ints = o;
    return t;
}

// Pillow Resample.c precompute_coeffs (bilinear filter, box = whole axis) + normalize_coeffs_8bpc
void axis_coeffs(int in_size, int out_size, int ksize, int* bounds, int* kk) {
    const double scale = (double)in_size / (double)out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 1.0 * filterscale;
    const double ss = 1.0 / filterscale;
    std::vector<double> w((size_t)ksize);
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = (xx + 0.5) * scale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            double a = (x + xmin - center + 0.5) * ss;
            if (a < 0.0) a = -a;
            const double v = a < 1.0 ? 1.0 - a : 0.0;
            w[x] = v;
            ww += v;
        }
        for (int x = 0; x < ksize; ++x) {
            double v = 0.0;
            if (x < xmax) v = ww != 0.0 ? w[x] / ww : w[x];
            kk[(size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << K3B_PREC)) : (int)(0.5 + v * (1 << K3B_PREC));
        }
        bounds[xx * 2] = xmin;
        bounds[xx * 2 + 1] = xmax;
    }
}

// Horizontal taps of one output column cover 3*cnt CONTIGUOUS source bytes.  The CTA stages source rows in shared
// memory with 128-bit loads (smem offset == global address mod 16), and a thread (= output column) reads its window
// as 32-bit words: four taps = 12 bytes = three words after one funnel shift per word, so a row costs 3-4 LDS.32 per
// four taps instead of 12 byte loads -- the byte-load form is LSU-bound.  Coefficient rows are padded with zeros to
// a multiple of four taps (NG groups); for NG <= 2 they live in registers.
// staged row: 3*w data bytes + 16 B alignment slack + the words the zero-coefficient taps of the last group may reach
__host__ __device__ inline int k3b_row_buf(int src_w, int ng) { return (src_w * 3 + 32 + 12 * ng + 15) & ~15; }

// byte `shift/8` of four coefficients, packed for dp4a
__device__ __forceinline__ unsigned k3b_pack(const int (&k)[4], int shift) {
    return ((unsigned)(k[0] >> shift) & 255u) | (((unsigned)(k[1] >> shift) & 255u) << 8) |
           (((unsigned)(k[2] >> shift) & 255u) << 16) | (((unsigned)(k[3] >> shift) & 255u) << 24);
}

template <typename OutT, int NG>              // NG: 4-tap groups per output column (0 = runtime, coefficients in smem)
__global__ void __launch_bounds__(K3B_THREADS)
k3b_resize(const uint8_t* __restrict__ imgs, long long img_stride, int row_stride, const int* __restrict__ tab,
           int swap_rb, const float* __restrict__ norm, OutT* __restrict__ dst, int stage_rows, int ng_rt,
           int band_rows, int max_src_rows) {
    extern __shared__ __align__(16) uint8_t k3b_sm[];
    __shared__ float s_lut[3][256];
    const int src_w = tab[1], dst_h = tab[2], dst_w = tab[3], kx = tab[4], ky = tab[5];
    const int* __restrict__ bx = tab + K3B_HDR;
    const int* __restrict__ kkx = bx + dst_w * 2;
    const int* __restrict__ by = kkx + dst_w * kx;
    const int* __restrict__ kky = by + dst_h * 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int img = blockIdx.y;
    const int oy0 = blockIdx.x * band_rows, oy1 = min(oy0 + band_rows, dst_h);
    const int pitch = (dst_w + 3) & ~3;                    // bytes per channel row of the tile
    const int ng = NG ? NG : ng_rt;
    const int row_buf = k3b_row_buf(src_w, ng);
    uint8_t* tile = k3b_sm;                                // [max_src_rows][3][pitch]
    uint8_t* stage = tile + (((size_t)max_src_rows * 3 * pitch + 15) & ~(size_t)15);   // [2][stage_rows][row_buf]
    int* kx_s = reinterpret_cast<int*>(stage + (size_t)2 * stage_rows * row_buf);        // [dst_w][ng*4] (NG == 0)

    // ToTensor + Normalize per OUTPUT channel: ((float)v / div - mean) / std with the reference's float32 ops
    for (int i = tid; i < 768; i += K3B_THREADS) {
        const int c = i >> 8, v = i & 255;
        s_lut[c][v] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, norm[c * 3 + 0]), norm[c * 3 + 1]), norm[c * 3 + 2]);
    }
    if (NG == 0) {
        // per (column, 4-tap group): the three byte planes of the four coefficients (see below), one pad word
        for (int i = tid; i < dst_w * ng; i += K3B_THREADS) {
            const int ox = i / ng, g = i - ox * ng;
            int k[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) k[t] = 4 * g + t < kx ? __ldg(kkx + ox * kx + 4 * g + t) : 0;
            uint4 pk;
            pk.x = k3b_pack(k, 0); pk.y = k3b_pack(k, 8); pk.z = k3b_pack(k, 16); pk.w = 0u;
            reinterpret_cast<uint4*>(kx_s)[i] = pk;
        }
    }
    // source rows this band taps (bounds are monotone in the output index)
    const int sy0 = by[oy0 * 2], sy1 = by[(oy1 - 1) * 2] + by[(oy1 - 1) * 2 + 1];
    const int nsrc = min(sy1 - sy0, max_src_rows);
    const uint8_t* ibase = imgs + (long long)img * img_stride;

    // ---- horizontal pass: source rows are staged stage_rows at a time, double-buffered with cp.async so the
    // next rows stream in while the current ones are filtered ---------------------------------------------
    auto issue_stage = [&](int r0, uint8_t* buf) {
        const int nr = min(stage_rows, nsrc - r0);
        // warp per row; 16-byte async copies for the aligned interior, bytes for the ragged ends; the tail is
        // zero-filled so the zero-coefficient taps of the last group read defined bytes
        for (int rr = warp; rr < nr; rr += K3B_THREADS / 32) {
            const uint8_t* p = ibase + (long long)(sy0 + r0 + rr) * row_stride;
            const int off = (int)(reinterpret_cast<uintptr_t>(p) & 15);
            uint8_t* sb = buf + (size_t)rr * row_buf;              // sb[off + i] = p[i]
            const int nbytes = src_w * 3;
            const int head = min(nbytes, (16 - off) & 15);
            const int nvec = (nbytes - head) >> 4;
            for (int i = lane; i < head; i += 32) sb[off + i] = __ldg(p + i);
            const uint8_t* pv = p + head;
            const uint32_t sv = qc_smem_u32(sb + off + head);
            for (int i = lane; i < nvec; i += 32)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sv + 16u * i), "l"(pv + 16 * (size_t)i) : "memory");
            for (int i = head + nvec * 16 + lane; i < nbytes; i += 32) sb[off + i] = __ldg(p + i);
            for (int i = off + nbytes + lane; i < row_buf; i += 32) sb[i] = 0;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // Coefficients are non-negative 22-bit integers: k = b0 + 256 b1 + 65536 b2.  With the four taps' bytes of one
    // plane packed in a word, a channel's four products are ONE dp4a per plane (3 per channel and group instead of 4
    // multiply-adds plus the byte extracts); the planes are recombined by shifts at the end -- the same integer sum, in
    // a different order.  A thread keeps the packed planes of its output column for the whole band.
    unsigned kb[NG ? NG : 1][3];
    auto load_kb = [&](int ox) {
#pragma unroll
        for (int g = 0; g < (NG ? NG : 1); ++g) {
            int k[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) k[t] = (ox < dst_w && 4 * g + t < kx) ? __ldg(kkx + ox * kx + 4 * g + t) : 0;
            kb[g][0] = k3b_pack(k, 0); kb[g][1] = k3b_pack(k, 8); kb[g][2] = k3b_pack(k, 16);
        }
    };
    if (NG) load_kb(tid);
    const int nstage = (nsrc + stage_rows - 1) / stage_rows;
    issue_stage(0, stage);
    for (int sti = 0; sti < nstage; ++sti) {
        const int r0 = sti * stage_rows;
        const int nr = min(stage_rows, nsrc - r0);
        uint8_t* cur = stage + (size_t)(sti & 1) * stage_rows * row_buf;
        if (sti + 1 < nstage) {
            issue_stage(r0 + stage_rows, stage + (size_t)((sti + 1) & 1) * stage_rows * row_buf);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();                                   // stage sti landed (and tables / LUT written)
        for (int ox = tid; ox < dst_w; ox += K3B_THREADS) {
            const int x0 = bx[ox * 2];
            if (NG && dst_w > K3B_THREADS) load_kb(ox);        // only when the output is wider than the CTA
            for (int rr = 0; rr < nr; ++rr) {
                const uint8_t* p = ibase + (long long)(sy0 + r0 + rr) * row_stride;
                const int s = (int)(reinterpret_cast<uintptr_t>(p) & 15) + x0 * 3;      // window start in the staged row
                const unsigned* wp = reinterpret_cast<const unsigned*>(cur + (size_t)rr * row_buf) + (s >> 2);
                const unsigned sh = (unsigned)(s & 3) * 8u;
                unsigned acc[3][3] = {{0u, 0u, 0u}, {0u, 0u, 0u}, {0u, 0u, 0u}};        // [channel][byte plane]
                unsigned wprev = wp[0];
#pragma unroll
                for (int g = 0; g < (NG ? NG : 1); ++g) {
                    for (int gg = g; gg < (NG ? g + 1 : ng); ++gg) {
                        const unsigned w1 = wp[3 * gg + 1], w2 = wp[3 * gg + 2], w3 = wp[3 * gg + 3];
                        const unsigned A0 = __funnelshift_r(wprev, w1, sh), A1 = __funnelshift_r(w1, w2, sh),
                                       A2 = __funnelshift_r(w2, w3, sh);
                        wprev = w3;
                        unsigned k0, k1, k2;
                        if (NG) { k0 = kb[g][0]; k1 = kb[g][1]; k2 = kb[g][2]; }
                        else {
                            const uint4 kv = *reinterpret_cast<const uint4*>(kx_s + (ox * ng + gg) * 4);
                            k0 = kv.x; k1 = kv.y; k2 = kv.z;
                        }
                        // 12 bytes = 4 RGB pixels: channel c of pixel t is byte 3t + c
                        const unsigned c0 = __byte_perm(__byte_perm(A0, A1, 0x0630), A2, 0x5210);   // bytes 0,3,6,9
                        const unsigned c1 = __byte_perm(__byte_perm(A0, A1, 0x0741), A2, 0x6210);   // bytes 1,4,7,10
                        const unsigned c2 = __byte_perm(__byte_perm(A0, A1, 0x0052), A2, 0x7410);   // bytes 2,5,8,11
                        acc[0][0] = __dp4a(c0, k0, acc[0][0]); acc[0][1] = __dp4a(c0, k1, acc[0][1]); acc[0][2] = __dp4a(c0, k2, acc[0][2]);
                        acc[1][0] = __dp4a(c1, k0, acc[1][0]); acc[1][1] = __dp4a(c1, k1, acc[1][1]); acc[1][2] = __dp4a(c1, k2, acc[1][2]);
                        acc[2][0] = __dp4a(c2, k0, acc[2][0]); acc[2][1] = __dp4a(c2, k1, acc[2][1]); acc[2][2] = __dp4a(c2, k2, acc[2][2]);
                    }
                }
                const int a0 = (1 << (K3B_PREC - 1)) + (int)(acc[0][0] + (acc[0][1] << 8) + (acc[0][2] << 16));
                const int a1 = (1 << (K3B_PREC - 1)) + (int)(acc[1][0] + (acc[1][1] << 8) + (acc[1][2] << 16));
                const int a2 = (1 << (K3B_PREC - 1)) + (int)(acc[2][0] + (acc[2][1] << 8) + (acc[2][2] << 16));
                uint8_t* t = tile + (size_t)(r0 + rr) * 3 * pitch + ox;
                t[0] = (uint8_t)min(max(a0 >> K3B_PREC, 0), 255);
                t[pitch] = (uint8_t)min(max(a1 >> K3B_PREC, 0), 255);
                t[2 * pitch] = (uint8_t)min(max(a2 >> K3B_PREC, 0), 255);
            }
        }
        __syncthreads();                                   // done with `cur` before stage sti+2 overwrites it
    }

    // ---- vertical pass + ToTensor + Normalize: 4 consecutive columns of one (row, channel) per thread -----
    const int quads = pitch >> 2;
    const int per_row = 3 * quads;
    const long long plane = (long long)dst_h * dst_w;
    const bool vec_ok = (dst_w & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
    for (int it = tid; it < (oy1 - oy0) * per_row; it += K3B_THREADS) {
        const int ry = it / per_row, rem = it - ry * per_row;
        const int c = rem / quads, q = rem - c * quads;
        const int oy = oy0 + ry;
        const int y0 = by[oy * 2] - sy0, cnt = by[oy * 2 + 1];
        const int* k = kky + oy * ky;
        int a0 = 1 << (K3B_PREC - 1), a1 = a0, a2 = a0, a3 = a0;
        const uint8_t* t = tile + (size_t)y0 * 3 * pitch + c * pitch + 4 * q;
        for (int y = 0; y < cnt; ++y) {
            const int kv = __ldg(k + y);
            const unsigned v4 = *reinterpret_cast<const unsigned*>(t + (size_t)y * 3 * pitch);
            a0 += (int)(v4 & 255u) * kv;
            a1 += (int)((v4 >> 8) & 255u) * kv;
            a2 += (int)((v4 >> 16) & 255u) * kv;
            a3 += (int)(v4 >> 24) * kv;
        }
        const int oc = swap_rb ? 2 - c : c;                 // tile channel c is source byte c
        const float* lut = s_lut[oc];
        const float f0 = lut[min(max(a0 >> K3B_PREC, 0), 255)], f1 = lut[min(max(a1 >> K3B_PREC, 0), 255)];
        const float f2 = lut[min(max(a2 >> K3B_PREC, 0), 255)], f3 = lut[min(max(a3 >> K3B_PREC, 0), 255)];
        OutT* o = dst + ((long long)img * 3 + oc) * plane + (long long)oy * dst_w + 4 * q;
        if (vec_ok) {
            if (sizeof(OutT) == 2) {
                const __nv_bfloat162 p0 = __floats2bfloat162_rn(f0, f1), p1 = __floats2bfloat162_rn(f2, f3);
                *reinterpret_cast<uint2*>(o) = make_uint2(*reinterpret_cast<const unsigned*>(&p0),
                                                          *reinterpret_cast<const unsigned*>(&p1));
            } else {
                *reinterpret_cast<float4*>(o) = make_float4(f0, f1, f2, f3);
            }
        } else {
            const float f[4] = {f0, f1, f2, f3};
            for (int j = 0; j < 4 && 4 * q + j < dst_w; ++j) {
                if (sizeof(OutT) == 2) reinterpret_cast<uint16_t*>(o)[j] = qc_f2bf(f[j]);
                else reinterpret_cast<float*>(o)[j] = f[j];
            }
        }
    }
}

}  // namespace

extern "C" size_t qc_pil_resize_table_bytes(int src_h, int src_w, int dst_h, int dst_w) {
    if (src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0) return 0;
    return layout(src_h, src_w, dst_h, dst_w).total_ints * sizeof(int32_t);
}

extern "C" int qc_pil_resize_make_tables(int src_h, int src_w, int dst_h, int dst_w, int32_t* host_tables) {
    if (src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0 || !host_tables) return QC_ERR_INVALID;
    Tables t = layout(src_h, src_w, dst_h, dst_w);
    axis_coeffs(src_w, dst_w, t.kx, host_tables + t.off_bx, host_tables + t.off_kx);
    axis_coeffs(src_h, dst_h, t.ky, host_tables + t.off_by, host_tables + t.off_ky);
    // output rows per CTA: as many as keep the uint8 tile of tapped source rows within ~64 KB (at least one)
    const int* by = host_tables + t.off_by;
    const int pitch = (dst_w + 3) & ~3;
    const int budget_rows = std::max(1, (64 * 1024) / (3 * pitch));
    int band = std::min(dst_h, 16), max_rows = 0;
    for (;; --band) {
        max_rows = 0;
        for (int o0 = 0; o0 < dst_h; o0 += band) {
            const int o1 = std::min(o0 + band, dst_h) - 1;
            max_rows = std::max(max_rows, by[o1 * 2] + by[o1 * 2 + 1] - by[o0 * 2]);
        }
        if (max_rows <= budget_rows || band == 1) break;
    }
    host_tables[0] = src_h; host_tables[1] = src_w; host_tables[2] = dst_h; host_tables[3] = dst_w;
    host_tables[4] = t.kx; host_tables[5] = t.ky; host_tables[6] = band; host_tables[7] = max_rows;
    return QC_OK;
}

extern "C" int qc_pil_resize(qc_ctx* ctx, const uint8_t* imgs, int n, int h, int w, int64_t img_stride,
                             int32_t row_stride, const int32_t* dev_tables, const int32_t* host_tables, int swap_rb,
                             const float* norm, int out_dtype, void* dst, void* stream) {
    qc_device_guard dev_guard(ctx);
    if (!ctx) return QC_ERR_INVALID;
    QC_ENTER(ctx, stream);
    if (n == 0) return QC_OK;
    if (!imgs || !dev_tables || !host_tables || !norm || !dst || n < 0 || h <= 0 || w <= 0 || row_stride < w * 3)
        QC_FAIL(ctx, QC_ERR_INVALID, "qc_pil_resize: bad argument");
    if (host_tables[0] != h || host_tables[1] != w) QC_FAIL(ctx, QC_ERR_INVALID, "qc_pil_resize: tables are for %dx%d", host_tables[0], host_tables[1]);
    if (n > 65535) QC_FAIL(ctx, QC_ERR_UNSUPPORTED, "qc_pil_resize: n > 65535 (split the call)");
    const int dst_h = host_tables[2], dst_w = host_tables[3], kx = host_tables[4], ky = host_tables[5];
    const int pitch = (dst_w + 3) & ~3;
    const int ng = (kx + 3) / 4;
    const int row_buf = k3b_row_buf(w, ng);
    const size_t coef_bytes = ng <= 2 ? 0 : (size_t)dst_w * ng * 16;
    // Output rows per CTA, chosen per launch: a band costs its tapped source rows (horizontal pass) plus its own rows
    // (vertical pass), and the grid should fill whole waves of the CTAs that fit an SM at that band's tile size.
    Tables tl = layout(h, w, dst_h, dst_w);
    const int* by = host_tables + tl.off_by;
    int band = 0, max_rows = 0, stage_rows = 4;
    size_t smem = 0;
    double best_cost = 0.0;
    (void)ky;
    for (int b = 4; b <= std::min(dst_h, 48); ++b) {
        int mr = 0;
        for (int o0 = 0; o0 < dst_h; o0 += b) {
            const int o1 = std::min(o0 + b, dst_h) - 1;
            mr = std::max(mr, by[o1 * 2] + by[o1 * 2 + 1] - by[o0 * 2]);
        }
        const size_t tile_b = ((size_t)mr * 3 * pitch + 15) & ~(size_t)15;
        int sr = 4;
        while (sr > 1 && tile_b + (size_t)2 * sr * row_buf + coef_bytes + 4096 > (size_t)ctx->max_smem_optin / 2) sr >>= 1;
        const size_t sm = tile_b + (size_t)2 * sr * row_buf + coef_bytes;
        if ((long long)sm + 4096 > ctx->max_smem_optin) continue;
        const int occ = (int)std::max<size_t>(1, std::min<size_t>(2048 / K3B_THREADS, (size_t)(228 * 1024) / (sm + 1024)));
        const long long ctas = (long long)((dst_h + b - 1) / b) * n;
        const long long waves = (ctas + (long long)occ * ctx->sm_count - 1) / ((long long)occ * ctx->sm_count);
        // per-CTA work ~ source rows filtered horizontally + a third of that per output row of the vertical pass
        // (the SM is issue-bound: CTAs resident together share it, so a wave takes occ times one CTA's work; below
        // three resident CTAs the staging latency is no longer hidden)
        const double cost = (double)waves * occ * ((double)mr + 0.35 * b + 6.0) * (occ >= 3 ? 1.0 : 1.3);
        if (band == 0 || cost < best_cost) { band = b; max_rows = mr; stage_rows = sr; smem = sm; best_cost = cost; }
    }
    if (band == 0)
        QC_FAIL(ctx, QC_ERR_UNSUPPORTED, "qc_pil_resize: %dx%d -> %dx%d does not fit shared memory", h, w, dst_h, dst_w);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)((dst_h + band - 1) / band), (unsigned)n);
#define K3B_LAUNCH(T, G)                                                                                          \
    do {                                                                                                          \
        if (int rc = qc_prepare(ctx, k3b_resize<T, G>, (int)smem, K3B_THREADS, nullptr)) return rc;               \
        k3b_resize<T, G><<<grid, K3B_THREADS, smem, st>>>(imgs, img_stride, row_stride, dev_tables, swap_rb, norm, \
                                                          (T*)dst, stage_rows, ng, band, max_rows);               \
    } while (0)
#define K3B_LAUNCH_NG(T)                          \
    do {                                          \
        if (ng == 1) K3B_LAUNCH(T, 1);            \
        else if (ng == 2) K3B_LAUNCH(T, 2);       \
        else K3B_LAUNCH(T, 0);                    \
    } while (0)
    if (out_dtype == QC_DT_BF16) K3B_LAUNCH_NG(__nv_bfloat16);
    else if (out_dtype == QC_DT_F32) K3B_LAUNCH_NG(float);
    else QC_FAIL(ctx, QC_ERR_INVALID, "qc_pil_resize: bad out_dtype");
    QC_LAUNCHED_AS(ctx, "k3b_resize", st);
    return QC_OK;
}
