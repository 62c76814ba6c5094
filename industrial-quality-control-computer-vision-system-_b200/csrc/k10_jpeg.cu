// k10_jpeg.cu -- K10: batched baseline-JPEG decode, bit-exact with cv2.imdecode(buf, IMREAD_COLOR).
//
// The reference decodes every upload on the host: app.py:190-194 wraps the file bytes in a 1-D uint8 array,
// inference/detector.py:208-234 (and models/yolo_model.py:74-76) run cv2.imdecode(..., IMREAD_COLOR) on it and then
// cv2.cvtColor(BGR2RGB).  cv2 4.13 links libjpeg-turbo 3.1; with its defaults that decoder is
//     Huffman decode -> dequantise + "islow" 13-bit integer IDCT (jidctint.c) -> "fancy" triangle chroma upsampling
//     (jdsample.c h2v1/h2v2_fancy_upsample; plain replication when the chroma plane is <= 2 samples wide)
//     -> integer YCbCr->RGB tables (jdcolor.c, SCALEBITS 16)
// all integer arithmetic with published rounding, restated here (oracle/jpeg.py restates it in numpy; both are pinned
// against cv2.imdecode).  Split of the work:
//   host  (qc_jpeg_info_host, entropy stage of qc_jpeg_decode): marker parsing and the Huffman bit stream -- serial by
//         construction within a restart interval -- run by a few host threads, one image each, into a caller-provided
//         pinned staging buffer as int16 coefficient blocks in natural order (not dequantised);
//   device (k10_idct, k10_color): one H2D copy of the staging buffer, then one thread per 8x8 block for the IDCT into
//         uint8 component planes, then upsampling + colour conversion + the reference's BGR2RGB (optional) into packed
//         HWC frames -- the layout K1/K3/K7 read.
// Unsupported streams (progressive, arithmetic coding, 12-bit, CMYK/RGB-coded, 4:4:0/4:1:1, multi-scan, EXIF
// orientation other than 1) return QC_ERR_UNSUPPORTED with a message; there is no fallback decoder in the library.
#include "common.cuh"
#include <string.h>
#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

namespace {

constexpr int K10_LOOK = 9;

const uint8_t k10_natural[64 + 16] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,
    6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
    39, 46, 53, 60, 61, 54, 47, 55, 62, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};

struct HuffTable {
    bool present = false;
    uint16_t lut[1 << K10_LOOK];      // (length << 8) | symbol for codes of <= K10_LOOK bits, 0 otherwise
    int maxcode[18];                  // largest code of each length, -1 if none
    int valoff[18];                   // index of the first symbol of a length minus its first code
    uint8_t vals[256];
    // AC tables only: when a code and the value bits that follow it fit the look-ahead together, one lookup yields
    // (value << 8) | (run << 4) | total bits; 0 = take the two-step path
    int16_t fast[1 << K10_LOOK];
};

struct Header {
    int w = 0, h = 0, nc = 0;
    int cid[3], hs[3], vs[3], tq[3], td[3], ta[3];
    uint16_t qt[4][64];
    bool qt_ok[4] = {false, false, false, false};
    HuffTable dc[4], ac[4];
    int restart = 0;
    size_t scan = 0;                  // offset of the first entropy-coded byte
    bool have_sof = false, have_sos = false, jfif = false, adobe = false;
    int adobe_transform = -1, orientation = 1;
};

// jdhuff.c jpeg_make_d_derived_tbl: canonical code assignment
bool build_huff(HuffTable& t, const uint8_t counts[16], const uint8_t* vals, int nvals) {
    memset(t.lut, 0, sizeof(t.lut));
    memcpy(t.vals, vals, (size_t)nvals);
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
        const int n = counts[l - 1];
        t.valoff[l] = k - code;
        if (n) {
            if (code + n > (1 << l)) return false;
            for (int i = 0; i < n; ++i, ++k, ++code) {
                if (l <= K10_LOOK) {
                    const int base = code << (K10_LOOK - l);
                    for (int f = 0; f < (1 << (K10_LOOK - l)); ++f) t.lut[base + f] = (uint16_t)((l << 8) | vals[k]);
                }
            }
            t.maxcode[l] = code - 1;
        } else {
            t.maxcode[l] = -1;
        }
        code <<= 1;
    }
    t.maxcode[17] = 0x7fffffff;
    for (int i = 0; i < (1 << K10_LOOK); ++i) {
        t.fast[i] = 0;
        const unsigned e = t.lut[i];
        if (!e) continue;
        const int len = (int)(e >> 8), run = (int)((e >> 4) & 15u), mag = (int)(e & 15u);
        if (mag == 0 || len + mag > K10_LOOK) continue;
        int v = ((i << len) & ((1 << K10_LOOK) - 1)) >> (K10_LOOK - mag);
        if (v < (1 << (mag - 1))) v = v - (1 << mag) + 1;
        if (v >= -128 && v <= 127) t.fast[i] = (int16_t)(v * 256 + run * 16 + len + mag);
    }
    t.present = true;
    return true;
}

inline int rd16(const uint8_t* p) { return (p[0] << 8) | p[1]; }

// EXIF orientation (tag 0x0112 of IFD0): cv2.imdecode(IMREAD_COLOR) rotates by it, so anything but 1 is refused
int exif_orientation(const uint8_t* p, size_t n) {
    if (n < 14 || memcmp(p, "Exif\0\0", 6) != 0) return 1;
    const uint8_t* t = p + 6;
    n -= 6;
    const bool le = t[0] == 'I' && t[1] == 'I', be = t[0] == 'M' && t[1] == 'M';
    if (!le && !be) return 1;
    auto u16 = [&](size_t o) -> unsigned { return le ? (unsigned)(t[o] | (t[o + 1] << 8)) : (unsigned)((t[o] << 8) | t[o + 1]); };
    auto u32 = [&](size_t o) -> unsigned {
        return le ? (unsigned)(t[o] | (t[o + 1] << 8) | (t[o + 2] << 16) | ((unsigned)t[o + 3] << 24))
                  : (unsigned)(((unsigned)t[o] << 24) | (t[o + 1] << 16) | (t[o + 2] << 8) | t[o + 3]);
    };
    if (n < 8) return 1;
    const size_t ifd = u32(4);
    if (ifd + 2 > n) return 1;
    const unsigned cnt = u16(ifd);
    for (unsigned i = 0; i < cnt; ++i) {
        const size_t e = ifd + 2 + 12 * (size_t)i;
        if (e + 12 > n) break;
        if (u16(e) == 0x0112) return (int)u16(e + 8);
    }
    return 1;
}

#define K10_BAD(rc, ...)                                                   \
    do {                                                                   \
        if (err) snprintf(err, errlen, __VA_ARGS__);                       \
        return (rc);                                                       \
    } while (0)

int parse_header(const uint8_t* d, size_t len, Header& H, char* err, size_t errlen) {
    if (!d || len < 4 || d[0] != 0xFF || d[1] != 0xD8) K10_BAD(QC_ERR_INVALID, "jpeg: no SOI marker");
    size_t p = 2;
    while (p + 4 <= len) {
        if (d[p] != 0xFF) { ++p; continue; }
        int m = d[p + 1];
        if (m == 0xFF) { ++p; continue; }                       // fill byte
        p += 2;
        if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;    // standalone markers
        if (m == 0xD9) break;
        if (p + 2 > len) break;
        const int seg = rd16(d + p);
        if (seg < 2 || p + (size_t)seg > len) K10_BAD(QC_ERR_INVALID, "jpeg: truncated segment 0x%02X", m);
        const uint8_t* s = d + p + 2;
        const int n = seg - 2;
        if (m == 0xDB) {                                         // DQT
            int o = 0;
            while (o < n) {
                const int pq = s[o] >> 4, tq = s[o] & 15;
                ++o;
                if (tq > 3 || o + (pq ? 128 : 64) > n) K10_BAD(QC_ERR_INVALID, "jpeg: bad DQT");
                for (int i = 0; i < 64; ++i) {
                    H.qt[tq][k10_natural[i]] = (uint16_t)(pq ? rd16(s + o + 2 * i) : s[o + i]);
                }
                o += pq ? 128 : 64;
                H.qt_ok[tq] = true;
            }
        } else if (m == 0xC4) {                                  // DHT
            int o = 0;
            while (o < n) {
                if (o + 17 > n) K10_BAD(QC_ERR_INVALID, "jpeg: bad DHT");
                const int tc = s[o] >> 4, th = s[o] & 15;
                int total = 0;
                for (int i = 0; i < 16; ++i) total += s[o + 1 + i];
                if (tc > 1 || th > 3 || total > 256 || o + 17 + total > n) K10_BAD(QC_ERR_INVALID, "jpeg: bad DHT");
                if (!build_huff(tc ? H.ac[th] : H.dc[th], s + o + 1, s + o + 17, total))
                    K10_BAD(QC_ERR_INVALID, "jpeg: bad Huffman code lengths");
                o += 17 + total;
            }
        } else if (m == 0xC0 || m == 0xC1) {                     // SOF0 / SOF1: sequential Huffman
            if (n < 6) K10_BAD(QC_ERR_INVALID, "jpeg: bad SOF");
            if (s[0] != 8) K10_BAD(QC_ERR_UNSUPPORTED, "jpeg: %d-bit samples are not supported", s[0]);
            H.h = rd16(s + 1); H.w = rd16(s + 3); H.nc = s[5];
            if (H.nc != 1 && H.nc != 3) K10_BAD(QC_ERR_UNSUPPORTED, "jpeg: %d components are not supported", H.nc);
            if (H.w <= 0 || H.h <= 0 || n < 6 + 3 * H.nc) K10_BAD(QC_ERR_INVALID, "jpeg: bad SOF");
            for (int c = 0; c < H.nc; ++c) {
                H.cid[c] = s[6 + 3 * c];
                H.hs[c] = s[7 + 3 * c] >> 4; H.vs[c] = s[7 + 3 * c] & 15;
                H.tq[c] = s[8 + 3 * c];
                if (H.tq[c] > 3 || H.hs[c] < 1 || H.vs[c] < 1) K10_BAD(QC_ERR_INVALID, "jpeg: bad SOF component");
            }
            H.have_sof = true;
        } else if (m == 0xC2 || (m >= 0xC3 && m <= 0xCF && m != 0xC8 && m != 0xCC)) {
            K10_BAD(QC_ERR_UNSUPPORTED, "jpeg: SOF%d (progressive / lossless / arithmetic) is not supported", m - 0xC0);
        } else if (m == 0xDD) {
            if (n < 2) K10_BAD(QC_ERR_INVALID, "jpeg: bad DRI");
            H.restart = rd16(s);
        } else if (m == 0xE0) {
            if (n >= 5 && memcmp(s, "JFIF\0", 5) == 0) H.jfif = true;
        } else if (m == 0xE1) {
            const int o = exif_orientation(s, (size_t)n);
            if (o >= 1 && o <= 8) H.orientation = o;
        } else if (m == 0xEE) {
            if (n >= 12 && memcmp(s, "Adobe", 5) == 0) { H.adobe = true; H.adobe_transform = s[11]; }
        } else if (m == 0xDA) {                                  // SOS
            if (!H.have_sof) K10_BAD(QC_ERR_INVALID, "jpeg: SOS before SOF");
            const int ns = s[0];
            if (ns != H.nc || n < 1 + 2 * ns + 3)
                K10_BAD(QC_ERR_UNSUPPORTED, "jpeg: multi-scan (non-interleaved) streams are not supported");
            for (int i = 0; i < ns; ++i) {
                int c = -1;
                for (int j = 0; j < H.nc; ++j) if (H.cid[j] == s[1 + 2 * i]) c = j;
                if (c != i) K10_BAD(QC_ERR_UNSUPPORTED, "jpeg: scan component order is not supported");
                H.td[c] = s[2 + 2 * i] >> 4; H.ta[c] = s[2 + 2 * i] & 15;
                if (H.td[c] > 3 || H.ta[c] > 3 || !H.dc[H.td[c]].present || !H.ac[H.ta[c]].present)
                    K10_BAD(QC_ERR_INVALID, "jpeg: scan refers to a missing Huffman table");
                if (!H.qt_ok[H.tq[c]]) K10_BAD(QC_ERR_INVALID, "jpeg: missing quantisation table");
            }
            H.scan = p + (size_t)seg;
            H.have_sos = true;
            break;
        }
        p += (size_t)seg;
    }
    if (!H.have_sos) K10_BAD(QC_ERR_INVALID, "jpeg: no scan found");
    if (H.orientation != 1)
        K10_BAD(QC_ERR_UNSUPPORTED, "jpeg: EXIF orientation %d (cv2.imdecode would rotate) is not supported", H.orientation);
    if (H.nc == 3) {
        // jdapimin.c default_decompress_parms: what libjpeg would treat as YCbCr
        bool ycc = true;
        if (!H.jfif) {
            if (H.adobe) ycc = H.adobe_transform == 1;
            else if (H.cid[0] == 'R' && H.cid[1] == 'G' && H.cid[2] == 'B') ycc = false;
        }
        if (!ycc) K10_BAD(QC_ERR_UNSUPPORTED, "jpeg: RGB-coded streams are not supported");
        if (H.hs[1] != 1 || H.vs[1] != 1 || H.hs[2] != 1 || H.vs[2] != 1 || H.vs[0] > 2 || H.hs[0] > 2 ||
            (H.hs[0] == 1 && H.vs[0] == 2))
            K10_BAD(QC_ERR_UNSUPPORTED, "jpeg: %dx%d luma sampling is not supported (4:4:4, 4:2:2, 4:2:0 only)", H.hs[0], H.vs[0]);
    }
    return QC_OK;
}

struct Geometry {
    int mode;                          // qc_jpeg_subsampling
    int mcux, mcuy;                    // MCUs per row / column
    int bw[3], bh[3];                  // blocks per component plane (MCU padded)
    long long blocks;                  // total 8x8 blocks
};

Geometry geometry(const Header& H) {
    Geometry g{};
    if (H.nc == 1) {                   // a single-component scan is never interleaved: blocks cover ceil(w/8) x ceil(h/8)
        g.mode = QC_JPEG_GRAY;
        g.mcux = (H.w + 7) / 8; g.mcuy = (H.h + 7) / 8;
        g.bw[0] = g.mcux; g.bh[0] = g.mcuy;
        g.blocks = (long long)g.bw[0] * g.bh[0];
        return g;
    }
    const int hm = H.hs[0], vm = H.vs[0];
    g.mode = hm == 1 ? QC_JPEG_444 : (vm == 1 ? QC_JPEG_422 : QC_JPEG_420);
    g.mcux = (H.w + 8 * hm - 1) / (8 * hm); g.mcuy = (H.h + 8 * vm - 1) / (8 * vm);
    for (int c = 0; c < 3; ++c) {
        g.bw[c] = g.mcux * H.hs[c]; g.bh[c] = g.mcuy * H.vs[c];
        g.blocks += (long long)g.bw[c] * g.bh[c];
    }
    return g;
}

// ---- entropy stage (host) ------------------------------------------------------------------------------------------
#define K10_INLINE __attribute__((always_inline)) inline
struct BitReader {
    const uint8_t* p; const uint8_t* end;
    uint64_t acc = 0; int cnt = 0; bool marker = false;
    // slow refill: byte by byte, un-stuffing 0xFF00 and stopping at a marker (then zeros are fed, like libjpeg)
    __attribute__((noinline)) void fill_slow() {
        while (cnt <= 56) {
            unsigned b = 0;
            if (!marker && p < end) {
                b = *p++;
                if (b == 0xFF) {
                    const unsigned b2 = p < end ? *p : 0xD9u;
                    if (b2 == 0) ++p;
                    else { marker = true; --p; b = 0; }       // leave p on the 0xFF of the marker
                }
            }
            acc = (acc << 8) | b;
            cnt += 8;
        }
    }
    // at least 33 valid bits afterwards: 16 for a code and 15 for the value that follows it
    K10_INLINE void fill() {
        if (cnt > 32) return;
        if (!marker && p + 4 <= end) {                            // four bytes at once when none of them is 0xFF
            unsigned v;
            memcpy(&v, p, 4);
            if (!has_ff(v)) {
                acc = (acc << 32) | __builtin_bswap32(v);
                cnt += 32;
                p += 4;
                return;
            }
        }
        fill_slow();
    }
    static K10_INLINE bool has_ff(unsigned v) {
        return (v & 0xffu) == 0xffu || ((v >> 8) & 0xffu) == 0xffu || ((v >> 16) & 0xffu) == 0xffu || (v >> 24) == 0xffu;
    }
    K10_INLINE unsigned peek(int n) const { return (unsigned)((acc >> (cnt - n)) & ((1u << n) - 1u)); }
};

K10_INLINE int huff_decode(BitReader& br, const HuffTable& t) {
    br.fill();
    const unsigned e = t.lut[br.peek(K10_LOOK)];
    if (__builtin_expect(e != 0, 1)) { br.cnt -= (int)(e >> 8); return (int)(e & 255u); }
    int l = K10_LOOK + 1;
    int code = (int)br.peek(l);
    while (l <= 16 && code > t.maxcode[l]) { ++l; code = (int)br.peek(l); }
    if (l > 16) return -1;
    br.cnt -= l;
    return t.vals[(t.valoff[l] + code) & 255];
}

K10_INLINE int receive_extend(BitReader& br, int s) {
    const int v = (int)br.peek(s);
    br.cnt -= s;
    return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
}

// one image: Huffman-decode the single interleaved scan into coefs ([comp][block row][block col][64] int16, natural
// order, MCU padded).  Returns 0 or a negative status with a message in err.
int entropy_decode(const uint8_t* d, size_t len, const Header& H, const Geometry& g, int16_t* coefs, char* err, size_t errlen) {
    long long off[3] = {0, 0, 0};
    for (int c = 1; c < H.nc; ++c) off[c] = off[c - 1] + (long long)g.bw[c - 1] * g.bh[c - 1] * 64;
    memset(coefs, 0, (size_t)g.blocks * 64 * sizeof(int16_t));
    BitReader br{d + H.scan, d + len};
    int pred[3] = {0, 0, 0};
    int togo = H.restart;
    int next_rst = 0;
    const int bh_n[3] = {H.nc == 1 ? 1 : H.hs[0], 1, 1}, bv_n[3] = {H.nc == 1 ? 1 : H.vs[0], 1, 1};
    for (int my = 0; my < g.mcuy; ++my) {
        for (int mx = 0; mx < g.mcux; ++mx) {
            if (H.restart && togo == 0) {
                // discard the partial byte, expect RSTn
                br.acc = 0; br.cnt = 0;
                if (!br.marker) {
                    while (br.p + 1 < br.end && !(br.p[0] == 0xFF && br.p[1] != 0 && br.p[1] != 0xFF)) ++br.p;
                }
                if (br.p + 1 >= br.end || br.p[1] != 0xD0 + next_rst) K10_BAD(QC_ERR_INVALID, "jpeg: missing restart marker");
                br.p += 2; br.marker = false;
                next_rst = (next_rst + 1) & 7;
                pred[0] = pred[1] = pred[2] = 0;
                togo = H.restart;
            }
            for (int c = 0; c < H.nc; ++c) {
                const HuffTable& dct = H.dc[H.td[c]];
                const HuffTable& act = H.ac[H.ta[c]];
                for (int v = 0; v < bv_n[c]; ++v) {
                    for (int hx = 0; hx < bh_n[c]; ++hx) {
                        const int by = my * bv_n[c] + v, bx = mx * bh_n[c] + hx;
                        int16_t* blk = coefs + off[c] + ((long long)by * g.bw[c] + bx) * 64;
                        int s = huff_decode(br, dct);
                        if (s < 0 || s > 15) K10_BAD(QC_ERR_INVALID, "jpeg: corrupt DC code");
                        if (s) pred[c] += receive_extend(br, s);
                        blk[0] = (int16_t)pred[c];
                        for (int k = 1; k < 64;) {
                            br.fill();
                            const int f = act.fast[br.peek(K10_LOOK)];
                            if (f) {                         // short code + short value: one lookup
                                k += (f >> 4) & 15;
                                br.cnt -= f & 15;
                                blk[k10_natural[k]] = (int16_t)(f >> 8);
                                ++k;
                                continue;
                            }
                            const int rs = huff_decode(br, act);
                            if (rs < 0) K10_BAD(QC_ERR_INVALID, "jpeg: corrupt AC code");
                            const int r = rs >> 4;
                            s = rs & 15;
                            if (s == 0) {
                                if (r != 15) break;
                                k += 16;
                                continue;
                            }
                            k += r;                    // fill() left >= 33 bits: 16 for the code, 15 for the value
                            blk[k10_natural[k]] = (int16_t)receive_extend(br, s);
                            ++k;
                        }
                    }
                }
            }
            if (H.restart) --togo;
        }
    }
    return QC_OK;
}

// ---- device side ---------------------------------------------------------------------------------------------------
struct K10Img {
    int w, h, nc, mode;
    int bw[3], bh[3];
    long long coef_off[3];            // int16 units from the coefficient base
    long long plane_off[3];           // bytes from the plane base; plane pitch = bw*8
    int dw, dh;                       // chroma plane size in samples (libjpeg's downsampled_width / height)
    int dst_stride;
    long long dst_off;
};

constexpr int K10_IDCT_THREADS = 128;
constexpr int K10_COLOR_THREADS = 256;

#define K10_DESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))

// jidctint.c jpeg_idct_islow on one 8-vector (CONST_BITS 13); SHIFT = 11 in pass 1, 18 in pass 2
template <int SHIFT>
__device__ __forceinline__ void k10_idct8(const int (&in)[8], int (&out)[8]) {
    int z2 = in[2], z3 = in[6];
    int z1 = (z2 + z3) * 4433;
    int tmp2 = z1 + z3 * (-15137);
    int tmp3 = z1 + z2 * 6270;
    int tmp0 = (in[0] + in[4]) << 13;
    int tmp1 = (in[0] - in[4]) << 13;
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = in[7]; tmp1 = in[5]; tmp2 = in[3]; tmp3 = in[1];
    z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
    int z4 = tmp1 + tmp3;
    const int z5 = (z3 + z4) * 9633;
    tmp0 *= 2446; tmp1 *= 16819; tmp2 *= 25172; tmp3 *= 12299;
    z1 *= -7373; z2 *= -20995; z3 *= -16069; z4 *= -3196;
    z3 += z5; z4 += z5;
    tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
    out[0] = K10_DESCALE(tmp10 + tmp3, SHIFT); out[7] = K10_DESCALE(tmp10 - tmp3, SHIFT);
    out[1] = K10_DESCALE(tmp11 + tmp2, SHIFT); out[6] = K10_DESCALE(tmp11 - tmp2, SHIFT);
    out[2] = K10_DESCALE(tmp12 + tmp1, SHIFT); out[5] = K10_DESCALE(tmp12 - tmp1, SHIFT);
    out[3] = K10_DESCALE(tmp13 + tmp0, SHIFT); out[4] = K10_DESCALE(tmp13 - tmp0, SHIFT);
}

// one thread per 8x8 block: dequantise, columns then rows, +128, clamp, eight 8-byte row stores
__global__ void __launch_bounds__(K10_IDCT_THREADS)
k10_idct(const K10Img* __restrict__ imgs, const uint16_t* __restrict__ quant, const int16_t* __restrict__ coefs,
         uint8_t* __restrict__ planes) {
    __shared__ int s_q[64];
    const int img = blockIdx.y / 3, c = blockIdx.y - img * 3;
    const K10Img& I = imgs[img];
    if (c >= I.nc) return;
    const int nb = I.bw[c] * I.bh[c];
    if (blockIdx.x * K10_IDCT_THREADS >= nb) return;
    if (threadIdx.x < 64) s_q[threadIdx.x] = quant[((size_t)img * 3 + c) * 64 + threadIdx.x];
    __syncthreads();
    const int b = blockIdx.x * K10_IDCT_THREADS + threadIdx.x;
    if (b >= nb) return;
    const uint4* src = reinterpret_cast<const uint4*>(coefs + I.coef_off[c] + (long long)b * 64);
    int ws[8][8];                                               // [row][col]
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const uint4 v = __ldg(src + r);
        const unsigned u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            ws[r][2 * j] = (int)(short)(u[j] & 0xffffu) * s_q[r * 8 + 2 * j];
            ws[r][2 * j + 1] = (int)(short)(u[j] >> 16) * s_q[r * 8 + 2 * j + 1];
        }
    }
#pragma unroll
    for (int col = 0; col < 8; ++col) {
        int in[8], out[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) in[r] = ws[r][col];
        k10_idct8<11>(in, out);
#pragma unroll
        for (int r = 0; r < 8; ++r) ws[r][col] = out[r];
    }
    const int by = b / I.bw[c], bx = b - by * I.bw[c];
    const int pitch = I.bw[c] * 8;
    uint8_t* dst = planes + I.plane_off[c] + (long long)by * 8 * pitch + bx * 8;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        int out[8];
        k10_idct8<18>(ws[r], out);
        unsigned lo = 0, hi = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            lo |= (unsigned)min(max(out[j] + 128, 0), 255) << (8 * j);
            hi |= (unsigned)min(max(out[4 + j] + 128, 0), 255) << (8 * j);
        }
        *reinterpret_cast<uint2*>(dst + (long long)r * pitch) = make_uint2(lo, hi);
    }
}

// jdsample.c: the upsampled chroma values of 8 consecutive pixels x0..x0+7 (x0 a multiple of 8) of row oy.  The four
// chroma samples under them come in as one aligned word per contributing row, the two neighbours as single bytes.
__device__ __forceinline__ void k10_chroma8(const uint8_t* __restrict__ pl, int pitch, int dw, int dh, int mode, int oy,
                                            int x0, int (&out)[8]) {
    if (mode == QC_JPEG_444) {
        const uint2 v = *reinterpret_cast<const uint2*>(pl + (long long)oy * pitch + x0);
#pragma unroll
        for (int j = 0; j < 4; ++j) { out[j] = (v.x >> (8 * j)) & 255u; out[4 + j] = (v.y >> (8 * j)) & 255u; }
        return;
    }
    const int c0 = x0 >> 1;                                   // a multiple of 4: aligned in the MCU-padded plane
    const int r0 = mode == QC_JPEG_422 ? oy : oy >> 1;
    const uint8_t* a = pl + (long long)r0 * pitch;
    const unsigned wa = *reinterpret_cast<const unsigned*>(a + c0);
    if (dw <= 2) {                                            // h2v1_upsample / h2v2_upsample: plain replication
#pragma unroll
        for (int j = 0; j < 8; ++j) out[j] = (wa >> (8 * (j >> 1))) & 255u;
        return;
    }
    const int cl = max(c0 - 1, 0), cr = min(c0 + 4, dw - 1);  // neighbours (clamped: the edge formulas ignore them)
    int v[6];
    if (mode == QC_JPEG_422) {                                // h2v1_fancy_upsample on the samples themselves
        v[0] = a[cl]; v[5] = a[cr];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[1 + j] = (wa >> (8 * j)) & 255u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + j, cur = v[1 + j];
            out[2 * j] = c == 0 ? cur : (3 * cur + v[j] + 1) >> 2;
            out[2 * j + 1] = c >= dw - 1 ? cur : (3 * cur + v[2 + j] + 2) >> 2;
        }
        return;
    }
    // h2v2_fancy_upsample: column sums 3*near + far of the two contributing rows (context rows replicate the first /
    // last real row), then the same 3:1 blend horizontally with roundings 8 and 7
    const int r1 = (oy & 1) ? min(r0 + 1, dh - 1) : max(r0 - 1, 0);
    const uint8_t* b = pl + (long long)r1 * pitch;
    const unsigned wb = *reinterpret_cast<const unsigned*>(b + c0);
    v[0] = 3 * a[cl] + b[cl]; v[5] = 3 * a[cr] + b[cr];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[1 + j] = 3 * (int)((wa >> (8 * j)) & 255u) + (int)((wb >> (8 * j)) & 255u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = c0 + j, cur = v[1 + j];
        out[2 * j] = c == 0 ? (cur * 4 + 8) >> 4 : (cur * 3 + v[j] + 8) >> 4;
        out[2 * j + 1] = c >= dw - 1 ? (cur * 4 + 7) >> 4 : (cur * 3 + v[2 + j] + 7) >> 4;
    }
}

// upsample + jdcolor.c ycc_rgb_convert (+ gray_rgb_convert), 8 pixels = 6 words per thread
__global__ void __launch_bounds__(K10_COLOR_THREADS)
k10_color(const K10Img* __restrict__ imgs, const uint8_t* __restrict__ planes, uint8_t* __restrict__ dst, int rgb) {
    const K10Img& I = imgs[blockIdx.y];
    const int wq = (I.w + 7) >> 3;
    const long long items = (long long)I.h * wq;
    const int p0 = I.bw[0] * 8, p1 = I.nc == 3 ? I.bw[1] * 8 : 0;
    const uint8_t* Y = planes + I.plane_off[0];
    const uint8_t* Cb = planes + I.plane_off[1];
    const uint8_t* Cr = planes + I.plane_off[2];
    for (long long it = (long long)blockIdx.x * K10_COLOR_THREADS + threadIdx.x; it < items;
         it += (long long)gridDim.x * K10_COLOR_THREADS) {
        const int y = (int)(it / wq), x0 = (int)(it - (long long)y * wq) * 8;
        const uint2 yw = *reinterpret_cast<const uint2*>(Y + (long long)y * p0 + x0);       // plane pitch is a multiple of 8
        int cb[8], cr[8];
        if (I.nc == 3) {
            k10_chroma8(Cb, p1, I.dw, I.dh, I.mode, y, x0, cb);
            k10_chroma8(Cr, p1, I.dw, I.dh, I.mode, y, x0, cr);
        }
        unsigned px[6] = {0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int yy = (int)(((j < 4 ? yw.x : yw.y) >> (8 * (j & 3))) & 255u);
            int r = yy, g = yy, b = yy;
            if (I.nc == 3) {
                const int u = cb[j] - 128, v = cr[j] - 128;
                r = min(max(yy + ((91881 * v + 32768) >> 16), 0), 255);
                b = min(max(yy + ((116130 * u + 32768) >> 16), 0), 255);
                g = min(max(yy + ((-22554 * u + 32768 - 46802 * v) >> 16), 0), 255);
            }
            const unsigned c3[3] = {(unsigned)(rgb ? r : b), (unsigned)g, (unsigned)(rgb ? b : r)};
#pragma unroll
            for (int c = 0; c < 3; ++c) px[(3 * j + c) >> 2] |= c3[c] << (8 * ((3 * j + c) & 3));
        }
        uint8_t* o = dst + I.dst_off + (long long)y * I.dst_stride + 3ll * x0;
        if (x0 + 8 <= I.w && ((reinterpret_cast<uintptr_t>(o) & 3) == 0)) {
            unsigned* o4 = reinterpret_cast<unsigned*>(o);
#pragma unroll
            for (int k = 0; k < 6; ++k) o4[k] = px[k];
        } else {
            const int nv = min(8, I.w - x0) * 3;
            for (int j = 0; j < nv; ++j) o[j] = (uint8_t)(px[j >> 2] >> (8 * (j & 3)));
        }
    }
}

// staging / workspace layout: [K10Img n][quant u16 n*3*64][coefs int16 ...] ; the device workspace appends the planes
struct Layout {
    size_t off_img, off_q, off_coef, staging, planes_off, total;
};
inline size_t al(size_t v) { return (v + 255) & ~(size_t)255; }

Layout layout_of(const qc_jpeg_info* infos, int n) {
    Layout L{};
    size_t o = 0;
    L.off_img = o; o = al(o + (size_t)n * sizeof(K10Img));
    L.off_q = o; o = al(o + (size_t)n * 3 * 64 * sizeof(uint16_t));
    L.off_coef = o;
    long long blocks = 0;
    for (int i = 0; i < n; ++i) blocks += infos[i].blocks;
    o = al(o + (size_t)blocks * 64 * sizeof(int16_t));
    L.staging = o;
    L.planes_off = o;
    o = al(o + (size_t)blocks * 64);
    L.total = o;
    return L;
}

}  // namespace

extern "C" int qc_jpeg_info_host(qc_ctx* ctx, const uint8_t* data, size_t len, qc_jpeg_info* info) {
    if (!info) return QC_ERR_INVALID;
    memset(info, 0, sizeof(*info));
    Header H;
    char msg[256] = {0};
    const int rc = parse_header(data, len, H, msg, sizeof(msg));
    if (rc != QC_OK) {
        if (ctx) snprintf(ctx->err, sizeof(ctx->err), "%s", msg);
        return rc;
    }
    const Geometry g = geometry(H);
    info->width = H.w; info->height = H.h; info->components = H.nc; info->subsampling = g.mode;
    info->restart_interval = H.restart; info->blocks = g.blocks;
    return QC_OK;
}

extern "C" int qc_jpeg_entropy_host(qc_ctx* ctx, const uint8_t* data, size_t len, int16_t* coefs, uint16_t* quant) {
    if (!coefs || !quant) return QC_ERR_INVALID;
    Header H;
    char msg[256] = {0};
    int rc = parse_header(data, len, H, msg, sizeof(msg));
    if (rc == QC_OK) {
        const Geometry g = geometry(H);
        rc = entropy_decode(data, len, H, g, coefs, msg, sizeof(msg));
        for (int c = 0; c < H.nc && rc == QC_OK; ++c)
            for (int k = 0; k < 64; ++k) quant[c * 64 + k] = H.qt[H.tq[c]][k];
    }
    if (rc != QC_OK && ctx) snprintf(ctx->err, sizeof(ctx->err), "%s", msg);
    return rc;
}

extern "C" size_t qc_jpeg_staging_bytes(const qc_jpeg_info* infos, int n) {
    if (!infos || n <= 0) return 0;
    return layout_of(infos, n).staging;
}

extern "C" size_t qc_jpeg_workspace_bytes(const qc_jpeg_info* infos, int n) {
    if (!infos || n <= 0) return 0;
    return layout_of(infos, n).total;
}

extern "C" int qc_jpeg_decode(qc_ctx* ctx, const uint8_t* const* jpegs, const size_t* lens, int n,
                              const qc_jpeg_info* infos, int host_threads, void* host_staging, size_t staging_bytes,
                              void* workspace, size_t workspace_bytes, uint8_t* dst, const int64_t* dst_offsets,
                              const int32_t* dst_row_strides, int rgb, void* stream) {
    qc_device_guard dev_guard(ctx);
    if (!ctx) return QC_ERR_INVALID;
    QC_ENTER(ctx, stream);
    if (n == 0) return QC_OK;
    if (!jpegs || !lens || !infos || n < 0 || !host_staging || !workspace || !dst || !dst_offsets || !dst_row_strides)
        QC_FAIL(ctx, QC_ERR_INVALID, "qc_jpeg_decode: bad argument");
    const Layout L = layout_of(infos, n);
    if (staging_bytes < L.staging || workspace_bytes < L.total)
        QC_FAIL(ctx, QC_ERR_WORKSPACE, "qc_jpeg_decode: staging %zu / workspace %zu bytes given, %zu / %zu needed",
                staging_bytes, workspace_bytes, L.staging, L.total);
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* hs = static_cast<uint8_t*>(host_staging);
    K10Img* himg = reinterpret_cast<K10Img*>(hs + L.off_img);
    uint16_t* hq = reinterpret_cast<uint16_t*>(hs + L.off_q);
    int16_t* hcoef = reinterpret_cast<int16_t*>(hs + L.off_coef);

    // ---- host: headers, geometry, offsets ----------------------------------------------------------
    std::vector<Header> hdr((size_t)n);
    std::vector<Geometry> geo((size_t)n);
    std::vector<long long> coef_base((size_t)n);
    long long cb = 0, pb = 0;
    int max_blocks = 1;
    long long max_items = 1;
    for (int i = 0; i < n; ++i) {
        char msg[256] = {0};
        const int rc = parse_header(jpegs[i], lens[i], hdr[i], msg, sizeof(msg));
        if (rc != QC_OK) QC_FAIL(ctx, rc, "qc_jpeg_decode: image %d: %s", i, msg);
        geo[i] = geometry(hdr[i]);
        const Header& H = hdr[i];
        const Geometry& g = geo[i];
        if (g.blocks != infos[i].blocks || H.w != infos[i].width || H.h != infos[i].height)
            QC_FAIL(ctx, QC_ERR_INVALID, "qc_jpeg_decode: image %d does not match its qc_jpeg_info", i);
        if (dst_row_strides[i] < 3 * H.w) QC_FAIL(ctx, QC_ERR_INVALID, "qc_jpeg_decode: image %d: row stride too small", i);
        K10Img& I = himg[i];
        memset(&I, 0, sizeof(I));
        I.w = H.w; I.h = H.h; I.nc = H.nc; I.mode = g.mode;
        coef_base[i] = cb;
        for (int c = 0; c < H.nc; ++c) {
            I.bw[c] = g.bw[c]; I.bh[c] = g.bh[c];
            I.coef_off[c] = cb; I.plane_off[c] = pb;
            cb += (long long)g.bw[c] * g.bh[c] * 64;
            pb += (long long)g.bw[c] * g.bh[c] * 64;
            max_blocks = std::max(max_blocks, g.bw[c] * g.bh[c]);
            for (int k = 0; k < 64; ++k) hq[((size_t)i * 3 + c) * 64 + k] = H.qt[H.tq[c]][k];
        }
        const int hm = H.nc == 3 ? H.hs[0] : 1, vm = H.nc == 3 ? H.vs[0] : 1;
        I.dw = (H.w + hm - 1) / hm; I.dh = (H.h + vm - 1) / vm;
        I.dst_stride = dst_row_strides[i]; I.dst_off = dst_offsets[i];
        max_items = std::max(max_items, (long long)H.h * ((H.w + 7) / 8));
    }
    // ---- host: entropy decode, one image per task; every finished image's coefficient blocks are sent to the device at
    // once (from the thread that decoded it), so the copies overlap the Huffman work of the other images ----------------
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    QC_ENTER(ctx, stream);
    QC_CUDA(ctx, cudaMemcpyAsync(ws, hs, L.off_coef, cudaMemcpyHostToDevice, st));          // image table + quantisers
    std::atomic<int> next{0}, failed{-1};
    std::vector<char> msgs((size_t)n * 128, 0);
    std::vector<int> rcs((size_t)n, 0);
    const int device = ctx->device;
    auto work = [&](bool set_device) {
        if (set_device) cudaSetDevice(device);
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= n) break;
            rcs[i] = entropy_decode(jpegs[i], lens[i], hdr[i], geo[i], hcoef + coef_base[i], &msgs[(size_t)i * 128], 128);
            if (rcs[i] == QC_OK) {
                const size_t off = L.off_coef + (size_t)coef_base[i] * sizeof(int16_t);
                const cudaError_t e = cudaMemcpyAsync(ws + off, hs + off, (size_t)geo[i].blocks * 64 * sizeof(int16_t),
                                                      cudaMemcpyHostToDevice, st);
                if (e != cudaSuccess) {
                    rcs[i] = QC_ERR_CUDA;
                    snprintf(&msgs[(size_t)i * 128], 128, "cudaMemcpyAsync: %s", cudaGetErrorString(e));
                }
            }
            if (rcs[i] != QC_OK) { int e = -1; failed.compare_exchange_strong(e, i); }
        }
    };
    const int nt = std::max(1, std::min(host_threads > 0 ? host_threads : 1, n));
    if (nt == 1) work(false);
    else {
        std::vector<std::thread> th;
        th.reserve((size_t)nt - 1);
        for (int t = 1; t < nt; ++t) th.emplace_back(work, true);
        work(false);
        for (auto& t : th) t.join();
    }
    if (failed.load() >= 0) {
        const int i = failed.load();
        QC_FAIL(ctx, rcs[i], "qc_jpeg_decode: image %d: %s", i, &msgs[(size_t)i * 128]);
    }
    // ---- device: IDCT, colour (the stream already holds every copy) ---------------------------------
    QC_ENTER(ctx, stream);                 // profiling boundary: the host stage above is not kernel time
    const K10Img* dimg = reinterpret_cast<const K10Img*>(ws + L.off_img);
    {
        dim3 grid((unsigned)((max_blocks + K10_IDCT_THREADS - 1) / K10_IDCT_THREADS), (unsigned)(3 * n));
        k10_idct<<<grid, K10_IDCT_THREADS, 0, st>>>(dimg, reinterpret_cast<const uint16_t*>(ws + L.off_q),
                                                   reinterpret_cast<const int16_t*>(ws + L.off_coef), ws + L.planes_off);
        QC_LAUNCHED_AS(ctx, "k10_idct", st);
    }
    {
        const long long want = (max_items + K10_COLOR_THREADS - 1) / K10_COLOR_THREADS;
        dim3 grid((unsigned)std::min<long long>(want, 4096), (unsigned)n);
        k10_color<<<grid, K10_COLOR_THREADS, 0, st>>>(dimg, ws + L.planes_off, dst, rgb);
        QC_LAUNCHED_AS(ctx, "k10_color", st);
    }
    return QC_OK;
}
