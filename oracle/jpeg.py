"""ORACLE (test infrastructure, never imported by the product): CPU restatement of cv2.imdecode(buf, IMREAD_COLOR) for
baseline JPEG, the decode the reference runs on every upload (inference/detector.py:214,229; models/yolo_model.py:74-76;
app.py:190-194).

The algorithm lives in a third-party dependency that is not in /root/reference: OpenCV 4.13's bundled libjpeg-turbo
3.1.2 (cv2.getBuildInformation(): "JPEG: build-libjpeg-turbo (ver 3.1.2-70)").  Its published integer pipeline is
restated here in numpy:
    jdhuff.c      decode_mcu / jpeg_make_d_derived_tbl       -> huffman_blocks()   (pure Python, small images only)
    jidctint.c    jpeg_idct_islow (CONST_BITS 13, PASS1 2)   -> idct_islow()
    jdsample.c    h2v1/h2v2_fancy_upsample, h2v1/h2v2_upsample for planes <= 2 samples wide;
    jdmainct.c    context rows replicate the first / last real sample row        -> upsample_chroma()
    jdcolor.c     build_ycc_rgb_table / ycc_rgb_convert, gray_rgb_convert          -> ycc_to_bgr()
Pinned (tests/test_jpeg_cpu.py) against cv2.imdecode itself on streams written by cv2.imencode over qualities,
samplings 4:4:4 / 4:2:2 / 4:2:0 / gray, odd sizes, restart intervals and optimised Huffman tables: byte-identical.
"""
import numpy as np

NATURAL = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,
                    6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45,
                    38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63] + [63] * 16)

GRAY, S444, S422, S420 = 0, 1, 2, 3


class Unsupported(ValueError):
    pass


def parse(data: bytes) -> dict:
    """Marker segments up to the first SOS (ITU T.81 B.2)."""
    d = bytes(data)
    if d[:2] != b"\xff\xd8":
        raise ValueError("no SOI")
    H = {"qt": {}, "dc": {}, "ac": {}, "restart": 0, "jfif": False, "adobe": None}
    p = 2
    while p + 4 <= len(d):
        if d[p] != 0xFF:
            p += 1
            continue
        m = d[p + 1]
        if m == 0xFF:
            p += 1
            continue
        p += 2
        if m == 0x01 or 0xD0 <= m <= 0xD7:
            continue
        if m == 0xD9:
            break
        n = (d[p] << 8) | d[p + 1]
        s = d[p + 2:p + n]
        if m == 0xDB:
            o = 0
            while o < len(s):
                pq, tq = s[o] >> 4, s[o] & 15
                o += 1
                q = np.zeros(64, np.int64)
                for i in range(64):
                    q[NATURAL[i]] = ((s[o + 2 * i] << 8) | s[o + 2 * i + 1]) if pq else s[o + i]
                o += 128 if pq else 64
                H["qt"][tq] = q
        elif m == 0xC4:
            o = 0
            while o < len(s):
                tc, th = s[o] >> 4, s[o] & 15
                counts = list(s[o + 1:o + 17])
                tot = sum(counts)
                H["ac" if tc else "dc"][th] = (counts, list(s[o + 17:o + 17 + tot]))
                o += 17 + tot
        elif m in (0xC0, 0xC1):
            if s[0] != 8:
                raise Unsupported("precision")
            H["h"], H["w"], H["nc"] = (s[1] << 8) | s[2], (s[3] << 8) | s[4], s[5]
            H["comp"] = [{"id": s[6 + 3 * c], "hs": s[7 + 3 * c] >> 4, "vs": s[7 + 3 * c] & 15, "tq": s[8 + 3 * c]}
                         for c in range(H["nc"])]
        elif 0xC2 <= m <= 0xCF and m not in (0xC4, 0xC8, 0xCC):
            raise Unsupported("SOF%d" % (m - 0xC0))
        elif m == 0xDD:
            H["restart"] = (s[0] << 8) | s[1]
        elif m == 0xE0 and s[:5] == b"JFIF\0":
            H["jfif"] = True
        elif m == 0xEE and s[:5] == b"Adobe":
            H["adobe"] = s[11]
        elif m == 0xDA:
            ns = s[0]
            if ns != H["nc"]:
                raise Unsupported("multi-scan")
            for i in range(ns):
                assert s[1 + 2 * i] == H["comp"][i]["id"]
                H["comp"][i]["td"], H["comp"][i]["ta"] = s[2 + 2 * i] >> 4, s[2 + 2 * i] & 15
            H["scan"] = p + n
            break
        p += n
    if H["nc"] == 1:
        H["mode"] = GRAY
        H["mcux"], H["mcuy"] = -(-H["w"] // 8), -(-H["h"] // 8)
        H["blocks"] = [(H["mcuy"], H["mcux"])]
    else:
        hm, vm = H["comp"][0]["hs"], H["comp"][0]["vs"]
        if any((c["hs"], c["vs"]) != (1, 1) for c in H["comp"][1:]) or (hm, vm) not in ((1, 1), (2, 1), (2, 2)):
            raise Unsupported("sampling")
        H["mode"] = {(1, 1): S444, (2, 1): S422, (2, 2): S420}[(hm, vm)]
        H["mcux"], H["mcuy"] = -(-H["w"] // (8 * hm)), -(-H["h"] // (8 * vm))
        H["blocks"] = [(H["mcuy"] * c["vs"], H["mcux"] * c["hs"]) for c in H["comp"]]
    return H


def _derived(counts, vals):
    """jdhuff.c jpeg_make_d_derived_tbl: {(length, code): symbol} for canonical codes."""
    table, code, k = {}, 0, 0
    for l in range(1, 17):
        for _ in range(counts[l - 1]):
            table[(l, code)] = vals[k]
            code += 1
            k += 1
        code <<= 1
    return table


def huffman_blocks(data: bytes, H: dict):
    """jdhuff.c decode_mcu, bit by bit: per component an int16 array [block rows][block cols][64] in natural order."""
    d = bytes(data)
    # un-stuff the entropy-coded segment into a list of bits per restart interval
    p = H["scan"]
    segs, cur = [], bytearray()
    while p < len(d):
        b = d[p]
        if b == 0xFF:
            b2 = d[p + 1] if p + 1 < len(d) else 0xD9
            if b2 == 0:
                cur.append(0xFF)
                p += 2
                continue
            if 0xD0 <= b2 <= 0xD7:
                segs.append(bytes(cur))
                cur = bytearray()
                p += 2
                continue
            if b2 == 0xFF:
                p += 1
                continue
            break
        cur.append(b)
        p += 1
    segs.append(bytes(cur))
    dc = {k: _derived(*v) for k, v in H["dc"].items()}
    ac = {k: _derived(*v) for k, v in H["ac"].items()}
    out = [np.zeros((bh, bw, 64), np.int16) for bh, bw in H["blocks"]]
    comps = H["comp"]
    nb = [(1, 1)] * len(comps) if H["nc"] == 1 else [(c["vs"], c["hs"]) for c in comps]

    class Bits:
        def __init__(self, seg):
            self.bits = np.unpackbits(np.frombuffer(seg, np.uint8)).tolist() if seg else []
            self.i = 0

        def get(self, n):
            v = 0
            for _ in range(n):
                v = (v << 1) | (self.bits[self.i] if self.i < len(self.bits) else 0)
                self.i += 1
            return v

        def sym(self, table):
            code = 0
            for l in range(1, 17):
                code = (code << 1) | self.get(1)
                if (l, code) in table:
                    return table[(l, code)]
            raise ValueError("bad code")

    def extend(v, s):
        return v - (1 << s) + 1 if v < (1 << (s - 1)) else v

    seg_i, br, pred = 0, Bits(segs[0]), [0] * len(comps)
    togo = H["restart"]
    for my in range(H["mcuy"]):
        for mx in range(H["mcux"]):
            if H["restart"] and togo == 0:
                seg_i += 1
                br, pred, togo = Bits(segs[seg_i]), [0] * len(comps), H["restart"]
            for c, comp in enumerate(comps):
                for v in range(nb[c][0]):
                    for hx in range(nb[c][1]):
                        blk = out[c][my * nb[c][0] + v, mx * nb[c][1] + hx]
                        s = br.sym(dc[comp["td"]])
                        if s:
                            pred[c] += extend(br.get(s), s)
                        blk[0] = pred[c]
                        k = 1
                        while k < 64:
                            rs = br.sym(ac[comp["ta"]])
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r != 15:
                                    break
                                k += 16
                                continue
                            k += r
                            blk[NATURAL[k]] = extend(br.get(s), s)
                            k += 1
            if H["restart"]:
                togo -= 1
    return out


def _idct8(x, shift):
    """jidctint.c: one 1-D pass over the last axis' 8 entries given as a list of 8 arrays."""
    def descale(v):
        return (v + (1 << (shift - 1))) >> shift
    z2, z3 = x[2], x[6]
    z1 = (z2 + z3) * 4433
    tmp2 = z1 + z3 * -15137
    tmp3 = z1 + z2 * 6270
    tmp0 = (x[0] + x[4]) << 13
    tmp1 = (x[0] - x[4]) << 13
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = x[7], x[5], x[3], x[1]
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * 9633
    tmp0, tmp1, tmp2, tmp3 = tmp0 * 2446, tmp1 * 16819, tmp2 * 25172, tmp3 * 12299
    z1, z2, z3, z4 = z1 * -7373, z2 * -20995, z3 * -16069 + z5, z4 * -3196 + z5
    tmp0, tmp1, tmp2, tmp3 = tmp0 + z1 + z3, tmp1 + z2 + z4, tmp2 + z2 + z3, tmp3 + z1 + z4
    return [descale(tmp10 + tmp3), descale(tmp11 + tmp2), descale(tmp12 + tmp1), descale(tmp13 + tmp0),
            descale(tmp13 - tmp0), descale(tmp12 - tmp1), descale(tmp11 - tmp2), descale(tmp10 - tmp3)]


def idct_islow(coefs: np.ndarray, quant: np.ndarray) -> np.ndarray:
    """coefs [bh][bw][64] (natural order) -> uint8 plane [bh*8][bw*8]: dequantise, columns (descale 11), rows
    (descale 18), +128, clamp."""
    bh, bw, _ = coefs.shape
    x = coefs.astype(np.int64).reshape(bh, bw, 8, 8) * quant.astype(np.int64).reshape(8, 8)
    cols = _idct8([x[:, :, r, :] for r in range(8)], 11)            # over rows index r, per column
    ws = np.stack(cols, axis=2)                                     # [bh][bw][row][col]
    rows = _idct8([ws[:, :, :, c] for c in range(8)], 18)
    px = np.clip(np.stack(rows, axis=3) + 128, 0, 255).astype(np.uint8)
    return px.transpose(0, 2, 1, 3).reshape(bh * 8, bw * 8)


def upsample_chroma(pl: np.ndarray, mode: int, w: int, h: int) -> np.ndarray:
    """Chroma plane (MCU padded) -> [h][w] at luma resolution (jdsample.c)."""
    if mode == S444:
        return pl[:h, :w].astype(np.int64)
    dw = -(-w // 2)
    if mode == S422:
        a = pl[:h, :dw].astype(np.int64)
        if dw <= 2:
            return np.repeat(a, 2, 1)[:, :w]
        out = np.empty((h, 2 * dw), np.int64)
        prev = np.concatenate([a[:, :1], a[:, :-1]], 1)
        nxt = np.concatenate([a[:, 1:], a[:, -1:]], 1)
        out[:, 0::2] = (3 * a + prev + 1) >> 2
        out[:, 1::2] = (3 * a + nxt + 2) >> 2
        out[:, 0] = a[:, 0]
        out[:, -1] = a[:, -1]
        return out[:, :w]
    dh = -(-h // 2)
    a = pl[:dh, :dw].astype(np.int64)
    if dw <= 2:
        return np.repeat(np.repeat(a, 2, 0), 2, 1)[:h, :w]
    up = np.concatenate([a[:1], a[:-1]], 0)                         # context above: first row replicated
    dn = np.concatenate([a[1:], a[-1:]], 0)                         # context below: last real row replicated
    out = np.empty((2 * dh, 2 * dw), np.int64)
    for v, other in ((0, up), (1, dn)):
        cs = 3 * a + other                                          # column sums
        prev = np.concatenate([cs[:, :1], cs[:, :-1]], 1)
        nxt = np.concatenate([cs[:, 1:], cs[:, -1:]], 1)
        ev = (3 * cs + prev + 8) >> 4
        od = (3 * cs + nxt + 7) >> 4
        ev[:, 0] = (4 * cs[:, 0] + 8) >> 4
        od[:, -1] = (4 * cs[:, -1] + 7) >> 4
        out[v::2, 0::2] = ev
        out[v::2, 1::2] = od
    return out[:h, :w]


def ycc_to_bgr(y, cb, cr) -> np.ndarray:
    """jdcolor.c build_ycc_rgb_table + ycc_rgb_convert (SCALEBITS 16), BGR order like cv2."""
    y, cb, cr = y.astype(np.int64), cb.astype(np.int64) - 128, cr.astype(np.int64) - 128
    r = y + ((91881 * cr + 32768) >> 16)
    b = y + ((116130 * cb + 32768) >> 16)
    g = y + ((-22554 * cb + 32768 - 46802 * cr) >> 16)
    return np.clip(np.stack([b, g, r], -1), 0, 255).astype(np.uint8)


def reconstruct(H: dict, coefs, quants) -> np.ndarray:
    """coefficient blocks per component -> the BGR frame cv2.imdecode(IMREAD_COLOR) returns."""
    w, h = H["w"], H["h"]
    planes = [idct_islow(c, q) for c, q in zip(coefs, quants)]
    if H["nc"] == 1:
        g = planes[0][:h, :w]
        return np.stack([g, g, g], -1)
    cb = upsample_chroma(planes[1], H["mode"], w, h)
    cr = upsample_chroma(planes[2], H["mode"], w, h)
    return ycc_to_bgr(planes[0][:h, :w], cb, cr)


def decode(data: bytes) -> np.ndarray:
    H = parse(data)
    coefs = huffman_blocks(data, H)
    return reconstruct(H, coefs, [H["qt"][c["tq"]] for c in H["comp"]])
