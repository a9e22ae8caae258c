"""TEST INFRASTRUCTURE ONLY (never imported by the product): CPU restatement of what
``PIL.Image.open(path)`` computes for the files SPEED ships -- baseline, Huffman-coded, 8-bit, single-component JPEG
(reference call sites: RV/datasets/speed.py:116 and :212, ``Image.open(img_path).convert('RGB')``).

The algorithm lives in a third-party dependency, not in /root/reference: Pillow (``pillow`` unpinned in
RV/requirements.txt; 12.2.0 here) -> libjpeg-turbo (libjpeg API 6.2 here).  Restated from the published standard and
the library's documented default:
  * entropy decoding: ITU-T T.81 Annex F.2.2 (DECODE with MINCODE / MAXCODE / VALPTR, RECEIVE + EXTEND, DC
    differences, EOB / ZRL runs, zigzag order Figure A.6), restart intervals E.2.4
  * inverse DCT: libjpeg's default JDCT_ISLOW (jidctint.c, jpeg_idct_islow): Loeffler-Ligtenberg-Moschytz with 13-bit
    integer constants, column pass descaled by 2^(13-2), row pass by 2^(13+2+3), result range-limited around 128
Pinned: ``tests/test_oracle.py::test_jpeg_ref_matches_pil`` checks this restatement against PIL's decoder bit for bit
(several sizes, qualities, optimised tables, restart markers).  PIL itself is the oracle at full frame size.
"""
import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7,
                   14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39,
                   46, 53, 60, 61, 54, 47, 55, 62, 63])


class Unsupported(ValueError):
    pass


def parse(data):
    """marker segments -> dict(width, height, restart, q[64] (zigzag order), dc/ac (counts[17], vals), scan bytes)"""
    d = memoryview(data)
    if len(d) < 4 or d[0] != 0xFF or d[1] != 0xD8:
        raise ValueError("not a JPEG file")
    qt, dc, ac = {}, {}, {}
    out = {"restart": 0}
    p = 2
    tq = td = ta = None
    while p + 4 <= len(d):
        assert d[p] == 0xFF
        while d[p] == 0xFF:
            p += 1
        m = d[p]
        p += 1
        if m == 0xD8 or 0xD0 <= m <= 0xD7 or m == 0x01:
            continue
        if m == 0xD9:
            break
        ln = (d[p] << 8) | d[p + 1]
        s = bytes(d[p + 2:p + ln])
        if m == 0xDB:
            i = 0
            while i < len(s):
                pq, t = s[i] >> 4, s[i] & 15
                i += 1
                if pq:
                    qt[t] = [(s[i + 2 * k] << 8) | s[i + 2 * k + 1] for k in range(64)]
                    i += 128
                else:
                    qt[t] = list(s[i:i + 64])
                    i += 64
        elif m == 0xC4:
            i = 0
            while i < len(s):
                tc, th = s[i] >> 4, s[i] & 15
                counts = [0] + list(s[i + 1:i + 17])
                n = sum(counts)
                vals = list(s[i + 17:i + 17 + n])
                i += 17 + n
                (ac if tc else dc)[th] = (counts, vals)
        elif m in (0xC0, 0xC1):
            if s[0] != 8:
                raise Unsupported("only 8-bit samples")
            out["height"], out["width"] = (s[1] << 8) | s[2], (s[3] << 8) | s[4]
            if s[5] != 1:
                raise Unsupported(f"{s[5]} components")
            tq = s[8]
        elif m == 0xC2 or (0xC3 <= m <= 0xCF and m not in (0xC4, 0xC8, 0xCC)):
            raise Unsupported("progressive / lossless / arithmetic")
        elif m == 0xDD:
            out["restart"] = (s[0] << 8) | s[1]
        elif m == 0xDA:
            td, ta = s[2] >> 4, s[2] & 15
            q = p + ln
            begin = q
            while q + 1 < len(d):
                if d[q] != 0xFF:
                    q += 1
                    continue
                mm = d[q + 1]
                if mm == 0 or 0xD0 <= mm <= 0xD7:
                    q += 2
                    continue
                if mm == 0xFF:
                    q += 1
                    continue
                break
            out["scan"] = bytes(d[begin:q])
            break
        p += ln
    out["q"] = qt[tq]
    out["dc"], out["ac"] = dc[td], ac[ta]
    return out


def _tables(counts, vals):
    """T.81 Annex C / F.2.2.3: MINCODE, MAXCODE, VALPTR per code length"""
    mincode, maxcode, valptr = [0] * 17, [-1] * 17, [0] * 17
    code = k = 0
    for l in range(1, 17):
        valptr[l] = k
        mincode[l] = code
        code += counts[l]
        k += counts[l]
        maxcode[l] = code - 1 if counts[l] else -1
        code <<= 1
    return mincode, maxcode, valptr, vals


class _Bits:
    def __init__(self, scan):
        self.s, self.p, self.buf, self.cnt, self.marker = scan, 0, 0, 0, False

    def bit(self):
        if self.cnt == 0:
            b = 0
            if not self.marker and self.p < len(self.s):
                b = self.s[self.p]
                if b == 0xFF:
                    b2 = self.s[self.p + 1] if self.p + 1 < len(self.s) else 0xD9
                    if b2 == 0:
                        self.p += 2
                    else:
                        self.marker, b = True, 0
                else:
                    self.p += 1
            self.buf, self.cnt = b, 8
        self.cnt -= 1
        return (self.buf >> self.cnt) & 1

    def receive(self, n):
        v = 0
        for _ in range(n):
            v = (v << 1) | self.bit()
        return v

    def restart(self):
        """E.2.4: drop the pad bits, step over the RSTn marker the interval ends with (the reader never reads past a
        marker, so `p` sits on it whether or not it has been seen yet)"""
        self.cnt = 0
        if self.p + 1 < len(self.s) and self.s[self.p] == 0xFF and 0xD0 <= self.s[self.p + 1] <= 0xD7:
            self.p += 2
        self.marker = False


def _decode(bits, tab):
    mincode, maxcode, valptr, vals = tab
    code, l = bits.bit(), 1
    while l <= 16 and code > maxcode[l]:
        code = (code << 1) | bits.bit()
        l += 1
    if l > 16:
        return 0
    return vals[valptr[l] + code - mincode[l]]


def _extend(v, t):
    return v - (1 << t) + 1 if v < (1 << (t - 1)) else v


def decode_coefficients(info):
    """-> int32 [blocks_h, blocks_w, 64] dequantised coefficients in natural (row-major) order"""
    bw, bh = (info["width"] + 7) // 8, (info["height"] + 7) // 8
    dct, act = _tables(*info["dc"]), _tables(*info["ac"])
    q = info["q"]
    bits = _Bits(info["scan"])
    coef = np.zeros((bh, bw, 64), np.int32)
    pred, left = 0, info["restart"]
    for by in range(bh):
        for bx in range(bw):
            if info["restart"]:
                if left == 0:
                    bits.restart()
                    pred, left = 0, info["restart"]
                left -= 1
            t = _decode(bits, dct)
            if t:
                pred += _extend(bits.receive(t), t)
            c = coef[by, bx]
            c[0] = pred * q[0]
            k = 1
            while k < 64:
                rs = _decode(bits, act)
                r, s = rs >> 4, rs & 15
                if s == 0:
                    if r != 15:
                        break
                    k += 16
                    continue
                k += r
                if k > 63:
                    break
                c[ZIGZAG[k]] = _extend(bits.receive(s), s) * q[k]
                k += 1
    return coef


def _idct_1d(v, shift):
    """jidctint.c, one pass over the last axis (8 values), int64 to stay clear of numpy overflow warnings; the values
    fit 32 bits exactly as in the C code"""
    v = v.astype(np.int64)
    F = dict(a=2446, b=3196, c=4433, d=6270, e=7373, f=9633, g=12299, h=15137, i=16069, j=16819, k=20995, l=25172)
    z2, z3 = v[..., 2], v[..., 6]
    z1 = (z2 + z3) * F["c"]
    tmp2 = z1 + z3 * (-F["h"])
    tmp3 = z1 + z2 * F["d"]
    z2, z3 = v[..., 0], v[..., 4]
    tmp0, tmp1 = (z2 + z3) << 13, (z2 - z3) << 13
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = v[..., 7], v[..., 5], v[..., 3], v[..., 1]
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * F["f"]
    tmp0, tmp1, tmp2, tmp3 = tmp0 * F["a"], tmp1 * F["j"], tmp2 * F["l"], tmp3 * F["g"]
    z1, z2, z3, z4 = z1 * -F["e"], z2 * -F["k"], z3 * -F["i"] + z5, z4 * -F["b"] + z5
    tmp0, tmp1, tmp2, tmp3 = tmp0 + z1 + z3, tmp1 + z2 + z4, tmp2 + z2 + z3, tmp3 + z1 + z4
    rnd = 1 << (shift - 1)
    out = np.stack([tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0, tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2,
                    tmp10 - tmp3], -1)
    return (out + rnd) >> shift


def idct_islow(coef):
    """[..., 64] dequantised coefficients -> uint8 [..., 8, 8] samples"""
    c = coef.reshape(coef.shape[:-1] + (8, 8))
    ws = np.swapaxes(_idct_1d(np.swapaxes(c, -1, -2), 13 - 2), -1, -2)      # pass 1 runs down the columns
    px = _idct_1d(ws, 13 + 2 + 3)
    i = px & 1023                                                             # libjpeg's range_limit table
    return np.where(i < 128, i + 128, np.where(i < 512, 255, np.where(i < 896, 0, i - 896))).astype(np.uint8)


def decode(data):
    """bytes of a baseline grayscale JPEG file -> uint8 [H, W]"""
    info = parse(data)
    blocks = idct_islow(decode_coefficients(info))
    bh, bw = blocks.shape[:2]
    img = blocks.transpose(0, 2, 1, 3).reshape(bh * 8, bw * 8)
    return np.ascontiguousarray(img[:info["height"], :info["width"]])
