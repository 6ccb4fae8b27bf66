"""Test infrastructure: a small FLAC *encoder* (RFC 9639 bit stream) used to build fixtures for the native reader
(csrc/flac.cpp).  No FLAC tool exists in the image, so the tests write their own streams: every subframe type
(CONSTANT, VERBATIM, FIXED 0-4, LPC with arbitrary coefficients), both Rice parameter widths, escape partitions,
wasted bits, the three stereo decorrelation modes, explicit / coded block sizes and sample rates, multi-byte frame
numbers, valid CRC-8 / CRC-16 and the STREAMINFO MD5.  Lossless by construction: residual = sample - prediction."""
import hashlib
import random

import numpy as np


class BitWriter:
    def __init__(self):
        self.acc, self.n, self.out = 0, 0, bytearray()

    def put(self, value, bits):
        if bits == 0:
            return
        self.acc = (self.acc << bits) | (int(value) & ((1 << bits) - 1))
        self.n += bits
        while self.n >= 8:
            self.n -= 8
            self.out.append((self.acc >> self.n) & 0xff)
        self.acc &= (1 << self.n) - 1

    def unary(self, zeros):
        while zeros >= 32:
            self.put(0, 32)
            zeros -= 32
        self.put(1, zeros + 1)

    def align(self):
        if self.n:
            self.put(0, 8 - self.n)


def crc8(data):
    c = 0
    for b in data:
        c ^= b
        for _ in range(8):
            c = ((c << 1) ^ 0x07) & 0xff if c & 0x80 else (c << 1) & 0xff
    return c


def crc16(data):
    c = 0
    for b in data:
        c ^= b << 8
        for _ in range(8):
            c = ((c << 1) ^ 0x8005) & 0xffff if c & 0x8000 else (c << 1) & 0xffff
    return c


def utf8_number(v):
    if v < 0x80:
        return bytes([v])
    out, n = [], 0
    while True:
        n += 1
        lead_bits = 6 - n
        if v < (1 << (6 * n + lead_bits)):
            break
    for _ in range(n):
        out.append(0x80 | (v & 0x3f))
        v >>= 6
    lead = (0xff << (7 - n)) & 0xff | v
    return bytes([lead] + out[::-1])


def _zigzag(r):
    return (r << 1) if r >= 0 else ((-r) << 1) - 1


def _write_residual(bw, res, blocksize, order, rng):
    method = rng.choice([0, 1])
    pbits, esc = (4, 15) if method == 0 else (5, 31)
    porders = [p for p in range(0, 9) if blocksize % (1 << p) == 0 and (blocksize >> p) >= max(order, 1)]
    porder = rng.choice(porders)
    bw.put(method, 2)
    bw.put(porder, 4)
    idx = 0
    for pt in range(1 << porder):
        count = (blocksize >> porder) - (order if pt == 0 else 0)
        part = res[idx:idx + count]
        idx += count
        if rng.random() < 0.15:
            need = max([1] + [int(abs(int(r))).bit_length() + 1 for r in part])
            bw.put(esc, pbits)
            bw.put(need, 5)
            for r in part:
                bw.put(int(r), need)
            continue
        mean = (sum(_zigzag(int(r)) for r in part) / max(1, len(part))) if len(part) else 0
        k = max(0, min(esc - 1, int(mean).bit_length() - 1 + rng.choice([-1, 0, 0, 1])))
        # keep the unary part bounded
        while any((_zigzag(int(r)) >> k) > 4000 for r in part) and k < esc - 1:
            k += 1
        bw.put(k, pbits)
        for r in part:
            u = _zigzag(int(r))
            bw.unary(u >> k)
            bw.put(u & ((1 << k) - 1), k)


def _write_subframe(bw, x, bps, rng):
    """x: python ints of one channel (already decorrelated); bps: bits of this subframe."""
    n = len(x)
    wasted = 0
    if any(x) and rng.random() < 0.5:
        while all((v >> wasted) & 1 == 0 for v in x) and wasted < bps - 1:
            wasted += 1
    if wasted:
        x = [v >> wasted for v in x]
    eff = bps - wasted
    kinds = ["verbatim", "fixed", "fixed", "lpc", "lpc"]
    if all(v == x[0] for v in x):
        kinds.append("constant")
        kinds.append("constant")
    kind = rng.choice(kinds)
    order, coefs, shift, prec = 0, [], 0, 0
    if kind == "fixed":
        order = rng.randint(0, min(4, n))
        coefs = {0: [], 1: [1], 2: [2, -1], 3: [3, -3, 1], 4: [4, -6, 4, -1]}[order]
    elif kind == "lpc":
        order = rng.randint(1, min(12, n))
        prec = rng.randint(2, 15)
        shift = rng.randint(max(0, prec - 2), min(15, prec + 2))     # |coef| / 2**shift <= 2: predictions stay in range
        lim = 1 << (prec - 1)
        coefs = [rng.randint(-lim, lim - 1) for _ in range(order)]
        if rng.random() < 0.5 and order >= 2:                        # close to a real second-order predictor
            coefs = [min(lim - 1, 2 << shift), -min(lim, 1 << shift)] + [0] * (order - 2)
    res = []
    if kind in ("fixed", "lpc"):
        for i in range(order, n):
            pred = sum(c * x[i - 1 - j] for j, c in enumerate(coefs)) >> shift
            res.append(x[i] - pred)
        if any(abs(r) >= (1 << 31) - 1 for r in res):                # residuals must fit 32 bits: code verbatim
            kind = "verbatim"
    code = {"constant": 0, "verbatim": 1}.get(kind, 8 + order if kind == "fixed" else 31 + order)
    bw.put(0, 1)
    bw.put(code, 6)
    if wasted:
        bw.put(1, 1)
        bw.unary(wasted - 1)
    else:
        bw.put(0, 1)
    if kind == "constant":
        bw.put(x[0], eff)
        return
    if kind == "verbatim":
        for v in x:
            bw.put(v, eff)
        return
    for v in x[:order]:
        bw.put(v, eff)
    if kind == "lpc":
        bw.put(prec - 1, 4)
        bw.put(shift, 5)
        for c in coefs:
            bw.put(c, prec)
    _write_residual(bw, res, n, order, rng)


def encode(pcm, sample_rate, bps, seed=0, blocksize=None, total_known=True, with_md5=True, id3=False):
    """pcm: int array [n, channels] -> FLAC bytes.  Frames use randomly chosen (valid) coding tools."""
    rng = random.Random(seed)
    pcm = np.asarray(pcm).astype(np.int64)
    n, ch = pcm.shape
    blocksize = blocksize or rng.choice([16, 192, 576, 1000, 1152, 4096, 4608])
    width = (bps + 7) // 8
    raw = pcm.astype({1: "i1", 2: "<i2", 3: "<i4", 4: "<i4"}[width]).tobytes()
    if width == 3:
        raw = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 4)[:, :3].tobytes()
    md5 = hashlib.md5(raw).digest() if with_md5 else bytes(16)
    si = BitWriter()
    si.put(blocksize, 16)
    si.put(blocksize, 16)
    si.put(0, 24)
    si.put(0, 24)
    si.put(sample_rate, 20)
    si.put(ch - 1, 3)
    si.put(bps - 1, 5)
    si.put(n if total_known else 0, 36)
    out = bytearray()
    if id3:
        out += b"ID3\x04\x00\x00" + bytes([0, 0, 0, 10]) + bytes(10)
    out += b"fLaC"
    pad = rng.random() < 0.5
    out += bytes([0x00 if pad else 0x80, 0, 0, 34]) + bytes(si.out) + md5
    if pad:
        out += bytes([0x81, 0, 0, 5]) + bytes(5)          # a PADDING block, last
    ss_code = {8: 1, 12: 2, 16: 4, 20: 5, 24: 6, 32: 7}.get(bps, 0)
    frame_no = 0
    for start in range(0, n, blocksize):
        block = pcm[start:start + blocksize]
        bs = len(block)
        bw = BitWriter()
        bw.put(0xfff8, 16)
        std = {192: 1, 576: 2, 1152: 3, 2304: 4, 4608: 5, 256: 8, 512: 9, 1024: 10, 2048: 11, 4096: 12, 8192: 13}
        if bs in std and rng.random() < 0.7:
            bs_code = std[bs]
        else:
            bs_code = 6 if bs <= 256 else 7
        sr_codes = {88200: 1, 176400: 2, 192000: 3, 8000: 4, 16000: 5, 22050: 6, 24000: 7, 32000: 8, 44100: 9, 48000: 10, 96000: 11}
        r = rng.random()
        if r < 0.4:
            sr_code = 0
        elif sample_rate in sr_codes and r < 0.8:
            sr_code = sr_codes[sample_rate]
        elif sample_rate % 1000 == 0 and sample_rate // 1000 < 256:
            sr_code = 12
        elif sample_rate < 65536:
            sr_code = 13
        elif sample_rate % 10 == 0 and sample_rate // 10 < 65536:
            sr_code = 14
        else:
            sr_code = 0
        mode = rng.choice([1, 8, 9, 10]) if ch == 2 else ch - 1
        bw.put(bs_code, 4)
        bw.put(sr_code, 4)
        bw.put(mode, 4)
        bw.put(ss_code if rng.random() < 0.7 else 0, 3)
        bw.put(0, 1)
        for b in utf8_number(frame_no):
            bw.put(b, 8)
        if bs_code == 6:
            bw.put(bs - 1, 8)
        elif bs_code == 7:
            bw.put(bs - 1, 16)
        if sr_code == 12:
            bw.put(sample_rate // 1000, 8)
        elif sr_code == 13:
            bw.put(sample_rate, 16)
        elif sr_code == 14:
            bw.put(sample_rate // 10, 16)
        bw.put(crc8(bytes(bw.out)), 8)
        cols = [[int(v) for v in block[:, c]] for c in range(ch)]
        if mode == 8:
            subs = [(cols[0], bps), ([a - b for a, b in zip(*cols)], bps + 1)]
        elif mode == 9:
            subs = [([a - b for a, b in zip(*cols)], bps + 1), (cols[1], bps)]
        elif mode == 10:
            subs = [([(a + b) >> 1 for a, b in zip(*cols)], bps), ([a - b for a, b in zip(*cols)], bps + 1)]
        else:
            subs = [(c, bps) for c in cols]
        for x, b in subs:
            _write_subframe(bw, x, b, rng)
        bw.align()
        bw.put(crc16(bytes(bw.out)), 16)
        out += bw.out
        frame_no += 1
    return bytes(out)
