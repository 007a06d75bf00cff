"""TEST INFRASTRUCTURE ONLY: a generator of syntactically valid FLAC streams (RFC 9639) that exercises what real encoders rarely
emit together -- every subframe type, LPC orders 1..32, wasted bits, escape-coded and 5-bit Rice partitions, variable block
sizes with sample numbers, all four stereo channel assignments, 8 / 12 / 16 / 20 / 24 bits, 1..8 channels, block-size and
sample-rate codes with trailing header bytes.  The streams are built from the DECODER's side (random warm-up samples,
contractive predictor coefficients, random residuals); their content is whatever comes out.  tests/ decode them with the oracle
(oracle/orc_flac.c), the real libavformat + libavcodec reader (oracle/ref_flac.py, where present) and the CUDA decoder."""
import numpy as np


def _crc_table(poly, bits):
    top = 1 << (bits - 1)
    mask = (1 << bits) - 1
    tab = []
    for i in range(256):
        c = i << (bits - 8)
        for _ in range(8):
            c = ((c << 1) ^ poly) & mask if c & top else (c << 1) & mask
        tab.append(c)
    return tab


_T8, _T16 = _crc_table(0x07, 8), _crc_table(0x8005, 16)


def crc8(b):
    c = 0
    for v in b:
        c = _T8[c ^ v]
    return c


def crc16(b):
    c = 0
    for v in b:
        c = ((c << 8) & 0xFFFF) ^ _T16[(c >> 8) ^ v]
    return c


class BitWriter:
    def __init__(self):
        self.out, self.v, self.n = bytearray(), 0, 0

    def put(self, nbits, val):
        if nbits:
            self.v = (self.v << nbits) | (val & ((1 << nbits) - 1))
            self.n += nbits
            if self.n >= 64:
                keep = self.n & 7
                self.out += (self.v >> keep).to_bytes((self.n - keep) // 8, "big")
                self.v &= (1 << keep) - 1
                self.n = keep

    def sput(self, nbits, val):
        self.put(nbits, val & ((1 << nbits) - 1))

    def unary(self, q):
        while q >= 32:
            self.put(32, 0)
            q -= 32
        self.put(q + 1, 1)

    def align(self):
        if self.n & 7:
            self.put(8 - (self.n & 7), 0)

    def bytes(self):
        self.align()
        if self.n:
            self.out += self.v.to_bytes(self.n // 8, "big")
            self.v, self.n = 0, 0
        return bytes(self.out)


def _utf8(v):
    if v < 0x80:
        return bytes([v])
    out, n = [], 0
    while True:
        n += 1
        lim = 1 << (6 * n + (6 - n))          # payload bits with n continuation bytes
        if v < lim:
            break
    first = ((0xFF << (7 - n)) & 0xFF) | (v >> (6 * n))
    out.append(first)
    for i in range(n - 1, -1, -1):
        out.append(0x80 | ((v >> (6 * i)) & 0x3F))
    return bytes(out)


_BS_CODES = {192: 1, 576: 2, 1152: 3, 2304: 4, 4608: 5, 256: 8, 512: 9, 1024: 10, 2048: 11, 4096: 12, 8192: 13, 16384: 14, 32768: 15}
_RATE_CODES = {88200: 1, 176400: 2, 192000: 3, 8000: 4, 16000: 5, 22050: 6, 24000: 7, 32000: 8, 44100: 9, 48000: 10, 96000: 11}
_SZ_CODES = {8: 1, 12: 2, 16: 4, 20: 5, 24: 6}


def _signal(rng, bs, lim):
    """a bounded integer signal: smoothed noise with a few steps, within +-lim"""
    x = rng.normal(0, 1, bs + 64)
    x = np.convolve(x, np.ones(int(rng.integers(1, 40))), "valid")[:bs]
    x = x / (np.abs(x).max() + 1e-9) * lim * rng.random()
    if rng.random() < 0.3:
        x[int(rng.integers(bs)):] += lim * 0.2 * rng.normal()
    return np.clip(np.round(x), -lim, lim - 1).astype(np.int64)


_FIXED = {0: [], 1: [1], 2: [2, -1], 3: [3, -3, 1], 4: [4, -6, 4, -1]}


def _subframe(w, rng, bs, sb, kinds):
    """writes one subframe of sb-bit samples: the signal is drawn first, then coded with RANDOM (valid, not good) decisions"""
    wasted = int(rng.integers(1, 4)) if (rng.random() < 0.2 and sb > 6) else 0
    sb -= wasted
    kind = kinds[int(rng.integers(len(kinds)))]
    # a quarter of the range: the decorrelated pair of a conforming stream adds up to in-range samples (libavcodec's SIMD
    # decorrelation saturates where C wraps, so only conforming streams have ONE right answer)
    lim = max(1, (1 << (sb - 1)) // 4)
    order = 0
    if kind == "constant":
        w.put(1, 0); w.put(6, 0)
    elif kind == "verbatim":
        w.put(1, 0); w.put(6, 1)
    else:
        order = min(int(rng.integers(0, 5)) if kind == "fixed" else int(rng.integers(1, 33)), bs)
        if kind == "lpc" and order == 0:
            order = 1
        w.put(1, 0); w.put(6, (8 | order) if kind == "fixed" else (32 | (order - 1)))
    if wasted:
        w.put(1, 1); w.unary(wasted - 1)
    else:
        w.put(1, 0)
    if kind == "constant":
        w.sput(sb, int(rng.integers(-lim, lim)))
        return
    x = _signal(rng, bs, lim)
    if kind == "verbatim":
        for v in x:
            w.sput(sb, int(v))
        return
    for v in x[:order]:
        w.sput(sb, int(v))
    shift = 0
    if kind == "lpc":
        prec = int(rng.integers(5, 16)); shift = int(rng.integers(0, min(prec, 15)))
        # keep sum |c| / 2^shift below ~2 so the residual of the bounded signal fits the sample width comfortably
        budget = 2.0 * (1 << shift)
        c = rng.integers(-(1 << (prec - 1)), 1 << (prec - 1), size=order).astype(np.float64)
        if np.abs(c).sum() > budget:
            c = np.trunc(c * budget / np.abs(c).sum())
        coef = [int(v) for v in c]
        w.put(4, prec - 1); w.sput(5, shift)
        for v in coef:
            w.sput(prec, v)
    else:
        coef = _FIXED[order]
    pred = np.zeros(bs, dtype=np.int64)
    for t, cv in enumerate(coef):
        pred[order:] += cv * x[order - 1 - t: bs - 1 - t]
    res = (x - (pred >> shift))[order:]
    # residual
    method = int(rng.random() < 0.3)
    pmax = 0
    while pmax < 8 and (bs >> (pmax + 1)) >= max(order, 1) and ((bs >> (pmax + 1)) << (pmax + 1)) == bs:
        pmax += 1
    porder = int(rng.integers(0, pmax + 1))
    w.put(2, method); w.put(4, porder)
    at = 0
    for part in range(1 << porder):
        cnt = (bs >> porder) - (order if part == 0 else 0)
        r = res[at: at + cnt]; at += cnt
        u = (r << 1) ^ (r >> 63)
        umax = int(u.max()) if cnt else 0
        need = 0 if (cnt == 0 or (r == 0).all()) else int(max(int(r.max()).bit_length(), int(-1 - int(r.min())).bit_length() if r.min() < 0 else 0)) + 1
        kmax = 30 if method else 14
        if rng.random() < 0.15 and need <= 31:
            nb = need if need == 0 or rng.random() < 0.5 else min(31, need + int(rng.integers(0, 3)))
            w.put(5 if method else 4, 31 if method else 15); w.put(5, nb)
            if nb:
                for v in r:
                    w.sput(nb, int(v))
        else:
            kopt = max(0, int(np.log2(u.mean() + 1))) if cnt else 0
            k = int(np.clip(kopt + int(rng.integers(-3, 3)), 0, kmax))
            while (umax >> k) > 3000 and k < kmax:
                k += 1
            w.put(5 if method else 4, k)
            for v in u:
                v = int(v)
                w.unary(v >> k); w.put(k, v & ((1 << k) - 1))


def make_stream(seed, n_frames=24, channels=2, bps=16, rate=44100, variable=False, kinds=("constant", "verbatim", "fixed", "lpc"),
                block_sizes=(4096, 1024, 576, 192, 256), odd_last=True, metadata_pad=0):
    """-> bytes of a complete stream (STREAMINFO total samples filled in)"""
    rng = np.random.default_rng(seed)
    frames, total = [], 0
    fixed_bs = int(block_sizes[0])
    sizes = []
    for f in range(n_frames):
        if variable:
            bs = int(block_sizes[int(rng.integers(len(block_sizes)))]) if rng.random() < 0.8 else int(rng.integers(16, 3000))
        else:
            bs = fixed_bs
        if f == n_frames - 1 and odd_last:
            bs = int(rng.integers(16, max(17, fixed_bs)))
        sizes.append(bs)
    max_bs = max(sizes)
    # how the rate and the sample size are coded is a property of the stream, not of the frame: libavcodec's flac parser scores a
    # header against its neighbours on the decoded header values (code 0 = "see STREAMINFO" reads as 0 there) and drops frames
    # whose fields "change" -- real encoders keep one coding, and so does this generator
    rate_code = _RATE_CODES.get(rate, 0)
    r = rng.random()
    if r < 0.15:
        rate_code = 0
    elif r < 0.3 and rate % 10 == 0 and rate // 10 < 65536:
        rate_code = 14
    elif r < 0.45 and rate < 65536:
        rate_code = 13
    elif r < 0.55 and rate % 1000 == 0 and rate // 1000 < 256:
        rate_code = 12
    if rate_code == 0 and rate not in _RATE_CODES:
        rate_code = 13 if rate < 65536 else 0
    size_code = _SZ_CODES[bps] if rng.random() < 0.8 else 0
    for f, bs in enumerate(sizes):
        w = BitWriter()
        w.put(14, 0x3FFE); w.put(1, 0); w.put(1, 1 if variable else 0)
        bsc = _BS_CODES.get(bs, 6 if bs <= 256 else 7)
        if bs in _BS_CODES and rng.random() < 0.2:
            bsc = 6 if bs <= 256 else 7                                  # the long form of a standard size
        rc = rate_code
        assign = channels - 1
        if channels == 2:
            assign = int(rng.choice([1, 8, 9, 10]))
        w.put(4, bsc); w.put(4, rc); w.put(4, assign)
        w.put(3, size_code); w.put(1, 0)
        for v in _utf8(total if variable else f):
            w.put(8, v)
        if bsc == 6:
            w.put(8, bs - 1)
        elif bsc == 7:
            w.put(16, bs - 1)
        if rc == 12:
            w.put(8, rate // 1000)
        elif rc == 13:
            w.put(16, rate)
        elif rc == 14:
            w.put(16, rate // 10)
        w.put(8, crc8(w.bytes()))
        for ch in range(channels):
            side = (assign == 8 and ch == 1) or (assign == 9 and ch == 0) or (assign == 10 and ch == 1)
            _subframe(w, rng, bs, bps + int(side), kinds)
        body = w.bytes()
        frames.append(body + crc16(body).to_bytes(2, "big"))
        total += bs
    si = BitWriter()
    si.put(16, min(sizes[:-1] or sizes) if variable else fixed_bs); si.put(16, max_bs)
    si.put(24, min(len(f) for f in frames)); si.put(24, max(len(f) for f in frames))
    si.put(20, rate); si.put(3, channels - 1); si.put(5, bps - 1); si.put(36, total); si.put(128, 0)
    out = b"fLaC"
    if metadata_pad:
        out += bytes([0x00, 0, 0, 34]) + si.bytes() + bytes([0x81]) + metadata_pad.to_bytes(3, "big") + bytes(metadata_pad)
    else:
        out += bytes([0x80, 0, 0, 34]) + si.bytes()
    return out + b"".join(frames)
