/*
 * ORACLE (test infrastructure only) -- FLAC (RFC 9639) for the chain's output format: mono, 16 bit, fixed block size
 * (the reference writes its result through libavcodec's FLAC encoder: mono / 44.1 kHz / s16 / 4096-sample frames,
 * internal/processor/encoder.go:92-101, processor.go:379-384).
 *
 *  orc_flac_encode_s16   sequential restatement of the encoder decisions the CUDA kernel (csrc/k_flac.cu) takes, so the two
 *                        streams can be compared BYTE for byte: CONSTANT subframe when all samples are equal; otherwise the
 *                        FIXED predictor order 0..4 with the smallest sum of |residual|; Rice partition order 0..6 (block
 *                        size a multiple of 64) with the parameter of each partition chosen as libavcodec's flacenc does
 *                        (find_optimal_param: k = log2((sum - n/2) / n)) and the order with the fewest exact bits; VERBATIM
 *                        when that is not smaller than 16 bits per sample.  STREAMINFO carries min / max frame size, the
 *                        sample count and an all-zero MD5 ("not known").
 *  orc_flac_decode_s16   independent decoder of the subset above (plus LPC subframes, for streams of other encoders).
 *  orc_flac_decode_pcm   the INPUT side (internal/audio/reader.go:29-188: libavformat flac demuxer + libavcodec flacdec.c): any
 *                        stream of RFC 9639 with 1..8 channels and 4..24 bits -- stereo decorrelation, wasted bits, variable
 *                        block size -- as interleaved int32 in the decoder's output scale (s16 streams: value << (16 - bps);
 *                        wider: value << (32 - bps), flacdec.c's sample_shift).  Checker of csrc/k_flac_dec.cu; pinned on the real
 *                        libavformat + libavcodec reader (oracle/ref_wav_probe.c) in tests/test_oracle_flac.py.
 *
 * Pinned: tests/test_oracle_flac.py decodes the encoder's streams with the REAL FFmpeg 8.0.1 libavcodec FLAC decoder found in
 * this image (oracle/ref_flac.py) and requires the PCM back bit for bit; the decoder is pinned by decoding the same streams.
 */
#include "orc.h"
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* ---- bit writer (MSB first) ---- */
typedef struct { uint8_t *p; int64_t cap, bit; int overflow; } BW;
static void bw_put(BW *w, int n, uint64_t v)
{
    for (int i = n - 1; i >= 0; i--) {
        int64_t byte = w->bit >> 3;
        if (byte >= w->cap) { w->overflow = 1; return; }
        if ((v >> i) & 1) w->p[byte] |= (uint8_t)(0x80 >> (w->bit & 7));
        w->bit++;
    }
}
static void bw_unary(BW *w, uint32_t q) { w->bit += q; bw_put(w, 1, 1); }      /* q zeros then a one: the buffer is pre-zeroed */

static uint8_t crc8(const uint8_t *p, int64_t n)
{
    uint8_t c = 0;
    for (int64_t i = 0; i < n; i++) { c ^= p[i]; for (int b = 0; b < 8; b++) c = (uint8_t)((c & 0x80) ? (c << 1) ^ 0x07 : (c << 1)); }
    return c;
}
static uint16_t crc16(const uint8_t *p, int64_t n)
{
    uint16_t c = 0;
    for (int64_t i = 0; i < n; i++) { c ^= (uint16_t)(p[i] << 8); for (int b = 0; b < 8; b++) c = (uint16_t)((c & 0x8000) ? (c << 1) ^ 0x8005 : (c << 1)); }
    return c;
}

static int rate_code(int rate)
{
    switch (rate) {
    case 88200: return 1; case 176400: return 2; case 192000: return 3; case 8000: return 4; case 16000: return 5; case 22050: return 6;
    case 24000: return 7; case 32000: return 8; case 44100: return 9; case 48000: return 10; case 96000: return 11;
    }
    return 0;                                   /* take it from STREAMINFO */
}
static int utf8_put(uint8_t *o, uint64_t v)    /* the "UTF-8" coded frame number, up to 36 bits */
{
    if (v < 0x80) { o[0] = (uint8_t)v; return 1; }
    int n = v < 0x800 ? 2 : v < 0x10000 ? 3 : v < 0x200000 ? 4 : v < 0x4000000 ? 5 : v < 0x80000000ull ? 6 : 7;
    static const uint8_t lead[8] = {0, 0, 0xC0, 0xE0, 0xF0, 0xF8, 0xFC, 0xFE};
    for (int i = n - 1; i > 0; i--) { o[i] = (uint8_t)(0x80 | (v & 0x3F)); v >>= 6; }
    o[0] = (uint8_t)(lead[n] | v);
    return n;
}

static inline uint32_t fold(int32_t r) { return ((uint32_t)r << 1) ^ (uint32_t)(r >> 31); }
static inline int ilog2_u32(uint32_t v) { int l = 0; while (v >>= 1) l++; return l; }
static int optimal_param(uint64_t sum, int n)  /* flacenc.c find_optimal_param, max 14 */
{
    if (sum <= (uint64_t)(n >> 1)) return 0;
    uint64_t q = (sum - (uint64_t)(n >> 1)) / (uint64_t)n;
    if (q > 0x7fffffffull) q = 0x7fffffffull;
    int k = ilog2_u32((uint32_t)q);
    return k > 14 ? 14 : k;
}

/* residual of the fixed predictor of `order` at sample i (i >= order) */
static inline int32_t fixed_res(const int16_t *x, int64_t i, int order)
{
    switch (order) {
    case 0: return x[i];
    case 1: return (int32_t)x[i] - x[i - 1];
    case 2: return (int32_t)x[i] - 2 * x[i - 1] + x[i - 2];
    case 3: return (int32_t)x[i] - 3 * x[i - 1] + 3 * x[i - 2] - x[i - 3];
    default: return (int32_t)x[i] - 4 * x[i - 1] + 6 * x[i - 2] - 4 * x[i - 3] + x[i - 4];
    }
}

#define ORC_FLAC_MAX_PORDER 6

/* ---- LPC subframes (RFC 9639 section 9.2.6): the analysis is integer up to the autocorrelation, so the CUDA encoder and this one
 * take the same decisions bit for bit.  Window: Welch, Q15, integer formula; windowed samples cut to 23 bits so that the 9 lags
 * of a 4096-sample block sum exactly in int64; Levinson-Durbin in double with every operation rounded on its own (no fused
 * multiply-add on either side); 12-bit coefficients with the shift and error feedback of flacenc.c's quantize_lpc_coefs, so the
 * prediction of 16-bit samples up to order 8 fits 32-bit arithmetic. ---- */
#define ORC_LPC_MAX 8
#define ORC_LPC_PREC 12
static int32_t welch_q15(int i, int n)
{
    const int64_t t = 2 * (int64_t)i - (n - 1), d = (int64_t)(n + 1) * (n + 1);
    const int64_t w = 32767 - (t * t * 32767) / d;
    return (int32_t)(w < 0 ? 0 : w);
}
/* coefficients (quantised) and shift of every order 1..maxo; returns the number of usable orders (0: no LPC for this block) */
static int lpc_analyse(const int16_t *x, int bs, int maxo, int32_t q[ORC_LPC_MAX + 1][ORC_LPC_MAX], int shift[ORC_LPC_MAX + 1])
{
    int64_t R[ORC_LPC_MAX + 1];
    int32_t *xw = (int32_t *)malloc(sizeof(int32_t) * (size_t)bs);
    for (int i = 0; i < bs; i++) xw[i] = ((int32_t)x[i] * welch_q15(i, bs)) >> 8;
    for (int k = 0; k <= maxo; k++) { int64_t a = 0; for (int i = k; i < bs; i++) a += (int64_t)xw[i] * xw[i - k]; R[k] = a; }
    free(xw);
    if (R[0] <= 0) return 0;
    volatile double err = (double)R[0];
    double lpc[ORC_LPC_MAX], tmp[ORC_LPC_MAX];
    int usable = 0;
    for (int i = 0; i < maxo; i++) {
        volatile double acc = (double)R[i + 1];
        for (int j = 0; j < i; j++) { volatile double m = lpc[j] * (double)R[i - j]; acc = acc - m; }
        const double k = acc / err;
        for (int j = 0; j < i; j++) { volatile double m = k * lpc[i - 1 - j]; tmp[j] = lpc[j] - m; }
        for (int j = 0; j < i; j++) lpc[j] = tmp[j];
        lpc[i] = k;
        { volatile double kk = k * k; volatile double om = 1.0 - kk; err = err * om; }
        if (!(err > 0.0)) break;
        /* quantise order i + 1 */
        const int o = i + 1, qmax = (1 << (ORC_LPC_PREC - 1)) - 1;
        double cmax = 0.0;
        for (int j = 0; j < o; j++) { const double a = lpc[j] < 0 ? -lpc[j] : lpc[j]; if (a > cmax) cmax = a; }
        int sh = 14;
        while (sh > 0 && cmax * (double)(1 << sh) > (double)qmax) sh--;
        volatile double e = 0.0;
        for (int j = 0; j < o; j++) {
            volatile double m = lpc[j] * (double)(1 << sh);
            e = e + m;
            long long v = llrint(e);
            if (v > qmax) v = qmax;
            if (v < -qmax) v = -qmax;
            q[o][j] = (int32_t)v;
            e = e - (double)v;
        }
        shift[o] = sh;
        usable = o;
    }
    return usable;
}
/* residual of the predictor (coefficients c[0..order-1] on x[i-1], x[i-2], ..., result >> sh) at sample i >= order */
static inline int32_t pred_res(const int16_t *x, int64_t i, int order, const int32_t *c, int sh)
{
    int32_t a = 0;
    for (int j = 0; j < order; j++) a += c[j] * (int32_t)x[i - 1 - j];
    return (int32_t)x[i] - (a >> sh);
}
static uint64_t est_bits(uint64_t sum_abs, int n, uint64_t overhead)
{
    const uint64_t S = 2 * sum_abs;
    const int k = optimal_param(S, n);
    return overhead + (uint64_t)n * (uint64_t)(k + 1) + (S >> k);
}

/* one frame; returns bytes written */
static int64_t encode_frame(const int16_t *x, int bs, int nominal_bs, int rate, uint64_t frame_no, uint8_t *out, int64_t cap)
{
    memset(out, 0, (size_t)cap);
    BW w = {out, cap, 0, 0};
    /* frame header */
    bw_put(&w, 14, 0x3FFE); bw_put(&w, 1, 0); bw_put(&w, 1, 0);                /* sync, reserved, fixed block size */
    const int bs_code = bs == 4096 ? 12 : 7;                                   /* 1100 = 4096, 0111 = 16-bit (bs - 1) follows */
    bw_put(&w, 4, (uint64_t)bs_code); bw_put(&w, 4, (uint64_t)rate_code(rate));
    bw_put(&w, 4, 0); bw_put(&w, 3, 4); bw_put(&w, 1, 0);                      /* mono, 16 bit, reserved */
    uint8_t u8[8]; int nu = utf8_put(u8, frame_no);
    for (int i = 0; i < nu; i++) bw_put(&w, 8, u8[i]);
    if (bs_code == 7) bw_put(&w, 16, (uint64_t)(bs - 1));
    bw_put(&w, 8, crc8(out, w.bit >> 3));
    (void)nominal_bs;
    /* subframe decision */
    int constant = 1;
    for (int i = 1; i < bs; i++) if (x[i] != x[0]) { constant = 0; break; }
    if (constant) {
        bw_put(&w, 8, 0x00); bw_put(&w, 16, (uint16_t)x[0]);
    } else {
        /* candidates: FIXED orders 0..4, then LPC orders 1..8; each scored by the Rice-cost estimate of its sum |residual|
         * plus its header; the first minimum wins */
        static const int32_t fixed_c[5][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};
        int32_t lq[ORC_LPC_MAX + 1][ORC_LPC_MAX]; int lsh[ORC_LPC_MAX + 1];
        const int n_lpc = bs > ORC_LPC_MAX + 1 ? lpc_analyse(x, bs, ORC_LPC_MAX, lq, lsh) : 0;
        int order = 0, is_lpc = 0, sh = 0; int32_t coef[ORC_LPC_MAX] = {0}; uint64_t best_est = ~0ull;
        for (int o = 0; o <= 4 && o < bs; o++) {
            uint64_t e = 0;
            for (int i = o; i < bs; i++) { int32_t r = pred_res(x, i, o, fixed_c[o], 0); e += (uint64_t)(r < 0 ? -(int64_t)r : r); }
            const uint64_t b = est_bits(e, bs - o, 16ull * (uint64_t)o);
            if (b < best_est) { best_est = b; order = o; is_lpc = 0; sh = 0; memset(coef, 0, sizeof(coef)); memcpy(coef, fixed_c[o], sizeof(int32_t) * 4); }
        }
        for (int o = 1; o <= n_lpc; o++) {
            uint64_t e = 0;
            for (int i = o; i < bs; i++) { int32_t r = pred_res(x, i, o, lq[o], lsh[o]); e += (uint64_t)(r < 0 ? -(int64_t)r : r); }
            const uint64_t b = est_bits(e, bs - o, 16ull * (uint64_t)o + 9 + (uint64_t)ORC_LPC_PREC * (uint64_t)o);
            if (b < best_est) { best_est = b; order = o; is_lpc = 1; sh = lsh[o]; memset(coef, 0, sizeof(coef)); memcpy(coef, lq[o], sizeof(int32_t) * (size_t)o); }
        }
        int pmax = 0;
        if (bs % 64 == 0) { pmax = ORC_FLAC_MAX_PORDER; while (pmax > 0 && (bs >> pmax) <= order) pmax--; }
        int best_p = 0; uint64_t best_bits = ~0ull; int best_k[1 << ORC_FLAC_MAX_PORDER];
        for (int p = 0; p <= pmax; p++) {
            const int psz = bs >> p; uint64_t bits = 0; int ks[1 << ORC_FLAC_MAX_PORDER];
            for (int j = 0; j < (1 << p); j++) {
                const int a = j * psz < order ? order : j * psz, b = (j + 1) * psz, n = b - a;
                uint64_t sum = 0;
                for (int i = a; i < b; i++) sum += fold(pred_res(x, i, order, coef, sh));
                const int k = optimal_param(sum, n);
                uint64_t shs = 0;
                for (int i = a; i < b; i++) shs += fold(pred_res(x, i, order, coef, sh)) >> k;
                bits += 4 + (uint64_t)n * (uint64_t)(k + 1) + shs;
                ks[j] = k;
            }
            if (bits < best_bits) { best_bits = bits; best_p = p; memcpy(best_k, ks, sizeof(int) * (size_t)(1 << p)); }
        }
        const uint64_t head_bits = 8 + 16ull * (uint64_t)order + (is_lpc ? 9 + (uint64_t)ORC_LPC_PREC * (uint64_t)order : 0) + 6;
        const uint64_t coded_bits = head_bits + best_bits, verbatim_bits = 8 + 16ull * (uint64_t)bs;
        if (coded_bits >= verbatim_bits) {
            bw_put(&w, 8, 0x02);
            for (int i = 0; i < bs; i++) bw_put(&w, 16, (uint16_t)x[i]);
        } else {
            bw_put(&w, 8, (uint64_t)((is_lpc ? (0x20 | (order - 1)) : (0x08 | order)) << 1));
            for (int i = 0; i < order; i++) bw_put(&w, 16, (uint16_t)x[i]);
            if (is_lpc) {
                bw_put(&w, 4, ORC_LPC_PREC - 1); bw_put(&w, 5, (uint64_t)sh);
                for (int j = 0; j < order; j++) bw_put(&w, ORC_LPC_PREC, (uint64_t)(uint32_t)coef[j] & ((1u << ORC_LPC_PREC) - 1));
            }
            bw_put(&w, 2, 0); bw_put(&w, 4, (uint64_t)best_p);
            const int psz = bs >> best_p;
            for (int j = 0; j < (1 << best_p); j++) {
                const int a = j * psz < order ? order : j * psz, b = (j + 1) * psz, k = best_k[j];
                bw_put(&w, 4, (uint64_t)k);
                for (int i = a; i < b; i++) { const uint32_t u = fold(pred_res(x, i, order, coef, sh)); bw_unary(&w, u >> k); if (k) bw_put(&w, k, u & ((1u << k) - 1)); }
            }
        }
    }
    if (w.bit & 7) w.bit += 8 - (w.bit & 7);
    if (w.overflow || (w.bit >> 3) + 2 > cap) return -1;
    const uint16_t c = crc16(out, w.bit >> 3);
    bw_put(&w, 16, c);
    return w.bit >> 3;
}

int64_t orc_flac_max_bytes(int64_t n, int block_size)
{
    const int64_t frames = (n + block_size - 1) / block_size;
    return 4 + 4 + 34 + frames * (16 + 2 * (int64_t)block_size + 2) + 64;
}

/* whole stream; returns bytes, or -1 when `cap` is short */
int64_t orc_flac_encode_s16(const int16_t *x, int64_t n, int rate, int block_size, uint8_t *out, int64_t cap)
{
    if (block_size < 16 || block_size > 65535 || cap < 42) return -1;
    const int64_t frames = (n + block_size - 1) / block_size;
    int64_t pos = 42; uint32_t min_fs = 0xFFFFFF, max_fs = 0;
    const int64_t fcap = 16 + 2 * (int64_t)block_size + 2;
    uint8_t *tmp = (uint8_t *)malloc((size_t)fcap);
    for (int64_t f = 0; f < frames; f++) {
        const int bs = (int)((n - f * block_size) < block_size ? (n - f * block_size) : block_size);
        const int64_t sz = encode_frame(x + f * block_size, bs, block_size, rate, (uint64_t)f, tmp, fcap);
        if (sz < 0 || pos + sz > cap) { free(tmp); return -1; }
        memcpy(out + pos, tmp, (size_t)sz); pos += sz;
        if ((uint32_t)sz < min_fs) min_fs = (uint32_t)sz;
        if ((uint32_t)sz > max_fs) max_fs = (uint32_t)sz;
    }
    free(tmp);
    if (frames == 0) min_fs = 0;
    memset(out, 0, 42);
    memcpy(out, "fLaC", 4);
    BW w = {out, 42, 32, 0};
    bw_put(&w, 1, 1); bw_put(&w, 7, 0); bw_put(&w, 24, 34);                    /* last metadata block, STREAMINFO, 34 bytes */
    bw_put(&w, 16, (uint64_t)block_size); bw_put(&w, 16, (uint64_t)block_size);
    bw_put(&w, 24, min_fs); bw_put(&w, 24, max_fs);
    bw_put(&w, 20, (uint64_t)rate); bw_put(&w, 3, 0); bw_put(&w, 5, 15); bw_put(&w, 36, (uint64_t)n);
    return pos;                                                                 /* MD5 stays zero: "not known" */
}

/* ---- decoder ---- */
typedef struct { const uint8_t *p; int64_t n, bit; int err; } BR;
static uint64_t br_get(BR *r, int n)
{
    uint64_t v = 0;
    for (int i = 0; i < n; i++) {
        const int64_t byte = r->bit >> 3;
        if (byte >= r->n) { r->err = 1; return 0; }
        v = (v << 1) | ((r->p[byte] >> (7 - (r->bit & 7))) & 1);
        r->bit++;
    }
    return v;
}
static int32_t br_sget(BR *r, int n) { uint64_t v = br_get(r, n); return (int32_t)((int64_t)(v << (64 - n)) >> (64 - n)); }
static uint32_t br_unary(BR *r) { uint32_t q = 0; while (!r->err && br_get(r, 1) == 0) q++; return q; }

/* returns samples decoded (mono, <= 16 bit), negative on a malformed stream: -2 CRC-8, -3 CRC-16, -4 syntax */
int64_t orc_flac_decode_s16(const uint8_t *in, int64_t nbytes, int16_t *out, int64_t cap, int *rate_out)
{
    if (nbytes < 42 || memcmp(in, "fLaC", 4)) return -4;
    BR r = {in, nbytes, 32, 0};
    int last = 0, rate = 0, bps = 16; int64_t total = 0;
    while (!last) {
        last = (int)br_get(&r, 1); const int type = (int)br_get(&r, 7); const int len = (int)br_get(&r, 24);
        if (type == 0) {
            br_get(&r, 16); br_get(&r, 16); br_get(&r, 24); br_get(&r, 24);
            rate = (int)br_get(&r, 20); const int ch = (int)br_get(&r, 3) + 1; bps = (int)br_get(&r, 5) + 1; total = (int64_t)br_get(&r, 36);
            if (ch != 1 || bps > 16) return -4;
            r.bit += 128;
        } else r.bit += 8ll * len;
        if (r.err) return -4;
    }
    if (rate_out) *rate_out = rate;
    int64_t pos = 0; int32_t *buf = (int32_t *)malloc(sizeof(int32_t) * 65536);
    while ((r.bit >> 3) < nbytes) {
        const int64_t f0 = r.bit >> 3;
        if (br_get(&r, 14) != 0x3FFE) { free(buf); return -4; }
        br_get(&r, 2);
        const int bsc = (int)br_get(&r, 4), src = (int)br_get(&r, 4), chc = (int)br_get(&r, 4), szc = (int)br_get(&r, 3);
        br_get(&r, 1);
        if (chc != 0 || (szc != 4 && szc != 0)) { free(buf); return -4; }
        int b0 = (int)br_get(&r, 8), extra = 0;                               /* coded number */
        if (b0 >= 0xC0) { int m = 0x20; extra = 1; while (b0 & m) { extra++; m >>= 1; } }
        for (int i = 0; i < extra; i++) br_get(&r, 8);
        int bs;
        if (bsc == 6) bs = (int)br_get(&r, 8) + 1; else if (bsc == 7) bs = (int)br_get(&r, 16) + 1;
        else if (bsc == 1) bs = 192; else if (bsc >= 2 && bsc <= 5) bs = 576 << (bsc - 2); else if (bsc >= 8) bs = 256 << (bsc - 8); else { free(buf); return -4; }
        if (src == 12) br_get(&r, 8); else if (src == 13 || src == 14) br_get(&r, 16);
        const uint8_t c8 = (uint8_t)br_get(&r, 8);
        if (r.err || crc8(in + f0, (r.bit >> 3) - 1 - f0) != c8) { free(buf); return -2; }
        /* subframe */
        if (br_get(&r, 1)) { free(buf); return -4; }
        const int type = (int)br_get(&r, 6); int wasted = 0;
        if (br_get(&r, 1)) wasted = (int)br_unary(&r) + 1;
        const int sb = bps - wasted;
        if (type == 0) { const int32_t v = br_sget(&r, sb); for (int i = 0; i < bs; i++) buf[i] = v; }
        else if (type == 1) { for (int i = 0; i < bs; i++) buf[i] = br_sget(&r, sb); }
        else if ((type >= 8 && type <= 12) || type >= 32) {
            const int lpc = type >= 32, order = lpc ? (type & 31) + 1 : type & 7;
            int32_t coef[32]; int shift = 0;
            for (int i = 0; i < order; i++) buf[i] = br_sget(&r, sb);
            if (lpc) { const int prec = (int)br_get(&r, 4) + 1; shift = br_sget(&r, 5); for (int i = 0; i < order; i++) coef[i] = br_sget(&r, prec); }
            const int method = (int)br_get(&r, 2), porder = (int)br_get(&r, 4), plen = method ? 5 : 4, esc = method ? 31 : 15;
            if (method > 1) { free(buf); return -4; }
            int i = order;
            for (int j = 0; j < (1 << porder); j++) {
                const int cnt = (bs >> porder) - (j == 0 ? order : 0), k = (int)br_get(&r, plen);
                if (k == esc) { const int nb = (int)br_get(&r, 5); for (int t = 0; t < cnt; t++) buf[i++] = nb ? br_sget(&r, nb) : 0; }
                else for (int t = 0; t < cnt; t++) { const uint32_t q = br_unary(&r); const uint32_t u = (q << k) | (uint32_t)br_get(&r, k); buf[i++] = (int32_t)(u >> 1) ^ -(int32_t)(u & 1); }
                if (r.err) { free(buf); return -4; }
            }
            for (i = order; i < bs; i++) {
                int64_t pred;
                if (lpc) { pred = 0; for (int t = 0; t < order; t++) pred += (int64_t)coef[t] * buf[i - 1 - t]; pred >>= shift; }
                else switch (order) {
                    case 0: pred = 0; break; case 1: pred = buf[i - 1]; break; case 2: pred = 2ll * buf[i - 1] - buf[i - 2]; break;
                    case 3: pred = 3ll * buf[i - 1] - 3ll * buf[i - 2] + buf[i - 3]; break;
                    default: pred = 4ll * buf[i - 1] - 6ll * buf[i - 2] + 4ll * buf[i - 3] - buf[i - 4];
                }
                buf[i] = (int32_t)(buf[i] + pred);
            }
        } else { free(buf); return -4; }
        if (r.bit & 7) r.bit += 8 - (r.bit & 7);
        const int64_t f1 = r.bit >> 3;
        const uint16_t c16 = (uint16_t)br_get(&r, 16);
        if (r.err || crc16(in + f0, f1 - f0) != c16) { free(buf); return -3; }
        for (int i = 0; i < bs; i++) { if (pos >= cap) { free(buf); return -1; } out[pos++] = (int16_t)(buf[i] * (1 << wasted)); }
    }
    free(buf);
    return (total && pos != total) ? -4 : pos;
}

/* ---- general decoder (input side) ---- */
static int dec_subframe(BR *r, int32_t *buf, int bs, int bps_sub)
{
    if (br_get(r, 1)) return -4;
    const int type = (int)br_get(r, 6); int wasted = 0;
    if (br_get(r, 1)) wasted = (int)br_unary(r) + 1;
    const int sb = bps_sub - wasted;
    if (sb < 1 || sb > 32) return -4;
    if (type == 0) { const int32_t v = br_sget(r, sb); for (int i = 0; i < bs; i++) buf[i] = v; }
    else if (type == 1) { for (int i = 0; i < bs; i++) buf[i] = br_sget(r, sb); }
    else if ((type >= 8 && type <= 12) || type >= 32) {
        const int lpc = type >= 32, order = lpc ? (type & 31) + 1 : type & 7;
        int32_t coef[32]; int shift = 0;
        if (order > bs) return -4;
        for (int i = 0; i < order; i++) buf[i] = br_sget(r, sb);
        if (lpc) {
            const int prec = (int)br_get(r, 4) + 1; if (prec == 16) return -4;
            shift = br_sget(r, 5); if (shift < 0) return -4;
            for (int i = 0; i < order; i++) coef[i] = br_sget(r, prec);
        }
        const int method = (int)br_get(r, 2), porder = (int)br_get(r, 4), plen = method ? 5 : 4, esc = method ? 31 : 15;
        if (method > 1 || (bs >> porder) < order || ((bs >> porder) << porder) != bs) return -4;
        int i = order;
        for (int j = 0; j < (1 << porder); j++) {
            const int cnt = (bs >> porder) - (j == 0 ? order : 0), k = (int)br_get(r, plen);
            if (k == esc) { const int nb = (int)br_get(r, 5); for (int t = 0; t < cnt; t++) buf[i++] = nb ? br_sget(r, nb) : 0; }
            else for (int t = 0; t < cnt; t++) { const uint32_t q = br_unary(r); const uint32_t u = (q << k) | (uint32_t)br_get(r, k); buf[i++] = (int32_t)(u >> 1) ^ -(int32_t)(u & 1); }
            if (r->err) return -4;
        }
        for (i = order; i < bs; i++) {
            int64_t pred;
            if (lpc) { pred = 0; for (int t = 0; t < order; t++) pred += (int64_t)coef[t] * buf[i - 1 - t]; pred >>= shift; }
            else switch (order) {
                case 0: pred = 0; break; case 1: pred = buf[i - 1]; break; case 2: pred = 2ll * buf[i - 1] - buf[i - 2]; break;
                case 3: pred = 3ll * buf[i - 1] - 3ll * buf[i - 2] + buf[i - 3]; break;
                default: pred = 4ll * buf[i - 1] - 6ll * buf[i - 2] + 4ll * buf[i - 3] - buf[i - 4];
            }
            buf[i] = (int32_t)(buf[i] + pred);
        }
    } else return -4;
    if (wasted) for (int i = 0; i < bs; i++) buf[i] = (int32_t)((uint32_t)buf[i] << wasted);
    return r->err ? -4 : 0;
}

/* returns FRAMES decoded (out: interleaved int32, cap frames), negative on a malformed stream: -2 CRC-8, -3 CRC-16, -4 syntax, -1 cap */
int64_t orc_flac_decode_pcm(const uint8_t *in, int64_t nbytes, int32_t *out, int64_t cap, int *rate_out, int *channels_out, int *bps_out)
{
    int64_t start = 0;
    if (nbytes >= 10 && !memcmp(in, "ID3", 3)) start = 10 + (((int64_t)(in[6] & 0x7F) << 21) | ((in[7] & 0x7F) << 14) | ((in[8] & 0x7F) << 7) | (in[9] & 0x7F));
    if (nbytes < start + 42 || memcmp(in + start, "fLaC", 4)) return -4;
    BR r = {in, nbytes, (start + 4) * 8, 0};
    int last = 0, rate = 0, bps = 16, ch = 1; int64_t total = 0;
    while (!last) {
        last = (int)br_get(&r, 1); const int type = (int)br_get(&r, 7); const int len = (int)br_get(&r, 24);
        if (type == 0) {
            br_get(&r, 16); br_get(&r, 16); br_get(&r, 24); br_get(&r, 24);
            rate = (int)br_get(&r, 20); ch = (int)br_get(&r, 3) + 1; bps = (int)br_get(&r, 5) + 1; total = (int64_t)br_get(&r, 36);
            r.bit += 128;
        } else r.bit += 8ll * len;
        if (r.err) return -4;
    }
    if (bps > 24 || bps < 4) return -4;
    if (rate_out) *rate_out = rate;
    if (channels_out) *channels_out = ch;
    if (bps_out) *bps_out = bps;
    const int out_shift = (bps <= 16 ? 16 : 32) - bps;
    int64_t pos = 0; int32_t *buf = (int32_t *)malloc(sizeof(int32_t) * 65536 * 8);
    while ((r.bit >> 3) < nbytes) {
        const int64_t f0 = r.bit >> 3;
        if (br_get(&r, 14) != 0x3FFE) { free(buf); return -4; }
        br_get(&r, 2);
        const int bsc = (int)br_get(&r, 4), src = (int)br_get(&r, 4), chc = (int)br_get(&r, 4), szc = (int)br_get(&r, 3);
        br_get(&r, 1);
        if (chc >= 11 || (chc < 8 ? chc + 1 : 2) != ch) { free(buf); return -4; }
        (void)szc;
        int b0 = (int)br_get(&r, 8), extra = 0;
        if (b0 >= 0xC0) { int m = 0x20; extra = 1; while (b0 & m) { extra++; m >>= 1; } }
        for (int i = 0; i < extra; i++) br_get(&r, 8);
        int bs;
        if (bsc == 6) bs = (int)br_get(&r, 8) + 1; else if (bsc == 7) bs = (int)br_get(&r, 16) + 1;
        else if (bsc == 1) bs = 192; else if (bsc >= 2 && bsc <= 5) bs = 576 << (bsc - 2); else if (bsc >= 8) bs = 256 << (bsc - 8); else { free(buf); return -4; }
        if (src == 12) br_get(&r, 8); else if (src == 13 || src == 14) br_get(&r, 16);
        const uint8_t c8 = (uint8_t)br_get(&r, 8);
        if (r.err || crc8(in + f0, (r.bit >> 3) - 1 - f0) != c8) { free(buf); return -2; }
        for (int c = 0; c < ch; c++) {
            const int side = (chc == 8 && c == 1) || (chc == 9 && c == 0) || (chc == 10 && c == 1);
            const int rc = dec_subframe(&r, buf + (size_t)c * 65536, bs, bps + side);
            if (rc) { free(buf); return rc; }
        }
        if (r.bit & 7) r.bit += 8 - (r.bit & 7);
        const int64_t f1 = r.bit >> 3;
        const uint16_t c16 = (uint16_t)br_get(&r, 16);
        if (r.err || crc16(in + f0, f1 - f0) != c16) { free(buf); return -3; }
        if (pos + bs > cap) { free(buf); return -1; }
        for (int i = 0; i < bs; i++) {
            if (ch == 2) {
                int32_t a = buf[i], s = buf[65536 + i], l, rr;
                if (chc == 8) { l = a; rr = a - s; }
                else if (chc == 9) { l = a + s; rr = s; }
                else if (chc == 10) { const int32_t m = (int32_t)(((uint32_t)a << 1) | ((uint32_t)s & 1u)); l = (m + s) >> 1; rr = (m - s) >> 1; }
                else { l = a; rr = s; }
                out[2 * (pos + i)] = (int32_t)((uint32_t)l << out_shift); out[2 * (pos + i) + 1] = (int32_t)((uint32_t)rr << out_shift);
            } else for (int c = 0; c < ch; c++) out[(size_t)ch * (pos + i) + c] = (int32_t)((uint32_t)buf[(size_t)c * 65536 + i] << out_shift);
        }
        pos += bs;
    }
    free(buf);
    return (total && pos != total) ? -4 : pos;
}
