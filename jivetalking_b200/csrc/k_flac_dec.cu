// FLAC (RFC 9639) decoder for the chain's INPUT: what the reference's audio.Reader gets from libavformat's flac demuxer +
// libavcodec's flac decoder (internal/audio/reader.go:29-188) -- interleaved s16 for streams of <= 16 bits per sample,
// interleaved s32 (samples shifted up to fill 32 bits, as flacdec.c does) above; SURVEY 8f-3.
//
// A FLAC stream has no index of its frames and a frame's length is only known once it has been decoded, so a sequential
// decoder walks frame after frame.  Here the walk is replaced by three data-parallel steps:
//   1. k_fd_scan    every byte position is tested for a frame header: sync code, no reserved values, block size / sample
//                   rate / channel / bits fields consistent with STREAMINFO, CRC-8 of the header.  Two passes (count per
//                   chunk, host scan of ~10^4 counts, ordered write) give the candidate list sorted by position.  False
//                   candidates (a header-shaped byte run inside a frame's payload) are ~10 per hour of audio.
//   2. k_fd_link    one thread per candidate runs the frame CRC-16 forward from its position and tests, at every later
//                   candidate position (and at the end of the stream), whether the two bytes before it are that CRC: the
//                   first hit is the frame's end.  The host follows the links from the first frame (a false candidate is
//                   never linked to), and prefix-sums the block sizes into output offsets.
//   3. k_fd_frames  one thread per frame decodes its subframes (CONSTANT, VERBATIM, FIXED and LPC predictors -- the fixed
//                   predictors run as LPC with binomial coefficients -- partitioned Rice with 4 / 5 bit parameters and
//                   escape partitions, wasted bits) into planar int32; k_fd_output undoes the stereo decorrelation
//                   (left/side, right/side, mid/side) and writes the interleaved result, coalesced.
// The residual of a frame is one long variable-length code, so a frame is inherently sequential; parallelism is across the
// ~39 000 frames of an hour of audio.
#include "jt_internal.h"
#include "jt_device.cuh"
#include <cstdio>
#include <cstring>

namespace {

struct FdInfo {                 // STREAMINFO
    int min_bs = 0, max_bs = 0, rate = 0, channels = 0, bps = 0;
    uint32_t min_fs = 0, max_fs = 0;
    long long total = 0, audio_off = 0;
};

struct FdHeader { int bs, assign, hdr_len; unsigned long long number; int variable; };

struct FdCand { unsigned long long pos; unsigned long long number; uint32_t bs; uint32_t assign; };

struct FdFrame { unsigned long long pos; long long out0; uint32_t bs; uint32_t assign; };

__host__ __device__ inline uint8_t fd_crc8(const uint8_t *p, int n)
{
    uint8_t c = 0;
    for (int i = 0; i < n; i++) { c ^= p[i]; for (int b = 0; b < 8; b++) c = (uint8_t)((c & 0x80) ? (c << 1) ^ 0x07 : (c << 1)); }
    return c;
}

// frame header at p (avail bytes readable): true when it is a valid header of THIS stream
__host__ __device__ inline bool fd_parse_header(const uint8_t *p, long long avail, const FdInfo &I, FdHeader &h)
{
    if (avail < 6 || p[0] != 0xFF || (p[1] & 0xFE) != 0xF8) return false;
    h.variable = p[1] & 1;
    const int bsc = p[2] >> 4, src = p[2] & 15, chc = p[3] >> 4, szc = (p[3] >> 1) & 7;
    if (bsc == 0 || src == 15 || chc >= 11 || (p[3] & 1) || szc == 3) return false;
    const int bits = szc == 0 ? I.bps : (szc == 1 ? 8 : szc == 2 ? 12 : szc == 4 ? 16 : szc == 5 ? 20 : szc == 6 ? 24 : 32);
    if (bits != I.bps) return false;
    const int nch = chc < 8 ? chc + 1 : 2;
    if (nch != I.channels) return false;
    // "UTF-8" coded frame / sample number: 1..7 bytes
    int pos = 4; const int b0 = p[4]; int extra = 0; unsigned long long v;
    if (b0 < 0x80) v = (unsigned)b0;
    else if (b0 < 0xC0 || b0 == 0xFF) return false;
    else { int m = 0x20; extra = 1; while (b0 & m) { extra++; m >>= 1; } v = (unsigned)(b0 & (m - 1)); }
    if (avail < 5 + extra + 1) return false;
    for (int i = 0; i < extra; i++) { const int cbyte = p[5 + i]; if ((cbyte & 0xC0) != 0x80) return false; v = (v << 6) | (unsigned)(cbyte & 0x3F); }
    pos = 5 + extra;
    h.number = v;
    int bs;
    if (bsc == 6) { if (avail < pos + 2) return false; bs = p[pos] + 1; pos += 1; }
    else if (bsc == 7) { if (avail < pos + 3) return false; bs = ((p[pos] << 8) | p[pos + 1]) + 1; pos += 2; }
    else if (bsc == 1) bs = 192;
    else if (bsc <= 5) bs = 576 << (bsc - 2);
    else bs = 256 << (bsc - 8);
    int rate = 0;
    switch (src) {
    case 0: rate = I.rate; break; case 1: rate = 88200; break; case 2: rate = 176400; break; case 3: rate = 192000; break;
    case 4: rate = 8000; break; case 5: rate = 16000; break; case 6: rate = 22050; break; case 7: rate = 24000; break;
    case 8: rate = 32000; break; case 9: rate = 44100; break; case 10: rate = 48000; break; case 11: rate = 96000; break;
    case 12: if (avail < pos + 2) return false; rate = p[pos] * 1000; pos += 1; break;
    case 13: if (avail < pos + 3) return false; rate = (p[pos] << 8) | p[pos + 1]; pos += 2; break;
    default: if (avail < pos + 3) return false; rate = ((p[pos] << 8) | p[pos + 1]) * 10; pos += 2; break;
    }
    if (rate != I.rate) return false;
    if (bs > I.max_bs || avail < pos + 1) return false;
    if (fd_crc8(p, pos) != p[pos]) return false;
    h.bs = bs; h.assign = chc; h.hdr_len = pos + 1;
    return true;
}

// ---- 1. sync scan -----------------------------------------------------------------------------------------------------
// chunk c = bytes [c * chunk, (c + 1) * chunk) of the audio part; WRITE = false counts, WRITE = true stores in position order
template <bool WRITE>
__global__ void __launch_bounds__(256)
k_fd_scan(const uint8_t *__restrict__ b, long long n, FdInfo I, int chunk, long long n_chunks, uint32_t *__restrict__ counts,
          const unsigned long long *__restrict__ offsets, FdCand *__restrict__ cands)
{
    __shared__ uint32_t wcount[8];
    __shared__ uint32_t running;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const long long lo = c * chunk, hi = min(n, lo + (long long)chunk);
        if (threadIdx.x == 0) running = 0;
        __syncthreads();
        for (long long t0 = lo; t0 < hi; t0 += 256) {
            const long long p = t0 + threadIdx.x;
            bool hit = false; FdHeader h;
            if (p < hi && p + 1 < n && b[p] == 0xFF && (b[p + 1] & 0xFE) == 0xF8) hit = fd_parse_header(b + p, n - p, I, h);
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (!__syncthreads_or(m != 0)) continue;                       // nearly every tile
            if (lane == 0) wcount[warp] = __popc(m);
            __syncthreads();
            uint32_t base = running, tot = 0;
            for (int w = 0; w < 8; w++) { if (w < warp) base += wcount[w]; tot += wcount[w]; }
            if (WRITE && hit) {
                FdCand cd; cd.pos = (unsigned long long)p; cd.number = h.number; cd.bs = (uint32_t)h.bs; cd.assign = (uint32_t)h.assign;
                cands[offsets[c] + base + __popc(m & ((1u << lane) - 1))] = cd;
            }
            __syncthreads();
            if (threadIdx.x == 0) running += tot;
        }
        __syncthreads();
        if (!WRITE && threadIdx.x == 0) counts[c] = running;
        __syncthreads();
    }
}

// ---- 2. CRC-16 linking ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint16_t fd_crc16_tab(const uint16_t *tab, uint16_t c, uint8_t v) { return (uint16_t)((c << 8) ^ tab[(c >> 8) ^ v]); }

__global__ void __launch_bounds__(128)
k_fd_link(const uint8_t *__restrict__ b, long long n, const FdCand *__restrict__ cands, long long n_cands, long long max_frame_bytes,
          int32_t *__restrict__ next)
{
    __shared__ uint16_t tab[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        uint16_t c = (uint16_t)(i << 8);
        for (int k = 0; k < 8; k++) c = (uint16_t)((c & 0x8000) ? (c << 1) ^ 0x8005 : (c << 1));
        tab[i] = c;
    }
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cands) return;
    const long long p0 = (long long)cands[i].pos;
    long long p = p0; uint16_t crc = 0; int32_t found = -1;
    for (long long j = i + 1; j <= n_cands; j++) {
        const long long target = j < n_cands ? (long long)cands[j].pos : n;
        if (target - p0 > max_frame_bytes) break;
        const long long stop = target - 2;                    // CRC covers [p0, stop), stored big-endian at [stop, stop + 2)
        if (stop < p) continue;                               // (two candidates closer than the CRC itself)
        while (p < stop && ((uintptr_t)(b + p) & 3)) crc = fd_crc16_tab(tab, crc, b[p++]);
        for (; p + 4 <= stop; p += 4) {
            const uint32_t w = *(const uint32_t *)(b + p);
            crc = fd_crc16_tab(tab, crc, (uint8_t)w); crc = fd_crc16_tab(tab, crc, (uint8_t)(w >> 8));
            crc = fd_crc16_tab(tab, crc, (uint8_t)(w >> 16)); crc = fd_crc16_tab(tab, crc, (uint8_t)(w >> 24));
        }
        while (p < stop) crc = fd_crc16_tab(tab, crc, b[p++]);
        if (crc == (uint16_t)((b[stop] << 8) | b[stop + 1])) { found = (int32_t)j; break; }
    }
    next[i] = found;
}

// ---- 3. frames --------------------------------------------------------------------------------------------------------
struct FdBits {
    const uint32_t *w, *wend; uint64_t acc; int nbits; int err;
    __device__ __forceinline__ void refill()
    {
        if (nbits <= 32) {
            const uint32_t v = w < wend ? __byte_perm(*w, 0, 0x0123) : 0u;   // zeros past the end of the stream
            w++;
            acc |= (uint64_t)v << (32 - nbits); nbits += 32;
        }
    }
    __device__ __forceinline__ void init(const uint8_t *b, long long pos, long long n)
    {
        const uintptr_t a = (uintptr_t)(b + pos);
        w = (const uint32_t *)(a & ~(uintptr_t)3);
        wend = (const uint32_t *)(((uintptr_t)(b + n) + 3) & ~(uintptr_t)3);
        acc = 0; nbits = 0; err = 0;
        refill(); refill();
        const int skip = (int)(a & 3) * 8;
        acc <<= skip; nbits -= skip;
    }
    __device__ __forceinline__ uint32_t get(int k)            // 0 <= k <= 32
    {
        refill();
        if (k == 0) return 0;
        const uint32_t v = (uint32_t)(acc >> (64 - k));
        acc <<= k; nbits -= k;
        return v;
    }
    __device__ __forceinline__ int32_t sget(int k)            // 1 <= k <= 32
    {
        refill();
        const int32_t v = (int32_t)((int64_t)acc >> (64 - k));
        acc <<= k; nbits -= k;
        return v;
    }
    __device__ __forceinline__ uint32_t unary()               // zeros before the next one bit (consumes the one)
    {
        uint32_t q = 0;
        for (;;) {
            refill();
            if (acc == 0) { q += nbits; nbits = 0; if (w > wend + 1 || q > (1u << 26)) { err = 1; return 0; } continue; }
            const int lz = __clzll((long long)acc);
            q += lz; acc <<= lz; acc <<= 1; nbits -= lz + 1;
            return q;
        }
    }
    __device__ __forceinline__ long long byte_pos(const uint8_t *b) const   // next unread byte once byte aligned
    {
        return (long long)((const uint8_t *)w - b) - (nbits >> 3);
    }
};

// residual decode + prediction with at most T taps (coefficients beyond the order are zero); hist[0] is the newest sample
template <int T>
__device__ __forceinline__ void fd_predict_run(FdBits &br, int32_t *__restrict__ dst, int bs, int order, const int32_t *coef_in, int shift,
                                               int wasted, int method_bits, int porder)
{
    int32_t coef[T], hist[T];
#pragma unroll
    for (int t = 0; t < T; t++) { coef[t] = t < order ? coef_in[t] : 0; hist[t] = t < order ? (dst[order - 1 - t] >> wasted) : 0; }
    const int esc = method_bits == 5 ? 31 : 15;
    int i = order;
    for (int part = 0; part < (1 << porder); part++) {
        int cnt = (bs >> porder) - (part == 0 ? order : 0);
        const int k = (int)br.get(method_bits);
        const int raw_bits = k == esc ? (int)br.get(5) : -1;
        for (; cnt > 0; cnt--, i++) {
            int32_t r;
            if (raw_bits >= 0) r = raw_bits ? br.sget(raw_bits) : 0;
            else { const uint32_t q = br.unary(); const uint32_t u = (q << k) | br.get(k); r = (int32_t)(u >> 1) ^ -(int32_t)(u & 1); }
            long long pred = 0;
#pragma unroll
            for (int t = 0; t < T; t++) pred += (long long)coef[t] * hist[t];
            const int32_t x = r + (int32_t)(pred >> shift);
#pragma unroll
            for (int t = T - 1; t > 0; t--) hist[t] = hist[t - 1];
            hist[0] = x;
            dst[i] = x << wasted;
        }
        if (br.err) return;
    }
}

__global__ void __launch_bounds__(64)
k_fd_frames(const uint8_t *__restrict__ b, long long n, const FdFrame *__restrict__ frames, long long n_frames, FdInfo I,
            long long total, int32_t *__restrict__ planar, int *__restrict__ err_flag)
{
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    const FdFrame fr = frames[f];
    FdHeader h;
    if (!fd_parse_header(b + fr.pos, n - (long long)fr.pos, I, h)) { atomicOr(err_flag, 1); return; }
    FdBits br; br.init(b, (long long)fr.pos + h.hdr_len, n);
    const int bs = h.bs;
    for (int ch = 0; ch < I.channels; ch++) {
        int32_t *dst = planar + (size_t)ch * total + fr.out0;
        // the side channel of a decorrelated pair carries one more bit
        const int side = (h.assign == 8 && ch == 1) || (h.assign == 9 && ch == 0) || (h.assign == 10 && ch == 1);
        if (br.get(1)) { atomicOr(err_flag, 2); return; }
        const int type = (int)br.get(6);
        int wasted = 0;
        if (br.get(1)) wasted = (int)br.unary() + 1;
        const int sb = I.bps + side - wasted;
        if (sb < 1 || sb > 32) { atomicOr(err_flag, 2); return; }
        if (type == 0) {
            const int32_t v = br.sget(sb) << wasted;
            for (int i = 0; i < bs; i++) dst[i] = v;
        } else if (type == 1) {
            for (int i = 0; i < bs; i++) dst[i] = br.sget(sb) << wasted;
        } else if ((type >= 8 && type <= 12) || type >= 32) {
            const bool lpc = type >= 32; const int order = lpc ? (type & 31) + 1 : (type & 7);
            if (order > bs) { atomicOr(err_flag, 2); return; }
            int32_t coef[32]; int shift = 0;
            for (int i = 0; i < order; i++) dst[i] = br.sget(sb) << wasted;
            if (lpc) {
                const int prec = (int)br.get(4) + 1;
                if (prec == 16) { atomicOr(err_flag, 2); return; }
                shift = (int)br.sget(5);
                if (shift < 0) { atomicOr(err_flag, 2); return; }
                for (int i = 0; i < order; i++) coef[i] = br.sget(prec);
            } else {
                // fixed predictors = LPC with binomial coefficients, no shift
                const int32_t fx[5][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};
                for (int i = 0; i < 4; i++) coef[i] = fx[order][i];
            }
            const int method = (int)br.get(2), porder = (int)br.get(4);
            if (method > 1 || (bs >> porder) < order || ((bs >> porder) << porder) != bs) { atomicOr(err_flag, 2); return; }
            const int mb = method ? 5 : 4;
            if (order <= 4) fd_predict_run<4>(br, dst, bs, order, coef, shift, wasted, mb, porder);
            else if (order <= 8) fd_predict_run<8>(br, dst, bs, order, coef, shift, wasted, mb, porder);
            else if (order <= 12) fd_predict_run<12>(br, dst, bs, order, coef, shift, wasted, mb, porder);
            else fd_predict_run<32>(br, dst, bs, order, coef, shift, wasted, mb, porder);
        } else { atomicOr(err_flag, 2); return; }
        if (br.err) { atomicOr(err_flag, 4); return; }
    }
    // zero padding to the byte boundary, then the CRC-16 the linking step has verified: the frame must end where its link says
    if (br.nbits & 7) br.get(br.nbits & 7);
    br.refill();
    const long long end = br.byte_pos(b) + 2;
    const long long expect = f + 1 < n_frames ? (long long)frames[f + 1].pos : n;
    if (end != expect) atomicOr(err_flag, 8);
}

// one CTA per frame: decorrelate, shift to the output width, interleave
template <class TO>
__global__ void __launch_bounds__(256)
k_fd_output(const int32_t *__restrict__ planar, const FdFrame *__restrict__ frames, long long total, int channels, int out_shift,
            TO *__restrict__ out)
{
    const FdFrame fr = frames[blockIdx.x];
    const int bs = (int)fr.bs;
    if (channels == 2) {
        const int32_t *c0 = planar + fr.out0, *c1 = planar + total + fr.out0;
        for (int i = threadIdx.x; i < bs; i += blockDim.x) {
            int32_t a = c0[i], s = c1[i], l, r;
            if (fr.assign == 8) { l = a; r = a - s; }                      // left / side
            else if (fr.assign == 9) { l = a + s; r = s; }                 // side / right
            else if (fr.assign == 10) { const int32_t m = (int32_t)(((uint32_t)a << 1) | ((uint32_t)s & 1u)); l = (m + s) >> 1; r = (m - s) >> 1; }
            else { l = a; r = s; }
            out[2 * (fr.out0 + i)] = (TO)(l << out_shift);
            out[2 * (fr.out0 + i) + 1] = (TO)(r << out_shift);
        }
    } else {
        for (int ch = 0; ch < channels; ch++) {
            const int32_t *src = planar + (size_t)ch * total + fr.out0;
            for (int i = threadIdx.x; i < bs; i += blockDim.x) out[(size_t)channels * (fr.out0 + i) + ch] = (TO)(src[i] << out_shift);
        }
    }
}

// "fLaC" + metadata blocks (an ID3v2 tag in front is skipped, as libavformat does); fetch(pos, len, dst) reads stream bytes
template <class F>
void fd_parse_stream(F fetch, long long n, FdInfo &I)
{
    uint8_t h[42];
    long long pos = 0;
    if (n >= 10) {
        fetch(0, 10, h);
        if (!memcmp(h, "ID3", 3)) pos = 10 + (((long long)(h[6] & 0x7F) << 21) | ((h[7] & 0x7F) << 14) | ((h[8] & 0x7F) << 7) | (h[9] & 0x7F));
    }
    if (n < pos + 42) JT_THROW(JT_ERR_INVALID_ARG, "not a FLAC stream (%lld bytes)", n);
    fetch(pos, 42, h);
    if (memcmp(h, "fLaC", 4) || (h[4] & 0x7F) != 0 || ((h[5] << 16) | (h[6] << 8) | h[7]) != 34) JT_THROW(JT_ERR_INVALID_ARG, "not a FLAC stream (no fLaC / STREAMINFO)");
    const uint8_t *s = h + 8;
    I.min_bs = (s[0] << 8) | s[1]; I.max_bs = (s[2] << 8) | s[3];
    I.min_fs = ((uint32_t)s[4] << 16) | (s[5] << 8) | s[6]; I.max_fs = ((uint32_t)s[7] << 16) | (s[8] << 8) | s[9];
    I.rate = (s[10] << 12) | (s[11] << 4) | (s[12] >> 4);
    I.channels = ((s[12] >> 1) & 7) + 1; I.bps = (((s[12] & 1) << 4) | (s[13] >> 4)) + 1;
    I.total = ((long long)(s[13] & 15) << 32) | ((long long)s[14] << 24) | (s[15] << 16) | (s[16] << 8) | s[17];
    if (I.rate <= 0 || I.max_bs < 16 || I.min_bs > I.max_bs) JT_THROW(JT_ERR_INVALID_ARG, "FLAC STREAMINFO: rate %d, block sizes %d..%d", I.rate, I.min_bs, I.max_bs);
    bool last = (h[4] & 0x80) != 0;
    pos += 42;
    while (!last) {
        if (pos + 4 > n) JT_THROW(JT_ERR_INVALID_ARG, "FLAC metadata runs past the end of the stream");
        fetch(pos, 4, h);
        last = (h[0] & 0x80) != 0;
        pos += 4 + ((h[1] << 16) | (h[2] << 8) | h[3]);
    }
    if (pos > n) JT_THROW(JT_ERR_INVALID_ARG, "FLAC metadata runs past the end of the stream");
    I.audio_off = pos;
}
}  // namespace

void jt_flac_info_host(const void *bytes, int64_t n_bytes, int *fmt, int *rate, int *channels, int *bps, int64_t *total, int64_t *audio_off)
{
    FdInfo I;
    const uint8_t *b = (const uint8_t *)bytes;
    fd_parse_stream([&](long long pos, int len, uint8_t *dst) { memcpy(dst, b + pos, (size_t)len); }, n_bytes, I);
    if (fmt) *fmt = I.bps <= 16 ? JT_FMT_S16 : JT_FMT_S32;
    if (rate) *rate = I.rate; if (channels) *channels = I.channels; if (bps) *bps = I.bps;
    if (total) *total = I.total; if (audio_off) *audio_off = I.audio_off;
}

// Decodes the stream at d_bytes (device); returns a device buffer of interleaved samples (*fmt says which) and the frame count.
void *jt_flac_decode_device(jt_ctx *c, const uint8_t *d_bytes, int64_t n_bytes, int *fmt, int *rate, int *channels, int64_t *n_frames_out)
{
    FdInfo I;
    fd_parse_stream([&](long long pos, int len, uint8_t *dst) {
        JT_CUDA(cudaMemcpyAsync(dst, d_bytes + pos, (size_t)len, cudaMemcpyDeviceToHost, c->stream));
        JT_CUDA(cudaStreamSynchronize(c->stream));
    }, n_bytes, I);
    if (I.bps > 24 || I.bps < 4) JT_THROW(JT_ERR_UNSUPPORTED, "FLAC with %d bits per sample (4..24)", I.bps);
    if (I.channels > 8) JT_THROW(JT_ERR_UNSUPPORTED, "FLAC with %d channels", I.channels);
    const int out_fmt = I.bps <= 16 ? JT_FMT_S16 : JT_FMT_S32;
    if (fmt) *fmt = out_fmt; if (rate) *rate = I.rate; if (channels) *channels = I.channels;
    const uint8_t *b = d_bytes + I.audio_off;
    const long long n = n_bytes - I.audio_off;
    if (n_frames_out) *n_frames_out = 0;
    if (n <= 0) return jt_dalloc_bytes(c, 64);

    // 1. candidates, sorted by position
    const int chunk = 16384;
    const long long n_chunks = (n + chunk - 1) / chunk;
    const int sgrid = (int)std::min<long long>(n_chunks, (long long)c->num_sms * 8);
    uint32_t *d_counts = jt_dalloc<uint32_t>(c, (size_t)n_chunks);
    { JtLaunch L(c, "flac_decode:scan"); k_fd_scan<false><<<sgrid, 256, 0, c->stream>>>(b, n, I, chunk, n_chunks, d_counts, nullptr, nullptr); }
    uint32_t *h_counts = jt_pinned<uint32_t>(c, (size_t)n_chunks);
    JT_CUDA(cudaMemcpyAsync(h_counts, d_counts, sizeof(uint32_t) * n_chunks, cudaMemcpyDeviceToHost, c->stream));
    JT_CUDA(cudaStreamSynchronize(c->stream));
    unsigned long long *h_off = jt_pinned<unsigned long long>(c, (size_t)n_chunks);
    long long n_cands = 0;
    for (long long i = 0; i < n_chunks; i++) { h_off[i] = (unsigned long long)n_cands; n_cands += h_counts[i]; }
    if (n_cands == 0 || n_cands > 0x7FFFFFF0ll) JT_THROW(JT_ERR_INVALID_ARG, "FLAC stream without a valid frame header");
    unsigned long long *d_off = jt_dalloc<unsigned long long>(c, (size_t)n_chunks);
    JT_CUDA(cudaMemcpyAsync(d_off, h_off, sizeof(unsigned long long) * n_chunks, cudaMemcpyHostToDevice, c->stream));
    FdCand *d_cands = jt_dalloc<FdCand>(c, (size_t)n_cands);
    { JtLaunch L(c, "flac_decode:scan"); k_fd_scan<true><<<sgrid, 256, 0, c->stream>>>(b, n, I, chunk, n_chunks, nullptr, d_off, d_cands); }

    // 2. link each candidate to the candidate its CRC-16 ends at
    // how far a frame can reach: STREAMINFO's maximum frame size when the encoder recorded it, else the size of a frame of
    // VERBATIM subframes, which no encoder exceeds (it would fall back to VERBATIM)
    const long long max_frame_bytes = I.max_fs ? (long long)I.max_fs
                                               : 16 + (long long)I.channels * ((long long)I.max_bs * ((I.bps + 1 + 7) / 8 + 1) + 8) + 2;
    int32_t *d_next = jt_dalloc<int32_t>(c, (size_t)n_cands);
    { JtLaunch L(c, "flac_decode:link"); k_fd_link<<<(int)((n_cands + 127) / 128), 128, 0, c->stream>>>(b, n, d_cands, n_cands, max_frame_bytes, d_next); }
    FdCand *h_cands = jt_pinned<FdCand>(c, (size_t)n_cands);
    int32_t *h_next = jt_pinned<int32_t>(c, (size_t)n_cands);
    JT_CUDA(cudaMemcpyAsync(h_cands, d_cands, sizeof(FdCand) * n_cands, cudaMemcpyDeviceToHost, c->stream));
    JT_CUDA(cudaMemcpyAsync(h_next, d_next, sizeof(int32_t) * n_cands, cudaMemcpyDeviceToHost, c->stream));
    JT_CUDA(cudaStreamSynchronize(c->stream));
    if (h_cands[0].pos != 0) JT_THROW(JT_ERR_INVALID_ARG, "FLAC: no frame at the start of the audio data");
    std::vector<FdFrame> fr;
    long long total = 0;
    for (long long i = 0; i >= 0 && i < n_cands;) {
        FdFrame f; f.pos = h_cands[i].pos; f.out0 = total; f.bs = h_cands[i].bs; f.assign = h_cands[i].assign;
        fr.push_back(f); total += f.bs;
        const long long j = h_next[i];
        if (j < 0) JT_THROW(JT_ERR_INVALID_ARG, "FLAC: frame %zu at byte %llu fails its CRC-16", fr.size() - 1, (unsigned long long)(f.pos + I.audio_off));
        i = j;                                                              // j == n_cands: the frame ends the stream
    }
    if (I.total && total != I.total) JT_THROW(JT_ERR_INVALID_ARG, "FLAC: frames hold %lld samples, STREAMINFO says %lld", total, I.total);
    const long long n_frames = (long long)fr.size();
    FdFrame *h_fr = jt_pinned<FdFrame>(c, (size_t)n_frames);
    memcpy(h_fr, fr.data(), sizeof(FdFrame) * n_frames);
    FdFrame *d_fr = jt_dalloc<FdFrame>(c, (size_t)n_frames);
    JT_CUDA(cudaMemcpyAsync(d_fr, h_fr, sizeof(FdFrame) * n_frames, cudaMemcpyHostToDevice, c->stream));

    // 3. decode
    int32_t *d_planar = jt_dalloc<int32_t>(c, (size_t)total * I.channels);
    int *d_err = jt_dalloc<int>(c, 1);
    JT_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), c->stream));
    { JtLaunch L(c, "flac_decode:frames"); k_fd_frames<<<(int)((n_frames + 63) / 64), 64, 0, c->stream>>>(b, n, d_fr, n_frames, I, total, d_planar, d_err); }
    void *d_out = jt_dalloc_bytes(c, (size_t)total * I.channels * (out_fmt == JT_FMT_S16 ? 2 : 4) + 16);
    {
        JtLaunch L(c, "flac_decode:output");
        if (out_fmt == JT_FMT_S16) k_fd_output<int16_t><<<(int)n_frames, 256, 0, c->stream>>>(d_planar, d_fr, total, I.channels, 16 - I.bps, (int16_t *)d_out);
        else k_fd_output<int32_t><<<(int)n_frames, 256, 0, c->stream>>>(d_planar, d_fr, total, I.channels, 32 - I.bps, (int32_t *)d_out);
    }
    int *h_err = jt_pinned<int>(c, 1);
    JT_CUDA(cudaMemcpyAsync(h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    JT_CUDA(cudaStreamSynchronize(c->stream));
    if (*h_err) JT_THROW(JT_ERR_INVALID_ARG, "FLAC: malformed frame (flags %d: 1 header, 2 subframe syntax, 4 ran out of data, 8 frame length)", *h_err);
    if (n_frames_out) *n_frames_out = total;
    return d_out;
}
