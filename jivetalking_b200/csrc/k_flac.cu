// FLAC (RFC 9639) encoder for the chain's output: mono, 16 bit, fixed block size -- the container the reference writes its
// result in (mono / 44.1 kHz / s16 / 4096-sample frames through libavcodec's encoder: internal/processor/encoder.go:92-101,
// processor.go:379-384; SURVEY 8f-3: at > 10 000 x realtime the CPU encoder, a few hundred x realtime per core, is the next
// bottleneck of a drop-in).
//
// Frames are independent, so ONE WARP encodes one frame: samples staged in shared memory (8 KB), the frame built in a
// shared-memory byte buffer with word-wide atomicOr (lanes own disjoint bit ranges; only boundary words are shared), then
// stored with coalesced word copies into a fixed-stride staging slot.  Decisions (identical, integer for integer, to the
// sequential oracle oracle/orc_flac.c, so the two streams are compared byte for byte):
//   CONSTANT subframe when all samples are equal; else the FIXED predictor order 0..4 with the smallest sum |residual|;
//   Rice partition order 0..6 (block size a multiple of 64), each partition's parameter as libavcodec's flacenc picks it
//   (k = log2((sum - n/2) / n), max 14), the order with the fewest exact bits; VERBATIM when that is not smaller.
// CRC-16 of a frame is computed in parallel: every lane takes 1/32 of the (right-aligned) frame from a zero state, and the
// partial CRCs are chained with the matrix of "append L zero bytes" (CRCs are linear over GF(2)).
// A host scan of the frame sizes gives the offsets; a second kernel packs the frames behind the 42-byte stream header.
// HBM traffic: 2 B/sample read + ~1 B/sample staged and re-read + ~1 B/sample written.
#include "jt_internal.h"
#include "jt_device.cuh"
#include <cstdio>
#include <cstring>
#include <vector>

namespace {
constexpr int FL_WARPS = 4;                              // frames per CTA
constexpr int FL_MAX_BS = 4096;
constexpr int FL_FRAME_CAP = 16 + 2 * FL_MAX_BS + 2 + 2; // header + verbatim subframe + CRC-16, multiple of 4 (8212)
constexpr int FL_FRAME_WORDS = FL_FRAME_CAP / 4;
constexpr int FL_MAX_PORDER = 6;
constexpr int FL_LPC_MAX = 8;                            // LPC orders 1..8 (the reference's compression_level 5), 12-bit coefficients
constexpr int FL_LPC_PREC = 12;
static_assert(FL_FRAME_CAP % 4 == 0, "frame slots are copied as words");
static_assert(FL_FRAME_WORDS >= 64 * 15, "the (u >> k) table borrows the frame buffer");

struct FlWarp {                                          // per-warp shared memory
    int16_t x[FL_MAX_BS + 64];                           // lane chunks skewed by one word each: see FL_X
    uint32_t frame[FL_FRAME_WORDS];
    uint16_t col[16];                                    // CRC-16 "append L zero bytes" operator, one column per state bit
    int32_t lq[FL_LPC_MAX + 1][FL_LPC_MAX];              // quantised LPC coefficients of every order (lane 0 writes, all read)
    int32_t lsh[FL_LPC_MAX + 1];
    int32_t n_lpc;
};

__device__ __forceinline__ uint32_t fl_fold(int32_t r) { return ((uint32_t)r << 1) ^ (uint32_t)(r >> 31); }

// Sample i of the staged block.  A lane walks its own contiguous chunk, so with the plain layout all 32 lanes would hit one
// bank (chunk = 128 int16 = 64 words: ncu showed 605 M conflicts in 636 M shared wavefronts); when the chunk is a power of two
// every chunk is shifted by one more word (2 int16), which puts the lanes on 32 different banks.
#define FL_X(i) (x[(i) + ((((i) >> xsh) << 1) & xpad)])

// for (i in [a, b)) with r = residual at i of the predictor (c1 .. c8 on the previous eight samples, >> sh; a >= order): the
// previous samples ride in registers, one shared-memory read per sample.  The fixed predictors are this with binomial
// coefficients and sh = 0; unused taps are zero.
#define FL_FOR_RES(a, b, BODY)                                                                                         \
    do {                                                                                                               \
        int i_ = (a);                                                                                                  \
        if (i_ < (b)) {                                                                                                \
            int32_t m1_ = i_ >= 1 ? FL_X(i_ - 1) : 0, m2_ = i_ >= 2 ? FL_X(i_ - 2) : 0, m3_ = i_ >= 3 ? FL_X(i_ - 3) : 0,    \
                    m4_ = i_ >= 4 ? FL_X(i_ - 4) : 0, m5_ = i_ >= 5 ? FL_X(i_ - 5) : 0, m6_ = i_ >= 6 ? FL_X(i_ - 6) : 0,    \
                    m7_ = i_ >= 7 ? FL_X(i_ - 7) : 0, m8_ = i_ >= 8 ? FL_X(i_ - 8) : 0;                                \
            for (; i_ < (b); i_++) {                                                                                   \
                const int32_t x0_ = FL_X(i_);                                                                          \
                const int32_t r = x0_ - ((pc1 * m1_ + pc2 * m2_ + pc3 * m3_ + pc4 * m4_ + pc5 * m5_ + pc6 * m6_ + pc7 * m7_ + pc8 * m8_) >> psh); \
                BODY;                                                                                                  \
                m8_ = m7_; m7_ = m6_; m6_ = m5_; m5_ = m4_; m4_ = m3_; m3_ = m2_; m2_ = m1_; m1_ = x0_;                \
            }                                                                                                          \
        }                                                                                                              \
    } while (0)

// Welch window in Q15 by an integer formula (the oracle's welch_q15)
__host__ __device__ inline int32_t fl_welch_q15(int i, int n)
{
    const long long t = 2 * (long long)i - (n - 1), d = (long long)(n + 1) * (n + 1);
    const long long w = 32767 - (t * t * 32767) / d;
    return (int32_t)(w < 0 ? 0 : w);
}

__device__ __forceinline__ unsigned long long fl_est_bits(unsigned long long sum_abs, int n, unsigned long long overhead);

__device__ __forceinline__ int fl_optimal_param(unsigned long long sum, int n)      // flacenc.c find_optimal_param
{
    if (sum <= (unsigned long long)(n >> 1)) return 0;
    unsigned long long q = (sum - (unsigned long long)(n >> 1)) / (unsigned long long)n;
    if (q > 0x7fffffffull) q = 0x7fffffffull;
    const int k = q ? 31 - __clz((unsigned)q) : 0;     // av_log2(0) == 0
    return k > 14 ? 14 : k;
}

__device__ __forceinline__ unsigned long long fl_est_bits(unsigned long long sum_abs, int n, unsigned long long overhead)
{
    const unsigned long long S = 2 * sum_abs;
    const int k = fl_optimal_param(S, n);
    return overhead + (unsigned long long)n * (unsigned long long)(k + 1) + (S >> k);
}

__device__ __forceinline__ unsigned long long fl_warp_sum_u64(unsigned long long v)
{
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// OR `n` bits (n + (bit & 7) <= 32) of v, MSB first, at bit position `bit` of a byte stream held in little-endian words
__device__ __forceinline__ void fl_put(uint32_t *buf, int bit, int n, uint32_t v)
{
    const int B = bit >> 3, o = bit & 7;
    const uint32_t w = v << (32 - o - n);                // big-endian window over bytes B .. B+3
    const uint32_t le = __byte_perm(w, 0, 0x0123);       // the same four bytes as a little-endian word
    const int A = B >> 2, sh = (B & 3) * 8;
    atomicOr(&buf[A], le << sh);
    if (sh) { const uint32_t hi = le >> (32 - sh); if (hi) atomicOr(&buf[A + 1], hi); }
}

__device__ __forceinline__ uint16_t fl_crc16_byte(uint16_t c, uint32_t byte)
{
    c ^= (uint16_t)(byte << 8);
#pragma unroll
    for (int b = 0; b < 8; b++) c = (uint16_t)((c & 0x8000) ? (c << 1) ^ 0x8005 : (c << 1));
    return c;
}

__device__ __forceinline__ int fl_rate_code(int rate)
{
    switch (rate) {
    case 88200: return 1; case 176400: return 2; case 192000: return 3; case 8000: return 4; case 16000: return 5; case 22050: return 6;
    case 24000: return 7; case 32000: return 8; case 44100: return 9; case 48000: return 10; case 96000: return 11;
    }
    return 0;
}

__global__ void __launch_bounds__(FL_WARPS * 32)
k_flac_frames(const int16_t *__restrict__ pcm, int64_t n, int block_size, int rate, int64_t n_frames,
              const int16_t *__restrict__ win_full, const int16_t *__restrict__ win_last,
              uint32_t *__restrict__ stage, uint32_t *__restrict__ sizes)
{
    extern __shared__ __align__(16) unsigned char fl_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    FlWarp &S = reinterpret_cast<FlWarp *>(fl_smem)[warp];
    for (int64_t f = (int64_t)blockIdx.x * FL_WARPS + warp; f < n_frames; f += (int64_t)gridDim.x * FL_WARPS) {
        const int64_t s0 = f * block_size;
        const int bs = (int)min((int64_t)block_size, n - s0);
        // ---- stage samples ----
        const int chunk = (bs + 31) / 32;
        const int xpad = (chunk >= 2 && (chunk & (chunk - 1)) == 0) ? -1 : 0, xsh = xpad ? 31 - __clz(chunk) : 0;
        int16_t *x = S.x;
        for (int i = lane; i < bs; i += 32) FL_X(i) = pcm[s0 + i];
        __syncwarp();
        uint8_t *fb = reinterpret_cast<uint8_t *>(S.frame);
        uint32_t *tab = S.frame;                          // until the frame is written its buffer holds the (u >> k) table: 64 cells x 15
        // frame header length: sync + flags (4), coded frame number, 16-bit block size unless 4096, CRC-8
        const unsigned long long fno = (unsigned long long)f;
        const int ulen = fno < 0x80 ? 1 : fno < 0x800 ? 2 : fno < 0x10000 ? 3 : fno < 0x200000 ? 4 : fno < 0x4000000 ? 5 : fno < 0x80000000ull ? 6 : 7;
        const int bs_code = bs == 4096 ? 12 : 7;
        const int hdr_bytes = 4 + ulen + (bs_code == 7 ? 2 : 0) + 1;
        const int sub0 = hdr_bytes * 8;                   // first bit of the subframe
        // ---- lane ranges: [lo, mid) and [mid, hi) are the lane's two cells when the block divides into 64 ----
        const bool cells = (bs % 64) == 0;
        const int lo = min(bs, lane * chunk), hi = min(bs, lo + chunk);
        const int mid = cells ? lo + chunk / 2 : hi;
        // ---- constant? fixed-order errors ----
        bool equal = true;
        unsigned long long e0 = 0, e1 = 0, e2 = 0, e3 = 0, e4 = 0;
        {
            const int32_t first = FL_X(0);
            int32_t m1 = lo >= 1 ? FL_X(lo - 1) : 0, m2 = lo >= 2 ? FL_X(lo - 2) : 0, m3 = lo >= 3 ? FL_X(lo - 3) : 0, m4 = lo >= 4 ? FL_X(lo - 4) : 0;
            for (int i = lo; i < hi; i++) {
                const int32_t x0 = FL_X(i);
                equal = equal && (x0 == first);
                e0 += (unsigned)abs(x0);
                if (i >= 1) e1 += (unsigned)abs(x0 - m1);
                if (i >= 2) e2 += (unsigned)abs(x0 - 2 * m1 + m2);
                if (i >= 3) e3 += (unsigned)abs(x0 - 3 * m1 + 3 * m2 - m3);
                if (i >= 4) e4 += (unsigned)abs(x0 - 4 * m1 + 6 * m2 - 4 * m3 + m4);
                m4 = m3; m3 = m2; m2 = m1; m1 = x0;
            }
        }
        const bool constant = __all_sync(0xffffffffu, equal);
        // ---- decisions (no frame bytes yet) ----
        int order = 0, best_p = 0, bk0 = 0, bk1 = 0, a0 = lo, a1 = mid, psh = 0;
        int32_t pc1 = 0, pc2 = 0, pc3 = 0, pc4 = 0, pc5 = 0, pc6 = 0, pc7 = 0, pc8 = 0;
        bool verbatim = false, open0 = false, open1 = false, is_lpc = false;
        unsigned lane_bits = 0;
        if (!constant) {
            e0 = fl_warp_sum_u64(e0); e1 = fl_warp_sum_u64(e1); e2 = fl_warp_sum_u64(e2); e3 = fl_warp_sum_u64(e3); e4 = fl_warp_sum_u64(e4);
            // ---- LPC analysis: windowed autocorrelation (integer, exact), Levinson-Durbin + quantisation on lane 0 ----
            int n_lpc = 0;
            if (bs > FL_LPC_MAX + 1) {
                const int16_t *win = bs == block_size ? win_full : win_last;      // nullptr: the ragged last block computes its window
                long long R0 = 0, R1 = 0, R2 = 0, R3 = 0, R4 = 0, R5 = 0, R6 = 0, R7 = 0, R8 = 0;
                {
                    int32_t h1 = 0, h2 = 0, h3 = 0, h4 = 0, h5 = 0, h6 = 0, h7 = 0, h8 = 0;
#define FL_XW(i) ((int32_t)(((int32_t)FL_X(i) * (win ? (int32_t)win[i] : fl_welch_q15((i), bs))) >> 8))
                    if (lo >= 1 && lo < hi) h1 = FL_XW(lo - 1); if (lo >= 2 && lo < hi) h2 = FL_XW(lo - 2); if (lo >= 3 && lo < hi) h3 = FL_XW(lo - 3);
                    if (lo >= 4 && lo < hi) h4 = FL_XW(lo - 4); if (lo >= 5 && lo < hi) h5 = FL_XW(lo - 5); if (lo >= 6 && lo < hi) h6 = FL_XW(lo - 6);
                    if (lo >= 7 && lo < hi) h7 = FL_XW(lo - 7); if (lo >= 8 && lo < hi) h8 = FL_XW(lo - 8);
                    for (int i = lo; i < hi; i++) {
                        const int32_t v = FL_XW(i);
                        R0 += (long long)v * v; R1 += (long long)v * h1; R2 += (long long)v * h2; R3 += (long long)v * h3; R4 += (long long)v * h4;
                        R5 += (long long)v * h5; R6 += (long long)v * h6; R7 += (long long)v * h7; R8 += (long long)v * h8;
                        h8 = h7; h7 = h6; h6 = h5; h5 = h4; h4 = h3; h3 = h2; h2 = h1; h1 = v;
                    }
#undef FL_XW
                }
                long long R[FL_LPC_MAX + 1] = {R0, R1, R2, R3, R4, R5, R6, R7, R8};
#pragma unroll
                for (int k = 0; k <= FL_LPC_MAX; k++) R[k] = (long long)fl_warp_sum_u64((unsigned long long)R[k]);
                if (lane == 0) {
                    int usable = 0;
                    if (R[0] > 0) {
                        double err = (double)R[0], lpc[FL_LPC_MAX], tmp[FL_LPC_MAX];
                        for (int i = 0; i < FL_LPC_MAX; i++) {
                            double acc = (double)R[i + 1];
                            for (int j = 0; j < i; j++) acc = __dsub_rn(acc, __dmul_rn(lpc[j], (double)R[i - j]));
                            const double k = __ddiv_rn(acc, err);
                            for (int j = 0; j < i; j++) tmp[j] = __dsub_rn(lpc[j], __dmul_rn(k, lpc[i - 1 - j]));
                            for (int j = 0; j < i; j++) lpc[j] = tmp[j];
                            lpc[i] = k;
                            err = __dmul_rn(err, __dsub_rn(1.0, __dmul_rn(k, k)));
                            if (!(err > 0.0)) break;
                            const int o = i + 1, qmax = (1 << (FL_LPC_PREC - 1)) - 1;
                            double cmax = 0.0;
                            for (int j = 0; j < o; j++) { const double a = lpc[j] < 0 ? -lpc[j] : lpc[j]; if (a > cmax) cmax = a; }
                            int sh = 14;
                            while (sh > 0 && __dmul_rn(cmax, (double)(1 << sh)) > (double)qmax) sh--;
                            double e = 0.0;
                            for (int j = 0; j < o; j++) {
                                e = __dadd_rn(e, __dmul_rn(lpc[j], (double)(1 << sh)));
                                long long v = __double2ll_rn(e);
                                if (v > qmax) v = qmax;
                                if (v < -qmax) v = -qmax;
                                S.lq[o][j] = (int32_t)v;
                                e = __dsub_rn(e, (double)v);
                            }
                            for (int j = o; j < FL_LPC_MAX; j++) S.lq[o][j] = 0;
                            S.lsh[o] = sh;
                            usable = o;
                        }
                    }
                    S.n_lpc = usable;
                }
                __syncwarp();
                n_lpc = S.n_lpc;
            }
            // ---- candidates: FIXED 0..4, then LPC 1..n_lpc, scored by the Rice-cost estimate of sum |residual| + header; first minimum wins ----
            unsigned long long best_est = ~0ull;
            {
                const unsigned long long ef[5] = {e0, e1, e2, e3, e4};
                const int32_t fc[5][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};
                for (int o = 0; o <= 4 && o < bs; o++) {
                    const unsigned long long b = fl_est_bits(ef[o], bs - o, 16ull * (unsigned)o);
                    if (b < best_est) { best_est = b; order = o; pc1 = fc[o][0]; pc2 = fc[o][1]; pc3 = fc[o][2]; pc4 = fc[o][3]; }
                }
            }
            for (int o = 1; o <= n_lpc; o++) {
                const int32_t c1 = S.lq[o][0], c2 = S.lq[o][1], c3 = S.lq[o][2], c4 = S.lq[o][3], c5 = S.lq[o][4], c6 = S.lq[o][5], c7 = S.lq[o][6], c8 = S.lq[o][7];
                const int sh = S.lsh[o];
                unsigned long long e = 0;
                {
                    int i = max(lo, o);
                    if (i < hi) {
                        int32_t m1 = FL_X(i - 1), m2 = i >= 2 ? FL_X(i - 2) : 0, m3 = i >= 3 ? FL_X(i - 3) : 0, m4 = i >= 4 ? FL_X(i - 4) : 0,
                                m5 = i >= 5 ? FL_X(i - 5) : 0, m6 = i >= 6 ? FL_X(i - 6) : 0, m7 = i >= 7 ? FL_X(i - 7) : 0, m8 = i >= 8 ? FL_X(i - 8) : 0;
                        for (; i < hi; i++) {
                            const int32_t x0 = FL_X(i);
                            e += (unsigned)abs(x0 - ((c1 * m1 + c2 * m2 + c3 * m3 + c4 * m4 + c5 * m5 + c6 * m6 + c7 * m7 + c8 * m8) >> sh));
                            m8 = m7; m7 = m6; m6 = m5; m5 = m4; m4 = m3; m3 = m2; m2 = m1; m1 = x0;
                        }
                    }
                }
                e = fl_warp_sum_u64(e);
                const unsigned long long b = fl_est_bits(e, bs - o, 16ull * (unsigned)o + 9ull + (unsigned long long)FL_LPC_PREC * (unsigned)o);
                if (b < best_est) { best_est = b; order = o; is_lpc = true; psh = sh; pc1 = c1; pc2 = c2; pc3 = c3; pc4 = c4; pc5 = c5; pc6 = c6; pc7 = c7; pc8 = c8; }
            }
            a0 = max(lo, order); a1 = max(mid, order);                    // residual samples of the two cells: [a0, mid), [a1, hi)
            // one pass: per cell, the sum of (u >> k) for every Rice parameter k (k = 0 is the plain sum that picks the parameter)
            {
                uint32_t t0[15], t1[15];
#pragma unroll
                for (int k = 0; k < 15; k++) { t0[k] = 0; t1[k] = 0; }
                FL_FOR_RES(a0, mid, { const uint32_t u = fl_fold(r);
                    _Pragma("unroll") for (int k = 0; k < 15; k++) t0[k] += u >> k; });
                FL_FOR_RES(a1, hi, { const uint32_t u = fl_fold(r);
                    _Pragma("unroll") for (int k = 0; k < 15; k++) t1[k] += u >> k; });
#pragma unroll
                for (int k = 0; k < 15; k++) { tab[(2 * lane) * 15 + k] = t0[k]; tab[(2 * lane + 1) * 15 + k] = t1[k]; }
            }
            __syncwarp();
            const unsigned cnt0 = (unsigned)max(0, mid - a0), cnt1 = (unsigned)max(0, hi - a1);
            int pmax = 0;
            if (cells) { pmax = FL_MAX_PORDER; while (pmax > 0 && (bs >> pmax) <= order) pmax--; }
            unsigned long long best_bits = ~0ull; unsigned best_lane = 0;
            for (int p = 0; p <= pmax; p++) {
                const int per = 64 >> p;                                  // cells per partition
                const int psz = bs >> p;
                int k0, k1;
                {
                    const int j = (2 * lane) / per;
                    unsigned long long s = 0;
                    for (int g = j * per; g < (j + 1) * per; g++) s += tab[g * 15];
                    k0 = fl_optimal_param(s, psz - (j == 0 ? order : 0));
                }
                {
                    const int j = (2 * lane + 1) / per;
                    unsigned long long s = 0;
                    for (int g = j * per; g < (j + 1) * per; g++) s += tab[g * 15];
                    k1 = fl_optimal_param(s, psz - (j == 0 ? order : 0));
                }
                const unsigned mine = cnt0 * (unsigned)(k0 + 1) + tab[(2 * lane) * 15 + k0] + cnt1 * (unsigned)(k1 + 1) + tab[(2 * lane + 1) * 15 + k1];
                const unsigned long long bits = fl_warp_sum_u64((unsigned long long)mine) + 4ull * (unsigned long long)(1 << p);
                if (bits < best_bits) { best_bits = bits; best_p = p; bk0 = k0; bk1 = k1; best_lane = mine; }
            }
            const unsigned long long coded_bits = 8ull + 16ull * (unsigned)order + (is_lpc ? 9ull + (unsigned long long)FL_LPC_PREC * (unsigned)order : 0ull) + 6ull + best_bits,
                                     verbatim_bits = 8ull + 16ull * (unsigned)bs;
            verbatim = coded_bits >= verbatim_bits;
            const int per = 64 >> best_p;
            // a cell opens a partition (4-bit parameter in front of it) when it is the partition's first cell
            open0 = cells ? ((2 * lane) % per) == 0 : lane == 0;
            open1 = cells ? ((2 * lane + 1) % per) == 0 : false;
            lane_bits = best_lane + (open0 ? 4u : 0u) + (open1 ? 4u : 0u);
        }
        __syncwarp();                                     // the table has been read: the buffer becomes the frame
        for (int i = lane; i < FL_FRAME_WORDS; i += 32) S.frame[i] = 0;
        __syncwarp();
        // ---- frame header (lane 0) ----
        if (lane == 0) {
            fb[0] = 0xFF; fb[1] = 0xF8; fb[2] = (uint8_t)((bs_code << 4) | fl_rate_code(rate)); fb[3] = 0x08;    // mono, 16 bit
            int p = 4;
            if (ulen == 1) fb[p++] = (uint8_t)fno;
            else {
                const uint8_t lead[8] = {0, 0, 0xC0, 0xE0, 0xF0, 0xF8, 0xFC, 0xFE};
                unsigned long long t = fno;
                for (int i = ulen - 1; i > 0; i--) { fb[p + i] = (uint8_t)(0x80 | (t & 0x3F)); t >>= 6; }
                fb[p] = (uint8_t)(lead[ulen] | t);
                p += ulen;
            }
            if (bs_code == 7) { fb[p++] = (uint8_t)((bs - 1) >> 8); fb[p++] = (uint8_t)(bs - 1); }
            uint8_t c = 0;
            for (int i = 0; i < p; i++) { c ^= fb[i]; for (int b = 0; b < 8; b++) c = (uint8_t)((c & 0x80) ? (c << 1) ^ 0x07 : (c << 1)); }
            fb[p] = c;
        }
        __syncwarp();
        // ---- subframe ----
        int total_bits;                                   // bits of header + subframe before padding
        if (constant) {
            if (lane == 0) { fl_put(S.frame, sub0, 8, 0x00); fl_put(S.frame, sub0 + 8, 16, (uint16_t)FL_X(0)); }
            total_bits = sub0 + 24;
        } else if (verbatim) {
            if (lane == 0) fl_put(S.frame, sub0, 8, 0x02);
            for (int i = lo; i < hi; i++) fl_put(S.frame, sub0 + 8 + 16 * i, 16, (uint16_t)FL_X(i));
            total_bits = sub0 + 8 + 16 * bs;
        } else {
            unsigned incl = lane_bits;
            for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            const int lpc_bits = is_lpc ? 9 + FL_LPC_PREC * order : 0;
            const int base = sub0 + 8 + 16 * order + lpc_bits + 6;
            int pos = base + (int)(incl - lane_bits);
            total_bits = base + (int)__shfl_sync(0xffffffffu, incl, 31);
            if (lane == 0) {
                fl_put(S.frame, sub0, 8, (uint32_t)((is_lpc ? (0x20 | (order - 1)) : (0x08 | order)) << 1));
                for (int i = 0; i < order; i++) fl_put(S.frame, sub0 + 8 + 16 * i, 16, (uint16_t)FL_X(i));
                if (is_lpc) {
                    int q = sub0 + 8 + 16 * order;
                    fl_put(S.frame, q, 4, (uint32_t)(FL_LPC_PREC - 1)); fl_put(S.frame, q + 4, 5, (uint32_t)psh); q += 9;
                    const int32_t cc[8] = {pc1, pc2, pc3, pc4, pc5, pc6, pc7, pc8};
                    for (int j = 0; j < order; j++) fl_put(S.frame, q + FL_LPC_PREC * j, FL_LPC_PREC, (uint32_t)cc[j] & ((1u << FL_LPC_PREC) - 1));
                }
                fl_put(S.frame, sub0 + 8 + 16 * order + lpc_bits, 6, (uint32_t)best_p);       // method 00 + partition order
            }
            if (open0) { fl_put(S.frame, pos, 4, (uint32_t)bk0); pos += 4; }
            FL_FOR_RES(a0, mid, {
                const uint32_t u = fl_fold(r);
                pos += (int)(u >> bk0);
                fl_put(S.frame, pos, bk0 + 1, (1u << bk0) | (u & ((1u << bk0) - 1)));          // the unary stop bit and the k low bits
                pos += bk0 + 1;
            });
            if (open1) { fl_put(S.frame, pos, 4, (uint32_t)bk1); pos += 4; }
            FL_FOR_RES(a1, hi, {
                const uint32_t u = fl_fold(r);
                pos += (int)(u >> bk1);
                fl_put(S.frame, pos, bk1 + 1, (1u << bk1) | (u & ((1u << bk1) - 1)));
                pos += bk1 + 1;
            });
        }
        __syncwarp();
        // ---- CRC-16 over the padded frame: lanes take 1/32 each of the right-aligned bytes, then chain ----
        const int nb = (total_bits + 7) >> 3;
        const int L = (nb + 31) / 32, Z = 32 * L - nb;
        uint16_t part = 0;
        for (int v = lane * L; v < (lane + 1) * L; v++) { const int r = v - Z; if (r >= 0) part = fl_crc16_byte(part, fb[r]); }
        if (lane < 16) { uint16_t c = (uint16_t)(1u << lane); for (int i = 0; i < L; i++) c = fl_crc16_byte(c, 0); S.col[lane] = c; }
        __syncwarp();
        uint16_t crc = 0;
        for (int l = 0; l < 32; l++) {
            const uint16_t pl = (uint16_t)__shfl_sync(0xffffffffu, (unsigned)part, l);
            uint16_t m = 0;
            for (int b = 0; b < 16; b++) if ((crc >> b) & 1) m ^= S.col[b];
            crc = m ^ pl;
        }
        if (lane == 0) { fb[nb] = (uint8_t)(crc >> 8); fb[nb + 1] = (uint8_t)crc; sizes[f] = (uint32_t)(nb + 2); }
        __syncwarp();
        uint32_t *dst = stage + (size_t)f * FL_FRAME_WORDS;
        const int words = (nb + 2 + 3) >> 2;
        for (int i = lane; i < words; i += 32) dst[i] = S.frame[i];
        __syncwarp();
    }
}

// pack the frames behind the stream header: one warp per frame, bytes (destinations are not word aligned)
__global__ void k_flac_pack(const uint8_t *__restrict__ stage, const uint32_t *__restrict__ sizes, const unsigned long long *__restrict__ offsets,
                            int64_t n_frames, uint8_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t wpb = blockDim.x >> 5;
    for (int64_t f = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); f < n_frames; f += (int64_t)gridDim.x * wpb) {
        const uint8_t *src = stage + (size_t)f * FL_FRAME_CAP;
        uint8_t *dst = out + offsets[f];
        const int sz = (int)sizes[f];
        for (int i = lane; i < sz; i += 32) dst[i] = src[i];
    }
}
}  // namespace

int64_t jt_flac_bound(int64_t n, int block_size)
{
    const int64_t frames = (n + block_size - 1) / block_size;
    return 42 + frames * (16 + 2 * (int64_t)block_size + 2) + 64;
}

// Encodes n mono s16 samples at d_pcm; returns a device buffer with the stream and its length.
void *jt_flac_encode_device(jt_ctx *c, const int16_t *d_pcm, int64_t n, int rate, int block_size, int64_t *n_bytes)
{
    if (block_size < 16 || block_size > FL_MAX_BS) JT_THROW(JT_ERR_UNSUPPORTED, "FLAC block size %d (16..%d)", block_size, FL_MAX_BS);
    if (rate <= 0 || rate >= (1 << 20)) JT_THROW(JT_ERR_INVALID_ARG, "FLAC sample rate %d", rate);
    const int64_t frames = (n + block_size - 1) / block_size;
    uint8_t hdr[42]; memset(hdr, 0, sizeof(hdr));
    uint32_t min_fs = 0, max_fs = 0;
    uint8_t *d_out = nullptr;
    int64_t total = 42;
    if (frames > 0) {
        uint32_t *d_stage = (uint32_t *)jt_dalloc_bytes(c, (size_t)frames * FL_FRAME_CAP);
        uint32_t *d_sizes = jt_dalloc<uint32_t>(c, (size_t)frames);
        const size_t smem = sizeof(FlWarp) * FL_WARPS;
        jt_smem_optin((const void *)k_flac_frames, smem);
        const int grid = (int)std::min<int64_t>((frames + FL_WARPS - 1) / FL_WARPS, (int64_t)c->num_sms * 12);
        // Welch window (Q15, integer formula) of the nominal block; a ragged last block computes its own in the kernel
        std::vector<int16_t> wf((size_t)block_size);
        for (int i = 0; i < block_size; i++) wf[i] = (int16_t)fl_welch_q15(i, block_size);
        const int16_t *d_wf = jt_dev_table(c, "flac_welch", wf), *d_wl = nullptr;
        {
            JtLaunch L(c, "flac:frames");
            k_flac_frames<<<grid, FL_WARPS * 32, smem, c->stream>>>(d_pcm, n, block_size, rate, frames, d_wf, d_wl, d_stage, d_sizes);
        }
        uint32_t *h_sizes = jt_pinned<uint32_t>(c, (size_t)frames);
        JT_CUDA(cudaMemcpyAsync(h_sizes, d_sizes, sizeof(uint32_t) * frames, cudaMemcpyDeviceToHost, c->stream));
        JT_CUDA(cudaStreamSynchronize(c->stream));
        unsigned long long *h_off = jt_pinned<unsigned long long>(c, (size_t)frames);
        min_fs = 0xFFFFFF;
        for (int64_t f = 0; f < frames; f++) {
            h_off[f] = (unsigned long long)total; total += h_sizes[f];
            min_fs = std::min(min_fs, h_sizes[f]); max_fs = std::max(max_fs, h_sizes[f]);
        }
        unsigned long long *d_off = jt_dalloc<unsigned long long>(c, (size_t)frames);
        JT_CUDA(cudaMemcpyAsync(d_off, h_off, sizeof(unsigned long long) * frames, cudaMemcpyHostToDevice, c->stream));
        d_out = (uint8_t *)jt_dalloc_bytes(c, (size_t)total + 16);
        {
            JtLaunch L(c, "flac:pack");
            const int pgrid = (int)std::min<int64_t>((frames + 7) / 8, (int64_t)c->num_sms * 8);
            k_flac_pack<<<pgrid, 256, 0, c->stream>>>((const uint8_t *)d_stage, d_sizes, d_off, frames, d_out);
        }
    } else d_out = (uint8_t *)jt_dalloc_bytes(c, 64);
    // "fLaC" + STREAMINFO (last metadata block): block sizes, frame sizes, rate / channels / bits, sample count, MD5 unknown
    memcpy(hdr, "fLaC", 4);
    hdr[4] = 0x80; hdr[7] = 34;
    hdr[8] = (uint8_t)(block_size >> 8); hdr[9] = (uint8_t)block_size; hdr[10] = hdr[8]; hdr[11] = hdr[9];
    hdr[12] = (uint8_t)(min_fs >> 16); hdr[13] = (uint8_t)(min_fs >> 8); hdr[14] = (uint8_t)min_fs;
    hdr[15] = (uint8_t)(max_fs >> 16); hdr[16] = (uint8_t)(max_fs >> 8); hdr[17] = (uint8_t)max_fs;
    // 20 bits rate | 3 bits channels-1 (0) | 5 bits bps-1 (15) | 36 bits total samples
    const unsigned long long v = ((unsigned long long)rate << 44) | (0ull << 41) | (15ull << 36) | ((unsigned long long)n & 0xFFFFFFFFFull);
    for (int i = 0; i < 8; i++) hdr[18 + i] = (uint8_t)(v >> (56 - 8 * i));
    uint8_t *h_hdr = jt_pinned<uint8_t>(c, 64);
    memcpy(h_hdr, hdr, 42);
    JT_CUDA(cudaMemcpyAsync(d_out, h_hdr, 42, cudaMemcpyHostToDevice, c->stream));
    if (n_bytes) *n_bytes = total;
    return d_out;
}
