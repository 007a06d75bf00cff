// jt_wav.cu -- input-side container parsing that needs no device: RIFF / WAVE (the format of the reference's
// own fixtures).  FLAC / other containers stay with the caller's decoder (the reference: libavformat, internal/audio/reader.go).
#include "../../include/jtdsp.h"
#include <cstdint>
#include <cstring>

// ---------------------------------------------------------------------------------------------------------------------
// RIFF / WAVE header of an input file (the reference's tests and fixtures are s16 WAVs, testutil_test.go:140-190; the
// reference itself decodes through libavformat, internal/audio/reader.go): where the PCM lies and what it is, so a caller
// can hand the sample bytes of a memory-mapped file straight to jt_analyse / jt_process_audio*.  Host-only.
// ---------------------------------------------------------------------------------------------------------------------
static uint32_t rd_u32(const unsigned char *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint32_t rd_u16(const unsigned char *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

// jt_wav_parse2 also accepts packed 24-bit PCM (reported as JT_FMT_S32: libavcodec's pcm_s24le decoder hands out s32, samples
// shifted up by 8) and says how many bits a stored sample has; jt_wav_parse keeps its "samples usable in place" contract and
// rejects 24 bit.  jt_wav_walk is the walk itself over the first n_readable bytes of a file of n_total bytes (a device-resident
// file image only brings its head to the host).
int jt_wav_walk(const void *bytes, int64_t n_readable, int64_t n_total, int *sample_fmt, int *sample_rate, int *channels, int *bits_per_sample,
                int64_t *data_offset, int64_t *n_frames)
{
    const unsigned char *b = (const unsigned char *)bytes;
    const int64_t n_bytes = n_readable;
    if (b && n_bytes >= 12 && (!memcmp(b, "RF64", 4) || !memcmp(b, "BW64", 4)) && !memcmp(b + 8, "WAVE", 4)) return JT_ERR_UNSUPPORTED;   // 64-bit sizes in a ds64 chunk
    if (!b || n_bytes < 12 || n_total < n_readable || memcmp(b, "RIFF", 4) || memcmp(b + 8, "WAVE", 4)) return JT_ERR_INVALID_ARG;
    int64_t pos = 12; bool have_fmt = false; int tag = 0, ch = 0, rate = 0, bits = 0, align = 0;
    while (pos + 8 <= n_bytes) {
        const unsigned char *c = b + pos; const int64_t len = rd_u32(c + 4);
        if (!memcmp(c, "fmt ", 4)) {
            if (len < 16 || pos + 8 + 16 > n_bytes) return JT_ERR_INVALID_ARG;
            tag = (int)rd_u16(c + 8); ch = (int)rd_u16(c + 10); rate = (int)rd_u32(c + 12); align = (int)rd_u16(c + 20); bits = (int)rd_u16(c + 22);
            if (tag == 0xFFFE && len >= 40 && pos + 8 + 40 <= n_bytes) tag = (int)rd_u16(c + 8 + 24);      // WAVE_FORMAT_EXTENSIBLE: SubFormat
            have_fmt = true;
        } else if (!memcmp(c, "data", 4)) {
            if (!have_fmt || ch <= 0 || rate <= 0) return JT_ERR_INVALID_ARG;
            int fmt;
            if (tag == 1 && bits == 16) fmt = JT_FMT_S16;
            else if (tag == 1 && (bits == 32 || bits == 24)) fmt = JT_FMT_S32;
            else if (tag == 3 && bits == 32) fmt = JT_FMT_FLT;
            else if (tag == 3 && bits == 64) fmt = JT_FMT_DBL;
            else return JT_ERR_UNSUPPORTED;                                                                    // 8 bit, compressed
            const int frame_bytes = ch * (bits / 8);
            if (align && align != frame_bytes) return JT_ERR_INVALID_ARG;
            const int64_t avail = n_total - (pos + 8);
            // streamed files (an encoder writing to a pipe leaves 0 or 0xFFFFFFFF) and truncated ones: libavformat's wav
            // demuxer reads such a data chunk to the end of the file
            const int64_t dl = (len == 0 || len == 0xFFFFFFFFll || len > avail) ? avail : len;
            if (bits_per_sample) *bits_per_sample = bits;
            if (sample_fmt) *sample_fmt = fmt; if (sample_rate) *sample_rate = rate; if (channels) *channels = ch;
            if (data_offset) *data_offset = pos + 8; if (n_frames) *n_frames = dl / frame_bytes;
            return JT_OK;
        }
        pos += 8 + len + (len & 1);
    }
    return n_readable < n_total ? JT_ERR_UNSUPPORTED : JT_ERR_INVALID_ARG;       // no data chunk within the readable head / at all
}

extern "C" int jt_wav_parse(const void *bytes, int64_t n_bytes, int *sample_fmt, int *sample_rate, int *channels,
                            int64_t *data_offset, int64_t *n_frames)
{
    int bits = 0;
    const int rc = jt_wav_walk(bytes, n_bytes, n_bytes, sample_fmt, sample_rate, channels, &bits, data_offset, n_frames);
    return (rc == JT_OK && bits == 24) ? JT_ERR_UNSUPPORTED : rc;
}

// ---------------------------------------------------------------------------------------------------------------------
// STREAMINFO's MD5 of the unencoded audio (RFC 9639 section 8.2: the samples as interleaved little-endian signed integers of
// the stream's sample size; libavcodec's encoder writes it, encoder.go:92-101).  MD5 is a serial chain over 64-byte blocks --
// nothing for a GPU -- so it is a host helper the caller runs next to the encode (0.5 s per hour of audio on one core);
// jt_flac_encode leaves the field zero ("unknown"), which decoders accept.  Host-only, no jt_ctx.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
struct Md5 {
    uint32_t a = 0x67452301u, b = 0xefcdab89u, c = 0x98badcfeu, d = 0x10325476u;
    static uint32_t rol(uint32_t v, int s) { return (v << s) | (v >> (32 - s)); }
    void block(const unsigned char *p)
    {
        static const uint32_t K[64] = {
            0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be,
            0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
            0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a, 0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c,
            0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
            0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1,
            0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
        static const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20,
                                  4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
        uint32_t m[16];
        for (int i = 0; i < 16; i++) m[i] = rd_u32(p + 4 * i);
        uint32_t A = a, B = b, C = c, D = d;
        for (int i = 0; i < 64; i++) {
            uint32_t f; int g;
            if (i < 16) { f = (B & C) | (~B & D); g = i; }
            else if (i < 32) { f = (D & B) | (~D & C); g = (5 * i + 1) & 15; }
            else if (i < 48) { f = B ^ C ^ D; g = (3 * i + 5) & 15; }
            else { f = C ^ (B | ~D); g = (7 * i) & 15; }
            const uint32_t t = D; D = C; C = B; B = B + rol(A + f + K[i] + m[g], S[i]); A = t;
        }
        a += A; b += B; c += C; d += D;
    }
};
}  // namespace

// MD5 of n_bytes at `data` (16 bytes out)
extern "C" int jt_md5(const void *data, int64_t n_bytes, unsigned char out[16])
{
    if ((!data && n_bytes > 0) || n_bytes < 0 || !out) return JT_ERR_INVALID_ARG;
    const unsigned char *p = (const unsigned char *)data;
    Md5 h; int64_t i = 0;
    for (; i + 64 <= n_bytes; i += 64) h.block(p + i);
    unsigned char tail[128]; memset(tail, 0, sizeof(tail));
    const int rem = (int)(n_bytes - i);
    if (rem) memcpy(tail, p + i, (size_t)rem);
    tail[rem] = 0x80;
    const int total = rem + 1 + 8 <= 64 ? 64 : 128;
    const uint64_t bits = (uint64_t)n_bytes * 8;
    for (int k = 0; k < 8; k++) tail[total - 8 + k] = (unsigned char)(bits >> (8 * k));
    h.block(tail); if (total == 128) h.block(tail + 64);
    const uint32_t v[4] = {h.a, h.b, h.c, h.d};
    for (int k = 0; k < 16; k++) out[k] = (unsigned char)(v[k / 4] >> (8 * (k % 4)));
    return JT_OK;
}

// writes the MD5 of the mono s16 samples into the STREAMINFO block of a stream made by jt_flac_encode
extern "C" int jt_flac_set_md5(void *stream, int64_t n_bytes, const int16_t *pcm, int64_t n_samples)
{
    unsigned char *s = (unsigned char *)stream;
    if (!s || n_bytes < 42 || memcmp(s, "fLaC", 4) || (s[4] & 0x7F) != 0 || (!pcm && n_samples > 0) || n_samples < 0) return JT_ERR_INVALID_ARG;
    return jt_md5(pcm, n_samples * 2, s + 26);            // x86-64 / aarch64 hosts are little-endian: the buffer is the s16le byte stream
}
