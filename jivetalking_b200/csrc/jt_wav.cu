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
