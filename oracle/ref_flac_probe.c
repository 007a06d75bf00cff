/*
 * TEST INFRASTRUCTURE ONLY -- decodes a FLAC stream with the REAL FFmpeg libavcodec (flac parser + flac decoder) and writes the
 * samples as raw s16.  Built by oracle/ref_flac.py against the reference's vendored headers
 * (third_party/ffmpeg-statigo/include, libavcodec major 62) and linked to the FFmpeg 8.0.1 shared libraries bundled with
 * opencv-python-headless in this image; the binary goes to oracle/_ref/ (git-ignored, travels to the GPU box).
 *
 *   ref_flac_decode in.flac out.raw     -> prints "samples=<n> rate=<r> channels=<c> fmt=<f>"; exit 0 on success
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <libavcodec/avcodec.h>
#include <libavutil/mem.h>

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 3;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    uint8_t *buf = av_mallocz((size_t)n + AV_INPUT_BUFFER_PADDING_SIZE);
    if (fread(buf, 1, (size_t)n, f) != (size_t)n) return 3;
    fclose(f);
    if (n < 42 || memcmp(buf, "fLaC", 4)) return 4;
    /* skip metadata blocks, keep STREAMINFO as extradata */
    long pos = 4; const uint8_t *si = NULL;
    for (;;) {
        int last = buf[pos] >> 7, type = buf[pos] & 0x7F; long len = (buf[pos + 1] << 16) | (buf[pos + 2] << 8) | buf[pos + 3];
        if (type == 0) si = buf + pos + 4;
        pos += 4 + len;
        if (last) break;
    }
    if (!si) return 4;
    const AVCodec *codec = avcodec_find_decoder(AV_CODEC_ID_FLAC);
    if (!codec) { fprintf(stderr, "no flac decoder in this libavcodec\n"); return 5; }
    AVCodecParserContext *parser = av_parser_init(AV_CODEC_ID_FLAC);
    if (!parser) { fprintf(stderr, "no flac parser in this libavcodec\n"); return 5; }
    AVCodecContext *ctx = avcodec_alloc_context3(codec);
    ctx->extradata = av_mallocz(34 + AV_INPUT_BUFFER_PADDING_SIZE); memcpy(ctx->extradata, si, 34); ctx->extradata_size = 34;
    if (avcodec_open2(ctx, codec, NULL) < 0) return 6;
    AVPacket *pkt = av_packet_alloc(); AVFrame *fr = av_frame_alloc();
    FILE *o = fopen(argv[2], "wb");
    long total = 0; int rate = 0, ch = 0, fmt = -1, err = 0;
    const uint8_t *p = buf + pos; long left = n - pos; int dry = 0;
    for (;;) {                                   /* the parser holds several frames back: keep calling it with no input until it is dry */
        int used = av_parser_parse2(parser, ctx, &pkt->data, &pkt->size, left > 0 ? p : NULL, left > 0 ? (int)(left > (1 << 20) ? (1 << 20) : left) : 0,
                                    AV_NOPTS_VALUE, AV_NOPTS_VALUE, 0);
        if (used < 0) { err = 7; break; }
        p += used; left -= used;
        if (pkt->size) {
            if (avcodec_send_packet(ctx, pkt) < 0) { err = 8; break; }
            while (avcodec_receive_frame(ctx, fr) == 0) {
                rate = fr->sample_rate; ch = fr->ch_layout.nb_channels; fmt = fr->format;
                if (fr->format == AV_SAMPLE_FMT_S16 || fr->format == AV_SAMPLE_FMT_S16P) fwrite(fr->data[0], 2, (size_t)fr->nb_samples, o);
                else if (fr->format == AV_SAMPLE_FMT_S32 || fr->format == AV_SAMPLE_FMT_S32P) {
                    for (int i = 0; i < fr->nb_samples; i++) { int16_t s = (int16_t)(((int32_t *)fr->data[0])[i] >> 16); fwrite(&s, 2, 1, o); }
                } else err = 9;
                total += fr->nb_samples;
            }
            dry = 0;
        } else if (left <= 0 && ++dry > 64) break;
    }
    avcodec_send_packet(ctx, NULL);
    while (avcodec_receive_frame(ctx, fr) == 0) { fwrite(fr->data[0], 2, (size_t)fr->nb_samples, o); total += fr->nb_samples; }
    fclose(o);
    printf("samples=%ld rate=%d channels=%d fmt=%d\n", total, rate, ch, fmt);
    return err;
}
