/*
 * TEST INFRASTRUCTURE ONLY -- encodes raw mono s16 with the REAL FFmpeg libavcodec FLAC encoder (LPC subframes, the
 * reference's compression_level 5 and 4096-sample frames: internal/processor/encoder.go:92-101) into a .flac stream, so the
 * oracle's decoder (oracle/orc_flac.c: fixed AND LPC subframes, escape-coded partitions) is pinned on streams it did not
 * write.  Built like ref_flac_probe.c (oracle/ref_flac.py) into oracle/_ref/.
 *
 *   ref_flac_encode in.raw out.flac <rate> <compression_level> [channels [bits: 16 | 24]]   (24: raw s32 input, samples << 8)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <libavcodec/avcodec.h>
#include <libavutil/channel_layout.h>
#include <libavutil/opt.h>

int main(int argc, char **argv)
{
    if (argc < 5) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 3;
    fseek(f, 0, SEEK_END); long nb = ftell(f); fseek(f, 0, SEEK_SET);
    int16_t *x = malloc((size_t)nb + 2);
    if (fread(x, 1, (size_t)nb, f) != (size_t)nb) return 3;
    fclose(f);
    const int nch = argc > 5 ? atoi(argv[5]) : 1, bits = argc > 6 ? atoi(argv[6]) : 16;
    const int bytes_ps = bits > 16 ? 4 : 2;
    const enum AVSampleFormat sf = bits > 16 ? AV_SAMPLE_FMT_S32 : AV_SAMPLE_FMT_S16;
    const long n = nb / bytes_ps / nch;                 /* frames */
    const AVCodec *codec = avcodec_find_encoder(AV_CODEC_ID_FLAC);
    if (!codec) { fprintf(stderr, "no flac encoder in this libavcodec\n"); return 5; }
    AVCodecContext *c = avcodec_alloc_context3(codec);
    c->sample_fmt = sf; c->bits_per_raw_sample = bits; c->sample_rate = atoi(argv[3]); c->compression_level = atoi(argv[4]);
    av_channel_layout_default(&c->ch_layout, nch);
    c->frame_size = 4096;
    if (avcodec_open2(c, codec, NULL) < 0) return 6;
    if (c->extradata_size != 34) return 7;
    FILE *o = fopen(argv[2], "wb");
    const unsigned char head[8] = {'f', 'L', 'a', 'C', 0x80, 0, 0, 34};
    fwrite(head, 1, 8, o); fwrite(c->extradata, 1, 34, o);
    AVFrame *fr = av_frame_alloc(); AVPacket *pkt = av_packet_alloc();
    long frames = 0, bytes = 0;
    for (long pos = 0; pos < n; pos += c->frame_size) {
        const int m = (int)((n - pos) < c->frame_size ? (n - pos) : c->frame_size);
        fr->nb_samples = m; fr->format = sf; fr->sample_rate = c->sample_rate;
        av_channel_layout_default(&fr->ch_layout, nch);
        if (av_frame_get_buffer(fr, 0) < 0) return 8;
        memcpy(fr->data[0], (const char *)x + (size_t)pos * bytes_ps * nch, (size_t)m * bytes_ps * nch);
        if (avcodec_send_frame(c, fr) < 0) return 9;
        av_frame_unref(fr);
        while (avcodec_receive_packet(c, pkt) == 0) { fwrite(pkt->data, 1, (size_t)pkt->size, o); bytes += pkt->size; frames++; av_packet_unref(pkt); }
    }
    avcodec_send_frame(c, NULL);
    while (avcodec_receive_packet(c, pkt) == 0) { fwrite(pkt->data, 1, (size_t)pkt->size, o); bytes += pkt->size; frames++; av_packet_unref(pkt); }
    fclose(o);
    printf("frames=%ld bytes=%ld\n", frames, bytes);
    return 0;
}
