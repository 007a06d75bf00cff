/*
 * TEST INFRASTRUCTURE ONLY -- reads an audio file the way the reference does (libavformat demuxer + libavcodec decoder,
 * internal/audio/reader.go) with the REAL FFmpeg libraries bundled in this image and writes the decoded interleaved samples
 * raw, so jt_wav_parse (csrc/jt_wav.cu) is pinned on the reader the reference actually uses.
 *
 *   ref_wav_read in.wav out.raw   -> prints "rate=<r> channels=<c> fmt=<AVSampleFormat> samples=<frames>"
 */
#include <stdio.h>
#include <string.h>
#include <libavformat/avformat.h>
#include <libavcodec/avcodec.h>
#include <libavutil/samplefmt.h>

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    AVFormatContext *fc = NULL;
    if (avformat_open_input(&fc, argv[1], NULL, NULL) < 0) return 3;
    if (avformat_find_stream_info(fc, NULL) < 0) return 4;
    int si = av_find_best_stream(fc, AVMEDIA_TYPE_AUDIO, -1, -1, NULL, 0);
    if (si < 0) return 5;
    const AVCodec *codec = avcodec_find_decoder(fc->streams[si]->codecpar->codec_id);
    if (!codec) return 6;
    AVCodecContext *c = avcodec_alloc_context3(codec);
    avcodec_parameters_to_context(c, fc->streams[si]->codecpar);
    if (avcodec_open2(c, codec, NULL) < 0) return 7;
    FILE *o = fopen(argv[2], "wb");
    AVPacket *pkt = av_packet_alloc(); AVFrame *fr = av_frame_alloc();
    long frames = 0; int fmt = -1;
    for (int eof = 0; !eof;) {
        if (av_read_frame(fc, pkt) < 0) { eof = 1; avcodec_send_packet(c, NULL); }
        else { if (pkt->stream_index == si) avcodec_send_packet(c, pkt); av_packet_unref(pkt); }
        while (avcodec_receive_frame(c, fr) == 0) {
            fmt = fr->format;
            if (av_sample_fmt_is_planar(fr->format) && fr->ch_layout.nb_channels > 1) return 8;      /* PCM decoders are packed */
            fwrite(fr->data[0], (size_t)av_get_bytes_per_sample(fr->format) * (size_t)fr->ch_layout.nb_channels, (size_t)fr->nb_samples, o);
            frames += fr->nb_samples;
        }
    }
    fclose(o);
    printf("rate=%d channels=%d fmt=%d samples=%ld\n", c->sample_rate, c->ch_layout.nb_channels, fmt, frames);
    return 0;
}
