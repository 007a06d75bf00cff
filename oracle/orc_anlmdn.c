/*
 * ORACLE (test infrastructure only) -- restatement of FFmpeg's anlmdn (non-local means
 * denoiser) as the reference instantiates it: "anlmdn=s=0.00001:p=0.0060:r=0.0020:m=3"
 * (internal/processor/filters.go:95-100, 811-816).  Follows libavfilter/af_anlmdn.c
 * (config_filter, filter_channel with compute_distance_ssd_c / compute_cache_c, activate
 * consuming hops of H = 2K+1 samples into a window of H + 2(K+S) samples); option->loop
 * mapping corroborated by scripts/anlmdn-matrix-spike.sh:292-320.  float32, sequential.
 * The filter has no latency compensation: output = input delayed by K+S samples.
 * Parity unpinned (see orc.h).
 */
#include "orc.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define WEIGHT_LUT_NBITS 20
#define WEIGHT_LUT_SIZE (1 << WEIGHT_LUT_NBITS)
#define SQR(x) ((x) * (x))

static int64_t rescale_near(int64_t a, int64_t b, int64_t c) { return (a * b + c / 2) / c; }

int orc_anlmdn(const float *x, float *y, int64_t n, int rate, double strength, double patch_s, double research_s, double smooth_m)
{
    const int K = (int)rescale_near(llround(patch_s * 1e6), rate, 1000000);
    const int S = (int)rescale_near(llround(research_s * 1e6), rate, 1000000);
    const int H = K * 2 + 1, N = H + (K + S) * 2, offset = N - H;
    const float a = (float)strength, m = (float)smooth_m;
    const float pdiff_lut_scale = 1.f / m * WEIGHT_LUT_SIZE;
    float *weight_lut = malloc(sizeof(float) * WEIGHT_LUT_SIZE);
    for (int i = 0; i < WEIGHT_LUT_SIZE; i++) { float w = -i / pdiff_lut_scale; weight_lut[i] = expf(w); }
    const float sw = (65536.f / (4 * K + 2)) / sqrtf(a);
    const float smooth = fminf(m, WEIGHT_LUT_SIZE / pdiff_lut_scale);
    float *src = calloc(N, sizeof(float)), *cache = calloc(2 * S, sizeof(float));
    const float *f = src + K;

    for (int64_t pos = 0; pos < n; pos += H) {
        const int nb = (int)(n - pos < H ? n - pos : H);
        memmove(src, &src[H], offset * sizeof(float));
        memcpy(&src[offset], x + pos, nb * sizeof(float));
        memset(&src[offset + nb], 0, (H - nb) * sizeof(float));
        for (int i = S; i < H + S; i++) {
            float P = 0.f, Q = 0.f;
            int v = 0;
            if (i == S) {
                for (int j = i - S; j <= i + S; j++) {
                    if (i == j) continue;
                    float distance = 0.;
                    for (int k = -K; k <= K; k++) distance += SQR(f[i + k] - f[j + k]);
                    cache[v++] = distance;
                }
            } else {
                for (int j = i - S, vv = 0; j < i; j++, vv++)
                    cache[vv] += -SQR(f[i - K - 1] - f[j - K - 1]) + SQR(f[i + K] - f[j + K]);
                for (int j = i + 1, vv = S; j < i + 1 + S; j++, vv++)
                    cache[vv] += -SQR(f[i - K - 1] - f[j - K - 1]) + SQR(f[i + K] - f[j + K]);
            }
            for (int j = 0; j < 2 * S; j++) {
                float distance = cache[j], w;
                unsigned idx;
                if (distance < 0.f) cache[j] = distance = 0.f;
                w = distance * sw;
                if (w >= smooth) continue;
                idx = w * pdiff_lut_scale;
                w = weight_lut[idx];
                P += w * f[i - S + j + (j >= S)];
                Q += w;
            }
            P += f[i];
            Q += 1;
            if (i - S < nb) y[pos + i - S] = P / Q;
        }
    }
    free(weight_lut); free(src); free(cache);
    return K + S;
}
