// Shared-memory FFT used by afftdn and aspectralstats (sm_100a).
#pragma once
#include <cuda_runtime.h>

__device__ __forceinline__ float2 af_cmul(float2 a, float2 w) { return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }

// Stockham autosort FFT of n = 2^k points in shared memory (natural order in and out): radix-4 passes and,
// for odd k, one final radix-2 pass, ping-ponging between `a` and `b`.  tw[j] = exp(-2*pi*i*j/n), j < n.
// Returns the buffer holding the result; ends with a barrier.
template <bool INV>
__device__ __forceinline__ float2 *af_fft(float2 *a, float2 *b, const float2 *__restrict__ tw, int n)
{
    float2 *src = a, *dst = b;
    int ns = 1;
    const int q = n >> 2;
    for (; ns * 4 <= n; ns <<= 2) {
        const int tstep = n / (ns * 4);
        for (int j = threadIdx.x; j < q; j += blockDim.x) {
            const int k = j & (ns - 1), t = k * tstep;
            float2 v0 = src[j], v1 = src[j + q], v2 = src[j + 2 * q], v3 = src[j + 3 * q];
            float2 w1 = tw[t], w2 = tw[2 * t], w3 = tw[3 * t];
            if (INV) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
            v1 = af_cmul(v1, w1); v2 = af_cmul(v2, w2); v3 = af_cmul(v3, w3);
            const float2 a0 = make_float2(v0.x + v2.x, v0.y + v2.y), a1 = make_float2(v0.x - v2.x, v0.y - v2.y);
            const float2 a2 = make_float2(v1.x + v3.x, v1.y + v3.y), d3 = make_float2(v1.x - v3.x, v1.y - v3.y);
            const float2 a3 = INV ? make_float2(-d3.y, d3.x) : make_float2(d3.y, -d3.x);      // (+/-)i * (v1 - v3)
            const int j0 = ((j - k) << 2) + k;
            dst[j0] = make_float2(a0.x + a2.x, a0.y + a2.y);
            dst[j0 + ns] = make_float2(a1.x + a3.x, a1.y + a3.y);
            dst[j0 + 2 * ns] = make_float2(a0.x - a2.x, a0.y - a2.y);
            dst[j0 + 3 * ns] = make_float2(a1.x - a3.x, a1.y - a3.y);
        }
        __syncthreads();
        float2 *t = src; src = dst; dst = t;
    }
    if (ns < n) {                                   // ns == n / 2
        const int h = n >> 1;
        for (int j = threadIdx.x; j < h; j += blockDim.x) {
            float2 w = tw[j]; if (INV) w.y = -w.y;
            const float2 v0 = src[j], v1 = af_cmul(src[j + h], w);
            dst[j] = make_float2(v0.x + v1.x, v0.y + v1.y);
            dst[j + h] = make_float2(v0.x - v1.x, v0.y - v1.y);
        }
        __syncthreads();
        float2 *t = src; src = dst; dst = t;
    }
    return src;
}

