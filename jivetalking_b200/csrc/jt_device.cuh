// Device-side helpers shared by the kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// ---- audioconvert.c element conversions -------------------------------------------------
template <class TI, class TO> __device__ __forceinline__ TO jt_conv(TI v);
template <> __device__ __forceinline__ float   jt_conv<int16_t, float>(int16_t v)  { return __fmul_rn((float)v, 1.0f / 32768.0f); }
template <> __device__ __forceinline__ double  jt_conv<int16_t, double>(int16_t v) { return (double)v * (1.0 / 32768.0); }
template <> __device__ __forceinline__ double  jt_conv<float, double>(float v)     { return (double)v; }
template <> __device__ __forceinline__ float   jt_conv<double, float>(double v)    { return (float)v; }
template <> __device__ __forceinline__ float   jt_conv<float, float>(float v)      { return v; }
template <> __device__ __forceinline__ double  jt_conv<double, double>(double v)   { return v; }
template <> __device__ __forceinline__ int16_t jt_conv<float, int16_t>(float v)
{   // av_clip_int16(lrintf(v * (1 << 15)))
    float s = __fmul_rn(v, 32768.0f);
    s = fminf(fmaxf(s, -40000.0f), 40000.0f);
    int r = __float2int_rn(s);
    return (int16_t)max(-32768, min(32767, r));
}
template <> __device__ __forceinline__ int16_t jt_conv<double, int16_t>(double v)
{   // av_clip_int16(lrint(v * (1 << 15)))
    double s = __dmul_rn(v, 32768.0);
    s = fmin(fmax(s, -40000.0), 40000.0);
    int r = __double2int_rn(s);
    return (int16_t)max(-32768, min(32767, r));
}

// 32-bit integer samples (24-bit FLAC / WAV decode to s32): audioconvert.c CONV_FUNC table, pinned on the real
// libswresample (tests/golden/swr_golden.npz)
template <> __device__ __forceinline__ float   jt_conv<int32_t, float>(int32_t v)  { return __fmul_rn((float)v, 1.0f / 2147483648.0f); }
template <> __device__ __forceinline__ double  jt_conv<int32_t, double>(int32_t v) { return (double)v * (1.0 / 2147483648.0); }
template <> __device__ __forceinline__ int16_t jt_conv<int32_t, int16_t>(int32_t v) { return (int16_t)(v >> 16); }
template <> __device__ __forceinline__ int32_t jt_conv<int16_t, int32_t>(int16_t v) { return (int32_t)v * 65536; }
template <> __device__ __forceinline__ int32_t jt_conv<int32_t, int32_t>(int32_t v) { return v; }
template <> __device__ __forceinline__ int32_t jt_conv<float, int32_t>(float v)
{   // av_clipl_int32(llrintf(v * (1U << 31)))
    const float s = __fmul_rn(v, 2147483648.0f);
    if (!(s < 2147483648.0f)) return 2147483647;
    if (s < -2147483648.0f) return (int32_t)0x80000000;
    return (int32_t)__float2ll_rn(s);
}
template <> __device__ __forceinline__ int32_t jt_conv<double, int32_t>(double v)
{   // av_clipl_int32(llrint(v * (1U << 31)))
    const double s = __dmul_rn(v, 2147483648.0);
    if (!(s < 2147483647.5)) return 2147483647;
    if (s < -2147483648.0) return (int32_t)0x80000000;
    const long long r = __double2ll_rn(s);
    return (int32_t)(r > 2147483647LL ? 2147483647LL : r);
}

// value normalised to [-1,1] the way the Go side does for raw frame statistics
// (analyser_metrics.go:310-349: s16 / 32768.0, float/double as is)
__device__ __forceinline__ double jt_norm_f64(int16_t v) { return (double)v / 32768.0; }
__device__ __forceinline__ double jt_norm_f64(int32_t v) { return (double)v / 2147483648.0; }
__device__ __forceinline__ double jt_norm_f64(float v)   { return (double)v; }
__device__ __forceinline__ double jt_norm_f64(double v)  { return v; }

// load any supported sample as the f64 a swr/aformat conversion to dbl would give
__device__ __forceinline__ double jt_as_f64(int16_t v) { return (double)v * (1.0 / 32768.0); }
__device__ __forceinline__ double jt_as_f64(int32_t v) { return (double)v * (1.0 / 2147483648.0); }
__device__ __forceinline__ double jt_as_f64(float v)   { return (double)v; }
__device__ __forceinline__ double jt_as_f64(double v)  { return v; }

// ---- reductions ---------------------------------------------------------------------------
__device__ __forceinline__ double jt_warp_sum(double v)
{
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float jt_warp_sum(float v)
{
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double jt_warp_max(double v)
{
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double jt_warp_min(double v)
{
    for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float jt_warp_max(float v)
{
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// atomic max / min on doubles.  Non-negative doubles order like their bit patterns; the
// general versions use a CAS loop.
__device__ __forceinline__ void jt_atomic_max_nonneg(double *addr, double v)
{
    atomicMax((unsigned long long *)addr, (unsigned long long)__double_as_longlong(v));
}
__device__ __forceinline__ void jt_atomic_min_nonneg(double *addr, double v)
{
    atomicMin((unsigned long long *)addr, (unsigned long long)__double_as_longlong(v));
}
__device__ __forceinline__ void jt_atomic_max_f64(double *addr, double v)
{
    unsigned long long *a = (unsigned long long *)addr, old = *a, assumed;
    do {
        assumed = old;
        if (__longlong_as_double((long long)assumed) >= v) break;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    } while (assumed != old);
}
__device__ __forceinline__ void jt_atomic_min_f64(double *addr, double v)
{
    unsigned long long *a = (unsigned long long *)addr, old = *a, assumed;
    do {
        assumed = old;
        if (__longlong_as_double((long long)assumed) <= v) break;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    } while (assumed != old);
}

// floor division / modulo for possibly negative numerators
__device__ __host__ __forceinline__ int64_t jt_floordiv(int64_t a, int64_t b)
{
    int64_t q = a / b, r = a % b;
    return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}
