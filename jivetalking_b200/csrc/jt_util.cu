// Context management, device memory, sample-format conversion, downmix, raw frame
// statistics, volume / gain / pad.  sm_100a.
#include "jt_internal.h"
#include "jt_device.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <iterator>

// ---------------------------------------------------------------------------------------
// context / memory / launch bookkeeping
// ---------------------------------------------------------------------------------------
static void *arena_take(jt_ctx *c, size_t bytes)
{
    for (auto &sl : c->slabs) {
        for (auto it = sl.free_list.begin(); it != sl.free_list.end(); ++it) {
            if (it->second < bytes) continue;
            const size_t off = it->first, len = it->second;
            sl.free_list.erase(it);
            if (len > bytes) sl.free_list[off + bytes] = len - bytes;
            return sl.base + off;
        }
    }
    return nullptr;
}

static void arena_give(jt_ctx *c, void *p, size_t bytes)
{
    for (auto &sl : c->slabs) {
        if ((char *)p < sl.base || (char *)p >= sl.base + sl.size) continue;
        size_t off = (char *)p - sl.base, len = bytes;
        auto nx = sl.free_list.lower_bound(off);
        if (nx != sl.free_list.end() && off + len == nx->first) { len += nx->second; nx = sl.free_list.erase(nx); }
        if (nx != sl.free_list.begin()) { auto pv = std::prev(nx); if (pv->first + pv->second == off) { off = pv->first; len += pv->second; sl.free_list.erase(pv); } }
        sl.free_list[off] = len;
        return;
    }
}

void *jt_dalloc_bytes(jt_ctx *c, size_t bytes)
{
    bytes = (std::max<size_t>(bytes, 1) + 511) & ~size_t(511);
    void *p = arena_take(c, bytes);
    if (!p) {
        // grow: a new slab at least as large as everything allocated so far (the arena doubles), so a stream of a
        // given length settles after its first step; jt_release_all then merges the slabs into one
        size_t total = 0; for (auto &sl : c->slabs) total += sl.size;
        const size_t want = std::max<size_t>(std::max<size_t>(bytes, total), (size_t)256 << 20);
        jt_ctx::Slab sl;
        cudaError_t e = cudaMalloc((void **)&sl.base, want);
        sl.size = want;
        if (e != cudaSuccess && want > bytes) { cudaGetLastError(); e = cudaMalloc((void **)&sl.base, bytes); sl.size = bytes; }
        if (e != cudaSuccess) { cudaGetLastError(); JT_THROW(JT_ERR_NOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); }
        sl.free_list[0] = sl.size;
        c->slabs.push_back(sl);
        p = arena_take(c, bytes);
    }
    c->live[p] = bytes; c->live_bytes += bytes; c->peak_bytes = std::max(c->peak_bytes, c->live_bytes);
    c->allocs.push_back(p);
    return p;
}

static void arena_release(jt_ctx *c, void *p)
{
    auto it = c->live.find(p);
    if (it == c->live.end()) return;
    c->live_bytes -= it->second;
    arena_give(c, p, it->second);
    c->live.erase(it);
}

void jt_release_all(jt_ctx *c)
{
    for (void *p : c->allocs) arena_release(c, p);
    c->allocs.clear();
    c->pin_block = 0; c->pin_used = 0; c->events_used = 0;       // the API call that owned them is over
    // several slabs (the arena grew during this call): replace them by one that holds the call's peak with headroom
    if (c->slabs.size() > 1 && c->live.empty()) {
        cudaStreamSynchronize(c->stream);
        size_t total = 0; for (auto &sl : c->slabs) { total += sl.size; cudaFree(sl.base); }
        c->slabs.clear();
        jt_ctx::Slab sl;
        size_t want = std::max(total, c->peak_bytes + c->peak_bytes / 4);
        if (cudaMalloc((void **)&sl.base, want) != cudaSuccess) {
            cudaGetLastError(); want = c->peak_bytes + ((size_t)64 << 20);
            if (cudaMalloc((void **)&sl.base, want) != cudaSuccess) { cudaGetLastError(); return; }
        }
        sl.size = want; sl.free_list[0] = want;
        c->slabs.push_back(sl);
    }
}

const void *jt_dev_table(jt_ctx *c, const char *tag, const void *host, size_t bytes)
{
    uint64_t h = 1469598103934665603ull;                           // FNV-1a over the content
    const unsigned char *b = (const unsigned char *)host;
    for (size_t i = 0; i < bytes; i++) { h ^= b[i]; h *= 1099511628211ull; }
    char key[160];
    snprintf(key, sizeof(key), "%s:%zu:%016llx", tag, bytes, (unsigned long long)h);
    auto it = c->dev_tables.find(key);
    if (it != c->dev_tables.end()) return it->second;
    void *d = nullptr;
    if (cudaMalloc(&d, bytes ? bytes : 16) != cudaSuccess) JT_THROW(JT_ERR_NOMEM, "cudaMalloc(%zu) for table %s", bytes, tag);
    // On the context's own stream and waited for: the stream is non-blocking, so a cudaMemcpy on the legacy stream is NOT
    // ordered before the kernels that read the table (from pageable memory it may return while the DMA is still queued --
    // harmless on an idle device, a stale table when other contexts keep the copy engine busy).  First use only: cached.
    if (bytes && (cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
                  cudaStreamSynchronize(c->stream) != cudaSuccess)) { cudaFree(d); JT_THROW(JT_ERR_CUDA, "table upload %s", tag); }
    c->dev_tables[key] = d;
    return d;
}

void *jt_pinned_bytes(jt_ctx *c, size_t bytes)
{
    bytes = (bytes + 255) & ~size_t(255);
    while (c->pin_block < c->pin_blocks.size()) {
        auto &blk = c->pin_blocks[c->pin_block];
        if (c->pin_used + bytes <= blk.second) { void *p = blk.first + c->pin_used; c->pin_used += bytes; return p; }
        c->pin_block++; c->pin_used = 0;
    }
    const size_t cap = std::max<size_t>(bytes, 8u << 20);
    char *p = nullptr;
    if (cudaHostAlloc((void **)&p, cap, cudaHostAllocDefault) != cudaSuccess) JT_THROW(JT_ERR_NOMEM, "cudaHostAlloc(%zu)", cap);
    c->pin_blocks.push_back({p, cap});
    c->pin_block = c->pin_blocks.size() - 1; c->pin_used = bytes;
    return p;
}

__global__ void k_copy_small(unsigned char *__restrict__ dst, const unsigned char *__restrict__ src, size_t bytes, int wide)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x, t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (wide) {
        const size_t n16 = bytes >> 4;
        for (size_t i = t; i < n16; i += stride) ((uint4 *)dst)[i] = ((const uint4 *)src)[i];
        for (size_t i = (n16 << 4) + t; i < bytes; i += stride) dst[i] = src[i];
    } else for (size_t i = t; i < bytes; i += stride) dst[i] = src[i];
}
void jt_copy_small(jt_ctx *c, void *dst, const void *src, size_t bytes)
{
    if (!bytes) return;
    const int wide = ((((uintptr_t)dst) | ((uintptr_t)src)) & 15) == 0;
    const size_t items = wide ? (bytes >> 4) + 16 : bytes;
    c->launches++;
    k_copy_small<<<(int)std::min<size_t>((items + 255) / 256, (size_t)c->num_sms * 2), 256, 0, c->stream>>>((unsigned char *)dst, (const unsigned char *)src, bytes, wide);
}

cudaEvent_t jt_record_event(jt_ctx *c)
{
    if (c->events_used == c->event_pool.size()) {
        cudaEvent_t e;
        JT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->event_pool.push_back(e);
    }
    cudaEvent_t e = c->event_pool[c->events_used++];
    JT_CUDA(cudaEventRecord(e, c->stream));
    return e;
}

void jt_release_since(jt_ctx *c, size_t mark, const void *keep)
{
    // `keep` may point INTO an allocation (a view of a buffer, a fused kernel's offset output): the allocation that
    // contains it survives; a non-null keep that no allocation made since `mark` contains is a caller bug
    std::vector<void *> kept(c->allocs.begin(), c->allocs.begin() + std::min(mark, c->allocs.size()));
    bool found = keep == nullptr;
    for (size_t i = mark; i < c->allocs.size(); i++) {
        void *p = c->allocs[i];
        auto it = c->live.find(p);
        const bool holds = keep && it != c->live.end() && (const char *)keep >= (const char *)p && (const char *)keep < (const char *)p + it->second;
        if (holds) { kept.push_back(p); found = true; }
        else arena_release(c, p);
    }
    c->allocs.swap(kept);
    if (!found) {
        for (void *p : c->allocs) {     // kept from before the mark: fine
            auto it = c->live.find(p);
            if (it != c->live.end() && (const char *)keep >= (const char *)p && (const char *)keep < (const char *)p + it->second) { found = true; break; }
        }
        if (!found) JT_THROW(JT_ERR_INVALID_ARG, "internal: jt_release_since would free the buffer that must survive");
    }
}

void jt_release_range(jt_ctx *c, size_t from, size_t to)
{
    to = std::min(to, c->allocs.size());
    if (from >= to) return;
    for (size_t i = from; i < to; i++) arena_release(c, c->allocs[i]);
    c->allocs.erase(c->allocs.begin() + (long)from, c->allocs.begin() + (long)to);
}

void jt_trace(jt_ctx *c, const char *label)
{
    if (c->trace < 0) c->trace = getenv("JT_TRACE") ? 1 : 0;
    if (!c->trace) return;
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    jt_ctx::TracePt p{label, ts.tv_sec + 1e-9 * ts.tv_nsec, nullptr};
    cudaEventCreate(&p.ev); cudaEventRecord(p.ev, c->stream);
    c->trace_pts.push_back(p);
}
void jt_trace_dump(jt_ctx *c)
{
    if (c->trace_pts.empty()) return;
    cudaDeviceSynchronize();
    const jt_ctx::TracePt &p0 = c->trace_pts[0];
    fprintf(stderr, "[jt_trace] %-28s %9s %9s   (ms since the first point: host reached it / main stream reached it)\n", "milestone", "host", "gpu");
    for (const jt_ctx::TracePt &p : c->trace_pts) {
        float g = 0; cudaEventElapsedTime(&g, p0.ev, p.ev);
        fprintf(stderr, "[jt_trace] %-28s %9.3f %9.3f\n", p.label.c_str(), (p.host - p0.host) * 1e3, g);
    }
    for (jt_ctx::TracePt &p : c->trace_pts) cudaEventDestroy(p.ev);
    c->trace_pts.clear();
}

void jt_check_cancel(jt_ctx *c)
{
    if (c->cancel.load(std::memory_order_relaxed)) JT_THROW(JT_ERR_CANCELLED, "cancelled");
}

JtLaunch::JtLaunch(jt_ctx *ctx, const char *name, int n) : c(ctx)
{
    c->launches += n;
    if (!c->timing) return;
    for (size_t i = 0; i < c->slots.size(); i++) if (c->slots[i].name == name) { slot = (int)i; break; }
    if (slot < 0) { c->slots.push_back(JtTimingSlot{name, 0, 0}); slot = (int)c->slots.size() - 1; }
    c->slots[slot].launches += n;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, c->stream);
}
JtLaunch::~JtLaunch()
{
    if (slot < 0) return;
    cudaEventRecord(b, c->stream);
    c->pending.push_back({a, b, slot});
}
#include <chrono>
static double host_now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
JtHost::JtHost(jt_ctx *ctx, const char *name) : c(ctx)
{
    if (!c || !c->timing) return;
    std::string nm = std::string("host:") + name;
    for (size_t i = 0; i < c->slots.size(); i++) if (c->slots[i].name == nm) { slot = (int)i; break; }
    if (slot < 0) { c->slots.push_back(JtTimingSlot{nm, 0, 0}); slot = (int)c->slots.size() - 1; }
    t0 = host_now_ms();
}
JtHost::~JtHost() { if (slot >= 0) { c->slots[slot].ms += host_now_ms() - t0; c->slots[slot].launches++; } }

void jt_flush_timing(jt_ctx *c)
{
    // pending[] is in stream order: besides each group's own duration, the device-idle time between one
    // group's end and the next group's start is booked to "gap:<next group>" (host work, syncs, copies)
    cudaEvent_t prev_end = nullptr;
    std::vector<cudaEvent_t> done;
    for (auto &p : c->pending) {
        float ms = 0;
        cudaEventSynchronize(p.b);
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) c->slots[p.slot].ms += ms;
        if (prev_end && cudaEventElapsedTime(&ms, prev_end, p.a) == cudaSuccess && ms > 0.f) {
            const std::string nm = "gap:" + c->slots[p.slot].name;
            int gs = -1;
            for (size_t i = 0; i < c->slots.size(); i++) if (c->slots[i].name == nm) { gs = (int)i; break; }
            if (gs < 0) { c->slots.push_back(JtTimingSlot{nm, 0, 0}); gs = (int)c->slots.size() - 1; }
            c->slots[gs].ms += ms; c->slots[gs].launches++;
            if (ms > 50.f && getenv("JT_DEBUG_GAPS"))
                fprintf(stderr, "[jtdsp] device idle %.1f ms before %s (kernel group #%zu of %zu since the last flush)\n", ms,
                        c->slots[p.slot].name.c_str(), (size_t)(&p - &c->pending[0]), c->pending.size());
        }
        done.push_back(p.a);
        prev_end = p.b;
    }
    for (auto &p : c->pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    c->pending.clear();
}

// ---------------------------------------------------------------------------------------
// format conversion (libswresample/audioconvert.c CONV_FUNC table)
// ---------------------------------------------------------------------------------------
template <class TI, class TO>
__global__ void k_convert(const TI *__restrict__ in, TO *__restrict__ out, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = jt_conv<TI, TO>(in[i]);
}

template <class TI, class TO>
static Sig convert_t(jt_ctx *c, const Sig &in, int out_fmt)
{
    Sig o = in; o.fmt = out_fmt; o.d = jt_dalloc<TO>(c, in.n);
    if (in.n > 0) {
        JtLaunch L(c, "convert");
        k_convert<TI, TO><<<jt_grid_for(in.n, 256, c->num_sms, 16), 256, 0, c->stream>>>((const TI *)in.d, (TO *)o.d, in.n);
    }
    return o;
}

Sig jt_convert(jt_ctx *c, const Sig &in, int out_fmt)
{
    if (in.fmt == out_fmt) return in;
    switch (in.fmt * 16 + out_fmt) {
    case JT_FMT_S16 * 16 + JT_FMT_FLT: return convert_t<int16_t, float>(c, in, out_fmt);
    case JT_FMT_S16 * 16 + JT_FMT_DBL: return convert_t<int16_t, double>(c, in, out_fmt);
    case JT_FMT_FLT * 16 + JT_FMT_DBL: return convert_t<float, double>(c, in, out_fmt);
    case JT_FMT_FLT * 16 + JT_FMT_S16: return convert_t<float, int16_t>(c, in, out_fmt);
    case JT_FMT_DBL * 16 + JT_FMT_FLT: return convert_t<double, float>(c, in, out_fmt);
    case JT_FMT_DBL * 16 + JT_FMT_S16: return convert_t<double, int16_t>(c, in, out_fmt);
    case JT_FMT_S32 * 16 + JT_FMT_FLT: return convert_t<int32_t, float>(c, in, out_fmt);
    case JT_FMT_S32 * 16 + JT_FMT_DBL: return convert_t<int32_t, double>(c, in, out_fmt);
    case JT_FMT_S32 * 16 + JT_FMT_S16: return convert_t<int32_t, int16_t>(c, in, out_fmt);
    case JT_FMT_S16 * 16 + JT_FMT_S32: return convert_t<int16_t, int32_t>(c, in, out_fmt);
    case JT_FMT_FLT * 16 + JT_FMT_S32: return convert_t<float, int32_t>(c, in, out_fmt);
    case JT_FMT_DBL * 16 + JT_FMT_S32: return convert_t<double, int32_t>(c, in, out_fmt);
    }
    JT_THROW(JT_ERR_UNSUPPORTED, "sample format conversion %d -> %d", in.fmt, out_fmt);
}

// ---------------------------------------------------------------------------------------
// aformat=channel_layouts=mono : swr rematrix (libswresample/rematrix.c).  Float formats use
// 1/sqrt(2) per channel un-normalised, integer formats the normalised 0.5/0.5.
// ---------------------------------------------------------------------------------------
__global__ void k_downmix2_f32(const float2 *__restrict__ in, float *__restrict__ out, int64_t n)
{
    const float cf = 0.70710678118654752440f;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) { float2 v = in[i]; out[i] = __fadd_rn(__fmul_rn(v.x, cf), __fmul_rn(v.y, cf)); }
}
__global__ void k_downmix2_s16(const short2 *__restrict__ in, int16_t *__restrict__ out, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        short2 v = in[i];
        int s = ((int)v.x * 16384 + (int)v.y * 16384 + 16384) >> 15;
        out[i] = (int16_t)max(-32768, min(32767, s));
    }
}

// s32 stereo -> s32 mono: swr picks FLTP as its internal format (swresample.c swr_init: 32-bit integer input with a
// rematrix), mixes with the normalised 0.5 / 0.5 matrix in float and converts back -- the low 8 bits of a 32-bit sample
// do not survive (pinned on the real library: tests/golden/swr_golden.npz "stereo_s32_to_mono")
__global__ void k_downmix2_s32(const int2 *__restrict__ in, int32_t *__restrict__ out, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int2 v = in[i];
        const float l = jt_conv<int32_t, float>(v.x), r = jt_conv<int32_t, float>(v.y);
        out[i] = jt_conv<float, int32_t>(__fadd_rn(__fmul_rn(l, 0.5f), __fmul_rn(r, 0.5f)));
    }
}

Sig jt_downmix(jt_ctx *c, const void *d_in, int64_t n, int channels, int fmt, int rate)
{
    Sig o; o.fmt = fmt; o.rate = rate; o.n = n;
    if (fmt != JT_FMT_S16 && fmt != JT_FMT_S32 && fmt != JT_FMT_FLT && fmt != JT_FMT_DBL)
        JT_THROW(JT_ERR_UNSUPPORTED, "input sample format %d", fmt);
    if (channels == 1) { o.d = const_cast<void *>(d_in); return o; }
    if (channels != 2) JT_THROW(JT_ERR_UNSUPPORTED, "downmix from %d channels", channels);
    if (fmt == JT_FMT_FLT) {
        o.d = jt_dalloc<float>(c, n);
        JtLaunch L(c, "downmix");
        k_downmix2_f32<<<jt_grid_for(n, 256, c->num_sms, 16), 256, 0, c->stream>>>((const float2 *)d_in, (float *)o.d, n);
    } else if (fmt == JT_FMT_S16) {
        o.d = jt_dalloc<int16_t>(c, n);
        JtLaunch L(c, "downmix");
        k_downmix2_s16<<<jt_grid_for(n, 256, c->num_sms, 16), 256, 0, c->stream>>>((const short2 *)d_in, (int16_t *)o.d, n);
    } else if (fmt == JT_FMT_S32) {
        o.d = jt_dalloc<int32_t>(c, n);
        JtLaunch L(c, "downmix");
        k_downmix2_s32<<<jt_grid_for(n, 256, c->num_sms, 16), 256, 0, c->stream>>>((const int2 *)d_in, (int32_t *)o.d, n);
    } else JT_THROW(JT_ERR_UNSUPPORTED, "stereo f64 downmix");
    return o;
}

// ---------------------------------------------------------------------------------------
// a2: per decoder frame sum of squares / peak over all channels pooled
// (frameSumSquaresAndPeak, analyser_metrics.go:273-358).  One warp per frame.
// ---------------------------------------------------------------------------------------
template <class T>
__global__ void k_raw_frame_stats(const T *__restrict__ in, int64_t n_total, int per_frame,
                                  double *__restrict__ sumsq, double *__restrict__ peak, int64_t n_frames)
{
    const int warps_per_block = blockDim.x >> 5;
    int64_t f = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int64_t fstride = (int64_t)gridDim.x * warps_per_block;
    for (; f < n_frames; f += fstride) {
        int64_t s0 = f * (int64_t)per_frame;
        int64_t s1 = min(s0 + per_frame, n_total);
        double ss = 0, pk = 0;
        for (int64_t i = s0 + lane; i < s1; i += 32) {
            double v = jt_norm_f64(in[i]);
            ss += v * v;
            pk = fmax(pk, fabs(v));
        }
        for (int o = 16; o; o >>= 1) {
            ss += __shfl_xor_sync(0xffffffffu, ss, o);
            pk = fmax(pk, __shfl_xor_sync(0xffffffffu, pk, o));
        }
        if (lane == 0) { sumsq[f] = ss; peak[f] = pk; }
    }
}

void jt_raw_frame_stats(jt_ctx *c, const void *d_in, int64_t n_frames, int channels, int fmt, int frame_size,
                        double *d_sumsq, double *d_peak, int64_t n_src_frames)
{
    if (n_src_frames <= 0) return;
    const int64_t n_total = n_frames * channels;
    const int per = frame_size * channels;
    const int grid = jt_grid_for(n_src_frames, 8, c->num_sms, 32);
    JtLaunch L(c, "raw_frame_stats");
    if (fmt == JT_FMT_S16) k_raw_frame_stats<int16_t><<<grid, 256, 0, c->stream>>>((const int16_t *)d_in, n_total, per, d_sumsq, d_peak, n_src_frames);
    else if (fmt == JT_FMT_S32) k_raw_frame_stats<int32_t><<<grid, 256, 0, c->stream>>>((const int32_t *)d_in, n_total, per, d_sumsq, d_peak, n_src_frames);
    else if (fmt == JT_FMT_FLT) k_raw_frame_stats<float><<<grid, 256, 0, c->stream>>>((const float *)d_in, n_total, per, d_sumsq, d_peak, n_src_frames);
    else k_raw_frame_stats<double><<<grid, 256, 0, c->stream>>>((const double *)d_in, n_total, per, d_sumsq, d_peak, n_src_frames);
}

// ---------------------------------------------------------------------------------------
// packed little-endian 24-bit PCM -> s32 with the sample in the top 24 bits (libavcodec pcm.c, pcm_s24le).  Four samples
// (12 bytes, three aligned words when the source is 4-byte aligned) per thread.
// ---------------------------------------------------------------------------------------
__global__ void k_unpack_s24(const uint8_t *__restrict__ in, int64_t n, int32_t *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool aligned = (((uintptr_t)in) & 3) == 0;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < n; q += stride) {
        const int64_t i = q * 4;
        if (aligned && i + 4 <= n) {
            const uint32_t *w = (const uint32_t *)(in + i * 3);
            const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
            out[i] = (int32_t)(w0 << 8);
            out[i + 1] = (int32_t)((w0 >> 24) << 8 | (w1 << 16));
            out[i + 2] = (int32_t)((w1 >> 16) << 8 | (w2 << 24));
            out[i + 3] = (int32_t)(w2 & 0xFFFFFF00u);
        } else {
            for (int64_t k = i; k < n && k < i + 4; k++) {
                const uint8_t *p = in + k * 3;
                out[k] = (int32_t)(((uint32_t)p[0] << 8) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 24));
            }
        }
    }
}
void jt_unpack_s24(jt_ctx *c, const uint8_t *d_bytes, int64_t n_samples, int32_t *d_out)
{
    if (n_samples <= 0) return;
    JtLaunch L(c, "wav_decode:s24");
    k_unpack_s24<<<jt_grid_for((n_samples + 3) / 4, 256, c->num_sms, 16), 256, 0, c->stream>>>(d_bytes, n_samples, d_out);
}

// ---------------------------------------------------------------------------------------
// volume (af_volume.c, precision=float: fltp * (float)volume) and loudnorm's linear gain
// ---------------------------------------------------------------------------------------
__global__ void k_scale_f32(const float *__restrict__ in, float *__restrict__ out, int64_t n, float g)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = __fmul_rn(in[i], g);
}
__global__ void k_scale_f64(const double *__restrict__ in, double *__restrict__ out, int64_t n, double g)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = __dmul_rn(in[i], g);
}

Sig jt_volume(jt_ctx *c, const Sig &in0, double volume)
{
    Sig in = jt_convert(c, in0, JT_FMT_FLT);
    Sig o = in; o.d = jt_dalloc<float>(c, in.n);
    if (in.n > 0) {
        JtLaunch L(c, "volume");
        k_scale_f32<<<jt_grid_for(in.n, 256, c->num_sms, 16), 256, 0, c->stream>>>((const float *)in.d, (float *)o.d, in.n, (float)volume);
    }
    return o;
}

Sig jt_gain_f64(jt_ctx *c, const Sig &in0, double gain)
{
    Sig in = jt_convert(c, in0, JT_FMT_DBL);
    Sig o = in; o.d = jt_dalloc<double>(c, in.n);
    if (in.n > 0) {
        JtLaunch L(c, "loudnorm_linear_gain");
        k_scale_f64<<<jt_grid_for(in.n, 256, c->num_sms, 16), 256, 0, c->stream>>>((const double *)in.d, (double *)o.d, in.n, gain);
    }
    return o;
}

Sig jt_slice(const Sig &in, int64_t start, int64_t count)
{
    Sig o = in;
    if (start < 0) start = 0;
    if (start > in.n) start = in.n;
    if (count > in.n - start) count = in.n - start;
    if (count < 0) count = 0;
    o.n = count;
    o.d = (char *)in.d + (size_t)start * jt_fmt_bytes(in.fmt);
    return o;
}

Sig jt_pad_zero(jt_ctx *c, const Sig &in, int64_t n_total)
{
    if (n_total <= in.n) return in;
    Sig o = in; o.n = n_total;
    const size_t b = jt_fmt_bytes(in.fmt);
    o.d = jt_dalloc_bytes(c, (size_t)n_total * b);
    JT_CUDA(cudaMemcpyAsync(o.d, in.d, (size_t)in.n * b, cudaMemcpyDeviceToDevice, c->stream));
    JT_CUDA(cudaMemsetAsync((char *)o.d + (size_t)in.n * b, 0, (size_t)(n_total - in.n) * b, c->stream));
    return o;
}

#include <mutex>
void jt_smem_optin(const void *kernel, size_t bytes)
{
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, size_t> granted;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(mu);
    size_t &have = granted[{dev, kernel}];
    if (bytes <= have) return;
    JT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    have = bytes;
}

// ---------------------------------------------------------------------------------------
// tensor map of a stream seen as rows of `seg` elements (jt_tiles.cuh).  cuTensorMapEncodeTiled is a driver entry point;
// it is looked up through the runtime so the library keeps linking against cudart only.
// ---------------------------------------------------------------------------------------
#include "jt_tiles.cuh"
#include <mutex>
typedef CUresult (*jt_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                       const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static jt_encode_tiled_fn jt_encode_tiled()
{
    static jt_encode_tiled_fn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (jt_encode_tiled_fn)p;
        cudaGetLastError();
    });
    return fn;
}
bool jt_lane_tensor_map(CUtensorMap *map, const void *base, int elem_bytes, int64_t n, int64_t seg, int lines_per_tile)
{
    jt_encode_tiled_fn enc = jt_encode_tiled();
    const int64_t rows = seg > 0 ? n / seg : 0;
    const int epl = 128 / elem_bytes;
    if (!enc || rows < 1 || rows > 0x7FFFFFFFll || (((uintptr_t)base) & 15) || seg % epl || (seg * elem_bytes) % 16 || lines_per_tile < 1 || lines_per_tile > 256) return false;
    const CUtensorMapDataType dt = elem_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16;
    const cuuint64_t dims[3] = {(cuuint64_t)epl, (cuuint64_t)rows, (cuuint64_t)(seg / epl)};
    const cuuint64_t strides[2] = {(cuuint64_t)seg * elem_bytes, 128};
    const cuuint32_t box[3] = {(cuuint32_t)epl, 32, (cuuint32_t)lines_per_tile};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, dt, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
