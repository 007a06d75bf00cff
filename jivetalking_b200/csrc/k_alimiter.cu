// alimiter (libavfilter/af_alimiter.c, Calf lookahead limiter port), f64 mono, latency=1:
//   "alimiter=limit=%.6f:attack=5:release=100:level_in=1:level_out=1:level=0:latency=1:asc=1:asc_level=0.8"
//   "alimiter=limit=0.803526:attack=1:release=50:..."   (reference: normalise.go:446-480)
// The limiter is a sequential state machine (gain `att`, slope `delta`, a queue of upcoming
// peaks inside the lookahead window).  Its state is re-anchored by every controlling peak and
// returns to exactly (att=1, delta=0, empty queue) one release time after the last over-limit
// sample, so the stream is cut into segments, one lane each, started 0.75 s early from the idle
// state.  The circular buffer of the C code is replaced by absolute sample indices (the
// buffer holds the last `buffer_size` inputs, i.e. x itself): the sample entering the lookahead
// window and the one leaving it are two TMA-staged views of the same stream, `latency` apart
// (jt_lanes.cuh); the peak queue lives in the lane's local memory and is touched only while
// limiting.  asc=1 only selects the release branch: with FFmpeg's negative asc_coeff the ASC
// term (a_att > patt) can never fire, so its accumulators are not kept.
#include "jt_internal.h"
#include "jt_device.cuh"
#include "jt_lanes.cuh"

#define LIM_MAXBUF 512
#define LIM_R 32
typedef LaneStage<double, LIM_R, 2> LimIn;      // 256-byte rows, double-buffered (x2 views) + output rows: 52 KB per warp
typedef LaneStore<double, LIM_R> LimOut;

__global__ void __launch_bounds__(64)
k_alimiter(const double *__restrict__ x, double *__restrict__ y, int64_t n, int seg, int warm, int rate, int bs,
           double limit, double release, double level_in, double level_out, double level, int auto_release)
{
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned char *wsm = smem + (size_t)(threadIdx.x >> 5) * (2 * LimIn::WARP_BYTES + LimOut::WARP_BYTES);
    const int64_t lane = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = lane * seg < n;
    const int64_t o0 = min(lane * (int64_t)seg, n), o1 = min(o0 + (int64_t)seg, n);
    const int latency = bs - 1;
    // step t consumes input t (zero past the end) and emits output t - latency
    const int64_t t_first = o0 + latency, t_last = o1 + latency;
    // warm is a multiple of the tile size and t0 may be negative (samples before the stream are
    // zeros, the limiter idles through them), so every tile is either all warm-up or all output
    const int64_t t0 = t_first - warm;
    const int warm_tiles = warm / LIM_R;
    const int64_t count = live ? t_last - t0 : 0;
    LimIn A, B; LimOut out;
    A.init(wsm, x + t0, count, max((int64_t)0, -t0), n - t0);                                               // entering sample x[t]
    B.init(wsm + LimIn::WARP_BYTES, x + t0 - latency, count, max((int64_t)0, latency - t0), n - t0 + latency);   // leaving sample x[t - latency]
    out.init(wsm + 2 * LimIn::WARP_BYTES, y + o0);

    int64_t qpos[LIM_MAXBUF]; double qdelta[LIM_MAXBUF];
    for (int i = 0; i < bs; i++) qpos[i] = -1;
    double att = 1.0, delta = 0.0; int nextiter = 0, nextlen = 0;
    int64_t head = -1;                                   // qpos[nextiter], cached
    const double rr = rate * release;                    // divided by, as the scalar code does
    auto inp = [&](int64_t s) -> double { return (s >= 0 && s < n) ? x[s] * level_in : 0.0; };

    A.prime(); B.prime();
    for (int tile = 0; tile < A.ntiles; tile++) {
        A.prefetch(); B.prefetch();
        const double *ra = A.wait(tile), *rb = B.wait(tile);
        const int nv = A.valid(tile);
        int64_t t = t0 + (int64_t)tile * LIM_R;
        double *orow = out.row();
        for (int k = 0; k < nv; k++, t++) {
            if (att == 1.0 && delta == 0.0 && nextlen == 0) {
                // idle limiter (the common case): unity gain until the next over-limit sample enters.
                // Blocks of 8 with all loads first, so the shared-memory latency is paid once per block.
                while (k + 8 <= nv) {
                    double a[8], b[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) { a[j] = ra[k + j]; b[j] = rb[k + j]; }
                    bool any = false;
#pragma unroll
                    for (int j = 0; j < 8; j++) any |= fabs(a[j] * level_in) > limit;
                    if (any) break;
#pragma unroll
                    for (int j = 0; j < 8; j++) orow[k + j] = fmin(fmax(b[j] * level_in, -limit), limit) * level * level_out;
                    k += 8; t += 8;
                }
                while (k < nv && !(fabs(ra[k] * level_in) > limit)) {
                    orow[k] = fmin(fmax(rb[k] * level_in, -limit), limit) * level * level_out;
                    k++; t++;
                }
                if (k >= nv) break;
            }
            const double sample = ra[k] * level_in;
            double peak = fabs(sample);
            if (peak > limit) {
                const double patt = fmin(limit / peak, 1.);
                const double rdelta = (1.0 - patt) / rr;
                const double d = (limit / peak - att) / bs;
                bool found = false; int i;
                if (d < delta) {
                    delta = d;
                    qpos[0] = t; qpos[1 % bs] = -1; qdelta[0] = rdelta;
                    nextlen = 1; nextiter = 0;
                } else {
                    for (i = nextiter; i < nextiter + nextlen; i++) {
                        const int j = i % bs;
                        const double ppeak = fabs(inp(qpos[j]));
                        const int dist = (int)((t - qpos[j]) % bs);
                        const double pdelta = (limit / peak - limit / ppeak) / (double)dist;
                        if (pdelta < qdelta[j]) { qdelta[j] = pdelta; found = true; break; }
                    }
                    if (found) {
                        nextlen = i - nextiter + 1;
                        qpos[(nextiter + nextlen) % bs] = t;
                        qdelta[(nextiter + nextlen) % bs] = rdelta;
                        qpos[(nextiter + nextlen + 1) % bs] = -1;
                        nextlen++;
                    }
                }
                head = qpos[nextiter];
            }
            // oldest sample in the lookahead buffer
            const double bufv = rb[k] * level_in;
            peak = fabs(bufv);
            att += delta;
            double o = bufv * att;
            // (pos + 1) % bs == nextpos[nextiter]  <=>  the queue head is the sample leaving next
            if (head >= 0 && ((t + 1 - head) % bs) == 0) {
                if (auto_release) {
                    delta = (1.0 - att) / rr;
                    if (nextlen > 1) {
                        const int64_t pn = qpos[(nextiter + 1) % bs];
                        const double ppeak = fabs(inp(pn));
                        int dist = (int)((pn - (t + 1)) % bs); if (dist < 0) dist += bs;
                        const double pdelta = (limit / ppeak - att) / (double)dist;
                        if (pdelta < delta) delta = pdelta;
                    }
                } else {
                    delta = qdelta[nextiter];
                    att = limit / peak;
                }
                nextlen -= 1;
                qpos[nextiter] = -1;
                nextiter = (nextiter + 1) % bs;
                head = qpos[nextiter];
            }
            if (att > 1.) { att = 1.; delta = 0.; nextiter = 0; nextlen = 0; qpos[0] = -1; head = -1; }
            if (att <= 0.) { att = 0.0000000000001; delta = (1.0 - att) / rr; }
            if (att != 1. && (1. - att) < 0.0000000000001) att = 1.;
            if (delta != 0. && fabs(delta) < 0.00000000000001) delta = 0.;
            o = fmin(fmax(o, -limit), limit) * level * level_out;
            orow[k] = o;
        }
        if (tile >= warm_tiles) out.commit(nv);
        A.release(); B.release();
    }
    out.finish();
}

Sig jt_alimiter(jt_ctx *c, const Sig &in, const LimiterParams &p)
{
    if (in.fmt != JT_FMT_DBL) JT_THROW(JT_ERR_INVALID_ARG, "alimiter expects f64 input");
    const double attack = p.attack_ms / 1000., release = p.release_ms / 1000.;
    const int bs = (int)(in.rate * attack);
    if (bs < 1) JT_THROW(JT_ERR_INVALID_ARG, "alimiter attack too short for %d Hz", in.rate);
    if (bs > LIM_MAXBUF) JT_THROW(JT_ERR_UNSUPPORTED, "alimiter lookahead of %d samples (max %d)", bs, LIM_MAXBUF);
    Sig o = in; o.d = jt_dalloc<double>(c, in.n);
    if (in.n <= 0) return o;
    // one wave: 4 warps per SM hold their staging at once (the walk is latency-bound per lane, and every lane pays
    // the 0.6 s warm-up, so more, shorter lanes finish sooner); segments grow before a second wave would start
    const int64_t slots = (int64_t)c->num_sms * 4 * 32;
    const int seg = (int)std::min<int64_t>(((std::max<int64_t>(4096, (in.n + slots - 1) / slots) + LIM_R - 1) / LIM_R) * LIM_R, 1 << 24);
    int warm = (int)(in.rate * (0.5 + 2 * release + 2 * attack)) + 2 * bs;
    warm = (warm + LIM_R - 1) / LIM_R * LIM_R;                  // tile-aligned (see the kernel)
    const int64_t lanes = (in.n + seg - 1) / seg;
    const double level = p.auto_level ? 1 / p.limit : 1;
    const size_t smem = 2 * (2 * LimIn::WARP_BYTES + LimOut::WARP_BYTES);
    jt_smem_optin((const void *)k_alimiter, (size_t)(smem));
    JtLaunch L(c, "alimiter");
    k_alimiter<<<(int)((lanes + 63) / 64), 64, smem, c->stream>>>((const double *)in.d, (double *)o.d, in.n, seg, warm, in.rate, bs,
                                                                    p.limit, release, p.level_in, p.level_out, level, p.asc);
    return o;
}
