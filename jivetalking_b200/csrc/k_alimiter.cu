// alimiter (libavfilter/af_alimiter.c, Calf lookahead limiter port), f64 mono, latency=1:
//   "alimiter=limit=%.6f:attack=5:release=100:level_in=1:level_out=1:level=0:latency=1:asc=1:asc_level=0.8"
//   "alimiter=limit=0.803526:attack=1:release=50:..."   (reference: normalise.go:446-480)
// The limiter is a sequential state machine (gain `att`, slope `delta`, a queue of upcoming
// peaks inside the lookahead window).  Its state is re-anchored by every controlling peak and
// returns to exactly (att=1, delta=0, empty queue) one release time after the last over-limit
// sample, so the stream is cut into segments, one lane each, started 0.75 s early from the idle
// state.  The circular buffer of the C code is replaced by absolute sample indices (the
// buffer holds the last `buffer_size` inputs, i.e. x itself); the peak queue lives in the
// lane's local memory.  asc=1 only selects the release branch: with FFmpeg's negative
// asc_coeff the ASC term (a_att > patt) can never fire, so its accumulators are not kept.
#include "jt_internal.h"
#include "jt_device.cuh"

#define LIM_MAXBUF 512

__global__ void __launch_bounds__(64)
k_alimiter(const double *__restrict__ x, double *__restrict__ y, int64_t n, int seg, int warm, int rate, int bs,
           double limit, double release, double level_in, double level_out, double level, int auto_release)
{
    const int64_t lane = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t o0 = lane * seg; if (o0 >= n) return;
    const int64_t o1 = min(o0 + (int64_t)seg, n);
    const int latency = bs - 1;
    int64_t qpos[LIM_MAXBUF]; double qdelta[LIM_MAXBUF];
    for (int i = 0; i < bs; i++) qpos[i] = -1;
    double att = 1.0, delta = 0.0; int nextiter = 0, nextlen = 0;
    // step t consumes input t (zero past the end) and emits output t - latency
    const int64_t t_first = o0 + latency, t_last = o1 + latency;
    int64_t t = max((int64_t)0, t_first - warm);
    auto inp = [&](int64_t s) -> double { return (s >= 0 && s < n) ? x[s] * level_in : 0.0; };
    for (; t < t_last; t++) {
        const double sample = inp(t);
        double peak = fabs(sample);
        if (peak > limit) {
            const double patt = fmin(limit / peak, 1.);
            const double rdelta = (1.0 - patt) / (rate * release);
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
        }
        // oldest sample in the lookahead buffer
        const int64_t t_old = t - latency;
        const double bufv = inp(t_old);
        peak = fabs(bufv);
        att += delta;
        double out = bufv * att;
        // (pos + 1) % bs == nextpos[nextiter]  <=>  the queue head is the sample leaving next
        if (qpos[nextiter] >= 0 && ((t + 1 - qpos[nextiter]) % bs) == 0) {
            if (auto_release) {
                delta = (1.0 - att) / (rate * release);
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
        }
        if (att > 1.) { att = 1.; delta = 0.; nextiter = 0; nextlen = 0; qpos[0] = -1; }
        if (att <= 0.) { att = 0.0000000000001; delta = (1.0 - att) / (rate * release); }
        if (att != 1. && (1. - att) < 0.0000000000001) att = 1.;
        if (delta != 0. && fabs(delta) < 0.00000000000001) delta = 0.;
        out = fmin(fmax(out, -limit), limit) * level * level_out;
        if (t >= t_first) y[t - latency] = out;
    }
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
    const int seg = 32768;
    const int warm = (int)(in.rate * (0.5 + 2 * release + 2 * attack)) + 2 * bs;
    const int64_t lanes = (in.n + seg - 1) / seg;
    const double level = p.auto_level ? 1 / p.limit : 1;
    JtLaunch L(c, "alimiter");
    k_alimiter<<<(int)((lanes + 63) / 64), 64, 0, c->stream>>>((const double *)in.d, (double *)o.d, in.n, seg, warm, in.rate, bs,
                                                                 p.limit, release, p.level_in, p.level_out, level, p.asc);
    return o;
}
