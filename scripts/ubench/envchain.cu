// micro-benchmark: cycles per sample of the envelope-follower step in several formulations, data from shared memory,
// W warps per CTA (one CTA per SM).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o envchain envchain.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int V, int LANES>
__global__ void k(double *out, long long *cyc, double a, double r, int n)
{
    __shared__ double sx[4][32][34];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    for (int i = 0; i < 34; i++) sx[w & 3][l][i] = 0.001 * ((l * 7 + i * 13) % 97);
    __syncthreads();
    const double ka = 1.0 - a, kr = 1.0 - r;
    double e[LANES];
    for (int q = 0; q < LANES; q++) e[q] = 0.01 * q;
    long long t0 = clock64();
    for (int it = 0; it < n; it += 32) {
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const double v = sx[w & 3][l][j];
#pragma unroll
            for (int q = 0; q < LANES; q++) {
                const double d = v * v + 1e-9 * q;
                if (V == 0) {            // scalar C form, DSETP
                    const double t = __dsub_rn(d, e[q]);
                    const double ea = __dadd_rn(e[q], __dmul_rn(t, a)), er = __dadd_rn(e[q], __dmul_rn(t, r));
                    e[q] = d > e[q] ? ea : er;
                } else if (V == 1) {     // fma form, integer compare
                    const double ea = fma(e[q], ka, a * d), er = fma(e[q], kr, r * d);
                    e[q] = __double_as_longlong(d) > __double_as_longlong(e[q]) ? ea : er;
                } else if (V == 2) {     // fma form, DSETP
                    const double ea = fma(e[q], ka, a * d), er = fma(e[q], kr, r * d);
                    e[q] = d > e[q] ? ea : er;
                } else if (V == 3) {     // fma form, max (a >= r)
                    const double ea = fma(e[q], ka, a * d), er = fma(e[q], kr, r * d);
                    e[q] = fmax(ea, er);
                } else if (V == 4) {     // fma form, compare on the high words only (approximate: for timing)
                    const double ea = fma(e[q], ka, a * d), er = fma(e[q], kr, r * d);
                    e[q] = __double2hiint(d) > __double2hiint(e[q]) ? ea : er;
                } else if (V == 5) {     // select the coefficient first
                    const bool up = __double_as_longlong(d) > __double_as_longlong(e[q]);
                    const double c = up ? a : r, kk = up ? ka : kr;
                    e[q] = fma(e[q], kk, c * d);
                }
            }
        }
    }
    long long t1 = clock64();
    double s = 0; for (int q = 0; q < LANES; q++) s += e[q];
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int V, int LANES> void run(const char *name, int warps, double *o, long long *c)
{
    const int n = 1 << 16;
    for (int rep = 0; rep < 2; rep++) { k<V, LANES><<<148, 32 * warps>>>(o, c, 1.0 / 60, 1.0 / 2400, n); cudaDeviceSynchronize(); }
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-34s lanes/thread %d warps/SM %d  %.2f cycles/sample/lane  (%.2f per thread-step)\n", name, LANES, warps, (double)h / n / LANES, (double)h / n);
}
int main()
{
    double *o; long long *c; cudaMalloc(&o, 8 * 148 * 1024); cudaMalloc(&c, 8);
    for (int warps : {1, 2, 4, 8}) {
        if (warps == 1) { run<0, 1>("scalar C form, DSETP", 1, o, c); run<1, 1>("fma, int64 compare", 1, o, c); run<2, 1>("fma, DSETP", 1, o, c); run<3, 1>("fma, fmax", 1, o, c);
                          run<4, 1>("fma, hi-word compare", 1, o, c); run<5, 1>("select coefficient first", 1, o, c);
                          run<1, 2>("fma, int64 compare", 1, o, c); run<1, 4>("fma, int64 compare", 1, o, c); run<3, 2>("fma, fmax", 1, o, c); run<3, 4>("fma, fmax", 1, o, c); run<2, 2>("fma, DSETP", 1, o, c); run<2, 4>("fma, DSETP", 1, o, c); }
        else { if (warps == 2) { run<1, 1>("fma, int64 compare", 2, o, c); run<1, 2>("fma, int64 compare", 2, o, c); }
               if (warps == 4) { run<1, 1>("fma, int64 compare", 4, o, c); run<1, 2>("fma, int64 compare", 4, o, c); run<1, 4>("fma, int64 compare", 4, o, c); }
               if (warps == 8) { run<1, 1>("fma, int64 compare", 8, o, c); run<1, 2>("fma, int64 compare", 8, o, c); } }
    }
    return 0;
}
