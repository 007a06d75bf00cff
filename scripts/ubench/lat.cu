// micro-benchmark: dependent-chain latency of f64 / f32 ops on this GPU (one warp)
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void k(double *out, long long *cyc, double a, double b, int n)
{
    double x = a; float xf = (float)a, bf = (float)b;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        if (OP == 0) x = fma(x, b, a);
        if (OP == 1) x = __dadd_rn(x, b);
        if (OP == 2) x = __dmul_rn(x, b);
        if (OP == 3) xf = fmaf(xf, bf, bf);
        if (OP == 4) { x = __dadd_rn(__dmul_rn(__dadd_rn(b, -x), a), x); }      // envelope step: 3 dependent ops
        if (OP == 5) { x = x > b ? x * a : x + a; }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = x + xf; cyc[0] = t1 - t0; }
}
int main()
{
    double *o; long long *c; cudaMalloc(&o, 8); cudaMalloc(&c, 8);
    const char *names[] = {"DFMA", "DADD", "DMUL", "FFMA", "env(3 dep f64)", "cmp/sel"};
    const int n = 100000;
    for (int op = 0; op < 6; op++) {
        for (int rep = 0; rep < 2; rep++) {
            switch (op) {
            case 0: k<0><<<1, 32>>>(o, c, 0.999, 1.0001, n); break;
            case 1: k<1><<<1, 32>>>(o, c, 0.999, 1.0001, n); break;
            case 2: k<2><<<1, 32>>>(o, c, 0.999, 1.0001, n); break;
            case 3: k<3><<<1, 32>>>(o, c, 0.999, 1.0001, n); break;
            case 4: k<4><<<1, 32>>>(o, c, 0.001, 0.5, n); break;
            case 5: k<5><<<1, 32>>>(o, c, 0.999, 1.0001, n); break;
            }
            cudaDeviceSynchronize();
        }
        long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("%-16s %.2f cycles/iter\n", names[op], (double)h / n);
    }
    return 0;
}
