// micro-benchmark: per-SM issue throughput of the ops the f64 stages are made of (whole GPU, 8 independent
// chains per thread), so DESIGN.md's compute rooflines use measured numbers.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o thr thr.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void __launch_bounds__(256) k(double *out, double a, double b, int n)
{
    __shared__ double s[256 + 64];
    s[threadIdx.x] = a * threadIdx.x; if (threadIdx.x < 64) s[256 + threadIdx.x] = b;
    __syncthreads();
    double x[8]; float f[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { x[j] = a + j * 1e-3 + threadIdx.x * 1e-6; f[j] = (float)x[j]; }
    const float bf = (float)b, af = (float)a;
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (OP == 0) x[j] = fma(x[j], b, a);
            if (OP == 1) x[j] = __dadd_rn(x[j], b);
            if (OP == 2) x[j] = __dmul_rn(x[j], b);
            if (OP == 3) f[j] = fmaf(f[j], bf, af);
            if (OP == 4) x[j] = __dadd_rn(__dmul_rn(x[j], b), a);
            if (OP == 5) x[j] += s[(threadIdx.x + j + i) & 255];                  // LDS.64 + DADD
            if (OP == 6) x[j] = fma(s[(threadIdx.x + j + i) & 255], b, x[j]);     // LDS.64 + DFMA (FIR inner loop)
            if (OP == 7) x[j] = fmax(x[j], fabs(x[j] - b));
        }
    }
    double r = 0; for (int j = 0; j < 8; j++) r += x[j] + f[j];
    if (r == 12345.678) out[0] = r;
}
int main()
{
    double *o; cudaMalloc(&o, 8);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const char *names[] = {"DFMA", "DADD", "DMUL", "FFMA", "DMUL+DADD", "LDS64+DADD", "LDS64+DFMA", "DADD+DMNMX"};
    const int opsPerIter[] = {1, 1, 1, 1, 2, 1, 1, 2};
    const int n = 4096, grid = p.multiProcessorCount * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("%s, %d SMs, max clock %d MHz\n", p.name, p.multiProcessorCount, khz / 1000);
    for (int op = 0; op < 8; op++) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            switch (op) {
            case 0: k<0><<<grid, 256>>>(o, 0.999, 1.0001, n); break;
            case 1: k<1><<<grid, 256>>>(o, 0.999, 1.0001, n); break;
            case 2: k<2><<<grid, 256>>>(o, 0.999, 1.0001, n); break;
            case 3: k<3><<<grid, 256>>>(o, 0.999, 1.0001, n); break;
            case 4: k<4><<<grid, 256>>>(o, 0.999, 1.0001, n); break;
            case 5: k<5><<<grid, 256>>>(o, 0.999, 1.0001, n); break;
            case 6: k<6><<<grid, 256>>>(o, 0.999, 1.0001, n); break;
            case 7: k<7><<<grid, 256>>>(o, 0.999, 1.0001, n); break;
            }
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double ops = (double)grid * 256 * n * 8 * opsPerIter[op];
        printf("%-12s %8.3f ms  %8.2f Gop/s  %7.2f thread-ops/clk/SM (at max clock)\n", names[op], best, ops / best / 1e6,
               ops / (best * 1e-3) / p.multiProcessorCount / (khz * 1e3));
    }
    return 0;
}
