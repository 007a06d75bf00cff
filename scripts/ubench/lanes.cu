// micro-benchmark: how fast can a warp of "sequential lanes" be fed?  Every lane walks its own contiguous range
// (segments far apart in memory) and only sums what it reads, so the time is the staging mechanism's:
//   mode 0  LaneStage (jt_lanes.cuh): one 1-D TMA bulk copy per lane per tile, double-buffered
//   mode 1  warp-cooperative cp.async: the warp copies row after row with 16-byte cp.async (coalesced), double-buffered
//   mode 2  plain per-lane global loads (strided across the warp)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../jivetalking_b200/csrc -o lanes lanes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "jt_lanes.cuh"

template <int R>
__global__ void __launch_bounds__(32) k_tma(const double *x, int64_t seg, double *out)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int64_t lane = (int64_t)blockIdx.x * 32 + threadIdx.x;
    LaneStage<double, R> in;
    in.init(smem, x + lane * seg, seg);
    double s = 0;
    in.prefetch();
    for (int t = 0; t < in.ntiles; t++) {
        in.prefetch();
        const double *row = in.wait(t);
#pragma unroll 8
        for (int k = 0; k < R; k++) s += row[k];
        in.release();
    }
    out[lane] = s;
}

template <int R>
__global__ void __launch_bounds__(32) k_cpasync(const double *x, int64_t seg, double *out)
{
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int ROW = R + 2;
    double *buf = (double *)smem;                                  // [2][32][ROW]
    const int lane = threadIdx.x;
    const int64_t w0 = (int64_t)blockIdx.x * 32;
    const int ntiles = (int)(seg / R);
    auto issue = [&](int t) {
        double *b = buf + (size_t)(t & 1) * 32 * ROW;
        for (int r = 0; r < 32; r++) {
            const double *src = x + (w0 + r) * seg + (int64_t)t * R;
            for (int c = lane * 2; c < R; c += 64) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(b + (size_t)r * ROW + c);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + c) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    double s = 0;
    issue(0);
    for (int t = 0; t < ntiles; t++) {
        if (t + 1 < ntiles) { issue(t + 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        const double *row = buf + ((size_t)(t & 1) * 32 + lane) * ROW;
#pragma unroll 8
        for (int k = 0; k < R; k++) s += row[k];
        __syncwarp();
    }
    out[w0 + lane] = s;
}

__global__ void __launch_bounds__(32) k_plain(const double *x, int64_t seg, double *out)
{
    const int64_t lane = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const double *p = x + lane * seg;
    double s = 0;
#pragma unroll 8
    for (int64_t k = 0; k < seg; k++) s += p[k];
    out[lane] = s;
}

int main()
{
    const int64_t seg = 32768;
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int wps = 1; wps <= 8; wps *= 2) {
        const int warps = p.multiProcessorCount * wps;
        const int64_t n = (int64_t)warps * 32 * seg;
        double *x, *o; cudaMalloc(&x, n * 8); cudaMalloc(&o, warps * 32 * 8); cudaMemset(x, 0, n * 8);
        auto run = [&](const char *name, auto launch) {
            float best = 1e30f;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            cudaError_t err = cudaGetLastError();
            printf("warps/SM %d  %-22s %8.3f ms  %7.1f GB/s  %6.1f cycles/sample/lane%s\n", wps, name, best, n * 8 / best / 1e6,
                   best * 1e-3 * 1.965e9 / seg, err ? cudaGetErrorString(err) : "");
        };
#define TMA(Rv) do { size_t sm = LaneStage<double, Rv>::WARP_BYTES; cudaFuncSetAttribute(k_tma<Rv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
        run("tma R=" #Rv, [&] { k_tma<Rv><<<warps, 32, sm>>>(x, seg, o); }); } while (0)
#define CPA(Rv) do { size_t sm = 2 * 32 * (Rv + 2) * 8; cudaFuncSetAttribute(k_cpasync<Rv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
        run("cp.async R=" #Rv, [&] { k_cpasync<Rv><<<warps, 32, sm>>>(x, seg, o); }); } while (0)
        TMA(32); TMA(64); TMA(128);
        if (wps <= 4) { CPA(32); CPA(64); CPA(128); } else { CPA(32); CPA(64); }
        run("plain loads", [&] { k_plain<<<warps, 32>>>(x, seg, o); });
        cudaFree(x); cudaFree(o);
    }
    return 0;
}
