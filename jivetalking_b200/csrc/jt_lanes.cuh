// Warp-cooperative staging for "sequential lane" kernels (sm_100a).
//
// A lane kernel gives every thread one long contiguous range of a stream to walk sequentially
// (a recurrence forbids anything else), so the 32 lanes of a warp read 32 far-apart addresses:
// un-coalesced, and with only a few independent loads in flight per lane the walk is bound by
// DRAM latency, not by the recurrence.  LaneStage fixes that with the TMA engine: each lane
// owns one row of a shared-memory tile and asks for the next R elements of its own range with a
// single 1-D bulk copy (cp.async.bulk.shared::cluster.global -> SASS UBLKCP) that completes on
// an mbarrier; tiles sit in a ring of NS buffers, NS - 1 rows in flight while one is being consumed.
// A lane's bulk copy takes 2700 (256 B) to 4100 (1 KB) cycles to land (scripts/ubench/lanes.cu), so a
// walk that spends c cycles per element needs (NS - 1) * R * c above that to stay compute-bound.  A lane's range is given in "virtual" coordinates: elements outside
// [lo, hi) read as zero (stream start / zero tail), rows near those edges are filled with plain
// loads by their own lane, and a source that is not 16-byte aligned is fetched from the aligned
// address below it (the row has 16 bytes of slack) and exposed with the matching offset.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t jt_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void jt_mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(jt_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void jt_mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(jt_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void jt_mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(jt_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void jt_tma_load_1d(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(jt_smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(jt_smem_u32(bar)) : "memory");
}

template <class T, int R, int NS = 2>
struct LaneStage {
    static_assert((R * sizeof(T)) % 16 == 0, "row must be a multiple of 16 bytes");
    static_assert(NS >= 2 && NS <= 8, "2..8 stages");
    static constexpr int PAD = 16 / (int)sizeof(T);
    static constexpr int ROW = R + PAD;                                  // 16 bytes of slack per row
    static constexpr size_t WARP_BYTES = NS * 32 * (size_t)ROW * sizeof(T) + 64;

    T *buf;                 // [NS][32][ROW]
    uint64_t *bar;          // [NS]
    const T *src;           // address of virtual element 0 (may lie outside the array)
    int64_t count, lo, hi;  // virtual length; real data only for lo <= i < hi, zero elsewhere
    int shift;              // elements between the 16-byte boundary below src and src
    int issued, ntiles;     // ntiles is warp-uniform: max over lanes of ceil(count / R)

    // smem: WARP_BYTES for this warp, 16-byte aligned.  All 32 lanes must call.
    __device__ __forceinline__ void init(unsigned char *smem, const T *virt0, int64_t lane_count, int64_t real_lo, int64_t real_hi)
    {
        buf = (T *)smem;
        bar = (uint64_t *)(smem + NS * 32 * (size_t)ROW * sizeof(T));
        src = virt0; count = lane_count < 0 ? 0 : lane_count; lo = real_lo; hi = real_hi; issued = 0;
        shift = (int)((((uintptr_t)virt0) & 15) / sizeof(T));
        int64_t nt = (count + R - 1) / R;
        for (int o = 16; o; o >>= 1) { int64_t v = __shfl_xor_sync(0xffffffffu, nt, o); nt = v > nt ? v : nt; }
        ntiles = (int)nt;
        if ((threadIdx.x & 31) < NS) jt_mbar_init(&bar[threadIdx.x & 31], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
    }
    __device__ __forceinline__ void init(unsigned char *smem, const T *lane_src, int64_t lane_count)
    {
        init(smem, lane_src, lane_count, 0, lane_count);
    }
    __device__ __forceinline__ int valid(int tile) const
    {
        const int64_t rem = count - (int64_t)tile * R;
        return rem <= 0 ? 0 : (rem < R ? (int)rem : R);
    }
    // issue the next tile (no-op past the end); the buffer it overwrites (tile - NS) must have been released
    __device__ __forceinline__ void prefetch()
    {
        if (issued >= ntiles) return;
        const int t = issued, b = t % NS, lane = threadIdx.x & 31;
        const int n = valid(t);
        T *row = buf + ((size_t)b * 32 + lane) * ROW;
        const int64_t v0 = (int64_t)t * R;                               // first virtual element of the tile
        const bool tma = n == R && v0 - shift >= lo && v0 + R + PAD - shift <= hi;
        const unsigned bytes = tma ? (unsigned)((R + (shift ? PAD : 0)) * sizeof(T)) : 0u;
        unsigned total = bytes;
        for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
        if (lane == 0) jt_mbar_expect_tx(&bar[b], total);
        __syncwarp();
        if (tma) jt_tma_load_1d(row, src + v0 - shift, bytes, &bar[b]);
        else for (int i = 0; i < n; i++) { const int64_t v = v0 + i; row[shift + i] = (v >= lo && v < hi) ? src[v] : (T)0; }
        issued++;
    }
    // fill the pipeline: NS - 1 tiles in flight; afterwards call prefetch() once per consumed tile, before wait()
    __device__ __forceinline__ void prime()
    {
#pragma unroll
        for (int i = 0; i < NS - 1; i++) prefetch();
    }
    // wait for `tile`, return this lane's row (element k of the tile at [k])
    __device__ __forceinline__ const T *wait(int tile)
    {
        const int b = tile % NS;
        jt_mbar_wait(&bar[b], (unsigned)((tile / NS) & 1));
        __syncwarp();
        return buf + ((size_t)b * 32 + (threadIdx.x & 31)) * ROW + shift;
    }
    __device__ __forceinline__ void release() { __syncwarp(); }
};

// The mirror image for results: a lane appends values to its shared-memory row; full rows leave
// with one TMA bulk store (cp.async.bulk.global.shared::cta, per-thread bulk group), so the lane's
// strided 4/8-byte stores become 256/512-byte bursts.  Two rows per lane: one fills while the
// other drains.
template <class T, int R>
struct LaneStore {
    static_assert((R * sizeof(T)) % 16 == 0, "row must be a multiple of 16 bytes");
    static constexpr int ROW = R + 16 / (int)sizeof(T);
    static constexpr size_t WARP_BYTES = 2 * 32 * (size_t)ROW * sizeof(T);
    T *rows;            // this lane's two rows: rows + b * 32 * ROW
    T *dst;             // next global element to write
    int filled, b;

    __device__ __forceinline__ void init(unsigned char *smem, T *lane_dst)
    {
        rows = (T *)smem + (size_t)(threadIdx.x & 31) * ROW;
        dst = lane_dst; filled = 0; b = 0;
    }
    __device__ __forceinline__ T *row() const { return rows + (size_t)b * 32 * ROW; }
    __device__ __forceinline__ void flush()
    {
        if (!filled) return;
        T *r = row();
        if (filled == R && (((uintptr_t)dst) & 15) == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(dst), "r"(jt_smem_u32(r)), "r"((unsigned)(R * sizeof(T))) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            b ^= 1;
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");    // the row we switch to has drained
        } else {
            for (int i = 0; i < filled; i++) dst[i] = r[i];
        }
        dst += filled; filled = 0;
    }
    // the caller wrote n elements straight into row(): hand the row over
    __device__ __forceinline__ void commit(int n) { filled = n; flush(); }
    __device__ __forceinline__ void put(T v)
    {
        row()[filled++] = v;
        if (filled == R) flush();
    }
    __device__ __forceinline__ void finish()
    {
        flush();
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
};
