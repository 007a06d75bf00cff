// Tensor-map TMA staging for "sequential lane" kernels (sm_100a) -- the successor of the per-lane bulk copies in jt_lanes.cuh.
//
// A lane kernel cuts a stream into equal segments of `seg` samples, one thread ("lane") each; the 32 lanes of a warp walk 32
// segments in lock step, starting `warm` samples early.  The addresses a warp needs at one step are therefore 32 rows of a
// matrix whose row pitch is `seg`: the stream, seen as a 3-D tensor
//       dim0 = the 128 / sizeof(T) elements of a 128-byte line            (stride sizeof(T))
//       dim1 = segment index (row)                                         (stride seg * sizeof(T))
//       dim2 = 128-byte line within the segment                            (stride 128 B)
// so ONE cp.async.bulk.tensor.3d (SASS UTMALDG, issued by one elected lane, completion on an mbarrier) brings the next
// CH lines of all 32 lanes: box (128 B, 32 rows, CH lines) -> shared memory [line][row][128 B] with the hardware's 128-byte
// swizzle, which puts the 16-byte chunk a lane reads at step k on a different bank group for each lane (a quarter warp's
// LDS.128 covers all 32 banks exactly once).  The warm-up walks columns c < 0 of a lane's own row, i.e. the tail of earlier
// rows: with 128-byte lines dividing both `seg` and `warm`, a tile never straddles two rows, so the box simply sits at
// (row0 + floor(c / seg), (c mod seg) / EPL).  Rows before the stream (and past its last full row) are out of bounds for the
// tensor map and arrive as zeros -- the rest state a lane starts from.  Results leave the same way: lanes fill a swizzled
// tile, one UTMASTG per tile stores 32 row segments, clipped to the tensor by the hardware.
// The per-lane bulk copies this replaces cost a 32-iteration issue loop per tile (UBLKCP takes uniform registers: ELECT +
// R2UR + UBLKCP per lane, ~35 % of the envelope kernel's stall samples, profiles/ncu_r2g) and 1 KB rows per lane per stage.
//
// What the tensor map cannot express is the ragged last row (n is not a multiple of seg): the lane that owns it patches its
// row of a landed tile with plain loads, and stores its results with plain stores.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "jt_lanes.cuh"

__device__ __forceinline__ void jt_tma_load_3d(void *dst_smem, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(jt_smem_u32(dst_smem)), "l"(map), "r"(jt_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void jt_tma_store_3d(const CUtensorMap *map, int c0, int c1, int c2, const void *src_smem)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(jt_smem_u32(src_smem)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// byte offset of 16-byte chunk `chunk` (0..7) of 128-byte line `line` of lane `lane` inside a [line][32][128 B] tile
__device__ __forceinline__ int jt_tile_off(int line, int lane, int chunk) { return line * 4096 + lane * 128 + ((chunk ^ (lane & 7)) << 4); }

// Input side.  T = element type, CH = 128-byte lines per lane per tile, NS = ring depth.  One instance per warp.
// Everything per tile is 32-bit and incremental: the walk is latency bound with one warp per scheduler, so a 64-bit division
// per tile (the first version: 300 of 630 instructions per tile) costs as much as the arithmetic it feeds.
template <class T, int CH, int NS>
struct LaneTileIn {
    static constexpr int EPL = 128 / (int)sizeof(T);         // elements per line
    static constexpr int R = CH * EPL;                        // elements per lane per tile
    static constexpr int TILE_BYTES = CH * 4096;
    static constexpr size_t WARP_BYTES = (size_t)NS * TILE_BYTES;          // tiles (1024-aligned); the NS barriers live elsewhere

    unsigned char *buf; uint64_t *bar;
    const CUtensorMap *map;
    const T *own;                          // this lane's own segment when it is the ragged last row, else nullptr
    int own_valid;                         // elements of that row inside the stream
    int pf_row, pf_line, lines_per_seg;    // box position of the next tile to request (lane 0 only)
    int issued, ntiles, own_tile0;         // own_tile0: first tile of the own segment (= warm / R)
    unsigned phase_bits;                   // parity of each ring slot

    // smem: WARP_BYTES, 1024-byte aligned; bars: NS mbarriers.  All 32 lanes must call.  warm and seg are multiples of R.
    __device__ __forceinline__ void init(unsigned char *smem, uint64_t *bars, const CUtensorMap *m, const T *stream, int64_t n_total, int64_t seg,
                                         int64_t warm, int64_t warp_row0, int64_t rows_full)
    {
        const int lane = threadIdx.x & 31;
        buf = smem; bar = bars; map = m;
        const int64_t row = warp_row0 + lane;
        own = (row >= rows_full && row * seg < n_total) ? stream + row * seg : nullptr;
        own_valid = own ? (int)min((int64_t)seg, n_total - row * seg) : 0;
        lines_per_seg = (int)(seg / EPL);
        const int64_t back = (warm + seg - 1) / seg;                      // rows the warm-up reaches back
        pf_row = (int)(warp_row0 - back); pf_line = (int)((back * seg - warm) / EPL);
        issued = 0; ntiles = (int)((warm + seg) / R); own_tile0 = (int)(warm / R); phase_bits = 0;
        if (lane < NS) jt_mbar_init(&bar[lane], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
    }
    __device__ __forceinline__ void prefetch()
    {
        if (issued >= ntiles) return;
        if ((threadIdx.x & 31) == 0) {
            const int b = issued % NS;
            jt_mbar_expect_tx(&bar[b], TILE_BYTES);
            jt_tma_load_3d(buf + b * TILE_BYTES, map, 0, pf_row, pf_line, &bar[b]);
            pf_line += CH;
            if (pf_line >= lines_per_seg) { pf_line = 0; pf_row++; }
        }
        issued++;
    }
    __device__ __forceinline__ void prime()
    {
#pragma unroll
        for (int i = 0; i < NS - 1; i++) prefetch();
    }
    // wait for `tile`; returns the tile's base.  The lane owning the ragged last row patches its row from global memory.
    __device__ __forceinline__ const unsigned char *wait(int tile)
    {
        const int b = tile % NS;
        jt_mbar_wait(&bar[b], (phase_bits >> b) & 1u);
        phase_bits ^= 1u << b;
        unsigned char *t = buf + b * TILE_BYTES;
        if (own != nullptr && tile >= own_tile0) {                        // at most one lane of the grid
            const int lane = threadIdx.x & 31, base = (tile - own_tile0) * R;
            for (int k = 0; k < R; k++) {
                const int line = k / EPL, e = k % EPL;
                *(T *)(t + jt_tile_off(line, lane, (e * (int)sizeof(T)) >> 4) + ((e * (int)sizeof(T)) & 15)) = base + k < own_valid ? own[base + k] : (T)0;
            }
        }
        __syncwarp();
        return t;
    }
    __device__ __forceinline__ void release() { __syncwarp(); }
};

// Output side: two tiles per warp, one fills while the other drains.
template <class T, int CH>
struct LaneTileOut {
    static constexpr int EPL = 128 / (int)sizeof(T);
    static constexpr int R = CH * EPL;
    static constexpr int TILE_BYTES = CH * 4096;
    static constexpr size_t WARP_BYTES = 2 * (size_t)TILE_BYTES;
    unsigned char *buf; const CUtensorMap *map; T *own; int own_valid, row0, line, b;

    __device__ __forceinline__ void init(unsigned char *smem, const CUtensorMap *m, T *stream, int64_t n_total, int64_t seg, int64_t warp_row0, int64_t rows_full)
    {
        buf = smem; map = m; row0 = (int)warp_row0; line = 0; b = 0;
        const int64_t row = warp_row0 + (threadIdx.x & 31);
        own = (row >= rows_full && row * seg < n_total) ? stream + row * seg : nullptr;
        own_valid = own ? (int)min((int64_t)seg, n_total - row * seg) : 0;
    }
    __device__ __forceinline__ unsigned char *tile() const { return buf + b * TILE_BYTES; }
    // the warp has filled tile() with the next R columns of its rows (tiles are committed in order, from column 0)
    __device__ __forceinline__ void commit()
    {
        const int lane = threadIdx.x & 31;
        unsigned char *t = tile();
        if (own != nullptr) {                                              // ragged last row: plain stores by its lane
            for (int k = 0; k < R; k++) {
                const int ln = k / EPL, e = k % EPL, v = line * EPL + k;
                if (v < own_valid) own[v] = *(const T *)(t + jt_tile_off(ln, lane, (e * (int)sizeof(T)) >> 4) + ((e * (int)sizeof(T)) & 15));
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            jt_tma_store_3d(map, 0, row0, line, t);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the tile we switch to has been read out
        }
        line += CH; b ^= 1;
        __syncwarp();
    }
    __device__ __forceinline__ void finish()
    {
        if ((threadIdx.x & 31) == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
    }
};

// host: tensor map of a stream of n elements of elem_bytes seen as rows of `seg` elements (full rows only); false when the
// stream cannot be described (fewer than one full row, base not 16-byte aligned, seg not a multiple of the 128-byte line)
bool jt_lane_tensor_map(CUtensorMap *map, const void *base, int elem_bytes, int64_t n, int64_t seg, int lines_per_tile);
