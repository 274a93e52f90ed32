// Row selection for the tensor-bound dense path (K2T): under a selective filter the GEMM would spend most of its
// flops on rows the filter drops (vector_store.py:462-530 builds ONE filter for both query_points calls; at BASELINE
// config 4 it passes 50 % of the rows).  When every query of the batch shares the filter, the rows that pass are
// copied ONCE per batch into a contiguous scratch matrix (coalesced 16-byte loads / stores, HBM-bound: read + write of
// the passing rows) and K2T runs over that matrix instead: half the tiles at 50 %, a hundredth at 1 %.
//   vb_rowsel_count_kernel    popcount of the filter words, one sum per 32768-row block
//   vb_rowsel_scatter_kernel  exclusive offsets (block sums + in-block scan) -> ascending row ids; count, decision
//   vb_rowsel_gather_kernel   scratch[p] = rows[ids[p]], inv_norm_c[p] = inv_norm[ids[p]] (skipped when the decision is "no")
// The decision is taken ON THE DEVICE (count * 100 <= pct * rows): no host round trip; K2T reads it and either walks
// the scratch matrix (row id of a candidate = ids[position]) or the rows in place with the filter bit in its epilogue.
// (TMA tile::gather4 was measured first — tools/gather4_probe.cu: box {64, 1}, four rows per operation land as half a
// SWIZZLE_128B atom, but 32 operations per 128-row tile sustain 2.2 TB/s chip-wide against 6.3 TB/s for one box per
// tile, and K2T re-reads a corpus tile once per query tile: the gather would bound the kernel, the copy does not.)
#pragma once
#include "common.cuh"

#define VB_SEL_THREADS 256u
#define VB_SEL_WORDS 1024u                 // filter words per block: 4 per thread, 32768 rows

struct VbRowSelArgs {
    const uint32_t* mask;                  // the shared filter's words for the whole shard
    uint32_t word_begin, word_end;         // words of the segment: [row_begin / 32, ceil(row_end / 32)); word_begin % 4 == 0
    uint32_t row_end;                      // bits at or above row_end are ignored
    uint32_t* block_sums;                  // [blocks]
    uint32_t* ids;                         // [round_up(rows, 128)] ascending ids of the passing rows, tail padded
    uint32_t* sel;                         // [0] passing rows, [1] 1 = use the compacted copy
    uint32_t pct;                          // compact when passing * 100 <= pct * rows of the segment
    uint32_t pad_row;                      // a valid row id for the padding slots (they are never scored)
    uint32_t cap_rows;                     // slots of `ids` (debug build: every write is checked against it)
};

__device__ __forceinline__ uint4 vb_rowsel_words(const VbRowSelArgs& a, uint32_t w0) {
    uint4 w = make_uint4(0u, 0u, 0u, 0u);
    if (w0 < a.word_end) {
        // (the filter's words start at mask + f * mask_words: 16-byte aligned only for f = 0 or a word count divisible by 4)
        if (w0 + 4u <= a.word_end && (reinterpret_cast<uintptr_t>(a.mask + w0) & 15u) == 0u) w = __ldg(reinterpret_cast<const uint4*>(a.mask + w0));
        else {
            w.x = a.mask[w0];
            if (w0 + 1u < a.word_end) w.y = a.mask[w0 + 1u];
            if (w0 + 2u < a.word_end) w.z = a.mask[w0 + 2u];
            if (w0 + 3u < a.word_end) w.w = a.mask[w0 + 3u];
        }
        uint32_t* p = &w.x;
#pragma unroll
        for (uint32_t i = 0; i < 4u; ++i) {
            const uint64_t first = (uint64_t)(w0 + i) * 32u;
            if (first >= a.row_end) p[i] = 0u;
            else if (first + 32u > a.row_end) p[i] &= (1u << (a.row_end - (uint32_t)first)) - 1u;
        }
    }
    return w;
}

// sum of v over the block (every thread gets it)
__device__ __forceinline__ uint32_t vb_rowsel_block_sum(uint32_t v, uint32_t* s_warp) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31u) == 0u) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t t = 0;
#pragma unroll
    for (uint32_t i = 0; i < VB_SEL_THREADS / 32u; ++i) t += s_warp[i];
    return t;
}

__global__ void __launch_bounds__(VB_SEL_THREADS)
vb_rowsel_count_kernel(const VbRowSelArgs a)
{
    __shared__ uint32_t s_warp[VB_SEL_THREADS / 32u];
    const uint4 w = vb_rowsel_words(a, a.word_begin + blockIdx.x * VB_SEL_WORDS + threadIdx.x * 4u);
    const uint32_t t = vb_rowsel_block_sum(__popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w), s_warp);
    if (threadIdx.x == 0) a.block_sums[blockIdx.x] = t;
}

__global__ void __launch_bounds__(VB_SEL_THREADS)
vb_rowsel_scatter_kernel(const VbRowSelArgs a)
{
    __shared__ uint32_t s_warp[VB_SEL_THREADS / 32u];
    __shared__ uint32_t s_scan[VB_SEL_THREADS / 32u];
    uint32_t before = 0;                                           // passing rows in the blocks before this one
    for (uint32_t i = threadIdx.x; i < blockIdx.x; i += VB_SEL_THREADS) before += a.block_sums[i];
    before = vb_rowsel_block_sum(before, s_warp);
    const uint32_t w0 = a.word_begin + blockIdx.x * VB_SEL_WORDS + threadIdx.x * 4u;
    const uint4 w = vb_rowsel_words(a, w0);
    const uint32_t mine = __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
    // exclusive scan of `mine` over the block
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += up;
    }
    __syncthreads();
    if (lane == 31u) s_scan[warp] = inc;
    __syncthreads();
    uint32_t warp_before = 0;
#pragma unroll
    for (uint32_t i = 0; i < VB_SEL_THREADS / 32u; ++i) warp_before += i < warp ? s_scan[i] : 0u;
    uint32_t pos = before + warp_before + inc - mine;
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (uint32_t i = 0; i < 4u; ++i) {
        uint32_t bits = ws[i];
        const uint32_t base = (w0 + i) * 32u;
        while (bits) {
            VB_CHECK(pos < a.cap_rows);
            a.ids[pos++] = base + (__ffs(bits) - 1u);
            bits &= bits - 1u;
        }
    }
    if (blockIdx.x == gridDim.x - 1u) {                            // the last block knows the total
        uint32_t block_total = 0;
#pragma unroll
        for (uint32_t i = 0; i < VB_SEL_THREADS / 32u; ++i) block_total += s_scan[i];
        const uint32_t total = before + block_total;
        const uint32_t padded = (total + 127u) & ~127u;
        VB_CHECK(padded <= a.cap_rows);
        for (uint32_t p = total + threadIdx.x; p < padded; p += VB_SEL_THREADS) a.ids[p] = a.pad_row;
        if (threadIdx.x == 0) {
            const uint64_t rows = (uint64_t)a.row_end - (uint64_t)a.word_begin * 32u;
            a.sel[0] = total;
            a.sel[1] = ((uint64_t)total * 100u <= (uint64_t)a.pct * rows) ? 1u : 0u;
        }
    }
}

// one warp per row, four rows in flight per warp; rows [count, round_up(count, 128)) are zero-filled
__global__ void __launch_bounds__(256)
vb_rowsel_gather_kernel(const uint4* __restrict__ rows, const float* __restrict__ inv_norm, const uint32_t* __restrict__ ids,
                        const uint32_t* __restrict__ sel, uint4* __restrict__ out, float* __restrict__ inv_out, uint32_t chunks)
{
    if (sel[1] == 0u) return;
    const uint32_t count = sel[0], padded = (count + 127u) & ~127u;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    constexpr uint32_t R = 4, C = 4;                               // rows per step, 16-byte chunks per lane and row (d_pad <= 1024)
    for (uint32_t p0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * R; p0 < padded; p0 += n_warps * R) {
        uint32_t r[R];
        float nv = 0.0f;
#pragma unroll
        for (uint32_t k = 0; k < R; ++k) r[k] = (p0 + k < padded) ? __ldg(ids + p0 + k) : 0u;     // padded is a multiple of 128
        if (lane < R && p0 + lane < count) {
            const uint32_t rr = lane == 0u ? r[0] : lane == 1u ? r[1] : lane == 2u ? r[2] : r[3];
            nv = __ldg(inv_norm + rr);
        }
        uint4 v[R][C];
#pragma unroll
        for (uint32_t k = 0; k < R; ++k)
#pragma unroll
            for (uint32_t c = 0; c < C; ++c) {
                const uint32_t ch = lane + 32u * c;
                v[k][c] = (ch < chunks && p0 + k < count) ? vb_ldg_stream(rows + (size_t)r[k] * chunks + ch) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
        for (uint32_t k = 0; k < R; ++k)
#pragma unroll
            for (uint32_t c = 0; c < C; ++c) {
                const uint32_t ch = lane + 32u * c;
                if (ch < chunks) out[(size_t)(p0 + k) * chunks + ch] = v[k][c];
            }
        if (lane < R) inv_out[p0 + lane] = nv;
    }
}
