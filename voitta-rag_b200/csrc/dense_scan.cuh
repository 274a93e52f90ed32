// K1 — single-query / small-batch dense scoring: a bandwidth-bound GEMV over the bf16 corpus
// with the filter bitmask and the top-k' threshold fused into the row loop.
// Replaces qdrant's cosine scoring for query_points(query=vec, limit=k', query_filter)
// (vector_store.py:612-617, :640-645; distances.py cosine_similarity in local mode).
//
//   - one warp owns a 32-row group (= one mask word); masked rows are never loaded, so at 1 %
//     selectivity only 1 % of the row bytes move;
//   - a row is read with coalesced 128-bit streaming loads (lane l reads 16-byte chunks
//     l, l+32, ...), 4 rows in flight per warp for memory-level parallelism;
//   - the query (fp32, unit norm) lives in registers; fp32 FMA accumulate, xor-butterfly reduce;
//   - score = dot * inv_norm[row]; rows beating the list's threshold tau are appended to the
//     candidate list (vb_push); the exact top-k' is selected by vb_compact_kernel.
// Roofline: HBM.  Algorithmic bytes per launch = P*(d_pad*2 + 4) + groups*4 (mask words),
// P = rows that pass the mask.
#pragma once
#include "common.cuh"

struct VbScanArgs {
    const uint4* rows;          // [n][d_pad] bf16 as 16-byte chunks
    const float* inv_norm;      // [n]
    const uint32_t* mask;       // [n_filters][mask_words] or nullptr
    const int32_t* mask_of;     // [B] filter index per query, -1 = no mask; nullptr = none
    const float* q_hat;         // [B][d_pad] fp32 unit-norm queries
    const float* tau;           // [n_lists]
    VbLists lists;
    uint32_t mask_words;
    uint32_t chunks;            // d_pad / 8
    uint32_t row_begin, row_end;// segment (row_begin % 32 == 0)
    uint32_t row_base;          // added to the row in the candidate key (shard offset)
    uint32_t q_begin;           // first query handled by blockIdx.y == 0
    uint32_t direct;            // 1: first segment — store the key at slot (row - row_begin), no atomics
};

__device__ __forceinline__ float vb_dot8(const uint4 v, const float* q, float acc) {
    acc = fmaf(__uint_as_float(v.x << 16), q[0], acc);
    acc = fmaf(__uint_as_float(v.x & 0xffff0000u), q[1], acc);
    acc = fmaf(__uint_as_float(v.y << 16), q[2], acc);
    acc = fmaf(__uint_as_float(v.y & 0xffff0000u), q[3], acc);
    acc = fmaf(__uint_as_float(v.z << 16), q[4], acc);
    acc = fmaf(__uint_as_float(v.z & 0xffff0000u), q[5], acc);
    acc = fmaf(__uint_as_float(v.w << 16), q[6], acc);
    acc = fmaf(__uint_as_float(v.w & 0xffff0000u), q[7], acc);
    return acc;
}

// NCH = ceil(chunks / 32) register-resident query chunks per lane (d_pad <= NCH*256).
template <int NCH>
__global__ void __launch_bounds__(256)
vb_dense_scan_kernel(const VbScanArgs a)
{
    constexpr int ROWS = 4;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t q_idx = a.q_begin + blockIdx.y;
    const uint32_t list = q_idx;  // dense lists come first
    const float tau = a.tau[list];
    const uint32_t* mask = nullptr;
    if (a.mask != nullptr && a.mask_of != nullptr) {
        const int32_t f = a.mask_of[q_idx];
        if (f >= 0) mask = a.mask + (size_t)f * a.mask_words;
    }

    float q[NCH][8];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const uint32_t ch = lane + 32u * c;
#pragma unroll
        for (int e = 0; e < 8; ++e)
            q[c][e] = ch < a.chunks ? a.q_hat[(size_t)q_idx * a.chunks * 8u + ch * 8u + e] : 0.0f;
    }

    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    const uint32_t g_begin = a.row_begin >> 5, g_end = (a.row_end + 31u) >> 5;
    for (uint32_t g = g_begin + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < g_end; g += warps) {
        uint32_t bits = mask ? mask[g] : 0xffffffffu;
        const uint32_t row0 = g << 5;
        if (row0 + 32u > a.row_end) bits &= (1u << (a.row_end - row0)) - 1u;
        if (bits == 0u) continue;
        const float invn = (row0 + lane < a.row_end) ? a.inv_norm[row0 + lane] : 0.0f;
        while (bits) {
            int r[ROWS];
#pragma unroll
            for (int k = 0; k < ROWS; ++k) {
                r[k] = bits ? (__ffs(bits) - 1) : -1;
                bits &= bits - 1u;
            }
            uint4 v[ROWS][NCH];
#pragma unroll
            for (int k = 0; k < ROWS; ++k) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const uint32_t ch = lane + 32u * c;
                    v[k][c] = (r[k] >= 0 && ch < a.chunks)
                                  ? vb_ldg_stream(a.rows + (size_t)(row0 + r[k]) * a.chunks + ch)
                                  : make_uint4(0u, 0u, 0u, 0u);
                }
            }
#pragma unroll
            for (int k = 0; k < ROWS; ++k) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) vb_keep_loaded(v[k][c]);
            }
            float acc[ROWS];
#pragma unroll
            for (int k = 0; k < ROWS; ++k) {
                acc[k] = 0.0f;
#pragma unroll
                for (int c = 0; c < NCH; ++c) acc[k] = vb_dot8(v[k][c], q[c], acc[k]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int k = 0; k < ROWS; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
            }
            // lane k finishes row r[k]
            float mine = acc[0];
            int myr = r[0];
#pragma unroll
            for (int k = 1; k < ROWS; ++k) {
                if (lane == (uint32_t)k) { mine = acc[k]; myr = r[k]; }
            }
            const float inv_r = __shfl_sync(0xffffffffu, invn, myr < 0 ? 0 : myr);
            if (lane < (uint32_t)ROWS && myr >= 0) {
                const float s = mine * inv_r;
                const uint32_t row = row0 + (uint32_t)myr;
                if (a.direct) a.lists.cand[(size_t)list * a.lists.cap + (row - a.row_begin)] = s > tau ? vb_pack_key(s, a.row_base + row) : 0ull;
                else if (s > tau) vb_push(a.lists, list, s, a.row_base + row);
            }
        }
    }
}

// Generic variant for d_pad > 1024: the query is re-read through L1 per chunk.
__global__ void __launch_bounds__(256)
vb_dense_scan_generic_kernel(const VbScanArgs a)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t q_idx = a.q_begin + blockIdx.y;
    const uint32_t list = q_idx;
    const float tau = a.tau[list];
    const uint32_t* mask = nullptr;
    if (a.mask != nullptr && a.mask_of != nullptr) {
        const int32_t f = a.mask_of[q_idx];
        if (f >= 0) mask = a.mask + (size_t)f * a.mask_words;
    }
    const float* qv = a.q_hat + (size_t)q_idx * a.chunks * 8u;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    const uint32_t g_begin = a.row_begin >> 5, g_end = (a.row_end + 31u) >> 5;
    for (uint32_t g = g_begin + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < g_end; g += warps) {
        uint32_t bits = mask ? mask[g] : 0xffffffffu;
        const uint32_t row0 = g << 5;
        if (row0 + 32u > a.row_end) bits &= (1u << (a.row_end - row0)) - 1u;
        while (bits) {
            const uint32_t r = __ffs(bits) - 1;
            bits &= bits - 1u;
            float acc = 0.0f;
            for (uint32_t ch = lane; ch < a.chunks; ch += 32u) {
                const uint4 v = vb_ldg_stream(a.rows + (size_t)(row0 + r) * a.chunks + ch);
                float qq[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) qq[e] = __ldg(qv + ch * 8u + e);
                acc = vb_dot8(v, qq, acc);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) {
                const float s = acc * a.inv_norm[row0 + r];
                const uint32_t row = row0 + r;
                if (a.direct) a.lists.cand[(size_t)list * a.lists.cap + (row - a.row_begin)] = s > tau ? vb_pack_key(s, a.row_base + row) : 0ull;
                else if (s > tau) vb_push(a.lists, list, s, a.row_base + row);
            }
        }
    }
}
