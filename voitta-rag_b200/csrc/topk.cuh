// Exact top-k' selection over a candidate list, and the K4 fusion merge kernel.
//
// vb_compact_kernel replaces the argsort + walk qdrant's local mode does after scoring
// (local_collection.py search(): `order = np.argsort(scores)[::-1]`, skip masked, stop at
// limit).  The scoring kernels append every row whose score beats the list's threshold tau
// (a lower bound on the k'-th best seen so far); this kernel reduces a list to its exact top-k'
// under the total order (score desc, row asc), sorted, and raises tau.  Because keys are
// unique, the result does not depend on the order in which candidates were appended.
//
// vb_fuse_kernel replaces voitta's own fusion (vector_store.py:659-697, min-max weighted sum)
// and adds Qdrant's RRF (qdrant_client/hybrid/fusion.py).  All arithmetic is fp64 with explicit
// round-to-nearest mul/add/div (no FMA contraction) so it reproduces the Python floats bit for bit.
#pragma once
#include "common.cuh"

#define VB_SORT_MAX 2048u
#define VB_COMPACT_THREADS 512

// descending bitonic sort of P (power of two <= VB_SORT_MAX) keys in shared memory
__device__ __forceinline__ void vb_bitonic_desc(uint64_t* s, uint32_t P) {
    for (uint32_t size = 2; size <= P; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (uint32_t t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                const uint32_t i = 2u * t - (t & (stride - 1u));
                const uint32_t j = i + stride;
                const bool up = (i & size) == 0u;
                const uint64_t x = s[i], y = s[j];
                if ((x < y) == up) { s[i] = y; s[j] = x; }
            }
        }
    }
    __syncthreads();
}

// One CTA per list.  Afterwards: cand[list][0..cnt) = top-min(k, #valid) keys sorted descending,
// cnt[list] = that count, tau[list] = score of the k-th (or -inf if fewer than k).
__global__ void __launch_bounds__(VB_COMPACT_THREADS)
vb_compact_kernel(uint64_t* __restrict__ cand, uint32_t* __restrict__ cnt, float* __restrict__ tau,
                  uint32_t* __restrict__ overflow, uint32_t cap, uint32_t k, uint32_t list_begin)
{
    __shared__ uint64_t s_keys[VB_SORT_MAX];
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_nsel;
    __shared__ uint64_t s_prefix;
    __shared__ uint32_t s_need;

    const uint32_t list = list_begin + blockIdx.x;
    uint64_t* keys = cand + (size_t)list * cap;
    const uint32_t raw = cnt[list];
    if (raw > cap && threadIdx.x == 0) overflow[list] = 1u;
    const uint32_t E = raw < cap ? raw : cap;

    uint32_t nsel;
    if (E <= VB_SORT_MAX) {
        for (uint32_t i = threadIdx.x; i < E; i += blockDim.x) s_keys[i] = keys[i];
        nsel = E;
    } else {
        // MSB-first radix select of the k-th largest key (8 bits per pass over 64-bit keys)
        if (threadIdx.x == 0) { s_prefix = 0ull; s_need = k; s_nsel = 0u; }
        uint64_t pmask = 0ull;
        for (int shift = 56; shift >= 0; shift -= 8) {
            if (threadIdx.x < 256) s_hist[threadIdx.x] = 0u;
            __syncthreads();
            const uint64_t prefix = s_prefix;
            for (uint32_t i = threadIdx.x; i < E; i += blockDim.x) {
                const uint64_t key = keys[i];
                if ((key & pmask) == prefix) atomicAdd(&s_hist[(uint32_t)(key >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t need = s_need, above = 0u;
                int d = 255;
                for (; d > 0; --d) {
                    if (above + s_hist[d] >= need) break;
                    above += s_hist[d];
                }
                s_need = need - above;           // rank inside bucket d
                s_prefix = prefix | ((uint64_t)d << shift);
            }
            pmask |= 0xffull << shift;
            __syncthreads();
        }
        const uint64_t pivot = s_prefix;         // the k-th largest key (0 if fewer than k valid)
        for (uint32_t i = threadIdx.x; i < E; i += blockDim.x) {
            const uint64_t key = keys[i];
            if (key != 0ull && key >= pivot) {
                const uint32_t slot = atomicAdd(&s_nsel, 1u);
                if (slot < VB_SORT_MAX) s_keys[slot] = key;
            }
        }
        __syncthreads();
        nsel = s_nsel < VB_SORT_MAX ? s_nsel : VB_SORT_MAX;
    }
    uint32_t P = 1;
    while (P < nsel) P <<= 1;
    if (P < 2) P = 2;
    for (uint32_t i = nsel + threadIdx.x; i < P; i += blockDim.x) s_keys[i] = 0ull;
    vb_bitonic_desc(s_keys, P);

    const uint32_t keep = nsel < k ? nsel : k;
    for (uint32_t i = threadIdx.x; i < keep; i += blockDim.x) keys[i] = s_keys[i];
    if (threadIdx.x == 0) {
        uint32_t valid = keep;                   // zero keys (padding) sort last
        while (valid > 0 && s_keys[valid - 1] == 0ull) --valid;
        cnt[list] = valid;
        tau[list] = (valid >= k) ? vb_key_score(s_keys[k - 1]) : -INFINITY;
    }
}

// Pack the first k keys of every list into out[list][k] (0-padded): the all-gather payload.
// If `overflow` is given, one extra word out[n_lists*k] carries "some list of this shard
// overflowed", so every rank learns it from the same all-gather and re-runs consistently.
__global__ void vb_export_kernel(const uint64_t* __restrict__ cand, const uint32_t* __restrict__ cnt,
                                 uint32_t cap, uint32_t k, uint64_t* __restrict__ out,
                                 const uint32_t* __restrict__ overflow, uint32_t n_lists)
{
    const uint32_t list = blockIdx.x;
    const uint32_t c = cnt[list] < k ? cnt[list] : k;
    for (uint32_t i = threadIdx.x; i < k; i += blockDim.x)
        out[(size_t)list * k + i] = i < c ? cand[(size_t)list * cap + i] : 0ull;
    if (overflow != nullptr && blockIdx.x == 0) {
        int any = 0;
        for (uint32_t i = threadIdx.x; i < n_lists; i += blockDim.x) any |= overflow[i] != 0u;
        any = __syncthreads_or(any);
        if (threadIdx.x == 0) out[(size_t)n_lists * k] = any ? 1ull : 0ull;
    }
}

// Scatter gathered[shard][list][k] (+1 flag word per shard) into cand[list][shard*k + i] and set
// cnt = n_shards*k (empty slots are 0 keys, which the compaction ignores).
__global__ void vb_import_kernel(const uint64_t* __restrict__ gathered, uint32_t n_shards, uint32_t n_lists,
                                 uint32_t k, uint32_t cap, uint64_t* __restrict__ cand, uint32_t* __restrict__ cnt,
                                 uint32_t* __restrict__ overflow)
{
    const uint32_t list = blockIdx.x;
    const size_t stride = (size_t)n_lists * k + 1;
    for (uint32_t i = threadIdx.x; i < n_shards * k; i += blockDim.x) {
        const uint32_t sh = i / k, j = i % k;
        cand[(size_t)list * cap + i] = gathered[sh * stride + (size_t)list * k + j];
    }
    if (threadIdx.x == 0) {
        cnt[list] = n_shards * k;
        if (list == 0)
            for (uint32_t sh = 0; sh < n_shards; ++sh)
                if (gathered[sh * stride + (size_t)n_lists * k] != 0ull) overflow[0] = 1u;
    }
}

// ---------------------------------------------------------------------------------------------
// K4 fusion.  One CTA per query.  Dense list = list q, sparse list = list B+q, both already
// sorted (score desc).  Output: rows + fp64 scores in rank order, count.
//   mode 0: dense only (vector_store.py:611-619): first `limit` dense hits, score = cosine
//   mode 1: weighted   (vector_store.py:659-697)
//   mode 2: rrf        (qdrant hybrid/fusion.py, ranking constant 2)
// Ties of the fused score keep first-seen order (dense list, then sparse-only ids).
// ---------------------------------------------------------------------------------------------
struct VbFuseArgs {
    const uint64_t* cand;
    const uint32_t* cnt;
    const int32_t* mode;      // [B]
    uint32_t cap, n_queries, k, limit;
    double w_sparse, w_dense;
    uint32_t* out_rows;       // [B][limit]
    double* out_scores;       // [B][limit]
    int32_t* out_cnt;         // [B]
};

__global__ void __launch_bounds__(128)
vb_fuse_kernel(const VbFuseArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t k = a.k;
    double* fin = reinterpret_cast<double*>(smem_raw);          // [2k] fused score
    uint32_t* row = reinterpret_cast<uint32_t*>(fin + 2 * k);   // [2k]
    uint8_t* valid = reinterpret_cast<uint8_t*>(row + 2 * k);   // [2k]

    const uint32_t q = blockIdx.x, B = a.n_queries;
    const uint64_t* dk = a.cand + (size_t)q * a.cap;
    const uint64_t* sk = a.cand + (size_t)(B + q) * a.cap;
    const uint32_t nd = a.cnt[q] < k ? a.cnt[q] : k;
    const int32_t mode = a.mode[q];

    if (mode == 0) {
        const uint32_t n = nd < a.limit ? nd : a.limit;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            a.out_rows[(size_t)q * a.limit + i] = vb_key_row(dk[i]);
            a.out_scores[(size_t)q * a.limit + i] = (double)vb_key_score(dk[i]);
        }
        if (threadIdx.x == 0) a.out_cnt[q] = (int32_t)n;
        return;
    }
    const uint32_t ns = a.cnt[B + q] < k ? a.cnt[B + q] : k;

    double dmin = 0.0, dspread = 0.0, smin = 0.0, sspread = 0.0;
    if (nd) { dmin = (double)vb_key_score(dk[nd - 1]); dspread = __dsub_rn((double)vb_key_score(dk[0]), dmin); }
    if (ns) { smin = (double)vb_key_score(sk[ns - 1]); sspread = __dsub_rn((double)vb_key_score(sk[0]), smin); }

    for (uint32_t i = threadIdx.x; i < 2 * k; i += blockDim.x) valid[i] = 0;
    __syncthreads();

    // dense candidates: slot i
    for (uint32_t i = threadIdx.x; i < nd; i += blockDim.x) {
        const uint32_t r = vb_key_row(dk[i]);
        int j = -1;
        for (uint32_t t = 0; t < ns; ++t) if (vb_key_row(sk[t]) == r) { j = (int)t; break; }
        double f;
        if (mode == 1) {
            const double dn = dspread > 0.0 ? __ddiv_rn(__dsub_rn((double)vb_key_score(dk[i]), dmin), dspread) : 1.0;
            const double sn = j < 0 ? 0.0
                              : (sspread > 0.0 ? __ddiv_rn(__dsub_rn((double)vb_key_score(sk[j]), smin), sspread) : 1.0);
            f = __dadd_rn(__dmul_rn(a.w_dense, dn), __dmul_rn(a.w_sparse, sn));
        } else {
            f = __ddiv_rn(1.0, (double)(2u + i));
            if (j >= 0) f = __dadd_rn(f, __ddiv_rn(1.0, (double)(2u + (uint32_t)j)));
        }
        fin[i] = f; row[i] = r; valid[i] = 1;
    }
    // sparse-only candidates: slot k + j
    for (uint32_t j = threadIdx.x; j < ns; j += blockDim.x) {
        const uint32_t r = vb_key_row(sk[j]);
        bool in_dense = false;
        for (uint32_t t = 0; t < nd; ++t) if (vb_key_row(dk[t]) == r) { in_dense = true; break; }
        if (in_dense) continue;
        double f;
        if (mode == 1) {
            const double sn = sspread > 0.0 ? __ddiv_rn(__dsub_rn((double)vb_key_score(sk[j]), smin), sspread) : 1.0;
            f = __dadd_rn(__dmul_rn(a.w_dense, 0.0), __dmul_rn(a.w_sparse, sn));
        } else {
            f = __ddiv_rn(1.0, (double)(2u + j));
        }
        fin[k + j] = f; row[k + j] = r; valid[k + j] = 1;
    }
    __syncthreads();

    // rank by counting: better(c', c) = f' > f or (f' == f and slot' < slot)
    uint32_t total = 0;
    for (uint32_t c = threadIdx.x; c < 2 * k; c += blockDim.x) {
        if (!valid[c]) continue;
        const double f = fin[c];
        uint32_t rank = 0;
        for (uint32_t o = 0; o < 2 * k; ++o) {
            if (!valid[o]) continue;
            const double g = fin[o];
            rank += (g > f || (g == f && o < c)) ? 1u : 0u;
        }
        if (rank < a.limit) {
            a.out_rows[(size_t)q * a.limit + rank] = row[c];
            a.out_scores[(size_t)q * a.limit + rank] = f;
        }
        ++total;
    }
    // total candidates = nd + sparse-only; count via block reduction
    __shared__ uint32_t s_total;
    if (threadIdx.x == 0) s_total = 0;
    __syncthreads();
    atomicAdd(&s_total, total);
    __syncthreads();
    if (threadIdx.x == 0) a.out_cnt[q] = (int32_t)(s_total < a.limit ? s_total : a.limit);
}
