// K3 — sparse (BM25-style) scoring over an inverted index staged through shared memory.
// Replaces qdrant's sparse dot product with the IDF modifier for
// query_points(query=SparseVector, using="bm25", limit=k', query_filter)
// (vector_store.py:647-656; sparse_distances.py sparse_dot_product in local mode).
//
// Index layout (built on the device from the row-major CSR the upserts append):
//   postings sorted by (term, row):  post_row[nnz] u32, post_val[nnz] f32
//   terms_sorted[T] u32 (distinct hashed term ids), term_ptr[T+1] u64
// Rows are cut into blocks of VB_ROWS_PER_BLOCK rows.  For a batch, vb_slice_kernel finds,
// for every query term and every row block, the slice of that term's postings that falls in
// the block (binary search on post_row).  vb_sparse_kernel then gives one CTA a (row block,
// query) pair: fp64 accumulators for the block's rows live in shared memory, the query's terms
// are walked in ascending term-id order and each slice is streamed with coalesced loads.
// Inside one term a row occurs at most once, so no atomics are needed and the summation
// order (ascending term id, explicit mul then add in fp64, one final rounding to fp32) is
// exactly the order of the reference's two-pointer merge: sparse scores are bit-identical to
// np.float32(sum of python floats).  Rows never touched keep a sentinel: "no shared index"
// rows are excluded, as in the reference.
// Roofline: HBM (L2 for slices shared by queries of the batch).
// Algorithmic bytes = sum over queries, terms: df_shard(term) * 8.
#pragma once
#include "common.cuh"
#include <cub/cub.cuh>

#define VB_ACC_SENTINEL 0x7ff8dead00000000ull   // a NaN payload no sum can produce

// ---- index build ------------------------------------------------------------------------------
// key = term << 32 | row for live rows, ~0 for tombstoned rows (sorted to the tail, then cut).
__global__ void __launch_bounds__(256)
vb_posting_keys_kernel(const int64_t* __restrict__ indptr, const uint32_t* __restrict__ term,
                       const uint32_t* __restrict__ alive, uint32_t n_rows, uint64_t* __restrict__ keys)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += warps) {
        const int64_t lo = indptr[r], hi = indptr[r + 1];
        const bool live = (alive[r >> 5] >> (r & 31u)) & 1u;
        for (int64_t p = lo + lane; p < hi; p += 32)
            keys[p] = live ? (((uint64_t)term[p] << 32) | r) : ~0ull;
    }
}

// After the sort: split keys into post_row, flag term starts.
__global__ void __launch_bounds__(256)
vb_posting_split_kernel(const uint64_t* __restrict__ keys, uint64_t nnz, uint32_t* __restrict__ post_row,
                        uint32_t* __restrict__ post_term)
{
    for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < nnz; p += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t k = keys[p];
        post_row[p] = (uint32_t)k;
        post_term[p] = (uint32_t)(k >> 32);
    }
}

// ---- per-batch slice table ----------------------------------------------------------------------
// off[j][b] = first posting of query-term j whose row >= b*VB_ROWS_PER_BLOCK  (b = 0..n_blocks)
__global__ void __launch_bounds__(256)
vb_slice_kernel(const uint32_t* __restrict__ post_row, const uint64_t* __restrict__ qt_lo,
                const uint64_t* __restrict__ qt_hi, uint32_t n_qterms, uint32_t n_blocks,
                uint64_t* __restrict__ off)
{
    const uint64_t total = (uint64_t)n_qterms * (n_blocks + 1u);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t j = (uint32_t)(i / (n_blocks + 1u));
        const uint32_t b = (uint32_t)(i % (n_blocks + 1u));
        uint64_t lo = qt_lo[j], hi = qt_hi[j];
        const uint64_t target = (uint64_t)b * VB_ROWS_PER_BLOCK;
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if ((uint64_t)post_row[mid] < target) lo = mid + 1; else hi = mid;
        }
        off[i] = lo;
    }
}

struct VbSparseArgs {
    const uint32_t* post_row;
    const float* post_val;
    const uint64_t* off;        // [n_qterms][n_blocks+1]
    const int64_t* q_indptr;    // [B+1] into the sorted query-term arrays
    const double* q_weight;     // [n_qterms] idf-scaled query values, ascending term id per query
    const uint32_t* mask;       // [n_filters][mask_words] or nullptr
    const int32_t* mask_of;     // [B] or nullptr
    const float* tau;
    uint64_t* cand;
    uint32_t* cnt;
    uint32_t mask_words;
    uint32_t n_blocks;          // row blocks in the whole index
    uint32_t blk_begin;         // first row block of this segment
    uint32_t n_queries;
    uint32_t n_rows;
    uint32_t row_base;
    uint32_t cap;
    uint32_t direct;            // 1: first segment — store keys at slot (row - segment begin), no atomics
};

#define VB_SPARSE_THREADS 128
#define VB_SPARSE_UNROLL 4

// grid.x = (#blocks in segment) * B ; CTA (blk, q) with q fastest so that concurrently running
// CTAs share a row block (and the posting slices of common terms hit L2).  16 KB of shared
// memory per CTA keeps ~14 CTAs resident per SM: the kernel is latency-bound (dependent global
// loads per term), so residency, not bandwidth, sets its speed.
__global__ void __launch_bounds__(VB_SPARSE_THREADS)
vb_sparse_kernel(const VbSparseArgs a)
{
    __shared__ unsigned long long acc[VB_ROWS_PER_BLOCK];       // fp64 bits; sentinel = untouched

    const uint32_t q = blockIdx.x % a.n_queries;
    const uint32_t blk = a.blk_begin + blockIdx.x / a.n_queries;
    const int64_t t_lo = a.q_indptr[q], t_hi = a.q_indptr[q + 1];
    if (t_lo == t_hi) return;                                   // dense-only query

    // any posting of this query in this block?
    int any = 0;
    for (int64_t j = t_lo + threadIdx.x; j < t_hi; j += blockDim.x) {
        const uint64_t* o = a.off + (size_t)j * (a.n_blocks + 1u) + blk;
        any |= (o[1] > o[0]);
    }
    if (!__syncthreads_or(any)) return;

    for (uint32_t r = threadIdx.x; r < VB_ROWS_PER_BLOCK; r += blockDim.x) acc[r] = VB_ACC_SENTINEL;
    __syncthreads();

    const uint32_t row0 = blk * VB_ROWS_PER_BLOCK;
    for (int64_t j = t_lo; j < t_hi; ++j) {
        const uint64_t* o = a.off + (size_t)j * (a.n_blocks + 1u) + blk;
        const uint64_t lo = o[0], hi = o[1];
        if (lo == hi) continue;                                 // uniform across the CTA
        const double w = a.q_weight[j];
        for (uint64_t p0 = lo; p0 < hi; p0 += VB_SPARSE_THREADS * VB_SPARSE_UNROLL) {
            uint32_t r[VB_SPARSE_UNROLL];
            float v[VB_SPARSE_UNROLL];
#pragma unroll
            for (int u = 0; u < VB_SPARSE_UNROLL; ++u) {        // all loads first (memory-level parallelism)
                const uint64_t p = p0 + (uint64_t)u * VB_SPARSE_THREADS + threadIdx.x;
                const bool ok = p < hi;
                r[u] = ok ? __ldg(a.post_row + p) : 0xffffffffu;
                v[u] = ok ? __ldg(a.post_val + p) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < VB_SPARSE_UNROLL; ++u) {        // inside one term every row occurs once
                if (r[u] == 0xffffffffu) continue;
                const uint32_t lr = r[u] - row0;
                const double prod = __dmul_rn(w, (double)v[u]);
                const unsigned long long cur = acc[lr];
                acc[lr] = (cur == VB_ACC_SENTINEL)
                              ? (unsigned long long)__double_as_longlong(__dadd_rn(0.0, prod))
                              : (unsigned long long)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)cur), prod));
            }
        }
        __syncthreads();
    }

    const uint32_t list = a.n_queries + q;                      // sparse lists follow the dense ones
    const float tau = a.tau[list];
    const uint32_t* mask = nullptr;
    if (a.mask != nullptr && a.mask_of != nullptr) {
        const int32_t f = a.mask_of[q];
        if (f >= 0) mask = a.mask + (size_t)f * a.mask_words;
    }
    const uint32_t seg_row0 = a.blk_begin * VB_ROWS_PER_BLOCK;
    for (uint32_t r = threadIdx.x; r < VB_ROWS_PER_BLOCK; r += blockDim.x) {
        const unsigned long long cur = acc[r];
        const uint32_t row = row0 + r;
        bool pass = cur != VB_ACC_SENTINEL && row < a.n_rows;
        if (pass && mask) pass = (mask[row >> 5] >> (row & 31u)) & 1u;
        float s = 0.0f;
        if (pass) { s = __double2float_rn(__longlong_as_double((long long)cur)); pass = s > tau; }
        if (a.direct) {
            if (row < a.n_rows) a.cand[(size_t)list * a.cap + (row - seg_row0)] = pass ? vb_pack_key(s, a.row_base + row) : 0ull;
        } else if (pass) {
            vb_push(a.cand, a.cnt, a.cap, list, s, a.row_base + row);
        }
    }
}
