// K3D — sparse scoring of the DELTA rows: rows appended since the inverted index was last built.
//
// The reference interleaves store_chunks batches of 100 (services/indexing.py:434,560 ->
// vector_store.py:311-313) and deletes (indexing.py:284, watcher.py:149-171) with searches.  Rebuilding the
// inverted index (a radix sort of every posting) after each of them would make nearly every search
// O(nnz); instead the sorted index stays immutable ("base", rows [0, base_rows)), deletes only clear
// alive bits (and adjust the host-side df table), and the few rows appended since are scored here straight
// from the forward CSR.  vb_upsert merges the delta into the base once it passes a size threshold, off
// the search path.
//
// One warp per delta row, the batch's queries side by side: the host inverts the QUERY batch
// (term -> the queries that contain it, with their weights); the row's terms are walked in ascending
// term id and each term's queries receive w*v in their fp64 accumulator (shared memory, one per query).
// Every query therefore sees its contributions in ascending term id with explicit mul then add — the
// reference's two-pointer merge order (sparse_distances.py sparse_dot_product) — so scores are
// bit-identical for any sign of weights and values.  Rows with no shared term keep the -0.0 sentinel
// and are excluded, as in the reference.
// Roofline: none worth quoting — a delta holds at most max(16384, base/32) rows.
#pragma once
#include "common.cuh"
#include "sparse.cuh"

struct VbDeltaArgs {
    const int64_t* sp_indptr;    // forward index
    const uint32_t* sp_term;
    const float* sp_val;
    const uint32_t* qt_term;     // [n_uterms] distinct terms of the query batch, ascending
    const uint32_t* qt_ptr;      // [n_uterms + 1] into qt_query / qt_weight
    const uint32_t* qt_query;    // [n_qterms]
    const double* qt_weight;     // [n_qterms]
    const uint32_t* mask;        // [n_filters][mask_words] or nullptr
    const int32_t* mask_of;      // [B] or nullptr
    const float* tau;
    VbLists lists;
    uint32_t mask_words, n_uterms, n_queries;
    uint32_t row_begin, row_end; // delta rows scored by this launch
    uint32_t row_base;
};

#define VB_DELTA_MAX_WARPS 8u

static uint32_t vb_delta_warps(uint32_t n_queries) {
    const uint32_t per_warp = n_queries * 8u;
    uint32_t w = (160u * 1024u) / (per_warp ? per_warp : 8u);
    if (w < 1u) w = 1u;
    return w > VB_DELTA_MAX_WARPS ? VB_DELTA_MAX_WARPS : w;
}

__global__ void __launch_bounds__(VB_DELTA_MAX_WARPS * 32)
vb_sparse_delta_kernel(const VbDeltaArgs a)
{
    extern __shared__ __align__(16) unsigned char vb_delta_smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t warps = blockDim.x >> 5;
    double* acc = reinterpret_cast<double*>(vb_delta_smem) + (size_t)warp * a.n_queries;
    const double neg_zero = __longlong_as_double((long long)VB_ACC_SENTINEL);
    for (uint32_t row = a.row_begin + blockIdx.x * warps + warp; row < a.row_end; row += gridDim.x * warps) {
        const int64_t p0 = __ldg(a.sp_indptr + row), p1 = __ldg(a.sp_indptr + row + 1);
        bool any = false;
        uint32_t lo = 0;                                          // both term lists ascend: the search range only shrinks
        for (int64_t p = p0; p < p1; ++p) {
            const uint32_t t = __ldg(a.sp_term + p);
            uint32_t hi = a.n_uterms;
            while (lo < hi) {
                const uint32_t mid = lo + ((hi - lo) >> 1);
                if (__ldg(a.qt_term + mid) < t) lo = mid + 1u; else hi = mid;
            }
            if (lo >= a.n_uterms) break;
            if (__ldg(a.qt_term + lo) != t) continue;
            if (!any) {                                           // first shared term of this row: arm the accumulators
                for (uint32_t q = lane; q < a.n_queries; q += 32u) acc[q] = neg_zero;
                __syncwarp();
                any = true;
            }
            const double v = (double)__ldg(a.sp_val + p);
            const uint32_t k1 = __ldg(a.qt_ptr + lo + 1u);
            for (uint32_t k = __ldg(a.qt_ptr + lo) + lane; k < k1; k += 32u) {     // a query holds a term once: no conflicts
                const uint32_t q = __ldg(a.qt_query + k);
                VB_CHECK(q < a.n_queries);
                acc[q] = __dadd_rn(acc[q], __dadd_rn(__dmul_rn(__ldg(a.qt_weight + k), v), 0.0));
            }
            __syncwarp();
        }
        if (!any) continue;
        for (uint32_t q = lane; q < a.n_queries; q += 32u) {
            const double s = acc[q];
            if ((unsigned long long)__double_as_longlong(s) == VB_ACC_SENTINEL) continue;
            if (a.mask != nullptr && a.mask_of != nullptr) {
                const int32_t f = __ldg(a.mask_of + q);
                if (f >= 0 && !((__ldg(a.mask + (size_t)f * a.mask_words + (row >> 5)) >> (row & 31u)) & 1u)) continue;
            }
            const float sf = __double2float_rn(s);
            const uint32_t list = a.n_queries + q;
            if (sf > a.tau[list])
                vb_push_sub(a.lists, list, ((row * 2654435761u) >> 20) & a.lists.sub_mask, sf, a.row_base + row);
        }
        __syncwarp();
    }
}
