// K3H — hash-accumulate MaxScore sparse scoring for LONG queries (more terms than K3M takes).
// Same replacement as sparse.cuh / sparse_ms.cuh: qdrant's sparse dot product with the IDF modifier
// (vector_store.py:647-656; sparse_distances.py sparse_dot_product), bit-identical results.
//
// A long query (the MCP replay's 17..64-term queries) has tens of ESSENTIAL terms.  K3M must rule them
// out one lookup at a time for every essential posting (ownership + bound: ~25 lookups per posting,
// measured 398 ms on the cfg5 shard); K3 accumulates without lookups but pays 2048 fp64 accumulators
// and every frequent-term column per (2048-row block, query) for ~180 postings (176 ms).  K3H
// accumulates like K3 but in a HASH TABLE sized to the postings, not to the rows:
//   * the plan (vb_ms_plan_kernel, sparse_ms.cuh) orders terms by descending ub, finds the
//     non-essential (NE) suffix under the list's threshold, and picks a row-range width W per query
//     so that a range holds ~1024 essential postings; a work unit = (query, row range);
//   * the CTA finds every essential term's slice of the range through the bucket tables (halving the
//     range if it holds more than the table can take), inserts the postings into a 4096-slot table in
//     shared memory (open addressing on the row, fp64 atomicAdd of w*v: an order-free sum of the
//     essential terms — complete, because every essential term was inserted);
//   * then every occupied slot is a candidate row: filter bit, partial + sum(ub of NE) < tau => drop
//     (almost all), otherwise the NE terms are looked up best-first while the row can still reach tau
//     (vb_ms_finish_row: dense column / bucket table), and the surviving order-free sum is verified
//     against the reference's rounding or re-scored from the forward index.
// Runs per row segment (after K3's direct first segment), compaction between segments, like K3.
// Roofline: shared-memory atomics and instruction issue; postings are read once (8 B each).
#pragma once
#include "sparse_ms.cuh"

#define VB_MH_SLOTS 4096u
#define VB_MH_MAX_POST 3072u        // postings a table takes before the row range is halved
#define VB_MH_EMPTY 0xffffffffu     // (VB_MH_TARGET and vb_mh_shift live in sparse_ms.cuh, next to the plan that uses them)

#ifdef __CUDACC__
struct VbMhArgs {
    const uint32_t* post_row;
    const float* post_val;
    const float* heavy_vals;
    uint32_t heavy_stride;
    const uint32_t* term_tab;
    const int64_t* sp_indptr;
    const uint32_t* sp_term;
    const float* sp_val;
    const int64_t* q_indptr;
    const uint32_t* q_term;
    const double* q_weight;
    const VbMsRec* rec;          // position order; [slo, shi) = the term's postings inside the segment (empty for NE)
    const VbMsQuery* qinfo;
    const uint32_t* hunit_prefix;// [B + 1] exclusive prefix of the work units per query
    uint32_t* counters;          // [3] next work unit
    const uint32_t* mask;
    const int32_t* mask_of;
    const float* tau;
    VbLists lists;
    uint32_t mask_words, n_queries, n_rows, row_base, nt_max;
    uint32_t seg_row0, seg_row1;
};

#define VB_MH_THREADS 256

static size_t vb_mh_smem_bytes(uint32_t nt_max) {
    return (size_t)VB_MH_SLOTS * 12u + (size_t)nt_max * (8u + 8u + 4u * 9u) + 64u;
}

__global__ void __launch_bounds__(VB_MH_THREADS)
vb_mh_score_kernel(const VbMhArgs a)
{
    extern __shared__ __align__(16) unsigned char vb_mh_smem[];
    double* s_acc = reinterpret_cast<double*>(vb_mh_smem);                  // [SLOTS]
    double* s_w = s_acc + VB_MH_SLOTS;                                       // [nt_max]
    double* s_suf = s_w + a.nt_max;                                          // [nt_max]
    uint32_t* s_key = reinterpret_cast<uint32_t*>(s_suf + a.nt_max);         // [SLOTS]
    int32_t* s_hidx = reinterpret_cast<int32_t*>(s_key + VB_MH_SLOTS);       // [nt_max]
    uint32_t* s_tab = reinterpret_cast<uint32_t*>(s_hidx + a.nt_max);        // [nt_max]
    uint32_t* s_shift = s_tab + a.nt_max;
    uint32_t* s_plo = s_shift + a.nt_max;
    uint32_t* s_phi = s_plo + a.nt_max;
    uint32_t* s_slo = s_phi + a.nt_max;                                      // [nt_max] segment range of the term
    uint32_t* s_shi = s_slo + a.nt_max;
    uint32_t* s_lo = s_shi + a.nt_max;                                       // [nt_max] slice of the current row range
    uint32_t* s_cum = s_lo + a.nt_max;                                       // [nt_max + 1]
    __shared__ uint32_t s_unit, s_total;
    const uint32_t tid = threadIdx.x;
    const uint32_t total_units = a.hunit_prefix[a.n_queries];
    for (;;) {
        __syncthreads();
        if (tid == 0) s_unit = atomicAdd(&a.counters[3], 1u);
        __syncthreads();
        const uint32_t u = s_unit;
        if (u >= total_units) break;
        uint32_t lo = 0, hi = a.n_queries;                                    // last query whose first unit is <= u
        while (hi - lo > 1u) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            if (__ldg(a.hunit_prefix + mid) <= u) lo = mid; else hi = mid;
        }
        const uint32_t q = lo;
        const uint32_t t_lo = (uint32_t)__ldg(a.q_indptr + q);
        const uint32_t nt = (uint32_t)__ldg(a.q_indptr + q + 1) - t_lo;
        const VbMsQuery qi = a.qinfo[q];
        const uint32_t n_ess = qi.n_ess;
        VB_CHECK(q < a.n_queries && nt <= a.nt_max && n_ess <= nt && qi.active == 2u);
        for (uint32_t i = tid; i < nt; i += VB_MH_THREADS) {
            const VbMsRec r = a.rec[t_lo + i];
            s_w[i] = r.w; s_suf[i] = r.suf; s_hidx[i] = r.hidx; s_tab[i] = r.tab; s_shift[i] = r.shift;
            s_plo[i] = r.plo; s_phi[i] = r.phi; s_slo[i] = r.slo; s_shi[i] = r.shi;
        }
        const uint32_t list = a.n_queries + q;
        VbMsCtx c;
        c.post_row = a.post_row; c.post_val = a.post_val; c.heavy_vals = a.heavy_vals; c.heavy_stride = a.heavy_stride;
        c.term_tab = a.term_tab; c.sp_indptr = a.sp_indptr; c.sp_term = a.sp_term; c.sp_val = a.sp_val;
        c.q_term = a.q_term + t_lo; c.q_weight = a.q_weight + t_lo;
        const uint32_t* mask = nullptr;
        if (a.mask != nullptr && a.mask_of != nullptr) {
            const int32_t f = __ldg(a.mask_of + q);
            if (f >= 0) mask = a.mask + (size_t)f * a.mask_words;
        }
        c.mask = nullptr;                                                     // the filter bit is tested below
        c.w = s_w; c.suf = s_suf; c.hidx = s_hidx; c.tab = s_tab; c.shift = s_shift; c.plo = s_plo; c.phi = s_phi;
        c.nt = nt; c.n_ess = n_ess; c.tau_lo = qi.tau_lo;
        c.delta = (double)(4u * nt) * 1.1102230246251565e-16;
        c.tau = a.tau[list];
        const uint32_t unit_row0 = a.seg_row0 + ((u - __ldg(a.hunit_prefix + q)) << qi.shift);
        const uint32_t unit_row1 = (uint32_t)min((unsigned long long)a.seg_row1, (unsigned long long)unit_row0 + (1ull << qi.shift));
        uint32_t width = unit_row1 - unit_row0;
        uint32_t pos = unit_row0;
        while (pos < unit_row1) {
            const uint32_t r1 = (uint32_t)min((unsigned long long)unit_row1, (unsigned long long)pos + width);
            __syncthreads();
            // every essential term's slice of rows [pos, r1)
            for (uint32_t i = tid; i < n_ess; i += VB_MH_THREADS) {
                const uint32_t lo_i = vb_ms_seg_bound(a.post_row, a.term_tab, s_slo[i], s_shi[i], s_tab[i], s_shift[i], a.n_rows, pos);
                const uint32_t hi_i = vb_ms_seg_bound(a.post_row, a.term_tab, s_slo[i], s_shi[i], s_tab[i], s_shift[i], a.n_rows, r1);
                s_lo[i] = lo_i;
                s_cum[i + 1u] = hi_i - lo_i;
            }
            for (uint32_t i = tid; i < VB_MH_SLOTS; i += VB_MH_THREADS) { s_key[i] = VB_MH_EMPTY; s_acc[i] = 0.0; }
            __syncthreads();
            if (tid == 0) {                                                   // n_ess <= 256: a serial prefix sum is a few hundred cycles
                uint32_t run = 0;
                s_cum[0] = 0u;
                for (uint32_t i = 0; i < n_ess; ++i) { run += s_cum[i + 1u]; s_cum[i + 1u] = run; }
                s_total = run;
            }
            __syncthreads();
            const uint32_t T = s_total;
            if (T > VB_MH_MAX_POST && width > 1u) { width = (width + 1u) >> 1; continue; }    // too many for the table: halve the range
            // insert: thread per posting of the concatenated slices
            uint32_t k = 0;
            for (uint32_t p = tid; p < T; p += VB_MH_THREADS) {
                while (s_cum[k + 1u] <= p) ++k;
                const uint32_t pp = s_lo[k] + (p - s_cum[k]);
                const uint32_t row = __ldg(a.post_row + pp);
                const double prod = __dmul_rn(s_w[k], (double)__ldg(a.post_val + pp));
                uint32_t slot = (row * 2654435761u) >> 20;                    // 12 bits
                for (;;) {
                    const uint32_t prev = atomicCAS(&s_key[slot], VB_MH_EMPTY, row);
                    if (prev == VB_MH_EMPTY || prev == row) break;
                    slot = (slot + 1u) & (VB_MH_SLOTS - 1u);
                }
                VB_CHECK(slot < VB_MH_SLOTS && pp < s_shi[k] && row >= pos && row < r1);
                atomicAdd(&s_acc[slot], prod);
            }
            __syncthreads();
            // finish: every occupied slot is a row with at least one essential term
            for (uint32_t sl = tid; sl < VB_MH_SLOTS; sl += VB_MH_THREADS) {
                const uint32_t row = s_key[sl];
                if (row == VB_MH_EMPTY) continue;
                if (mask != nullptr && !((__ldg(mask + (row >> 5)) >> (row & 31u)) & 1u)) continue;
                float score;
                if (vb_ms_finish_row(c, n_ess, row, s_acc[sl], score))
                    vb_push_sub(a.lists, list, ((row * 2654435761u) >> 20) & a.lists.sub_mask, score, a.row_base + row);
            }
            pos = r1;
        }
    }
}
#endif  // __CUDACC__
