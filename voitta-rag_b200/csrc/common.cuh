// Shared device helpers: candidate keys, error macros.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>

#define VB_ROWS_PER_BLOCK 2048u   // sparse row block (smem accumulators) and segment alignment

// Bounds checks of our own (compute-sanitizer is closed on the GPU pool): a debug build (-DVB_DEBUG_BOUNDS,
// libvoitta_b200_dbg.so, tools/gpu_sanitize.sh) turns every VB_CHECK into a device-side assert — a failing check traps
// the kernel and the next CUDA call fails loudly.  The release build compiles them away.
#ifdef VB_DEBUG_BOUNDS
#include <assert.h>
#define VB_CHECK(cond) assert(cond)
#else
#define VB_CHECK(cond) ((void)0)
#endif

// ---------------------------------------------------------------------------------------------
// Candidate key: one u64 orders candidates by (score desc, row asc).  0 is never a valid key
// (it would need a NaN score), so 0 marks an empty slot.
//   hi 32 bits: monotone map of the fp32 score;  lo 32 bits: ~row (smaller row = larger key)
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t vb_f32_ordered(uint32_t u) {
    return (u & 0x80000000u) ? ~u : (u ^ 0x80000000u);
}
__host__ __device__ __forceinline__ uint32_t vb_ordered_f32(uint32_t o) {
    return (o & 0x80000000u) ? (o ^ 0x80000000u) : ~o;
}
__device__ __forceinline__ uint64_t vb_pack_key(float score, uint32_t row) {
    score += 0.0f;  // -0.0 -> +0.0 so equal scores have equal keys
    return ((uint64_t)vb_f32_ordered(__float_as_uint(score)) << 32) | (uint64_t)(~row);
}
__host__ __device__ __forceinline__ uint32_t vb_key_row(uint64_t key) { return ~(uint32_t)key; }
__device__ __forceinline__ float vb_key_score(uint64_t key) {
    return __uint_as_float(vb_ordered_f32((uint32_t)(key >> 32)));
}

// A candidate list of `cap` slots is cut into up to VB_SUB sub-ranges of `sub_cap` slots, each with
// its own append counter (cnt[list*VB_SUB + sub]): a CTA appends to sub-range blockIdx.x % nsub, so
// the atomics of one list are spread over nsub addresses instead of serialising on one.
// After a compaction the list lives at the front of sub-range 0.
#define VB_SUB 8u

struct VbLists {
    uint64_t* cand;       // [n_lists][cap]
    uint32_t* cnt;        // [n_lists][VB_SUB]
    uint32_t n_lists;     // (bounds checks)
    uint32_t cap;         // slots per list
    uint32_t sub_cap;     // slots per sub-range = cap / nsub
    uint32_t sub_mask;    // nsub - 1 (nsub is a power of two <= VB_SUB; 1 in safe mode)
};

// Append a surviving candidate to list `list`.  A counter may run past sub_cap: the compaction
// kernel turns that into the overflow flag and the host re-runs the batch in safe mode.
__device__ __forceinline__ void vb_push(const VbLists& L, uint32_t list, float score, uint32_t row) {
    const uint32_t sub = blockIdx.x & L.sub_mask;
    const uint32_t slot = atomicAdd(&L.cnt[list * VB_SUB + sub], 1u);
    if (slot < L.sub_cap) L.cand[(size_t)list * L.cap + (size_t)sub * L.sub_cap + slot] = vb_pack_key(score, row);
}

__device__ __forceinline__ void vb_push_sub(const VbLists& L, uint32_t list, uint32_t sub, float score, uint32_t row) {
    VB_CHECK(list < L.n_lists && sub <= L.sub_mask && (size_t)(sub + 1u) * L.sub_cap <= L.cap);
    const uint32_t slot = atomicAdd(&L.cnt[list * VB_SUB + sub], 1u);
    if (slot < L.sub_cap) L.cand[(size_t)list * L.cap + (size_t)sub * L.sub_cap + slot] = vb_pack_key(score, row);
}

// v[j] for a run-time j (the array lives in registers: a select tree, not an indexed load)
__device__ __forceinline__ uint32_t vb_sel16(const uint32_t (&v)[16], uint32_t j) {
    const bool b0 = j & 1u, b1 = j & 2u, b2 = j & 4u, b3 = j & 8u;
    const uint32_t a0 = b0 ? v[1] : v[0], a1 = b0 ? v[3] : v[2], a2 = b0 ? v[5] : v[4], a3 = b0 ? v[7] : v[6];
    const uint32_t a4 = b0 ? v[9] : v[8], a5 = b0 ? v[11] : v[10], a6 = b0 ? v[13] : v[12], a7 = b0 ? v[15] : v[14];
    const uint32_t c0 = b1 ? a1 : a0, c1 = b1 ? a3 : a2, c2 = b1 ? a5 : a4, c3 = b1 ? a7 : a6;
    const uint32_t d0 = b2 ? c1 : c0, d1 = b2 ? c3 : c2;
    return b3 ? d1 : d0;
}

// Append the survivors of one 16-column chunk of a GEMM epilogue (a warp = 32 rows x 16 queries).  mm = columns with a
// survivor of the pre-threshold somewhere in the warp (warp-uniform), m = this lane's columns.  A run-time loop over the
// set bits of mm: exact comparison on the final score, one slot reservation per column for the whole warp (ballot),
// store.  The loop is deliberately NOT unrolled: this path runs for a few per cent of the chunks, its code is cold in
// the instruction cache every time, and the epilogue warps have ~5 us per tile before the tensor pipe waits for them.
// Measured on cfg4's 7.3M-row segment (a survivor in ~15 % of the chunks): 8.9 ms with a 16x unrolled per-column body
// (~400 instructions), 10.1 ms with a three-loop per-chunk reservation (~600), 13-15 ms with a deferred, buffered
// append (more still) — the smaller the code, the faster the segment, whatever the number of atomic round trips.
template <class ScoreFn, class KeepFn>
__device__ __forceinline__ void vb_append_flagged(const VbLists& L, uint32_t list0, uint32_t sub, uint32_t mm, uint32_t m,
                                                  uint32_t lane, uint32_t lane_lt, uint32_t row_id, ScoreFn score_of, KeepFn keep_of) {
#pragma unroll 1
    for (uint32_t todo = mm; todo != 0u; todo &= todo - 1u) {
        const uint32_t j = (uint32_t)__ffs((int)todo) - 1u;              // warp-uniform
        const float fs = score_of(j);
        const bool keep = ((m >> j) & 1u) && keep_of(j, fs);
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (bal == 0u) continue;
        const uint32_t leader = (uint32_t)__ffs((int)bal) - 1u;
        uint32_t base = 0u;
        if (lane == leader) {
            VB_CHECK(list0 + j < L.n_lists);
            base = atomicAdd(L.cnt + (size_t)(list0 + j) * VB_SUB + sub, (uint32_t)__popc(bal));
        }
        base = __shfl_sync(0xffffffffu, base, leader);
        if (keep) {
            const uint32_t slot = base + (uint32_t)__popc(bal & lane_lt);
            if (slot < L.sub_cap) L.cand[(size_t)(list0 + j) * L.cap + (size_t)sub * L.sub_cap + slot] = vb_pack_key(fs, row_id);
        }
    }
}

// ---- deferred appends (K2T epilogue) --------------------------------------------------------------------------
// A slot reservation is a global atomic whose result the store needs: a ~2 us round trip under load, per flagged
// column, in a warp that has ~5 us per tile before the tensor pipe waits for it.  The query-tiled kernel therefore
// parks its survivors in a per-warp buffer in shared memory (key + list, VB_PEND slots) and appends them 32 at a time
// — one round trip per 32 candidates, each lane its own — from an out-of-line function (cold code stays out of the
// epilogue loop; see vb_append_flagged for what code size does to this path).
#define VB_PEND 64u                  // slots per warp: flushed past 32, and one column adds at most 32
__device__ __noinline__ uint32_t vb_flush_pending(uint64_t* cand, uint32_t* cnt, uint32_t cap, uint32_t sub_cap, uint32_t sub,
                                                  uint32_t pk_addr, uint32_t pl_addr, uint32_t pn, uint32_t lane) {
    __syncwarp();
    while (pn != 0u) {
        const uint32_t take = pn < 32u ? pn : 32u;
        if (lane < take) {
            const uint32_t i = pn - take + lane;
            uint32_t list, klo, khi;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(list) : "r"(pl_addr + i * 4u));
            asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(klo), "=r"(khi) : "r"(pk_addr + i * 8u));
            const uint32_t slot = atomicAdd(cnt + (size_t)list * VB_SUB + sub, 1u);
            if (slot < sub_cap) cand[(size_t)list * cap + (size_t)sub * sub_cap + slot] = ((uint64_t)khi << 32) | klo;
        }
        pn -= take;
    }
    __syncwarp();
    return 0u;
}

// vb_append_flagged with the appends deferred: returns the new number of pending entries (warp-uniform)
template <class ScoreFn, class KeepFn>
__device__ __forceinline__ uint32_t vb_park_flagged(const VbLists& L, uint32_t list0, uint32_t sub, uint32_t mm, uint32_t m,
                                                    uint32_t lane_lt, uint32_t row_id, uint32_t pk_addr, uint32_t pl_addr, uint32_t pn,
                                                    uint32_t lane, ScoreFn score_of, KeepFn keep_of) {
#pragma unroll 1
    for (uint32_t todo = mm; todo != 0u; todo &= todo - 1u) {
        const uint32_t j = (uint32_t)__ffs((int)todo) - 1u;              // warp-uniform
        const float fs = score_of(j);
        const bool keep = ((m >> j) & 1u) && keep_of(j, fs);
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const uint32_t i = pn + (uint32_t)__popc(bal & lane_lt);
            const uint64_t key = vb_pack_key(fs, row_id);
            asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(pk_addr + i * 8u), "r"((uint32_t)key), "r"((uint32_t)(key >> 32)) : "memory");
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(pl_addr + i * 4u), "r"(list0 + j) : "memory");
        }
        pn += (uint32_t)__popc(bal);
        if (pn > 32u) pn = vb_flush_pending(L.cand, L.cnt, L.cap, L.sub_cap, sub, pk_addr, pl_addr, pn, lane);
    }
    return pn;
}

// List set-up of one search, done by the first n threads of the query-prep launch (it was a launch of its own).
// cnt0: slots of the first (direct) segment at the front of every list.  `no_direct` (optional, one flag per query): the
// sparse list of such a query is never written by a direct segment (K3M scores it in stages).  gtau[n..2n) are the
// single-pass scan's ticket counters.
struct VbListInit {
    float* tau; uint32_t* cnt; uint32_t* overflow; uint32_t* gtau;
    uint32_t n, cnt0;
    const uint8_t* no_direct;
    uint32_t n_queries, dense_direct;
};
__device__ __forceinline__ void vb_init_list(const VbListInit& a, uint32_t i) {
    if (i >= a.n) return;
    a.tau[i] = -INFINITY; a.overflow[i] = 0u; a.gtau[i] = 0u; a.gtau[a.n + i] = 0u;
    uint32_t c0 = (a.no_direct != nullptr && i >= a.n_queries && a.no_direct[i - a.n_queries] == 1) ? 0u : a.cnt0;
    if (i < a.n_queries && !a.dense_direct) c0 = 0u;                 // the single-pass scan appends to an empty list
    for (uint32_t s = 0; s < VB_SUB; ++s) a.cnt[i * VB_SUB + s] = s == 0 ? c0 : 0u;
}

__device__ __forceinline__ uint4 vb_ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// Scheduling fence: every load issued before it (volatile asm keeps program order) must be in
// flight before any consumer of `v` runs, so the compiler cannot sink a row's FMAs between the
// loads of different rows (which would serialise the memory requests of a warp).
__device__ __forceinline__ void vb_keep_loaded(uint4& v) {
    asm volatile("" : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w));
}

// One evaluated filter, device side (see vb_filter in include/voitta_b200.h).
struct VbFilterDev {
    const uint32_t* scope_bits;  // device, or nullptr
    uint32_t scope_words;
    int32_t ts_field;            // 0 none, 1 created, 2 modified
    int64_t ts_lo, ts_hi;
};
