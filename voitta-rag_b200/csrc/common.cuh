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
