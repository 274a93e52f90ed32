// Shared device helpers: candidate keys, error macros.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>

#define VB_ROWS_PER_BLOCK 2048u   // sparse row block (smem accumulators) and segment alignment

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

// Append a surviving candidate to list `list`.  cnt may run past cap: the compaction kernel
// turns that into the overflow flag and the host re-runs the batch in safe mode.
__device__ __forceinline__ void vb_push(uint64_t* __restrict__ cand, uint32_t* __restrict__ cnt,
                                        uint32_t cap, uint32_t list, float score, uint32_t row) {
    uint32_t slot = atomicAdd(&cnt[list], 1u);
    if (slot < cap) cand[(size_t)list * cap + slot] = vb_pack_key(score, row);
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
