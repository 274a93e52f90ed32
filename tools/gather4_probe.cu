// Probe of cp.async.bulk.tensor.2d ... tile::gather4 on sm_100a (no public docs in this image): which box shape the
// tensor map needs, how four gathered rows land in shared memory under SWIZZLE_128B, and what a stream of 32 gather4
// operations per 128-row tile sustains against one 128-row box load.  Build + run: tools/gpu_gather4_probe.sh
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, int32_t c0, int32_t r0, int32_t r1, int32_t r2, int32_t r3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ void tma_box(uint32_t dst, const CUtensorMap* map, int32_t c0, int32_t c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

// layout probe: two gather4 operations fill one 8-row swizzle atom
__global__ void layout_kernel(const __grid_constant__ CUtensorMap map, uint16_t* out, int8_t r0, int8_t r1, int8_t r2, int8_t r3) {
    extern __shared__ unsigned char raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)raw + 1023u) & ~(uintptr_t)1023u);
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 512; ++i) reinterpret_cast<uint16_t*>(smem)[i] = 0xdead;
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect(smem_u32(&bar), 1024);
        tma_gather4(smem_u32(smem), &map, 64, r0, r1, r2, r3, smem_u32(&bar));
        tma_gather4(smem_u32(smem) + 512, &map, 64, r3 + 20, r2 + 20, r1 + 20, r0 + 20, smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        for (int i = 0; i < 512; ++i) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
    }
}

// throughput probe: every CTA streams `tiles` 128-row x 64-col tiles (16 KB) through a 4-stage ring,
// either as one box load or as 32 gather4 operations issued by the 32 lanes of one warp
template <bool GATHER>
__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap map_box, const __grid_constant__ CUtensorMap map_g,
                                                       const uint32_t* ids, uint32_t tiles_per_cta, uint32_t k_blocks, uint32_t* sink) {
    extern __shared__ unsigned char raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)raw + 1023u) & ~(uintptr_t)1023u);
    __shared__ uint64_t full[4], empty[4];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 4; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t stage = 0, phase = 0, acc = 0;
    if (warp == 0) {
        for (uint32_t t = 0; t < tiles_per_cta; ++t) {
            const uint32_t tile = blockIdx.x * tiles_per_cta + t;
            uint4 my = make_uint4(0, 0, 0, 0);
            if (GATHER) my = reinterpret_cast<const uint4*>(ids)[(size_t)tile * 32u + lane];
            for (uint32_t kb = 0; kb < k_blocks; ++kb) {
                mbar_wait(smem_u32(&empty[stage]), phase ^ 1u);
                if (lane == 0) mbar_expect(smem_u32(&full[stage]), 16384);
                __syncwarp();
                const uint32_t dst = smem_u32(smem + stage * 16384u);
                if (GATHER) tma_gather4(dst + lane * 512u, &map_g, (int32_t)(kb * 64u), (int32_t)my.x, (int32_t)my.y, (int32_t)my.z, (int32_t)my.w, smem_u32(&full[stage]));
                else if (lane == 0) tma_box(dst, &map_box, (int32_t)(kb * 64u), (int32_t)(tile * 128u), smem_u32(&full[stage]));
                if (++stage == 4) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        for (uint32_t t = 0; t < tiles_per_cta; ++t)
            for (uint32_t kb = 0; kb < k_blocks; ++kb) {
                mbar_wait(smem_u32(&full[stage]), phase);
                acc += reinterpret_cast<volatile uint32_t*>(smem + stage * 16384u)[lane * 128u];
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[stage])) : "memory");
                if (++stage == 4) { stage = 0; phase ^= 1u; }
            }
        if (acc == 0x12345678u) sink[0] = acc;
    }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

int main() {
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qr));
    const uint64_t rows = 1u << 20, cols = 768;                 // 1.6 GB of bf16: larger than L2
    uint16_t* d = nullptr;
    CK(cudaMalloc(&d, rows * cols * 2));
    {   // value = (row & 0xff) << 8 | (col & 0xff) for the first 256 rows / columns; enough for the layout probe
        std::vector<uint16_t> h(256 * cols);
        for (uint64_t r = 0; r < 256; ++r) for (uint64_t c = 0; c < cols; ++c) h[r * cols + c] = (uint16_t)((r << 8) | (c & 0xff));
        CK(cudaMemset(d, 0, rows * cols * 2));
        CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    }
    uint16_t* out = nullptr;
    CK(cudaMalloc(&out, 1024));
    CK(cudaFuncSetAttribute(layout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096));
    CUtensorMap good{};
    bool have_good = false;
    for (uint32_t box_rows : {1u}) {   // box rows = 4 encodes, but the instruction then traps (illegal instruction): the box is ONE row
        CUtensorMap map;
        cuuint64_t dims[2] = {cols, rows};
        cuuint64_t strides[1] = {cols * 2};
        cuuint32_t box[2] = {64, box_rows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("box rows %u: encode rc %d\n", box_rows, (int)r);
        if (r != CUDA_SUCCESS) continue;
        CK(cudaMemset(out, 0, 1024));
        layout_kernel<<<1, 32, 4096>>>(map, out, 5, 17, 2, 9);
        cudaError_t e = cudaDeviceSynchronize();
        printf("box rows %u: kernel %s\n", box_rows, cudaGetErrorString(e));
        if (e != cudaSuccess) { printf("(context lost: stop)\n"); return 2; }
        uint16_t h[512];
        CK(cudaMemcpy(h, out, 1024, cudaMemcpyDeviceToHost));
        bool ok = true;
        const int want_rows[8] = {5, 17, 2, 9, 29, 22, 37, 25};
        for (int i = 0; i < 8; ++i) {
            printf("  smem row %d:", i);
            for (int c = 0; c < 8; ++c) {
                const uint16_t v = h[i * 64 + c * 8];
                printf(" (r%u c%u)", v >> 8, v & 0xff);
                const int src_chunk = c ^ (i & 7);                 // SWIZZLE_128B: 16-byte chunk index XOR row-in-atom
                if ((v >> 8) != want_rows[i] || (v & 0xff) != 64 + src_chunk * 8) ok = false;
            }
            printf("\n");
        }
        printf("box rows %u: layout %s the 8-row SWIZZLE_128B atom\n", box_rows, ok ? "MATCHES" : "does NOT match");
        if (ok && !have_good) { good = map; have_good = true; }
    }
    if (!have_good) { printf("no working gather4 tensor map\n"); return 3; }
    // throughput
    CUtensorMap map_box;
    {
        cuuint64_t dims[2] = {cols, rows};
        cuuint64_t strides[1] = {cols * 2};
        cuuint32_t box[2] = {64, 128};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map_box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("box map failed\n"); return 4; }
    }
    const uint32_t ctas = 148, tiles_per_cta = 24, k_blocks = 12;     // 148 x 24 x 128 rows = 454,656 rows, 698 MB
    std::vector<uint32_t> ids((size_t)ctas * tiles_per_cta * 128);
    uint32_t x = 12345u, row = 0;
    for (auto& v : ids) { x = x * 1664525u + 1013904223u; row += 1u + ((x >> 16) & 3u); v = row % (uint32_t)rows; }   // ascending, ~40 % of the rows
    uint32_t *d_ids = nullptr, *sink = nullptr;
    CK(cudaMalloc(&d_ids, ids.size() * 4));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemcpy(d_ids, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384 + 1024));
    CK(cudaFuncSetAttribute(stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384 + 1024));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 3; ++rep)
        for (int g = 0; g < 2; ++g) {
            CK(cudaEventRecord(e0));
            if (g) stream_kernel<true><<<ctas, 64, 4 * 16384 + 1024>>>(map_box, good, d_ids, tiles_per_cta, k_blocks, sink);
            else stream_kernel<false><<<ctas, 64, 4 * 16384 + 1024>>>(map_box, good, d_ids, tiles_per_cta, k_blocks, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            const double bytes = (double)ctas * tiles_per_cta * k_blocks * 16384.0;
            printf("%s: %.3f ms, %.0f GB/s, %.1f M TMA ops/s per SM\n", g ? "gather4 x32 per tile" : "one box per tile   ", ms, bytes / ms / 1e6,
                   (double)tiles_per_cta * k_blocks * (g ? 32 : 1) / ms / 1e3);
        }
    return 0;
}
