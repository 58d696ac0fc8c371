// Microbenchmark (measurement tool, not product): how fast can ONE SM stream an L2-resident f64 matrix into shared memory?
// Variants: 2-D TMA tiles of different box shapes (with / without the 128-byte swizzle), 1-D bulk copies, plain coalesced loads.
// Decides the tile shape of the batched decision kernel's terr_ip0 weight stream.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream tma_stream.cu && ./tma_stream
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
constexpr int ROWS = 64, COLS = 5984;
constexpr int STAGES = 8;

__device__ __forceinline__ unsigned su32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* b, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(void* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(void* b, unsigned parity) {
    asm volatile("{\n .reg .pred p;\n W_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D_%=;\n bra W_%=;\n D_%=:\n}\n" ::"r"(su32(b)), "r"(parity) : "memory");
}

// 2-D TMA: each stage = `per` boxes of [box_rows][box_cols] f64; thread 0 produces, everybody waits (no compute: pure delivery rate)
__global__ void k_tma2d(const __grid_constant__ CUtensorMap map, int box_rows, int box_cols, int per, int iters, unsigned long long* sink) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ unsigned long long full[STAGES];
    const int box_bytes = box_rows * box_cols * 8, stage_bytes = box_bytes * per;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int ncol_boxes = COLS / box_cols, nrow_boxes = ROWS / box_rows, nboxes = ncol_boxes * nrow_boxes;
    auto issue = [&](int it) {
        const int s = it % STAGES;
        mbar_expect(&full[s], stage_bytes);
        for (int p = 0; p < per; ++p) {
            const int b = (it * per + p) % nboxes;
            const int c0 = (b % ncol_boxes) * box_cols, c1 = (b / ncol_boxes) * box_rows;
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                             su32(sm + (size_t)s * stage_bytes + (size_t)p * box_bytes)),
                         "l"(&map), "r"(su32(&full[s])), "r"(c0), "r"(c1)
                         : "memory");
        }
    };
    if (threadIdx.x == 0) for (int it = 0; it < STAGES && it < iters; ++it) issue(it);
    unsigned long long acc = 0;
    for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES;
        mbar_wait(&full[s], (it / STAGES) & 1);
        acc += *(volatile unsigned long long*)(sm + (size_t)s * stage_bytes + (threadIdx.x & 63) * 8);
        __syncthreads();                               // every thread has touched the stage: it may be refilled
        if (threadIdx.x == 0 && it + STAGES < iters) issue(it + STAGES);
    }
    if (acc == 0x1234567) sink[blockIdx.x] = acc;
}
// 1-D bulk copies of `bytes` contiguous bytes per stage
__global__ void k_bulk1d(const double* src, int bytes, int iters, unsigned long long* sink) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ unsigned long long full[STAGES];
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t total = (size_t)ROWS * COLS * 8;
    auto issue = [&](int it) {
        const int s = it % STAGES;
        const size_t off = ((size_t)it * bytes) % (total - bytes) & ~(size_t)15;
        mbar_expect(&full[s], bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(su32(sm + (size_t)s * bytes)),
                     "l"((const char*)src + off), "r"(bytes), "r"(su32(&full[s]))
                     : "memory");
    };
    if (threadIdx.x == 0) for (int it = 0; it < STAGES && it < iters; ++it) issue(it);
    unsigned long long acc = 0;
    for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES;
        mbar_wait(&full[s], (it / STAGES) & 1);
        acc += *(volatile unsigned long long*)(sm + (size_t)s * bytes + (threadIdx.x & 63) * 8);
        __syncthreads();
        if (threadIdx.x == 0 && it + STAGES < iters) issue(it + STAGES);
    }
    if (acc == 0x1234567) sink[blockIdx.x] = acc;
}
// plain loads: every thread keeps `UNR` independent 8-byte loads in flight (B-fragment pattern: 8 rows x 32 B per warp instruction)
template <int UNR>
__global__ void k_ldg(const double* __restrict__ src, int iters, double* sink) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t4 = lane & 3;
    const double* row = src + (size_t)((warp & 7) * 8 + g) * COLS + t4;
    double acc = 0;
    const int kbase = (warp >> 3) * 2992;
    for (int it = 0; it < iters; ++it) {
        double v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) v[u] = row[kbase + ((it * UNR + u) * 4) % 2988];
#pragma unroll
        for (int u = 0; u < UNR; ++u) acc += v[u];
    }
    if (acc == 1.2345) sink[blockIdx.x] = acc;
}

static float timed(void (*launch)(void*), void* ctx) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(ctx); cudaDeviceSynchronize();
    cudaEventRecord(e0); launch(ctx); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    double* d = nullptr; unsigned long long* sink = nullptr;
    CK(cudaMalloc(&d, (size_t)ROWS * COLS * 8)); CK(cudaMemset(d, 0, (size_t)ROWS * COLS * 8)); CK(cudaMalloc(&sink, 4096 * 8));
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)p;
    CK(cudaFuncSetAttribute(k_tma2d, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_bulk1d, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    struct Cfg { int br, bc, per; CUtensorMapSwizzle sw; const char* name; };
    const Cfg cfgs[] = {{64, 16, 1, CU_TENSOR_MAP_SWIZZLE_128B, "box 64x16 (128 B rows) swizzle128, 8 KB/stage"},
                        {64, 16, 1, CU_TENSOR_MAP_SWIZZLE_NONE, "box 64x16 (128 B rows) no swizzle, 8 KB/stage"},
                        {64, 16, 2, CU_TENSOR_MAP_SWIZZLE_128B, "2 x box 64x16 swizzle128, 16 KB/stage"},
                        {32, 16, 1, CU_TENSOR_MAP_SWIZZLE_128B, "box 32x16 swizzle128, 4 KB/stage"},
                        {64, 32, 1, CU_TENSOR_MAP_SWIZZLE_NONE, "box 64x32 (256 B rows) no swizzle, 16 KB/stage"},
                        {16, 64, 1, CU_TENSOR_MAP_SWIZZLE_NONE, "box 16x64 (512 B rows) no swizzle, 8 KB/stage"},
                        {64, 64, 1, CU_TENSOR_MAP_SWIZZLE_NONE, "box 64x64 (512 B rows) no swizzle, 32 KB/stage... 16 stages too big -> 8"},
                        {8, 176, 1, CU_TENSOR_MAP_SWIZZLE_NONE, "box 8x176 (1408 B rows) no swizzle, 11 KB/stage"},
                        {4, 256, 1, CU_TENSOR_MAP_SWIZZLE_NONE, "box 4x256 (2 KB rows) no swizzle, 8 KB/stage"}};
    const int iters = 2000;
    for (int nblk : {1, 8, 148}) {
        printf("---- %d CTA(s), one per SM, %d stages in flight\n", nblk, STAGES);
        for (const Cfg& c : cfgs) {
            if (COLS % c.bc) continue;
            CUtensorMap map;
            const cuuint64_t dims[2] = {COLS, ROWS}; const cuuint64_t strides[1] = {COLS * 8};
            const cuuint32_t box[2] = {(cuuint32_t)c.bc, (cuuint32_t)c.br}; const cuuint32_t es[2] = {1, 1};
            if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("  encode failed: %s\n", c.name); continue; }
            const size_t smem = (size_t)STAGES * c.br * c.bc * 8 * c.per + 1024;
            if (smem > 200 * 1024) continue;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            k_tma2d<<<nblk, 128, smem>>>(map, c.br, c.bc, c.per, iters, sink); CK(cudaDeviceSynchronize());
            cudaEventRecord(e0); k_tma2d<<<nblk, 128, smem>>>(map, c.br, c.bc, c.per, iters, sink); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double bytes = (double)iters * c.br * c.bc * 8 * c.per;
            printf("  TMA 2-D %-72s %7.1f GB/s per SM  (%.2f us per stage)\n", c.name, bytes / (ms * 1e-3) / 1e9, ms * 1e3 / iters);
        }
        for (int bytes : {2048, 8192, 16384}) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            const size_t smem = (size_t)STAGES * bytes + 1024;
            k_bulk1d<<<nblk, 128, smem>>>(d, bytes, iters, sink); CK(cudaDeviceSynchronize());
            cudaEventRecord(e0); k_bulk1d<<<nblk, 128, smem>>>(d, bytes, iters, sink); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("  bulk 1-D %6d B per copy %58s %7.1f GB/s per SM  (%.2f us per stage)\n", bytes, "", (double)iters * bytes / (ms * 1e-3) / 1e9, ms * 1e3 / iters);
        }
        {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            const int it2 = 400;
            k_ldg<8><<<nblk, 512>>>(d, it2, (double*)sink); CK(cudaDeviceSynchronize());
            cudaEventRecord(e0); k_ldg<8><<<nblk, 512>>>(d, it2, (double*)sink); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("  LDG.64 fragments, 512 threads x 8 in flight %43s %7.1f GB/s per SM\n", "", (double)it2 * 8 * 512 * 8 / (ms * 1e-3) / 1e9);
            cudaEventRecord(e0); k_ldg<16><<<nblk, 512>>>(d, it2, (double*)sink); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            cudaEventElapsedTime(&ms, e0, e1);
            printf("  LDG.64 fragments, 512 threads x 16 in flight %42s %7.1f GB/s per SM\n", "", (double)it2 * 16 * 512 * 8 / (ms * 1e-3) / 1e9);
        }
    }
    return 0;
}
