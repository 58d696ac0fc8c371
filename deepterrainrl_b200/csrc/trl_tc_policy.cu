// deepterrainrl_b200 -- EXPERIMENT, not on the product path: the policy's wide inner product (terr_ip0, 5984 -> 64,
// data/policies/dog/nets/dog_mace3_deploy.prototxt; cNeuralNet::Eval, learning/NeuralNet.cpp:352-375) as a tcgen05 GEMM over a
// batch of decisions, in SPLIT precision: every f64 operand is the sum of up to three narrow parts (bf16 x 3 or tf32 x 2) and the
// product is the sum of the part-by-part products the caller selects, each one a tcgen05.mma with FP32 accumulation in TMEM.
//
// Why it exists: the north star names a tcgen05 GEMM for the policy forward pass; the reference evaluates it in f64 and the parity
// bar is 1e-12, which no tcgen05 kind can hold (there is no f64 kind).  This kernel settles with data what the narrow kinds cost:
// tools/tc_policy_probe.py feeds it the states of ~1e6 real decisions and counts how often the arg-max over the critics flips
// against the f64 network (profiles/tc_policy_r02.json).  The decision path of the library stays f64 (trl_decide2.cuh).
//
// Shape of one launch: D[M][64] = sum over selected (i, j) of A_i[M][K] * B_j[64][K]^T, K = 5984.
//   * operands are K-major planes in HBM: A parts stacked as [parts * M][K], B parts as [parts * 64][K]; TMA (128-byte swizzle)
//     brings one k-block = 128 bytes of K per row of every part into shared memory: 3 x (16 KB + 8 KB) per stage, 3 stages;
//   * warp 0 lane 0 issues the TMA loads, warp 1 lane 0 issues tcgen05.mma (M 128, N 64, K = 32 bytes per instruction, 4 per
//     k-block and pair), tcgen05.commit hands the stage back; the accumulator is 64 TMEM columns x 128 lanes;
//   * all four warps read the accumulator with tcgen05.ld (lane = row) and write FP32 rows (k-split > 1: red.add).
// Every mbarrier wait is bounded: a wrong descriptor must end the launch with an error code, not hang the GPU.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>

namespace {

constexpr int kBM = 128, kBN = 64, kRowBytes = 128;
constexpr int kMaxParts = 3, kStages = 3;
constexpr int kATile = kBM * kRowBytes, kBTile = kBN * kRowBytes;          // 16 KB, 8 KB
constexpr int kStageBytes = kMaxParts * (kATile + kBTile);                  // 72 KB
constexpr int kSmemBytes = kStages * kStageBytes + 1024;                    // + alignment slack
constexpr int kThreads = 128;
constexpr unsigned kTmemCols = 64;
constexpr long long kSpinLimit = 1ll << 24;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: false = gave up (sets *bail so that every role leaves)
__device__ __forceinline__ bool mbar_wait(void* bar, unsigned parity, volatile int* bail) {
    for (long long spin = 0; spin < kSpinLimit; ++spin) {
        unsigned ok;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return true;
        if ((spin & 1023) == 1023 && *bail) return false;
    }
    *bail = 1;
    return false;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, void* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
// shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart
// (start address >> 4 in bits [0,14), stride byte offset >> 4 in [32,46), descriptor version 1 in [46,48), layout type 2 in [61,64))
__device__ __forceinline__ uint64_t umma_desc(unsigned smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D f32 (bits [4,6) = 1), A / B format in [7,10) / [10,13) (1 bf16, 2 tf32), both K-major, N >> 3 in [17,23), M >> 4 in [24,29)
__host__ __device__ constexpr uint32_t umma_idesc(int fmt) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}
template <int KIND>
__device__ __forceinline__ void umma(unsigned tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, unsigned accumulate) {
    if (KIND == 0)
        asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc),
                     "l"(bdesc), "r"(idesc), "r"(accumulate)
                     : "memory");
    else
        asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc),
                     "l"(bdesc), "r"(idesc), "r"(accumulate)
                     : "memory");
}
__device__ __forceinline__ void umma_commit(void* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct Params {
    int M, parts, pairs, nkb, ksplit, elems_per_row;       // pairs: bit (3 i + j) selects A_i x B_j
    float* out;                                             // [M][64]
    int* err;                                               // device flag: 1 = a wait gave up
};

template <int KIND>
__global__ void __launch_bounds__(kThreads, 1)
trl_tc_fc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, Params p) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar_full[kStages], bar_empty[kStages], bar_done;
    __shared__ unsigned tmem_slot;
    __shared__ int bail;
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * kBM;
    const int kb0 = (int)((long long)p.nkb * blockIdx.y / p.ksplit), kb1 = (int)((long long)p.nkb * (blockIdx.y + 1) / p.ksplit);

    if (tid == 0) {
        bail = 0;
        for (int s = 0; s < kStages; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        mbar_init(&bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = tmem_slot;

    if (warp == 0 && lane == 0) {
        // ---- TMA producer
        for (int kb = kb0, it = 0; kb < kb1; ++kb, ++it) {
            const int s = it % kStages;
            if (it >= kStages && !mbar_wait(&bar_empty[s], ((it / kStages) - 1) & 1, &bail)) break;
            mbar_expect_tx(&bar_full[s], (unsigned)(p.parts * (kATile + kBTile)));
            unsigned char* st = smem + s * kStageBytes;
            for (int q = 0; q < p.parts; ++q) {
                tma_load_2d(st + q * kATile, &map_a, &bar_full[s], kb * p.elems_per_row, q * p.M + m0);
                tma_load_2d(st + kMaxParts * kATile + q * kBTile, &map_b, &bar_full[s], kb * p.elems_per_row, q * kBN);
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ---- MMA issuer
        constexpr uint32_t idesc = umma_idesc(KIND == 0 ? 1 : 2);
        unsigned acc = 0;
        bool ok = true;
        for (int kb = kb0, it = 0; kb < kb1 && ok; ++kb, ++it) {
            const int s = it % kStages;
            ok = mbar_wait(&bar_full[s], (it / kStages) & 1, &bail);
            if (!ok) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned a0 = smem_u32(smem + s * kStageBytes), b0 = a0 + kMaxParts * kATile;
            for (int i = 0; i < p.parts; ++i)
                for (int j = 0; j < p.parts; ++j) {
                    if (!((p.pairs >> (3 * i + j)) & 1)) continue;
                    const uint64_t ad = umma_desc(a0 + i * kATile), bd = umma_desc(b0 + j * kBTile);
#pragma unroll
                    for (int k = 0; k < kRowBytes / 32; ++k) {          // 32 bytes of K per instruction: +2 in the (>> 4) start address
                        umma<KIND>(tmem, ad + 2 * k, bd + 2 * k, idesc, acc);
                        acc = 1;
                    }
                }
            umma_commit(&bar_empty[s]);      // the stage is free once these MMAs have read it
        }
        umma_commit(&bar_done);
    }
    __syncwarp();
    // ---- epilogue: every warp reads its 32 lanes (rows) of the 64 accumulator columns
    const bool done = mbar_wait(&bar_done, 0, &bail);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (done && kb1 > kb0) {
        const int row = m0 + warp * 32 + lane;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t v[32];
            const unsigned taddr = tmem + ((unsigned)(warp * 32) << 16) + (unsigned)(half * 32);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
                "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                  "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                  "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                  "=r"(v[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < p.M) {
                float* o = p.out + (size_t)row * kBN + half * 32;
                if (p.ksplit == 1) {
#pragma unroll
                    for (int c = 0; c < 32; c += 4)
                        *reinterpret_cast<float4*>(o + c) = make_float4(__uint_as_float(v[c]), __uint_as_float(v[c + 1]), __uint_as_float(v[c + 2]), __uint_as_float(v[c + 3]));
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) atomicAdd(o + c, __uint_as_float(v[c]));
                }
            }
        }
    }
    if (tid == 0 && bail) *p.err = 1;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
}

// f64 rows -> narrow parts (the residual of every part goes to the next one): bf16 x parts (round to nearest even) or tf32 x parts
// (FP32 with the low 13 mantissa bits rounded away -- the tf32 kind ignores them).  One thread per element, planes [parts][rows][K].
template <int KIND>
__global__ void trl_tc_split_kernel(const double* __restrict__ x, long long n, int parts, void* __restrict__ planes) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r = x[i];
    for (int q = 0; q < parts; ++q) {
        float f = (float)r;
        uint32_t u = __float_as_uint(f);
        if (KIND == 0) {
            u += 0x7fffu + ((u >> 16) & 1u);        // round to nearest even at bit 16
            u &= 0xffff0000u;
            reinterpret_cast<uint16_t*>(planes)[(size_t)q * n + i] = (uint16_t)(u >> 16);
        } else {
            u += 0xfffu + ((u >> 13) & 1u);         // round to nearest even at bit 13
            u &= 0xffffe000u;
            reinterpret_cast<uint32_t*>(planes)[(size_t)q * n + i] = u;
        }
        r -= (double)__uint_as_float(u);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encoder() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult r;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)q;
    }
    return fn;
}
thread_local std::string g_tc_err;
int tc_fail(const std::string& m) { g_tc_err = m; return 1; }
int encode(CUtensorMap* map, int kind, const void* base, uint64_t rows, uint64_t K, uint32_t box_rows) {
    EncodeTiledFn enc = encoder();
    if (!enc) return tc_fail("cuTensorMapEncodeTiled is not available from this driver");
    const uint32_t esz = kind == 0 ? 2 : 4;
    const cuuint64_t dims[2] = {K, rows};
    const cuuint64_t strides[1] = {K * esz};
    const cuuint32_t box[2] = {(cuuint32_t)(kRowBytes / esz), box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, kind == 0 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : tc_fail("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
}

}  // namespace

extern "C" {

const char* trl_tc_last_error(void) { return g_tc_err.c_str(); }

// x[rows][K] f64 (device) -> planes[parts][rows][K] (device; bf16 for kind 0, tf32-rounded FP32 for kind 1)
int trl_tc_split(const double* x, long long rows, int K, int kind, int parts, void* planes, void* stream) {
    if (parts < 1 || parts > kMaxParts || (kind != 0 && kind != 1)) return tc_fail("trl_tc_split: bad kind / parts");
    const long long n = rows * K;
    const int threads = 256;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    if (kind == 0) trl_tc_split_kernel<0><<<blocks, threads, 0, (cudaStream_t)stream>>>(x, n, parts, planes);
    else trl_tc_split_kernel<1><<<blocks, threads, 0, (cudaStream_t)stream>>>(x, n, parts, planes);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : tc_fail(std::string("trl_tc_split: ") + cudaGetErrorString(e));
}

// out[M][64] (FP32, device) = sum over the selected part pairs of A_i[M][K] B_j[64][K]^T on tcgen05.  a_planes: [parts][M][K],
// b_planes: [parts][64][K].  pairs: bit (3 i + j).  ksplit > 1 splits K over gridDim.y (out must be zeroed by the caller).
// err_flag (device int, zeroed by the caller) is set when a bounded barrier wait gave up.
int trl_tc_fc(const void* a_planes, const void* b_planes, int M, int K, int kind, int parts, int pairs, int ksplit, float* out, int* err_flag, void* stream) {
    if (parts < 1 || parts > kMaxParts || (kind != 0 && kind != 1) || M < 1 || ksplit < 1) return tc_fail("trl_tc_fc: bad arguments");
    const int esz = kind == 0 ? 2 : 4;
    if (((size_t)K * esz) % 16 != 0) return tc_fail("trl_tc_fc: the K extent must be a multiple of 16 bytes (TMA global stride)");
    CUtensorMap ma, mb;
    if (encode(&ma, kind, a_planes, (uint64_t)parts * M, (uint64_t)K, kBM)) return 1;
    if (encode(&mb, kind, b_planes, (uint64_t)parts * kBN, (uint64_t)K, kBN)) return 1;
    Params p;
    p.M = M; p.parts = parts; p.pairs = pairs; p.elems_per_row = kRowBytes / esz;
    p.nkb = (K + p.elems_per_row - 1) / p.elems_per_row;
    p.ksplit = ksplit < p.nkb ? ksplit : p.nkb;
    p.out = out; p.err = err_flag;
    dim3 grid((M + kBM - 1) / kBM, p.ksplit);
    cudaError_t e;
    if (kind == 0) {
        e = cudaFuncSetAttribute(trl_tc_fc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e == cudaSuccess) trl_tc_fc_kernel<0><<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(ma, mb, p);
    } else {
        e = cudaFuncSetAttribute(trl_tc_fc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e == cudaSuccess) trl_tc_fc_kernel<1><<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(ma, mb, p);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    return e == cudaSuccess ? 0 : tc_fail(std::string("trl_tc_fc: ") + cudaGetErrorString(e));
}

}  // extern "C"
