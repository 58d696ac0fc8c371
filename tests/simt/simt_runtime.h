// TEST INFRASTRUCTURE ONLY -- a SIMT emulator that lets the *unmodified* kernel sources of deepterrainrl_b200/csrc
// (trl_step.cu, trl_decide.cuh, trl_terrain.cuh, trl_train.cu) and the host code behind the C ABI (trl_host.cu) be compiled
// with g++ and executed on a CPU, thread for thread: every CUDA thread is a fiber, warp collectives (__shfl_sync,
// __ballot_sync, __syncwarp), __syncthreads and cluster.sync() are barriers between fibers, shared memory is per-CTA storage,
// distributed shared memory is addressable across the CTAs of a cluster, streams execute in issue order and graph capture
// records / replays closures.
//
// Purpose: (1) run the GPU parity suite's checks against the kernel *source* in the `-m "not gpu"` tier (this container has
// no GPU), (2) verify experimental kernel variants (e.g. -DTRL_ACCUM_SMEM=1) bit for bit against the default build before
// any GPU time is spent on them.  A collective that not every live lane of a warp reaches deadlocks the fibers and is
// reported -- the emulator therefore also checks the convergence assumptions of the warp-per-environment kernels.
//
// This is NOT a CPU fallback of the product: nothing under deepterrainrl_b200/ loads it, the library it produces lives under
// tests/simt/_build/ and only tests/test_simt_*.py dlopen it.  Arithmetic: compiled with -ffp-contract=off, so results differ
// from the GPU's (FMA-contracted) by rounding; rsqrt is 1/sqrt.
#pragma once
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <type_traits>
#include <utility>
#include <vector>

// ------------------------------------------------------------------------------------------------ qualifiers
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __constant__
#define __shared__ static thread_local
#define __restrict__
#define __launch_bounds__(...)
#define __cluster_dims__(...)
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
    unsigned x, y, z;
    constexpr dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct double2 { double x, y; } __attribute__((aligned(16)));
struct float2 { float x, y; };
struct float4 { float x, y, z, w; } __attribute__((aligned(16)));
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

// ------------------------------------------------------------------------------------------------ fibers
namespace simt {

struct Warp {
    int alive = 0, arrived = 0;
    unsigned gen = 0;
    uint64_t slot[2][32];
    bool present[2][32];
};
struct Block {
    int alive = 0, arrived = 0;
    unsigned gen = 0;
    int or_acc[2] = {0, 0};
    char* dyn_smem = nullptr;
    int rank_in_cluster = 0;
};
struct Cluster {
    int alive = 0, arrived = 0;
    unsigned gen = 0;
    int nblocks = 1;
    Block* blocks = nullptr;
};
struct Thread {
    dim3 tid, bid;
    int lane = 0;
    Warp* warp = nullptr;
    Block* block = nullptr;
    Cluster* cluster = nullptr;
    // scheduling
    void* sp = nullptr;
    bool done = false;
    const volatile unsigned* wait_gen = nullptr;
    unsigned wait_val = 0;
    Thread* warp_base = nullptr;   // lane 0 of this thread's warp (the lanes are contiguous)
    int warp_lanes = 0;
    Thread* group_base = nullptr;  // first thread of the co-scheduled group (CTA, or all CTAs of a cluster)
    int group_size = 0, group_index = 0;
};

extern thread_local Thread* cur;

void yield_until_changed(const volatile unsigned* gen, unsigned val);   // suspends the calling fiber
void run_grid(dim3 grid, dim3 block, size_t dyn_smem, int cluster_size, const std::function<void()>& body);
long long counters(int which);   // 0: launches executed, 1: fiber switches, 2.. : lane-calls of the collectives (kColl*)
enum { kCollShfl64 = 2, kCollShfl32 = 3, kCollBallot = 4, kCollSyncwarp = 5, kCollSyncthreads = 6, kCollCluster = 7, kCollEnd = 8 };
extern long long coll_count[kCollEnd];

// ---- barriers
inline void warp_arrive_and_wait(Warp* w, unsigned g) {
    if (++w->arrived >= w->alive) { w->arrived = 0; w->gen = g + 1; return; }
    yield_until_changed(&w->gen, g);
}
// publish a 64-bit payload, wait for every live lane of the warp, return the generation's buffer index
inline int warp_publish(uint64_t v) {
    Thread* t = cur;
    Warp* w = t->warp;
    const unsigned g = w->gen;
    const int b = g & 1;
    if (w->arrived == 0) for (int l = 0; l < 32; ++l) w->present[b][l] = false;
    w->slot[b][t->lane] = v;
    w->present[b][t->lane] = true;
    warp_arrive_and_wait(w, g);
    return b;
}
inline void block_barrier(int pred, int* or_out) {
    Thread* t = cur;
    Block* bl = t->block;
    const unsigned g = bl->gen;
    if (pred) bl->or_acc[g & 1] = 1;
    if (++bl->arrived >= bl->alive) { bl->arrived = 0; bl->or_acc[(g + 1) & 1] = 0; bl->gen = g + 1; }
    else yield_until_changed(&bl->gen, g);
    if (or_out) *or_out = bl->or_acc[g & 1];
}
inline void cluster_barrier() {
    Thread* t = cur;
    Cluster* c = t->cluster;
    const unsigned g = c->gen;
    if (++c->arrived >= c->alive) { c->arrived = 0; c->gen = g + 1; }
    else yield_until_changed(&c->gen, g);
}
inline void* dyn_smem() { return cur->block->dyn_smem; }

template <typename T>
inline uint64_t to_bits(T v) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    uint64_t u = 0;
    std::memcpy(&u, &v, sizeof(T));
    return u;
}
template <typename T>
inline T from_bits(uint64_t u) {
    T v;
    std::memcpy(&v, &u, sizeof(T));
    return v;
}

// ---- launch plumbing (streams execute in issue order; a capturing thread records closures instead)
struct Graph { std::vector<std::function<void()>> ops; };
extern thread_local Graph* capturing;
void submit(std::function<void()> op);

}  // namespace simt

// the scheduler loads the resumed fiber's indices into these before every switch
extern thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

using std::isfinite;
using std::isnan;
using std::isinf;
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline double min(double a, double b) { return std::fmin(a, b); }
inline double max(double a, double b) { return std::fmax(a, b); }

// ------------------------------------------------------------------------------------------------ device intrinsics
template <typename T>
inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    (void)width;
    ++simt::coll_count[sizeof(T) == 8 ? simt::kCollShfl64 : simt::kCollShfl32];
    const int b = simt::warp_publish(simt::to_bits(v));
    simt::Warp* w = simt::cur->warp;
    const int s = src & 31;
    return w->present[b][s] ? simt::from_bits<T>(w->slot[b][s]) : T(0);
}
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
    (void)width;
    ++simt::coll_count[sizeof(T) == 8 ? simt::kCollShfl64 : simt::kCollShfl32];
    const int b = simt::warp_publish(simt::to_bits(v));
    simt::Warp* w = simt::cur->warp;
    const int s = (simt::cur->lane ^ m) & 31;
    return w->present[b][s] ? simt::from_bits<T>(w->slot[b][s]) : v;
}
template <typename T>
inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
    (void)width;
    const int b = simt::warp_publish(simt::to_bits(v));
    simt::Warp* w = simt::cur->warp;
    const int s = simt::cur->lane + (int)d;
    return (s < 32 && w->present[b][s]) ? simt::from_bits<T>(w->slot[b][s]) : v;
}
template <typename T>
inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
    (void)width;
    const int b = simt::warp_publish(simt::to_bits(v));
    simt::Warp* w = simt::cur->warp;
    const int s = simt::cur->lane - (int)d;
    return (s >= 0 && w->present[b][s]) ? simt::from_bits<T>(w->slot[b][s]) : v;
}
inline unsigned __ballot_sync(unsigned, int pred) {
    ++simt::coll_count[simt::kCollBallot];
    const int b = simt::warp_publish(pred ? 1u : 0u);
    simt::Warp* w = simt::cur->warp;
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) if (w->present[b][l] && w->slot[b][l]) m |= 1u << l;
    return m;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, !pred) == 0; }
inline void __syncwarp(unsigned = 0xffffffffu) { ++simt::coll_count[simt::kCollSyncwarp]; simt::warp_publish(0); }
inline void __syncthreads() { ++simt::coll_count[simt::kCollSyncthreads]; simt::block_barrier(0, nullptr); }
inline int __syncthreads_or(int pred) { int r; simt::block_barrier(pred, &r); return r; }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
// round-to-nearest arithmetic that must not be contracted: the emulator is built with -ffp-contract=off
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __double2float_rn(double a) { return (float)a; }
template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline T __ldcg(const T* p) { return *p; }

template <typename T>
inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline double atomicAdd(double* p, double v) {
    double old = *p, des;
    do { des = old + v; } while (!__atomic_compare_exchange(p, &old, &des, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    return old;
}
inline float atomicAdd(float* p, float v) {
    float old = *p, des;
    do { des = old + v; } while (!__atomic_compare_exchange(p, &old, &des, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    return old;
}
template <typename T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
template <typename T> inline T atomicCAS(T* p, T cmp, T v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED); return cmp; }
template <typename T> inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }

// ------------------------------------------------------------------------------------------------ host runtime
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
typedef struct simtStream_* cudaStream_t;
struct simtEvent_ { std::chrono::steady_clock::time_point t; };
typedef simtEvent_* cudaEvent_t;
typedef simt::Graph* cudaGraph_t;
typedef simt::Graph* cudaGraphExec_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaEventDefault = 0 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal, cudaStreamCaptureModeThreadLocal, cudaStreamCaptureModeRelaxed };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { char name[256]; int multiProcessorCount; int major, minor; size_t totalGlobalMem; size_t sharedMemPerBlockOptin; };

inline const char* cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : "simt emulator error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    std::memset(p, 0, sizeof(*p));
    std::snprintf(p->name, sizeof(p->name), "SIMT emulator (tests only)");
    p->multiProcessorCount = 4; p->major = 10; p->minor = 0; p->totalGlobalMem = (size_t)1 << 34; p->sharedMemPerBlockOptin = 232448;
    return cudaSuccess;
}
inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -1; return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
template <typename T> inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)std::aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? cudaSuccess : 2; }
template <typename T> inline cudaError_t cudaMallocHost(T** p, size_t n) { *p = (T*)std::aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? cudaSuccess : 2; }
template <typename T> inline cudaError_t cudaHostAlloc(T** p, size_t n, unsigned) { return cudaMallocHost(p, n); }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2D(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind) {
    for (size_t r = 0; r < h; ++r) std::memmove((char*)d + r * dp, (const char*)s + r * sp, w);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) {
    simt::submit([=] { std::memmove(d, s, n); });
    return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) {
    simt::submit([=] { std::memset(d, v, n); });
    return cudaSuccess;
}
template <typename T>
inline cudaError_t cudaMemcpyToSymbol(T& sym, const void* s, size_t n, size_t off = 0, cudaMemcpyKind = cudaMemcpyHostToDevice) {
    std::memcpy((char*)&sym + off, s, n);
    return cudaSuccess;
}
inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = (cudaStream_t) new int(0); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { return cudaStreamCreate(s); }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { return cudaStreamCreate(s); }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete (int*)s; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new simtEvent_(); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) {
    if (!simt::capturing) e->t = std::chrono::steady_clock::now();
    return cudaSuccess;
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { simt::capturing = new simt::Graph(); return cudaSuccess; }
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { *g = simt::capturing; simt::capturing = nullptr; return cudaSuccess; }
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, unsigned long long = 0) { *e = new simt::Graph(*g); return cudaSuccess; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t) { for (auto& op : e->ops) op(); return cudaSuccess; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t g) { delete g; return cudaSuccess; }
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

// cudaLaunchKernelEx (trl_train.cu's programmatic-dependent-launch helper): attributes are ignored, the order is serial anyway
enum cudaLaunchAttributeID { cudaLaunchAttributeProgrammaticStreamSerialization = 1 };
struct cudaLaunchAttribute { cudaLaunchAttributeID id; struct { int programmaticStreamSerializationAllowed; } val; };
struct cudaLaunchConfig_t { dim3 gridDim, blockDim; size_t dynamicSmemBytes; cudaStream_t stream; cudaLaunchAttribute* attrs; unsigned numAttrs; };
template <typename... KArgs, typename... Args>
inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, void (*kern)(KArgs...), Args... args) {
    const dim3 g = cfg->gridDim, b = cfg->blockDim;
    const size_t sm = cfg->dynamicSmemBytes;
    simt::submit([=] { simt::run_grid(g, b, sm, 1, [=] { kern(args...); }); });
    return cudaSuccess;
}

// kernel<<<grid, block, smem, stream>>>(args...) is spelled TRL_LAUNCH(...) in the sources (trl_types.h)
#define SIMT_LAUNCH(cluster, kern, grid, block, smem, st, ...)                                              \
    do {                                                                                                    \
        const dim3 g_(grid), b_(block);                                                                     \
        const size_t sm_ = (smem);                                                                          \
        simt::submit([=] { simt::run_grid(g_, b_, sm_, (cluster), [=] { kern(__VA_ARGS__); }); });          \
    } while (0)
