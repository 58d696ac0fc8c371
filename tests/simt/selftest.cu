// TEST INFRASTRUCTURE ONLY -- self-test kernels of the SIMT emulator: the emulator has to get the collectives right, see a
// missing __syncwarp as a schedule-dependent result, and address distributed shared memory across a cluster.
#include <cuda_runtime.h>
#include <cooperative_groups.h>

#include <cstdint>

#define TRL_SIMT_SELFTEST 1
#include "trl_types.h"   // TRL_LAUNCH / TRL_DYN_SHARED

namespace cg = cooperative_groups;

// out[warp*32 + lane] = value of lane (lane+1)%32, + ballot popcount of odd lanes, + xor-butterfly sum of lane ids
__global__ void k_collectives(int* out) {
    const int lane = threadIdx.x & 31, gid = blockIdx.x * blockDim.x + threadIdx.x;
    int v = __shfl_sync(0xffffffffu, 100 * (int)blockIdx.x + (int)threadIdx.x, (lane + 1) & 31);
    unsigned odd = __ballot_sync(0xffffffffu, lane & 1);
    int s = lane;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    double d = __shfl_sync(0xffffffffu, 0.5 * lane, 7);
    out[gid] = v + 1000 * __popc(odd) + 100000 * s + (d == 3.5 ? 0 : 7);
}

// shared-memory neighbour exchange inside a warp; `with_sync` = 0 leaves out the __syncwarp a correct kernel needs
__global__ void k_handoff(int* out, int with_sync) {
    __shared__ int buf[64];
    const int t = threadIdx.x;
    buf[t] = -1;
    __syncthreads();
    for (int round = 0; round < 4; ++round) {
        buf[t] = 10 * round + t;
        if (with_sync) __syncwarp();
        int got = buf[(t & 32) | ((t + 1) & 31)];
        if (with_sync) __syncwarp();
        out[round * 64 + t] = got;
    }
}

// every CTA of the cluster writes its rank pattern into dynamic shared memory; rank r reads rank (r+1)'s block through DSMEM
__global__ void __cluster_dims__(4, 1, 1) k_cluster(int* out) {
    TRL_DYN_SHARED(int, sh);
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), t = threadIdx.x;
    sh[t] = 1000 * (int)blockIdx.x + t;
    cluster.sync();
    const int* peer = cluster.map_shared_rank(sh, (rank + 1) % 4);
    out[blockIdx.x * blockDim.x + t] = peer[(t + 3) % blockDim.x];
    cluster.sync();
}

// block-wide OR and early-exiting warps: warps >= 2 leave before the barrier the others use
__global__ void k_block(int* out) {
    const int t = threadIdx.x;
    if (t >= 64) return;
    int any = __syncthreads_or(t == 37);
    __shared__ int acc;
    if (t == 0) acc = 0;
    __syncthreads();
    atomicAdd(&acc, t);
    __syncthreads();
    out[t] = any * 100000 + acc;
}

extern "C" int simt_selftest(int which, int* out, int arg) {
    int* d = nullptr;
    cudaMalloc(&d, 4096 * sizeof(int));
    cudaMemset(d, 0, 4096 * sizeof(int));
    cudaStream_t st = nullptr;
    switch (which) {
        case 0: TRL_LAUNCH(k_collectives, 3, 64, 0, st, d); break;
        case 1: TRL_LAUNCH(k_handoff, 1, 64, 0, st, d, arg); break;
        case 2: TRL_LAUNCH_CLUSTER(4, k_cluster, 8, 96, 96 * sizeof(int), st, d); break;
        case 3: TRL_LAUNCH(k_block, 2, 128, 0, st, d); break;
        default: return 1;
    }
    cudaMemcpy(out, d, 4096 * sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return 0;
}
