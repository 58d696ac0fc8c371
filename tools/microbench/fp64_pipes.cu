// Microbenchmark (measurement tool, not product): FP64 issue rates on one B200 -- DFMA (vector pipe) against
// mma.sync.m8n8k4.f64 (DMMA), per SM and whole chip, at 1 / 2 / 4 / 8 resident warps per scheduler.  Decides whether the batched
// decision kernel's inner product should be DMMA or register-tiled DFMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu && ./fp64_pipes
#include <cuda_runtime.h>
#include <cstdio>

__global__ void k_dfma(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_dmma(double* out, int iters) {
    double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0000001;
    for (int i = 0; i < iters; ++i) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(a), "d"(b));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}

template <typename K>
double time_ms(K kern, int grid, int block, double* out, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<grid, block>>>(out, iters);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kern<<<grid, block>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double* out;
    cudaMalloc(&out, (size_t)sms * 1024 * 8 * 8);
    const int iters = 20000;
    printf("%s, %d SMs\n", p.name, sms);
    printf("warps/SM   DFMA GFMA/s  (TFLOP/s)   DMMA GMAC/s (TFLOP/s)\n");
    for (int warps : {4, 8, 16, 32}) {
        const int block = warps * 32;
        double t1 = time_ms(k_dfma, sms, block, out, iters);
        double t2 = time_ms(k_dmma, sms, block, out, iters);
        double fma = (double)sms * block * 8.0 * iters / (t1 * 1e-3);                 // thread-level FMAs / s
        double mac = (double)sms * warps * 4.0 * 256.0 * iters / (t2 * 1e-3);         // m8n8k4 = 256 MAC per warp instruction
        printf("%8d   %10.1f  (%6.2f)   %10.1f  (%6.2f)\n", warps, fma / 1e9, 2 * fma / 1e12, mac / 1e9, 2 * mac / 1e12);
    }
    return 0;
}
