// Microbenchmark (development tool): FP64 throughput of mma.sync.m8n8k4.f64 against DFMA on one
// GPU.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_rate dmma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dmma_kernel(double *out, int iters, double seed)
{
    double c[CHAINS][2];
    for (int k = 0; k < CHAINS; ++k) c[k][0] = c[k][1] = seed + k;
    double a = 1.0000001 + threadIdx.x * 1e-9, b = 0.9999999;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int k = 0; k < CHAINS; ++k) s += c[k][0] + c[k][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA rate against the number of independent accumulator chains in flight per SM sub-partition
template <int CHAINS>
void sweep(double *out, int sms, cudaEvent_t e0, cudaEvent_t e1)
{
    for (int warps_per_smsp = 1; warps_per_smsp <= 8; warps_per_smsp *= 2) {
        const int threads = 128, blocks = sms * warps_per_smsp, iters = 2048;  // 4 warps per block = 1 per SMSP
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            dmma_kernel<CHAINS><<<blocks, threads>>>(out, iters, 1.0);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        const double fma = (double)blocks * threads / 32 * iters * CHAINS * 256.0;
        printf("DMMA chains/warp %d, warps/SMSP %d: %.2f TFLOP/s\n", CHAINS, warps_per_smsp, 2 * fma / (best * 1e-3) / 1e12);
    }
}

__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double seed)
{
    double c[8];
    for (int k = 0; k < 8; ++k) c[k] = seed + k + threadIdx.x * 1e-6;
    const double m = 1.0000001, d = 1e-7;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) c[k] = __fma_rn(c[k], m, d);
    double s = 0;
    for (int k = 0; k < 8; ++k) s += c[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 4096;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int which = 0; which < 2; ++which) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            if (which == 0) dmma_kernel<8><<<blocks, threads>>>(out, iters, 1.0);
            else dfma_kernel<<<blocks, threads>>>(out, iters, 1.0);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        // one m8n8k4 mma = 8*8*4 FMA per warp; one DFMA = 32 FMA per warp
        const double warps = (double)blocks * threads / 32;
        const double fma = which == 0 ? warps * iters * 8 * 256.0 : warps * iters * 8 * 32.0;
        printf("%s: %.3f ms, %.2f TFLOP/s (2 flop per FMA), %.2f T warp-instr/s\n", which == 0 ? "DMMA m8n8k4" : "DFMA", best,
               2 * fma / (best * 1e-3) / 1e12, warps * iters * 8 / (best * 1e-3) / 1e12);
    }
    sweep<1>(out, p.multiProcessorCount, e0, e1);
    sweep<2>(out, p.multiProcessorCount, e0, e1);
    sweep<4>(out, p.multiProcessorCount, e0, e1);
    sweep<8>(out, p.multiProcessorCount, e0, e1);
    printf("error state: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
