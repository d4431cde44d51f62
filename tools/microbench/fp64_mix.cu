// Microbenchmark (development tool): issue behaviour of FP64 instructions next to the other pipes
// on one SM sub-partition.  Does a DFMA warp-instruction block the issue port for its two pipe
// cycles, or can integer / FP32 instructions issue in the gap?  Decides whether the walk kernels
// are bound by FP64 x 2 + everything else (only fewer instructions help) or by max(issue slots,
// busiest pipe) (more ILP would help).  All operands are registers (values loaded at run time), as
// in the walk kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix fp64_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

// ND DFMA, NI IMAD (FMA-heavy pipe), NA LOP3 (ALU pipe), NF FFMA per loop iteration, in independent chains
template <int ND, int NI, int NA, int NF>
__global__ void __launch_bounds__(128) mix_kernel(double *out, const double *in, int iters)
{
    double d[ND > 0 ? ND : 1];
    unsigned x[NI > 0 ? NI : 1], a[NA > 0 ? NA : 1];
    float f[NF > 0 ? NF : 1];
    const double m = in[0], c = in[1];
    const unsigned im = (unsigned)in[2], ic = (unsigned)in[3];
    const float fm = (float)in[4], fc = (float)in[5];
#pragma unroll
    for (int k = 0; k < ND; ++k) d[k] = in[6] + k + threadIdx.x * 1e-6;
#pragma unroll
    for (int k = 0; k < NI; ++k) x[k] = ic + k * 977u + threadIdx.x;
#pragma unroll
    for (int k = 0; k < NA; ++k) a[k] = im + k * 131u + threadIdx.x;
#pragma unroll
    for (int k = 0; k < NF; ++k) f[k] = fm + k + threadIdx.x * 1e-3f;
    constexpr int NMAX = (ND > NI ? ND : NI) > (NA > NF ? NA : NF) ? (ND > NI ? ND : NI) : (NA > NF ? NA : NF);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < NMAX; ++k) {   // interleaved by hand so that the mix stays fine-grained
            if (k < ND) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[k]) : "d"(m), "d"(c));
            if (k < NI) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(im), "r"(ic));
            if (k < NA) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(im), "r"(ic));
            if (k < NF) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[k]) : "f"(fm), "f"(fc));
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ND; ++k) s += d[k];
#pragma unroll
    for (int k = 0; k < NI; ++k) s += (double)x[k];
#pragma unroll
    for (int k = 0; k < NA; ++k) s += (double)a[k];
#pragma unroll
    for (int k = 0; k < NF; ++k) s += (double)f[k];
    out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static double g_ghz;
static int g_sms;
static double *g_out, *g_in;

template <int ND, int NI, int NA, int NF>
void run()
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = g_sms * 8, iters = 4096;  // 8 blocks x 4 warps = 8 warps per SMSP
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        mix_kernel<ND, NI, NA, NF><<<blocks, 128>>>(g_out, g_in, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    // cycles per SMSP per loop iteration of one warp: 8 warps per SMSP share the port
    const double cyc = best * 1e-3 * g_ghz * 1e9 / (8.0 * iters);
    const int n = ND + NI + NA + NF;
    printf("DFMA %2d  IMAD %2d  LOP3 %2d  FFMA %2d : %6.2f cycles / warp-iteration   [%d instr; FP64 pipe %d; +FP64 twice %d]\n", ND, NI,
           NA, NF, cyc, n, 2 * ND, n + ND);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    g_ghz = clk_khz * 1e-6;
    g_sms = p.multiProcessorCount;
    printf("%s, %d SMs, %.3f GHz nominal\n", p.name, g_sms, g_ghz);
    cudaMalloc(&g_out, sizeof(double) * g_sms * 8 * 128);
    const double h[8] = {1.0000001, 1e-7, 1664525.0, 1013904223.0, 1.0000001, 1e-7, 1.0, 0.0};
    cudaMalloc(&g_in, sizeof h);
    cudaMemcpy(g_in, h, sizeof h, cudaMemcpyHostToDevice);
    run<8, 0, 0, 0>();
    run<0, 16, 0, 0>();
    run<0, 0, 16, 0>();
    run<0, 0, 0, 16>();
    run<8, 8, 0, 0>();
    run<8, 0, 8, 0>();
    run<8, 0, 0, 8>();
    run<8, 0, 16, 0>();
    run<8, 0, 0, 16>();
    run<8, 4, 8, 4>();
    run<8, 8, 8, 0>();
    run<8, 0, 8, 8>();
    run<4, 4, 8, 4>();
    run<0, 8, 8, 0>();
    run<0, 0, 8, 8>();
    run<0, 8, 0, 8>();
    cudaFree(g_out);
    return 0;
}
