// FP32 FMA issue-rate probe (development aid): scalar FFMA, packed FFMA2 (fma.rn.f32x2) and mixtures, 16 independent
// accumulator chains per thread, 32 warps per SM.  Prints lane-FMAs per clock per SM for each mix.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fma_probe tools/fma_probe.cu && ./tools/fma_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float ffma(float a, float b, float c)
{
    float d;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// NS scalar chains and NP packed chains per loop iteration
template <int NS, int NP>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float seed)
{
    float s[NS > 0 ? NS : 1];
    unsigned long long p[NP > 0 ? NP : 1];
    for (int i = 0; i < NS; ++i) s[i] = seed + i + threadIdx.x;
    for (int i = 0; i < NP; ++i) p[i] = ((unsigned long long)__float_as_uint(seed + i) << 32) | __float_as_uint(seed + threadIdx.x);
    const float m = 1.0000001f;
    const unsigned long long m2 = ((unsigned long long)__float_as_uint(m) << 32) | __float_as_uint(m);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < NS; ++i) s[i] = ffma(s[i], m, 0.5f);
#pragma unroll
            for (int i = 0; i < NP; ++i) p[i] = ffma2(p[i], m2, m2);
        }
    }
    float acc = 0.f;
    for (int i = 0; i < NS; ++i) acc += s[i];
    for (int i = 0; i < NP; ++i) acc += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// NP packed (or NS scalar) FMA chains plus NI integer (ALU-pipe) chains per loop iteration: do other pipes' instructions
// issue in the second cycle of an FFMA2?
template <int NS, int NP, int NI>
__global__ void __launch_bounds__(256) probe_mix(float* out, int iters, float seed)
{
    float s[NS > 0 ? NS : 1];
    unsigned long long p[NP > 0 ? NP : 1];
    unsigned q[NI > 0 ? NI : 1];
    for (int i = 0; i < NS; ++i) s[i] = seed + i + threadIdx.x;
    for (int i = 0; i < NP; ++i) p[i] = ((unsigned long long)__float_as_uint(seed + i) << 32) | __float_as_uint(seed + threadIdx.x);
    for (int i = 0; i < NI; ++i) q[i] = threadIdx.x * 2654435761u + i;
    const float m = 1.0000001f;
    const unsigned long long m2 = ((unsigned long long)__float_as_uint(m) << 32) | __float_as_uint(m);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < (NS > NP ? NS : NP) || i < NI; ++i) {
                if (i < NS) s[i] = ffma(s[i], m, 0.5f);
                if (i < NP) p[i] = ffma2(p[i], m2, m2);
                if (i < NI) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(q[(i + 1) % (NI > 0 ? NI : 1)]), "r"(0x9e3779b9u + it));
            }
        }
    }
    float acc = 0.f;
    for (int i = 0; i < NS; ++i) acc += s[i];
    for (int i = 0; i < NP; ++i) acc += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    for (int i = 0; i < NI; ++i) acc += (float)q[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int NS, int NP, int NI>
void run_mix(const char* name, float* out, int sms)
{
    const int iters = 20000;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    probe_mix<NS, NP, NI><<<sms * 4, 256>>>(out, 100, 1.f);
    cudaEventRecord(a);
    probe_mix<NS, NP, NI><<<sms * 4, 256>>>(out, iters, 1.f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double cyc = ms * 1e-3 * clk_khz * 1e3;
    const double per_it = cyc / ((double)iters * 4 * 8);          // cycles per SMSP per (iteration x r) for its 8 warps
    printf("%-34s %8.3f ms  %6.2f cycles per warp-round (FMA floor %d, ALU floor %d, issue floor %d)\n", name, ms, per_it,
           NS + 2 * NP, 2 * NI, NS + NP + NI);
}

template <int NS, int NP>
void run(const char* name, float* out, int sms)
{
    const int iters = 20000;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    probe<NS, NP><<<sms * 4, 256>>>(out, 100, 1.f);
    cudaEventRecord(a);
    probe<NS, NP><<<sms * 4, 256>>>(out, iters, 1.f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double fmas = (double)sms * 4 * 256 * iters * 4 * (NS + 2.0 * NP);
    const double instr = (double)sms * 4 * 8 * iters * 4 * (NS + NP);          // warp instructions
    printf("%-28s %8.3f ms  %7.1f lane-FMA/clk/SM (at %d MHz)  %5.2f warp-instr/clk/SMSP\n", name, ms,
           fmas / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1000, instr / (ms * 1e-3) / sms / 4 / (clk_khz * 1e3));
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out;
    cudaMalloc(&out, (size_t)sms * 4 * 256 * sizeof(float));
    run<16, 0>("scalar FFMA x16", out, sms);
    run<0, 8>("FFMA2 x8", out, sms);
    run<0, 16>("FFMA2 x16", out, sms);
    run<8, 8>("FFMA x8 + FFMA2 x8", out, sms);
    run<8, 4>("FFMA x8 + FFMA2 x4", out, sms);
    run<4, 8>("FFMA x4 + FFMA2 x8", out, sms);
    run<16, 8>("FFMA x16 + FFMA2 x8", out, sms);
    run_mix<0, 8, 0>("FFMA2 x8", out, sms);
    run_mix<0, 8, 4>("FFMA2 x8 + LOP3 x4", out, sms);
    run_mix<0, 8, 8>("FFMA2 x8 + LOP3 x8", out, sms);
    run_mix<16, 0, 0>("FFMA x16", out, sms);
    run_mix<16, 0, 4>("FFMA x16 + LOP3 x4", out, sms);
    run_mix<16, 0, 8>("FFMA x16 + LOP3 x8", out, sms);
    run_mix<0, 0, 8>("LOP3 x8", out, sms);
    cudaFree(out);
    return 0;
}
