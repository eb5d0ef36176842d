// Integer-pipe mix microbenchmark: how many 32-bit integer warp instructions per clock can an SM sub-partition
// issue when ALU-pipe (LOP3) and FMA-pipe (IMAD) instructions are mixed in different patterns?
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o int_mix int_mix.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int PATTERN, int CHAINS>
__global__ void __launch_bounds__(256) k(uint32_t iters, uint32_t seed, uint32_t* sink)
{
    uint32_t r[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) r[c] = seed * (2 * c + 1) + threadIdx.x + blockIdx.x;
    const uint32_t k1 = seed | 0x9e3779b1u, k2 = seed ^ 0x7f4a7c15u;
#pragma unroll 1
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int c = 0; c < CHAINS; c++) {
                bool alu;
                const int idx = u * CHAINS + c;
                if (PATTERN == 0) alu = true;                 // ALU only
                else if (PATTERN == 1) alu = false;           // FMA only
                else if (PATTERN == 2) alu = idx & 1;         // alternate
                else if (PATTERN == 3) alu = (idx >> 1) & 1;  // pairs
                else if (PATTERN == 4) alu = (idx % 3) != 0;  // 2 ALU : 1 FMA
                else alu = (idx % 3) == 0;                    // 1 ALU : 2 FMA
                if (alu) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[c]) : "r"(k1), "r"(k2));
                else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[c]) : "r"(k1), "r"(k2));
            }
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) x ^= r[c];
    if (x == 0x12345u) sink[0] = x;
}

__global__ void __launch_bounds__(256) kw(uint32_t iters, uint32_t seed, uint32_t* sink)
{
    unsigned long long r[8];
#pragma unroll
    for (int c = 0; c < 8; c++) r[c] = seed * (2 * c + 1) + threadIdx.x + blockIdx.x;
    const uint32_t k1 = seed | 0x9e3779b1u;
#pragma unroll 1
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int c = 0; c < 8; c++) {
                uint32_t lo = (uint32_t)r[c];
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r[c]) : "r"(lo), "r"(k1));
            }
        }
    }
    unsigned long long x = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) x ^= r[c];
    if (x == 0x12345u) sink[0] = (uint32_t)x;
}

__global__ void __launch_bounds__(256) kw2(uint32_t iters, uint32_t seed, uint32_t* sink)
{
    uint32_t r[16];
#pragma unroll
    for (int c = 0; c < 16; c++) r[c] = seed * (2 * c + 1) + threadIdx.x + blockIdx.x;
    const uint32_t k1 = seed | 0x9e3779b1u;
    unsigned long long acc = 0;
#pragma unroll 1
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int c = 0; c < 16; c++) {
                unsigned long long w;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(r[c]), "r"(k1));
                r[c] = (uint32_t)w;                  // chain through the low word; the high word is kept alive below
                acc ^= w >> 32;
            }
        }
    }
    uint32_t x = (uint32_t)acc;
#pragma unroll
    for (int c = 0; c < 16; c++) x ^= r[c];
    if (x == 0x12345u) sink[0] = x;
}

template <int P, int C>
double run(int ctas_per_sm, int sms, uint32_t* sink)
{
    const uint32_t iters = 1 << 11;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k<P, C><<<sms * ctas_per_sm, 256>>>(iters, rep + 3, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)sms * ctas_per_sm * 256.0 * iters * 4.0 * C;
        if (rep) best = ops / (ms * 1e-3) > best ? ops / (ms * 1e-3) : best;
    }
    return best;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    uint32_t* sink; cudaMalloc(&sink, 64);
    const int sms = p.multiProcessorCount;
    const double clk = p.clockRate * 1e3;
    printf("%s, %d SMs, %.0f MHz: thread-instructions/s (T) and warp-instructions/clk/SMSP\n", p.name, sms, clk / 1e6);
    const char* names[] = {"ALU only (LOP3)", "FMA only (IMAD)", "alternate", "pairs AA FF", "2 ALU : 1 FMA", "1 ALU : 2 FMA"};
#define ROW(P) { double a = run<P, 8>(8, sms, sink), b = run<P, 8>(4, sms, sink), c = run<P, 16>(4, sms, sink), d = run<P, 4>(8, sms, sink); \
    printf("%-18s  8ch x64w %6.2f (%.3f)   8ch x32w %6.2f (%.3f)   16ch x32w %6.2f (%.3f)   4ch x64w %6.2f (%.3f)\n", names[P], a / 1e12, a / 32 / (sms * 4 * clk), \
           b / 1e12, b / 32 / (sms * 4 * clk), c / 1e12, c / 32 / (sms * 4 * clk), d / 1e12, d / 32 / (sms * 4 * clk)); }
    ROW(0) ROW(1) ROW(2) ROW(3) ROW(4) ROW(5)
    // alternating pattern: independent chains per thread x warps per SM
    printf("alternate, chains x warps/SM -> warp-instructions/clk/SMSP\n");
#define CELL(C, W) { double a = run<2, C>(W / 8, sms, sink); printf("  %2dch x %2dw  %.3f", C, W, a / 32 / (sms * 4 * clk)); }
    CELL(4, 16) CELL(4, 32) CELL(4, 64) printf("\n");
    CELL(8, 16) CELL(8, 32) CELL(8, 64) printf("\n");
    CELL(16, 8) CELL(16, 16) CELL(16, 32) printf("\n");
    CELL(32, 8) CELL(32, 16) CELL(32, 32) printf("\n");
    // IMAD.WIDE alone (64-bit result in a register pair): how many per clock?
    {
        double best = 0;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            kw<<<sms * 4, 256>>>(1 << 11, rep + 3, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double ops = (double)sms * 4 * 256.0 * (1 << 11) * 4.0 * 8;
            if (rep && ops / (ms * 1e-3) > best) best = ops / (ms * 1e-3);
        }
        printf("IMAD.WIDE only      8ch x32w %6.2f T/s (%.3f warp-instructions/clk/SMSP)\n", best / 1e12, best / 32 / (sms * 4 * clk));
    }
    {
        double best = 0;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            kw2<<<sms * 4, 256>>>(1 << 11, rep + 3, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double ops = (double)sms * 4 * 256.0 * (1 << 11) * 4.0 * 16;
            if (rep && ops / (ms * 1e-3) > best) best = ops / (ms * 1e-3);
        }
        printf("mul.wide + xor      16ch x32w %6.2f T wide-multiplies/s (each followed by one LOP3 on the ALU pipe)\n", best / 1e12);
    }
    return 0;
}
