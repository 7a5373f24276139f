// tools/pipe_rates.cu -- issue-rate microbenchmarks for the integer ops the fill kernels use.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/pipe_rates tools/pipe_rates.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(256) k(int iters, int seed, int* sink)
{
    int a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x * (i + 1);
    const int s = seed | 1, f = -seed;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (OP == 0) a[i] = __viaddmax_s32(a[i], s, f + i);
                if (OP == 1) a[i] = max(a[i], f + i + u);                 // VIMNMX
                if (OP == 2) a[i] = __vimax3_s32(a[i], f + i, s + u);     // VIMNMX3
                if (OP == 3) a[i] = a[i] + (s ^ (i + u));                 // IADD3 / VIADD
                if (OP == 4) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(s), "r"(0x3210 + i));
                if (OP == 5) a[i] = a[i] * 3 + s;                          // IMAD
                if (OP == 6) a[i] = (a[i] & 0x7ffffffe) | (s & i);         // LOP3
                if (OP == 7) a[i] = __viaddmax_s16x2(a[i], s, f + i);      // packed
                if (OP == 8) asm volatile("mad.lo.s32 %0, %0, 1, %1;" : "+r"(a[i]) : "r"(s + i)); // IMAD as add
                if (OP == 9) a[i] = (a[i] > f + i + u) ? s : a[i];        // ISETP+SEL
                if (OP == 10) a[i] = __vadd2(a[i], s + i);                 // VIADD.16x2
                if (OP == 11) a[i] = __vmaxs2(a[i], f + i + u);            // VIMNMX.S16x2
                if (OP == 12) a[i] = __vimax3_s16x2(a[i], f + i, s + u);   // VIMNMX3.S16x2
                if (OP == 13) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(s + i), "r"(f));  // LOP3 (xor3)
                if (OP == 14) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(s), "r"(f + i));     // IMAD
                if (OP == 15) asm volatile("add.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(s + i));                   // IADD
                if (OP == 16) asm volatile("shl.b32 %0, %0, 1;" : "+r"(a[i]));                                 // SHF/IMAD.SHL
            }
        }
    }
    int r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r ^= a[i];
    if (r == 0x7fffffff) sink[0] = r;
}

template <int OP> void run(const char* name)
{
    int* sink; cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8, iters = 2048;
    k<OP><<<blocks, 256>>>(16, 3, sink);
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); k<OP><<<blocks, 256>>>(iters, 3, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double ops = (double)blocks * 256 * iters * 16 * 8;
    printf("%-12s %.2f Tlaneop/s  (%.1f lanes/clk/SM at 1.965 GHz)\n", name, ops / best / 1e9, ops / (best * 1e-3) / 148 / 1.965e9);
}

int main()
{
    run<0>("viaddmax"); run<1>("vimnmx"); run<2>("vimnmx3"); run<3>("iadd"); run<4>("prmt");
    run<5>("imad"); run<6>("lop3"); run<7>("viaddmax16x2"); run<8>("imad_add"); run<9>("isetp+sel");
    run<10>("viadd16x2"); run<11>("vimnmx16x2"); run<12>("vimnmx3_16x2"); run<13>("lop3_asm"); run<14>("imad_asm");
    run<15>("iadd_asm"); run<16>("shl_asm");
    return 0;
}
