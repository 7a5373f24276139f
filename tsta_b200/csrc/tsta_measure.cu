// tsta_measure.cu -- on-device measurement of the integer/DPX issue roofline.
//
// The fill kernels are bound by integer-ALU issue, and the B200's DPX rate is
// not documented, so the roofline denominator is measured: a dependency-free stream of
// VIADDMNMX (__viaddmax_s32) over 8 independent accumulators per thread, enough warps per
// SM to saturate every scheduler, timed with CUDA events.
#include "common.cuh"

namespace tsta {

constexpr int kDpxChains = 8;
constexpr int kDpxUnroll = 16;

__global__ void __launch_bounds__(256) dpx_peak_kernel(int iters, int seed, int* sink)
{
    int acc[kDpxChains];
#pragma unroll
    for (int k = 0; k < kDpxChains; k++) acc[k] = seed + threadIdx.x * (k + 1);
    const int step = seed | 1, floor_ = -seed;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < kDpxUnroll; u++) {
#pragma unroll
            for (int k = 0; k < kDpxChains; k++) acc[k] = __viaddmax_s32(acc[k], step, floor_ + k);
        }
    }
    int r = 0;
#pragma unroll
    for (int k = 0; k < kDpxChains; k++) r ^= acc[k];
    if (r == 0x7fffffff) sink[0] = r;  // keeps the chains alive without a store in practice
}

}  // namespace tsta

using namespace tsta;

extern "C" int tsta_measure_dpx_peak(double* lane_ops_per_s, float* ms_out)
{
    if (int rc = ensure_device()) return rc;
    int dev = 0, sms = 0;
    TSTA_CUDA_TRY(cudaGetDevice(&dev));
    TSTA_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int* sink = nullptr;
    TSTA_CUDA_TRY(cudaMalloc(&sink, sizeof(int)));
    cudaEvent_t e0, e1;
    TSTA_CUDA_TRY(cudaEventCreate(&e0));
    TSTA_CUDA_TRY(cudaEventCreate(&e1));
    const int blocks = sms * 8, threads = 256, iters = 4096;
    dpx_peak_kernel<<<blocks, threads>>>(64, 3, sink);  // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        TSTA_CUDA_TRY(cudaEventRecord(e0));
        dpx_peak_kernel<<<blocks, threads>>>(iters, 3, sink);
        TSTA_CUDA_TRY(cudaEventRecord(e1));
        TSTA_CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        TSTA_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    TSTA_CUDA_TRY(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    const double ops = (double)blocks * threads * (double)iters * kDpxUnroll * kDpxChains;
    if (lane_ops_per_s) *lane_ops_per_s = ops / (best * 1e-3);
    if (ms_out) *ms_out = best;
    return TSTA_OK;
}
