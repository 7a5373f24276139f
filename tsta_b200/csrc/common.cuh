// common.cuh -- shared helpers for the tsta_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <limits.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>

#include <nvtx3/nvToolsExt.h>   // header-only; a no-op unless a profiler injects its library

#include "../../include/tsta_cuda.h"

namespace tsta {

// ---------------------------------------------------------------------------------------
// error plumbing (thread-local message, C-ABI style return codes)
// ---------------------------------------------------------------------------------------
extern thread_local std::string g_last_error;

inline int set_error(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
inline int set_error(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define TSTA_CUDA_TRY(expr)                                                                  \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return ::tsta::set_error(e__ == cudaErrorMemoryAllocation ? TSTA_ENOMEM          \
                                                                      : TSTA_ECUDA,          \
                                     "%s failed: %s (%s:%d)", #expr,                         \
                                     cudaGetErrorString(e__), __FILE__, __LINE__);           \
    } while (0)

int ensure_device();  // TSTA_OK when a CUDA device is usable, TSTA_ENODEV otherwise

// NVTX range over the ENQUEUE of a phase (fill / traceback / graph DP): Nsight timelines then name the
// kernels of one alignment; the reference has no tracing at all.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

// Growable device buffer (never shrinks); freed by the owner's destructor.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return TSTA_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return set_error(TSTA_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        }
        cap = bytes;
        return TSTA_OK;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Checkpointed traceback of one long pair (psa_fill16.cuh / psa_trace.cuh): direction words exist only
// for TILES = (band, range of 128 wavefront steps).  One row of the tile table describes the tiles of
// one step range: bands [bmin, bmax] in consecutive slots starting at slot0; bmin > bmax = none.
struct PsaTileRow {
    int bmin, bmax;
    long long slot0;
};
constexpr int kTileSteps = 128;

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__
constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ int ld_acquire_gpu(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// PRMT with the full 4-bit selector semantics (bit 3 of a nibble = replicate sign).
__device__ __forceinline__ int prmt(uint32_t lo, uint32_t hi, uint32_t sel)
{
    int d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "r"(sel));
    return d;
}
__device__ __forceinline__ int warp_max(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(kFullMask, v, o));
    return v;
}
#endif

}  // namespace tsta
