// tsta_psa_batch.cu -- many independent pairs in one call (include/tsta_cuda.h: tsta_psa_batch).
//
// The reference aligns one pair per process (psa/psa.c:192-193 reads record 1 only); this
// entry point is the batch front-end BASELINE.json's config 4 needs.  Per pair it computes
// exactly what tsta_psa_align computes; pairs are independent, so persistent warps pull
// whole pairs from a ticket and there is no inter-warp dependency.
#include <algorithm>
#include <thread>
#include <chrono>
#include <vector>

#include "common.cuh"
#include "psa_fill.cuh"
#include "psa_fill16.cuh"
#include "psa_trace.cuh"

namespace tsta {
int check_params(const tsta_params* p);
cudaError_t launch_byte_presence(const uint8_t* p, long long nbytes, unsigned* present, int blocks, cudaStream_t st);
int build_lut(const unsigned present[8], uint8_t lut[256]);

template <int C, bool TRACE, int WARPS>
static void launch_batch(bool prmt, int blocks, cudaStream_t st, const PsaBatchArgs& a)
{
    if (prmt) psa_batch_kernel<C, TRACE, true, WARPS><<<blocks, 32 * WARPS, 0, st>>>(a);
    else psa_batch_kernel<C, TRACE, false, WARPS><<<blocks, 32 * WARPS, 0, st>>>(a);
}
}  // namespace tsta

using namespace tsta;

namespace {
// Device buffers, stream and events of the calling thread, kept across calls: a caller that
// streams batches pays for the allocations once (cudaMalloc/cudaFree of hundreds of MB cost
// more than the kernels of a 1M x 250 bp batch).
constexpr int kMaxChunks = 16;
struct BatchCtx {
    int device = -1;
    cudaStream_t st = nullptr;             // kernels
    cudaStream_t cp = nullptr, dh = nullptr;   // host -> device copies of the chunks ahead; device -> host of finished chunks
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaEvent_t evc[kMaxChunks] = {};      // chunk c uploaded and its alphabet scanned
    DevBuf s0, s1, off0, off1, len0, len1, ms, corner, ctl, scratch, lastrow, ops, ops_off, alen, dirs_off, dirs, p0, p1, scan;
    int* h_present = nullptr;   // pinned, 8 words per chunk
    void* h_scan = nullptr;     // pinned: the descriptor scan's per-block results
    long long* h_dirs_off = nullptr;   // pinned, grown on demand: traced batches, offset of every pair's direction matrix in its slice
    size_t h_dirs_off_cap = 0;
    std::vector<cudaEvent_t> ev;   // timing events, two per kernel slice, grown on demand
    std::vector<size_t> pair_bytes;   // traced batches: bytes of every pair's direction matrix
    void release()
    {
        DevBuf* all[] = {&s0, &s1, &off0, &off1, &len0, &len1, &ms, &corner, &ctl, &scratch, &lastrow, &ops, &ops_off, &alen, &dirs_off, &dirs, &p0, &p1, &scan};
        for (DevBuf* b : all) b->release();
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        for (auto& e : evc) { if (e) cudaEventDestroy(e); e = nullptr; }
        for (auto& e : ev) cudaEventDestroy(e);
        ev.clear();
        if (st) cudaStreamDestroy(st);
        if (cp) cudaStreamDestroy(cp);
        if (dh) cudaStreamDestroy(dh);
        if (h_present) cudaFreeHost(h_present);
        if (h_scan) cudaFreeHost(h_scan);
        if (h_dirs_off) cudaFreeHost(h_dirs_off);
        st = cp = dh = nullptr; e0 = e1 = nullptr; h_present = nullptr; h_scan = nullptr; h_dirs_off = nullptr; h_dirs_off_cap = 0; device = -1;
    }
    ~BatchCtx() { /* process exit: the CUDA context may already be gone; leave the buffers to it */ }
};
thread_local BatchCtx g_ctx;

}  // namespace

namespace tsta {
// tsta_psa_release_cache(): the calling thread's batch buffers, streams and pinned staging go back to the driver
void batch_release_ctx()
{
    if (g_ctx.device < 0) return;
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(g_ctx.device);
    g_ctx.release();
    cudaSetDevice(cur);
    std::vector<size_t>().swap(g_ctx.pair_bytes);
}
}  // namespace tsta

namespace {
struct Guard {   // shim over the persistent context: same call sites as a scoped allocator
    BatchCtx& c;
    cudaStream_t& st;
    cudaEvent_t &e0, &e1;
    bool done = false;   // set on the success path, which has synchronised already
    explicit Guard(BatchCtx& ctx) : c(ctx), st(ctx.st), e0(ctx.e0), e1(ctx.e1) {}
    // every early error return leaves copies and kernels in flight that read or write the CALLER's buffers
    // (asynchronous when they are page-locked): wait for them before the caller may free those buffers
    ~Guard()
    {
        if (done) return;
        if (c.cp) cudaStreamSynchronize(c.cp);
        if (c.st) cudaStreamSynchronize(c.st);
        if (c.dh) cudaStreamSynchronize(c.dh);
    }
    template <class T> int alloc(T** out, size_t count, DevBuf& buf)
    {
        if (int rc = buf.reserve(std::max<size_t>(count, 1) * sizeof(T))) return rc;
        *out = buf.as<T>();
        return TSTA_OK;
    }
};
constexpr int kBatchWarps = 4;
}  // namespace

static int psa_batch_impl(const tsta_params* prm, int64_t npairs,
                          const char* seqs0, const int64_t* off0, const int32_t* len0,
                          const char* seqs1, const int64_t* off1, const int32_t* len1,
                          int want_trace, int32_t* maxsorce, int32_t* corner,
                          uint8_t* ops, const int64_t* ops_off, int32_t* alen,
                          float* kernel_ms, const bool pack2, const bool bits2 = false);

extern "C" int tsta_psa_batch(const tsta_params* prm, int64_t npairs,
                              const char* seqs0, const int64_t* off0, const int32_t* len0,
                              const char* seqs1, const int64_t* off1, const int32_t* len1,
                              int want_trace, int32_t* maxsorce, int32_t* corner,
                              uint8_t* ops, const int64_t* ops_off, int32_t* alen,
                              float* kernel_ms)
{
    return psa_batch_impl(prm, npairs, seqs0, off0, len0, seqs1, off1, len1, want_trace, maxsorce, corner, ops, ops_off, alen, kernel_ms, false);
}

extern "C" int tsta_psa_batch_packed(const tsta_params* prm, int64_t npairs,
                                     const char* seqs0, const int64_t* off0, const int32_t* len0,
                                     const char* seqs1, const int64_t* off1, const int32_t* len1,
                                     int32_t* maxsorce, int32_t* corner,
                                     uint8_t* ops2, const int64_t* ops_off, int32_t* alen,
                                     float* kernel_ms)
{
    return psa_batch_impl(prm, npairs, seqs0, off0, len0, seqs1, off1, len1, 1, maxsorce, corner, ops2, ops_off, alen, kernel_ms, true);
}

extern "C" int tsta_psa_batch_2bit(const tsta_params* prm, int64_t npairs,
                                   const uint8_t* bases0, const int64_t* off0, const int32_t* len0,
                                   const uint8_t* bases1, const int64_t* off1, const int32_t* len1,
                                   int want_trace, int32_t* maxsorce, int32_t* corner,
                                   uint8_t* ops2, const int64_t* ops_off, int32_t* alen,
                                   float* kernel_ms)
{
    return psa_batch_impl(prm, npairs, reinterpret_cast<const char*>(bases0), off0, len0, reinterpret_cast<const char*>(bases1), off1, len1,
                          want_trace ? 1 : 0, maxsorce, corner, ops2, ops_off, alen, kernel_ms, want_trace != 0, true);
}

// 2-bit wire format -> the byte reads the kernels index their tables with: base i of the stream is bits 2(i%4)..
// of byte i/4, codes 0..3 = A C G T.  One thread per four output bytes.
__global__ void __launch_bounds__(256) unpack_2bit_kernel(const uint8_t* __restrict__ packed, uint8_t* __restrict__ out, long long lo, long long hi)
{
    const long long q0 = lo >> 2, q1 = (hi + 3) >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = q0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; q < q1; q += stride) {
        const unsigned b = packed[q];
        const unsigned w = __byte_perm(0x54474341u, 0u, (b & 3u) | (((b >> 2) & 3u) << 4) | (((b >> 4) & 3u) << 8) | (((b >> 6) & 3u) << 12));   // "ACGT"
        const long long i = q << 2;
        if (i >= lo && i + 4 <= hi) *reinterpret_cast<uint32_t*>(out + i) = w;
        else
            for (int k = 0; k < 4; k++)
                if (i + k >= lo && i + k < hi) out[i + k] = (uint8_t)(w >> (8 * k));
    }
}

// Descriptor scan of a batch ON THE DEVICE: validation, the extents every chunk of the upload needs, totals.  The
// descriptors have to be on the device anyway; reading 28 MB of them on the host cost 3.5-4.9 ms per 1M pairs even on
// four threads (a third of the kernel time of a score-only 1M x 250 bp batch), here it is one pass at HBM speed.
// One block of threads per scan block; chunk boundaries are scan-block boundaries (power-of-two chunk counts).
constexpr int kScanBlocks = 64;
struct BlockScan { int64_t lo0, hi0, lo1, hi1, max0, max1, ops_lo, ops_hi, bad; double cells; };
__global__ void __launch_bounds__(256) batch_scan_kernel(const long long* __restrict__ off0, const int* __restrict__ len0,
                                                         const long long* __restrict__ off1, const int* __restrict__ len1,
                                                         const long long* __restrict__ ops_off, const int pack2,
                                                         const long long npairs, BlockScan* __restrict__ out)
{
    const int b = blockIdx.x;
    const long long k0 = npairs * b / kScanBlocks, k1 = npairs * (b + 1) / kScanBlocks;
    long long lo0 = LLONG_MAX, hi0 = 0, lo1 = LLONG_MAX, hi1 = 0, mx0 = 0, mx1 = 0, olo = LLONG_MAX, ohi = 0, bad = LLONG_MAX;
    double cells = 0.0;
    for (long long k = k0 + threadIdx.x; k < k1; k += blockDim.x) {
        const long long a = len0[k], c = len1[k], o0 = off0[k], o1 = off1[k];
        bool isbad = a < c || c < 0 || o0 < 0 || o1 < 0 || a + c > 4000000;
        lo0 = min(lo0, o0); hi0 = max(hi0, o0 + a);
        lo1 = min(lo1, o1); hi1 = max(hi1, o1 + c);
        mx0 = max(mx0, a); mx1 = max(mx1, c);
        cells += (double)a * (double)c;
        if (ops_off) {
            const long long oo = ops_off[k];
            isbad = isbad || oo < 0 || (pack2 && (oo & 3));
            olo = min(olo, oo); ohi = max(ohi, oo + a + c);
        }
        if (isbad) bad = min(bad, k);
    }
    __shared__ BlockScan part[8];
    for (int d = 16; d > 0; d >>= 1) {
        lo0 = min(lo0, __shfl_xor_sync(0xffffffffu, lo0, d)); hi0 = max(hi0, __shfl_xor_sync(0xffffffffu, hi0, d));
        lo1 = min(lo1, __shfl_xor_sync(0xffffffffu, lo1, d)); hi1 = max(hi1, __shfl_xor_sync(0xffffffffu, hi1, d));
        mx0 = max(mx0, __shfl_xor_sync(0xffffffffu, mx0, d)); mx1 = max(mx1, __shfl_xor_sync(0xffffffffu, mx1, d));
        olo = min(olo, __shfl_xor_sync(0xffffffffu, olo, d)); ohi = max(ohi, __shfl_xor_sync(0xffffffffu, ohi, d));
        bad = min(bad, __shfl_xor_sync(0xffffffffu, bad, d));
        cells += __shfl_xor_sync(0xffffffffu, cells, d);
    }
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = BlockScan{(int64_t)lo0, (int64_t)hi0, (int64_t)lo1, (int64_t)hi1, (int64_t)mx0, (int64_t)mx1, (int64_t)olo, (int64_t)ohi, (int64_t)bad, cells};
    __syncthreads();
    if (threadIdx.x == 0) {
        BlockScan r = part[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) {
            const BlockScan q = part[w];
            auto mn = [](int64_t x, int64_t y) { return x < y ? x : y; };
            auto mx = [](int64_t x, int64_t y) { return x > y ? x : y; };
            r.lo0 = mn(r.lo0, q.lo0); r.hi0 = mx(r.hi0, q.hi0); r.lo1 = mn(r.lo1, q.lo1); r.hi1 = mx(r.hi1, q.hi1);
            r.max0 = mx(r.max0, q.max0); r.max1 = mx(r.max1, q.max1);
            r.ops_lo = mn(r.ops_lo, q.ops_lo); r.ops_hi = mx(r.ops_hi, q.ops_hi); r.bad = mn(r.bad, q.bad);
            r.cells += q.cells;
        }
        if (r.bad == LLONG_MAX) r.bad = -1;
        out[b] = r;
    }
}

static int psa_batch_impl(const tsta_params* prm, int64_t npairs,
                          const char* seqs0, const int64_t* off0, const int32_t* len0,
                          const char* seqs1, const int64_t* off1, const int32_t* len1,
                          int want_trace, int32_t* maxsorce, int32_t* corner,
                          uint8_t* ops, const int64_t* ops_off, int32_t* alen,
                          float* kernel_ms, const bool pack2, const bool bits2)
{
    if (kernel_ms) *kernel_ms = 0.f;
    // TSTA_BATCH_TIMING=1: host-side phase times of this call on stderr (where the end-to-end time beyond the kernels goes)
    const bool timing = getenv("TSTA_BATCH_TIMING") != nullptr;
    const auto t_call = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (timing) fprintf(stderr, "[tsta_psa_batch] %8.3f ms  %s\n",
                            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count(), what);
    };
    if (int rc = check_params(prm)) return rc;
    if (npairs < 0 || !maxsorce || !corner) return set_error(TSTA_EINVAL, "bad batch arguments");
    if (npairs == 0) return TSTA_OK;
    if (!seqs0 || !seqs1 || !off0 || !off1 || !len0 || !len1) return set_error(TSTA_EINVAL, "NULL batch input");
    if (want_trace && (!ops || !ops_off || !alen)) return set_error(TSTA_EINVAL, "traceback requested without ops/ops_off/alen");
    if (int rc = ensure_device()) return rc;

    int dev = 0, sms = 0;
    TSTA_CUDA_TRY(cudaGetDevice(&dev));
    TSTA_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (g_ctx.device != dev) {
        if (g_ctx.device >= 0) { cudaSetDevice(g_ctx.device); g_ctx.release(); cudaSetDevice(dev); }
        TSTA_CUDA_TRY(cudaStreamCreateWithFlags(&g_ctx.st, cudaStreamNonBlocking));
        TSTA_CUDA_TRY(cudaStreamCreateWithFlags(&g_ctx.cp, cudaStreamNonBlocking));
        TSTA_CUDA_TRY(cudaStreamCreateWithFlags(&g_ctx.dh, cudaStreamNonBlocking));
        TSTA_CUDA_TRY(cudaEventCreate(&g_ctx.e0));
        TSTA_CUDA_TRY(cudaEventCreate(&g_ctx.e1));
        for (auto& e : g_ctx.evc) TSTA_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        TSTA_CUDA_TRY(cudaMallocHost((void**)&g_ctx.h_present, kMaxChunks * 8 * sizeof(int)));
        TSTA_CUDA_TRY(cudaMallocHost(&g_ctx.h_scan, kScanBlocks * sizeof(BlockScan)));
        g_ctx.device = dev;
    }
    Guard g(g_ctx);

    // ---- the descriptors go to the device in one piece and are scanned there (batch_scan_kernel) -----------------
    constexpr int kBlocks = kScanBlocks;
    long long *d_off0, *d_off1, *d_ops_off = nullptr;
    int *d_len0, *d_len1;
    BlockScan* d_scan;
    if (int rc = g.alloc(&d_off0, (size_t)npairs, g_ctx.off0)) return rc;
    if (int rc = g.alloc(&d_off1, (size_t)npairs, g_ctx.off1)) return rc;
    if (int rc = g.alloc(&d_len0, (size_t)npairs, g_ctx.len0)) return rc;
    if (int rc = g.alloc(&d_len1, (size_t)npairs, g_ctx.len1)) return rc;
    if (want_trace) { if (int rc = g.alloc(&d_ops_off, (size_t)npairs, g_ctx.ops_off)) return rc; }
    if (int rc = g.alloc(&d_scan, (size_t)kScanBlocks, g_ctx.scan)) return rc;
    static_assert(sizeof(long long) == sizeof(int64_t), "offset width");
    TSTA_CUDA_TRY(cudaMemcpyAsync(d_off0, off0, (size_t)npairs * 8, cudaMemcpyHostToDevice, g_ctx.cp));
    TSTA_CUDA_TRY(cudaMemcpyAsync(d_len0, len0, (size_t)npairs * 4, cudaMemcpyHostToDevice, g_ctx.cp));
    TSTA_CUDA_TRY(cudaMemcpyAsync(d_off1, off1, (size_t)npairs * 8, cudaMemcpyHostToDevice, g_ctx.cp));
    TSTA_CUDA_TRY(cudaMemcpyAsync(d_len1, len1, (size_t)npairs * 4, cudaMemcpyHostToDevice, g_ctx.cp));
    if (want_trace) TSTA_CUDA_TRY(cudaMemcpyAsync(d_ops_off, ops_off, (size_t)npairs * 8, cudaMemcpyHostToDevice, g_ctx.cp));
    batch_scan_kernel<<<kScanBlocks, 256, 0, g_ctx.cp>>>(d_off0, d_len0, d_off1, d_len1, d_ops_off, pack2 ? 1 : 0, npairs, d_scan);
    TSTA_CUDA_TRY(cudaGetLastError());
    TSTA_CUDA_TRY(cudaMemcpyAsync(g_ctx.h_scan, d_scan, kScanBlocks * sizeof(BlockScan), cudaMemcpyDeviceToHost, g_ctx.cp));
    TSTA_CUDA_TRY(cudaStreamSynchronize(g_ctx.cp));
    const BlockScan* const bs = static_cast<const BlockScan*>(g_ctx.h_scan);
    mark("descriptor scan done");
    int64_t tot0 = 0, tot1 = 0, max0 = 0, max1 = 0, ops_total = 0;
    double cells = 0.0;
    for (int b = 0; b < kBlocks; b++) {
        if (bs[b].bad >= 0) {
            const int64_t k = bs[b].bad;
            if (len0[k] < len1[k] || len1[k] < 0 || off0[k] < 0 || off1[k] < 0)
                return set_error(TSTA_EINVAL, "pair %lld: need len0 >= len1 >= 0 (the caller swaps, psa.c:194-199)", (long long)k);
            if ((int64_t)len0[k] + len1[k] > 4000000) return set_error(TSTA_EINVAL, "pair %lld too long", (long long)k);
            if (ops_off[k] < 0) return set_error(TSTA_EINVAL, "pair %lld: negative ops_off", (long long)k);
            return set_error(TSTA_EINVAL, "pair %lld: packed moves need ops_off to be a multiple of 4", (long long)k);
        }
        tot0 = std::max(tot0, bs[b].hi0); tot1 = std::max(tot1, bs[b].hi1);
        max0 = std::max(max0, bs[b].max0); max1 = std::max(max1, bs[b].max1);
        ops_total = std::max(ops_total, bs[b].ops_hi);
        cells += bs[b].cells;
    }
    const int S = want_trace ? 2 : 1;

    uint8_t *d_s0, *d_s1;
    int *d_ms, *d_corner, *d_ctl;
    if (int rc = g.alloc(&d_s0, (size_t)tot0, g_ctx.s0)) return rc;
    if (int rc = g.alloc(&d_s1, (size_t)tot1, g_ctx.s1)) return rc;
    uint8_t *d_p0 = nullptr, *d_p1 = nullptr;       // 2-bit wire format: the packed bytes as they arrive
    if (bits2) {
        if (int rc = g.alloc(&d_p0, (size_t)(tot0 + 3) / 4 + 4, g_ctx.p0)) return rc;
        if (int rc = g.alloc(&d_p1, (size_t)(tot1 + 3) / 4 + 4, g_ctx.p1)) return rc;
    }
    if (int rc = g.alloc(&d_ms, (size_t)npairs, g_ctx.ms)) return rc;
    if (int rc = g.alloc(&d_corner, (size_t)npairs, g_ctx.corner)) return rc;
    if (int rc = g.alloc(&d_ctl, 16 + 8 * kMaxChunks, g_ctx.ctl)) return rc;

    // ---- chunks: the reads of chunk c+1 travel to the device while chunk c is computed -----------
    // (a 1M x 250 bp batch is 0.5 GB of reads against 15 ms of kernel time).  Every chunk gets its
    // own alphabet scan, so its kernels can start as soon as its bytes have arrived.
    // Worth it only when the copy is not negligible against the kernels (short reads: ~125 cells per
    // byte; 10 kbp reads: 5000) and every chunk still has many pairs per resident warp.
    int nch = (int)std::min<int64_t>(kMaxChunks, std::max<int64_t>(1, (tot0 + tot1) / ((int64_t)32 << 20)));
    nch = (int)std::min<int64_t>(nch, std::max<int64_t>(1, npairs / ((int64_t)sms * 24 * 16)));
    if (cells > 1000.0 * (double)(tot0 + tot1)) nch = 1;
    if (const char* e = getenv("TSTA_BATCH_CHUNKS")) nch = std::max(1, std::min(kMaxChunks, atoi(e)));
    nch = (int)std::min<int64_t>(nch, npairs);
    while (nch & (nch - 1)) nch &= nch - 1;          // a power of two: chunk boundaries are block boundaries of the scan above
    int64_t kb[kMaxChunks + 1];
    int64_t ext[kMaxChunks][4] = {};                  // 2-bit wire format: base extents of each chunk's two streams
    for (int c = 0; c <= nch; c++) kb[c] = npairs * (c * (kBlocks / nch)) / kBlocks;
    unsigned* d_present = reinterpret_cast<unsigned*>(d_ctl + 16);
    TSTA_CUDA_TRY(cudaMemsetAsync(d_present, 0, 8 * kMaxChunks * sizeof(unsigned), g_ctx.cp));
    for (int c = 0; c < nch; c++) {
        int64_t lo0 = INT64_MAX, hi0 = 0, lo1 = INT64_MAX, hi1 = 0;
        for (int b = c * (kBlocks / nch); b < (c + 1) * (kBlocks / nch); b++) {
            lo0 = std::min(lo0, bs[b].lo0); hi0 = std::max(hi0, bs[b].hi0);
            lo1 = std::min(lo1, bs[b].lo1); hi1 = std::max(hi1, bs[b].hi1);
        }
        cudaStream_t cp = g_ctx.cp;
        if (bits2) {
            // a quarter of the bytes cross PCIe; the device expands them to the byte reads (0.1 ms per 500 MB)
            // (the expansion runs on the kernel stream, in front of the chunk's fill: behind a persistent fill kernel on
            // another stream it would wait for SM slots)
            if (hi0 > lo0) TSTA_CUDA_TRY(cudaMemcpyAsync(d_p0 + (lo0 >> 2), seqs0 + (lo0 >> 2), (size_t)(((hi0 + 3) >> 2) - (lo0 >> 2)), cudaMemcpyHostToDevice, cp));
            if (hi1 > lo1) TSTA_CUDA_TRY(cudaMemcpyAsync(d_p1 + (lo1 >> 2), seqs1 + (lo1 >> 2), (size_t)(((hi1 + 3) >> 2) - (lo1 >> 2)), cudaMemcpyHostToDevice, cp));
            ext[c][0] = lo0; ext[c][1] = hi0; ext[c][2] = lo1; ext[c][3] = hi1;
        } else {
        if (hi0 > lo0) TSTA_CUDA_TRY(cudaMemcpyAsync(d_s0 + lo0, seqs0 + lo0, (size_t)(hi0 - lo0), cudaMemcpyHostToDevice, cp));
        if (hi1 > lo1) TSTA_CUDA_TRY(cudaMemcpyAsync(d_s1 + lo1, seqs1 + lo1, (size_t)(hi1 - lo1), cudaMemcpyHostToDevice, cp));
        }
        if (bits2) {          // the alphabet is known: A C G T
            unsigned* hp = reinterpret_cast<unsigned*>(g_ctx.h_present + 8 * c);
            for (int w = 0; w < 8; w++) hp[w] = 0;
            for (const char ch : {'A', 'C', 'G', 'T'}) hp[(unsigned char)ch >> 5] |= 1u << ((unsigned char)ch & 31);
        } else {
            if (hi0 > lo0) TSTA_CUDA_TRY(launch_byte_presence(d_s0 + lo0, hi0 - lo0, d_present + 8 * c, sms * 4, cp));
            if (hi1 > lo1) TSTA_CUDA_TRY(launch_byte_presence(d_s1 + lo1, hi1 - lo1, d_present + 8 * c, sms * 4, cp));
            TSTA_CUDA_TRY(cudaMemcpyAsync(g_ctx.h_present + 8 * c, d_present + 8 * c, 8 * sizeof(unsigned), cudaMemcpyDeviceToHost, cp));
        }
        TSTA_CUDA_TRY(cudaEventRecord(g_ctx.evc[c], cp));
    }

    mark("uploads queued");
    PsaBatchArgs a{};
    a.sc.M = prm->M; a.sc.X = prm->X; a.sc.E = prm->E; a.sc.O = prm->O;
    // strip widths depend on the longest read of the whole batch; the largest band any variant uses
    // is 512 columns (per-chunk choices below only pick among them)
    constexpr int kMaxBandCols = 512;

    // persistent grid: ~24 warps per SM (register-limited), never more warps than pairs
    int blocks = sms * 6;
    blocks = (int)std::min<int64_t>(blocks, (npairs + kBatchWarps - 1) / kBatchWarps);
    const long long nwarps = (long long)blocks * kBatchWarps;
    unsigned long long* d_scratch;
    int* d_lastrow;
    if (int rc = g.alloc(&d_scratch, (size_t)nwarps * (size_t)std::max<int64_t>(max1, 1), g_ctx.scratch)) return rc;
    if (int rc = g.alloc(&d_lastrow, (size_t)nwarps * kMaxBandCols, g_ctx.lastrow)) return rc;

    a.seqs0 = d_s0; a.seqs1 = d_s1;
    a.scratch = d_scratch; a.scratch_stride = std::max<int64_t>(max1, 1);
    a.ticket = reinterpret_cast<unsigned long long*>(d_ctl);
    a.lastrow = d_lastrow;

    // traceback: direction matrices are large (len1 * ceil(len0/band)*band/2 bytes per pair),
    // so pairs go through in slices that fit a fixed HBM budget
    long long* dirs_off = nullptr;            // pinned and kept across calls (a fresh 8 MB vector per call was 2 ms of page faults)
    std::vector<size_t>& pair_bytes = g_ctx.pair_bytes;
    uint8_t* d_dirs = nullptr;
    long long* d_dirs_off = nullptr;
    uint8_t* d_ops = nullptr;
    int* d_alen = nullptr;
    size_t dirs_budget = 0;
    bool budget_known = false;
    if (want_trace) {
        if (int rc = g.alloc(&d_ops, (size_t)(pack2 ? (ops_total + 3) / 4 + 4 : ops_total), g_ctx.ops)) return rc;
        if (int rc = g.alloc(&d_alen, (size_t)npairs, g_ctx.alen)) return rc;
        if (int rc = g.alloc(&d_dirs_off, (size_t)npairs, g_ctx.dirs_off)) return rc;
        if ((size_t)npairs > g_ctx.h_dirs_off_cap) {
            if (g_ctx.h_dirs_off) cudaFreeHost(g_ctx.h_dirs_off);
            g_ctx.h_dirs_off = nullptr; g_ctx.h_dirs_off_cap = 0;
            TSTA_CUDA_TRY(cudaMallocHost((void**)&g_ctx.h_dirs_off, (size_t)npairs * sizeof(long long)));
            g_ctx.h_dirs_off_cap = (size_t)npairs;
        }
        dirs_off = g_ctx.h_dirs_off;
        // slices of direction matrices: as large as HBM allows -- a slice with fewer pairs than resident warps (3552)
        // leaves sub-partitions idle.  What is already allocated is a budget that needs no driver call; free memory
        // is only asked for (cudaMemGetInfo, ~1 ms) when a chunk does not fit it.
        dirs_budget = g_ctx.dirs.cap;
    }
    auto full_budget = [&]() -> int {
        if (budget_known) return TSTA_OK;
        size_t free_b = 0, total_b = 0;
        TSTA_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        dirs_budget = (size_t)((double)(free_b + g_ctx.dirs.cap) * 0.8);
        budget_known = true;
        return TSTA_OK;
    };

    // No host synchronisation inside the loop: the host prepares chunk c+1 (slice layout, alphabet)
    // while the kernels of chunk c run; kernel time is summed from event pairs at the end.
    size_t nev = 0;
    auto next_event = [&](cudaEvent_t* out) -> cudaError_t {
        if (nev == g_ctx.ev.size()) {
            cudaEvent_t e;
            cudaError_t rc = cudaEventCreate(&e);
            if (rc != cudaSuccess) return rc;
            g_ctx.ev.push_back(e);
        }
        *out = g_ctx.ev[nev++];
        return cudaSuccess;
    };
    mark("buffers ready");
    for (int c = 0; c < nch; c++) {
        // ---- this chunk's alphabet decides its kernels -------------------------------------------
        if (!bits2) TSTA_CUDA_TRY(cudaEventSynchronize(g_ctx.evc[c]));       // (its alphabet has to be read on the host)
        TSTA_CUDA_TRY(cudaStreamWaitEvent(g.st, g_ctx.evc[c], 0));
        if (bits2) {
            if (ext[c][1] > ext[c][0]) unpack_2bit_kernel<<<sms * 8, 256, 0, g.st>>>(d_p0, d_s0, ext[c][0], ext[c][1]);
            if (ext[c][3] > ext[c][2]) unpack_2bit_kernel<<<sms * 8, 256, 0, g.st>>>(d_p1, d_s1, ext[c][2], ext[c][3]);
            TSTA_CUDA_TRY(cudaGetLastError());
        }
        mark("chunk: reads on the device");
        const unsigned* present = reinterpret_cast<const unsigned*>(g_ctx.h_present + 8 * c);
        const int K = build_lut(present, a.sc.lut);
        const bool prmt = (K <= kPadCode) && abs(S * prm->M) <= 127 && abs(S * prm->X) <= 127;
        // packed int16 kernels (psa_fill16.cuh) when alphabet and score range allow; 64*C columns per band
        // (traced: at most 4 columns per virtual lane, the direction ring of wider strips would
        // not leave room for enough warps per SM)
        const int C16 = (max0 <= 128) ? 2 : ((max0 <= 256 || want_trace) ? 4 : 8);
        const char* off16 = getenv("TSTA_PSA_PACKED");
        const bool packed = !(off16 && atoi(off16) == 0) &&
                            psa16_fits(prm->M, prm->X, prm->E, prm->O, want_trace ? 4 : 1, C16, K);
        const char* offd = getenv("TSTA_PSA_DUAL");
        const bool dual = packed && max0 <= 256 && !(offd && atoi(offd) == 0) &&
                          psa16_fits(prm->M, prm->X, prm->E, prm->O, want_trace ? 4 : 1, 8, K);
        const int Cd = max0 <= 128 ? 4 : 8;                          // dual-pair mode: 32*Cd columns, one band
        const int C = dual ? Cd : (packed ? C16 : ((max0 <= 128) ? 4 : 8));
        const int band_cols = dual ? 32 * Cd : (packed ? 64 : 32) * C;
        auto dirs_bytes = [&](int64_t k) -> size_t {
            const size_t nb = (size_t)((len0[k] + band_cols - 1) / band_cols);
            return (size_t)len1[k] * nb * (size_t)(band_cols / 2);
        };
        const int64_t kend = kb[c + 1];
        size_t slice_target = dirs_budget;   // equal slices: a short last slice would leave most warps idle
        if (want_trace) {
            // bytes of every pair's direction matrix: one division per pair, on four host threads for large chunks
            pair_bytes.resize((size_t)npairs);
            {
                const int64_t a0 = kb[c], n = kend - kb[c];
                const int nthreads = n >= 200000 ? 4 : 1;
                auto part = [&](int t) {
                    for (int64_t k = a0 + n * t / nthreads; k < a0 + n * (t + 1) / nthreads; k++)
                        pair_bytes[(size_t)k] = (dirs_bytes(k) + 255) & ~(size_t)255;
                };
                std::thread th[3];
                for (int t = 1; t < nthreads; t++) th[t - 1] = std::thread(part, t);
                part(0);
                for (int t = 1; t < nthreads; t++) th[t - 1].join();
            }
            size_t total = 0;
            for (int64_t k = kb[c]; k < kend; k++) total += pair_bytes[(size_t)k];
            if (total > dirs_budget) { if (int rc = full_budget()) return rc; }
            const size_t nslices = std::max<size_t>(1, (total + dirs_budget - 1) / std::max<size_t>(dirs_budget, 1));
            slice_target = std::min(dirs_budget, (total + nslices - 1) / nslices + ((size_t)64 << 20));
        }
        int64_t k0 = kb[c];
        cudaEvent_t last_done = nullptr;
        while (k0 < kend) {
            int64_t k1 = kend;
            size_t slice_bytes = 0;
            if (want_trace) {
                k1 = k0;
                while (k1 < kend) {
                    const size_t b = pair_bytes[(size_t)k1];
                    if (k1 > k0 && slice_bytes + b > slice_target) break;
                    dirs_off[(size_t)k1] = (long long)slice_bytes;
                    slice_bytes += b;
                    k1++;
                }
                if (slice_bytes > g_ctx.dirs.cap) TSTA_CUDA_TRY(cudaStreamSynchronize(g.st));   // (the old buffer may still be read)
                if (int rc = g_ctx.dirs.reserve(std::max<size_t>(slice_bytes, 256))) return rc;
                d_dirs = g_ctx.dirs.as<uint8_t>();
                TSTA_CUDA_TRY(cudaMemcpyAsync(d_dirs_off + k0, dirs_off + k0, (size_t)(k1 - k0) * 8,
                                              cudaMemcpyHostToDevice, g.st));
            }
            a.off0 = d_off0 + k0; a.off1 = d_off1 + k0; a.len0 = d_len0 + k0; a.len1 = d_len1 + k0;
            a.npairs = k1 - k0;
            a.dirs = d_dirs; a.dirs_off = want_trace ? d_dirs_off + k0 : nullptr;
            a.maxsorce = d_ms + k0; a.corner = d_corner + k0;
            cudaEvent_t ea, eb;
            TSTA_CUDA_TRY(next_event(&ea));
            TSTA_CUDA_TRY(next_event(&eb));
            TSTA_CUDA_TRY(cudaMemsetAsync(d_ctl, 0, 8, g.st));
            TSTA_CUDA_TRY(cudaEventRecord(ea, g.st));
            if (want_trace) {
                if (packed && dual) {   // short reads: two pairs per warp; the step ring is 16 KB (Cd = 4) or 32 KB (Cd = 8) per warp
                    if (Cd == 4) psa_batch16_dual_kernel<4, true, 4><<<blocks, 128, 0, g.st>>>(a);
                    else psa_batch16_dual_kernel<8, true, 2><<<blocks * 2, 64, 0, g.st>>>(a);
                } else if (packed) {   // two warps per CTA: the step ring of direction words is 16 KB per warp
                    if (C == 2) psa_batch16_kernel<2, true, 2><<<blocks * 2, 64, 0, g.st>>>(a);
                    else psa_batch16_kernel<4, true, 2><<<blocks * 2, 64, 0, g.st>>>(a);
                } else if (C == 4) launch_batch<4, true, kBatchWarps>(prmt, blocks, g.st, a);
                else launch_batch<8, true, kBatchWarps>(prmt, blocks, g.st, a);
                PsaBatchTraceArgs t{};
                t.enc = packed ? (kBatchTM == 3 ? 2 : 1) : 0;      // how psa_batch_trace_kernel reads the codes
                t.dirs = d_dirs; t.dirs_off = d_dirs_off + k0;
                t.len0 = d_len0 + k0; t.len1 = d_len1 + k0; t.npairs = k1 - k0;
                t.cols_per_band = band_cols;
                t.ops = d_ops; t.ops_off = d_ops_off + k0; t.alen = d_alen + k0;
                if (pack2) psa_batch_trace_kernel<true><<<(unsigned)((k1 - k0 + 127) / 128), 128, 0, g.st>>>(t);
                else psa_batch_trace_kernel<false><<<(unsigned)((k1 - k0 + 127) / 128), 128, 0, g.st>>>(t);
            } else if (packed && dual) {
                // short reads: two pairs per warp (psa_sweep_dual16), 32*Cd columns each
                if (Cd == 4) psa_batch16_dual_kernel<4, false, kBatchWarps><<<blocks, 32 * kBatchWarps, 0, g.st>>>(a);
                else psa_batch16_dual_kernel<8, false, kBatchWarps><<<blocks, 32 * kBatchWarps, 0, g.st>>>(a);
            } else if (packed) {
                if (C == 2) psa_batch16_kernel<2, false, kBatchWarps><<<blocks, 32 * kBatchWarps, 0, g.st>>>(a);
                else if (C == 4) psa_batch16_kernel<4, false, kBatchWarps><<<blocks, 32 * kBatchWarps, 0, g.st>>>(a);
                else psa_batch16_kernel<8, false, kBatchWarps><<<blocks, 32 * kBatchWarps, 0, g.st>>>(a);
            } else {
                if (C == 4) launch_batch<4, false, kBatchWarps>(prmt, blocks, g.st, a);
                else launch_batch<8, false, kBatchWarps>(prmt, blocks, g.st, a);
            }
            cudaError_t le = cudaGetLastError();
            TSTA_CUDA_TRY(cudaEventRecord(eb, g.st));
            TSTA_CUDA_TRY(le);
            last_done = eb;
            k0 = k1;
        }
        mark("chunk: kernels queued");
        // ---- results of a finished chunk travel back while the next one is computed --------------
        if (want_trace && nch > 1) {
            int64_t lo = INT64_MAX, hi = 0;
            for (int b = c * (kBlocks / nch); b < (c + 1) * (kBlocks / nch); b++) {
                lo = std::min(lo, bs[b].ops_lo);
                hi = std::max(hi, bs[b].ops_hi);
            }
            if (last_done) TSTA_CUDA_TRY(cudaStreamWaitEvent(g_ctx.dh, last_done, 0));
            if (pack2) { lo >>= 2; hi = (hi + 3) >> 2; }      // (slots start at multiples of four moves: chunks do not share a byte)
            if (hi > lo) TSTA_CUDA_TRY(cudaMemcpyAsync(ops + lo, d_ops + lo, (size_t)(hi - lo), cudaMemcpyDeviceToHost, g_ctx.dh));
        }
    }
    TSTA_CUDA_TRY(cudaMemcpyAsync(maxsorce, d_ms, (size_t)npairs * 4, cudaMemcpyDeviceToHost, g.st));
    TSTA_CUDA_TRY(cudaMemcpyAsync(corner, d_corner, (size_t)npairs * 4, cudaMemcpyDeviceToHost, g.st));
    if (want_trace) {
        TSTA_CUDA_TRY(cudaMemcpyAsync(alen, d_alen, (size_t)npairs * 4, cudaMemcpyDeviceToHost, g.st));
        if (nch == 1) TSTA_CUDA_TRY(cudaMemcpyAsync(ops, d_ops, (size_t)(pack2 ? (ops_total + 3) / 4 : ops_total), cudaMemcpyDeviceToHost, g.st));
    }
    mark("all queued");
    TSTA_CUDA_TRY(cudaStreamSynchronize(g.st));
    mark("kernel stream drained");
    TSTA_CUDA_TRY(cudaStreamSynchronize(g_ctx.dh));
    mark("result copies drained");
    float total_ms = 0.f;
    for (size_t i = 0; i + 1 < nev; i += 2) {
        float ms = 0.f;
        TSTA_CUDA_TRY(cudaEventElapsedTime(&ms, g_ctx.ev[i], g_ctx.ev[i + 1]));
        total_ms += ms;
    }
    if (kernel_ms) *kernel_ms = total_ms;
    g.done = true;
    return TSTA_OK;
}
