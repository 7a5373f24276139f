// tsta_poa.cu -- C-ABI host side of the POA path (include/tsta_cuda.h: tsta_poa_*).
//
// Plays the role of control() between readseq() and node_fuse() (reference
// msa/c-t-simd.c:979-1047): the per-node trace allocation (1003-1017) becomes one dense
// device matrix, the tile loop (1019-1043) one kernel launch, and the trace walk of
// node_fuse() (751-785) a second kernel that returns the compact path.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "poa_fill.cuh"
#include "poa_rows.cuh"

namespace tsta {
int check_params(const tsta_params* p);
}
using namespace tsta;

struct tsta_poa {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    DevBuf d_rowptr, d_pred, d_base, d_hneg1, d_read, d_HE, d_trace, d_bnd, d_ctl, d_last, d_path, d_desc;
    std::vector<uint4> h_desc;
    unsigned epoch = 0;
    int nrows = 0, len = 0, nbands = 0, C = kPoaC;
    bool rows_kernel = true;       // last fill ran poa_rows_kernel (row-synchronous) rather than poa_fill_kernel (skewed)
    long long stride = 0;
    bool filled = false;
    int launches = 0;
    double t_desc = 0, t_alloc = 0, t_run = 0;   // host seconds per phase (TSTA_POA_TIMING=1 prints them)
    long long n_fill_rows = 0, n_fill_skew = 0;   // fills by kernel
    long long n_rows = 0, n_wide = 0, n_far = 0, n_keep = 0, n_deg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

static double poa_now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static void poa_free(tsta_poa* p)
{
    if (!p) return;
    cudaSetDevice(p->device);
    DevBuf* bufs[] = {&p->d_rowptr, &p->d_pred, &p->d_base, &p->d_hneg1, &p->d_read, &p->d_HE,
                      &p->d_trace, &p->d_bnd, &p->d_ctl, &p->d_last, &p->d_path, &p->d_desc};
    for (DevBuf* b : bufs) b->release();
    for (auto& e : p->ev) if (e) cudaEventDestroy(e);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

extern "C" {

int tsta_poa_create(tsta_poa** out)
{
    if (!out) return set_error(TSTA_EINVAL, "context pointer is NULL");
    *out = nullptr;
    if (int rc = ensure_device()) return rc;
    tsta_poa* p = new (std::nothrow) tsta_poa();
    if (!p) return set_error(TSTA_ENOMEM, "out of host memory");
    cudaError_t e = cudaGetDevice(&p->device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
    for (int k = 0; k < 2 && e == cudaSuccess; k++) e = cudaEventCreate(&p->ev[k]);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(poa_fill_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PoaSmemC<4>));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(poa_fill_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PoaSmemC<2>));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(poa_rows_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PoaRowsSmem<4>));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(poa_rows_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PoaRowsSmem<2>));
    if (e != cudaSuccess) {
        poa_free(p);
        return set_error(TSTA_ECUDA, "POA context creation failed: %s", cudaGetErrorString(e));
    }
    *out = p;
    return TSTA_OK;
}

void tsta_poa_destroy(tsta_poa* p)
{
    if (p && getenv("TSTA_POA_TIMING"))
        fprintf(stderr, "tsta_poa host phases: descriptors %.3f s, device buffers %.3f s, copies + kernel + sync %.3f s; fills: %lld poa_rows_kernel, %lld poa_fill_kernel\n",
                p->t_desc, p->t_alloc, p->t_run, p->n_fill_rows, p->n_fill_skew);
    if (p && getenv("TSTA_POA_TIMING"))
        fprintf(stderr, "tsta_poa rows %lld: in-degree>8 %lld, with a far predecessor %lld, spilled %lld; in-degree histogram 1:%lld 2:%lld 3:%lld 4:%lld 5:%lld 6:%lld 7+:%lld\n",
                p->n_rows, p->n_wide, p->n_far, p->n_keep, p->n_deg[1], p->n_deg[2], p->n_deg[3], p->n_deg[4], p->n_deg[5], p->n_deg[6], p->n_deg[7]);
    poa_free(p);
}

int tsta_poa_fill(tsta_poa* p, const tsta_params* prm, int L, int nrows,
                  const int32_t* row_ptr, const int32_t* pred_row, const uint8_t* base, const int32_t* hneg1,
                  const char* read, int len, int32_t* lastsorce, float* kernel_ms)
{
    if (kernel_ms) *kernel_ms = 0.f;
    if (!p) return set_error(TSTA_EINVAL, "context is NULL");
    if (int rc = check_params(prm)) return rc;
    if (nrows <= 0 || len <= 0 || !row_ptr || !pred_row || !base || !hneg1 || !read || !lastsorce)
        return set_error(TSTA_EINVAL, "bad POA arguments");
    {   // the kernel carries scores times 128 (key packing, poa_fill.cuh)
        const long long amax = std::max(std::max(std::abs(prm->M), std::abs(prm->X)), std::abs(prm->E) + std::abs(prm->O));
        if (((long long)nrows + len + 2) * amax >= (1LL << 22))
            return set_error(TSTA_EINVAL, "POA job too large for the int32 key packing: (rows + read length) * max|score| must stay below 2^22");
    }
    const int nedges = row_ptr[nrows];
    for (int r = 0; r < nrows; r++)
        if (row_ptr[r + 1] - row_ptr[r] < 1 || row_ptr[r + 1] - row_ptr[r] > 41)
            return set_error(TSTA_EINVAL, "row %d: in-degree must be 1..41 (reference limit: predecessor index < 42)", r);
    TSTA_CUDA_TRY(cudaSetDevice(p->device));
    p->filled = false;
    const double t0 = poa_now();
    // row descriptors (see poa_fill.cuh): base, in-degree, flags, history slots of up to three predecessors
    p->h_desc.assign((size_t)nrows, make_uint4(0, 0, 0, 0));
    for (int r = 0; r < nrows; r++) {
        const int e0 = row_ptr[r], np = row_ptr[r + 1] - e0;
        unsigned flags = 0, jf = kPoaNoFwd, slot[kPoaFast] = {kPoaSlotNone, kPoaSlotNone, kPoaSlotNone, kPoaSlotNone,
                                              kPoaSlotNone, kPoaSlotNone, kPoaSlotNone, kPoaSlotNone};
        static_assert(kPoaFast == 8, "descriptor words y and z hold eight slot bytes");
        if (pred_row[e0] < 0) {
            flags |= kPoaHead;
            if (L > 0 && r >= L) flags |= kPoaLate;                 // quirk Q1
            slot[0] = (flags & kPoaLate) ? kPoaSlotLate : kPoaSlotHead;
            if (np != 1) return set_error(TSTA_EINVAL, "row %d: a head must have the single entry -1", r);
        } else {
            if (np > kPoaFast) { flags |= kPoaSlow; p->n_wide++; }
            bool far = false;
            for (int j = 0; j < np; j++) {
                const int pr = pred_row[e0 + j];
                if (pr < 0 || pr >= r) return set_error(TSTA_EINVAL, "row %d: predecessor %d is not an earlier row", r, pr);
                const int dist = r - pr;
                if (dist > kPoaNear) {
                    p->h_desc[(size_t)pr].x |= kPoaKeep << 16;      // it must also spill its keys to HBM
                    flags |= kPoaSlow;
                    far = true;
                }
                if (j < kPoaFast) slot[j] = (unsigned)(pr & (kPoaHist - 1));
                if (pr == r - 1 && jf != kPoaNoFwd) flags |= kPoaSlow;   // (a duplicated edge: generic loop reads the ring)
                if (j < kPoaFast && pr == r - 1) jf = (unsigned)j;     // poa_rows_kernel takes this one from registers
            }
            if (far) p->n_far++;
        }
        p->n_deg[np < 7 ? np : 7]++;
        p->n_rows++;
        uint4& q = p->h_desc[(size_t)r];
        q.x |= (unsigned)base[r] | ((unsigned)np << 8) | (flags << 16) | (jf << 28);
        q.y = slot[0] | (slot[1] << 8) | (slot[2] << 16) | (slot[3] << 24);
        q.z = (flags & kPoaSlow) ? (unsigned)e0 : (slot[4] | (slot[5] << 8) | (slot[6] << 16) | (slot[7] << 24));
    }
    // rows with a successor further away than the history ring get a row of the HBM spill
    size_t nkeep = 0;
    for (int r = 0; r < nrows; r++) {
        uint4& q = p->h_desc[(size_t)r];
        q.w = ((q.x >> 16) & kPoaKeep) ? (unsigned)nkeep++ : 0u;
        // poa_rows_kernel: the only predecessor is row r-1 and nothing else is special -> register-only row
        if (((q.x >> 8) & 0xffu) == 1u && (q.x >> 28) == 0u && !((q.x >> 16) & (kPoaHead | kPoaKeep | kPoaSlow | kPoaLate)))
            q.x |= kPoaPlainRow << 16;
    }
    p->n_keep += (long long)nkeep;
    // Strip width.  The kernel is a template (2 or 4 columns per lane).  Measured on 16 x 50 kbp: 4 columns (391 bands)
    // fill in 398 ms, 2 columns (782 bands, two per SM sub-partition) in 445 ms -- the fixed cost of a step (shuffle,
    // history round trip, descriptor) does not shrink with the strip, and twice the bands double the start-up.
    p->C = 4;
    if (const char* e = getenv("TSTA_POA_C")) { const int c = atoi(e); if (c == 2 || c == 4) p->C = c; }
    // Kernel.  poa_rows_kernel (row-synchronous, F by a max-plus scan; poa_rows.cuh) is the default; the skewed
    // poa_fill_kernel keeps the reads with 'N' columns (quirk Q3 is a second horizontal chain) and strip widths
    // that do not divide L (the late-head re-base must fall on a lane's first column).  TSTA_POA_KERNEL=skew forces it.
    p->rows_kernel = (L % p->C == 0) && memchr(read, 'N', (size_t)len) == nullptr;
    if (const char* e = getenv("TSTA_POA_KERNEL")) { if (!strcmp(e, "skew")) p->rows_kernel = false; }
    const int band_cols = 32 * p->C;
    p->nbands = (len + band_cols - 1) / band_cols;
    p->stride = (long long)p->nbands * band_cols;
    p->nrows = nrows;
    p->len = len;
    const double t1 = poa_now();
    p->t_desc += t1 - t0;

    // The graph gains rows every round and a cudaFree + cudaMalloc pair costs milliseconds whatever
    // the size (tools/malloc_time.py), so every buffer grows with 60 % headroom: a 16-read job
    // reallocates once or twice instead of on every call.
    auto grow = [](DevBuf& b, size_t bytes) {
        if (bytes <= b.cap) return (int)TSTA_OK;
        if (b.reserve(bytes + bytes * 3 / 5 + 4096) == TSTA_OK) return (int)TSTA_OK;
        return b.reserve(bytes);                                    // no room for headroom: exact size
    };
    if (int rc = grow(p->d_desc, (size_t)nrows * sizeof(uint4))) return rc;
    if (int rc = grow(p->d_rowptr, (size_t)(nrows + 1) * 4)) return rc;
    if (int rc = grow(p->d_pred, (size_t)nedges * 4)) return rc;
    if (int rc = grow(p->d_base, (size_t)nrows)) return rc;
    if (int rc = grow(p->d_hneg1, (size_t)nrows * 4)) return rc;
    if (int rc = grow(p->d_read, (size_t)len)) return rc;
    if (int rc = grow(p->d_last, (size_t)nrows * 4)) return rc;
    if (int rc = p->d_ctl.reserve(64)) return rc;
    if (int rc = grow(p->d_path, (size_t)(nrows + len + 1) * sizeof(PoaStep))) return rc;
    {   // grow generously: the graph gains rows every round
        const size_t cells = (size_t)nrows * (size_t)p->stride;
        const size_t spill = std::max<size_t>(nkeep, 1) * (size_t)p->stride * sizeof(int2);
        if (int rc = grow(p->d_HE, spill)) return rc;
        if (int rc = grow(p->d_trace, cells * 2)) return rc;
        const size_t need = (size_t)p->nbands * (size_t)nrows * 8;
        if (need > p->d_bnd.cap) {
            if (int rc = grow(p->d_bnd, need)) return rc;
            TSTA_CUDA_TRY(cudaMemsetAsync(p->d_bnd.p, 0, p->d_bnd.cap, p->stream));
            p->epoch = 0;
        }
    }
    if (++p->epoch > 0x7fffu) {
        TSTA_CUDA_TRY(cudaMemsetAsync(p->d_bnd.p, 0, p->d_bnd.cap, p->stream));
        p->epoch = 1;
    }
    const double t2 = poa_now();
    p->t_alloc += t2 - t1;
    TSTA_CUDA_TRY(cudaMemcpyAsync(p->d_desc.p, p->h_desc.data(), (size_t)nrows * sizeof(uint4), cudaMemcpyHostToDevice, p->stream));
    TSTA_CUDA_TRY(cudaMemcpyAsync(p->d_rowptr.p, row_ptr, (size_t)(nrows + 1) * 4, cudaMemcpyHostToDevice, p->stream));
    TSTA_CUDA_TRY(cudaMemcpyAsync(p->d_pred.p, pred_row, (size_t)nedges * 4, cudaMemcpyHostToDevice, p->stream));
    TSTA_CUDA_TRY(cudaMemcpyAsync(p->d_base.p, base, (size_t)nrows, cudaMemcpyHostToDevice, p->stream));
    TSTA_CUDA_TRY(cudaMemcpyAsync(p->d_hneg1.p, hneg1, (size_t)nrows * 4, cudaMemcpyHostToDevice, p->stream));
    TSTA_CUDA_TRY(cudaMemcpyAsync(p->d_read.p, read, (size_t)len, cudaMemcpyHostToDevice, p->stream));
    TSTA_CUDA_TRY(cudaMemsetAsync(p->d_ctl.p, 0, 64, p->stream));

    PoaFillArgs a{};
    a.desc = p->d_desc.as<uint4>();
    a.row_ptr = p->d_rowptr.as<int>(); a.pred_row = p->d_pred.as<int>();
    a.base = p->d_base.as<uint8_t>(); a.hneg1 = p->d_hneg1.as<int>(); a.read = p->d_read.as<uint8_t>();
    a.len = len; a.nrows = nrows; a.nbands = p->nbands; a.L = L;
    a.M = prm->M; a.X = prm->X; a.E = prm->E; a.O = prm->O;
    a.HE = p->d_HE.as<int2>(); a.stride = p->stride;
    a.trace = p->d_trace.as<uint16_t>();
    a.bnd = p->d_bnd.as<unsigned long long>(); a.epoch = p->epoch;
    a.ticket = p->d_ctl.as<unsigned>();
    a.lastsorce = p->d_last.as<int>();
    TSTA_CUDA_TRY(cudaEventRecord(p->ev[0], p->stream));
    {
        NvtxRange range("tsta_poa fill");
        (p->rows_kernel ? p->n_fill_rows : p->n_fill_skew)++;
        if (p->rows_kernel) {
            if (p->C == 2) poa_rows_kernel<2><<<p->nbands, 32, sizeof(PoaRowsSmem<2>), p->stream>>>(a);
            else poa_rows_kernel<4><<<p->nbands, 32, sizeof(PoaRowsSmem<4>), p->stream>>>(a);
        } else if (p->C == 2) poa_fill_kernel<2><<<p->nbands, 32, sizeof(PoaSmemC<2>), p->stream>>>(a);
        else poa_fill_kernel<4><<<p->nbands, 32, sizeof(PoaSmemC<4>), p->stream>>>(a);
    }
    TSTA_CUDA_TRY(cudaGetLastError());
    TSTA_CUDA_TRY(cudaEventRecord(p->ev[1], p->stream));
    TSTA_CUDA_TRY(cudaMemcpyAsync(lastsorce, p->d_last.p, (size_t)nrows * 4, cudaMemcpyDeviceToHost, p->stream));
    TSTA_CUDA_TRY(cudaStreamSynchronize(p->stream));
    if (kernel_ms) TSTA_CUDA_TRY(cudaEventElapsedTime(kernel_ms, p->ev[0], p->ev[1]));
    p->filled = true;
    p->launches = 1;
    p->t_run += poa_now() - t2;
    return TSTA_OK;
}

int tsta_poa_walk(tsta_poa* p, int start_row, tsta_poa_step* path, int* nsteps, float* kernel_ms)
{
    if (kernel_ms) *kernel_ms = 0.f;
    if (!p || !p->filled || !path || !nsteps) return set_error(TSTA_EINVAL, "POA walk before fill, or NULL argument");
    if (start_row < 0 || start_row >= p->nrows) return set_error(TSTA_EINVAL, "start row out of range");
    static_assert(sizeof(tsta_poa_step) == sizeof(PoaStep), "step layout");
    TSTA_CUDA_TRY(cudaSetDevice(p->device));
    PoaWalkArgs w{};
    w.trace = p->d_trace.as<uint16_t>(); w.stride = p->stride;
    w.row_ptr = p->d_rowptr.as<int>(); w.pred_row = p->d_pred.as<int>();
    w.base = p->d_base.as<uint8_t>(); w.read = p->d_read.as<uint8_t>();
    w.len = p->len; w.nrows = p->nrows; w.start_row = start_row;
    w.path = p->d_path.as<PoaStep>();
    w.nsteps = p->d_ctl.as<int>() + 4;
    TSTA_CUDA_TRY(cudaEventRecord(p->ev[0], p->stream));
    { NvtxRange range("tsta_poa walk"); poa_walk_kernel<<<1, kPoaWalkThreads, 0, p->stream>>>(w); }
    TSTA_CUDA_TRY(cudaGetLastError());
    TSTA_CUDA_TRY(cudaEventRecord(p->ev[1], p->stream));
    int n = 0;
    TSTA_CUDA_TRY(cudaMemcpyAsync(&n, w.nsteps, 4, cudaMemcpyDeviceToHost, p->stream));
    TSTA_CUDA_TRY(cudaStreamSynchronize(p->stream));
    if (n < 0 || n > p->nrows + p->len) return set_error(TSTA_ECUDA, "walker returned an impossible path length %d", n);
    if (n > 0) TSTA_CUDA_TRY(cudaMemcpy(path, p->d_path.p, (size_t)n * sizeof(PoaStep), cudaMemcpyDeviceToHost));
    *nsteps = n;
    if (kernel_ms) TSTA_CUDA_TRY(cudaEventElapsedTime(kernel_ms, p->ev[0], p->ev[1]));
    p->launches = 2;
    return TSTA_OK;
}

int tsta_poa_debug_trace(tsta_poa* p, uint16_t* out_rows_by_len)
{
    if (!p || !p->filled || !out_rows_by_len) return set_error(TSTA_EINVAL, "POA trace export before fill");
    TSTA_CUDA_TRY(cudaSetDevice(p->device));
    TSTA_CUDA_TRY(cudaMemcpy2D(out_rows_by_len, (size_t)p->len * 2, p->d_trace.p, (size_t)p->stride * 2,
                               (size_t)p->len * 2, (size_t)p->nrows, cudaMemcpyDeviceToHost));
    return TSTA_OK;
}

}  // extern "C"
