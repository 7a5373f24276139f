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
#include "poa_build.cuh"

namespace tsta {
int check_params(const tsta_params* p);
}
using namespace tsta;

struct tsta_poa {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    DevBuf d_rowptr, d_pred, d_base, d_hneg1, d_read, d_HE, d_trace, d_bnd, d_ctl, d_last, d_path, d_desc;
    // device-resident graph (tsta_poa_graph_update / tsta_poa_fill_resident, poa_build.cuh)
    DevBuf d_nmeta, d_npre, d_nsub, d_sort, d_keep, d_recs, d_edge;
    int node_cap = 0;              // nodes the resident table has room for
    bool fixed8 = false;           // the last fill's predecessor table is at stride 8 (no row_ptr)
    long long h2d_bytes = 0;       // graph + read bytes uploaded by the fills of this context
    std::vector<uint4> h_desc;
    std::vector<int> h_stage;      // transport form of changed node records / the packed permutation
    unsigned epoch = 0;
    int nrows = 0, len = 0, nbands = 0, C = kPoaC;
    bool rows_kernel = true;       // last fill ran poa_rows_kernel (row-synchronous) rather than poa_fill_kernel (skewed)
    long long stride = 0;
    bool filled = false;
    int launches = 0;
    double t_desc = 0, t_alloc = 0, t_run = 0, t_update = 0, t_update_first = 0, t_walk = 0;   // host seconds per phase (TSTA_POA_TIMING=1 prints them)
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
                      &p->d_trace, &p->d_bnd, &p->d_ctl, &p->d_last, &p->d_path, &p->d_desc,
                      &p->d_nmeta, &p->d_npre, &p->d_nsub, &p->d_sort, &p->d_keep, &p->d_recs, &p->d_edge};
    for (DevBuf* b : bufs) b->release();
    for (auto& e : p->ev) if (e) cudaEventDestroy(e);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

// Buffers grow with 60 % headroom: the graph gains rows every round and a cudaFree + cudaMalloc pair costs
// milliseconds whatever the size (tools/malloc_time.py), so a 16-read job reallocates once or twice.
static int poa_grow(DevBuf& b, size_t bytes)
{
    if (bytes <= b.cap) return (int)TSTA_OK;
    if (b.reserve(bytes + bytes * 3 / 5 + 4096) == TSTA_OK) return (int)TSTA_OK;
    return b.reserve(bytes);                                    // no room for headroom: exact size
}

// Same, keeping the first `used` bytes (the resident node table).
static int poa_grow_keep(DevBuf& b, size_t bytes, size_t used, cudaStream_t st)
{
    if (bytes <= b.cap) return (int)TSTA_OK;
    DevBuf nb;
    if (nb.reserve(bytes + bytes * 3 / 5 + 4096) != TSTA_OK) {
        if (int rc = nb.reserve(bytes)) return rc;
    }
    if (used > 0 && b.p) {
        cudaError_t e = cudaMemcpyAsync(nb.p, b.p, used, cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { nb.release(); return set_error(TSTA_ECUDA, "node table copy failed: %s", cudaGetErrorString(e)); }
    }
    b.release();
    b = nb;
    return TSTA_OK;
}

// Strip width and kernel for this read (see tsta_poa_fill).
static void poa_choose(tsta_poa* p, int L, int nrows, const char* read, int len)
{
    // Kernel.  poa_rows_kernel (row-synchronous, F by a max-plus scan; poa_rows.cuh) is the default; the skewed
    // poa_fill_kernel keeps the reads with 'N' columns (quirk Q3 is a second horizontal chain).
    // TSTA_POA_KERNEL=skew forces it.
    p->rows_kernel = memchr(read, 'N', (size_t)len) == nullptr;
    if (const char* e = getenv("TSTA_POA_KERNEL")) { if (!strcmp(e, "skew")) p->rows_kernel = false; }
    // Strip width (columns per lane; the kernels are templates).  Rows kernel: 3 -- a 50 kbp read is 521 bands, one
    // per SM sub-partition (592), and the per-row time falls with the strip (4: 526 cycles, 3: ~470, 2: 405 alone
    // but 498 once 782 bands share sub-partitions); skewed kernel: 4 (2 also compiled).  TSTA_POA_C overrides.
    p->C = p->rows_kernel ? 3 : 4;
    if (const char* e = getenv("TSTA_POA_C")) { const int c = atoi(e); if (c == 2 || c == 4 || (c == 3 && p->rows_kernel)) p->C = c; }
    const int band_cols = 32 * p->C;
    p->nbands = (len + band_cols - 1) / band_cols;
    p->stride = (long long)p->nbands * band_cols;
    p->nrows = nrows;
    p->len = len;
}

// Everything after the graph tables are in place: working buffers, the read, the fill launch, lastsorce back.
// pred_tab / hneg1_dev: the predecessor rows the slow path reads and H(r,-1) (NULL: computed by band 0).
static int poa_launch(tsta_poa* p, const tsta_params* prm, int L, int nrows, size_t nkeep, const int* pred_tab, const int* hneg1_dev,
                      const char* read, int len, int32_t* lastsorce, float* kernel_ms)
{
    const double t1 = poa_now();
    if (int rc = poa_grow(p->d_read, (size_t)len)) return rc;
    if (int rc = poa_grow(p->d_last, (size_t)nrows * 4)) return rc;
    if (int rc = p->d_ctl.reserve(64)) return rc;
    if (int rc = poa_grow(p->d_path, (size_t)(nrows + len + 1) * sizeof(PoaStep))) return rc;
    {
        const size_t cells = (size_t)nrows * (size_t)p->stride;
        const size_t spill = std::max<size_t>(nkeep, 1) * (size_t)p->stride * sizeof(int2);
        if (int rc = poa_grow(p->d_HE, spill)) return rc;
        if (int rc = poa_grow(p->d_edge, std::max<size_t>(nkeep, 1) * 4)) return rc;
        if (int rc = poa_grow(p->d_trace, cells * 2)) return rc;
        const size_t need = (size_t)p->nbands * (size_t)nrows * 8;
        if (need > p->d_bnd.cap) {
            if (int rc = poa_grow(p->d_bnd, need)) return rc;
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
    TSTA_CUDA_TRY(cudaMemcpyAsync(p->d_read.p, read, (size_t)len, cudaMemcpyHostToDevice, p->stream));
    p->h2d_bytes += len;
    TSTA_CUDA_TRY(cudaMemsetAsync(p->d_ctl.p, 0, 32, p->stream));

    PoaFillArgs a{};
    a.desc = p->d_desc.as<uint4>();
    a.row_ptr = p->d_rowptr.as<int>(); a.pred_row = pred_tab;
    a.base = p->d_base.as<uint8_t>(); a.hneg1 = hneg1_dev; a.read = p->d_read.as<uint8_t>();
    a.edge_spill = p->d_edge.as<int>();
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
            else if (p->C == 3) poa_rows_kernel<3><<<p->nbands, 32, sizeof(PoaRowsSmem<3>), p->stream>>>(a);
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
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(poa_rows_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PoaRowsSmem<3>));
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
        fprintf(stderr, "tsta_poa host phases: descriptors %.3f s, device buffers %.3f s, copies + kernel + sync %.3f s, graph updates %.3f s (the first one %.3f s; copy + kernel + sync %.3f s); fills: %lld poa_rows_kernel, %lld poa_fill_kernel\n",
                p->t_desc, p->t_alloc, p->t_run, p->t_update, p->t_update_first, p->t_walk, p->n_fill_rows, p->n_fill_skew);
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
    poa_choose(p, L, nrows, read, len);
    p->fixed8 = false;
    const double t1 = poa_now();
    p->t_desc += t1 - t0;
    if (int rc = poa_grow(p->d_desc, (size_t)nrows * sizeof(uint4))) return rc;
    if (int rc = poa_grow(p->d_rowptr, (size_t)(nrows + 1) * 4)) return rc;
    if (int rc = poa_grow(p->d_pred, (size_t)nedges * 4)) return rc;
    if (int rc = poa_grow(p->d_hneg1, (size_t)nrows * 4)) return rc;
    TSTA_CUDA_TRY(cudaMemcpyAsync(p->d_desc.p, p->h_desc.data(), (size_t)nrows * sizeof(uint4), cudaMemcpyHostToDevice, p->stream));
    TSTA_CUDA_TRY(cudaMemcpyAsync(p->d_rowptr.p, row_ptr, (size_t)(nrows + 1) * 4, cudaMemcpyHostToDevice, p->stream));
    TSTA_CUDA_TRY(cudaMemcpyAsync(p->d_pred.p, pred_row, (size_t)nedges * 4, cudaMemcpyHostToDevice, p->stream));
    TSTA_CUDA_TRY(cudaMemcpyAsync(p->d_hneg1.p, hneg1, (size_t)nrows * 4, cudaMemcpyHostToDevice, p->stream));
    p->h2d_bytes += (long long)nrows * (sizeof(uint4) + 8) + 4 + (long long)nedges * 4;
    return poa_launch(p, prm, L, nrows, nkeep, p->d_pred.as<int>(), p->d_hneg1.as<int>(), read, len, lastsorce, kernel_ms);
}

// ---------------------------------------------------------------------------------------
// device-resident graph: changed node records in, row tables built on the device (poa_build.cuh)
// ---------------------------------------------------------------------------------------
int tsta_poa_graph_update(tsta_poa* p, int nnodes, int n_narrow, const tsta_poa_node2* narrow, int n_wide, const tsta_poa_node8* wide)
{
    if (!p || nnodes <= 0 || n_narrow < 0 || n_wide < 0 || (n_narrow > 0 && !narrow) || (n_wide > 0 && !wide))
        return set_error(TSTA_EINVAL, "bad POA graph update");
    static_assert(sizeof(tsta_poa_node2) == sizeof(PoaNodeNarrow) && sizeof(tsta_poa_node8) == sizeof(PoaNodeWide), "node record layout");
    for (int i = 0; i < n_narrow; i++)
        if (narrow[i].id < 0 || narrow[i].id >= nnodes || ((narrow[i].base_in >> 8) & 0xffu) > 2u)
            return set_error(TSTA_EINVAL, "narrow node record %d: id out of range or more than two predecessors", i);
    for (int i = 0; i < n_wide; i++)
        if (wide[i].id < 0 || wide[i].id >= nnodes || ((wide[i].base_in >> 8) & 0xffu) > (unsigned)kPoaNodePreds)
            return set_error(TSTA_EINVAL, "wide node record %d: id out of range or more than eight predecessors", i);
    TSTA_CUDA_TRY(cudaSetDevice(p->device));
    const double t_in = poa_now();
    if (nnodes > p->node_cap) {
        // Node ids are never reused and every read adds one provisional node per base, so the id space of a job
        // grows by a read length per read (16 x 50 kbp: 50 k -> 850 k ids).  The tables are 40 bytes per id: reserve
        // four times what is asked for, at least a million ids -- a job then reallocates them once or twice instead
        // of at every 1.6 x step (cudaMalloc / cudaFree next to a multi-GB trace buffer cost up to 100 ms each on a
        // shared GPU: the outliers of the driver loop).
        const size_t used = (size_t)p->node_cap;
        const size_t want = std::max<size_t>((size_t)nnodes * 4, (size_t)1 << 20);
        if (int rc = poa_grow_keep(p->d_nmeta, want * 4, used * 4, p->stream)) return rc;
        if (int rc = poa_grow_keep(p->d_npre, want * 4 * kPoaNodePreds, used * 4 * kPoaNodePreds, p->stream)) return rc;
        if (int rc = poa_grow_keep(p->d_nsub, want * 4, 0, p->stream)) return rc;
        p->node_cap = (int)std::min<size_t>(p->d_nmeta.cap / 4, std::min(p->d_npre.cap / (4 * kPoaNodePreds), p->d_nsub.cap / 4));
    }
    // transport form (poa_build.cuh): 12-byte records for in-degree <= 1, 16 bytes for 2, 40 bytes up to 8
    std::vector<int>& st = p->h_stage;
    st.clear();
    int n1 = 0, n2 = 0;
    for (int i = 0; i < n_narrow; i++)
        if (((narrow[i].base_in >> 8) & 0xffu) <= 1u) { st.push_back(narrow[i].id); st.push_back((int)narrow[i].base_in); st.push_back(narrow[i].pre[0]); n1++; }
    for (int i = 0; i < n_narrow; i++)
        if (((narrow[i].base_in >> 8) & 0xffu) == 2u) {
            st.push_back(narrow[i].id); st.push_back((int)narrow[i].base_in); st.push_back(narrow[i].pre[0]); st.push_back(narrow[i].pre[1]); n2++;
        }
    for (int i = 0; i < n_wide; i++) {
        st.push_back(wide[i].id); st.push_back((int)wide[i].base_in);
        for (int j = 0; j < kPoaNodePreds; j++) st.push_back(wide[i].pre[j]);
    }
    const size_t bytes = st.size() * sizeof(int);
    if (bytes == 0) return TSTA_OK;
    const double t_staged = poa_now();
    if (int rc = poa_grow(p->d_recs, bytes)) return rc;
    PoaNodeTable t{p->d_nmeta.as<unsigned>(), p->d_npre.as<int>(), p->d_nsub.as<int>()};
    TSTA_CUDA_TRY(cudaMemcpyAsync(p->d_recs.p, st.data(), bytes, cudaMemcpyHostToDevice, p->stream));
    const int nrec = n1 + n2 + n_wide;
    poa_apply_records_kernel<<<(nrec + 255) / 256, 256, 0, p->stream>>>(t, p->d_recs.as<int>(), n1, n2, n_wide);
    TSTA_CUDA_TRY(cudaGetLastError());
    // (the caller's records may be reused as soon as this returns)
    TSTA_CUDA_TRY(cudaStreamSynchronize(p->stream));
    p->h2d_bytes += (long long)bytes;
    if (p->t_update == 0) p->t_update_first = poa_now() - t_in;
    p->t_update += poa_now() - t_in;
    p->t_walk += poa_now() - t_staged;      // (TSTA_POA_TIMING: copy + kernel + sync share of the updates)
    return TSTA_OK;
}

int tsta_poa_fill_resident(tsta_poa* p, const tsta_params* prm, int L, int nrows, const int32_t* sort, const int32_t* hneg1,
                           const char* read, int len, int32_t* lastsorce, float* kernel_ms)
{
    if (kernel_ms) *kernel_ms = 0.f;
    if (!p) return set_error(TSTA_EINVAL, "context is NULL");
    if (int rc = check_params(prm)) return rc;
    if (nrows <= 0 || len <= 0 || !sort || !read || !lastsorce) return set_error(TSTA_EINVAL, "bad POA arguments");
    if (p->node_cap <= 0) return set_error(TSTA_EINVAL, "tsta_poa_fill_resident before tsta_poa_graph_update");
    {
        const long long amax = std::max(std::max(std::abs(prm->M), std::abs(prm->X)), std::abs(prm->E) + std::abs(prm->O));
        if (((long long)nrows + len + 2) * amax >= (1LL << 22))
            return set_error(TSTA_EINVAL, "POA job too large for the int32 key packing: (rows + read length) * max|score| must stay below 2^22");
    }
    for (int r = 0; r < nrows; r++)
        if (sort[r] < 0 || sort[r] >= p->node_cap) return set_error(TSTA_EINVAL, "row %d: node id %d is not in the resident table", r, sort[r]);
    TSTA_CUDA_TRY(cudaSetDevice(p->device));
    p->filled = false;
    const double t0 = poa_now();
    poa_choose(p, L, nrows, read, len);
    if (!p->rows_kernel && !hneg1)
        return set_error(TSTA_EINVAL, "this read runs on the skewed kernel ('N' columns): it needs hneg1");
    p->fixed8 = true;
    if (int rc = poa_grow(p->d_desc, (size_t)nrows * sizeof(uint4))) return rc;
    if (int rc = poa_grow(p->d_sort, (size_t)nrows * 4)) return rc;
    if (int rc = poa_grow(p->d_keep, (size_t)nrows * 4)) return rc;
    if (int rc = poa_grow(p->d_pred, (size_t)nrows * 4 * kPoaNodePreds)) return rc;
    if (int rc = p->d_ctl.reserve(64)) return rc;
    // the permutation crosses PCIe as 3-byte ids
    const int id_bytes = p->node_cap < (1 << 24) ? 3 : 4;
    if (int rc = poa_grow(p->d_recs, (size_t)nrows * id_bytes + 4)) return rc;
    p->h_stage.resize(((size_t)nrows * id_bytes + 3) / 4 + 1);
    {
        uint8_t* q = reinterpret_cast<uint8_t*>(p->h_stage.data());
        for (int r = 0; r < nrows; r++, q += id_bytes) {
            const unsigned id = (unsigned)sort[r];
            q[0] = (uint8_t)id; q[1] = (uint8_t)(id >> 8); q[2] = (uint8_t)(id >> 16);
            if (id_bytes == 4) q[3] = (uint8_t)(id >> 24);
        }
    }
    TSTA_CUDA_TRY(cudaMemcpyAsync(p->d_recs.p, p->h_stage.data(), (size_t)nrows * id_bytes, cudaMemcpyHostToDevice, p->stream));
    p->h2d_bytes += (long long)nrows * id_bytes;
    const int* hneg1_dev = nullptr;
    if (!p->rows_kernel) {
        if (int rc = poa_grow(p->d_hneg1, (size_t)nrows * 4)) return rc;
        TSTA_CUDA_TRY(cudaMemcpyAsync(p->d_hneg1.p, hneg1, (size_t)nrows * 4, cudaMemcpyHostToDevice, p->stream));
        p->h2d_bytes += (long long)nrows * 4;
        hneg1_dev = p->d_hneg1.as<int>();
    }
    TSTA_CUDA_TRY(cudaMemsetAsync(p->d_ctl.as<int>() + 8, 0, 4, p->stream));
    PoaNodeTable t{p->d_nmeta.as<unsigned>(), p->d_npre.as<int>(), p->d_nsub.as<int>()};
    const int nb = (nrows + 255) / 256;
    {
        NvtxRange range("tsta_poa build rows");
        poa_rows_of_nodes_kernel<<<nb, 256, 0, p->stream>>>(t, p->d_recs.as<uint8_t>(), id_bytes, p->d_sort.as<int>(), nrows, p->d_keep.as<int>());
        poa_build_rows_kernel<<<nb, 256, 0, p->stream>>>(t, p->d_sort.as<int>(), nrows, L, p->d_desc.as<uint4>(), p->d_pred.as<int>(), p->d_keep.as<int>());
        poa_build_keep_kernel<<<nb, 256, 0, p->stream>>>(nrows, p->d_desc.as<uint4>(), p->d_keep.as<int>(), p->d_ctl.as<int>() + 8);
    }
    TSTA_CUDA_TRY(cudaGetLastError());
    int nkeep = 0;     // sizes the spill: a handful of rows per job
    TSTA_CUDA_TRY(cudaMemcpyAsync(&nkeep, p->d_ctl.as<int>() + 8, 4, cudaMemcpyDeviceToHost, p->stream));
    TSTA_CUDA_TRY(cudaStreamSynchronize(p->stream));
    p->n_keep += nkeep;
    p->n_rows += nrows;
    p->t_desc += poa_now() - t0;
    return poa_launch(p, prm, L, nrows, (size_t)nkeep, p->d_pred.as<int>(), hneg1_dev, read, len, lastsorce, kernel_ms);
}

long long tsta_poa_h2d_bytes(const tsta_poa* p) { return p ? p->h2d_bytes : 0; }

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
    w.desc = p->d_desc.as<uint4>(); w.read = p->d_read.as<uint8_t>();
    w.fixed8 = p->fixed8 ? 1 : 0;
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
