// tsta_psa.cu -- C-ABI host side of the pairwise path (include/tsta_cuda.h).
//
// Plays the role of the reference main() between readseq() and printf("maxsorce")
// (psa/psa.c:588-666): state set-up (blockmatrix_init, 219-232, happens inside the fill
// kernel), the fill, the traceback and the '-'-padded strings of trace() (467-515).
#include <algorithm>
#include <cmath>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "psa_fill.cuh"
#include "psa_fill16.cuh"
#include "psa_trace.cuh"
#include "psa_chase.cuh"

namespace tsta {

thread_local std::string g_last_error;
long long* g_trace_dbg = nullptr;   // TSTA_TRACE_DEBUG=1: cycle counters of the traceback kernel

int ensure_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return set_error(TSTA_ENODEV, "no CUDA device available (%s); tsta_b200 has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    return TSTA_OK;
}

int check_params(const tsta_params* p)
{
    if (!p) return set_error(TSTA_EINVAL, "params is NULL");
    if (p->O > 0 || p->E > 0 || p->X > 0)
        return set_error(TSTA_EINVAL, "parameters outside the parity domain: need O<=0, E<=0, X<=0 (got O=%d E=%d X=%d)",
                         p->O, p->E, p->X);
    const int lim = 127;
    if (abs(p->M) > lim || abs(p->X) > lim || abs(p->E) > lim || abs(p->O) > lim)
        return set_error(TSTA_EINVAL, "parameters outside the parity domain: |M|,|X|,|E|,|O| must be <= 127");
    return TSTA_OK;
}

// ---------------------------------------------------------------------------------------
// alphabet: which byte values occur (device-side, so batches of 1e6 reads cost microseconds)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) byte_presence_kernel(const uint8_t* p, long long nbytes, unsigned* present)
{
    __shared__ unsigned s_present[8];
    if (threadIdx.x < 8) s_present[threadIdx.x] = 0;
    __syncthreads();
    unsigned loc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += stride) {
        const unsigned b = p[i];
#pragma unroll
        for (int k = 0; k < 8; k++) loc[k] |= ((b >> 5) == (unsigned)k) ? (1u << (b & 31)) : 0u;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        unsigned v = loc[k];
        for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(kFullMask, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicOr(&s_present[k], v);
    }
    __syncthreads();
    if (threadIdx.x < 8 && s_present[threadIdx.x]) atomicOr(&present[threadIdx.x], s_present[threadIdx.x]);
}

// Builds the byte -> code LUT from the presence bitmap; returns the number of symbols.
int build_lut(const unsigned present[8], uint8_t lut[256])
{
    int K = 0;
    for (int b = 0; b < 256; b++) {
        lut[b] = 0;
        if (present[b >> 5] & (1u << (b & 31))) lut[b] = (uint8_t)std::min(K++, 255);
    }
    return K;
}

// Host wrapper so other translation units can use the kernel without -rdc.
// `present` points at 8 zero-initialised unsigned words in device memory.
cudaError_t launch_byte_presence(const uint8_t* p, long long nbytes, unsigned* present, int blocks, cudaStream_t st)
{
    if (nbytes > 0) byte_presence_kernel<<<blocks, 256, 0, st>>>(p, nbytes, present);
    return cudaGetLastError();
}

}  // namespace tsta

using namespace tsta;

// ---------------------------------------------------------------------------------------
// plan: one long pair, device resident
// ---------------------------------------------------------------------------------------
struct tsta_psa_plan {
    int want_trace = 0;
    int device = 0;
    int sms = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    PsaScoring sc{};
    int64_t n = 0, m = 0;
    int n_pad = 0, nbands = 0, C = 4, CH = 16;
    bool prmt = true;
    bool packed = false;           // int16 x 2 kernels (psa_fill16.cuh); else int32 (psa_fill.cuh)
    int scale = 1;                 // score scale of the device results (1 score-only, 2 int32 traced, 4 packed traced)
    bool skew = false;             // traced + packed: skewed direction layout, 2-bit moves (psa_trace_skew_kernel)
    // traced + packed + whole pair on this GPU: checkpointed traceback (psa_fill16.cuh) -- score-only fill that
    // leaves checkpoints, then the traced form of the tiles in a corridor around the path, then the walk
    bool ckpt = false;
    int64_t corridor_w = 2048;     // columns added on both sides of the expected path
    long long max_slots = 0;       // tile budget of one corridor
    long long slots = 0;           // tiles of the current corridor
    int rounds = 0;                // corridors the last run needed (1 = the walk never left the first one)
    long long tiles_run = 0;       // tiles re-run by the last run, all corridors
    bool chase = true;             // parallel traceback over the tiles (psa_chase.cuh); false: the serial window walker
    DevBuf d_ckpt, d_tilerows, d_summary, d_visits, d_counts, d_vslots;
    int max_visits = 0;
    PsaTileRow* h_rows = nullptr;  // pinned, [rows_cap]
    size_t rows_cap = 0;
    // column slice of a pair spread over several GPUs (tsta_psa_plan_upload_slice)
    int64_t col0 = 0, n_total = 0;
    bool has_left = false, has_right = false;
    DevBuf d_inbox;                // [m] tagged boundary elements written by the left neighbour GPU
    size_t inbox_rows = 0;
    unsigned long long* outbox = nullptr;   // the right neighbour's inbox, opened through CUDA IPC
    bool outbox_open = false;
    int nsym = 0;
    bool uploaded = false, ran = false;
    int launches = 0;
    unsigned epoch = 0;            // boundary tag of the last run (see psa_fill.cuh: bnd_pack)
    std::vector<char> h_s0, h_s1;  // host copies for string emission
    DevBuf d_s0, d_s1, d_bnd, d_ctl, d_dirs, d_ops, d_lastrow, d_timeline;
    // d_ctl layout (ints): [0] ticket, [1] out_max, [2..4] trace out, [8..15] presence
    int* h_small = nullptr;  // pinned, 16 ints
};

static void plan_free(tsta_psa_plan* p)
{
    if (!p) return;
    cudaSetDevice(p->device);
    p->d_s0.release(); p->d_s1.release(); p->d_bnd.release(); p->d_ctl.release();
    p->d_dirs.release(); p->d_ops.release(); p->d_lastrow.release(); p->d_inbox.release(); p->d_timeline.release();
    p->d_ckpt.release(); p->d_tilerows.release();
    p->d_summary.release(); p->d_visits.release(); p->d_counts.release(); p->d_vslots.release();
    if (p->h_rows) cudaFreeHost(p->h_rows);
    if (p->outbox_open) cudaIpcCloseMemHandle(p->outbox);
    for (auto& e : p->ev) if (e) cudaEventDestroy(e);
    if (p->stream) cudaStreamDestroy(p->stream);
    if (p->h_small) cudaFreeHost(p->h_small);
    delete p;
}

template <int C, bool TRACE>
static void launch_pair(const tsta_psa_plan* p, const PsaPairArgs& fa)
{
    // hand-off chunk: 16 rows keeps the inter-band lag short; 32 is kept for experiments
    if (p->CH == 32) {
        if (p->prmt) psa_pair_kernel<C, TRACE, true, 32><<<fa.nbands, 32, 0, p->stream>>>(fa);
        else psa_pair_kernel<C, TRACE, false, 32><<<fa.nbands, 32, 0, p->stream>>>(fa);
    } else if (p->CH == 8) {
        if (p->prmt) psa_pair_kernel<C, TRACE, true, 8><<<fa.nbands, 32, 0, p->stream>>>(fa);
        else psa_pair_kernel<C, TRACE, false, 8><<<fa.nbands, 32, 0, p->stream>>>(fa);
    } else {
        if (p->prmt) psa_pair_kernel<C, TRACE, true, 16><<<fa.nbands, 32, 0, p->stream>>>(fa);
        else psa_pair_kernel<C, TRACE, false, 16><<<fa.nbands, 32, 0, p->stream>>>(fa);
    }
}

template <int C, int TM>
static void launch_pair16(const tsta_psa_plan* p, const PsaPairArgs& fa)
{
    // 32-row chunks (re-centring, row tables); the left boundary is imported in groups of 8 rows
    if (p->CH == 16) psa_pair16_kernel<C, TM, 16><<<fa.nbands, 32, 0, p->stream>>>(fa);
    else psa_pair16_kernel<C, TM, 32><<<fa.nbands, 32, 0, p->stream>>>(fa);
}

template <int C>
static cudaError_t launch_trace_skew(const PsaTraceArgs& ta, cudaStream_t st)
{
    // (function attributes are per device: set it on every launch, it costs nothing)
    cudaError_t e = cudaFuncSetAttribute(psa_trace_skew_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, PsaSkewGeom<C>::kSmemBytes);
    if (e != cudaSuccess) return e;
    psa_trace_skew_kernel<C><<<1, kSkewThreads, PsaSkewGeom<C>::kSmemBytes, st>>>(ta);
    return cudaGetLastError();
}

// Packed kernels: 64*C columns per band; the smallest strip that keeps one band per SM
// sub-partition, as below (traced pairs too: the skewed direction layout takes any strip width).
static int choose_strip16(int64_t n, int sms)
{
    const int64_t slots = (int64_t)sms * 4;
    for (int c : {2, 3, 4, 6}) if (n <= 64 * c * slots) return c;
    return 8;
}

// Column slices over several GPUs: the strip width that minimises the time model of DESIGN.md 5,
//     T(C) = tau(C) * (rows + bands of the WHOLE pair * lag),   tau(C) = 39 + 11 C cycles,   lag = 118 steps,
// among the widths that give this GPU at most one band per sub-partition.  (The slice's own length alone would pick
// narrow strips -- fewer columns per GPU -- and with them more bands in the chain: 0.86 x at two GPUs on 100 kbp.)
static int choose_strip16_slice(int64_t n_slice, int64_t n_total, int64_t m, int sms)
{
    const int64_t slots = (int64_t)sms * 4;
    int best = 8;
    double best_t = 1e300;
    for (int c : {2, 3, 4, 6, 8}) {
        if ((n_slice + 64 * c - 1) / (64 * c) > slots && c != 8) continue;
        const double t = (39.0 + 11.0 * c) * ((double)m + (double)((n_total + 64 * c - 1) / (64 * c)) * 118.0);
        if (t < best_t) { best_t = t; best = c; }
    }
    return best;
}

// Lane strip width for one long pair: the pipeline runs at the speed of its slowest band,
// so prefer at most one band (warp) per SM sub-partition (4 per SM) and widen the strips
// when the read is long enough to need more.
static int choose_strip(int64_t n, int sms, bool trace)
{
    const int64_t slots = (int64_t)sms * 4;
    if (n <= 32 * 4 * slots) return 4;
    if (!trace && n <= 32 * 6 * slots) return 6;   // 6-column strips do not pack to whole direction words
    if (n <= 32 * 8 * slots) return 8;
    return 16;
}

// Plans cached by tsta_psa_align, one per (thread, trace mode); freed at thread exit or on request.
struct AlignCache {
    tsta_psa_plan* plan[2] = {nullptr, nullptr};
    void release() { for (auto& q : plan) { plan_free(q); q = nullptr; } }
    ~AlignCache() { release(); }
};
static thread_local AlignCache g_align_cache;

// Tiles around the path of a walk that starts in (i, j), checkpointed traceback: for every range of 128 steps the
// bands between the two straight lines the path is expected to stay near -- the diagonal through (i, j) (related
// reads keep their offset) and the line to the origin (unrelated reads drift back to it) -- widened by W columns on
// both sides.  Ranges are filled from the start of the walk upwards until the slot budget is used up.
static long long build_corridor(const tsta_psa_plan* p, int64_t i, int64_t j, PsaTileRow* rows, int* k_low, int* width)
{
    const int K = (int)psa_tile_rows(p->m);
    // Margin of corridor number `rounds`: the second corridor keeps the margin of the first (it is centred on the
    // point the walk got to, which is what the first one lacked), every later one doubles it (a path that has left
    // two corridors wanders further than expected).  Measured on 200 / 300 kbp pairs that need two corridors: 4.4-7.1
    // ms of traceback against 5.4-8.9 ms when the second corridor already doubles (profiles/r02_corridor_sweep.md).
    // (TSTA_PSA_CORRIDOR_GROWTH=<percent>: experiments, a plain geometric growth from the second corridor on.)
    static const int growth_pct = getenv("TSTA_PSA_CORRIDOR_GROWTH") ? std::max(100, atoi(getenv("TSTA_PSA_CORRIDOR_GROWTH"))) : 0;
    int64_t W = p->corridor_w;
    if (growth_pct) { for (int r = 0; r < std::min(p->rounds, 12) && W < (int64_t)1 << 30; r++) W = W * growth_pct / 100; }
    else if (p->rounds >= 2) W <<= std::min(p->rounds - 1, 6);
    const int64_t band_cols = 64 * p->C;
    long long slots = 0;
    bool full = false;
    *k_low = K;
    *width = 0;
    for (int k = K - 1; k >= 0; k--) {
        rows[k] = PsaTileRow{1, 0, 0};
        // rows whose cells can lie in the range's steps (step = row + virtual lane, 64 virtual lanes)
        const int64_t r_lo = std::max<int64_t>(0, (int64_t)k * kTileSteps - 63), r_hi = std::min<int64_t>(i, (int64_t)k * kTileSteps + kTileSteps - 1);
        if (r_lo > i || full) continue;
        auto diag = [&](int64_t r) { return j - (i - r); };
        auto orig = [&](int64_t r) { return j * (r + 1) / (i + 1); };
        int64_t c_lo = std::min(diag(r_lo), orig(r_lo)) - W, c_hi = std::max(diag(r_hi), orig(r_hi)) + W;
        c_lo = std::min(std::max<int64_t>(c_lo, 0), j);
        c_hi = std::min(std::max<int64_t>(c_hi, 0), j);
        const int bmin = (int)(c_lo / band_cols), bmax = (int)(c_hi / band_cols);
        if (slots + (bmax - bmin + 1) > p->max_slots) { full = true; continue; }
        rows[k] = PsaTileRow{bmin, bmax, slots};
        slots += bmax - bmin + 1;
        *k_low = k;
        *width = std::max(*width, bmax - bmin + 1);
    }
    return slots;
}

template <int C>
static void launch_retrace(const PsaRetraceArgs& ra, int width, int nranges, cudaStream_t st)
{
    psa_retrace16_kernel<C><<<dim3((unsigned)width, (unsigned)nranges), 32, 0, st>>>(ra);
}

// Parallel traceback over the corridor's tiles: summaries, chase, per-visit walks, compaction (psa_chase.cuh).
static cudaEvent_t g_chase_ev[6];      // TSTA_TRACE_DEBUG=1: time stamps around the kernels of the last corridor
static bool g_chase_dbg = false;

template <int C>
static void launch_chase(const PsaChaseArgs& ca, int width, int nranges, cudaStream_t st)
{
    auto stamp = [&](int k) { if (g_chase_dbg) cudaEventRecord(g_chase_ev[k], st); };
    stamp(1);
    psa_tile_summary_kernel<C><<<dim3((unsigned)width, (unsigned)nranges), 32, 0, st>>>(ca);
    stamp(2);
    psa_chase_kernel<C><<<1, 32, 0, st>>>(ca);
    stamp(3);
    psa_visit_walk_kernel<C><<<(unsigned)((ca.max_visits + 63) / 64), 64, 0, st>>>(ca);
    stamp(4);
    psa_visit_compact_kernel<C><<<1, 1024, 0, st>>>(ca);
    stamp(5);
}
static size_t chase_sum_words(int C) { return (size_t)3 * C * 64 + 64 + (size_t)kTileSteps * 3; }
static size_t chase_visit_moves(int C) { return (size_t)kTileSteps + 64 * (size_t)C; }

static int launch_walk(tsta_psa_plan* p, int64_t start_i, int64_t start_j, int start_state, int64_t start_nops);

// Checkpointed traceback: the traced form of the corridor's tiles, then the walk from (i, j).
static int launch_corridor_walk(tsta_psa_plan* p, int64_t i, int64_t j, int state, int64_t nops)
{
    NvtxRange range("tsta_psa corridor");
    const int K = (int)psa_tile_rows(p->m);
    int k_low = K, width = 0;
    p->slots = build_corridor(p, i, j, p->h_rows, &k_low, &width);
    if (p->slots <= 0) return set_error(TSTA_ECUDA, "empty traceback corridor at (%lld, %lld)", (long long)i, (long long)j);
    if ((size_t)p->slots * (size_t)psa_tile_bytes(p->C) > p->d_dirs.cap) {
        TSTA_CUDA_TRY(cudaStreamSynchronize(p->stream));
        if (int rc = p->d_dirs.reserve((size_t)(p->slots + p->slots / 4) * (size_t)psa_tile_bytes(p->C))) return rc;
    }
    TSTA_CUDA_TRY(cudaMemcpyAsync(p->d_tilerows.p, p->h_rows, (size_t)K * sizeof(PsaTileRow), cudaMemcpyHostToDevice, p->stream));
    PsaRetraceArgs ra{};
    ra.sc = p->sc;
    ra.s0 = p->d_s0.as<uint8_t>(); ra.s1 = p->d_s1.as<uint8_t>();
    ra.n = (int)p->n; ra.m = (int)p->m; ra.nbands = p->nbands;
    ra.bnd = p->d_bnd.as<unsigned long long>(); ra.bnd_stride = psa_bnd_stride(p->m);
    ra.epoch = p->epoch;
    ra.ckpt = p->d_ckpt.as<uint32_t>(); ra.ckpt_band_words = psa_ckpt_band_words(p->m, p->C);
    ra.rows = p->d_tilerows.as<PsaTileRow>(); ra.nrows = K;
    ra.dirs = p->d_dirs.as<uint8_t>();
    ra.abort_flag = p->d_ctl.as<int>() + 6;
    ra.timeout_ns = 2000000000ull;
    const int nranges = K - k_low;
    if (getenv("TSTA_TRACE_DEBUG") && !g_chase_dbg) {
        for (auto& e : g_chase_ev) cudaEventCreate(&e);
        g_chase_dbg = true;
    }
    if (g_chase_dbg) cudaEventRecord(g_chase_ev[0], p->stream);
    if (p->C == 2) launch_retrace<2>(ra, width, nranges, p->stream);
    else if (p->C == 3) launch_retrace<3>(ra, width, nranges, p->stream);
    else if (p->C == 4) launch_retrace<4>(ra, width, nranges, p->stream);
    else if (p->C == 6) launch_retrace<6>(ra, width, nranges, p->stream);
    else launch_retrace<8>(ra, width, nranges, p->stream);
    TSTA_CUDA_TRY(cudaGetLastError());
    p->launches++;
    p->rounds++;
    p->tiles_run += p->slots;
    if (!p->chase) {
        if (int rc = launch_walk(p, i, j, state, nops)) return rc;
        p->launches++;
        return TSTA_OK;
    }
    if (p->slots * chase_sum_words(p->C) * sizeof(uint32_t) > p->d_summary.cap) {
        TSTA_CUDA_TRY(cudaStreamSynchronize(p->stream));
        if (int rc = p->d_summary.reserve((size_t)(p->slots + p->slots / 4) * chase_sum_words(p->C) * sizeof(uint32_t))) return rc;
    }
    PsaChaseArgs ca{};
    ca.dirs = p->d_dirs.as<uint8_t>(); ca.tile_bytes = psa_tile_bytes(p->C);
    ca.summary = p->d_summary.as<uint32_t>();
    ca.rows = p->d_tilerows.as<PsaTileRow>(); ca.nrows = K;
    ca.n = (int)p->n; ca.m = (int)p->m;
    ca.start_i = (int)i; ca.start_j = (int)j; ca.start_arr = state; ca.start_nops = (int)nops;
    ca.visits = p->d_visits.as<int4>(); ca.max_visits = p->max_visits;
    ca.nvisits = p->d_counts.as<int>();                       // [0] number of visits, [4..6] exits of the start cell, then counts, then offsets
    ca.start_exit = p->d_counts.as<uint32_t>() + 4;
    ca.counts = p->d_counts.as<int>() + 8;
    ca.offsets = p->d_counts.as<int>() + 8 + p->max_visits;
    ca.slots = p->d_vslots.as<uint8_t>();
    ca.ops = p->d_ops.as<uint8_t>();
    ca.out = p->d_ctl.as<int>() + 2;
    {
        NvtxRange r2("tsta_psa traceback");
        if (p->C == 2) launch_chase<2>(ca, width, nranges, p->stream);
        else if (p->C == 3) launch_chase<3>(ca, width, nranges, p->stream);
        else if (p->C == 4) launch_chase<4>(ca, width, nranges, p->stream);
        else if (p->C == 6) launch_chase<6>(ca, width, nranges, p->stream);
        else launch_chase<8>(ca, width, nranges, p->stream);
    }
    TSTA_CUDA_TRY(cudaGetLastError());
    p->launches += 4;
    return TSTA_OK;
}

// Traceback kernel of a plan (whole pair or one column slice).
static int launch_walk(tsta_psa_plan* p, int64_t start_i, int64_t start_j, int start_state, int64_t start_nops)
{
    NvtxRange range("tsta_psa traceback");
    int* ctl = p->d_ctl.as<int>();
    PsaTraceArgs ta{};
    ta.dirs = p->d_dirs.as<uint8_t>();
    ta.dirs_stride = p->n_pad / 2;
    ta.n = (int)p->n; ta.m = (int)p->m;
    ta.ops = p->d_ops.as<uint8_t>();
    ta.out = ctl + 2;
    ta.enc = p->packed ? 1 : 0;
    ta.start_i = (int)start_i; ta.start_j = (int)start_j; ta.start_forced = start_state;
    ta.start_nops = (int)start_nops;
    if (getenv("TSTA_TRACE_DEBUG")) { if (!tsta::g_trace_dbg) cudaMalloc(&tsta::g_trace_dbg, 64); ta.dbg = tsta::g_trace_dbg; }
    if (p->skew) {
        ta.skew_band_bytes = psa_skew_band_bytes(p->m, p->C);
        ta.skew_steps = psa_skew_steps(p->m);
        ta.skew_bands = p->nbands;
        if (p->ckpt) {
            ta.tile_rows = p->d_tilerows.as<PsaTileRow>();
            ta.tile_bytes = psa_tile_bytes(p->C);
        }
        cudaError_t e;
        if (p->C == 2) e = launch_trace_skew<2>(ta, p->stream);
        else if (p->C == 3) e = launch_trace_skew<3>(ta, p->stream);
        else if (p->C == 4) e = launch_trace_skew<4>(ta, p->stream);
        else if (p->C == 6) e = launch_trace_skew<6>(ta, p->stream);
        else e = launch_trace_skew<8>(ta, p->stream);
        TSTA_CUDA_TRY(e);
    } else {
        // (function attributes are per device: set it on every launch, it costs nothing)
        TSTA_CUDA_TRY(cudaFuncSetAttribute(psa_trace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTraceSmemBytes));
        psa_trace_kernel<<<1, kTraceThreads, kTraceSmemBytes, p->stream>>>(ta);
        TSTA_CUDA_TRY(cudaGetLastError());
    }
    return TSTA_OK;
}

// Moves of the last walk as one byte each (1 diagonal, 0 left, 2 up), walk order.  The skewed-layout
// walker writes the tags (3 diagonal, 2 left, 1 up).
static int fetch_moves(tsta_psa_plan* p, uint8_t* out, int64_t nops)
{
    if (nops <= 0) return TSTA_OK;
    TSTA_CUDA_TRY(cudaMemcpy(out, p->d_ops.p, (size_t)nops, cudaMemcpyDeviceToHost));
    if (p->skew) {
        static const uint8_t kMove[4] = {1, 2, 0, 1};   // tag -> move (tag 0 never occurs)
        for (int64_t k = 0; k < nops; k++) out[k] = kMove[out[k] & 3u];
    }
    return TSTA_OK;
}

namespace tsta { void batch_release_ctx(); }   // tsta_psa_batch.cu: the calling thread's batch context

extern "C" {

const char* tsta_last_error(void) { return g_last_error.c_str(); }

void* tsta_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (ensure_device() != TSTA_OK) return nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        set_error(TSTA_ENOMEM, "cudaMallocHost(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}
void tsta_host_free(void* p) { if (p) cudaFreeHost(p); }
const char* tsta_version(void) { return "tsta_b200 0.2 (sm_100a)"; }

int tsta_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error(TSTA_ENODEV, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    }
    return n;
}

int tsta_set_device(int device)
{
    if (int rc = ensure_device()) return rc;
    TSTA_CUDA_TRY(cudaSetDevice(device));
    return TSTA_OK;
}

int tsta_psa_plan_create(tsta_psa_plan** out, int want_trace)
{
    if (!out) return set_error(TSTA_EINVAL, "plan pointer is NULL");
    *out = nullptr;
    if (int rc = ensure_device()) return rc;
    tsta_psa_plan* p = new (std::nothrow) tsta_psa_plan();
    if (!p) return set_error(TSTA_ENOMEM, "out of host memory");
    p->want_trace = want_trace ? 1 : 0;
    cudaError_t e = cudaGetDevice(&p->device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&p->sms, cudaDevAttrMultiProcessorCount, p->device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
    for (int k = 0; k < 4 && e == cudaSuccess; k++) e = cudaEventCreate(&p->ev[k]);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&p->h_small, 16 * sizeof(int));
    if (e != cudaSuccess) {
        plan_free(p);
        return set_error(TSTA_ECUDA, "plan creation failed: %s", cudaGetErrorString(e));
    }
    *out = p;
    return TSTA_OK;
}

void tsta_psa_plan_destroy(tsta_psa_plan* p) { plan_free(p); }

static int plan_upload(tsta_psa_plan* p, const tsta_params* prm, const char* s0, int64_t n, const char* s1, int64_t m,
                       int64_t col0, int64_t n_total, bool has_left, bool has_right);

int tsta_psa_plan_upload(tsta_psa_plan* p, const tsta_params* prm,
                         const char* s0, int64_t n, const char* s1, int64_t m)
{
    return plan_upload(p, prm, s0, n, s1, m, 0, n, false, false);
}

int tsta_psa_plan_upload_slice(tsta_psa_plan* p, const tsta_params* prm, const char* s0_slice, int64_t n_slice,
                               int64_t col0, int64_t n_total, const char* s1, int64_t m, int has_left, int has_right)
{
    if (!p) return set_error(TSTA_EINVAL, "plan is NULL");
    if (n_slice <= 0 || m <= 0 || col0 < 0 || col0 + n_slice > n_total || (has_left != 0) != (col0 > 0) ||
        (has_right != 0) != (col0 + n_slice < n_total))
        return set_error(TSTA_EINVAL, "bad slice: columns [%lld, %lld) of %lld", (long long)col0, (long long)(col0 + n_slice), (long long)n_total);
    if (n_total < m) return set_error(TSTA_EINVAL, "s0 must be the longer read");
    return plan_upload(p, prm, s0_slice, n_slice, s1, m, col0, n_total, has_left != 0, has_right != 0);
}

int tsta_psa_plan_inbox_handle(tsta_psa_plan* p, unsigned char handle[64])
{
    if (!p || !handle || !p->d_inbox.p) return set_error(TSTA_EINVAL, "plan has no inbox (upload a slice with has_left first)");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    TSTA_CUDA_TRY(cudaSetDevice(p->device));
    cudaIpcMemHandle_t h;
    TSTA_CUDA_TRY(cudaIpcGetMemHandle(&h, p->d_inbox.p));
    memcpy(handle, &h, 64);
    return TSTA_OK;
}

int tsta_psa_plan_connect_right(tsta_psa_plan* p, const unsigned char handle[64])
{
    if (!p || !handle) return set_error(TSTA_EINVAL, "NULL argument");
    TSTA_CUDA_TRY(cudaSetDevice(p->device));
    if (p->outbox_open) { cudaIpcCloseMemHandle(p->outbox); p->outbox_open = false; p->outbox = nullptr; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* ptr = nullptr;
    TSTA_CUDA_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    p->outbox = static_cast<unsigned long long*>(ptr);
    p->outbox_open = true;
    return TSTA_OK;
}

int tsta_psa_plan_connect_right_local(tsta_psa_plan* p, tsta_psa_plan* right)
{
    // same process, both plans on devices with peer access (or the same device)
    if (!p || !right || !right->d_inbox.p) return set_error(TSTA_EINVAL, "right plan has no inbox");
    if (p->device != right->device) {
        TSTA_CUDA_TRY(cudaSetDevice(p->device));
        cudaError_t e = cudaDeviceEnablePeerAccess(right->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) TSTA_CUDA_TRY(e);
        cudaGetLastError();
    }
    p->outbox = right->d_inbox.as<unsigned long long>();
    p->outbox_open = false;
    return TSTA_OK;
}

static int plan_upload(tsta_psa_plan* p, const tsta_params* prm, const char* s0, int64_t n, const char* s1, int64_t m,
                       int64_t col0, int64_t n_total, bool has_left, bool has_right)
{
    if (!p) return set_error(TSTA_EINVAL, "plan is NULL");
    if (int rc = check_params(prm)) return rc;
    if (n < 0 || m < 0 || (n > 0 && !s0) || (m > 0 && !s1)) return set_error(TSTA_EINVAL, "bad sequence arguments");
    if (n_total < m && !getenv("TSTA_PSA_ALLOW_TALL"))   // (experiments only: tools/calib.py times few bands over many rows)
        return set_error(TSTA_EINVAL, "s0 must be the longer read (n=%lld < m=%lld); swap like psa.c:194-199",
                                      (long long)n_total, (long long)m);
    // int32 headroom: doubled scores, worst step 2*127 per row/column
    if ((n_total + m) > (int64_t)4000000) return set_error(TSTA_EINVAL, "n+m=%lld exceeds the supported 4e6", (long long)(n_total + m));
    TSTA_CUDA_TRY(cudaSetDevice(p->device));
    p->col0 = col0; p->n_total = n_total; p->has_left = has_left; p->has_right = has_right;
    p->sc.M = prm->M; p->sc.X = prm->X; p->sc.E = prm->E; p->sc.O = prm->O;
    p->n = n;
    p->m = m;
    p->uploaded = false;
    p->ran = false;
    p->h_s0.assign(s0, s0 + n);
    p->h_s1.assign(s1, s1 + m);
    if (n == 0 || m == 0) {  // degenerate: nothing to fill
        p->uploaded = true;
        return TSTA_OK;
    }
    const int S = p->want_trace ? 2 : 1;
    p->C = choose_strip(n, p->sms, p->want_trace != 0);
    p->CH = 0;   // 0 = the kernel family's default (16 rows for int32, 32 for packed)
    // tuning overrides for experiments (not part of the ABI)
    if (const char* e = getenv("TSTA_PSA_C")) { const int c = atoi(e); if (c == 4 || c == 8 || c == 16 || (c == 6 && !p->want_trace)) p->C = c; }
    if (const char* e = getenv("TSTA_PSA_CH")) { const int c = atoi(e); if (c == 8 || c == 16 || c == 32) p->CH = c; }
    if (int rc = p->d_s0.reserve((size_t)n)) return rc;
    if (int rc = p->d_s1.reserve((size_t)m)) return rc;
    if (int rc = p->d_ctl.reserve(16 * sizeof(int))) return rc;
    int* ctl = p->d_ctl.as<int>();
    TSTA_CUDA_TRY(cudaMemcpyAsync(p->d_s0.p, s0, (size_t)n, cudaMemcpyHostToDevice, p->stream));
    TSTA_CUDA_TRY(cudaMemcpyAsync(p->d_s1.p, s1, (size_t)m, cudaMemcpyHostToDevice, p->stream));
    // alphabet: distinct bytes of both reads -> codes 0..K-1 (comparison is raw-byte equality,
    // psa.c:253 mm_cmpeq_epi8, so any relabelling that keeps bytes distinct is exact)
    TSTA_CUDA_TRY(cudaMemsetAsync(ctl + 8, 0, 8 * sizeof(int), p->stream));
    TSTA_CUDA_TRY(launch_byte_presence(p->d_s0.as<uint8_t>(), n, reinterpret_cast<unsigned*>(ctl + 8), p->sms, p->stream));
    TSTA_CUDA_TRY(launch_byte_presence(p->d_s1.as<uint8_t>(), m, reinterpret_cast<unsigned*>(ctl + 8), p->sms, p->stream));
    TSTA_CUDA_TRY(cudaMemcpyAsync(p->h_small + 8, ctl + 8, 8 * sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    TSTA_CUDA_TRY(cudaStreamSynchronize(p->stream));
    const int K = build_lut(reinterpret_cast<unsigned*>(p->h_small + 8), p->sc.lut);
    p->nsym = K;
    p->prmt = (K <= kPadCode) && abs(S * prm->M) <= 127 && abs(S * prm->X) <= 127;
    // packed int16 kernels when the alphabet and the score range allow (psa16_fits)
    p->packed = false;
    {
        const bool tr = p->want_trace != 0;
        int c16 = (has_left || has_right) ? choose_strip16_slice(n, n_total, m, p->sms) : choose_strip16(n, p->sms);
        if (const char* e = getenv("TSTA_PSA_C16")) {
            const int c = atoi(e);
            if (c == 2 || c == 3 || c == 4 || c == 6 || c == 8) c16 = c;
        }
        const char* off = getenv("TSTA_PSA_PACKED");
        if (!(off && atoi(off) == 0) && psa16_fits(prm->M, prm->X, prm->E, prm->O, tr ? 4 : 1, c16, K)) {
            p->packed = true;
            p->C = c16;
        }
    }
    p->skew = p->packed && p->want_trace;
    {
        // TSTA_PSA_TRACE_MODE=full keeps the direction words of the WHOLE matrix (what column slices do)
        const char* tm = getenv("TSTA_PSA_TRACE_MODE");
        p->ckpt = p->skew && !(has_left || has_right) && !(tm && strcmp(tm, "full") == 0);
    }
    p->scale = p->ckpt ? 1 : (p->skew ? 4 : S);
    if (p->ckpt) p->CH = 32;
    const int band_cols = (p->packed ? 64 : 32) * p->C;
    p->nbands = (int)((n + band_cols - 1) / band_cols);
    p->n_pad = p->nbands * band_cols;
    if (has_left || has_right) {
        if (!p->packed) return set_error(TSTA_EINVAL, "column slices need the packed kernels (alphabet of <= 4 symbols, small scores)");
        if (has_right && n % band_cols != 0)
            return set_error(TSTA_EINVAL, "a slice with a right neighbour must be a multiple of %d columns (got %lld)", band_cols, (long long)n);
        if (has_left && (size_t)m > p->inbox_rows) {
            if (int rc = p->d_inbox.reserve(((size_t)m + kBndPadRows) * sizeof(unsigned long long))) return rc;
            TSTA_CUDA_TRY(cudaMemsetAsync(p->d_inbox.p, 0, p->d_inbox.cap, p->stream));
            p->inbox_rows = (size_t)m;
        }
    }

    {   // boundary columns: tagged elements, so fresh memory must start out all-zero (tag 0)
        // (+ kBndPadRows: the streamed import of psa_sweep_band16 requests a few rows past the end of a column)
        const size_t need = ((size_t)p->nbands * (size_t)psa_bnd_stride(m) + kBndPadRows) * sizeof(unsigned long long);
        if (need > p->d_bnd.cap) {
            if (int rc = p->d_bnd.reserve(need + need / 4)) return rc;
            TSTA_CUDA_TRY(cudaMemsetAsync(p->d_bnd.p, 0, p->d_bnd.cap, p->stream));
            // (the epoch stays monotonic: fresh memory carries tag 0, which no run uses, and the slices of
            // one pair on other GPUs keep counting in step with this plan)
        }
    }
    if (int rc = p->d_lastrow.reserve((size_t)p->n_pad * sizeof(int))) return rc;
    if (p->ckpt) {
        const size_t K = (size_t)psa_tile_rows(m);
        if (int rc = p->d_ckpt.reserve((size_t)p->nbands * (size_t)psa_ckpt_band_words(m, p->C) * sizeof(uint32_t))) return rc;
        if (int rc = p->d_tilerows.reserve(K * sizeof(PsaTileRow))) return rc;
        if (K > p->rows_cap) {
            if (p->h_rows) cudaFreeHost(p->h_rows);
            p->h_rows = nullptr;
            p->rows_cap = 0;
            TSTA_CUDA_TRY(cudaMallocHost((void**)&p->h_rows, (K + K / 4) * sizeof(PsaTileRow)));
            p->rows_cap = K + K / 4;
        }
        // margin: the optimal path of two unrelated reads wanders ~ length^(2/3) away from the straight line
        // (measured: +-1000 columns at 100 kbp); 2048 columns at 100 kbp, scaled by that law for longer reads
        p->corridor_w = 2048;
        if (n > 100000) p->corridor_w = (int64_t)(2048.0 * pow((double)n / 100000.0, 2.0 / 3.0));
        if (const char* e = getenv("TSTA_PSA_CORRIDOR")) p->corridor_w = atoll(e);
        p->corridor_w = std::max<int64_t>(p->corridor_w, 1024);      // (a window of the walk is 512 columns wide: progress per corridor)
        long long budget_mb = 8192;
        if (const char* e = getenv("TSTA_PSA_TILE_BUDGET_MB")) budget_mb = atoll(e);
        budget_mb = std::max<long long>(budget_mb, 16);
        p->max_slots = std::max<long long>((budget_mb << 20) / psa_tile_bytes(p->C), 2 * (p->nbands + 1));
        int k_low = 0, width = 0;
        p->rounds = 0;
        const long long first = build_corridor(p, m - 1, n - 1, p->h_rows, &k_low, &width);
        if (int rc = p->d_dirs.reserve((size_t)first * (size_t)psa_tile_bytes(p->C))) return rc;
        {
            const char* w = getenv("TSTA_PSA_WALK");           // "serial": the one-thread window walker instead of the tile chase
            p->chase = !(w && strcmp(w, "serial") == 0);
        }
        if (p->chase) {
            p->max_visits = 4 * ((int)K + p->nbands) + 64;
            if (int rc = p->d_summary.reserve((size_t)first * chase_sum_words(p->C) * sizeof(uint32_t))) return rc;
            if (int rc = p->d_visits.reserve((size_t)p->max_visits * sizeof(int4))) return rc;
            if (int rc = p->d_counts.reserve((size_t)(2 * p->max_visits + 8) * sizeof(int))) return rc;
            if (int rc = p->d_vslots.reserve((size_t)p->max_visits * chase_visit_moves(p->C))) return rc;
        }
        if (int rc = p->d_ops.reserve((size_t)(n + m) + 64)) return rc;
    } else if (p->want_trace) {
        const size_t dir_bytes = p->skew ? (size_t)p->nbands * (size_t)psa_skew_band_bytes(m, p->C)
                                         : (size_t)m * (size_t)(p->n_pad / 2);
        if (int rc = p->d_dirs.reserve(dir_bytes)) return rc;
        if (int rc = p->d_ops.reserve((size_t)(n + m) + 64)) return rc;
    }
    p->uploaded = true;
    return TSTA_OK;
}

int tsta_psa_plan_run(tsta_psa_plan* p, float* fill_ms, float* trace_ms)
{
    if (int rc = tsta_psa_plan_launch(p)) return rc;
    return tsta_psa_plan_wait(p, fill_ms, trace_ms);
}

int tsta_psa_plan_wait(tsta_psa_plan* p, float* fill_ms, float* trace_ms)
{
    if (!p || !p->uploaded) return set_error(TSTA_EINVAL, "plan has no uploaded pair");
    if (fill_ms) *fill_ms = 0.f;
    if (trace_ms) *trace_ms = 0.f;
    if (p->n == 0 || p->m == 0) { p->ran = true; return TSTA_OK; }
    TSTA_CUDA_TRY(cudaSetDevice(p->device));
    for (;;) {
        TSTA_CUDA_TRY(cudaMemcpyAsync(p->h_small, p->d_ctl.as<int>(), 8 * sizeof(int), cudaMemcpyDeviceToHost, p->stream));
        TSTA_CUDA_TRY(cudaStreamSynchronize(p->stream));
        if (p->h_small[6] != 0)
            return set_error(TSTA_ECUDA, "boundary hand-off timed out: a band (or column slice) waited for its left neighbour in vain "
                                         "(slices of one pair must be launched together and run the same number of times)");
        if (!p->want_trace || p->h_small[7] == 0) break;
        if (!p->ckpt || p->h_small[7] == 3)
            return set_error(TSTA_ECUDA, "traceback ran into cells without a direction code at (%d, %d)", p->h_small[3], p->h_small[4]);
        // the walk left the corridor (or the corridor's tile budget ended): a new one around the point it got to
        if (p->rounds > 100000) return set_error(TSTA_ECUDA, "traceback does not advance at (%d, %d)", p->h_small[3], p->h_small[4]);
        TSTA_CUDA_TRY(cudaMemsetAsync(p->d_ctl.as<int>() + 7, 0, sizeof(int), p->stream));
        if (int rc = launch_corridor_walk(p, p->h_small[3], p->h_small[4], p->h_small[5], p->h_small[2])) return rc;
        TSTA_CUDA_TRY(cudaEventRecord(p->ev[2], p->stream));
    }
    if (p->ckpt && p->chase && g_chase_dbg) {
        float t[5] = {0, 0, 0, 0, 0};
        for (int k = 0; k < 5; k++) cudaEventElapsedTime(&t[k], g_chase_ev[k], g_chase_ev[k + 1]);
        int nv = 0;
        cudaMemcpy(&nv, p->d_counts.p, sizeof(int), cudaMemcpyDeviceToHost);
        fprintf(stderr, "trace dbg: last corridor of %d: %lld tiles; retrace %.3f summary %.3f chase %.3f (%d visits) walks %.3f compact %.3f ms\n",
                p->rounds, p->slots, t[0], t[1], t[2], nv, t[3], t[4]);
    } else if (p->want_trace && getenv("TSTA_TRACE_DEBUG")) {
        long long h[4] = {0, 0, 0, 0};
        using tsta::g_trace_dbg;
        if (g_trace_dbg) { cudaMemcpy(h, g_trace_dbg, 32, cudaMemcpyDeviceToHost);
            fprintf(stderr, "trace dbg: walk %lld cyc, total %lld cyc, windows %lld, hits %lld\n", h[0], h[1], h[2], h[3]); }
    }
    float ms = 0.f;
    if (fill_ms) { TSTA_CUDA_TRY(cudaEventElapsedTime(&ms, p->ev[0], p->ev[1])); *fill_ms = ms; }
    if (trace_ms) { TSTA_CUDA_TRY(cudaEventElapsedTime(&ms, p->ev[1], p->ev[2])); *trace_ms = ms; }
    p->ran = true;
    return TSTA_OK;
}

int tsta_psa_plan_launch(tsta_psa_plan* p)
{
    if (!p || !p->uploaded) return set_error(TSTA_EINVAL, "plan has no uploaded pair");
    p->launches = 0;
    if (p->n == 0 || p->m == 0) return TSTA_OK;
    TSTA_CUDA_TRY(cudaSetDevice(p->device));
    int* ctl = p->d_ctl.as<int>();
    // ticket = 0; out_max = INT_MIN; a fresh boundary tag for this run
    TSTA_CUDA_TRY(cudaMemsetAsync(ctl, 0, 8 * sizeof(int), p->stream));
    if (++p->epoch > 0xffffu) {
        TSTA_CUDA_TRY(cudaMemsetAsync(p->d_bnd.p, 0, p->d_bnd.cap, p->stream));
        if (p->d_inbox.p) TSTA_CUDA_TRY(cudaMemsetAsync(p->d_inbox.p, 0, p->d_inbox.cap, p->stream));
        p->epoch = 1;
    }
    p->h_small[0] = INT_MIN;
    TSTA_CUDA_TRY(cudaMemcpyAsync(ctl + 1, p->h_small, sizeof(int), cudaMemcpyHostToDevice, p->stream));

    PsaPairArgs fa{};
    fa.sc = p->sc;
    fa.s0 = p->d_s0.as<uint8_t>();
    fa.s1 = p->d_s1.as<uint8_t>();
    fa.n = (int)p->n; fa.m = (int)p->m; fa.nbands = p->nbands;
    fa.bnd = p->d_bnd.as<unsigned long long>();
    fa.bnd_stride = psa_bnd_stride(p->m);
    fa.epoch = p->epoch;
    fa.ticket = reinterpret_cast<unsigned*>(ctl);
    fa.dirs = p->d_dirs.as<uint8_t>();
    fa.dirs_stride = p->skew ? psa_skew_band_bytes(p->m, p->C) : p->n_pad / 2;
    fa.abort_flag = ctl + 6;
    fa.timeout_ns = 10000000000ull;   // a band that waits this long for its left neighbour gives up (TSTA_ECUDA instead of a hang)
    if (const char* e = getenv("TSTA_HANDOFF_TIMEOUT_MS")) fa.timeout_ns = (unsigned long long)atoll(e) * 1000000ull;
    fa.timeline = nullptr;
    if (getenv("TSTA_PSA_TIMELINE")) {   // experiments: per-band time stamps, read back with tsta_psa_plan_timeline
        if (int rc = p->d_timeline.reserve((size_t)p->nbands * 16 * sizeof(unsigned long long))) return rc;
        TSTA_CUDA_TRY(cudaMemsetAsync(p->d_timeline.p, 0, (size_t)p->nbands * 128, p->stream));
        fa.timeline = p->d_timeline.as<unsigned long long>();
    }
    fa.ckpt = p->ckpt ? p->d_ckpt.as<uint32_t>() : nullptr;
    fa.ckpt_band_words = p->ckpt ? psa_ckpt_band_words(p->m, p->C) : 0;
    p->rounds = 0;
    p->tiles_run = 0;
    fa.out_max = ctl + 1;
    fa.lastrow = p->d_lastrow.as<int>();
    fa.col_off = (int)p->col0;
    fa.inbox = p->has_left ? p->d_inbox.as<unsigned long long>() : nullptr;
    fa.outbox = p->has_right ? p->outbox : nullptr;
    if (p->has_right && !p->outbox) return set_error(TSTA_EINVAL, "slice has a right neighbour but tsta_psa_plan_connect_right was not called");

    TSTA_CUDA_TRY(cudaEventRecord(p->ev[0], p->stream));
    nvtxRangePushA("tsta_psa fill");
    if (p->want_trace && p->packed && !p->ckpt) {
        if (p->C == 2) launch_pair16<2, 2>(p, fa);
        else if (p->C == 3) launch_pair16<3, 2>(p, fa);
        else if (p->C == 4) launch_pair16<4, 2>(p, fa);
        else if (p->C == 6) launch_pair16<6, 2>(p, fa);
        else launch_pair16<8, 2>(p, fa);
    } else if (p->want_trace && !p->ckpt) {
        if (p->C == 4) launch_pair<4, true>(p, fa);
        else if (p->C == 8) launch_pair<8, true>(p, fa);
        else launch_pair<16, true>(p, fa);
    } else if (p->packed) {
        if (p->C == 2) launch_pair16<2, 0>(p, fa);
        else if (p->C == 3) launch_pair16<3, 0>(p, fa);
        else if (p->C == 4) launch_pair16<4, 0>(p, fa);
        else if (p->C == 6) launch_pair16<6, 0>(p, fa);
        else launch_pair16<8, 0>(p, fa);
    } else {
        if (p->C == 4) launch_pair<4, false>(p, fa);
        else if (p->C == 6) launch_pair<6, false>(p, fa);
        else if (p->C == 8) launch_pair<8, false>(p, fa);
        else launch_pair<16, false>(p, fa);
    }
    nvtxRangePop();
    TSTA_CUDA_TRY(cudaGetLastError());
    p->launches++;
    TSTA_CUDA_TRY(cudaEventRecord(p->ev[1], p->stream));
    if (p->ckpt) {
        if (int rc = launch_corridor_walk(p, fa.m - 1, fa.n - 1, 0, 0)) return rc;
    } else if (p->want_trace && !(p->has_left || p->has_right)) {   // (slices are walked later, right to left: tsta_psa_plan_trace_from)
        if (int rc = launch_walk(p, fa.m - 1, fa.n - 1, 0, 0)) return rc;
        p->launches++;
    }
    TSTA_CUDA_TRY(cudaEventRecord(p->ev[2], p->stream));
    return TSTA_OK;
}

int tsta_psa_plan_fetch(tsta_psa_plan* p, tsta_psa_result* out)
{
    if (!p || !out) return set_error(TSTA_EINVAL, "NULL argument");
    if (!p->ran) return set_error(TSTA_EINVAL, "plan has not been run");
    memset(out, 0, sizeof *out);
    const int64_t n = p->n, m = p->m;
    const int S = p->scale;
    std::vector<uint8_t> ops;
    int64_t nops = 0, fi = m - 1, fj = n - 1;
    out->maxsorce = -100;  // reference seed, psa.c:110,590
    if (n > 0 && m > 0) {
        TSTA_CUDA_TRY(cudaSetDevice(p->device));
        int* ctl = p->d_ctl.as<int>();
        TSTA_CUDA_TRY(cudaMemcpyAsync(p->h_small, ctl, 8 * sizeof(int), cudaMemcpyDeviceToHost, p->stream));
        TSTA_CUDA_TRY(cudaMemcpyAsync(p->h_small + 7, p->d_lastrow.as<int>() + (n - 1), sizeof(int),
                                      cudaMemcpyDeviceToHost, p->stream));
        TSTA_CUDA_TRY(cudaStreamSynchronize(p->stream));
        out->maxsorce = std::max(-100, p->h_small[1] / S);
        out->corner = p->h_small[7] / S;
        if (p->want_trace && !(p->has_left || p->has_right)) {
            nops = p->h_small[2];
            fi = p->h_small[3];
            fj = p->h_small[4];
            ops.resize((size_t)(n + m));
            if (int rc = fetch_moves(p, ops.data(), nops)) return rc;
        }
    }
    if (p->want_trace && !(p->has_left || p->has_right)) {   // (slices: tsta_psa_plan_trace_from / _ops)
        ops.resize((size_t)(n + m));
        // leftovers: rest of s0 against gaps, then rest of s1 (psa.c:497-510)
        while (fj >= 0) { ops[(size_t)nops++] = 0; fj--; }
        while (fi >= 0) { ops[(size_t)nops++] = 2; fi--; }
        char* a = (char*)malloc((size_t)nops + 1);
        char* b = (char*)malloc((size_t)nops + 1);
        if (!a || !b) { free(a); free(b); return set_error(TSTA_ENOMEM, "out of host memory"); }
        int64_t i = m - 1, j = n - 1, pos = nops;
        for (int64_t k = 0; k < nops; k++) {
            pos--;
            const uint8_t d = ops[(size_t)k];
            if (d == 1) { a[pos] = p->h_s0[(size_t)j--]; b[pos] = p->h_s1[(size_t)i--]; }
            else if (d == 0) { a[pos] = p->h_s0[(size_t)j--]; b[pos] = '-'; }
            else { a[pos] = '-'; b[pos] = p->h_s1[(size_t)i--]; }
        }
        a[nops] = b[nops] = '\0';
        out->a = a; out->b = b; out->alen = nops;
    }
    return TSTA_OK;
}

int tsta_psa_plan_lastrow(tsta_psa_plan* p, int32_t* out)
{
    if (!p || !out || !p->ran) return set_error(TSTA_EINVAL, "plan not run or NULL argument");
    if (p->n == 0 || p->m == 0) return TSTA_OK;
    TSTA_CUDA_TRY(cudaSetDevice(p->device));
    TSTA_CUDA_TRY(cudaMemcpy(out, p->d_lastrow.p, (size_t)p->n * sizeof(int), cudaMemcpyDeviceToHost));
    if (p->scale > 1) for (int64_t j = 0; j < p->n; j++) out[j] /= p->scale;
    return TSTA_OK;
}

int tsta_psa_plan_dirs(tsta_psa_plan* p, uint8_t* out)
{
    if (!p || !out || !p->ran || !p->want_trace) return set_error(TSTA_EINVAL, "plan not run with traceback");
    if (p->n == 0 || p->m == 0) return TSTA_OK;
    TSTA_CUDA_TRY(cudaSetDevice(p->device));
    const int64_t n = p->n, m = p->m;
    if (p->skew) {
        // skewed layout: band b, step t = row + virtual lane, lane; the cell holds back tag | wf << 2 | we << 3.
        // Reported in the oracle's form: cF(i,j) = wf(i,j-1), cE(i,j) = we(i-1,j) (see psa_fill16.cuh).
        const int C = p->C, W = C > 4 ? 2 : 1;
        const size_t band_bytes = (size_t)psa_skew_band_bytes(m, C);
        const int64_t T = psa_skew_steps(m);
        if (p->ckpt) {
            // checkpointed traceback keeps tiles around the path only: for this export EVERY tile is re-run
            // (a corridor as wide as the matrix; the walk that follows finds the same path again)
            const long long saved_budget = p->max_slots, saved_w = p->corridor_w;
            p->max_slots = psa_tile_rows(m) * p->nbands;
            p->corridor_w = 8000000;
            TSTA_CUDA_TRY(cudaMemsetAsync(p->d_ctl.as<int>() + 2, 0, 6 * sizeof(int), p->stream));
            const int rc = launch_corridor_walk(p, m - 1, n - 1, 0, 0);
            p->max_slots = saved_budget;
            p->corridor_w = saved_w;
            if (rc) return rc;
            TSTA_CUDA_TRY(cudaStreamSynchronize(p->stream));
        }
        std::vector<uint32_t> raw(p->ckpt ? (size_t)p->slots * (size_t)psa_tile_bytes(C) / 4 : (size_t)p->nbands * band_bytes / 4);
        TSTA_CUDA_TRY(cudaMemcpy(raw.data(), p->d_dirs.p, raw.size() * 4, cudaMemcpyDeviceToHost));
        std::vector<PsaTileRow> rows;
        if (p->ckpt) rows.assign(p->h_rows, p->h_rows + psa_tile_rows(m));
        auto cell = [&](int64_t i, int64_t j) -> unsigned {
            const int64_t b = j / (64 * C), v = (j % (64 * C)) / C, c = j % C;
            size_t w;
            if (p->ckpt) {
                const int64_t t = i + v;
                const PsaTileRow& r = rows[(size_t)(t / kTileSteps)];
                if (b < r.bmin || b > r.bmax) return 0u;
                w = ((size_t)(r.slot0 + (b - r.bmin)) * kTileSteps + (size_t)(t % kTileSteps)) * 32 * W + (size_t)(v & 31) * W + (W == 2 ? (size_t)(v >> 5) : 0);
            } else
            w = ((size_t)b * (size_t)T + (size_t)(i + v)) * 32 * W + (size_t)(v & 31) * W + (W == 2 ? (size_t)(v >> 5) : 0);
            const unsigned sh = (unsigned)(4 * c + (W == 1 ? 16 * (v >> 5) : 0));
            return (raw[w] >> sh) & 15u;
        };
        for (int64_t i = 0; i < m; i++)
            for (int64_t j = 0; j < n; j++) {
                const unsigned c = cell(i, j);
                unsigned code = (0x48u >> (2 * (c & 3))) & 3u;                   // tags 3 / 2 / 1 -> 1 diagonal, 0 left, 2 up
                if (j >= 1 && (cell(i, j - 1) & 4u)) code |= 4u;
                if (i >= 1 && (cell(i - 1, j) & 8u)) code |= 8u;
                out[(size_t)i * n + j] = (uint8_t)code;
            }
        return TSTA_OK;
    }
    const size_t stride = (size_t)p->n_pad / 2;
    std::vector<uint8_t> packed((size_t)p->m * stride);
    TSTA_CUDA_TRY(cudaMemcpy(packed.data(), p->d_dirs.p, packed.size(), cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < p->m; i++)
        for (int64_t j = 0; j < p->n; j++) {
            const uint8_t b = packed[(size_t)i * stride + (size_t)(j >> 1)];
            uint8_t c = (j & 1) ? (b >> 4) : (b & 15);
            if (p->packed) c = (uint8_t)((c & 12) | ((0x48 >> (2 * (c & 3))) & 3));   // packed kernels tag diagonal/left/up as 3/2/1
            out[(size_t)i * p->n + j] = c;
        }
    return TSTA_OK;
}

int tsta_psa_plan_info(tsta_psa_plan* p, int* launches, int* cols_per_lane, int* bands)
{
    if (!p) return set_error(TSTA_EINVAL, "plan is NULL");
    if (launches) *launches = p->launches;
    if (cols_per_lane) *cols_per_lane = p->C;
    if (bands) *bands = p->nbands;
    return TSTA_OK;
}

int tsta_psa_plan_packed(tsta_psa_plan* p) { return p && p->packed ? 1 : 0; }

int tsta_psa_plan_trace_stats(tsta_psa_plan* p, int* corridors, long long* tiles, long long* tiles_of_matrix)
{
    if (!p) return set_error(TSTA_EINVAL, "plan is NULL");
    if (corridors) *corridors = p->ckpt ? p->rounds : 0;
    if (tiles) *tiles = p->ckpt ? p->tiles_run : 0;
    if (tiles_of_matrix) *tiles_of_matrix = p->ckpt ? psa_tile_rows(p->m) * p->nbands : 0;
    return TSTA_OK;
}

int tsta_psa_plan_timeline(tsta_psa_plan* p, unsigned long long* out)   // [bands][16]: 8 globaltimer stamps (ns), then import statistics (TSTA_PSA_TIMELINE=1)
{
    if (!p || !out || !p->d_timeline.p) return set_error(TSTA_EINVAL, "no timeline recorded (set TSTA_PSA_TIMELINE=1)");
    TSTA_CUDA_TRY(cudaSetDevice(p->device));
    TSTA_CUDA_TRY(cudaMemcpy(out, p->d_timeline.p, (size_t)p->nbands * 128, cudaMemcpyDeviceToHost));
    return TSTA_OK;
}

int tsta_psa_plan_trace_from(tsta_psa_plan* p, int64_t start_i, int start_forced,
                             int64_t* end_i, int* end_forced, int64_t* nops, float* ms)
{
    if (!p || !p->ran || !p->want_trace) return set_error(TSTA_EINVAL, "plan not run with traceback");
    if (!end_i || !end_forced || !nops) return set_error(TSTA_EINVAL, "NULL argument");
    if (start_i < 0 || start_i >= p->m || start_forced < 0 || start_forced > 2)
        return set_error(TSTA_EINVAL, "bad start of the walk (row %lld, state %d)", (long long)start_i, start_forced);
    TSTA_CUDA_TRY(cudaSetDevice(p->device));
    int* ctl = p->d_ctl.as<int>();
    TSTA_CUDA_TRY(cudaEventRecord(p->ev[1], p->stream));
    if (int rc = launch_walk(p, start_i, p->n - 1, start_forced, 0)) return rc;
    TSTA_CUDA_TRY(cudaEventRecord(p->ev[2], p->stream));
    TSTA_CUDA_TRY(cudaMemcpyAsync(p->h_small, ctl, 8 * sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    TSTA_CUDA_TRY(cudaStreamSynchronize(p->stream));
    *nops = p->h_small[2];
    *end_i = p->h_small[3];
    *end_forced = p->h_small[5];
    if (ms) TSTA_CUDA_TRY(cudaEventElapsedTime(ms, p->ev[1], p->ev[2]));
    return TSTA_OK;
}

int tsta_psa_plan_ops(tsta_psa_plan* p, uint8_t* out, int64_t nops)
{
    if (!p || !out || !p->want_trace || nops < 0 || nops > p->n + p->m) return set_error(TSTA_EINVAL, "bad arguments");
    TSTA_CUDA_TRY(cudaSetDevice(p->device));
    return fetch_moves(p, out, nops);
}

void tsta_psa_result_free(tsta_psa_result* r)
{
    if (!r) return;
    free(r->a);
    free(r->b);
    r->a = r->b = nullptr;
}

int tsta_psa_align(const tsta_params* prm, const char* s0, int64_t n, const char* s1, int64_t m,
                   int want_trace, tsta_psa_result* out)
{
    if (!out) return set_error(TSTA_EINVAL, "result pointer is NULL");
    // one cached plan per (thread, trace mode): repeated calls reuse device buffers (a traced 100 kbp
    // pair keeps its ~7 GB direction matrix; tsta_psa_release_cache() or thread exit returns it)
    const int k = want_trace ? 1 : 0;
    tsta_psa_plan*& plan = g_align_cache.plan[k];
    if (!plan) {
        if (int rc = tsta_psa_plan_create(&plan, k)) return rc;
    }
    if (int rc = tsta_psa_plan_upload(plan, prm, s0, n, s1, m)) return rc;
    if (int rc = tsta_psa_plan_run(plan, nullptr, nullptr)) return rc;
    return tsta_psa_plan_fetch(plan, out);
}

void tsta_psa_release_cache(void) { g_align_cache.release(); tsta::batch_release_ctx(); }

}  // extern "C"
