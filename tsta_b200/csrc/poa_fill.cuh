// poa_fill.cuh -- sequence-to-graph DP (POA fill) and its traceback walker on sm_100a.
//
// Replaces the tile loop of control() (reference msa/c-t-simd.c:1019-1043) that feeds
// block_alignment()/block_line_alignment() (165-657) to the pthread pool, and the part of
// node_fuse() that walks the per-node trace rows (751-785).
//
// Rows are graph nodes in topological (p->sort[]) order, columns are read positions.  Every
// predecessor of a row has a smaller row index, so the plain skew "row + column strip" is a
// legal wavefront -- the same band/lane decomposition as the pairwise fill:
//   * the read is cut into bands of 32*C columns; one warp sweeps one band over all rows;
//   * lane k owns C columns and runs one row behind lane k-1; the strip's right edge (H, Fg)
//     moves to the next lane by shuffle and to the next band through tagged 64-bit elements
//     in L2 (no fences);
//   * unlike the pairwise case a row's inputs are ARBITRARY earlier rows (the node's pre[]),
//     so every lane keeps the (H, Eout) of its last D rows in a shared-memory history ring;
//     predecessors further back than the ring (rare: ~2 % of edges span more than 4 rows) are
//     read back from the global (H, Eout) matrix that every row also writes.
// Recurrences, tie-breaks and the two quirks (late heads Q1, 'N' columns) are the ones
// restated in oracle/poa_oracle.c; predecessor order decides ties (lowest index wins).
//
// Trace: 2 bytes per cell, row-major, written with 128-bit stores from a shared ring:
//   byte 0  bits 0-6 `source` (42*type + predecessor index; 126 = horizontal), bit 7 = cH:
//           "a horizontal move from this cell forces the next cell horizontal"  (767-769)
//   byte 1  bits 0-5 j* (index of the predecessor whose Eout is largest), bit 6 = that Eout
//           was opened at the predecessor, bit 7 = its extension ties with opening here
//           (the sign of the reference's esource, 509-538)
#pragma once
#include "common.cuh"
#include "psa_fill.cuh"   // ld_relaxed_u64 / st_relaxed_u64_if, lds/sts helpers

namespace tsta {

constexpr int kPoaC = 4;            // columns per lane
constexpr int kPoaHist = 32;        // history ring depth in rows (power of two)
constexpr int kPoaNear = kPoaHist - 2;   // predecessors at most this many rows back are read from the ring
constexpr int kPoaChunk = 16;       // rows per staging chunk
constexpr int kPoaBndRing = 64;     // boundary ring (rows)
constexpr int kPoaTrRing = 64;      // trace ring (rows)
constexpr int kPoaNegInf = -1000000000;

// Row descriptor, built on the host from the CSR graph and staged through shared memory a chunk
// ahead of the wavefront, so that the inner loop never waits for a dependent global load:
//   x: base | in-degree << 8 | flags << 16      y, z, w: six predecessor distances (r - p), 16 bit each
constexpr unsigned kPoaHead = 1u;        // in == 0: the single "predecessor" is the virtual row
constexpr unsigned kPoaKeep = 2u;        // some successor is further away than the history ring: also store (H, Eout) in HBM
constexpr unsigned kPoaSlow = 4u;        // more than six predecessors or a distance >= 65536: read the CSR lists instead
constexpr int kPoaInline = 6;
constexpr int kPoaDescRing = 64;

struct PoaFillArgs {
    const uint4* desc;        // [nrows] row descriptors
    const int* row_ptr;       // [nrows+1]
    const int* pred_row;      // [nedges], -1 = virtual predecessor of a head
    const uint8_t* base;      // [nrows]
    const int* hneg1;         // [nrows] H(r,-1)
    const uint8_t* read;      // [len]
    int len, nrows, nbands, L;
    int M, X, E, O;
    int2* HE;                 // [nrows][stride] (H, Eout) of every cell
    long long stride;         // padded read length (multiple of 32*C)
    uint16_t* trace;          // [nrows][stride]
    unsigned long long* bnd;  // [nbands][nrows] tagged right-boundary elements
    unsigned epoch;           // 1..32767
    unsigned* ticket;
    int* lastsorce;           // [nrows] H(r, len-1)
};

struct PoaSmem {
    int2 hist[kPoaHist][32][kPoaC];          // (H, Eout) of the last kPoaHist rows, per lane
    uint16_t tr[kPoaTrRing][32 * kPoaC];     // trace ring
    int2 bring[kPoaBndRing];                 // left boundary (H(r, j0-1), Fg(r, j0)) by row
    uint4 desc[kPoaDescRing];                // row descriptors
    uint8_t bneg[kPoaBndRing];               // fsource(r, j0-1) < 0
};

// boundary element: low word H; high word: bits 0-15 Fg-H, bit 16 fsource sign, bits 17-31 epoch
__device__ __forceinline__ unsigned long long poa_bnd_pack(int H, int F, int fneg, unsigned epoch)
{
    const unsigned hi = ((unsigned)(F - H) & 0xffffu) | ((unsigned)fneg << 16) | (epoch << 17);
    return ((unsigned long long)hi << 32) | (unsigned)H;
}

__global__ void __launch_bounds__(32) poa_fill_kernel(const __grid_constant__ PoaFillArgs a)
{
    constexpr int C = kPoaC;
    extern __shared__ __align__(16) unsigned char poa_smem_raw[];
    PoaSmem& sm = *reinterpret_cast<PoaSmem*>(poa_smem_raw);
    const int lane = threadIdx.x;
    int band = 0;
    if (lane == 0) band = (int)atomicAdd(a.ticket, 1u);
    band = __shfl_sync(kFullMask, band, 0);

    const int E = a.E, O = a.O, OE = a.O + a.E, M = a.M, X = a.X;
    const int R = a.nrows, len = a.len, L = a.L;
    const int j0 = (band * 32 + lane) * C;
    uint8_t rc[C];                       // read bytes of this lane's columns (0 = pad: matches nothing)
#pragma unroll
    for (int c = 0; c < C; c++) rc[c] = (j0 + c < len) ? a.read[j0 + c] : 0;
    const bool owns_last = (len - 1 >= j0) && (len - 1 < j0 + C);
    bool hasN = false;                   // quirk Q3 applies to this lane's strip at all?
#pragma unroll
    for (int c = 0; c < C; c++) hasN |= rc[c] == 'N';
    const unsigned long long* bnd_in = a.bnd + (long long)(band - 1) * R;
    unsigned long long* bnd_out = a.bnd + (long long)band * R;
    const bool store_right = (band + 1 < a.nbands) && lane == 31;

    int Hout = 0, Fout = 0, Nout = 0;    // right edge of the previous step: H, Fg(next col), fsource<0
    const int nchunks = (R + 31 + kPoaChunk - 1) / kPoaChunk;
    unsigned long long pre_b = 0;
    if (band > 0 && lane < kPoaChunk && lane < R) pre_b = ld_relaxed_u64(bnd_in + lane);
    // descriptors: lanes 16..31 stage rows [t0+16, t0+32) at the top of chunk t0 (one chunk ahead)
    if (lane < kPoaChunk && lane < R) sm.desc[lane] = __ldg(a.desc + lane);
    uint4 pre_d = make_uint4(0, 0, 0, 0);
    if (lane >= 16 && lane - 16 + kPoaChunk < R) pre_d = __ldg(a.desc + (lane - 16 + kPoaChunk));

    for (int ch = 0; ch < nchunks; ch++) {
        const int t0 = ch * kPoaChunk;
        // ---- stage next chunk's descriptors, this chunk's left boundary --------------------------
        if (lane >= 16) {
            const int rown = t0 + kPoaChunk + (lane - 16);
            if (rown < R) sm.desc[rown & (kPoaDescRing - 1)] = pre_d;
            const int rownn = rown + kPoaChunk;
            if (rownn < R) pre_d = __ldg(a.desc + rownn);
        }
        {
            const int row = lane < kPoaChunk ? t0 + lane : R;
            if (band > 0) {
                while (!__all_sync(kFullMask, row >= R || (unsigned)(pre_b >> 49) == a.epoch)) {
                    __nanosleep(20);
                    if (row < R && (unsigned)(pre_b >> 49) != a.epoch) pre_b = ld_relaxed_u64(bnd_in + row);
                }
                if (row < R) {
                    const int H = (int)(unsigned)pre_b;
                    const unsigned hi = (unsigned)(pre_b >> 32);
                    sm.bring[row & (kPoaBndRing - 1)] = make_int2(H, H + (int)(short)(hi & 0xffffu));
                    sm.bneg[row & (kPoaBndRing - 1)] = (uint8_t)((hi >> 16) & 1u);
                }
                const int rowp = row + kPoaChunk;
                if (rowp < R) pre_b = ld_relaxed_u64(bnd_in + rowp);
            } else if (row < R) {
                const int h = a.hneg1[row];
                sm.bring[row & (kPoaBndRing - 1)] = make_int2(h, h + OE);   // Fg(r,0) = H(r,-1)+O+E
                sm.bneg[row & (kPoaBndRing - 1)] = 0;
            }
        }
        __syncwarp();

#pragma unroll 1
        for (int s = 0; s < kPoaChunk; s++) {
            const int r = t0 + s - lane;
            int Hleft = __shfl_up_sync(kFullMask, Hout, 1);
            int F = __shfl_up_sync(kFullMask, Fout, 1);
            int fneg_left = __shfl_up_sync(kFullMask, Nout, 1);
            __syncwarp();   // orders last step's history stores before this step's loads
            if ((unsigned)r < (unsigned)R) {
                if (lane == 0) {
                    const int2 v = sm.bring[r & (kPoaBndRing - 1)];
                    Hleft = v.x;
                    F = v.y;
                    fneg_left = sm.bneg[r & (kPoaBndRing - 1)];
                }
                const uint4 dsc = sm.desc[r & (kPoaDescRing - 1)];
                const int b = (int)(dsc.x & 0xffu);
                const int npre = (int)((dsc.x >> 8) & 0xffu);
                const unsigned flags = dsc.x >> 16;
                const bool head = flags & kPoaHead;
                const bool late = head && L > 0 && r >= L;

                // Reduction over the predecessors with KEYS: the history holds H*64 and
                // Eout*128 + "opened", so  max(H*64 + (63-j))  and  max(Eout*128 + 2*(63-j) + opened)
                // give the maximum, the LOWEST predecessor index that attains it (the reference's
                // tie-break) and the opened flag of exactly that predecessor in one VIADDMNMX each.
                int kd[C], ke[C];
                if (head) {
#pragma unroll
                    for (int c = 0; c < C; c++) {
                        const int col = j0 + c;
                        int Hd, Hp;
                        if (late && col <= L - 1) { Hd = col * E; Hp = (col + 1) * E; }       // quirk Q1
                        else { Hd = col == 0 ? 0 : O + col * E; Hp = O + (col + 1) * E; }
                        kd[c] = Hd * 64 + 63;                       // predecessor index 0
                        ke[c] = (Hp + OE) * 128 + 2 * 63 + 1;       // Eout_v = Hv + O + E: opened
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < C; c++) { kd[c] = INT_MIN / 2; ke[c] = INT_MIN / 2; }
                    const int e0 = (flags & kPoaSlow) ? __ldg(a.row_ptr + r) : 0;
                    unsigned dw = dsc.y;
                    for (int j = 0; j < npre; j++) {
                        int p;
                        if (flags & kPoaSlow) p = __ldg(a.pred_row + e0 + j);
                        else {
                            p = r - (int)(dw & 0xffffu);
                            dw = (j & 1) ? (j == 1 ? dsc.z : dsc.w) : (dw >> 16);
                        }
                        int4 v01, v23;
                        int Hd0;
                        if (r - p <= kPoaNear) {
                            const int slot = p & (kPoaHist - 1);
                            v01 = *reinterpret_cast<const int4*>(&sm.hist[slot][lane][0]);
                            v23 = *reinterpret_cast<const int4*>(&sm.hist[slot][lane][2]);
                            const int hl = sm.hist[slot][lane > 0 ? lane - 1 : 0][C - 1].x;
                            const int hb = sm.bring[p & (kPoaBndRing - 1)].x * 64;          // (band 0: H(p,-1))
                            Hd0 = lane > 0 ? hl : hb;
                        } else {
                            const int2* src = a.HE + (long long)p * a.stride + j0;
                            v01 = __ldcg(reinterpret_cast<const int4*>(src));
                            v23 = __ldcg(reinterpret_cast<const int4*>(src + 2));
                            if (lane > 0) Hd0 = __ldcg(src - 1).x;
                            else if (band > 0) Hd0 = (int)(unsigned)ld_relaxed_u64(bnd_in + p) * 64;
                            else Hd0 = __ldg(a.hneg1 + p) * 64;
                        }
                        const int wd = 63 - j, we = 2 * (63 - j);
                        kd[0] = __viaddmax_s32(Hd0, wd, kd[0]);
                        kd[1] = __viaddmax_s32(v01.x, wd, kd[1]);
                        kd[2] = __viaddmax_s32(v01.z, wd, kd[2]);
                        kd[3] = __viaddmax_s32(v23.x, wd, kd[3]);
                        ke[0] = __viaddmax_s32(v01.y, we, ke[0]);
                        ke[1] = __viaddmax_s32(v01.w, we, ke[1]);
                        ke[2] = __viaddmax_s32(v23.y, we, ke[2]);
                        ke[3] = __viaddmax_s32(v23.w, we, ke[3]);
                    }
                }

                // ---- the row itself: serial along the strip --------------------------------------
                int2 outv[C];
                uint16_t tw[C];
                int Hlast = 0;
#pragma unroll
                for (int c = 0; c < C; c++) {
                    const int col = j0 + c;
                    if (late && col == L) F += O;                              // quirk Q1: carried F re-based
                    const int bd = kd[c] >> 6, bdj = 63 - (kd[c] & 63);
                    const int be = ke[c] >> 7, bej = 63 - ((ke[c] >> 1) & 63), eopen = ke[c] & 1;
                    const bool match = (int)rc[c] == b;
                    const int h = bd + (match ? M : X);
                    // candidates tagged with the tie-break priority of c-t-simd.c:484-506:
                    // diagonal (3) beats vertical E (2) beats horizontal F (1)
                    const int Hk = __vimax3_s32(h * 4 + 3, be * 4 + 2, F * 4 + 1);
                    int H = Hk >> 2;
                    const int s3 = (match ? 42 : 84) + bdj;
                    int src;
                    if (hasN && rc[c] == 'N') {                                // quirk Q3: 'N' columns copy H from the left
                        H = Hleft;
                        src = 126;
                        if (be == H) src = bej;
                        if (h == H) src = s3;
                    } else {
                        const int tag = Hk & 3;
                        src = tag == 3 ? s3 : (tag == 2 ? bej : 126);
                    }
                    const int t = H + OE, ae = be + E;
                    const int eneg = (ae == t);
                    const int fabs2 = (F - Hleft == OE);                       // |fsource| == 2
                    const int fneg = (F + E == t);
                    const int cH = (!fabs2) || fneg_left;                      // (767-769), column >= 2 checked by the walker
                    tw[c] = (uint16_t)(src | (cH << 7) | (bej << 8) | (eopen << 14) | (eneg << 15));
                    // Eout = max(H+O+E, be+E) with "opened" (ties count as opened) in bit 0 of the doubled max
                    const int z = max(2 * ae, 2 * t + 1);
                    outv[c] = make_int2(H * 64, (z & ~1) * 64 + (z & 1));      // (H*64, Eout*128 + opened)
                    F = max(F + E, t);
                    Hleft = H;
                    fneg_left = fneg;
                    if (j0 + c == len - 1) Hlast = H;
                }
                Hout = Hleft; Fout = F; Nout = fneg_left;
                // history ring, HBM spill for far successors, trace ring
                {
                    const int slot = r & (kPoaHist - 1);
                    *reinterpret_cast<int4*>(&sm.hist[slot][lane][0]) = make_int4(outv[0].x, outv[0].y, outv[1].x, outv[1].y);
                    *reinterpret_cast<int4*>(&sm.hist[slot][lane][2]) = make_int4(outv[2].x, outv[2].y, outv[3].x, outv[3].y);
                    if (flags & kPoaKeep) {    // a successor further away than the history ring will read it back
                        int2* g = a.HE + (long long)r * a.stride + j0;
                        *reinterpret_cast<int4*>(g) = make_int4(outv[0].x, outv[0].y, outv[1].x, outv[1].y);
                        *reinterpret_cast<int4*>(g + 2) = make_int4(outv[2].x, outv[2].y, outv[3].x, outv[3].y);
                    }
                    *reinterpret_cast<uint2*>(&sm.tr[r & (kPoaTrRing - 1)][lane * C]) =
                        make_uint2((uint32_t)tw[0] | ((uint32_t)tw[1] << 16), (uint32_t)tw[2] | ((uint32_t)tw[3] << 16));
                }
                if (owns_last) a.lastsorce[r] = Hlast;
                st_relaxed_u64_if(bnd_out + r, poa_bnd_pack(Hout, Fout, Nout, a.epoch), store_right);
            }
        }

        // ---- flush finished trace rows [lo, hi): 256 B per band row, 16 B per lane-unit ---------
        __syncwarp();
        {
            const int hi = min(R, max(0, t0 + kPoaChunk - 31));
            const int lo = min(R, max(0, t0 - 31));
            constexpr int kUnits = 32 * C * 2 / 16;    // 16 units per row
            const int units = (hi - lo) * kUnits;
            for (int u = lane; u < units; u += 32) {
                const int row = lo + u / kUnits, part = u % kUnits;
                const uint4 v = *reinterpret_cast<const uint4*>(&sm.tr[row & (kPoaTrRing - 1)][part * 8]);
                __stcs(reinterpret_cast<uint4*>(a.trace + (long long)row * a.stride + (long long)band * 32 * C + part * 8), v);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// traceback walker (node_fuse control flow, c-t-simd.c:751-956, with the forced
// continuations folded into the automaton).  One CTA; a window of the trace matrix and the
// window's slice of the CSR graph are staged in shared memory, one thread walks.
// ---------------------------------------------------------------------------------------
constexpr int kPoaWinRows = 128;
constexpr int kPoaWinCols = 128;
constexpr int kPoaWinEdges = 2048;
constexpr int kPoaWalkThreads = 256;

struct PoaStep {
    int row, col;
    uint8_t kind, pred;
    uint16_t pad;
};

struct PoaWalkArgs {
    const uint16_t* trace;
    long long stride;
    const int* row_ptr;
    const int* pred_row;
    int len, nrows, start_row;
    PoaStep* path;
    int* nsteps;
};

__global__ void __launch_bounds__(kPoaWalkThreads) poa_walk_kernel(const PoaWalkArgs a)
{
    __shared__ __align__(16) uint16_t win[kPoaWinRows][kPoaWinCols];
    __shared__ int s_ptr[kPoaWinRows + 1];
    __shared__ int s_pred[kPoaWinEdges];
    __shared__ int s_row, s_col, s_forced, s_n;
    if (threadIdx.x == 0) { s_row = a.start_row; s_col = a.len - 1; s_forced = -1; s_n = 0; }
    __syncthreads();
    while (true) {
        const int r0 = s_row, c0 = s_col;
        if (r0 < 0 || c0 < 0) break;
        // window: rows (r0-128, r0], columns [w0, w0+128) with c0 near the right end
        int w0 = ((c0 + 8) & ~7) - kPoaWinCols;
        if (w0 < 0) w0 = 0;
        const int rlo = max(0, r0 - kPoaWinRows + 1);
        for (int u = threadIdx.x; u < kPoaWinRows * (kPoaWinCols / 8); u += kPoaWalkThreads) {
            const int rr = u / (kPoaWinCols / 8), q = u % (kPoaWinCols / 8);
            const int row = r0 - rr;
            const long long col = (long long)w0 + q * 8;
            if (row >= 0 && col < a.stride)
                *reinterpret_cast<uint4*>(&win[rr][q * 8]) =
                    __ldcs(reinterpret_cast<const uint4*>(a.trace + (long long)row * a.stride + col));
        }
        for (int u = threadIdx.x; u <= r0 - rlo + 1; u += kPoaWalkThreads) s_ptr[u] = a.row_ptr[rlo + u];
        __syncthreads();
        const int ebase = s_ptr[0];
        const int ecount = min(s_ptr[r0 - rlo + 1] - ebase, kPoaWinEdges);
        for (int u = threadIdx.x; u < ecount; u += kPoaWalkThreads) s_pred[u] = a.pred_row[ebase + u];
        __syncthreads();
        if (threadIdx.x == 0) {
            int r = r0, c = c0, forced = s_forced, n = s_n;
            while (r >= rlo && r >= 0 && c >= w0) {
                const uint16_t w = win[r0 - r][c - w0];
                const int S = forced >= 0 ? forced : (w & 0x7f);
                const int kind = S / 42, pj = S % 42;
                PoaStep st;
                st.row = r; st.col = c; st.kind = (uint8_t)kind; st.pred = (uint8_t)(kind == 3 ? 0 : pj); st.pad = 0;
                a.path[n++] = st;
                if (kind == 3) {
                    forced = (c - 1 > 0 && (w & 0x80)) ? 126 : -1;
                    c--;
                    continue;
                }
                const int e = s_ptr[r - rlo] - ebase + pj;
                const int p = e < kPoaWinEdges ? s_pred[e] : a.pred_row[ebase + e];
                if (kind == 0) {
                    // vertical: the predecessor keeps going up when this cell's best E was an
                    // extension there, or was opened there and ties with extending (776-782)
                    forced = -1;
                    if (p > 0) {
                        const int eopen = (w >> 14) & 1;
                        const uint16_t wp = (p >= rlo) ? win[r0 - p][c - w0]
                                                      : a.trace[(long long)p * a.stride + c];
                        if (!eopen || (wp >> 15)) forced = (wp >> 8) & 0x3f;
                    }
                    r = p;
                } else {
                    forced = -1;
                    r = p;
                    c--;
                }
                if (r < 0) break;
            }
            s_row = r; s_col = c; s_forced = forced; s_n = n;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *a.nsteps = s_n;
}

}  // namespace tsta
