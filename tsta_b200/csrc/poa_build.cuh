// poa_build.cuh -- the row tables of the POA fill built ON THE DEVICE from a graph that stays resident in HBM.
//
// Reference: every read re-sorts the graph and the DP re-reads all of it (msa/msa.c:132-133, msa/topo.c:243-260).
// The per-read host export + re-upload of round 1 (CSR, base, hneg1, 16-byte descriptors: ~30 bytes per row)
// becomes: the node table (base, in-degree, up to eight predecessor NODE IDS) lives on the device; after a read is
// fused the host sends only the records of the nodes that changed and the permutation t_sort produced (4 bytes per
// row); these kernels turn that into what the fill and the walker read -- row descriptors (poa_fill.cuh), a
// fixed-stride predecessor table in row indices, spill slots.  H(r,-1) is computed by the fill itself
// (poa_rows.cuh, EDGE0).
#pragma once
#include "poa_fill.cuh"
#include "poa_rows.cuh"

namespace tsta {

constexpr int kPoaNodePreds = 8;             // predecessors a resident node record holds (= kPoaFast)

// changed nodes, two record widths (most nodes have one or two predecessors)
struct PoaNodeNarrow { int id; unsigned base_in; int pre[2]; };              // 16 bytes
struct PoaNodeWide { int id; unsigned base_in; int pre[kPoaNodePreds]; };    // 40 bytes

struct PoaNodeTable {
    unsigned* meta;          // [nodes] base | in-degree << 8   (in-degree 0: a head)
    int* pre;                // [nodes][8] predecessor node ids
    int* sub;                // [nodes] row of the node in the current order
};

__global__ void __launch_bounds__(256) poa_apply_narrow_kernel(PoaNodeTable t, const PoaNodeNarrow* recs, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const PoaNodeNarrow r = recs[i];
    t.meta[r.id] = r.base_in;
    t.pre[(size_t)r.id * kPoaNodePreds] = r.pre[0];
    t.pre[(size_t)r.id * kPoaNodePreds + 1] = r.pre[1];
}

__global__ void __launch_bounds__(256) poa_apply_wide_kernel(PoaNodeTable t, const PoaNodeWide* recs, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const PoaNodeWide& r = recs[i];
    t.meta[r.id] = r.base_in;
#pragma unroll
    for (int j = 0; j < kPoaNodePreds; j++) t.pre[(size_t)r.id * kPoaNodePreds + j] = r.pre[j];
}

__global__ void __launch_bounds__(256) poa_rows_of_nodes_kernel(PoaNodeTable t, const int* sort, int nrows, int* keep)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    t.sub[sort[r]] = r;
    keep[r] = 0;
}

// One thread per row: what the host loop of tsta_poa_fill does (descriptor words x, y, z; w comes from the next
// kernel), plus the predecessor rows at stride 8 (the "CSR" of the slow path and of the walker: row_ptr[r] = 8 r).
__global__ void __launch_bounds__(256) poa_build_rows_kernel(PoaNodeTable t, const int* sort, int nrows, int L,
                                                             uint4* desc, int* pred_tab, int* keep)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const int node = sort[r];
    const unsigned meta = t.meta[node];
    const int in = (int)((meta >> 8) & 0xffu);
    const int np = in > 0 ? in : 1;
    unsigned flags = 0, jf = kPoaNoFwd;
    unsigned slot[kPoaFast];
#pragma unroll
    for (int j = 0; j < kPoaFast; j++) slot[j] = kPoaSlotNone;
    int* prow = pred_tab + (size_t)r * kPoaNodePreds;
    if (in == 0) {
        flags |= kPoaHead;
        if (L > 0 && r >= L) flags |= kPoaLate;
        slot[0] = (flags & kPoaLate) ? kPoaSlotLate : kPoaSlotHead;
        prow[0] = -1;
    } else {
#pragma unroll
        for (int j = 0; j < kPoaNodePreds; j++) {
            if (j < np) {
                const int pr = t.sub[t.pre[(size_t)node * kPoaNodePreds + j]];
                prow[j] = pr;
                if (r - pr > kPoaNear) {
                    atomicOr(keep + pr, 1);
                    flags |= kPoaSlow;
                }
                if (pr == r - 1 && jf != kPoaNoFwd) flags |= kPoaSlow;
                slot[j] = (unsigned)(pr & (kPoaHist - 1));
                if (pr == r - 1) jf = (unsigned)j;
            }
        }
    }
    uint4 q;
    q.x = (meta & 0xffu) | ((unsigned)np << 8) | (flags << 16) | (jf << 28);
    q.y = slot[0] | (slot[1] << 8) | (slot[2] << 16) | (slot[3] << 24);
    q.z = (flags & kPoaSlow) ? (unsigned)(r * kPoaNodePreds) : (slot[4] | (slot[5] << 8) | (slot[6] << 16) | (slot[7] << 24));
    q.w = 0;
    desc[r] = q;
}

// Spill slots for the rows a far successor reads back (any unique numbering will do), and the "plain row" flag,
// which needs to know that the row is not spilled.
__global__ void __launch_bounds__(256) poa_build_keep_kernel(int nrows, uint4* desc, const int* keep, int* nkeep)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    unsigned x = desc[r].x;
    if (keep[r]) {
        desc[r].w = (unsigned)atomicAdd(nkeep, 1);
        x |= kPoaKeep << 16;
    } else if (((x >> 8) & 0xffu) == 1u && (x >> 28) == 0u && !((x >> 16) & (kPoaHead | kPoaSlow | kPoaLate)))
        x |= kPoaPlainRow << 16;
    desc[r].x = x;
}

}  // namespace tsta
