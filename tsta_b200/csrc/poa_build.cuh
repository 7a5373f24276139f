// poa_build.cuh -- the row tables of the POA fill built ON THE DEVICE from a graph that stays resident in HBM.
//
// Reference: every read re-sorts the graph and the DP re-reads all of it (msa/msa.c:132-133, msa/topo.c:243-260).
// The per-read host export + re-upload of round 1 (CSR, base, hneg1, 16-byte descriptors: ~30 bytes per row)
// becomes: the node table (base, in-degree, up to eight predecessor NODE IDS) lives on the device; after a read is
// fused the host sends only the records of the nodes that changed and the permutation t_sort produced (3 bytes per
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

// Changed node records as they cross PCIe: n1 records of 12 bytes (in-degree <= 1: id, base_in, pre0), then n2 of
// 16 bytes (in-degree 2), then n8 of 40 bytes (up to eight predecessors), back to back.
__global__ void __launch_bounds__(256) poa_apply_records_kernel(PoaNodeTable t, const int* buf, int n1, int n2, int n8)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n1) {
        const int* r = buf + (size_t)i * 3;
        t.meta[r[0]] = (unsigned)r[1];
        t.pre[(size_t)r[0] * kPoaNodePreds] = r[2];
    } else if (i < n1 + n2) {
        const int* r = buf + (size_t)n1 * 3 + (size_t)(i - n1) * 4;
        t.meta[r[0]] = (unsigned)r[1];
        t.pre[(size_t)r[0] * kPoaNodePreds] = r[2];
        t.pre[(size_t)r[0] * kPoaNodePreds + 1] = r[3];
    } else if (i < n1 + n2 + n8) {
        const int* r = buf + (size_t)n1 * 3 + (size_t)n2 * 4 + (size_t)(i - n1 - n2) * 10;
        t.meta[r[0]] = (unsigned)r[1];
#pragma unroll
        for (int j = 0; j < kPoaNodePreds; j++) t.pre[(size_t)r[0] * kPoaNodePreds + j] = r[2 + j];
    }
}

// The permutation arrives as 3-byte node ids (4-byte when the graph has 2^24 nodes or more): unpack it, note the row
// of every node, clear the spill marks.
__global__ void __launch_bounds__(256) poa_rows_of_nodes_kernel(PoaNodeTable t, const uint8_t* packed, int id_bytes, int* sort,
                                                                int nrows, int* keep)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const uint8_t* q = packed + (size_t)r * id_bytes;
    int id = (int)q[0] | ((int)q[1] << 8) | ((int)q[2] << 16);
    if (id_bytes == 4) id |= (int)q[3] << 24;
    sort[r] = id;
    t.sub[id] = r;
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
