/* poa_graph.c -- see poa_graph.h.  Host-side partial-order graph of TSTA_msa. */
#include "poa_graph.h"

#include <limits.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * node storage
 * ---------------------------------------------------------------------------------------- */
static int node_new(poa_graph* g, unsigned char base)
{
    if (g->nnode == g->cap) {
        g->cap = g->cap ? g->cap * 2 : 1024;
        g->node = (poa_node*)realloc(g->node, (size_t)g->cap * sizeof(poa_node));
    }
    poa_node* n = &g->node[g->nnode];
    memset(n, 0, sizeof *n);
    n->base = base;
    n->sub = -1;
    n->alive = 1;
    return g->nnode++;
}

/* pool of small arrays: power-of-two size classes (in 8-byte units), bump allocation from 1 MiB
 * blocks, per-class free lists threaded through the freed arrays themselves */
#define POOL_BLOCK ((size_t)1 << 20)
static int pool_class(size_t bytes)
{
    int k = 0;
    while (((size_t)8 << k) < bytes) k++;
    return k;
}
static void* pool_get(poa_graph* g, size_t bytes)
{
    const int k = pool_class(bytes);
    const size_t sz = (size_t)8 << k;
    if (g->pool_free[k]) {
        void* p = g->pool_free[k];
        g->pool_free[k] = *(void**)p;
        return p;
    }
    if (sz > POOL_BLOCK / 4) {                       /* huge: its own block, kept on the block list */
        void** b = (void**)malloc(sz + 16);
        b[0] = g->pool_blocks; g->pool_blocks = b;
        return (char*)b + 16;
    }
    if (g->pool_left < sz) {
        void** b = (void**)malloc(POOL_BLOCK);
        b[0] = g->pool_blocks; g->pool_blocks = b;
        g->pool_cur = (char*)b + 16;
        g->pool_left = POOL_BLOCK - 16;
    }
    void* p = g->pool_cur;
    g->pool_cur += sz; g->pool_left -= sz;
    return p;
}
static void pool_put(poa_graph* g, void* p, size_t bytes)
{
    if (!p) return;
    const int k = pool_class(bytes);
    *(void**)p = g->pool_free[k];
    g->pool_free[k] = p;
}
static int* pool_grow(poa_graph* g, int* old, int used, int old_cap, int new_cap)
{
    int* p = (int*)pool_get(g, (size_t)new_cap * sizeof(int));
    if (used > 0) memcpy(p, old, (size_t)used * sizeof(int));
    pool_put(g, old, (size_t)old_cap * sizeof(int));
    return p;
}
static uint8_t* passing_new(poa_graph* g)
{
    uint8_t* p = (uint8_t*)pool_get(g, (size_t)g->nseq);
    memset(p, 0, (size_t)g->nseq);
    return p;
}

static void push_pre(poa_graph* g, poa_node* n, int id, int weight)
{
    if (n->in + 1 > n->pre_cap) {
        const int cap = n->pre_cap ? n->pre_cap * 2 : 2;
        /* (a head keeps pre[0] = sentinel with in == 0: copy the whole old array) */
        n->pre = pool_grow(g, n->pre, n->pre_cap, n->pre_cap, cap);
        n->ew = pool_grow(g, n->ew, n->pre_cap, n->pre_cap, cap);
        n->pre_cap = cap;
    }
    n->pre[n->in] = id;
    n->ew[n->in] = weight;
    n->in++;
}

static void push_next(poa_graph* g, poa_node* n, int id)
{
    if (n->out + 1 > n->next_cap) {
        const int cap = n->next_cap ? n->next_cap * 2 : 2;
        n->next = pool_grow(g, n->next, n->out, n->next_cap, cap);
        n->next_cap = cap;
    }
    n->next[n->out++] = id;
}

static void node_kill(poa_graph* g, int id)
{
    poa_node* n = &g->node[id];
    pool_put(g, n->pre, (size_t)n->pre_cap * sizeof(int));
    pool_put(g, n->ew, (size_t)n->pre_cap * sizeof(int));
    pool_put(g, n->next, (size_t)n->next_cap * sizeof(int));
    pool_put(g, n->passing_seq, (size_t)g->nseq);
    n->pre = n->ew = n->next = NULL;
    n->passing_seq = NULL;
    n->pre_cap = n->next_cap = 0;
    n->alive = 0;
}

static void order_reserve(poa_graph* g, int want)
{
    if (want <= g->order_cap) return;
    while (g->order_cap < want) g->order_cap = g->order_cap ? g->order_cap * 2 : 1024;
    g->unsort = (int*)realloc(g->unsort, (size_t)g->order_cap * sizeof(int));
    g->sort = (int*)realloc(g->sort, (size_t)g->order_cap * sizeof(int));
}

/* the virtual predecessor of heads behaves like a node with these values (c-t-simd.c:81-94) */
static inline int pre_sorce(const poa_graph* g, int id) { return id == POA_SENTINEL ? 0 : g->node[id].node_sorce; }
static inline int pre_len(const poa_graph* g, int id) { return id == POA_SENTINEL ? 0 : g->node[id].node_base_len; }
static inline int pre_sub(const poa_graph* g, int id) { return id == POA_SENTINEL ? -1 : g->node[id].sub; }

/* ------------------------------------------------------------------------------------------
 * build from the first read: a chain (c-t-simd.c:72-163)
 * ---------------------------------------------------------------------------------------- */
void poa_graph_init(poa_graph* g, const char* a, int nseq)
{
    memset(g, 0, sizeof *g);
    g->nseq = nseq;
    const int len = (int)strlen(a);
    order_reserve(g, len > 0 ? len : 1);
    {
        /* Node ids are never reused and every read adds one provisional node per base: a job ends with about
         * (reads + 1) x read length node records.  Reserve them now (untouched pages cost nothing) instead of
         * doubling -- and copying up to ~50 MB -- four or five times while reads are fused. */
        const size_t want = (size_t)(nseq > 0 ? nseq + 1 : 2) * (size_t)(len > 0 ? len : 1) + 1024;
        if (want < ((size_t)1 << 27)) {
            poa_node* arr = (poa_node*)malloc(want * sizeof(poa_node));
            if (arr) { g->node = arr; g->cap = (int)want; }
        }
    }
    for (int i = 0; i < len; i++) {
        const int id = node_new(g, (unsigned char)a[i]);
        poa_node* n = &g->node[id];
        n->passing_seq = passing_new(g);
        n->passing_seq[0] = 1;
        if (i == 0) {
            push_pre(g, n, POA_SENTINEL, 0);
            n->in = 0;                 /* a head: pre[0] is the sentinel, weight 0 */
            n->node_base_len = 1;
        } else {
            push_pre(g, n, id - 1, 1);
        }
        if (i + 1 < len) push_next(g, n, id + 1);
        n->sub = i;
        g->unsort[i] = g->sort[i] = id;
    }
    g->len = len;
    g->last_node_num = 1;
}

void poa_graph_free(poa_graph* g)
{
    for (void* b = g->pool_blocks; b;) { void* nx = *(void**)b; free(b); b = nx; }
    free(g->node); free(g->unsort); free(g->sort);
    memset(g, 0, sizeof *g);
}

/* ------------------------------------------------------------------------------------------
 * CSR export in row order
 * ---------------------------------------------------------------------------------------- */
void poa_graph_export(const poa_graph* g, int E, int O, poa_csr* c)
{
    memset(c, 0, sizeof *c);
    const int R = g->len;
    c->nrows = R;
    c->row_ptr = (int32_t*)malloc((size_t)(R + 1) * sizeof(int32_t));
    c->base = (uint8_t*)malloc((size_t)R + 1);
    c->is_sink = (uint8_t*)malloc((size_t)R + 1);
    c->hneg1 = (int32_t*)malloc((size_t)(R + 1) * sizeof(int32_t));
    int ne = 0;
    for (int r = 0; r < R; r++) {
        const poa_node* n = &g->node[g->sort[r]];
        ne += n->in > 0 ? n->in : 1;
    }
    c->nedges = ne;
    c->pred_row = (int32_t*)malloc((size_t)(ne > 0 ? ne : 1) * sizeof(int32_t));
    int e = 0;
    for (int r = 0; r < R; r++) {
        const poa_node* n = &g->node[g->sort[r]];
        c->row_ptr[r] = e;
        c->base[r] = n->base;
        c->is_sink[r] = n->out == 0;
        if (n->in == 0) {
            c->pred_row[e++] = POA_SENTINEL;
            c->hneg1[r] = O + E;                               /* c-t-simd.c:187 */
        } else {
            int best = INT_MIN;
            for (int j = 0; j < n->in; j++) {
                const int pr = g->node[n->pre[j]].sub;
                c->pred_row[e++] = pr;
                if (c->hneg1[pr] + E > best) best = c->hneg1[pr] + E;   /* c-t-simd.c:196-199 */
            }
            c->hneg1[r] = best;
        }
    }
    c->row_ptr[R] = e;
}

void poa_csr_free(poa_csr* c)
{
    free(c->row_ptr); free(c->pred_row); free(c->base); free(c->is_sink); free(c->hneg1);
    memset(c, 0, sizeof *c);
}

/* ------------------------------------------------------------------------------------------
 * delta export: node records that changed since the device last saw them
 * ---------------------------------------------------------------------------------------- */
int poa_graph_delta(const poa_graph* g, poa_shadow* sh)
{
    if (sh->cap < g->nnode) {
        int cap = sh->cap ? sh->cap : 1024;
        while (cap < g->nnode) cap *= 2;
        sh->meta = (uint32_t*)realloc(sh->meta, (size_t)cap * sizeof(uint32_t));
        sh->pre = (int32_t*)realloc(sh->pre, (size_t)cap * 8 * sizeof(int32_t));
        memset(sh->meta + sh->cap, 0xff, (size_t)(cap - sh->cap) * sizeof(uint32_t));
        sh->cap = cap;
    }
    sh->n2 = sh->n8 = 0;
    sh->nedges = 0;
    for (int r = 0; r < g->len; r++) {
        const int id = g->sort[r];
        const poa_node* n = &g->node[id];
        sh->nedges += n->in > 0 ? n->in : 1;
        if (n->in > 8) return -1;
        const uint32_t meta = (uint32_t)n->base | ((uint32_t)n->in << 8);
        int32_t* sp = sh->pre + (size_t)id * 8;
        int same = sh->meta[id] == meta;
        for (int j = 0; same && j < n->in; j++) same = sp[j] == n->pre[j];
        if (same) continue;
        sh->meta[id] = meta;
        for (int j = 0; j < n->in; j++) sp[j] = n->pre[j];
        if (n->in <= 2) {
            if (sh->n2 == sh->cap2) {
                sh->cap2 = sh->cap2 ? sh->cap2 * 2 : 4096;
                sh->r2 = (poa_rec2*)realloc(sh->r2, (size_t)sh->cap2 * sizeof(poa_rec2));
            }
            poa_rec2* q = &sh->r2[sh->n2++];
            q->id = id; q->base_in = meta;
            q->pre[0] = n->in > 0 ? n->pre[0] : -1;
            q->pre[1] = n->in > 1 ? n->pre[1] : -1;
        } else {
            if (sh->n8 == sh->cap8) {
                sh->cap8 = sh->cap8 ? sh->cap8 * 2 : 1024;
                sh->r8 = (poa_rec8*)realloc(sh->r8, (size_t)sh->cap8 * sizeof(poa_rec8));
            }
            poa_rec8* q = &sh->r8[sh->n8++];
            q->id = id; q->base_in = meta;
            for (int j = 0; j < 8; j++) q->pre[j] = j < n->in ? n->pre[j] : -1;
        }
    }
    return 0;
}

void poa_shadow_free(poa_shadow* sh)
{
    free(sh->meta); free(sh->pre); free(sh->r2); free(sh->r8);
    memset(sh, 0, sizeof *sh);
}

int poa_graph_pick_start(const poa_graph* g, const int32_t* lastsorce)
{
    int num1 = g->len - 1, best = INT_MIN, seen = 0;
    for (int i = g->len - 1; i > 0; i--) {                     /* c-t-simd.c:737-750 */
        if (g->node[g->sort[i]].out == 0) {
            if (best <= lastsorce[i]) { best = lastsorce[i]; num1 = i; }
            seen++;
        }
        if (seen >= g->last_node_num) break;
    }
    return num1;
}

/* ------------------------------------------------------------------------------------------
 * fuse one read along its traceback path (c-t-simd.c:659-977)
 * ---------------------------------------------------------------------------------------- */
static void realize(poa_graph* g, int id, int num, int* l, int set_sub)
{
    poa_node* n = &g->node[id];
    n->passing_seq = passing_new(g);
    n->passing_seq[num] = 1;
    n->ew[0] = 1;
    order_reserve(g, g->len + *l + 1);
    g->unsort[g->len + *l] = id;
    if (set_sub) n->sub = g->len + *l;
    (*l)++;
}

int poa_graph_fuse(poa_graph* g, const char* b, int num, const poa_step* path, int nsteps)
{
    const int len_b = (int)strlen(b);
    if (len_b == 0) return 0;
    int overflow = 0;
    /* one provisional node per read base, chained (c-t-simd.c:685-726) */
    int* seq = (int*)malloc((size_t)len_b * sizeof(int));
    for (int i = 0; i < len_b; i++) seq[i] = node_new(g, (unsigned char)b[i]);
    for (int i = 0; i < len_b; i++) {
        poa_node* n = &g->node[seq[i]];
        if (i == 0) { push_pre(g, n, POA_SENTINEL, 0); n->in = 0; n->node_base_len = 1; }
        else push_pre(g, n, seq[i - 1], 0);
        if (i + 1 < len_b) push_next(g, n, seq[i + 1]);
    }
    order_reserve(g, g->len + len_b);

    int l = 0, cont = 0, s2 = 1;
    int num2 = len_b - 1;
    int ended_at_sentinel = 0;
    for (int k = 0; k < nsteps; k++) {
        const poa_step st = path[k];
        num2 = st.col;
        const int cur = g->sort[st.row];
        const int next_is_match = (k + 1 < nsteps) && path[k + 1].kind == POA_STEP_MATCH;
        if (st.kind == POA_STEP_HORIZ) {                                   /* 753-772 */
            cont = 0;
            realize(g, seq[num2], num, &l, 1);
            num2--;
            continue;
        }
        if (st.kind == POA_STEP_VERT) {                                    /* 773-785 */
            cont = 3;
            continue;
        }
        const int P = g->node[cur].pre[st.pred];
        const int P_sub = pre_sub(g, P);
        if (st.kind == POA_STEP_MATCH) {                                   /* 786-861 */
            poa_node* c = &g->node[cur];
            if (num2 == len_b - 1) {
                if (next_is_match) {
                    c->ew[st.pred]++;
                    node_kill(g, seq[num2]);
                } else {
                    if (num2 >= 1) {
                        push_pre(g, c, seq[num2 - 1], 1);
                        poa_node* q = &g->node[seq[num2 - 1]];
                        q->next[q->out - 1] = cur;
                    }
                    node_kill(g, seq[num2]);
                    seq[num2] = cur;
                }
            } else if (num2 == 0) {
                if (cont == 1 || cont == 5) {
                    node_kill(g, seq[0]);
                    seq[0] = cur;
                } else {
                    push_next(g, c, seq[1]);
                    poa_node* q = &g->node[seq[1]];
                    q->pre[(q->in > 0 ? q->in : 1) - 1] = cur;
                    node_kill(g, seq[0]);          /* the reference leaks it; it is in no list */
                    seq[0] = cur;
                }
            } else {
                if (P_sub != -1 && next_is_match) {
                    c->ew[st.pred]++;
                } else {
                    push_pre(g, c, seq[num2 - 1], 1);
                    poa_node* q = &g->node[seq[num2 - 1]];
                    q->next[q->out - 1] = cur;
                }
                if (!(cont == 1 || cont == 5)) {
                    c = &g->node[cur];
                    push_next(g, c, seq[num2 + 1]);
                    poa_node* q = &g->node[seq[num2 + 1]];
                    q->pre[(q->in > 0 ? q->in : 1) - 1] = cur;
                }
                node_kill(g, seq[num2]);
                seq[num2] = cur;
            }
            cont = 1;
            g->node[cur].passing_seq[num] = 1;
            num2--;
            if (P_sub == -1) ended_at_sentinel = 1;
            continue;
        }
        /* POA_STEP_MISMATCH                                                862-956 */
        int s4 = 0;
        for (int s = 0; s < g->node[cur].mm_num; s++) {
            const int MN = g->node[cur].mm[s];
            if (g->node[seq[num2]].base != g->node[MN].base) continue;
            if (num2 != 0) {
                if (P_sub != -1 && next_is_match) {
                    poa_node* mn = &g->node[MN];
                    for (int ss = 0; ss < mn->in; ss++)
                        if (mn->pre[ss] == P) { mn->ew[ss]++; s2 = -1; }
                }
                if (s2 != -1) {
                    push_pre(g, &g->node[MN], seq[num2 - 1], 1);
                    poa_node* q = &g->node[seq[num2 - 1]];
                    q->next[q->out - 1] = MN;
                }
            }
            s4 = 1;
            if ((cont == 1 || cont == 4) && num2 + 1 < len_b) {
                poa_node* q = &g->node[seq[num2 + 1]];
                const int qin = q->in;
                for (int ss = 0; ss < qin; ss++)
                    if (q->pre[ss] == MN) {
                        s4 = 2;
                        q->ew[ss]++;
                        q->in--;               /* drops the provisional link appended last */
                    }
            }
            if (s4 == 1 && num2 != len_b - 1) {
                poa_node* q = &g->node[seq[num2 + 1]];
                q->pre[(q->in > 0 ? q->in : 1) - 1] = MN;
                push_next(g, &g->node[MN], seq[num2 + 1]);
            }
            g->node[MN].passing_seq[num] = 1;
            cont = (s2 == -1) ? 5 : 4;
            s2 = 0;
            node_kill(g, seq[num2]);
            seq[num2] = MN;
        }
        if (s4 == 0) {
            cont = 2;
            realize(g, seq[num2], num, &l, 1);
            poa_node* c = &g->node[cur];
            poa_node* nn = &g->node[seq[num2]];
            /* The reference keeps the group in mismatch_node[4] (msa/poa.h:100-134) and writes past it when a
             * sixth distinct symbol lands in one column (soft-masked FASTA, IUPAC codes, proteins); here the
             * read is refused instead of corrupting the graph. */
            if (c->mm_num >= POA_MM_MAX) { overflow = 1; break; }
            c->mm_num++;
            c->mm[c->mm_num - 1] = seq[num2];
            nn->mm_num = c->mm_num;
            nn->mm[nn->mm_num - 1] = cur;
            for (int s = 0; s < c->mm_num - 1; s++) {
                poa_node* o = &g->node[c->mm[s]];
                o->mm_num++;
                o->mm[c->mm_num - 1] = seq[num2];
                nn->mm[s] = c->mm[s];
            }
        }
        num2--;
        if (P_sub == -1) ended_at_sentinel = 1;
    }
    (void)ended_at_sentinel;
    if (overflow) { free(seq); return -1; }     /* (the graph is left half fused: the caller stops) */
    /* read bases the path never reached become a new head chain (959-972) */
    while (num2 > -1) {
        realize(g, seq[num2], num, &l, 0);
        num2--;
    }
    g->len += l;
    free(seq);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * topological orders (msa/topo.c).  The reference recurses once per node along chains, so
 * the sorts run on a thread with a large stack.
 * ---------------------------------------------------------------------------------------- */
static int tp(poa_graph* g, int id, int subs)                              /* topo.c:179-203 */
{
    poa_node* p = &g->node[id];
    g->sort[subs] = id;
    p->node_logo = 0;
    p->sub = subs;
    p->in_temp = -1;
    subs++;
    for (int j = 0; j < p->out; j++) {
        poa_node* q = &g->node[p->next[j]];
        if (q->out == 0 && q->passing == 1 && q->in_temp - 1 == 0) {
            q->in_temp--;
            if (q->in_temp == 0) subs = tp(g, p->next[j], subs);
            p = &g->node[id];
        }
    }
    for (int j = 0; j < p->out; j++) {
        poa_node* q = &g->node[p->next[j]];
        q->in_temp--;
        if (q->in_temp == 0 && q->passing != 2) subs = tp(g, p->next[j], subs);
        p = &g->node[id];
    }
    return subs;
}

static void mark_passing(poa_graph* g, int with_one)
{
    for (int i = 0; i < g->len; i++) {
        poa_node* u = &g->node[g->unsort[i]];
        u->in_temp = u->in;
        u->passing = 0;
        if (u->out == 0 && u->mm_num > 0) {
            int any = 0;
            for (int j = 0; j < u->mm_num; j++)
                if (g->node[u->mm[j]].out != 0) { any = 1; if (with_one) u->passing = 1; }
            if (!any) u->passing = 2;
        }
    }
}

static void toposort(poa_graph* g)                                         /* topo.c:205-241 */
{
    mark_passing(g, 1);
    int subs = 0;
    while (subs < g->len) {
        int progressed = 0;
        for (int i = 0; i < g->len; i++)
            if (g->node[g->unsort[i]].in_temp == 0) {
                subs = tp(g, g->unsort[i], subs);
                if (subs + g->last_node_num == g->len)
                    for (int k = 0; k < g->len; k++)
                        if (g->node[g->unsort[k]].in_temp == 0) subs = tp(g, g->unsort[k], subs);
                progressed = 1;
                break;
            }
        if (!progressed) break;   /* malformed graph: the reference would spin forever */
    }
}

static void best_pre(const poa_graph* g, const poa_node* n, int* max_out, int* max_i_out)
{
    int max = 0, max_i = 0;                                                /* topo.c:17-32,151-166 */
    for (int i = 0; i < n->in; i++) {
        if (pre_sorce(g, n->pre[i]) >= 0) {
            if (max < n->ew[i]) { max = n->ew[i]; max_i = i; }
            else if (max == n->ew[i] && pre_sorce(g, n->pre[max_i]) <= pre_sorce(g, n->pre[i])) { max = n->ew[i]; max_i = i; }
        }
    }
    *max_out = max;
    *max_i_out = max_i;
}

static int group_ready(const poa_graph* g, const poa_node* q)
{
    int c = 0;
    for (int s = 0; s < q->mm_num; s++)
        if (g->node[q->mm[s]].in_temp == 0) c++;
    return c == q->mm_num;
}

static int tp1(poa_graph* g, int id, int subs)                             /* topo.c:7-64 */
{
    poa_node* p = &g->node[id];
    g->sort[subs] = id;
    p->node_logo = 0;
    p->sub = subs;
    int max, max_i;
    best_pre(g, p, &max, &max_i);
    p->node_sorce = pre_sorce(g, p->pre[max_i]) + max;
    p->node_base_len = pre_len(g, p->pre[max_i]) + 1;
    p->node_sorce_source = pre_sub(g, p->pre[max_i]);
    p->in_temp = -1;
    subs++;
    for (int j = 0; j < p->out; j++) {
        const int qid = p->next[j];
        poa_node* q = &g->node[qid];
        q->in_temp--;
        if (q->in_temp == 0 && q->mm_num == 0 && q->passing != 2) {
            subs = tp1(g, qid, subs);
        } else if (q->in_temp == 0 && q->mm_num > 0 && q->passing != 2) {
            if (group_ready(g, q)) {
                subs = tp1(g, qid, subs);
                q = &g->node[qid];
                for (int ss = 0; ss < q->mm_num; ss++)
                    if (g->node[q->mm[ss]].in_temp == 0) subs = tp1(g, q->mm[ss], subs);
            }
        }
        p = &g->node[id];
    }
    return subs;
}

static void toposort1(poa_graph* g)                                        /* topo.c:66-117 */
{
    mark_passing(g, 0);
    int subs = 0;
    while (subs < g->len) {
        int progressed = 0;
        for (int i = 0; i < g->len; i++) {
            const int uid = g->unsort[i];
            poa_node* u = &g->node[uid];
            if (u->in_temp != 0) continue;
            if (u->mm_num == 0) {
                subs = tp1(g, uid, subs);
                progressed = 1;
                break;
            }
            if (group_ready(g, u)) {
                subs = tp1(g, uid, subs);
                u = &g->node[uid];
                for (int ss = 0; ss < u->mm_num; ss++)
                    if (g->node[u->mm[ss]].in_temp == 0) subs = tp1(g, u->mm[ss], subs);
                progressed = 1;
                break;
            }
        }
        if (!progressed) break;
    }
}

static void modify(poa_graph* g)                                           /* topo.c:119-177 */
{
    for (;;) {
        int max = INT_MIN, max_i = 0;
        for (int i = 0; i < g->len; i++)
            if (max <= g->node[g->sort[i]].node_sorce) { max = g->node[g->sort[i]].node_sorce; max_i = i; }
        poa_node* M = &g->node[g->sort[max_i]];
        if (M->out == 0) return;
        for (int i = 0; i < M->out; i++) {
            poa_node* nx = &g->node[M->next[i]];
            for (int j = 0; j < nx->in; j++) {
                poa_node* pj = &g->node[nx->pre[j]];
                if (pj->node_sorce < M->node_sorce && pj->node_sorce > 0) pj->node_sorce = -pj->node_sorce;
            }
            nx->node_logo = 4;
        }
        for (int i = max_i + 1; i < g->len; i++) {
            poa_node* n = &g->node[g->sort[i]];
            if (n->node_sorce >= 0 || n->node_logo == 4) {
                int Max, Max_i;
                best_pre(g, n, &Max, &Max_i);
                n->node_sorce = pre_sorce(g, n->pre[Max_i]) + Max;
                n->node_base_len = pre_len(g, n->pre[Max_i]) + 1;
                n->node_sorce_source = pre_sub(g, n->pre[Max_i]);
                n->node_logo = 0;
            }
        }
    }
}

typedef struct { poa_graph* g; int final; } sort_job;

static void* sort_thread(void* arg)
{
    sort_job* j = (sort_job*)arg;
    poa_graph* g = j->g;
    g->last_node_num = 0;                                                  /* topo.c:243-260 */
    for (int i = 0; i < g->len; i++)
        if (g->node[g->unsort[i]].out == 0) g->last_node_num++;
    if (!j->final) {
        toposort(g);
    } else {
        toposort1(g);
        modify(g);
    }
    for (int i = 0; i < g->len; i++) g->unsort[i] = g->sort[i];
    return NULL;
}

/* The sorts recurse as deep as the graph is long, so they run on a thread with a 1 GiB stack.  The thread is created once
 * and kept: a fresh one per call costs ~2.7 ms (mmap / munmap of the stack and its page faults all over again), as
 * much as sorting 70 000 nodes. */
static pthread_mutex_t g_sort_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_sort_go = PTHREAD_COND_INITIALIZER, g_sort_done = PTHREAD_COND_INITIALIZER;
static sort_job* g_sort_job;
static int g_sort_worker;        /* 0: not started, 1: running, -1: could not be created (sort on the caller's stack) */

static void* sort_worker(void* arg)
{
    (void)arg;
    pthread_mutex_lock(&g_sort_mu);
    for (;;) {
        while (g_sort_job == NULL) pthread_cond_wait(&g_sort_go, &g_sort_mu);
        sort_job* job = g_sort_job;
        pthread_mutex_unlock(&g_sort_mu);
        sort_thread(job);
        pthread_mutex_lock(&g_sort_mu);
        g_sort_job = NULL;
        pthread_cond_signal(&g_sort_done);
    }
    return NULL;
}

void poa_graph_sort(poa_graph* g, int final)
{
    sort_job job = {g, final};
    pthread_mutex_lock(&g_sort_mu);
    if (g_sort_worker == 0) {
        pthread_attr_t at;
        pthread_attr_init(&at);
        pthread_attr_setstacksize(&at, (size_t)1 << 30);   /* recursion depth ~ number of nodes */
        pthread_attr_setdetachstate(&at, PTHREAD_CREATE_DETACHED);
        pthread_t th;
        g_sort_worker = pthread_create(&th, &at, sort_worker, NULL) == 0 ? 1 : -1;
        pthread_attr_destroy(&at);
    }
    if (g_sort_worker < 0) {
        pthread_mutex_unlock(&g_sort_mu);
        sort_thread(&job);
        return;
    }
    g_sort_job = &job;
    pthread_cond_signal(&g_sort_go);
    while (g_sort_job != NULL) pthread_cond_wait(&g_sort_done, &g_sort_mu);
    pthread_mutex_unlock(&g_sort_mu);
}

/* ------------------------------------------------------------------------------------------
 * output (msa/result.c)
 * ---------------------------------------------------------------------------------------- */
void poa_graph_write(poa_graph* g, FILE* res)
{
    const int num = g->nseq;
    char** r = (char**)malloc((size_t)num * sizeof(char*));
    for (int i = 0; i < num; i++) {
        r[i] = (char*)malloc((size_t)g->len + 1);
        memset(r[i], '-', (size_t)g->len);
    }
    int col = 0;
    for (int i = 0; i < g->len; i++) {                                     /* result.c:52-66 */
        poa_node* n = &g->node[g->sort[i]];
        if (n->node_logo != 0) continue;
        for (int j = 0; j < num; j++)
            if (n->passing_seq[j] == 1) r[j][col] = (char)n->base;
        for (int l = 0; l < n->mm_num; l++) {
            poa_node* o = &g->node[n->mm[l]];
            for (int j = 0; j < num; j++)
                if (o->passing_seq[j] == 1) { r[j][col] = (char)o->base; o->node_logo = 1; }
        }
        col++;
    }
    for (int i = 0; i < num; i++) {
        r[i][col] = '\0';
        fprintf(res, ">%d\n", i + 1);
        fputs(r[i], res);
        fputs("\n", res);
        free(r[i]);
    }
    free(r);
    /* consensus: heaviest path ending in the best of the last sinks (result.c:6-40) */
    if (g->len == 0) { fprintf(res, "consensus:\n"); return; }
    int s = 0;
    int max = g->node[g->sort[g->len - 1]].node_sorce, max_i = g->len - 1;
    for (int i = g->len - 2; i >= 0; i--) {
        poa_node* n = &g->node[g->sort[i]];
        if (n->out == 0) {
            if (max <= n->node_sorce) { max = n->node_sorce; max_i = i; }
            s++;
        }
        if (s >= g->last_node_num) break;
    }
    int len = g->node[g->sort[max_i]].node_base_len;
    char* c = (char*)malloc((size_t)len + 1);
    c[len] = '\0';
    for (int k = len - 1; k > -1; k--) {
        poa_node* n = &g->node[g->sort[max_i]];
        c[k] = (char)n->base;
        max_i = n->node_sorce_source;
        if (max_i < 0 && k > 0) { memmove(c, c + k, (size_t)(len - k) + 1); break; }
    }
    fprintf(res, "consensus:\n");
    fputs(c, res);
    free(c);
}
