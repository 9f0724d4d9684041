/*
 * vdj_oracle.c -- sequential CPU restatement of V'DJer's de Bruijn graph build.
 *
 * TEST INFRASTRUCTURE ONLY (see vdj_oracle.h).  This file restates, in plain C and in the
 * reference's own record-by-record order, what these reference functions compute
 * (all in /root/reference/src/main/c/assembler2_vdj.c):
 *
 *   include_kmer          :240-259   window gate (no 'N', every phred >= 20)
 *   add_to_table          :322-367   pass 1: count / multi-read flag / per-position quality sums
 *   build_pre_graph       :369-409   record loop over one buffer
 *   is_base_quality_good  :454-465   \  prune
 *   prune_pre_graph       :467-484   /
 *   new_node              :190-204   \
 *   increment_node_freq   :261-265    | pass 2: nodes in creation order, saturating frequency,
 *   is_node_in_list       :210-221    | head-inserted toNodes / fromNodes lists
 *   link_nodes            :223-237    |
 *   add_to_graph          :267-320    |
 *   build_graph2          :412-452   /
 *   assemble() block      :1381-1415 orchestration: primary then secondary buffer, both passes
 *
 * The reference stores its tables in Google sparsehash keyed by a pointer to the first window
 * that created the entry, hashing/comparing kmer_size characters (hash_utils.h:13-28).  Only the
 * *content* of the tables is restated here (a private chained hash map with the same key
 * semantics); sparsehash iteration order is not observable in any output of this stage.
 *
 * Parity pin: compared bit-for-bit against the compiled reference (oracle/_ref/libvdjref.so,
 * built by oracle/Makefile from the sources under /root/reference) in tests/test_oracle_vs_ref.py
 * and against the committed reference outputs under tests/golden/.
 */
#define _POSIX_C_SOURCE 200809L
#include "vdj_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <time.h>

/* constants, assembler2_vdj.c:66-76 */
#define REF_MAX_FREQUENCY 32766
#define REF_MAX_QUAL_SUM 255
#define REF_MIN_BASE_QUALITY 20

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* phred33, assembler2_vdj.c:150-152 (unsigned char arithmetic, wraps like the reference) */
static unsigned char phred(char c) { return (unsigned char)(c - '!'); }

/* ------------------------------------------------------------------------------------------
 * private map: key = pointer to k characters; equality = strncmp over k (hash_utils.h:13-20)
 * ---------------------------------------------------------------------------------------- */
typedef struct pre_entry { /* struct pre_node, assembler2_vdj.c:127-133 */
    const char *key;       /* map key: first gated window of this k-mer */
    const char *first_read;/* contributingRead */
    unsigned char qsum[VDJ_ORACLE_MAX_KMER];
    unsigned short freq;
    char multi;            /* hasMultipleUniqueReads */
    char strand;           /* contributing_strand */
    char alive;
    uint32_t next;         /* chain */
} pre_entry;

typedef struct graph_node { /* struct node, assembler2_vdj.c:109-125 (fields this stage sets) */
    const char *kmer;
    uint32_t id;            /* creation rank from 1 */
    unsigned short freq;
    uint32_t to_head, from_head; /* index into cells, NIL = none */
    uint32_t next;          /* chain */
} graph_node;

typedef struct list_cell { uint32_t node, next; } list_cell; /* struct linked_node :135-138 */

#define NIL 0xFFFFFFFFu

static uint64_t hash_k(const char *s, int k) {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < k; i++) { h ^= (unsigned char)s[i]; h *= 0x100000001B3ull; h ^= h >> 29; }
    return h;
}

typedef struct ctx {
    int L, k, w;
    /* pre table */
    pre_entry *pre; uint64_t n_pre, cap_pre;
    uint32_t *pre_buckets; uint64_t n_pre_buckets;
    /* nodes */
    graph_node *nodes; uint64_t n_nodes, cap_nodes;
    uint32_t *node_buckets; uint64_t n_node_buckets;
    list_cell *cells; uint64_t n_cells, cap_cells;
    uint64_t n_gated, n_hits;
} ctx;

static void pre_rehash(ctx *c) {
    uint64_t nb = c->n_pre_buckets ? c->n_pre_buckets * 2 : 1024;
    free(c->pre_buckets);
    c->pre_buckets = (uint32_t *)malloc(nb * sizeof(uint32_t));
    memset(c->pre_buckets, 0xFF, nb * sizeof(uint32_t));
    c->n_pre_buckets = nb;
    for (uint64_t i = 0; i < c->n_pre; i++) {
        uint64_t b = hash_k(c->pre[i].key, c->k) & (nb - 1);
        c->pre[i].next = c->pre_buckets[b];
        c->pre_buckets[b] = (uint32_t)i;
    }
}

static pre_entry *pre_find(ctx *c, const char *kmer) {
    if (!c->n_pre_buckets) return NULL;
    uint64_t b = hash_k(kmer, c->k) & (c->n_pre_buckets - 1);
    for (uint32_t i = c->pre_buckets[b]; i != NIL; i = c->pre[i].next)
        if (strncmp(c->pre[i].key, kmer, (size_t)c->k) == 0) return &c->pre[i];
    return NULL;
}

static pre_entry *pre_insert(ctx *c, const char *kmer) {
    if (c->n_pre == c->cap_pre) {
        c->cap_pre = c->cap_pre ? c->cap_pre * 2 : 4096;
        c->pre = (pre_entry *)realloc(c->pre, c->cap_pre * sizeof(pre_entry));
    }
    if (c->n_pre >= c->n_pre_buckets) pre_rehash(c);
    pre_entry *e = &c->pre[c->n_pre];
    memset(e, 0, sizeof(*e));
    e->key = kmer;
    e->alive = 1;
    uint64_t b = hash_k(kmer, c->k) & (c->n_pre_buckets - 1);
    e->next = c->pre_buckets[b];
    c->pre_buckets[b] = (uint32_t)c->n_pre;
    c->n_pre++;
    return e;
}

static void node_rehash(ctx *c) {
    uint64_t nb = c->n_node_buckets ? c->n_node_buckets * 2 : 1024;
    free(c->node_buckets);
    c->node_buckets = (uint32_t *)malloc(nb * sizeof(uint32_t));
    memset(c->node_buckets, 0xFF, nb * sizeof(uint32_t));
    c->n_node_buckets = nb;
    for (uint64_t i = 0; i < c->n_nodes; i++) {
        uint64_t b = hash_k(c->nodes[i].kmer, c->k) & (nb - 1);
        c->nodes[i].next = c->node_buckets[b];
        c->node_buckets[b] = (uint32_t)i;
    }
}

static uint32_t node_find(ctx *c, const char *kmer) {
    if (!c->n_node_buckets) return NIL;
    uint64_t b = hash_k(kmer, c->k) & (c->n_node_buckets - 1);
    for (uint32_t i = c->node_buckets[b]; i != NIL; i = c->nodes[i].next)
        if (strncmp(c->nodes[i].kmer, kmer, (size_t)c->k) == 0) return i;
    return NIL;
}

/* new_node, :190-204: pool bump, frequency 1, id = running counter from 1 */
static uint32_t node_create(ctx *c, const char *kmer) {
    if (c->n_nodes == c->cap_nodes) {
        c->cap_nodes = c->cap_nodes ? c->cap_nodes * 2 : 4096;
        c->nodes = (graph_node *)realloc(c->nodes, c->cap_nodes * sizeof(graph_node));
    }
    if (c->n_nodes >= c->n_node_buckets) node_rehash(c);
    uint32_t idx = (uint32_t)c->n_nodes;
    graph_node *n = &c->nodes[idx];
    n->kmer = kmer;
    n->freq = 1;
    n->id = idx + 1;
    n->to_head = n->from_head = NIL;
    uint64_t b = hash_k(kmer, c->k) & (c->n_node_buckets - 1);
    n->next = c->node_buckets[b];
    c->node_buckets[b] = idx;
    c->n_nodes++;
    return idx;
}

/* is_node_in_list, :210-221: membership by k-mer content */
static int list_has(ctx *c, uint32_t head, uint32_t node) {
    for (uint32_t p = head; p != NIL; p = c->cells[p].next) {
        const char *a = c->nodes[c->cells[p].node].kmer, *b = c->nodes[node].kmer;
        if (a == b || strncmp(a, b, (size_t)c->k) == 0) return 1;
    }
    return 0;
}

static uint32_t cell_new(ctx *c, uint32_t node, uint32_t next) {
    if (c->n_cells == c->cap_cells) {
        c->cap_cells = c->cap_cells ? c->cap_cells * 2 : 8192;
        c->cells = (list_cell *)realloc(c->cells, c->cap_cells * sizeof(list_cell));
    }
    c->cells[c->n_cells].node = node;
    c->cells[c->n_cells].next = next;
    return (uint32_t)c->n_cells++;
}

/* link_nodes, :223-237: unseen successor / predecessor goes to the HEAD of the list */
static void link(ctx *c, uint32_t from, uint32_t to) {
    if (!list_has(c, c->nodes[from].to_head, to))
        c->nodes[from].to_head = cell_new(c, to, c->nodes[from].to_head);
    if (!list_has(c, c->nodes[to].from_head, from))
        c->nodes[to].from_head = cell_new(c, from, c->nodes[to].from_head);
}

/* include_kmer, :240-259 */
static int window_passes_gate(const char *seq, const char *qual, int idx, int k) {
    for (int i = idx; i < idx + k; i++) {
        if (seq[i] == 'N') return 0;
        if (phred(qual[i]) < REF_MIN_BASE_QUALITY) return 0;
    }
    return 1;
}

/* add_to_table, :322-367 */
static void pass1_record(ctx *c, const char *seq, const char *qual, int strand) {
    const int k = c->k;
    for (int i = 0; i <= c->L - k; i++) {
        if (!window_passes_gate(seq, qual, i, k)) continue;
        c->n_gated++;
        const char *kmer = seq + i;
        const char *kq = qual + i;
        pre_entry *e = pre_find(c, kmer);
        if (!e) {
            e = pre_insert(c, kmer);
            e->first_read = seq;
            e->freq = 1;
            e->multi = 0;
            e->strand = (char)strand;
            /* :337-339 -- seeded from the RECORD's first k qualities (qual[j]), not the
             * window's (kq[j]); the reference's inner loop index shadows i.  Parity keeps it. */
            for (int j = 0; j < k; j++) e->qsum[j] = phred(qual[j]);
        } else {
            if (e->freq < REF_MAX_FREQUENCY - 1) e->freq++; /* :345-347 */
            /* :349-352, compare_read :142-144 = strncmp over read_length */
            if (!e->multi) {
                int same = (e->first_read == seq) || strncmp(e->first_read, seq, (size_t)c->L) == 0;
                if (!same || e->strand != (char)strand) e->multi = 1;
            }
            /* :354-361 */
            for (int j = 0; j < k; j++) {
                unsigned char q = phred(kq[j]);
                if ((int)e->qsum[j] + (int)q < REF_MAX_QUAL_SUM - 41) e->qsum[j] = (unsigned char)(e->qsum[j] + q);
                else e->qsum[j] = REF_MAX_QUAL_SUM;
            }
        }
    }
}

/* add_to_graph, :267-320 */
static void pass2_record(ctx *c, const char *seq) {
    uint32_t prev = NIL;
    for (int i = 0; i <= c->L - c->k; i++) {
        const char *kmer = seq + i;
        pre_entry *e = pre_find(c, kmer);
        if (e && e->alive) {
            c->n_hits++;
            uint32_t cur = node_find(c, kmer);
            if (cur == NIL) cur = node_create(c, kmer);
            else if (c->nodes[cur].freq < REF_MAX_FREQUENCY - 1) c->nodes[cur].freq++; /* :261-265 */
            if (prev != NIL) link(c, prev, cur);
            prev = cur;
        } else {
            prev = NIL;
        }
    }
}

/* record loops of build_pre_graph :369-409 / build_graph2 :412-452 */
static int scan_buffer(ctx *c, const char *buf, int pass) {
    size_t len = strlen(buf);
    size_t rec_len = (size_t)c->L * 2 + 1;
    size_t n = len / rec_len;
    for (size_t r = 0; r < n; r++) {
        const char *p = buf + r * rec_len;
        int strand;
        if (p[0] == '0') strand = 0;
        else if (p[0] == '1') strand = 1;
        else return -2; /* reference: exit(-1), :388-390 */
        if (pass == 1) pass1_record(c, p + 1, p + 1 + c->L, strand);
        else pass2_record(c, p + 1);
    }
    return 0;
}

typedef struct bufmap { const char *p, *s; size_t np, ns; int L, w; } bufmap;

static uint64_t stamp_of(const bufmap *m, const char *kmer) {
    size_t rec_len = (size_t)m->L * 2 + 1;
    const char *base; uint64_t rec0;
    if (kmer >= m->p && kmer < m->p + m->np * rec_len) { base = m->p; rec0 = 0; }
    else { base = m->s; rec0 = m->np; }
    size_t off = (size_t)(kmer - base);
    return (rec0 + off / rec_len) * (uint64_t)m->w + (off % rec_len - 1);
}

static int cmp_u64_pair(const void *a, const void *b) {
    uint64_t x = ((const uint64_t *)a)[0], y = ((const uint64_t *)b)[0];
    return x < y ? -1 : x > y;
}

int vdj_oracle_build(const char *primary, const char *secondary, int read_length, int kmer_size,
                     int min_node_freq, int min_base_quality, vdj_oracle_result *out) {
    memset(out, 0, sizeof(*out));
    if (!primary || !secondary) return -1;
    if (kmer_size < 1 || kmer_size > VDJ_ORACLE_MAX_KMER || kmer_size > read_length) return -1;
    ctx c;
    memset(&c, 0, sizeof(c));
    c.L = read_length; c.k = kmer_size; c.w = read_length - kmer_size + 1;
    bufmap m = { primary, secondary, strlen(primary) / ((size_t)read_length * 2 + 1),
                 strlen(secondary) / ((size_t)read_length * 2 + 1), read_length, c.w };
    int rc;

    /* pass 1: primary then secondary, :1388-1390 */
    double t0 = now_s();
    if ((rc = scan_buffer(&c, primary, 1)) || (rc = scan_buffer(&c, secondary, 1))) goto fail;
    double t1 = now_s();
    out->n_pre_total = c.n_pre;

    /* prune_pre_graph :467-484 with is_base_quality_good :454-465.
     * main() clamps --mq to <= 254 before anything runs (:1514-1516). */
    int mq = min_base_quality > REF_MAX_QUAL_SUM - 1 ? REF_MAX_QUAL_SUM - 1 : min_base_quality;
    uint64_t n_keep = 0;
    for (uint64_t i = 0; i < c.n_pre; i++) {
        pre_entry *e = &c.pre[i];
        int good = 1;
        for (int j = 0; j < c.k; j++) if ((int)e->qsum[j] < mq) { good = 0; break; }
        if ((int)e->freq < min_node_freq || !e->multi || !good) e->alive = 0;
        else n_keep++;
    }
    double t2 = now_s();

    /* pass 2: primary then secondary, :1402-1408 */
    if ((rc = scan_buffer(&c, primary, 2)) || (rc = scan_buffer(&c, secondary, 2))) goto fail;
    double t3 = now_s();

    out->t_pass1 = t1 - t0; out->t_prune = t2 - t1; out->t_pass2 = t3 - t2;
    out->n_records = m.np + m.ns;
    out->n_windows = out->n_records * (uint64_t)c.w;
    out->n_gated = c.n_gated;
    out->n_hits = c.n_hits;

    /* ---- export: pruned pre table sorted by first gated stamp ---- */
    out->n_pre = n_keep;
    {
        uint64_t *ord = (uint64_t *)malloc((n_keep ? n_keep : 1) * 2 * sizeof(uint64_t));
        uint64_t j = 0;
        for (uint64_t i = 0; i < c.n_pre; i++)
            if (c.pre[i].alive) { ord[2 * j] = stamp_of(&m, c.pre[i].key); ord[2 * j + 1] = i; j++; }
        qsort(ord, n_keep, 2 * sizeof(uint64_t), cmp_u64_pair);
        out->pre_first_pos = (uint64_t *)malloc((n_keep ? n_keep : 1) * sizeof(uint64_t));
        out->pre_freq = (uint16_t *)malloc((n_keep ? n_keep : 1) * sizeof(uint16_t));
        out->pre_qual_sums = (uint8_t *)malloc((n_keep ? n_keep : 1) * (size_t)c.k);
        for (j = 0; j < n_keep; j++) {
            pre_entry *e = &c.pre[ord[2 * j + 1]];
            out->pre_first_pos[j] = ord[2 * j];
            out->pre_freq[j] = e->freq;
            memcpy(out->pre_qual_sums + j * (size_t)c.k, e->qsum, (size_t)c.k);
        }
        free(ord);
    }
    /* ---- export: nodes in creation order ---- */
    {
        uint64_t n = c.n_nodes, na = n ? n : 1;
        out->n_nodes = n;
        out->node_first_pos = (uint64_t *)malloc(na * sizeof(uint64_t));
        out->node_freq = (uint16_t *)malloc(na * sizeof(uint16_t));
        out->out_deg = (uint8_t *)calloc(na, 1);
        out->in_deg = (uint8_t *)calloc(na, 1);
        out->out_succ = (uint32_t *)malloc(na * 4 * sizeof(uint32_t));
        out->in_pred = (uint32_t *)malloc(na * 4 * sizeof(uint32_t));
        memset(out->out_succ, 0xFF, na * 4 * sizeof(uint32_t));
        memset(out->in_pred, 0xFF, na * 4 * sizeof(uint32_t));
        for (uint64_t i = 0; i < n; i++) {
            graph_node *g = &c.nodes[i];
            out->node_first_pos[i] = stamp_of(&m, g->kmer);
            out->node_freq[i] = g->freq;
            int d = 0;
            for (uint32_t p = g->to_head; p != NIL; p = c.cells[p].next) {
                if (d < 4) out->out_succ[i * 4 + d] = c.cells[p].node;
                d++;
            }
            out->out_deg[i] = (uint8_t)d;
            d = 0;
            for (uint32_t p = g->from_head; p != NIL; p = c.cells[p].next) {
                if (d < 4) out->in_pred[i * 4 + d] = c.cells[p].node;
                d++;
            }
            out->in_deg[i] = (uint8_t)d;
        }
    }
    rc = 0;
fail:
    free(c.pre); free(c.pre_buckets); free(c.nodes); free(c.node_buckets); free(c.cells);
    if (rc) vdj_oracle_free(out);
    return rc;
}

void vdj_oracle_free(vdj_oracle_result *r) {
    if (!r) return;
    free(r->pre_first_pos); free(r->pre_freq); free(r->pre_qual_sums);
    free(r->node_first_pos); free(r->node_freq); free(r->out_deg); free(r->out_succ);
    free(r->in_deg); free(r->in_pred);
    memset(r, 0, sizeof(*r));
}
