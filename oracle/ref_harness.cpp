/*
 * ref_harness.cpp -- drives the REFERENCE's own graph-build functions, compiled in place from
 * /root/reference (never copied into this repo), and dumps their tables in the layout of
 * vdj_oracle.h.  TEST INFRASTRUCTURE ONLY; built into oracle/_ref/libvdjref.so by oracle/Makefile.
 *
 * How: the hot-path structs are private to assembler2_vdj.c, so this TU textually includes that
 * file (found through -I$(REF)/src/main/c) with its main() renamed, then calls, in the order of
 * the orchestrating block assembler2_vdj.c:1381-1415:
 *     build_pre_graph(primary) ; build_pre_graph(secondary) ; prune_pre_graph ;
 *     build_graph2(primary)    ; build_graph2(secondary)
 * bam_read.c (htslib) is not linked: the two functions assembler2_vdj.c imports from it are
 * stubbed below; they are never reached from the functions above.
 *
 * No reference source text appears in this file.
 */
#define main vdjer_reference_main
#include "assembler2_vdj.c"
#undef main

#include "vdj_oracle.h"

void set_default_params(params *p); /* params.c:53 */

/* bam_read.c stubs (declared extern at assembler2_vdj.c:33-36) */
void extract(char *, char *, char *, char *, char *&, char *&) { abort(); }
int get_read_length(char *) { abort(); return 0; }

namespace {

typedef dense_hash_map<const char *, pre_node, my_hash, eqstr> pre_map_t;
typedef dense_hash_map<const char *, struct node *, my_hash, eqstr> node_map_t;

double mono_s() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

bool g_vjf_ready = false;

struct Stamper {
    const char *p, *s;
    size_t np, ns, rec_len;
    int w;
    uint64_t operator()(const char *kmer) const {
        const char *base = s;
        uint64_t rec0 = np;
        if (kmer >= p && kmer < p + np * rec_len) { base = p; rec0 = 0; }
        size_t off = (size_t)(kmer - base);
        return (rec0 + off / rec_len) * (uint64_t)w + (off % rec_len - 1);
    }
};

} // namespace

extern "C" int vdjref_build(const char *primary, const char *secondary, int L, int k, int mf,
                            int mq, const char *scratch_dir, vdj_oracle_result *out) {
    memset(out, 0, sizeof(*out));
    set_default_params(&p);
    p.kmer = k;
    p.min_node_freq = mf;
    p.min_base_quality = mq > MAX_QUAL_SUM - 1 ? MAX_QUAL_SUM - 1 : mq; /* main(), :1514-1516 */
    if (!g_vjf_ready) {
        /* matches_vmer/jmer dereference global sets that only vjf_init allocates */
        std::string v = std::string(scratch_dir) + "/empty_v_index";
        std::string j = std::string(scratch_dir) + "/empty_j_index";
        FILE *f = fopen(v.c_str(), "w"); if (!f) return -3; fclose(f);
        f = fopen(j.c_str(), "w"); if (!f) return -3; fclose(f);
        vjf_init((char *)v.c_str(), (char *)j.c_str(), 4, 10, 90, 'W', 486, 162);
        g_vjf_ready = true;
    }
    read_length = L;
    kmer_size = k;
    node_id = 1;

    Stamper stamp = { primary, secondary, strlen(primary) / (size_t)(2 * L + 1),
                      strlen(secondary) / (size_t)(2 * L + 1), (size_t)(2 * L + 1), L - k + 1 };

    struct_pool pool;
    memset(&pool, 0, sizeof(pool));
    node_map_t *nodes = new node_map_t();
    nodes->set_empty_key(NULL);
    {
        pre_map_t pre_nodes;
        pre_nodes.set_empty_key(NULL);
        char *deleted_key = (char *)calloc(k, 1);
        pre_nodes.set_deleted_key(deleted_key);

        double t0 = mono_s();
        build_pre_graph(primary, pre_nodes);
        build_pre_graph(secondary, pre_nodes);
        double t1 = mono_s();
        out->n_pre_total = pre_nodes.size();
        prune_pre_graph(pre_nodes);
        double t2 = mono_s();

        pool.nodes = (struct node *)calloc(pre_nodes.size() + 1, sizeof(struct node));
        pool.idx = 0;
        pool.size = pre_nodes.size() + 3;
        build_graph2(primary, nodes, &pool, 1, pre_nodes);
        build_graph2(secondary, nodes, &pool, 0, pre_nodes);
        double t3 = mono_s();
        out->t_pass1 = t1 - t0; out->t_prune = t2 - t1; out->t_pass2 = t3 - t2;

        /* dump the pruned pre table sorted by the stamp of its key pointer */
        size_t n = pre_nodes.size();
        out->n_pre = n;
        std::vector<std::pair<uint64_t, const pre_node *> > ord;
        ord.reserve(n);
        for (pre_map_t::const_iterator it = pre_nodes.begin(); it != pre_nodes.end(); ++it)
            ord.push_back(std::make_pair(stamp(it->first), &it->second));
        std::sort(ord.begin(), ord.end());
        size_t na = n ? n : 1;
        out->pre_first_pos = (uint64_t *)malloc(na * sizeof(uint64_t));
        out->pre_freq = (uint16_t *)malloc(na * sizeof(uint16_t));
        out->pre_qual_sums = (uint8_t *)malloc(na * (size_t)k);
        for (size_t i = 0; i < n; i++) {
            out->pre_first_pos[i] = ord[i].first;
            out->pre_freq[i] = ord[i].second->frequency;
            memcpy(out->pre_qual_sums + i * (size_t)k, ord[i].second->qual_sums, (size_t)k);
        }
        free(deleted_key);
    }

    out->n_records = stamp.np + stamp.ns;
    out->n_windows = out->n_records * (uint64_t)(L - k + 1);
    size_t n = (size_t)pool.idx, na = n ? n : 1;
    out->n_nodes = n;
    out->node_first_pos = (uint64_t *)malloc(na * sizeof(uint64_t));
    out->node_freq = (uint16_t *)malloc(na * sizeof(uint16_t));
    out->out_deg = (uint8_t *)calloc(na, 1);
    out->in_deg = (uint8_t *)calloc(na, 1);
    out->out_succ = (uint32_t *)malloc(na * 4 * sizeof(uint32_t));
    out->in_pred = (uint32_t *)malloc(na * 4 * sizeof(uint32_t));
    memset(out->out_succ, 0xFF, na * 4 * sizeof(uint32_t));
    memset(out->in_pred, 0xFF, na * 4 * sizeof(uint32_t));
    int bad = 0;
    for (size_t i = 0; i < n; i++) {
        struct node *g = &pool.nodes[i];
        if (g->id != (int)i + 1) bad = 1;
        out->node_first_pos[i] = stamp(g->kmer);
        out->node_freq[i] = g->frequency;
        int d = 0;
        for (linked_node *c = g->toNodes; c; c = c->next, d++)
            if (d < 4) out->out_succ[i * 4 + d] = (uint32_t)(c->node - pool.nodes);
        out->out_deg[i] = (uint8_t)d;
        d = 0;
        for (linked_node *c = g->fromNodes; c; c = c->next, d++)
            if (d < 4) out->in_pred[i * 4 + d] = (uint32_t)(c->node - pool.nodes);
        out->in_deg[i] = (uint8_t)d;
    }
    /* release what the reference leaks */
    for (size_t i = 0; i < n; i++) {
        for (linked_node *c = pool.nodes[i].toNodes; c;) { linked_node *x = c->next; free(c); c = x; }
        for (linked_node *c = pool.nodes[i].fromNodes; c;) { linked_node *x = c->next; free(c); c = x; }
    }
    free(pool.nodes);
    delete nodes;
    return bad ? -4 : 0;
}

extern "C" void vdjref_free(vdj_oracle_result *r) {
    if (!r) return;
    free(r->pre_first_pos); free(r->pre_freq); free(r->pre_qual_sums);
    free(r->node_first_pos); free(r->node_freq); free(r->out_deg); free(r->out_succ);
    free(r->in_deg); free(r->in_pred);
    memset(r, 0, sizeof(*r));
}
