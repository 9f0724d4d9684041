/*
 * glue_harness.cpp -- TEST INFRASTRUCTURE ONLY.  Proves that the arrays libvdjgraph exports are
 * sufficient for, and that glue/vdjgraph_glue.inc correctly rebuilds, what the reference's
 * downstream code consumes: the reference's own identify_root_nodes (:653), condense_graph (:598)
 * and dump_graph (:1133, "vdjer.dot") are run on
 *   (a) the graph built by the reference's own build_pre_graph/prune_pre_graph/build_graph2, and
 *   (b) the graph rebuilt by the glue from a vdjgraph_result,
 * and the two vdjer.dot files must be byte-identical (node ids, sparsehash iteration order, edge
 * list order, condensed sequences, root flags).
 *
 * Like ref_harness.cpp this TU includes the reference's assembler2_vdj.c in place (never copied);
 * built into oracle/_ref/libvdjglue.so by oracle/Makefile.
 */
#define main vdjer_reference_main
#include "assembler2_vdj.c"
#undef main

#define VDJGRAPH_GLUE_NO_LIBRARY   /* the result arrays are handed in by the test */
#include "../glue/vdjgraph_glue.inc"

void set_default_params(params *p); /* params.c:53 */
void extract(char *, char *, char *, char *, char *&, char *&) { abort(); }
int get_read_length(char *) { abort(); return 0; }

namespace {
typedef dense_hash_map<const char *, pre_node, my_hash, eqstr> pre_map_t;
typedef dense_hash_map<const char *, struct node *, my_hash, eqstr> node_map_t;
bool g_vjf_ready = false;
}

/* res == NULL: build with the reference's own functions.  Returns the node count or < 0. */
extern "C" long vdjglue_dot(const char *primary, const char *secondary, int L, int k, int mf, int mq,
                            const char *scratch_dir, const vdjgraph_result *res, const char *dot_path,
                            long *n_roots) {
    set_default_params(&p);
    p.kmer = k;
    p.min_node_freq = mf;
    p.min_base_quality = mq > MAX_QUAL_SUM - 1 ? MAX_QUAL_SUM - 1 : mq; /* main(), :1514-1516 */
    if (!g_vjf_ready) {
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
    struct_pool pool;
    memset(&pool, 0, sizeof(pool));
    node_map_t *nodes = new node_map_t();
    nodes->set_empty_key(NULL);
    if (res) {
        int rc = vdjgraph_rebuild_nodes(res, primary, strlen(primary) / (size_t)(2 * L + 1), secondary, nodes, &pool);
        if (rc) return rc;
    } else {
        pre_map_t pre_nodes;
        pre_nodes.set_empty_key(NULL);
        char *deleted_key = (char *)calloc(k, 1);
        pre_nodes.set_deleted_key(deleted_key);
        build_pre_graph(primary, pre_nodes);
        build_pre_graph(secondary, pre_nodes);
        prune_pre_graph(pre_nodes);
        pool.nodes = (struct node *)calloc(pre_nodes.size() + 1, sizeof(struct node));
        pool.idx = 0;
        pool.size = pre_nodes.size() + 3;
        build_graph2(primary, nodes, &pool, 1, pre_nodes);
        build_graph2(secondary, nodes, &pool, 0, pre_nodes);
        free(deleted_key);
    }
    struct linked_node *roots = identify_root_nodes(nodes);
    long nr = 0;
    for (struct linked_node *r = roots; r; r = r->next) nr++;
    if (n_roots) *n_roots = nr;
    condense_graph(nodes);
    dump_graph(nodes, dot_path);
    long n = (long)nodes->size();
    delete nodes;   /* the rest is leaked like in the reference; the harness is short-lived */
    return n;
}

/* Wall time of vdjgraph_rebuild_nodes alone (milliseconds; < 0 on error): what the glue costs a
 * caller after vdjgraph_build has returned.  The rebuilt graph is leaked like above. */
extern "C" double vdjglue_rebuild_ms(const char *primary, const char *secondary, int L, int k,
                                     const char *scratch_dir, const vdjgraph_result *res) {
    set_default_params(&p);
    p.kmer = k;
    if (!g_vjf_ready) {
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
    struct_pool pool;
    memset(&pool, 0, sizeof(pool));
    node_map_t *nodes = new node_map_t();
    nodes->set_empty_key(NULL);
    struct timespec t0, t1;
    const size_t n_primary = strlen(primary) / (size_t)(2 * L + 1);   /* (the caller's own strlen, :370-374: not part of the rebuild) */
    clock_gettime(CLOCK_MONOTONIC, &t0);
    int rc = vdjgraph_rebuild_nodes(res, primary, n_primary, secondary, nodes, &pool);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    delete nodes;
    if (rc) return (double)rc;
    return (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
}

/* The iteration order of the reference's `nodes` map after its own build_graph2 (node creation ranks,
 * i.e. id - 1, in dense_hash_map bucket order) and the bucket count: what the library's
 * hashmap-layout export (vdjgraph_result.hm_order) must reproduce.  Returns the node count or < 0. */
extern "C" long vdjglue_iteration_order(const char *primary, const char *secondary, int L, int k, int mf, int mq,
                                        const char *scratch_dir, uint32_t *out_ids, long cap, uint64_t *n_buckets) {
    set_default_params(&p);
    p.kmer = k;
    p.min_node_freq = mf;
    p.min_base_quality = mq > MAX_QUAL_SUM - 1 ? MAX_QUAL_SUM - 1 : mq;
    if (!g_vjf_ready) {
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
    struct_pool pool;
    memset(&pool, 0, sizeof(pool));
    node_map_t *nodes = new node_map_t();
    nodes->set_empty_key(NULL);
    pre_map_t pre_nodes;
    pre_nodes.set_empty_key(NULL);
    char *deleted_key = (char *)calloc(k, 1);
    pre_nodes.set_deleted_key(deleted_key);
    build_pre_graph(primary, pre_nodes);
    build_pre_graph(secondary, pre_nodes);
    prune_pre_graph(pre_nodes);
    pool.nodes = (struct node *)calloc(pre_nodes.size() + 1, sizeof(struct node));
    pool.idx = 0;
    pool.size = pre_nodes.size() + 3;
    build_graph2(primary, nodes, &pool, 1, pre_nodes);
    build_graph2(secondary, nodes, &pool, 0, pre_nodes);
    free(deleted_key);
    long n = 0;
    for (node_map_t::const_iterator it = nodes->begin(); it != nodes->end(); ++it) {
        if (n < cap) out_ids[n] = (uint32_t)(it->second->id - 1);
        n++;
    }
    if (n_buckets) *n_buckets = (uint64_t)nodes->bucket_count();
    delete nodes;
    return n;
}
