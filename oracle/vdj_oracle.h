/*
 * vdj_oracle.h -- CPU oracle for V'DJer's de Bruijn graph build.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker or the timed CPU baseline.  The product (libvdjgraph.so) never
 * links, loads or calls it.
 *
 * The result layout below is shared by
 *   - liboracle.so     (oracle/vdj_oracle.c  : our sequential restatement, "port"), and
 *   - _ref/libvdjref.so (oracle/ref_harness.cpp: the reference's own functions, "reference"),
 * so that one comparison routine checks both against the CUDA path.
 *
 * Parity pin: the reference has no golden vectors of its own; the restatement is
 * pinned against outputs of the compiled reference (oracle/_ref) on seeded inputs; those outputs
 * are committed under tests/golden/ together with the generating script.
 */
#ifndef VDJ_ORACLE_H
#define VDJ_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VDJ_ORACLE_MAX_KMER 50 /* MAX_KMER_LEN, assembler2_vdj.c:70 */

/*
 * Positions ("stamps"): records are numbered in processing order, all primary-buffer records
 * first, then all secondary-buffer records (assembler2_vdj.c:1388-1390, 1402-1408).  Window i of
 * record r has stamp r*w + i with w = read_length - kmer_size + 1.
 */
typedef struct vdj_oracle_result {
    /* ---- pass 1 + prune (pre_nodes) ---- */
    uint64_t n_records;       /* primary + secondary */
    uint64_t n_windows;       /* n_records * w */
    uint64_t n_gated;         /* windows passing include_kmer */
    uint64_t n_pre_total;     /* "Pre Num nodes": distinct gated k-mers before pruning */
    uint64_t n_pre;           /* "pre nodes after pruning" */
    /* survivors sorted by the stamp of their first gated occurrence (the map key pointer) */
    uint64_t *pre_first_pos;  /* [n_pre] */
    uint16_t *pre_freq;       /* [n_pre] pre_node.frequency */
    uint8_t  *pre_qual_sums;  /* [n_pre * kmer_size] pre_node.qual_sums */

    /* ---- pass 2 (nodes) in creation order: node i has id i+1 ---- */
    uint64_t n_nodes;
    uint64_t n_hits;          /* pass-2 windows found in pre_nodes (uncapped) */
    uint64_t *node_first_pos; /* [n_nodes] stamp node->kmer points at */
    uint16_t *node_freq;      /* [n_nodes] node->frequency */
    uint8_t  *out_deg;        /* [n_nodes] length of toNodes */
    uint32_t *out_succ;       /* [n_nodes*4] toNodes in list order (head first), 0-based node index, 0xFFFFFFFF pad */
    uint8_t  *in_deg;         /* [n_nodes] length of fromNodes */
    uint32_t *in_pred;        /* [n_nodes*4] fromNodes in list order (head first) */

    /* ---- timing of the stages, seconds (monotonic clock) ---- */
    double t_pass1, t_prune, t_pass2;
} vdj_oracle_result;

/*
 * primary / secondary: NUL-terminated record buffers in the reference's format
 * (bam_read.c:206-244): record = strand char ('0'|'1') + read_length bases + read_length
 * phred+33 qualities.  Either may be "" (not NULL).
 * Returns 0, or a negative code for input the reference would exit(-1) on.
 */
int vdj_oracle_build(const char *primary, const char *secondary, int read_length, int kmer_size,
                     int min_node_freq, int min_base_quality, vdj_oracle_result *out);

void vdj_oracle_free(vdj_oracle_result *r);

#ifdef __cplusplus
}
#endif
#endif
