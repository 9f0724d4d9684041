/*
 * synth.c -- seeded synthetic paired-end Ig-locus reads in V'DJer's record-buffer format.
 *
 * Workload generator for tests and bench.py; not part of the graph build.
 * Output format = what the reference's add_to_buffer writes (bam_read.c:206-244): per mate a
 * forward record and its reverse-complement record (qualities reversed), each
 *     '0' + L bases + L phred+33 qualities,
 * concatenated and NUL-terminated; two buffers, "primary" (V-region / anchor-hit reads) and
 * "secondary" (C-region / unmapped reads), processed in that order (assembler2_vdj.c:1388-1390).
 *
 * Model: a clone library of transcripts  V(300, germline gene + 1-5 % somatic mutations)
 * + CDR3(random, cdr3_min..cdr3_max) + J(48) + constant(352); clone abundance ~ Zipf(s);
 * fragments N(insert_mean, insert_sd) clipped to [L, 400]; per-read quality profiles with
 * low-quality tails; substitution probability tied to quality; rare 'N'.
 *
 * Everything is a pure function of (seed, pair index), so the output does not depend on the
 * thread count.
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct vdjsynth_params {
    uint64_t seed;
    uint64_t n_pairs;
    int32_t read_length;
    int32_t n_clones;
    int32_t n_v_genes;
    int32_t n_j_genes;
    int32_t cdr3_min, cdr3_max;
    double zipf_s;          /* clone abundance exponent; 0 = flat */
    double frac_secondary;  /* share of pairs written to the secondary buffer */
    double insert_mean, insert_sd;
    double frac_bad_tail;   /* reads whose quality collapses towards the 3' end */
    double p_low_base;      /* isolated low-quality bases in good reads */
    double p_n;             /* 'N' rate */
    int32_t threads;
    uint64_t pair_offset;   /* pairs are numbered pair_offset .. pair_offset + n_pairs - 1 in the (seed-defined) stream:
                               several callers can draw disjoint read sets from the SAME clone library */
    int32_t forward_only;   /* 1: only the forward record of every mate (2 records per pair instead of 4): the even
                               records of the default output, i.e. what vdjgraph_stage_forward takes */
    int32_t reserved;
} vdjsynth_params;

#define V_LEN 300
#define J_LEN 48
#define C_LEN 352
#define MAX_T (V_LEN + 64 + J_LEN + C_LEN)

static inline uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
typedef struct rng { uint64_t s; } rng;
static inline rng rng_make(uint64_t seed, uint64_t stream, uint64_t idx) {
    rng r = { mix64(seed ^ mix64(stream * 0xD1342543DE82EF95ull + idx)) };
    return r;
}
static inline uint64_t rng_next(rng *r) { r->s += 0x9E3779B97F4A7C15ull; return mix64(r->s); }
static inline double rng_unif(rng *r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint32_t rng_below(rng *r, uint32_t n) { return (uint32_t)(((rng_next(r) >> 32) * (uint64_t)n) >> 32); }

static const char BASES[5] = "ACGT";

typedef struct library {
    char *transcripts;   /* n_clones * MAX_T */
    int32_t *t_len;
    double *cdf;         /* n_clones */
    int32_t n_clones;
} library;

static void build_library(const vdjsynth_params *p, library *lib) {
    int nv = p->n_v_genes, nj = p->n_j_genes;
    char *v = (char *)malloc((size_t)nv * V_LEN), *j = (char *)malloc((size_t)nj * J_LEN);
    char c[C_LEN];
    /* germline genes: related to a common ancestor (70 % shared bases) like a real V family */
    rng g = rng_make(p->seed, 1, 0);
    char anc[V_LEN];
    for (int i = 0; i < V_LEN; i++) anc[i] = BASES[rng_below(&g, 4)];
    for (int a = 0; a < nv; a++)
        for (int i = 0; i < V_LEN; i++)
            v[a * V_LEN + i] = rng_unif(&g) < 0.7 ? anc[i] : BASES[rng_below(&g, 4)];
    for (int a = 0; a < nj * J_LEN; a++) j[a] = BASES[rng_below(&g, 4)];
    for (int i = 0; i < C_LEN; i++) c[i] = BASES[rng_below(&g, 4)];

    lib->n_clones = p->n_clones;
    lib->transcripts = (char *)malloc((size_t)p->n_clones * MAX_T);
    lib->t_len = (int32_t *)malloc((size_t)p->n_clones * sizeof(int32_t));
    lib->cdf = (double *)malloc((size_t)p->n_clones * sizeof(double));
    double acc = 0;
    for (int cl = 0; cl < p->n_clones; cl++) {
        rng r = rng_make(p->seed, 2, (uint64_t)cl);
        char *t = lib->transcripts + (size_t)cl * MAX_T;
        int vg = (int)rng_below(&r, (uint32_t)nv), jg = (int)rng_below(&r, (uint32_t)nj);
        double mut = 0.01 + 0.04 * rng_unif(&r);
        int n = 0;
        for (int i = 0; i < V_LEN; i++) {
            char b = v[vg * V_LEN + i];
            if (rng_unif(&r) < mut) b = BASES[rng_below(&r, 4)];
            t[n++] = b;
        }
        int cl3 = p->cdr3_min + (int)rng_below(&r, (uint32_t)(p->cdr3_max - p->cdr3_min + 1));
        for (int i = 0; i < cl3; i++) t[n++] = BASES[rng_below(&r, 4)];
        memcpy(t + n, j + jg * J_LEN, J_LEN); n += J_LEN;
        memcpy(t + n, c, C_LEN); n += C_LEN;
        lib->t_len[cl] = n;
        acc += pow((double)(cl + 1), -p->zipf_s);
        lib->cdf[cl] = acc;
    }
    for (int cl = 0; cl < p->n_clones; cl++) lib->cdf[cl] /= acc;
    free(v); free(j);
}

static inline char comp(char b) {
    switch (b) { case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; default: return b; }
}

static void make_read(const vdjsynth_params *p, rng *r, const char *src, int L, char *seq, char *qual) {
    int tail_from = L;
    if (rng_unif(r) < p->frac_bad_tail) tail_from = (int)(L * (0.35 + 0.6 * rng_unif(r)));
    for (int i = 0; i < L; i++) {
        int q;
        if (i >= tail_from && rng_unif(r) < 0.7) q = 2 + (int)rng_below(r, 18);
        else if (rng_unif(r) < p->p_low_base) q = 2 + (int)rng_below(r, 18);
        else q = 30 + (int)rng_below(r, 11);
        char b = src[i];
        double perr = q >= 30 ? 0.001 : fmin(0.2, pow(10.0, -q / 10.0));
        if (rng_unif(r) < perr) b = BASES[(rng_below(r, 3) + 1 + (uint32_t)(strchr(BASES, b) - BASES)) & 3];
        if (rng_unif(r) < p->p_n) { b = 'N'; q = 2; }
        seq[i] = b;
        qual[i] = (char)(q + 33);
    }
}

static void write_records(char *dst, const char *seq, const char *qual, int L, int forward_only) {
    /* forward record, then reverse complement with reversed qualities */
    dst[0] = '0';
    memcpy(dst + 1, seq, (size_t)L);
    memcpy(dst + 1 + L, qual, (size_t)L);
    if (forward_only) return;
    char *d2 = dst + 2 * L + 1;
    d2[0] = '0';
    for (int i = 0; i < L; i++) {
        d2[1 + i] = comp(seq[L - 1 - i]);
        d2[1 + L + i] = qual[L - 1 - i];
    }
}

typedef struct job {
    const vdjsynth_params *p;
    const library *lib;
    char *primary, *secondary;
    uint64_t n_primary_pairs, lo, hi;
} job;

static void *worker(void *arg) {
    job *jb = (job *)arg;
    const vdjsynth_params *p = jb->p;
    const library *lib = jb->lib;
    const int L = p->read_length;
    const size_t rec = (size_t)2 * L + 1;
    char seq[512], qual[512];
    for (uint64_t pair = jb->lo; pair < jb->hi; pair++) {
        rng r = rng_make(p->seed, 3, p->pair_offset + pair);
        double u = rng_unif(&r);
        int lo = 0, hi = lib->n_clones - 1;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (lib->cdf[mid] < u) lo = mid + 1; else hi = mid; }
        const char *t = lib->transcripts + (size_t)lo * MAX_T;
        int tl = lib->t_len[lo];
        /* Box-Muller insert size */
        double g = sqrt(-2.0 * log(rng_unif(&r) + 1e-300)) * cos(6.283185307179586 * rng_unif(&r));
        int ins = (int)lrint(p->insert_mean + p->insert_sd * g);
        if (ins < L) ins = L;
        if (ins > 400) ins = 400;
        if (ins > tl) ins = tl;
        int start = (int)rng_below(&r, (uint32_t)(tl - ins + 1));
        const size_t per_mate = p->forward_only ? rec : 2 * rec;   /* bytes one mate occupies */
        char *dst = pair < jb->n_primary_pairs ? jb->primary + pair * 2 * per_mate
                                                : jb->secondary + (pair - jb->n_primary_pairs) * 2 * per_mate;
        make_read(p, &r, t + start, L, seq, qual);
        write_records(dst, seq, qual, L, p->forward_only);
        make_read(p, &r, t + start + ins - L, L, seq, qual);
        write_records(dst + per_mate, seq, qual, L, p->forward_only);
    }
    return NULL;
}

/* Sizes the caller must allocate: bytes including the terminating NUL. */
void vdjsynth_sizes(const vdjsynth_params *p, uint64_t *primary_bytes, uint64_t *secondary_bytes,
                    uint64_t *primary_records, uint64_t *secondary_records) {
    uint64_t n_sec = (uint64_t)((double)p->n_pairs * p->frac_secondary);
    uint64_t n_pri = p->n_pairs - n_sec;
    uint64_t rec = (uint64_t)2 * p->read_length + 1;
    uint64_t per_pair = p->forward_only ? 2 : 4;
    *primary_records = n_pri * per_pair; *secondary_records = n_sec * per_pair;
    *primary_bytes = n_pri * per_pair * rec + 1; *secondary_bytes = n_sec * per_pair * rec + 1;
}

int vdjsynth_generate(const vdjsynth_params *p, char *primary, char *secondary) {
    if (p->read_length < 20 || p->read_length > 255 || p->n_clones < 1 || p->n_v_genes < 1 ||
        p->n_j_genes < 1 || p->cdr3_min < 0 || p->cdr3_max > 64 || p->cdr3_min > p->cdr3_max)
        return -1;
    library lib;
    build_library(p, &lib);
    uint64_t pb, sb, pr, sr;
    vdjsynth_sizes(p, &pb, &sb, &pr, &sr);
    int nt = p->threads > 0 ? p->threads : 1;
    if (nt > 256) nt = 256;
    pthread_t th[256];
    job jobs[256];
    for (int i = 0; i < nt; i++) {
        jobs[i].p = p; jobs[i].lib = &lib; jobs[i].primary = primary; jobs[i].secondary = secondary;
        jobs[i].n_primary_pairs = pr / (p->forward_only ? 2 : 4);
        jobs[i].lo = p->n_pairs * (uint64_t)i / (uint64_t)nt;
        jobs[i].hi = p->n_pairs * (uint64_t)(i + 1) / (uint64_t)nt;
        pthread_create(&th[i], NULL, worker, &jobs[i]);
    }
    for (int i = 0; i < nt; i++) pthread_join(th[i], NULL);
    primary[pb - 1] = 0;
    secondary[sb - 1] = 0;
    free(lib.transcripts); free(lib.t_len); free(lib.cdf);
    return 0;
}
