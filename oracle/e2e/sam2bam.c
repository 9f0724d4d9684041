/* sam2bam -- TEST INFRASTRUCTURE: SAM text -> coordinate-sorted BAM + .bai, with the htslib that is
 * vendored in the reference tree (compiled by oracle/build_e2e.py).  The test's SAM is written
 * already sorted.  usage: sam2bam in.sam out.bam */
#include <stdio.h>
#include "htslib/sam.h"

int main(int argc, char **argv) {
    if (argc != 3) { fprintf(stderr, "usage: sam2bam in.sam out.bam\n"); return 2; }
    samFile *in = sam_open(argv[1], "r");
    if (!in) { perror(argv[1]); return 1; }
    bam_hdr_t *h = sam_hdr_read(in);
    samFile *out = sam_open(argv[2], "wb");
    if (!h || !out || sam_hdr_write(out, h) < 0) { fprintf(stderr, "cannot write %s\n", argv[2]); return 1; }
    bam1_t *b = bam_init1();
    long n = 0;
    while (sam_read1(in, h, b) >= 0) { if (sam_write1(out, h, b) < 0) return 1; n++; }
    bam_destroy1(b);
    sam_close(out);
    sam_close(in);
    if (bam_index_build(argv[2], 0) < 0) { fprintf(stderr, "index failed\n"); return 1; }
    fprintf(stderr, "sam2bam: %ld records\n", n);
    return 0;
}
