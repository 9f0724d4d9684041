#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Builds, from the sources where they lie under /root/reference, the two
whole-program binaries that tests/test_e2e_binary.py compares, plus a SAM->BAM helper:

    oracle/_ref/vdjer_ref   the reference program (its own graph build on one host core)
    oracle/_ref/vdjer_gpu   the same program with the block assembler2_vdj.c:1381-1415 replaced by
                            libvdjgraph through glue/vdjgraph_glue.inc, exactly the change shown in
                            INTEGRATION.md
    oracle/_ref/sam2bam     oracle/e2e/sam2bam.c + the vendored htslib

Nothing from the reference is copied into the repository: sources are compiled in place; the few
files that need a patch are patched COPIES in a temporary directory that is deleted afterwards:
  * four functions that fall off their end without `return` (g++ 13 emits a trap there):
    params.c parse_params, bam_read.c rc/reverse, assembler2_vdj.c worker_thread;
  * the worker-exit race of worker_thread (:1099 reads the queue size before the "all roots handed
    out" flag, so up to 5 queued roots per thread can be dropped): the two operands of
    the loop condition are swapped, in BOTH binaries, so that vdj_contigs.fa is reproducible;
  * vdjer_gpu only: the INTEGRATION.md diffs (the assemble() block; the two record-buffer callocs
    of bam_read.c:386-388 taken from page-locked memory; add_to_buffer notes every forward record).
htslib is compiled from its .c files with plain gcc commands (no reference build system is run).
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("VDJER_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "src", "main", "c")
HTS = os.path.join(REF, "samtools-1.2", "htslib-1.2.1")
OUT = os.path.join(HERE, "_ref")

HTS_TUS = ("kfunc knetfile kstring bgzf faidx hfile hfile_net hts regidx sam synced_bcf_reader vcf_sweep tbx vcf vcfutils "
           "cram/cram_codecs cram/cram_decode cram/cram_encode cram/cram_index cram/cram_io cram/cram_samtools "
           "cram/cram_stats cram/files cram/mFILE cram/md5 cram/open_trace_file cram/pooled_alloc cram/rANS_static "
           "cram/sam_header cram/string_alloc cram/thread_pool cram/vlen cram/zfio").split()
REF_TUS = "assembler2_vdj seq_score vj_filter seq_to_kmer hash_utils bam_read quick_map3 coverage status params".split()


def run(cmd, **kw):
    r = subprocess.run(cmd, capture_output=True, text=True, **kw)
    if r.returncode != 0:
        raise RuntimeError(" ".join(cmd) + "\n" + r.stdout[-3000:] + r.stderr[-3000:])


def sub_once(text: str, pattern: str, repl: str, what: str, count: int = 1) -> str:
    new, n = re.subn(pattern, repl, text, count=0, flags=re.S)
    if n != count:
        raise RuntimeError(f"patch '{what}': expected {count} match(es), found {n}")
    return new


def patched_sources(tmp: str, gpu: bool) -> list[str]:
    """Paths of the ten translation units; patched copies live in tmp."""
    out = []
    for tu in REF_TUS:
        path = os.path.join(SRC, tu + ".c")
        text = open(path).read()
        if tu == "params":
            text = sub_once(text, r"(\tvalidate_params\(p\);\n)(\})", r"\1\treturn 0;\n\2", "parse_params return")
        elif tu == "bam_read":
            text = sub_once(text, r"(\toutput\[strlen\(input\)\] = '\\0';\n)(\})", r"\1\treturn 0;\n\2", "rc/reverse return", 2)
            if gpu:
                # the record buffers in page-locked memory (bam_read.c:386-388)
                text = sub_once(text, r"primary_buf = \(char\*\) calloc\((primary_reads\.size\(\) \* \(read_len\*8 \+ 4\) \+ 1), sizeof\(char\)\);",
                                r"primary_buf = (char*) vdjgraph_records_calloc(\1, 0);", "primary record buffer")
                text = sub_once(text, r"secondary_buf = \(char\*\) calloc\((secondary_reads\.size\(\) \* \(read_len\*8 \+ 4\) \+ 1), sizeof\(char\)\);",
                                r"secondary_buf = (char*) vdjgraph_records_calloc(\1, 1);", "secondary record buffer")
                # section 3c: every read is also kept once (add_to_buffer, :206-244)
                text = sub_once(text, r"\nvoid add_to_buffer\(bam1_t \*b, char\*& buf_ptr,",
                                "\nchar* vdjgraph_records_calloc(size_t, int);\nvoid vdjgraph_records_note_read(const char*, size_t);\n\n"
                                "void add_to_buffer(bam1_t *b, char*& buf_ptr,", "glue declarations")
                text = sub_once(text, r"(\tbuf_ptr\[0\] = '0';\n\tbuf_ptr \+= 1;\n\tstrncpy\(buf_ptr, seq, read_len\);)",
                                r"\tchar* vdjgraph_record_start = buf_ptr;\n\1", "forward record start")
                text = sub_once(text, r"(\t// Now add the reverse alignment\n)",
                                r"\tvdjgraph_records_note_read(vdjgraph_record_start, 2 * read_len + 1);\n\1", "forward record noted")
        elif tu == "assembler2_vdj":
            text = sub_once(text, r"while \(num_roots_in_thread\(thread\) > 0 \|\| !all_roots_processed\) \{",
                            "while (!all_roots_processed || num_roots_in_thread(thread) > 0) {", "worker exit race")
            text = sub_once(text, r"(//\t\t\tfprintf\(stderr, \"\\n\"\);\n\t\t\}\n\t\}\n)(\})", r"\1\treturn NULL;\n\2", "worker_thread return")
            if gpu:
                text = sub_once(text, r"\nchar\* assemble\(const char\* input,",
                                "\n#include \"vdjgraph_glue.inc\"\n\nchar* assemble(const char* input,", "glue include")
                block = (r"(\t\tdense_hash_map<const char\*, pre_node, my_hash, eqstr> pre_nodes;\n.*?"
                         r"\t\tpre_nodes\.resize\(0\);\n)(\t\} // End pre_node block)")
                text = sub_once(text, block,
                                "#ifdef VDJER_WITH_VDJGRAPH\n"
                                "\t\tif (vdjgraph_assemble_block(input, unaligned_input, nodes, pool) != 0)\n\t\t\texit(-1);\n"
                                "\t\troot_nodes = identify_root_nodes(nodes);\n"
                                "#else\n\\1#endif\n\\2", "assemble() block")
        else:
            out.append(path)
            continue
        dst = os.path.join(tmp, ("gpu_" if gpu else "ref_") + tu + ".c")
        open(dst, "w").write(text)
        out.append(dst)
    return out


def main():
    if not os.path.exists(os.path.join(SRC, "assembler2_vdj.c")):
        print("reference tree not found; nothing built", file=sys.stderr)
        return 0
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="vdjer_e2e_")
    try:
        # htslib objects
        open(os.path.join(tmp, "version.h"), "w").write('#define HTS_VERSION "1.2.1"\n')
        objs = []
        for tu in HTS_TUS:
            o = os.path.join(tmp, tu.replace("/", "_") + ".o")
            run(["gcc", "-O2", "-w", "-fPIC", "-I", tmp, "-I", HTS, "-c", os.path.join(HTS, tu + ".c"), "-o", o])
            objs.append(o)
        lib = os.path.join(tmp, "libhts.a")
        run(["ar", "rcs", lib] + objs)
        run(["gcc", "-O2", "-w", "-I", HTS, os.path.join(HERE, "e2e", "sam2bam.c"), lib, "-lz", "-lpthread", "-lm",
             "-o", os.path.join(OUT, "sam2bam")])
        inc = ["-I", SRC, "-I", os.path.join(REF, "samtools-1.2"), "-I", HTS]
        run(["g++", "-O2", "-g", "-w", "-pthread"] + inc + patched_sources(tmp, False) +
            [lib, "-lz", "-lpthread", "-o", os.path.join(OUT, "vdjer_ref")])
        gpu_lib = os.path.join(ROOT, "vdjer_b200", "libvdjgraph.so")
        if os.path.exists(gpu_lib):
            run(["g++", "-O2", "-g", "-w", "-pthread", "-DVDJER_WITH_VDJGRAPH", "-I", os.path.join(ROOT, "include"),
                 "-I", os.path.join(ROOT, "glue")] + inc + patched_sources(tmp, True) +
                [lib, "-lz", "-lpthread", "-L", os.path.join(ROOT, "vdjer_b200"), "-lvdjgraph",
                 "-Wl,-rpath,$ORIGIN/../../vdjer_b200", "-Wl,-rpath," + os.path.join(ROOT, "vdjer_b200"),
                 "-o", os.path.join(OUT, "vdjer_gpu")])
        else:
            print("libvdjgraph.so not built yet: vdjer_gpu skipped", file=sys.stderr)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
