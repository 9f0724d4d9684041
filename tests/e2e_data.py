"""Synthetic whole-program input for the V'DJer binary: a coordinate-sorted SAM
with paired-end reads of a few B-cell clones (shared V and J, clone-specific CDR3) placed inside
the IGH locus window of --chain IGH (params.c:14), plus the --ref-dir files.  The reference's demo
BAM (demo/star.sort.bam) is not in the tree, so this stands in for BASELINE configs[0]."""
from __future__ import annotations

import os

import numpy as np

CODE = {"A": 0, "T": 1, "C": 2, "G": 3}           # seq_to_int, seq_to_kmer.c:6-29
COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}
STOPS = {"TAA", "TAG", "TGA"}


def seq_to_int(s: str) -> int:
    v = 0
    for ch in s[:16]:
        v = (v << 2) | CODE[ch]
    return v


def _orf(rng, n_codons: int) -> str:
    out = []
    while len(out) < n_codons:
        c = "".join("ACGT"[i] for i in rng.integers(0, 4, 3))
        if c not in STOPS:
            out.append(c)
    return "".join(out)


def make_case(workdir: str, n_clones: int = 4, pairs_per_clone: int = 2500, read_length: int = 50, seed: int = 7):
    """Writes workdir/x.sam and workdir/ref/*; returns the SAM path and the transcripts."""
    rng = np.random.default_rng(seed)
    V, J = _orf(rng, 110), _orf(rng, 140)             # 330 / 420 bases, stop-free in frame
    transcripts = [V + "TGT" + _orf(rng, 13) + "TGG" + J for _ in range(n_clones)]
    ref = os.path.join(workdir, "ref")
    os.makedirs(ref, exist_ok=True)
    T = transcripts[0]
    # anchors: Cys (330) must lie in [v, v+30), Trp (372) in [j-16, j+14)  (vj_filter.c:140,152)
    open(os.path.join(ref, "v_index"), "w").write(f"{seq_to_int(T[312:328])}\t0\n")
    open(os.path.join(ref, "j_index"), "w").write(f"{seq_to_int(T[378:394])}\t0\n")
    open(os.path.join(ref, "ig_vdj.fa"), "w").write("".join(f">t{i}\n{t[280:420]}\n" for i, t in enumerate(transcripts)))
    open(os.path.join(ref, "v_region.fa"), "w").write(">v\n" + V + "\n")

    L = read_length
    base = 105_600_000                               # inside chr14:105566277-105939754 (c_region and v_region)
    recs = []
    qual = "I" * L
    n = 0
    for ci, t in enumerate(transcripts):
        for _ in range(pairs_per_clone):
            ins = int(rng.integers(150, 201))
            start = int(rng.integers(0, len(t) - ins + 1))
            r1 = t[start:start + L]
            frag_end = start + ins
            r2_fwd = t[frag_end - L:frag_end]         # stored as it aligns to the forward strand
            name = f"r{n}"
            n += 1
            p1, p2 = base + start + 1, base + frag_end - L + 1
            recs.append((p1, f"{name}\t99\tchr14\t{p1}\t60\t{L}M\t=\t{p2}\t{ins}\t{r1}\t{qual}"))
            recs.append((p2, f"{name}\t147\tchr14\t{p2}\t60\t{L}M\t=\t{p1}\t{-ins}\t{r2_fwd}\t{qual}"))
    recs.sort(key=lambda x: x[0])
    sam = os.path.join(workdir, "x.sam")
    with open(sam, "w") as f:
        f.write("@HD\tVN:1.4\tSO:coordinate\n@SQ\tSN:chr14\tLN:107043718\n")
        for _, line in recs:
            f.write(line + "\n")
        # extract() sizes its records from the first read it sees AFTER the region queries
        # (bam_read.c:355-362): there has to be one behind the locus, e.g. an unmapped pair
        u = "".join("ACGT"[i] for i in rng.integers(0, 4, L))
        f.write(f"u0\t77\t*\t0\t0\t*\t*\t0\t0\t{u}\t{qual}\n")
        f.write(f"u0\t141\t*\t0\t0\t*\t*\t0\t0\t{u[::-1]}\t{qual}\n")
    return sam, transcripts
