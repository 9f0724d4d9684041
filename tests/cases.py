"""Seeded parity cases shared by the golden generator, the CPU tests and the GPU tests.

L/k/mf/mq are read_length / --k / --mf / --mq.  Defaults of the reference: k=35 mf=3 mq=90
(params.c:53-73); "sensitive" is README's --k 25 --mq 60 --mf 2."""
from __future__ import annotations

import numpy as np

from vdjer_b200 import synth

CASES = {
    # BASELINE configs[1] shape (2x50, default flags), small
    "igh_default_k35": dict(L=50, k=35, mf=3, mq=90, gen=dict(n_pairs=6000, seed=101, n_clones=120)),
    # configs[2]: sensitive mode on a flat, diverse repertoire
    "igh_sensitive_k25": dict(L=50, k=25, mf=2, mq=60, gen=dict(n_pairs=3000, seed=102, n_clones=400, zipf_s=0.3)),
    # permissive: many branching nodes, exercises toNodes/fromNodes ordering
    "permissive_k25": dict(L=50, k=25, mf=1, mq=20, gen=dict(n_pairs=5000, seed=103, n_clones=60)),
    # configs[3] shape: IGK/IGL 2x75
    "igk_2x75_k35": dict(L=75, k=35, mf=3, mq=90, gen=dict(n_pairs=3000, seed=104, n_clones=80, n_v_genes=30,
                                                           cdr3_min=24, cdr3_max=36)),
    # configs[4] shape: 2x100
    "pooled_2x100_k35": dict(L=100, k=35, mf=3, mq=90, gen=dict(n_pairs=2500, seed=105, n_clones=100)),
    # key-width boundaries: 2k = 62 / 64 / 66 bits, the longest k the reference allows, and k <= SEQ_LEN
    "k31": dict(L=50, k=31, mf=2, mq=60, gen=dict(n_pairs=2500, seed=106, n_clones=50)),
    "k32": dict(L=50, k=32, mf=2, mq=60, gen=dict(n_pairs=2500, seed=107, n_clones=50)),
    "k33": dict(L=50, k=33, mf=2, mq=60, gen=dict(n_pairs=2500, seed=108, n_clones=50)),
    "k50_L100": dict(L=100, k=50, mf=3, mq=90, gen=dict(n_pairs=2000, seed=109, n_clones=60)),
    "k16": dict(L=50, k=16, mf=3, mq=90, gen=dict(n_pairs=2000, seed=110, n_clones=40)),
    "k_eq_L": dict(L=36, k=36, mf=2, mq=40, gen=dict(n_pairs=3000, seed=111, n_clones=10)),
    # quality-sum thresholds around the saturation point (214 -> 255) and the 254 clamp
    "mq214": dict(L=50, k=35, mf=3, mq=214, gen=dict(n_pairs=5000, seed=112, n_clones=80)),
    "mq215": dict(L=50, k=35, mf=3, mq=215, gen=dict(n_pairs=5000, seed=112, n_clones=80)),
    "mq999": dict(L=50, k=35, mf=3, mq=999, gen=dict(n_pairs=5000, seed=112, n_clones=80)),
    "mq150_mf1": dict(L=50, k=35, mf=1, mq=150, gen=dict(n_pairs=5000, seed=113, n_clones=80)),
    "mq0": dict(L=50, k=35, mf=2, mq=0, gen=dict(n_pairs=3000, seed=114, n_clones=80)),
    "mq_negative": dict(L=50, k=35, mf=0, mq=-5, gen=dict(n_pairs=3000, seed=115, n_clones=80)),
    # many low qualities and Ns: most windows fail the gate
    "noisy": dict(L=50, k=25, mf=2, mq=60, gen=dict(n_pairs=4000, seed=116, n_clones=30, frac_bad_tail=0.6,
                                                    p_low_base=0.05, p_n=0.01)),
    # all reads in the secondary buffer / all in the primary buffer
    "secondary_only": dict(L=50, k=35, mf=3, mq=90, gen=dict(n_pairs=2000, seed=117, n_clones=20, frac_secondary=1.0)),
    "primary_only": dict(L=50, k=35, mf=3, mq=90, gen=dict(n_pairs=2000, seed=118, n_clones=20, frac_secondary=0.0)),
    "k_L_minus_1": dict(L=36, k=35, mf=2, mq=40, gen=dict(n_pairs=3000, seed=120, n_clones=10)),
    # count saturation at 32765 (MAX_FREQUENCY-1, :66/:262/:345) in both passes
    "saturating": dict(L=20, k=8, mf=3, mq=90, hand="saturate"),
    # hand-made edge cases (inputs stored in the fixture)
    "hand_edges": dict(L=20, k=8, mf=2, mq=40, store_input=True, hand="edges"),
    "hand_duplicates_only": dict(L=20, k=8, mf=2, mq=40, store_input=True, hand="dups"),
    "hand_strand1": dict(L=20, k=8, mf=2, mq=40, store_input=True, hand="strand1"),
}


def _hand(kind: str):
    rng = np.random.default_rng(7)

    def rnd(n):
        return "".join("ACGT"[i] for i in rng.integers(0, 4, n))

    if kind == "edges":
        t = rnd(60)
        reads, quals = [], []
        for s in range(0, 41, 3):
            reads.append(t[s:s + 20]); quals.append("I" * 20)
        for s in range(1, 40, 4):  # the same positions again with mixed qualities
            reads.append(t[s:s + 20]); quals.append("".join("5I#?"[(s + j) % 4] for j in range(20)))
        reads.append("A" * 20); quals.append("I" * 20)          # homopolymer: self loop
        reads.append("A" * 20); quals.append("5" * 20)
        reads.append("A" * 19 + "C"); quals.append("I" * 20)
        reads.append("ACGT" * 5); quals.append("I" * 20)         # period-4 repeat: cycle
        reads.append("CGTA" * 5); quals.append("?" * 20)
        reads.append(t[5:14] + "N" + t[15:25]); quals.append("I" * 20)   # N in the middle
        reads.append("N" * 20); quals.append("I" * 20)                   # all N
        reads.append(t[10:30]); quals.append("!" * 20)                   # all quality 0
        reads.append(t[10:30]); quals.append("5" * 19 + "4")             # phred 20 vs 19 boundary
        reads.append(t[10:30]); quals.append("~" * 20)                   # phred 93
        p = synth.records_from_reads(reads[:20], quals[:20])
        s = synth.records_from_reads(reads[20:], quals[20:])
        return p, s
    if kind == "dups":
        # every k-mer is seen only in identical records -> hasMultipleUniqueReads stays 0
        # (except palindromic overlap between a read and its reverse complement)
        t = rnd(20)
        u = rnd(20)
        p = synth.records_from_reads([t] * 6 + [u] * 5, ["I" * 20] * 11)
        return p, np.zeros(1, np.uint8)
    if kind == "strand1":
        # identical sequences that differ only in the strand byte count as different reads (:350)
        t = rnd(20)
        recs = ("0" + t + "I" * 20) + ("1" + t + "I" * 20) + ("0" + t + "I" * 20)
        p = np.frombuffer(recs.encode() + b"\0", dtype=np.uint8).copy()
        return p, np.zeros(1, np.uint8)
    if kind == "saturate":
        # 5 overlapping reads x 9000 copies: the central k-mers occur 45000 times (> 32765)
        t = rnd(24)
        reads = [t[s:s + 20] for s in range(5)] * 9000
        quals = ["I" * 20, "?" * 20, "5" * 20] * 15000
        return synth.records_from_reads(reads[:30000], quals[:30000]), synth.records_from_reads(reads[30000:], quals[30000:])
    raise ValueError(kind)


def make_inputs(case: dict):
    if "hand" in case:
        return _hand(case["hand"])
    return synth.generate(read_length=case["L"], threads=2, **case["gen"])
