"""Pins BASELINE.json's full-size configs bit-exactly: runs the COMPILED REFERENCE
(oracle/_ref/libvdjref.so, built by oracle/Makefile from /root/reference) once on the exact
workload bench.py times (vdjer_b200.synth.CONFIGS[name], seed 12345) and writes the counters and a
SHA-256 of every result array -- not the arrays -- to tests/golden/full_digests.json.

    python tests/golden/make_full_digest.py igh_2x50_5M [igh_sensitive_2x50_5M igk_2x75_20M ...]

Build container only (needs the compiled reference; minutes of one host core and a few GB per
config: igh_2x50_5M ~ 5 min).  tests/test_parity_gpu.py::test_full_size_matches_reference_digest
and bench.py compare the CUDA result with these digests.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402
from tests.util import GOLDEN_DIR, digest_of, sha  # noqa: E402
from vdjer_b200 import synth  # noqa: E402

OUT = os.path.join(GOLDEN_DIR, "full_digests.json")

def main():
    assert loader.have_reference(), "build oracle/_ref first: make -C oracle ref"
    names = sys.argv[1:] or ["igh_2x50_5M"]
    for name in names:
        wl = dict(synth.CONFIGS[name])
        L, k, mf, mq = wl["read_length"], wl["k"], wl["mf"], wl["mq"]
        gen = {kk: v for kk, v in wl.items() if kk not in ("k", "mf", "mq")}
        primary, secondary = synth.generate(seed=12345, **gen)
        t0 = time.perf_counter()
        ref = loader.build(primary, secondary, L, k, mf, mq, kind="reference")
        dt = time.perf_counter() - t0
        entry = {
            "workload": name, "seed": 12345, "read_length": L, "k": k, "mf": mf, "mq": mq, "n_pairs": wl["n_pairs"],
            "input_sha256": sha(primary, secondary),
            "n_records": ref["n_records"], "n_windows": ref["n_windows"], "n_gated": ref["n_gated"],
            "n_pre_total": ref["n_pre_total"], "n_pre": ref["n_pre"], "n_nodes": ref["n_nodes"], "n_hits": ref["n_hits"],
            "sha256": digest_of(lambda n: ref[n]),
            "reference_seconds": {"pass1": ref["t_pass1"], "prune": ref["t_prune"], "pass2": ref["t_pass2"], "wall": dt},
            "source": "oracle/_ref/libvdjref.so (the reference's own build_pre_graph/prune_pre_graph/build_graph2, -O2)",
        }
        all_ = json.load(open(OUT)) if os.path.exists(OUT) else {}
        all_[name] = entry
        with open(OUT, "w") as f:
            json.dump(all_, f, indent=1, sort_keys=True)
        print(f"{name}: nodes={ref['n_nodes']} pre_total={ref['n_pre_total']} hits={ref['n_hits']} "
              f"{dt:.0f} s (pass1 {ref['t_pass1']:.0f} prune {ref['t_prune']:.0f} pass2 {ref['t_pass2']:.0f})", flush=True)


if __name__ == "__main__":
    main()
