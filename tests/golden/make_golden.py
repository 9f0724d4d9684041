"""Regenerates tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref/libvdjref.so, built by
oracle/Makefile from /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py

Each fixture holds the case parameters, a sha256 of the generated input buffers (so generator
drift is detected), and every table the reference produced: the pruned pre_nodes
(first position, frequency, qual_sums) and the node pool in creation order (position, frequency,
ordered toNodes / fromNodes).  Hand-made cases store their input text too.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402
from tests.cases import CASES, make_inputs  # noqa: E402
from tests.util import GOLDEN_DIR, sha  # noqa: E402


def main():
    assert loader.have_reference(), "build oracle/_ref first: make -C oracle ref"
    for name, case in CASES.items():
        primary, secondary = make_inputs(case)
        ref = loader.build(primary, secondary, case["L"], case["k"], case["mf"], case["mq"], kind="reference")
        out = {k: v for k, v in ref.items() if isinstance(v, np.ndarray)}
        out["n_pre_total"] = np.uint64(ref["n_pre_total"])
        out["input_sha256"] = np.array(sha(primary, secondary))
        if case.get("store_input"):
            out["primary"] = primary
            out["secondary"] = secondary
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: nodes={ref['n_nodes']} pre={ref['n_pre']} pre_total={ref['n_pre_total']} "
              f"branching={(ref['out_deg'] > 1).sum()} -> {os.path.getsize(path)} B")


if __name__ == "__main__":
    main()
