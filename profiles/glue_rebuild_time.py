"""CPU measurement (no GPU needed): what the reference-side glue costs after the library has
returned, i.e. vdjgraph_rebuild_nodes of glue/vdjgraph_glue.inc (node pool through the reference's
new_node, sparsehash `nodes` map in the reference's insertion order, toNodes / fromNodes lists).
Arrays come from the oracle (same layout as vdjgraph_result).

    python profiles/glue_rebuild_time.py [pairs]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import loader  # noqa: E402
from vdjer_b200 import synth  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
wl = dict(synth.CONFIGS["igh_2x50_5M"])
L, k, mf, mq = wl["read_length"], wl["k"], wl["mf"], wl["mq"]
gen = {kk: v for kk, v in wl.items() if kk not in ("k", "mf", "mq", "n_pairs")}
primary, secondary = synth.generate(n_pairs=pairs, seed=12345, **gen)
t0 = time.perf_counter()
g = loader.build(primary, secondary, L, k, mf, mq, kind="port")
t_oracle = time.perf_counter() - t0
ms = min(loader.glue_rebuild_ms(primary, secondary, L, k, g) for _ in range(3))
print(json.dumps({"workload": "igh_2x50_5M generator", "pairs": pairs, "windows": g["n_windows"], "nodes": g["n_nodes"],
                  "glue_rebuild_ms": round(ms, 2), "ns_per_node": round(ms * 1e6 / max(1, g["n_nodes"]), 1),
                  "oracle_build_s": round(t_oracle, 2)}))
