"""CPU only (build container): BASELINE.md section 3, item 3 -- the WHOLE reference binary
(oracle/_ref/vdjer_ref, built from /root/reference by oracle/build_e2e.py) on a synthetic BAM, with
--t 1 and --t $(nproc): the ELAPSED_SECS lines (status.c:25, one-second resolution) between
POST_READ_EXTRACT and POST_BUILD_GRAPH2 bracket the graph block (assembler2_vdj.c:1381-1415) and show
that --t does not change it (the block is single-threaded, --t only sizes the traversal pool).

    python profiles/whole_binary_time.py [clones] [pairs_per_clone] > profiles/whole_binary_r2.json
"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import e2e_data  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")


def main():
    clones = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    per = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
    out = {"input": f"tests/e2e_data.make_case: {clones} clones x {per} pairs, 2x50, --chain IGH, k=35 mf=3 mq=90", "runs": []}
    with tempfile.TemporaryDirectory() as work:
        sam, _ = e2e_data.make_case(work, n_clones=clones, pairs_per_clone=per)
        bam = os.path.join(work, "x.bam")
        subprocess.run([os.path.join(REF, "sam2bam"), sam, bam], check=True, capture_output=True)
        for t in (1, os.cpu_count() or 1):
            cwd = os.path.join(work, f"t{t}")
            os.makedirs(cwd)
            t0 = time.time()
            r = subprocess.run([os.path.join(REF, "vdjer_ref"), "--in", bam, "--t", str(t), "--ins", "175", "--chain", "IGH",
                                "--ref-dir", os.path.join(work, "ref")], cwd=cwd, capture_output=True, text=True)
            wall = time.time() - t0
            tags = {}
            for ln in r.stderr.splitlines():
                if ln.startswith("ELAPSED_SECS"):
                    f = ln.split("\t")
                    tags[f[1]] = int(f[2])
            nodes = [ln for ln in r.stderr.splitlines() if ln.startswith(("Num nodes", "pre nodes after", "Total nodes"))]
            out["runs"].append({"t": t, "rc": r.returncode, "wall_s": round(wall, 1),
                                "graph_block_s": tags.get("POST_BUILD_GRAPH2", 0) - tags.get("POST_READ_EXTRACT", 0),
                                "traversal_s": tags.get("THREADS_DONE", 0) - tags.get("POST_CONDENSE_GRAPH", 0),
                                "elapsed_secs": tags, "log": nodes[-2:]})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
