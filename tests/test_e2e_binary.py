"""GPU, whole program: the V'DJer binary with its graph block replaced by libvdjgraph
(oracle/_ref/vdjer_gpu = the reference's sources + the INTEGRATION.md change, linked against
vdjer_b200/libvdjgraph.so) against the reference binary (oracle/_ref/vdjer_ref) on the same BAM,
same CLI: byte-identical vdjer.dot, vdj_contigs.fa and SAM output, same ROOT_INIT lines.

Both binaries carry the same two test-only fixes (missing `return`s that g++ 13 turns into traps;
the worker-exit race that makes the reference drop roots at random); see
oracle/build_e2e.py.  Input: tests/e2e_data.py (the reference's demo BAM is not in its tree)."""
import os
import re
import subprocess

import pytest

from tests import e2e_data

REF_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
BIN = {n: os.path.join(REF_DIR, n) for n in ("vdjer_ref", "vdjer_gpu", "sam2bam")}
needs_bins = pytest.mark.skipif(not all(os.path.exists(p) for p in BIN.values()),
                                reason="oracle/_ref/vdjer_ref, vdjer_gpu, sam2bam not built (no /root/reference)")


def _run(binary, bam, ref, cwd, extra, env=None):
    os.makedirs(cwd, exist_ok=True)
    r = subprocess.run([binary, "--in", bam, "--t", "1", "--ins", "175", "--chain", "IGH", "--ref-dir", ref] + extra,
                       cwd=cwd, capture_output=True, text=True, timeout=900, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stderr[-3000:]
    keep = [ln for ln in r.stderr.splitlines()
            if re.match(r"(ROOT_INIT|num root nodes|pre nodes after pruning|Total nodes|Num traversable|Num condensed)", ln)]
    return dict(dot=open(os.path.join(cwd, "vdjer.dot"), "rb").read(),
                contigs=open(os.path.join(cwd, "vdj_contigs.fa"), "rb").read(), sam=r.stdout, log=keep, err=r.stderr)


@needs_bins
@pytest.mark.gpu
@pytest.mark.parametrize("extra,clones,expect_contigs", [
    ([], 4, True),                                            # default flags: k=35 mf=3 mq=90
    (["--mq", "60", "--mf", "2"], 6, True),
    (["--k", "25", "--mq", "60", "--mf", "2"], 6, False),     # README's sensitive mode (no window passes the V(D)J filter here)
])
def test_vdjer_binary_with_gpu_graph_matches_reference_binary(built, tmp_path, extra, clones, expect_contigs):
    work = str(tmp_path)
    sam, _ = e2e_data.make_case(work, n_clones=clones)
    bam = os.path.join(work, "x.bam")
    subprocess.run([BIN["sam2bam"], sam, bam], check=True, capture_output=True)
    ref = _run(BIN["vdjer_ref"], bam, os.path.join(work, "ref"), os.path.join(work, "cpu"), extra)
    gpu = _run(BIN["vdjer_gpu"], bam, os.path.join(work, "ref"), os.path.join(work, "gpu"), extra)
    assert gpu["log"] == ref["log"], "\n".join(gpu["log"][:5] + ref["log"][:5])
    assert gpu["dot"] == ref["dot"], "vdjer.dot differs"
    assert gpu["contigs"] == ref["contigs"], "vdj_contigs.fa differs"
    assert gpu["sam"] == ref["sam"], "SAM output differs"
    if expect_contigs:   # the pipeline really produced contigs and mapped reads to them
        assert len(ref["contigs"]) > 100 and ref["sam"].count("\n") > 100
    assert "vdjgraph" not in ref["err"]
    # vdjer_gpu took the forward-reads path (add_to_buffer keeps every read once,
    # the reverse-complement records are derived on the device); the text path gives the same files
    assert "vdjgraph: forward reads only" in gpu["err"]
    txt = _run(BIN["vdjer_gpu"], bam, os.path.join(work, "ref"), os.path.join(work, "gpu_text"), extra, env={"VDJGRAPH_FORWARD": "0"})
    assert "vdjgraph: forward reads only" not in txt["err"]
    assert txt["log"] == ref["log"] and txt["dot"] == ref["dot"] and txt["contigs"] == ref["contigs"] and txt["sam"] == ref["sam"]


@needs_bins
@pytest.mark.gpu
def test_vdjer_binary_over_several_devices_matches_reference_binary(built, tmp_path):
    """VDJGRAPH_DEVICES: the glue builds the graph through vdjgraph_multi_* (here: this GPU named twice, one host
    thread each); vdjer.dot, contigs and SAM stay byte-identical."""
    work = str(tmp_path)
    sam, _ = e2e_data.make_case(work, n_clones=4)
    bam = os.path.join(work, "x.bam")
    subprocess.run([BIN["sam2bam"], sam, bam], check=True, capture_output=True)
    ref = _run(BIN["vdjer_ref"], bam, os.path.join(work, "ref"), os.path.join(work, "cpu"), [])
    gpu = _run(BIN["vdjer_gpu"], bam, os.path.join(work, "ref"), os.path.join(work, "gpu2"), [], env={"VDJGRAPH_DEVICES": "0,0"})
    assert "vdjgraph: one graph over 2 devices" in gpu["err"]
    assert gpu["log"] == ref["log"] and gpu["dot"] == ref["dot"] and gpu["contigs"] == ref["contigs"] and gpu["sam"] == ref["sam"]
    assert len(ref["contigs"]) > 100


@needs_bins
def test_reference_binary_runs_on_the_synthetic_bam(built, tmp_path):
    """CPU: the stand-in for BASELINE configs[0] drives the unmodified pipeline to contigs."""
    work = str(tmp_path)
    sam, _ = e2e_data.make_case(work)
    bam = os.path.join(work, "x.bam")
    subprocess.run([BIN["sam2bam"], sam, bam], check=True, capture_output=True)
    ref = _run(BIN["vdjer_ref"], bam, os.path.join(work, "ref"), os.path.join(work, "cpu"), [])
    assert ref["contigs"].count(b">") >= 2 and any(ln.startswith("ROOT_INIT") for ln in ref["log"])
