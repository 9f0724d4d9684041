"""GPU: the sharded build gives exactly the single-device result.

G ranks run as G contexts inside this process on ONE GPU (peer buffers are plain device pointers;
across processes the same phases run over CUDA IPC mappings, see vdjer_b200/shard.py and
tests/test_shard_dist.py): every tuple is written by the scatter kernel into the buffer of the
rank that owns its hash partition, reads are compared and quality rows fetched across ranks, and
rank 0 finishes over the gathered survivor records."""
import numpy as np
import pytest

from oracle import loader
from tests.util import assert_graph_equal
from vdjer_b200 import GraphBuilder, shard, synth

pytestmark = pytest.mark.gpu


def _sharded(primary, secondary, L, k, mf, mq, G, **kw):
    rb = 2 * L + 1
    total = (primary.size // rb) + (secondary.size // rb)
    parts = [shard.split_records(primary, secondary, L, lo, hi) for lo, hi in shard.shard_ranges(total, G)]
    builders = [GraphBuilder(L, k, mf, mq, device=0, **kw) for _ in range(G)]
    try:
        g = shard.build_local(builders, parts)
        stats = [b.fetch_stats() for b in builders]
    finally:
        for b in builders:
            b.close()
    return g, stats


@pytest.mark.parametrize("G", [2, 4, 8])
@pytest.mark.parametrize("L,k,mf,mq,pairs,clones,seed", [
    (50, 35, 3, 90, 40000, 800, 201),
    (50, 25, 1, 20, 20000, 150, 202),      # heavy branching: edge order across owners
    (100, 50, 2, 120, 10000, 300, 203),    # wide tuples
])
def test_sharded_equals_single_device_and_oracle(built, G, L, k, mf, mq, pairs, clones, seed):
    primary, secondary = synth.generate(n_pairs=pairs, read_length=L, seed=seed, n_clones=clones, threads=4)
    want = loader.build(primary, secondary, L, k, mf, mq, kind="port")
    got, stats = _sharded(primary, secondary, L, k, mf, mq, G)
    assert_graph_equal(got, want, f"G={G}")
    assert sum(s["n_pre_total"] for s in stats) == want["n_pre_total"]      # k-mers are owned by exactly one rank
    assert sum(s["n_gated"] for s in stats) == want["n_gated"]
    assert sum(s["n_hits"] for s in stats) == want["n_hits"]
    with GraphBuilder(L, k, mf, mq) as gb:
        single = gb.build(primary, secondary)
    for name in ["first_pos", "frequency", "out_deg", "in_deg", "out_succ", "in_pred"]:
        assert np.array_equal(getattr(got, name), getattr(single, name)), name


@pytest.mark.parametrize("G", [2, 8])
def test_sharded_duplicates_and_strand_across_ranks(built, G):
    """Identical reads that land on different ranks must still count as ONE read for
    hasMultipleUniqueReads (:349-352): the exact comparison reads the peer's packed reads."""
    L, k = 50, 35
    primary, secondary = synth.generate(n_pairs=6000, read_length=L, seed=205, n_clones=60, threads=4)
    rb = 2 * L + 1
    body = primary[:-1].reshape(-1, rb)
    primary = np.concatenate([body, body, body[::-1]]).reshape(-1)      # every record three times, far apart
    primary = np.concatenate([primary, np.zeros(1, np.uint8)])
    want = loader.build(primary, secondary, L, k, 2, 60, kind="port")
    got, _ = _sharded(primary, secondary, L, k, 2, 60, G)
    assert_graph_equal(got, want, f"dups G={G}")


def test_sharded_empty_rank(built):
    """A rank without records (fewer records than ranks) takes part in every phase."""
    L, k = 50, 35
    primary, secondary = synth.generate(n_pairs=3000, read_length=L, seed=206, n_clones=30, threads=2)
    want = loader.build(primary, secondary, L, k, 2, 60, kind="port")
    rb = 2 * L + 1
    total = primary.size // rb + secondary.size // rb
    parts = [shard.split_records(primary, secondary, L, 0, total)] + [(np.zeros(0, np.uint8), np.zeros(0, np.uint8))] * 3
    builders = [GraphBuilder(L, k, 2, 60, device=0) for _ in range(4)]
    try:
        got = shard.build_local(builders, parts)
    finally:
        for b in builders:
            b.close()
    assert_graph_equal(got, want, "empty ranks")


def test_multi_process_parity_over_nvlink(built):
    """One process per GPU (torchrun), peer buffers mapped with CUDA IPC, tuples written over NVLink
    by the scatter kernel: bit-exact against the oracle (tests/mgpu_parity.py).  Needs >= 2 GPUs."""
    import os
    import subprocess
    import sys

    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs on the box")
    world = 8 if n >= 8 else 4 if n >= 4 else 2
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tests", "mgpu_parity.py")],
                       capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("PARITY OK") == 3, r.stdout[-2000:]


@pytest.mark.parametrize("G", [2, 4])
def test_sharded_from_forward_reads_only(built, G):
    """vdjgraph_shard_stage_forward: every rank stages forward reads only, the reverse-complement
    records are derived on its device; record numbers stay those of the doubled buffers."""
    from vdjer_b200 import forward_reads
    L, k, mf, mq = 50, 35, 3, 90
    primary, secondary = synth.generate(n_pairs=40000, read_length=L, seed=207, n_clones=600, threads=4)
    want = loader.build(primary, secondary, L, k, mf, mq, kind="port")
    rb = 2 * L + 1
    total = primary.size // rb + secondary.size // rb
    parts = []
    for lo, hi in shard.shard_ranges(total, G):          # even boundaries: a read stays with its reverse complement
        p, s = shard.split_records(primary, secondary, L, lo, hi)
        parts.append((forward_reads(p, L), forward_reads(s, L)))
    builders = [GraphBuilder(L, k, mf, mq, device=0) for _ in range(G)]
    try:
        got = shard.build_local(builders, parts, forward=True)
    finally:
        for b in builders:
            b.close()
    assert_graph_equal(got, want, f"forward G={G}")
