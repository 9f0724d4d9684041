"""GPU: super-partition rounds (vdjgraph_params.rounds) give exactly the one-round result.

For inputs whose tuples do not fit HBM (BASELINE configs[4]) the build walks the hash space in S
rounds: each round scatters only the windows of its 1/S of the hash partitions, runs pass 1, prune
and pass 2 on them, and appends the survivors as records; the finish then runs once over all
records.  Every per-k-mer quantity belongs to exactly one round, so nothing may change."""
import numpy as np
import pytest

from oracle import loader
from tests.util import assert_graph_equal
from vdjer_b200 import GraphBuilder, VdjGraphError, shard, synth

pytestmark = pytest.mark.gpu

FIELDS = ["first_pos", "frequency", "out_deg", "in_deg", "out_succ", "in_pred"]


@pytest.mark.parametrize("rounds", [2, 8, 64])
@pytest.mark.parametrize("L,k,mf,mq,pairs,clones,seed", [
    (50, 35, 3, 90, 40000, 800, 301),
    (50, 25, 1, 20, 20000, 150, 302),      # heavy branching: successors live in other rounds
    (100, 50, 2, 120, 10000, 300, 303),    # wide tuples
])
def test_rounds_equal_one_round_and_oracle(built, rounds, L, k, mf, mq, pairs, clones, seed):
    primary, secondary = synth.generate(n_pairs=pairs, read_length=L, seed=seed, n_clones=clones, threads=4)
    want = loader.build(primary, secondary, L, k, mf, mq, kind="port")
    with GraphBuilder(L, k, mf, mq, rounds=rounds, export_keys=True) as gb:
        got = gb.build(primary, secondary)
        again = gb.build(primary, secondary)          # buffers of the first build are reused
    assert got.stats["rounds"] == rounds and got.stats["partitions"] >= rounds
    assert_graph_equal(got, want, f"S={rounds}")
    for name in ["n_gated", "n_pre_total", "n_hits"]:
        assert got.stats[name] == want[name], name
    with GraphBuilder(L, k, mf, mq, rounds=1, export_keys=True) as gb:
        single = gb.build(primary, secondary)
    assert single.stats["rounds"] == 1
    for name in FIELDS + ["kmer_lo", "kmer_hi"]:
        assert np.array_equal(getattr(got, name), getattr(single, name)), name
        assert np.array_equal(getattr(again, name), getattr(single, name)), name
    assert got.stats["n_hits_ungated"] == single.stats["n_hits_ungated"]


def test_auto_rounds_follow_the_memory_budget(built, monkeypatch):
    """rounds = 0: one round when the working set fits the device, more when it does not (the budget
    is the device's memory; the test shrinks it)."""
    L, k, mf, mq = 50, 35, 3, 90
    primary, secondary = synth.generate(n_pairs=40000, read_length=L, seed=305, n_clones=500, threads=4)
    with GraphBuilder(L, k, mf, mq) as gb:
        one = gb.build(primary, secondary)
    assert one.stats["rounds"] == 1
    # the packed reads of 160 k records alone are ~15 MB, their runs ~11 MB, tables and graph ~15 MB
    monkeypatch.setenv("VDJGRAPH_MEM_BUDGET_MB", "20")
    with GraphBuilder(L, k, mf, mq) as gb:
        many = gb.build(primary, secondary)
    assert many.stats["rounds"] > 1
    for name in FIELDS:
        assert np.array_equal(getattr(many, name), getattr(one, name)), name


@pytest.mark.parametrize("G,rounds", [(2, 2), (4, 8), (8, 32)])
def test_rounds_in_a_sharded_build(built, G, rounds):
    """Devices x rounds: partition p belongs to device p mod G and round (p / G) mod S."""
    L, k, mf, mq = 50, 35, 2, 60
    primary, secondary = synth.generate(n_pairs=30000, read_length=L, seed=306, n_clones=400, threads=4)
    want = loader.build(primary, secondary, L, k, mf, mq, kind="port")
    rb = 2 * L + 1
    total = primary.size // rb + secondary.size // rb
    parts = [shard.split_records(primary, secondary, L, lo, hi) for lo, hi in shard.shard_ranges(total, G)]
    builders = [GraphBuilder(L, k, mf, mq, device=0, rounds=rounds) for _ in range(G)]
    try:
        got = shard.build_local(builders, parts)
        stats = [b.fetch_stats() for b in builders]
    finally:
        for b in builders:
            b.close()
    assert all(s["rounds"] == rounds for s in stats)
    assert_graph_equal(got, want, f"G={G} S={rounds}")
    assert sum(s["n_pre_total"] for s in stats) == want["n_pre_total"]
    assert sum(s["n_hits"] for s in stats) == want["n_hits"]


def test_rounds_parameter_errors(built):
    with pytest.raises(VdjGraphError) as e:
        GraphBuilder(50, 35, 3, 90, rounds=3)
    assert e.value.code == -1
    # rounds x devices may not exceed the 256 hash buckets of the window histogram
    L = 50
    primary, secondary = synth.generate(n_pairs=2000, read_length=L, seed=307, n_clones=20, threads=2)
    rb = 2 * L + 1
    total = primary.size // rb + secondary.size // rb
    parts = [shard.split_records(primary, secondary, L, lo, hi) for lo, hi in shard.shard_ranges(total, 2)]
    builders = [GraphBuilder(L, 35, 2, 60, device=0, rounds=256) for _ in range(2)]
    try:
        with pytest.raises(VdjGraphError) as e:
            shard.build_local(builders, parts)
        assert e.value.code == -1
    finally:
        for b in builders:
            b.close()
    # the pruned pass-1 table only exists as a whole in a one-round build
    with GraphBuilder(L, 35, 2, 60, rounds=4) as gb:
        gb.build(primary, secondary)
        with pytest.raises(VdjGraphError) as e:
            gb.pre_table()
        assert e.value.code == -7


def test_large_graph_finish_frees_the_dead_tables(built, monkeypatch):
    """A merged finish of a large graph gives the last round's tables back before it allocates its own
    buffers (configs[4] at full size needs that to fit one round); the threshold is lowered here so
    that a small build takes that path, twice on the same context."""
    L, k, mf, mq = 50, 35, 3, 90
    primary, secondary = synth.generate(n_pairs=30000, read_length=L, seed=306, n_clones=500, threads=4)
    want = loader.build(primary, secondary, L, k, mf, mq, kind="port")
    monkeypatch.setenv("VDJGRAPH_FREE_TABLES_MB", "0")
    with GraphBuilder(L, k, mf, mq, rounds=2) as gb:
        assert_graph_equal(gb.build(primary, secondary), want, "first build")
        assert_graph_equal(gb.build(primary, secondary), want, "second build (tables re-allocated)")
