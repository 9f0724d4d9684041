"""GPU: the CUDA path (through the C ABI) against the oracle and the committed reference outputs.
Bit-exact: integer/byte/index work only."""
import os

import numpy as np
import pytest

from oracle import loader
from tests.cases import CASES, make_inputs
from tests.util import GOLDEN_DIR, assert_graph_equal, assert_pre_table_equal, digest_of, graph_invariants, sha
from vdjer_b200 import GraphBuilder, VdjGraphError, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(CASES))
def test_matches_reference_golden(built, name):
    """CUDA result == what the compiled reference produced (tests/golden), every field."""
    case = CASES[name]
    gold = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    gold["n_nodes"] = len(gold["first_pos"])
    gold["n_pre"] = len(gold["pre_first_pos"])
    primary, secondary = make_inputs(case)
    with GraphBuilder(case["L"], case["k"], case["mf"], case["mq"], export_keys=True) as gb:
        got = gb.build(primary, secondary)
        pre = gb.pre_table()
    assert got.stats["n_pre_total"] == int(gold["n_pre_total"])
    assert_pre_table_equal(pre, gold, primary, secondary, case["L"], case["k"], name)
    assert_graph_equal(got, gold, name)


@pytest.mark.parametrize("seed,L,k,mf,mq,pairs,clones", [
    (31, 50, 35, 3, 90, 60000, 2000),     # configs[1] shape
    (32, 50, 25, 2, 60, 60000, 20000),    # configs[2] shape, flat repertoire
    (33, 75, 35, 3, 90, 40000, 1500),     # configs[3] shape
    (34, 100, 35, 3, 90, 30000, 1500),    # configs[4] shape
    (35, 100, 50, 2, 120, 20000, 500),
    (36, 50, 25, 1, 20, 40000, 200),      # heavy branching
    (37, 150, 41, 3, 90, 10000, 300),
    (38, 255, 50, 3, 90, 4000, 100),      # longest read the reference stores
    (39, 20, 3, 2, 40, 500, 5),           # k below the minimizer length: the k-mer is its own minimizer
    (40, 30, 1, 2, 40, 300, 5),           # k = 1: four possible nodes
    (41, 64, 10, 3, 90, 800, 20),         # k = m: one m-mer per window
    (42, 64, 11, 3, 90, 800, 20),         # k = m + 1: two
    (43, 40, 9, 2, 60, 800, 20),
])
def test_matches_oracle_seeded(built, seed, L, k, mf, mq, pairs, clones):
    primary, secondary = synth.generate(n_pairs=pairs, read_length=L, seed=seed, n_clones=clones, threads=4)
    want = loader.build(primary, secondary, L, k, mf, mq, kind="port")
    with GraphBuilder(L, k, mf, mq) as gb:
        got = gb.build(primary, secondary)
        pre = gb.pre_table()
    assert got.stats["n_records"] == want["n_records"] and got.stats["n_windows"] == want["n_windows"]
    assert got.stats["n_gated"] == want["n_gated"]
    assert got.stats["n_pre_total"] == want["n_pre_total"]
    assert got.stats["n_hits"] == want["n_hits"]
    assert_pre_table_equal(pre, want, primary, secondary, L, k, f"seed{seed}")
    assert_graph_equal(got, want, f"seed{seed}")


@pytest.mark.parametrize("partitions,wide", [(1, False), (2, False), (16, True), (256, False), (64, True)])
@pytest.mark.parametrize("L,k,mf,mq", [(50, 35, 3, 90), (100, 50, 2, 120), (50, 25, 1, 20)])
def test_partition_count_and_tuple_format_do_not_change_the_result(built, partitions, wide, L, k, mf, mq):
    """The hash-partition count and the 16/24-byte tuple format are performance knobs only."""
    primary, secondary = synth.generate(n_pairs=15000, read_length=L, seed=61 + L + k, n_clones=300, threads=4)
    want = loader.build(primary, secondary, L, k, mf, mq, kind="port")
    with GraphBuilder(L, k, mf, mq, partitions=partitions, wide_tuples=wide) as gb:
        got = gb.build(primary, secondary)
        pre = gb.pre_table()
    # narrow tuples need room for the stamp and a few read-fingerprint bits; k=50 leaves none
    assert got.stats["partitions"] == partitions and got.stats["tuple_bytes"] == (24 if wide or k == 50 else 16)
    assert_pre_table_equal(pre, want, primary, secondary, L, k, f"P{partitions}")
    assert_graph_equal(got, want, f"P{partitions} wide{wide}")
    assert got.stats["n_hits"] == want["n_hits"] and got.stats["n_gated"] == want["n_gated"]


@pytest.mark.parametrize("knobs", [
    {"VDJGRAPH_HOT_T": "1", "VDJGRAPH_HOT_FLUSH": "1"},           # every fast-path increment goes through the warp cache
    {"VDJGRAPH_HOT_T": "2000000000"},                               # ... none does
    {"VDJGRAPH_LOAD1": "0.85", "VDJGRAPH_LOAD2": "0.85"},           # long probe chains in both tables
    {"VDJGRAPH_SLICE_MB": "0.05"},                                  # as many partitions as there can be
    {"VDJGRAPH_QFLUSH1": "32", "VDJGRAPH_QDENSE1": "16"},           # pass-1 drains re-queue stragglers
    {"VDJGRAPH_QFLUSH2": "1", "VDJGRAPH_QDENSE2": "0"},             # pass-2 drains after every batch, to the end
    {"VDJGRAPH_QFLUSH2": "96", "VDJGRAPH_QDENSE2": "31"},           # ... or give up early and re-queue
])
def test_performance_knobs_do_not_change_the_result(built, knobs, monkeypatch):
    """The environment knobs of DESIGN.md section 4 only move work around."""
    for k_, v in knobs.items():
        monkeypatch.setenv(k_, v)
    for (L, k, mf, mq, seed) in [(50, 35, 3, 90, 81), (50, 25, 1, 20, 82)]:
        primary, secondary = synth.generate(n_pairs=20000, read_length=L, seed=seed, n_clones=200, threads=4)
        want = loader.build(primary, secondary, L, k, mf, mq, kind="port")
        with GraphBuilder(L, k, mf, mq) as gb:
            got = gb.build(primary, secondary)
            pre = gb.pre_table()
        assert_pre_table_equal(pre, want, primary, secondary, L, k, str(knobs))
        assert_graph_equal(got, want, str(knobs))
        assert got.stats["n_hits"] == want["n_hits"]


@pytest.mark.parametrize("fp_bits", ["0", "2"])
def test_weak_read_fingerprints_fall_back_to_the_exact_comparison(built, fp_bits, monkeypatch):
    """hasMultipleUniqueReads (:349-352) is decided by a read fingerprint carried in the tuples and,
    when fingerprints are equal, by comparing the packed reads.  With 0 or 2 fingerprint bits almost
    every decision takes the exact path; the result must not change.  Duplicated reads included."""
    monkeypatch.setenv("VDJGRAPH_FP_BITS", fp_bits)
    primary, secondary = synth.generate(n_pairs=12000, read_length=50, seed=77, n_clones=150, threads=4)
    rl = 2 * 50 + 1
    body = primary[:-1].reshape(-1, rl)
    primary = np.concatenate([body, body[:4000], body[1000:2000]]).reshape(-1)   # duplicated records
    primary = np.concatenate([primary, np.zeros(1, np.uint8)])
    want = loader.build(primary, secondary, 50, 35, 2, 60, kind="port")
    with GraphBuilder(50, 35, 2, 60) as gb:
        got = gb.build(primary, secondary)
        pre = gb.pre_table()
    assert_pre_table_equal(pre, want, primary, secondary, 50, 35, f"fp{fp_bits}")
    assert_graph_equal(got, want, f"fp{fp_bits}")


def test_context_reuse_and_param_changes(built):
    """One context, several builds with different --k/--mf/--mq and inputs: no state leaks."""
    primary, secondary = synth.generate(n_pairs=8000, read_length=50, seed=41, n_clones=100, threads=4)
    p2, s2 = synth.generate(n_pairs=3000, read_length=50, seed=42, n_clones=30, threads=4)
    with GraphBuilder(50, 35, 3, 90) as gb:
        for (pp, ss, k, mf, mq) in [(primary, secondary, 35, 3, 90), (p2, s2, 25, 2, 60), (primary, secondary, 35, 3, 90),
                                     (primary, secondary, 35, 1, 214), (p2, s2, 35, 3, 90)]:
            gb.set_params(k=k, mf=mf, mq=mq)
            got = gb.build(pp, ss)
            want = loader.build(pp, ss, 50, k, mf, mq, kind="port")
            assert_graph_equal(got, want, f"k{k} mf{mf} mq{mq}")
        # stage once, run with different pruning flags without re-staging
        gb.set_params(k=35, mf=3, mq=90)
        gb.stage(primary, secondary)
        for mf, mq in [(3, 90), (2, 60), (5, 150)]:
            gb.set_params(mf=mf, mq=mq)
            gb.run()
            assert_graph_equal(gb.fetch(), loader.build(primary, secondary, 50, 35, mf, mq, kind="port"), f"rerun mf{mf}")


def test_tiny_table_capacity_grows(built):
    """An undersized table_capacity hint must be recovered from (the reference's dense_hash_map
    grows by rehashing, internal/densehashtable.h:631-653)."""
    primary, secondary = synth.generate(n_pairs=5000, read_length=50, seed=43, n_clones=100, threads=4)
    want = loader.build(primary, secondary, 50, 35, 3, 90, kind="port")
    with GraphBuilder(50, 35, 3, 90, table_capacity=1024) as gb:
        got = gb.build(primary, secondary)
    assert got.stats["table1_slots"] > want["n_pre_total"]
    assert_graph_equal(got, want, "grown")


def test_low_redundancy_input_fills_the_occurrence_log(built):
    """Random, non-overlapping reads: millions of distinct k-mers that each occur once (per strand), so
    every gated window claims its own occurrence-log block -- the log, not the table, is the
    structure under pressure (the log used to be sized for clonal data only)."""
    rng = np.random.default_rng(77)
    n, L, k = 350_000, 50, 35
    codes = rng.integers(0, 4, (n, L), dtype=np.uint8)
    fwd = np.frombuffer(b"ACGT", np.uint8)[codes]
    rc = np.frombuffer(b"TGCA", np.uint8)[codes[:, ::-1]]
    rec = np.empty((n, 2, 2 * L + 1), np.uint8)
    rec[:, :, 0] = ord("0")
    rec[:, 0, 1:1 + L] = fwd
    rec[:, 1, 1:1 + L] = rc
    rec[:, :, 1 + L:] = ord("I")
    primary = np.concatenate([rec.reshape(-1), np.zeros(1, np.uint8)])
    want = loader.build(primary, np.zeros(1, np.uint8), L, k, 1, 40, kind="port")
    assert want["n_pre_total"] > 5_000_000
    with GraphBuilder(L, k, 1, 40) as gb:
        got = gb.build(primary, b"")
    assert got.stats["n_pre_total"] == want["n_pre_total"]
    assert_graph_equal(got, want, "low redundancy")


def test_empty_and_degenerate_inputs(built):
    with GraphBuilder(50, 35, 3, 90) as gb:
        g = gb.build(b"", b"")
        assert g.n_nodes == 0 and g.stats["n_windows"] == 0
        g = gb.build(np.zeros(1, np.uint8), np.zeros(1, np.uint8))
        assert g.n_nodes == 0
        one = synth.records_from_reads(["ACGTTGCAAC" * 5], both_strands=False)
        g = gb.build(one, b"")
        # period-10 sequence: 16 gated windows, 10 distinct k-mers, one read only -> nothing survives
        assert g.n_nodes == 0 and g.stats["n_pre_total"] == 10 and g.stats["n_gated"] == 16
        alln = synth.records_from_reads(["N" * 50] * 7)
        g = gb.build(alln, alln)
        assert g.n_nodes == 0 and g.stats["n_gated"] == 0 and g.stats["n_pre_total"] == 0


def test_input_errors_are_reported_not_fatal(built):
    good = synth.records_from_reads(["ACGTTGCAAC" * 5] * 4)
    with GraphBuilder(50, 35, 3, 90) as gb:
        bad = good.copy(); bad[101 * 3] = ord("2")           # strand byte (reference: exit(-1), :383-391)
        with pytest.raises(VdjGraphError) as e:
            gb.build(bad, b"")
        assert e.value.code == -2 and "record 3" in str(e.value)
        bad = good.copy(); bad[101 * 5 + 7] = ord("M")        # IUPAC base
        with pytest.raises(VdjGraphError) as e:
            gb.build(b"", bad)
        assert e.value.code == -3 and "record 5" in str(e.value)
        with pytest.raises(VdjGraphError) as e:
            gb.fetch()
        assert e.value.code == -7
        # the context stays usable
        assert gb.build(good, b"").stats["n_windows"] == 8 * 16


def test_large_invariants_and_cross_check(built):
    """A mid-size run (1.3 M records): size-independent properties, plus agreement with the
    oracle on the counters and the whole graph."""
    L, k = 50, 35
    primary, secondary = synth.generate(n_pairs=320000, read_length=L, seed=51, n_clones=20000, threads=8)
    with GraphBuilder(L, k, 3, 90) as gb:
        got = gb.build(primary, secondary)
    graph_invariants(got, L, k, primary, secondary)
    want = loader.build(primary, secondary, L, k, 3, 90, kind="port")
    assert_graph_equal(got, want, "mid-size")
    assert got.stats["n_hits"] == want["n_hits"]


def _digest(g):
    return sha(g.first_pos, g.frequency, g.out_deg, g.in_deg, g.out_succ, g.in_pred)


def _full_digests():
    import json
    path = os.path.join(GOLDEN_DIR, "full_digests.json")
    return json.load(open(path)) if os.path.exists(path) else {}


@pytest.mark.parametrize("workload", sorted(_full_digests()))
def test_full_size_matches_reference_digest(built, workload):
    """BASELINE.json configs at FULL size, bit for bit: the compiled reference ran once on exactly
    this workload (tests/golden/make_full_digest.py, minutes of CPU) and left the counters and a
    SHA-256 of every result array in tests/golden/full_digests.json; the CUDA result must hash the
    same.  This is the workload bench.py times."""
    want = _full_digests()[workload]
    wl = dict(synth.CONFIGS[workload])
    L, k, mf, mq = wl["read_length"], wl["k"], wl["mf"], wl["mq"]
    gen = {kk: v for kk, v in wl.items() if kk not in ("k", "mf", "mq")}
    primary, secondary = synth.generate(seed=want["seed"], **gen)
    assert sha(primary, secondary) == want["input_sha256"], "the generator no longer produces the pinned input"
    with GraphBuilder(L, k, mf, mq) as gb:
        got = gb.build(primary, secondary, copy=False)
        for name in ["n_records", "n_windows", "n_pre_total", "n_pre", "n_nodes"]:
            assert got.stats[name] == want[name], name
        assert digest_of(lambda n: getattr(got, n)) == want["sha256"]


@pytest.mark.parametrize("workload", ["igh_2x50_5M", "igh_sensitive_2x50_5M", "igk_2x75_20M"])
def test_full_size_baseline_configs_properties(built, workload):
    """BASELINE.json configs[1..3] at FULL size (the oracle would need minutes there), through
    size-independent properties:
      * structural invariants of any correct graph (creation order, edge symmetry, k-1 overlap, ...);
      * determinism and independence from the performance knobs (partition count);
      * monotonicity against the oracle on a PREFIX of the records: counts, the multi-read flag and
        the quality sums only grow with more records and the first occurrence of a k-mer does not
        move, so every node of the prefix graph is a node of the full graph with the same first_pos,
        a frequency at least as large, in the same relative creation order, and every prefix edge
        is an edge of the full graph."""
    from tests.util import sha  # noqa: F401
    wl = dict(synth.CONFIGS[workload])
    L, k, mf, mq = wl["read_length"], wl["k"], wl["mf"], wl["mq"]
    gen = {kk: v for kk, v in wl.items() if kk not in ("k", "mf", "mq")}
    primary, secondary = synth.generate(seed=12345, **gen)
    rb = 2 * L + 1
    with GraphBuilder(L, k, mf, mq) as gb:
        full = gb.build(primary, secondary)
        stats = dict(full.stats)
        W = stats["n_windows"]
        assert W == (primary.size // rb + secondary.size // rb) * (L - k + 1)
        assert 0 < stats["n_gated"] <= W and stats["n_hits"] <= W
        assert stats["n_nodes"] == stats["n_pre"] <= stats["n_pre_total"] <= stats["n_gated"]
        # frequency is the capped count of N-free occurrences
        assert int(full.frequency.astype(np.int64).sum()) <= stats["n_hits"]
        assert int(full.frequency.max()) <= 32765
        graph_invariants(full, L, k, primary, secondary)
        d0 = _digest(full)
        again = gb.build(primary, secondary)
        assert _digest(again) == d0, "two builds of the same input differ"
    with GraphBuilder(L, k, mf, mq, partitions=max(1, stats["partitions"] // 4)) as gb:
        other = gb.build(primary, secondary)
        assert other.stats["partitions"] != stats["partitions"]
        assert _digest(other) == d0, "result depends on the partition count"
    # ... and from the number of super-partition rounds (the route for inputs whose tuples exceed HBM)
    with GraphBuilder(L, k, mf, mq, rounds=4) as gb:
        rounds4 = gb.build(primary, secondary)
        assert rounds4.stats["rounds"] == 4
        assert _digest(rounds4) == d0, "result depends on the number of rounds"
        for name in ["n_gated", "n_pre_total", "n_hits", "n_hits_ungated"]:
            assert rounds4.stats[name] == stats[name], name
    # prefix monotonicity against the oracle: the first 150k primary records
    n_pre_rec = min(150_000, (primary.size // rb) & ~1)
    prefix = np.concatenate([primary[: n_pre_rec * rb], np.zeros(1, np.uint8)])
    want = loader.build(prefix, np.zeros(1, np.uint8), L, k, mf, mq, kind="port")
    pos = np.searchsorted(full.first_pos, want["first_pos"])
    assert np.all(pos < full.n_nodes) and np.array_equal(full.first_pos[pos], want["first_pos"]), \
        "a node of the prefix graph is missing from the full graph"
    assert np.all(np.diff(pos.astype(np.int64)) > 0)
    assert np.all(full.frequency[pos] >= want["frequency"])
    # prefix edges survive: successor sets (mapped to full node numbers) are subsets
    nil = 0xFFFFFFFF
    sel = np.arange(0, want["n_nodes"], max(1, want["n_nodes"] // 200000))
    for j in range(4):
        s = want["out_succ"][sel, j]
        m = s != nil
        tgt = pos[s[m]]
        src = pos[sel[m]]
        assert np.all((full.out_succ[src] == tgt[:, None].astype(np.uint32)).any(axis=1)), "a prefix edge is missing"


def test_page_locked_record_buffers_are_staged_without_bounce_copies(built):
    """Record buffers from vdjgraph_host_alloc / registered with vdjgraph_host_register are DMAed
    straight to the device; pageable ones are bounced by the staging threads.
    Same graph either way, also when the chunk size does not divide the record counts and a chunk
    spans the primary/secondary boundary."""
    from vdjer_b200 import PinnedRecords, host_alloc, host_free
    L, k, mf, mq = 50, 35, 3, 90
    primary, secondary = synth.generate(n_pairs=150000, read_length=L, seed=91, n_clones=1500, threads=4)
    want = loader.build(primary, secondary, L, k, mf, mq, kind="port")
    with GraphBuilder(L, k, mf, mq) as gb:
        pageable = gb.build(primary, secondary)
        assert_graph_equal(pageable, want, "pageable")
        with PinnedRecords(primary, secondary):
            registered = gb.build(primary, secondary)
        assert_graph_equal(registered, want, "registered")
        p2, s2 = host_alloc(primary.size), host_alloc(secondary.size)
        try:
            p2[:] = primary
            s2[:] = secondary
            allocated = gb.build(p2, s2)
            assert_graph_equal(allocated, want, "host_alloc")
            mixed = gb.build(p2, secondary)          # one pageable buffer: the bounce path takes both
            assert_graph_equal(mixed, want, "mixed")
            only_secondary = gb.build(np.zeros(1, np.uint8), s2)
        finally:
            host_free(p2)
            host_free(s2)
        want_s = loader.build(np.zeros(1, np.uint8), secondary, L, k, mf, mq, kind="port")
        assert_graph_equal(only_secondary, want_s, "secondary only")
        for g in (pageable, registered, allocated, mixed):
            assert g.stats["h2d_bytes"] == (primary.size - 1) + (secondary.size - 1)


@pytest.mark.parametrize("L,k,mf,mq,pairs,clones,seed", [
    (50, 35, 3, 90, 150000, 1500, 92),
    (100, 50, 2, 120, 20000, 400, 93),     # wide tuples, two mask words
    (75, 25, 1, 20, 30000, 200, 94),
])
def test_forward_reads_only_give_the_graph_of_the_doubled_buffers(built, L, k, mf, mq, pairs, clones, seed):
    """vdjgraph_build_forward: from the forward reads alone the device derives every
    reverse-complement record (bam_read.c:230-243) and builds the graph of the reference's doubled
    buffers: same nodes, same positions in the DOUBLED numbering, same edges."""
    from vdjer_b200 import PinnedRecords, forward_reads
    primary, secondary = synth.generate(n_pairs=pairs, read_length=L, seed=seed, n_clones=clones, threads=4)
    want = loader.build(primary, secondary, L, k, mf, mq, kind="port")
    fp, fs = forward_reads(primary, L), forward_reads(secondary, L)
    assert fp.size - 1 == (primary.size - 1) // 2
    with GraphBuilder(L, k, mf, mq, export_keys=True) as gb:
        text = gb.build(primary, secondary)
        got = gb.build_forward(fp, fs)
        pre = gb.pre_table()
        assert_graph_equal(got, want, "forward")
        assert_pre_table_equal(pre, want, primary, secondary, L, k, "forward")
        for name in ["first_pos", "frequency", "out_deg", "in_deg", "out_succ", "in_pred", "kmer_lo", "kmer_hi"]:
            assert np.array_equal(getattr(got, name), getattr(text, name)), name
        for name in ["n_records", "n_windows", "n_gated", "n_pre_total", "n_hits"]:
            assert got.stats[name] == text.stats[name], name
        assert got.stats["h2d_bytes"] * 2 == text.stats["h2d_bytes"]
        with PinnedRecords(fp, fs):                      # page-locked: the direct-DMA path
            pinned = gb.build_forward(fp, fs)
        assert_graph_equal(pinned, want, "forward, page-locked")
        only_primary = gb.build_forward(fp)
    want_p = loader.build(primary, np.zeros(1, np.uint8), L, k, mf, mq, kind="port")
    assert_graph_equal(only_primary, want_p, "forward, primary only")


def test_forward_reads_errors(built):
    """A bad base or strand byte in a forward read is reported with its record number in the doubled
    numbering; N reads stay N in the derived record."""
    from vdjer_b200 import forward_reads
    L, k = 50, 35
    primary, secondary = synth.generate(n_pairs=4000, read_length=L, seed=95, n_clones=40, threads=2)
    fp = forward_reads(primary, L)
    rb = 2 * L + 1
    with GraphBuilder(L, k, 2, 60) as gb:
        bad = fp.copy()
        bad[7 * rb + 5] = ord("X")
        with pytest.raises(VdjGraphError) as e:
            gb.build_forward(bad)
        assert e.value.code == -3 and "record 14 " in str(e.value)
        bad = fp.copy()
        bad[3 * rb] = ord("2")
        with pytest.raises(VdjGraphError) as e:
            gb.build_forward(bad)
        assert e.value.code == -2 and "record 6 " in str(e.value)
        # reads with N: both strands carry it at mirrored positions
        withn = primary.copy()
        recs = withn[:-1].reshape(-1, rb)
        recs[0::2, 1 + 10] = ord("N")
        recs[1::2, 1 + L - 1 - 10] = ord("N")
        want = loader.build(withn, np.zeros(1, np.uint8), L, k, 2, 60, kind="port")
        got = gb.build_forward(forward_reads(withn, L))
        assert_graph_equal(got, want, "forward with N")
