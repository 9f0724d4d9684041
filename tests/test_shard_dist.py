"""CPU, world_size 2 over gloo: the host plumbing of the sharded build (vdjer_b200/shard.py).
No GPU here, so the library's phases are played by a recording stand-in; what is checked is what
the plumbing is responsible for: every rank hands the SAME merged inputs to the plan phase, record
bases follow rank order, peer tables are complete and exclude nobody, every rank's exchange buffer
reaches every peer, phases run in the documented order and nobody deadlocks.
The device side of the same flow is covered on a GPU by tests/test_shard_gpu.py."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from vdjer_b200 import shard  # noqa: E402
from vdjer_b200.graph import BUF_GATHER, SHARD_HIST, SHARD_HLL, SHARD_NBUF  # noqa: E402


class FakeBuilder:
    """Stands in for GraphBuilder's shard_* methods; pointers are tagged integers."""

    def __init__(self, rank, n_records, rounds=1):
        self.rank, self.n, self.log = rank, n_records, []
        self.gather = 0
        self.rounds = rounds

    def shard_rounds(self):
        return self.rounds

    def _n_records(self, primary, secondary):
        return self.n

    def shard_stage(self, p, s, G, rank, base, total):
        self.log.append(("stage", G, rank, base, total))

    def shard_count(self):
        rng = np.random.default_rng(100 + self.rank)
        self.hist = rng.integers(0, 1000, SHARD_HIST).astype(np.uint64)
        self.hll = rng.integers(0, 30, SHARD_HLL).astype(np.uint8)
        self.log.append(("count",))
        return self.hist, self.hll

    def shard_plan(self, hist_all, hll, counts):
        self.log.append(("plan", hist_all.copy(), hll.copy(), counts.copy()))

    def shard_buffers(self):
        ptrs = [1000 * (self.rank + 1) + i for i in range(SHARD_NBUF)]
        ptrs[BUF_GATHER] = self.gather
        return ptrs, [0] * SHARD_NBUF

    def shard_set_peers(self, table):
        self.log.append(("peers", [list(r) for r in table]))

    def shard_scatter(self):
        self.log.append(("scatter",))

    def shard_passes(self):
        self.log.append(("passes",))
        self.surv = getattr(self, "surv", 0) + 10 + self.rank      # survivors accumulate over the rounds
        return self.surv

    def shard_gather_plan(self, surv):
        self.log.append(("gather_plan", list(surv)))
        self.gather = 777 + 100 * self.rank

    def shard_finish_bytes(self, surv, rank):
        return 64 * sum(surv) + 8 * surv[rank]

    def shard_finish_step(self, step, device_barrier):
        self.log.append(("step", step, device_barrier))

    def shard_finish(self):
        self.log.append(("finish",))

    def shard_release_retired(self):
        self.log.append(("release",))

    def fetch(self, copy=True):
        return "graph"


def _worker(rank, world, port, out_dir, rounds=1, exchange="auto"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # identity "IPC": a handle is the pointer's decimal text
    shard._Peers.export = lambda self, ptr: str(ptr).encode().ljust(64, b" ") if ptr else b""
    shard._Peers.map = lambda self, r, i, h: int(h) + 5 if h else 0      # +5: a mapping is a different address
    shard._Peers.close = lambda self: None
    b = FakeBuilder(rank, 100 + 20 * rank, rounds)
    g = shard.build_distributed(b, None, None, dist=dist, exchange=exchange)
    np.save(os.path.join(out_dir, f"log{rank}.npy"), np.array([b.log, g], dtype=object), allow_pickle=True)
    dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("exchange", ["shm", "torch"])
def test_two_rank_plumbing_over_gloo(tmp_path, exchange):
    """exchange: the small host messages through shared memory (ranks of one host, the default) or through
    torch.distributed collectives"""
    world, port = 2, 29500 + os.getpid() % 2000 + (7 if exchange == "shm" else 0)
    mp.spawn(_worker, args=(world, port, str(tmp_path), 1, exchange), nprocs=world, join=True)
    logs = [np.load(tmp_path / f"log{r}.npy", allow_pickle=True) for r in range(world)]
    (l0, g0), (l1, g1) = logs
    assert g0 == "graph" and g1 is None
    order = ["stage", "count", "plan", "peers", "release", "scatter", "passes", "gather_plan", "peers", "step", "step", "step",
             "finish", "release"]
    assert [e[0] for e in l0] == order and [e[0] for e in l1] == order
    # the three steps of the finish, in order, meeting in peer memory (the default)
    assert [e[1:] for e in l0 if e[0] == "step"] == [(0, True), (1, True), (2, True)]
    # record bases follow rank order; totals agree
    assert l0[0] == ("stage", 2, 0, 0, 220) and l1[0] == ("stage", 2, 1, 100, 220)
    # both ranks plan from identical merged inputs
    for a, b in zip(l0[2][1:], l1[2][1:]):
        assert np.array_equal(a, b)
    hist_all, hll, counts = l0[2][1:]
    assert hist_all.shape == (2, SHARD_HIST) and list(counts) == [100, 120]
    rng0, rng1 = np.random.default_rng(100), np.random.default_rng(101)
    h0, m0 = rng0.integers(0, 1000, SHARD_HIST), rng0.integers(0, 30, SHARD_HLL)
    h1, m1 = rng1.integers(0, 1000, SHARD_HIST), rng1.integers(0, 30, SHARD_HLL)
    assert np.array_equal(hist_all[0], h0) and np.array_equal(hist_all[1], h1)
    assert np.array_equal(hll, np.maximum(m0, m1))           # HyperLogLog registers merge by max
    # peer tables: the other rank's buffers mapped (pointer + 5), own row left to the library
    t0, t1 = l0[3][1], l1[3][1]
    assert t0[0] == [0] * SHARD_NBUF and t1[1] == [0] * SHARD_NBUF
    assert t0[1][:BUF_GATHER] == [2000 + i + 5 for i in range(BUF_GATHER)]
    assert t1[0][:BUF_GATHER] == [1000 + i + 5 for i in range(BUF_GATHER)]
    # survivor counts all-gathered; each rank's exchange buffer reaches the other
    assert l0[7][1] == [10, 11] and l1[7][1] == [10, 11]
    assert l1[8][1][0][BUF_GATHER] == 777 + 5 and l0[8][1][1][BUF_GATHER] == 877 + 5


@pytest.mark.timeout(120)
def test_two_rank_rounds_over_gloo(tmp_path):
    """Three super-partition rounds: scatter + passes once per round on every rank, the cumulative
    survivor counts of the LAST round are what is gathered."""
    world, port = 2, 31500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path), 3), nprocs=world, join=True)
    logs = [np.load(tmp_path / f"log{r}.npy", allow_pickle=True) for r in range(world)]
    (l0, _), (l1, _) = logs
    order = ["stage", "count", "plan", "peers", "release"] + ["scatter", "passes"] * 3 + ["gather_plan", "peers"] + ["step"] * 3 + ["finish", "release"]
    assert [e[0] for e in l0] == order and [e[0] for e in l1] == order
    gp = order.index("gather_plan")
    assert l0[gp][1] == [30, 33] and l1[gp][1] == [30, 33]


def _failing_worker(rank, world, port, out_dir, where):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard._Peers.export = lambda self, ptr: str(ptr).encode().ljust(64, b" ") if ptr else b""
    shard._Peers.map = lambda self, r, i, h: int(h) + 5 if h else 0
    shard._Peers.close = lambda self: None
    b = FakeBuilder(rank, 100 + 20 * rank)
    if rank == 1:   # this rank's library phase fails (e.g. VDJGRAPH_ERR_NOMEM in vdjgraph_shard_plan)
        def boom(*a, **k):
            raise MemoryError("rank 1 ran out of memory")
        setattr(b, where, boom)
    outcome = "finished"
    db = shard.DistributedBuilder(b, dist)
    try:
        db.build(None, None)
    except shard.ShardAborted as e:
        outcome = "aborted: " + str(e)
    except MemoryError as e:
        outcome = "own error: " + str(e)
    open(os.path.join(out_dir, f"outcome{rank}.txt"), "w").write(outcome)
    dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("where", ["shard_plan", "shard_passes", "shard_finish_step"])
def test_a_failing_rank_aborts_every_rank(tmp_path, where):
    """a rank that raises half-way through run() must not leave the others waiting in a
    barrier.  The failing rank re-raises its own error, the other one raises ShardAborted; both return."""
    world, port = 2, 33500 + os.getpid() % 2000 + {"shard_plan": 0, "shard_passes": 1, "shard_finish_step": 2}[where]
    mp.spawn(_failing_worker, args=(world, port, str(tmp_path), where), nprocs=world, join=True)
    o0 = open(tmp_path / "outcome0.txt").read()
    o1 = open(tmp_path / "outcome1.txt").read()
    assert o0.startswith("aborted") and "[1]" in o0, o0
    assert o1.startswith("own error"), o1


def _hostbar_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard._Peers.export = lambda self, ptr: str(ptr).encode().ljust(64, b" ") if ptr else b""
    shard._Peers.map = lambda self, r, i, h: int(h) + 5 if h else 0
    shard._Peers.close = lambda self: None
    b = FakeBuilder(rank, 100 + 20 * rank)
    db = shard.DistributedBuilder(b, dist, device_barriers=False)
    g = db.build(None, None)
    db.close()
    np.save(os.path.join(out_dir, f"log{rank}.npy"), np.array([b.log, g], dtype=object), allow_pickle=True)
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_host_barriers_between_the_steps_of_the_finish(tmp_path):
    """device_barriers=False (also VDJGRAPH_HOST_BARRIERS=1): every step is asked to synchronise its stream and
    the ranks meet on the host after each; same phase order, nobody hangs."""
    world, port = 2, 37500 + os.getpid() % 2000
    mp.spawn(_hostbar_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        log, _ = np.load(tmp_path / f"log{r}.npy", allow_pickle=True)
        assert [e[1:] for e in log if e[0] == "step"] == [(0, False), (1, False), (2, False)]
        assert [e[0] for e in log][-5:] == ["step", "step", "step", "finish", "release"]


def _exchange_worker(rank, world, port, out_dir):
    import time
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = shard._HostExchange(dist, None, rank, world, timeout_s=60)
    rng = np.random.default_rng(rank)
    ok = True
    for it in range(400):
        n = 1 + (it * 37) % 5000                              # every rank sends n bytes that depend on (rank, it)
        mine = ((np.arange(n) * (rank + 3) + it) % 251).astype(np.uint8)
        if rng.random() < 0.05:
            time.sleep(rng.random() * 0.003)                  # ranks drift apart: one is an exchange ahead at most
        got = x.all_gather(mine)
        for r in range(world):
            ok &= bool(np.array_equal(got[r], ((np.arange(n) * (r + 3) + it) % 251).astype(np.uint8)))
    big = np.full(shard._HostExchange.SLOT, rank, np.uint8)   # a full slot
    got = x.all_gather(big)
    ok &= all(bool((got[r] == r).all()) for r in range(world))
    x.close()
    open(os.path.join(out_dir, f"x{rank}.txt"), "w").write("ok" if ok else "bad")
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_host_exchange_through_shared_memory(tmp_path):
    """_HostExchange: 400 all-gathers of varying size between three drifting processes return every rank's
    payload of THAT exchange (double-buffered slots), and the backing file is gone from /dev/shm."""
    world, port = 3, 35500 + os.getpid() % 2000
    before = set(os.listdir("/dev/shm"))
    mp.spawn(_exchange_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert [open(tmp_path / f"x{r}.txt").read() for r in range(world)] == ["ok"] * world
    assert not [f for f in set(os.listdir("/dev/shm")) - before if f.startswith("vdjgraph_")]


def test_shard_ranges_and_split():
    assert shard.shard_ranges(10, 4) == [(0, 4), (4, 8), (8, 10), (10, 10)]
    for total, G in [(0, 2), (7, 8), (1000, 8), (1001, 4)]:
        r = shard.shard_ranges(total, G)
        assert r[0][0] == 0 and r[-1][1] == total and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert all(lo % 2 == 0 for lo, hi in r if hi > lo)   # a read and its reverse complement stay together
    L = 3
    rb = 2 * L + 1
    p = np.arange(4 * rb, dtype=np.uint8)
    s = np.arange(100, 100 + 3 * rb, dtype=np.uint8)
    parts = [shard.split_records(p, s, L, lo, hi) for lo, hi in [(0, 2), (2, 6), (6, 7)]]
    assert np.array_equal(np.concatenate([a for a, _ in parts]), p)
    assert np.array_equal(np.concatenate([b for _, b in parts]), s)
    assert parts[1][0].size == 2 * rb and parts[1][1].size == 2 * rb
