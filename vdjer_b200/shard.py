"""Sharded graph build over G = 1, 2, 4 or 8 B200s (phases in include/vdjgraph.h).

Records are split into contiguous ranges (rank order = record order), k-mers are owner-computed:
every hash unit (a group of minimizer buckets) belongs to one rank, chosen by the library from the
all-gathered window counts.  This module is only the host plumbing between the library's phases:
a handful of small exchanges per build (histograms + HyperLogLog registers + IPC handles in one all-gather,
a flag gather that doubles as a barrier, one barrier per round, survivor counts, one barrier after
the finish).  The bulk exchange is inside the kernels: k_scatter writes every run (a
stretch of consecutive windows, 32 bytes) straight into the owner's buffer through peer-mapped memory
(NVLink / NVSwitch), the exact read comparison and the quality rows of border k-mers are peer
loads; the finish is distributed too: every rank ranks and links its own survivors (neighbours
owned by a peer are peer loads), the finished node rows are stored into rank 0's result buffer,
and the three steps of it are separated by barriers the devices keep in peer memory.  The small
exchanges themselves go through a shared-memory file when all ranks share the host (_HostExchange).

Two ways to run it:
  * one process per GPU (torchrun): `build_distributed(builder, primary, secondary)` with
    torch.distributed initialised (any backend: only small host objects are exchanged); peer
    buffers are mapped with CUDA IPC handles;
  * several ranks inside one process (`build_local`): G contexts on one or several devices, peer
    buffers are plain device pointers.  Used by the tests to check on ONE GPU that the sharded
    result is identical to the single-device result.  (The library's own one-process driver, with
    one host thread per device, is vdjgraph_multi_* / `MultiBuilder`.)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .graph import BUF_GATHER, SHARD_NBUF, GraphBuilder, load_library

RECORD_BYTES = lambda L: 2 * L + 1  # noqa: E731


def plan_inputs(hists: list[np.ndarray], hlls: list[np.ndarray], counts: list[int]):
    """What every rank hands to vdjgraph_shard_plan, from the all-gathered per-rank pieces."""
    hist_all = np.stack([np.asarray(h, np.uint64) for h in hists])
    hll = np.maximum.reduce([np.asarray(h, np.uint8) for h in hlls])   # byte registers: element-wise max
    return hist_all, hll, np.asarray(counts, np.uint64)


def shard_ranges(total_records: int, n_ranks: int) -> list[tuple[int, int]]:
    """Contiguous, balanced record ranges [lo, hi) per rank, even-sized so that a read and its
    reverse complement (stored as consecutive records, bam_read.c:206-244) stay together."""
    per = -(-total_records // n_ranks)
    per += per & 1
    return [(min(total_records, r * per), min(total_records, (r + 1) * per)) for r in range(n_ranks)]


def split_records(primary, secondary, L: int, lo: int, hi: int):
    """Records [lo, hi) of primary ++ secondary as (primary_part, secondary_part) byte arrays."""
    rb = RECORD_BYTES(L)
    p = np.asarray(primary, np.uint8)
    s = np.asarray(secondary, np.uint8)
    n_p = p.size // rb
    a = p[min(lo, n_p) * rb: min(hi, n_p) * rb]
    b = s[max(0, lo - n_p) * rb: max(0, hi - n_p) * rb]
    return a, b


# ------------------------------------------------------------------------------------------------
# several ranks in one process
# ------------------------------------------------------------------------------------------------
def build_local(builders: list[GraphBuilder], parts: list[tuple], devices: list[int] | None = None, copy: bool = True,
                forward: bool = False):
    """Run the sharded build with rank r = builders[r] on parts[r] = (primary, secondary); all in
    this process.  Returns rank 0's Graph.  forward: the parts hold forward reads only (two packed
    records per read, vdjgraph_shard_stage_forward)."""
    G = len(builders)
    lib = load_library()
    counts = [b._n_records(p, s) * (2 if forward else 1) for b, (p, s) in zip(builders, parts)]
    total = sum(counts)
    base = np.concatenate([[0], np.cumsum(counts)])
    if devices:
        for a in set(devices):
            for b in set(devices):
                rc = lib.vdjgraph_enable_peer_access(a, b)
                if rc:
                    raise RuntimeError(f"peer access {a}->{b}: {lib.vdjgraph_last_error().decode()}")
    for r, (b, (p, s)) in enumerate(zip(builders, parts)):
        b.shard_stage(p, s, G, r, int(base[r]), total, forward=forward)
    pieces = [b.shard_count() for b in builders]
    hist_all, hll, cnt = plan_inputs([x[0] for x in pieces], [x[1] for x in pieces], counts)
    for b in builders:
        b.shard_plan(hist_all, hll, cnt)
    table = [b.shard_buffers()[0] for b in builders]
    for b in builders:
        b.shard_set_peers(table)
    # one scatter (= the exchange) + passes per super-partition round; every rank plans the same count
    for _ in range(builders[0].shard_rounds()):
        for b in builders:
            b.shard_scatter()
        surv = [b.shard_passes() for b in builders]
    for b in builders:
        b.shard_gather_plan(surv)
    table = [b.shard_buffers()[0] for b in builders]
    for b in builders:
        b.shard_set_peers(table)
    # the ranks share this thread: a step is queued and waited for on every rank before the next begins
    for step in range(3):
        for b in builders:
            b.shard_finish_step(step, False)
    for b in builders:
        b.shard_finish()
    return builders[0].fetch(copy=copy)


# ------------------------------------------------------------------------------------------------
# one process per GPU
# ------------------------------------------------------------------------------------------------
class _Peers:
    """CUDA IPC mappings of the other ranks' buffers.  Mapping a multi-GB allocation costs far more
    than a build step, so mappings are kept across builds: a (rank, buffer) is re-mapped only when
    its handle changes, i.e. when the owner had to grow the buffer (the owner parks the old
    allocation until vdjgraph_shard_release_retired, so closing it late is safe)."""

    def __init__(self):
        self.lib = load_library()
        self.cache = {}   # (rank, buffer) -> (handle, mapped pointer)

    def export(self, ptr: int) -> bytes:
        if not ptr:
            return b""
        h = (C.c_ubyte * 64)()
        if self.lib.vdjgraph_ipc_export(C.c_void_p(ptr), h):
            raise RuntimeError("vdjgraph_ipc_export: " + self.lib.vdjgraph_last_error().decode())
        return bytes(h)

    def map(self, rank: int, buf: int, handle: bytes) -> int:
        old = self.cache.get((rank, buf))
        if old and old[0] == handle:
            return old[1]
        if old:
            self.lib.vdjgraph_ipc_close(C.c_void_p(old[1]))
            del self.cache[(rank, buf)]
        if not handle:
            return 0
        out = C.c_void_p()
        raw = (C.c_ubyte * 64).from_buffer_copy(handle)
        if self.lib.vdjgraph_ipc_open(raw, C.byref(out)):
            raise RuntimeError("vdjgraph_ipc_open: " + self.lib.vdjgraph_last_error().decode())
        self.cache[(rank, buf)] = (handle, int(out.value))
        return int(out.value)

    def close(self):
        for _, ptr in self.cache.values():
            self.lib.vdjgraph_ipc_close(C.c_void_p(ptr))
        self.cache = {}


class ShardAborted(RuntimeError):
    """Another rank of the sharded build failed; this rank stopped with it."""


class _HostExchange:
    """all-gather of small byte strings between the ranks of ONE host through a shared-memory file
    (/dev/shm): a rank writes its payload into its own slot, then the slot's sequence number, and reads
    the others' once their sequence numbers have arrived.  A build makes seven of these exchanges; through
    torch.distributed each costs about 0.3 ms (tensor round trip through the device, NCCL launch), here a
    few microseconds.  Slots are double-buffered by the parity of the sequence number: a rank can only be
    one exchange ahead of the slowest one (it needs everybody's payload to get out of an exchange), so the
    slot of exchange n is not rewritten before everybody has read it.  The file is unlinked as soon as
    every rank has mapped it.  Stores become visible in program order (x86)."""

    SLOT = 1 << 17          # payload bytes per rank (histograms + HyperLogLog registers + handles: 40 kB)
    HEAD = 64

    def __init__(self, dist, group, rank: int, G: int, timeout_s: float = 300.0):
        import mmap
        import secrets
        self.rank, self.G, self.timeout_s = rank, G, timeout_s
        self.seq = 0
        size = G * 2 * (self.SLOT + self.HEAD)
        names = [f"/dev/shm/vdjgraph_{os.getpid()}_{secrets.token_hex(8)}" if rank == 0 else None]
        kw = {"group": group} if group is not None else {}
        if rank == 0:
            with open(names[0], "wb") as f:
                f.truncate(size)
        dist.broadcast_object_list(names, src=0, **kw)
        with open(names[0], "r+b") as f:
            self.mm = mmap.mmap(f.fileno(), size)
        dist.barrier(**kw)
        if rank == 0:
            os.unlink(names[0])
        raw = np.frombuffer(self.mm, np.uint8).reshape(G, 2, self.SLOT + self.HEAD)
        self.head = raw[:, :, :self.HEAD].view(np.uint64)      # [G, 2, 8]: sequence number, payload bytes
        self.body = raw[:, :, self.HEAD:]

    def all_gather(self, payload: np.ndarray) -> np.ndarray:
        import time
        n = payload.size
        if n > self.SLOT:
            raise ValueError(f"exchange payload of {n} bytes")
        self.seq += 1
        par = self.seq & 1
        self.body[self.rank, par, :n] = payload
        self.head[self.rank, par, 1] = n
        self.head[self.rank, par, 0] = self.seq            # last: the payload is complete
        out = np.empty((self.G, n), np.uint8)
        t0 = None
        for r in range(self.G):
            spins = 0
            while self.head[r, par, 0] != self.seq:
                spins += 1
                if spins % 64 == 0:
                    os.sched_yield()                           # fewer cores than ranks: let the late rank run
                if spins % 4096 == 0:
                    t0 = t0 or time.perf_counter()
                    if time.perf_counter() - t0 > self.timeout_s:
                        raise ShardAborted(f"sharded build aborted: rank {r} did not reach exchange {self.seq}")
            if int(self.head[r, par, 1]) != n:
                raise ShardAborted(f"sharded build aborted: rank {r} sent {int(self.head[r, par, 1])} bytes, expected {n}")
            out[r] = self.body[r, par, :n]
        return out

    def close(self):
        self.head = self.body = None
        try:
            self.mm.close()
        except BufferError:
            pass


class DistributedBuilder:
    """One rank of the sharded build (one process per GPU).  `dist`: torch.distributed (initialised);
    `group`: process group for the small host exchanges (default group if None); `device`: where
    the exchanged tensors live ("cuda:N" with an NCCL group: a few microseconds per exchange over
    NVLink; None = CPU tensors, e.g. a gloo group); `device_barriers`: False = the steps of the finish
    are separated by host barriers instead of the ones the devices keep in peer memory (also:
    VDJGRAPH_HOST_BARRIERS=1); `exchange`: "auto" = the small host exchanges go through shared memory
    when every rank runs on this host (_HostExchange) and through torch.distributed otherwise, "torch" =
    always the latter (also: VDJGRAPH_EXCHANGE=torch).  This process' records are its `primary` /
    `secondary`; rank order = record order.  close() before the GraphBuilder is closed: it unmaps
    the peers' buffers."""

    def __init__(self, builder: GraphBuilder, dist=None, group=None, device=None, device_barriers: bool | None = None,
                 exchange: str = "auto"):
        if dist is None:
            import torch.distributed as dist  # noqa: PLW0642
        self.b, self.dist, self.group, self.device = builder, dist, group, device
        self.G = dist.get_world_size(group) if group is not None else dist.get_world_size()
        self.rank = dist.get_rank(group) if group is not None else dist.get_rank()
        self.counts = None
        self._err = None
        self._gather_mapped = False
        # small host exchanges: shared memory when every rank runs on this host, torch.distributed otherwise
        self.xchg = None
        if exchange == "auto":
            exchange = os.environ.get("VDJGRAPH_EXCHANGE", "auto")
        if exchange in ("auto", "shm") and self.G > 1:
            import socket
            hosts = [None] * self.G
            kw = {"group": group} if group is not None else {}
            dist.all_gather_object(hosts, socket.gethostname(), **kw)
            if len(set(hosts)) == 1 and os.path.isdir("/dev/shm"):
                self.xchg = _HostExchange(dist, group, self.rank, self.G)
            elif exchange == "shm":
                raise RuntimeError("exchange='shm' needs every rank on one host")
        # the three steps of the finish meet at barriers in peer memory (default) or at host barriers
        self.device_barriers = os.environ.get("VDJGRAPH_HOST_BARRIERS", "0") != "1" if device_barriers is None else device_barriers
        self.peers = _Peers()
        self.table = [[0] * SHARD_NBUF for _ in range(self.G)]

    def _try(self, fn, *args, default=None):
        """A library phase of this rank.  When it fails the exception is parked, the rank keeps taking
        part in the collectives with placeholder data, and the next collective makes EVERY rank raise:
        no rank is left waiting in a barrier for one that has gone."""
        if self._err is not None:
            return default
        try:
            return fn(*args)
        except Exception as e:   # noqa: BLE001
            self._err = e
            return default

    def _gather(self, arr: np.ndarray) -> np.ndarray:
        """all-gather a small fixed-shape array: result [G, *arr.shape].  Every exchange also carries
        each rank's failure flag (see _try)."""
        a = np.ascontiguousarray(arr)
        payload = np.concatenate([a.view(np.uint8).reshape(-1), np.array([1 if self._err is not None else 0], np.uint8)])
        if self.xchg is not None:
            flat = self.xchg.all_gather(payload)
        else:
            import torch
            t = torch.from_numpy(payload)
            if self.device is not None:
                t = t.to(self.device)
            out = [torch.empty_like(t) for _ in range(self.G)]
            kw = {"group": self.group} if self.group is not None else {}
            self.dist.all_gather(out, t, **kw)
            flat = torch.stack(out).cpu().numpy()
        bad = [r for r in range(self.G) if flat[r, -1]]
        if bad:
            err, self._err = self._err, None
            if err is not None:
                raise err
            raise ShardAborted(f"sharded build aborted: rank(s) {bad} failed")
        return np.ascontiguousarray(flat[:, :-1]).view(a.dtype).reshape((self.G,) + a.shape)

    def _barrier(self):
        self._gather(np.zeros(0, np.uint8))

    def _handles(self, only) -> np.ndarray:
        """IPC handles of this rank's buffers in `only`: [NBUF*64 handle bytes | NBUF present flags]."""
        ptrs, _ = self._try(self.b.shard_buffers, default=([0] * SHARD_NBUF, None))
        mine = np.zeros((SHARD_NBUF, 64), np.uint8)
        have = np.zeros(SHARD_NBUF, np.uint8)
        for i in only:
            h = self._try(self.peers.export, ptrs[i], default=b"")
            if h:
                mine[i] = np.frombuffer(h, np.uint8)
                have[i] = 1
        return np.concatenate([mine.reshape(-1), have])

    def _install(self, handles, only):
        """map the peers' buffers in `only` from their all-gathered handles and hand the table to the library
        (a mapping is kept as long as its handle does not change, see _Peers)"""
        for r in range(self.G):
            if r != self.rank:
                hs, hv = handles[r][:SHARD_NBUF * 64].reshape(SHARD_NBUF, 64), handles[r][SHARD_NBUF * 64:]
                for i in only:
                    self.table[r][i] = self._try(self.peers.map, r, i, hs[i].tobytes() if hv[i] else b"", default=0)
        self._try(self.b.shard_set_peers, self.table)

    def _exchange(self, only):
        """all-gather the IPC handles of this rank's buffers in `only`; map the peers'."""
        self._install(self._gather(self._handles(only)), only)

    def stage(self, primary, secondary=b"", forward: bool = False):
        """forward: this rank's buffers hold forward reads only (every rank alike)."""
        n_local = self.b._n_records(primary, secondary) * (2 if forward else 1)
        self.counts = [int(x) for x in self._gather(np.array([n_local], np.int64))[:, 0]]
        base = int(sum(self.counts[:self.rank]))
        if forward:
            self._try(lambda: self.b.shard_stage(primary, secondary, self.G, self.rank, base, int(sum(self.counts)), forward=True))
        else:
            self._try(self.b.shard_stage, primary, secondary, self.G, self.rank, base, int(sum(self.counts)))

    def run(self):
        """count -> plan -> scatter (= the all-to-all) -> passes -> distributed finish on the staged
        records.  Up to the finish every phase returns with its stream synchronised, so the host
        barriers order the devices.  The graph then sits on rank 0's device: fetch() copies it to the host.
        self.phase_ms holds the wall time of each phase of the last run."""
        import time
        b = self.b
        t = [time.perf_counter()]

        def mark():
            t.append(time.perf_counter())

        from .graph import SHARD_HIST, SHARD_HLL
        NH = SHARD_NBUF * 65
        data = [i for i in range(SHARD_NBUF) if i != BUF_GATHER]
        hist, hll = self._try(b.shard_count, default=(np.zeros(SHARD_HIST, np.uint64), np.zeros(SHARD_HLL, np.uint8)))
        before, _ = self._try(b.shard_buffers, default=([0] * SHARD_NBUF, None))
        mark()
        # one exchange: histograms, cardinality registers and the handles of the buffers as they are now
        pieces = self._gather(np.concatenate([hist.view(np.uint8), hll, self._handles(data)]))
        mark()
        n_h = hist.size * 8
        hist_all, hll_m, cnt = plan_inputs([x[:n_h].copy().view(np.uint64) for x in pieces], [x[n_h:-NH] for x in pieces], self.counts)
        self._try(b.shard_plan, hist_all, hll_m, cnt)
        mark()
        # the plan may have grown this rank's run buffer: only then are handles exchanged again.  This
        # gather is also the barrier "every peer buffer exists" and carries the number of rounds.
        after, sizes = self._try(b.shard_buffers, default=([0] * SHARD_NBUF, [0] * SHARD_NBUF))
        grown = any(before[i] != after[i] for i in data)
        n_rounds = self._try(b.shard_rounds, default=0)
        flags = self._gather(np.array([1 if grown else 0, n_rounds], np.int64))
        n_rounds = int(flags[:, 1].max())   # a failed rank still walks the rounds
        if flags[:, 0].any():
            self._exchange(data)
            self._barrier()                 # ... and is mapped everywhere
        else:
            self._install([x[-NH:] for x in pieces], data)
        self._try(b.shard_release_retired)
        mark()
        t_sc = t_pa = 0.0
        n_surv = 0
        for rnd in range(n_rounds):
            if rnd:
                self._barrier()             # the owners have consumed the previous round's runs
            t0 = time.perf_counter()
            self._try(b.shard_scatter)
            self._barrier()                 # every rank's runs of this round have arrived
            t1 = time.perf_counter()
            n_surv = self._try(b.shard_passes, default=0)
            t_sc += t1 - t0
            t_pa += time.perf_counter() - t1
        t.append(t[-1] + t_sc)
        t.append(t[-1] + t_pa)
        # survivor counts, and how much every rank's exchange buffer holds: every rank can tell whether one of
        # them has to grow (only then do the handles travel again)
        got = self._gather(np.array([n_surv, sizes[BUF_GATHER] if sizes else 0, 1 if self._gather_mapped else 0], np.int64))
        surv = [int(x) for x in got[:, 0]]
        self._try(b.shard_gather_plan, surv)
        need = [self._try(b.shard_finish_bytes, surv, r, default=0) for r in range(self.G)]
        if any(need[r] > int(got[r, 1]) for r in range(self.G)) or not got[:, 2].all():
            self._exchange([BUF_GATHER])
            self._gather_mapped = True
        else:
            self._try(self.b.shard_set_peers, self.table)
        # (a rank's barrier flags are in place before its handle travels, or since the plan of this build)
        mark()
        for step in range(3):
            self._try(b.shard_finish_step, step, self.device_barriers)
            if not self.device_barriers:
                self._barrier()
        self._try(b.shard_finish)
        self._barrier()                     # nobody returns before the finish is known to have worked
        self._try(b.shard_release_retired)  # ... and every peer has let go of a buffer that was replaced
        mark()
        names = ["count", "x_hist", "plan", "x_peers", "scatter+barrier", "passes", "x_surv", "finish"]
        self.phase_ms = {n: (t[i + 1] - t[i]) * 1e3 for i, n in enumerate(names)}

    def fetch(self, copy: bool = True):
        """The graph on rank 0 (None elsewhere)."""
        return self.b.fetch(copy=copy) if self.rank == 0 else None

    def build(self, primary, secondary=b"", copy: bool = True, forward: bool = False):
        self.stage(primary, secondary, forward=forward)
        self.run()
        return self.fetch(copy=copy)

    def close(self):
        self._barrier()                     # nobody is still reading a peer buffer
        self.peers.close()
        self._barrier()                     # everything is unmapped before the owners free
        if self.xchg is not None:
            self.xchg.close()
            self.xchg = None


def build_distributed(builder: GraphBuilder, primary, secondary, dist=None, copy: bool = True, exchange: str = "auto"):
    """stage + run of one rank; see DistributedBuilder."""
    db = DistributedBuilder(builder, dist, exchange=exchange)
    try:
        return db.build(primary, secondary, copy=copy)
    finally:
        db.close()
