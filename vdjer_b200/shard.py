"""Sharded graph build over G = 1, 2, 4 or 8 B200s (phases in include/vdjgraph.h).

Records are split into contiguous ranges (rank order = record order), k-mers are owner-computed:
hash partition p belongs to rank p mod G.  This module is only the host plumbing between the
library's phases: three small all-gathers (window histograms + record counts + HyperLogLog
registers, device pointers, survivor counts) and three barriers.  The bulk exchange is inside the
kernels: k_scatter writes every tuple straight into the owner's buffer through peer-mapped memory
(NVLink / NVSwitch), the exact read comparison and the quality rows of border k-mers are peer
loads, and the survivors travel to rank 0 as one device-to-device copy per rank.

Two ways to run it:
  * one process per GPU (torchrun): `build_distributed(builder, primary, secondary)` with
    torch.distributed initialised (any backend: only small host objects are exchanged); peer
    buffers are mapped with CUDA IPC handles;
  * several ranks inside one process (`build_local`): G contexts on one or several devices, peer
    buffers are plain device pointers.  Used by the tests to check on ONE GPU that the sharded
    result is identical to the single-device result.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .graph import BUF_GATHER, SHARD_NBUF, GraphBuilder, load_library

RECORD_BYTES = lambda L: 2 * L + 1  # noqa: E731


def plan_inputs(hists: list[np.ndarray], hlls: list[np.ndarray], counts: list[int]):
    """What every rank hands to vdjgraph_shard_plan, from the all-gathered per-rank pieces."""
    hist_all = np.stack([np.asarray(h, np.uint64) for h in hists])
    hll = np.maximum.reduce([np.asarray(h, np.uint32) for h in hlls])
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
def build_local(builders: list[GraphBuilder], parts: list[tuple], devices: list[int] | None = None, copy: bool = True):
    """Run the sharded build with rank r = builders[r] on parts[r] = (primary, secondary); all in
    this process.  Returns rank 0's Graph."""
    G = len(builders)
    lib = load_library()
    counts = [b._n_records(p, s) for b, (p, s) in zip(builders, parts)]
    total = sum(counts)
    base = np.concatenate([[0], np.cumsum(counts)])
    if devices:
        for a in set(devices):
            for b in set(devices):
                rc = lib.vdjgraph_enable_peer_access(a, b)
                if rc:
                    raise RuntimeError(f"peer access {a}->{b}: {lib.vdjgraph_last_error().decode()}")
    for r, (b, (p, s)) in enumerate(zip(builders, parts)):
        b.shard_stage(p, s, G, r, int(base[r]), total)
    pieces = [b.shard_count() for b in builders]
    hist_all, hll, cnt = plan_inputs([x[0] for x in pieces], [x[1] for x in pieces], counts)
    for b in builders:
        b.shard_plan(hist_all, hll, cnt)
    table = [b.shard_buffers()[0] for b in builders]
    for b in builders:
        b.shard_set_peers(table)
    for b in builders:
        b.shard_scatter()
    surv = [b.shard_passes() for b in builders]
    for b in builders:
        b.shard_gather_plan(surv)
    table = [b.shard_buffers()[0] for b in builders]
    for b in builders:
        b.shard_set_peers(table)
    for b in builders:
        b.shard_send()
    builders[0].shard_finish()
    return builders[0].fetch(copy=copy)


# ------------------------------------------------------------------------------------------------
# one process per GPU
# ------------------------------------------------------------------------------------------------
class _Peers:
    """CUDA IPC mappings of the other ranks' buffers, closed before the owners may free them."""

    def __init__(self):
        self.lib = load_library()
        self.open = []

    def export(self, ptr: int) -> bytes:
        if not ptr:
            return b""
        h = (C.c_ubyte * 64)()
        if self.lib.vdjgraph_ipc_export(C.c_void_p(ptr), h):
            raise RuntimeError("vdjgraph_ipc_export: " + self.lib.vdjgraph_last_error().decode())
        return bytes(h)

    def map(self, handle: bytes) -> int:
        if not handle:
            return 0
        out = C.c_void_p()
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        if self.lib.vdjgraph_ipc_open(buf, C.byref(out)):
            raise RuntimeError("vdjgraph_ipc_open: " + self.lib.vdjgraph_last_error().decode())
        self.open.append(out.value)
        return int(out.value)

    def close(self):
        for p in self.open:
            self.lib.vdjgraph_ipc_close(C.c_void_p(p))
        self.open = []


def _all_gather(obj, dist):
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def build_distributed(builder: GraphBuilder, primary, secondary, dist=None, copy: bool = True, timings: dict | None = None):
    """One rank of the sharded build: this process' records are `primary`/`secondary`, rank order =
    record order.  Returns the Graph on rank 0 and None elsewhere.  `dist`: torch.distributed
    (initialised) or any object with get_rank/get_world_size/all_gather_object/barrier."""
    if dist is None:
        import torch.distributed as dist  # noqa: PLW0642
    G, rank = dist.get_world_size(), dist.get_rank()
    peers = _Peers()
    try:
        n_local = builder._n_records(primary, secondary)
        counts = _all_gather(n_local, dist)
        base = int(sum(counts[:rank]))
        builder.shard_stage(primary, secondary, G, rank, base, int(sum(counts)))
        hist, hll = builder.shard_count()
        pieces = _all_gather((hist, hll), dist)
        hist_all, hll_m, cnt = plan_inputs([x[0] for x in pieces], [x[1] for x in pieces], counts)
        builder.shard_plan(hist_all, hll_m, cnt)

        def exchange(only=None):
            ptrs, _ = builder.shard_buffers()
            mine = [peers.export(p) if (only is None or i in only) else b"" for i, p in enumerate(ptrs)]
            handles = _all_gather(mine, dist)
            return [[0] * SHARD_NBUF if r == rank else [peers.map(h) for h in handles[r]] for r in range(G)]

        table = exchange(only=set(range(SHARD_NBUF)) - {BUF_GATHER})
        builder.shard_set_peers(table)
        dist.barrier()                      # every peer buffer exists and is mapped
        builder.shard_scatter()
        dist.barrier()                      # every rank's tuples have arrived
        surv = _all_gather(builder.shard_passes(), dist)
        builder.shard_gather_plan(surv)
        gather = exchange(only={BUF_GATHER})
        for r in range(G):
            table[r][BUF_GATHER] = gather[r][BUF_GATHER]
        builder.shard_set_peers(table)
        builder.shard_send()
        dist.barrier()                      # rank 0 holds every survivor record
        graph = None
        if rank == 0:
            builder.shard_finish()
            graph = builder.fetch(copy=copy)
        dist.barrier()                      # peers stay mapped until everybody is done with them
        return graph
    finally:
        peers.close()
