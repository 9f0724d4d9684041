"""Shared helpers for the parity tests (CPU oracle vs CUDA path)."""
from __future__ import annotations

import hashlib
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i


def sha(*bufs) -> str:
    h = hashlib.sha256()
    for b in bufs:
        h.update(np.ascontiguousarray(b).tobytes())
    return h.hexdigest()


RESULT_ARRAYS = dict(first_pos=np.uint64, frequency=np.uint16, out_deg=np.uint8, in_deg=np.uint8,
                     out_succ=np.uint32, in_pred=np.uint32)


def digest_of(get) -> dict:
    """SHA-256 per result array (little-endian, C order, the dtypes of include/vdjgraph.h);
    get(name) returns the array."""
    return {name: sha(np.ascontiguousarray(get(name), dtype=dt)) for name, dt in RESULT_ARRAYS.items()}


def kmer_codes_at(primary, secondary, L, k, stamps):
    """2-bit codes [n,k] of the k-mers whose first base is at the given stamps (r*w + o)."""
    w = L - k + 1
    rec = 2 * L + 1
    stamps = np.asarray(stamps, dtype=np.uint64)
    n_p = (np.asarray(primary).size) // rec
    r = (stamps // np.uint64(w)).astype(np.int64)
    o = (stamps % np.uint64(w)).astype(np.int64)
    both = np.concatenate([np.asarray(primary, np.uint8)[: n_p * rec], np.asarray(secondary, np.uint8)])
    start = r * rec + 1 + o
    idx = start[:, None] + np.arange(k)[None, :]
    return _CODE[both[idx]]


def pack_codes(codes):
    """[n,k] 2-bit codes -> (lo, hi) uint64 with base j at bits 2j (the library's key layout)."""
    n, k = codes.shape
    lo = np.zeros(n, dtype=np.uint64)
    hi = np.zeros(n, dtype=np.uint64)
    for j in range(k):
        c = codes[:, j].astype(np.uint64)
        if j < 32:
            lo |= c << np.uint64(2 * j)
        else:
            hi |= c << np.uint64(2 * (j - 32))
    return lo, hi


def assert_graph_equal(got, want: dict, what=""):
    """got: vdjer_b200.Graph ; want: oracle.loader.build() dict.  Bit-exact, every field."""
    assert got.n_nodes == want["n_nodes"], f"{what}: n_nodes {got.n_nodes} != {want['n_nodes']}"
    for name in ["first_pos", "frequency", "out_deg", "in_deg", "out_succ", "in_pred"]:
        a, b = getattr(got, name), want[name]
        if not np.array_equal(a, b):
            bad = np.nonzero((a != b).reshape(len(a), -1).any(axis=1))[0]
            raise AssertionError(f"{what}: {name} differs at {bad.size} nodes, first {bad[:5]}: "
                                 f"got {a[bad[:3]]} want {b[bad[:3]]}")


def assert_pre_table_equal(pre, want: dict, primary, secondary, L, k, what=""):
    """Pruned pass-1 table: same k-mer set and the same pre_node.frequency for every k-mer."""
    assert len(pre.kmer_lo) == want["n_pre"], f"{what}: n_pre {len(pre.kmer_lo)} != {want['n_pre']}"
    if want["n_pre"] == 0:
        return
    lo, hi = pack_codes(kmer_codes_at(primary, secondary, L, k, want["pre_first_pos"]))
    w_order = np.lexsort((lo, hi))
    g_order = np.lexsort((pre.kmer_lo, pre.kmer_hi))
    assert np.array_equal(lo[w_order], pre.kmer_lo[g_order]) and np.array_equal(hi[w_order], pre.kmer_hi[g_order]), \
        f"{what}: surviving k-mer sets differ"
    assert np.array_equal(want["pre_freq"][w_order], pre.frequency[g_order]), f"{what}: pre_node.frequency differs"


def graph_invariants(g, L, k, primary, secondary):
    """Size-independent properties of any correct result (used at full benchmark sizes)."""
    n = g.n_nodes
    fp = g.first_pos.astype(np.uint64)
    assert np.all(fp[1:] > fp[:-1]), "nodes must be in strictly increasing first-occurrence order"
    assert np.all(g.frequency >= 1) and np.all(g.frequency <= 32765)
    nil = np.uint32(0xFFFFFFFF)
    for deg, adj in ((g.out_deg, g.out_succ), (g.in_deg, g.in_pred)):
        cols = np.arange(4)[None, :]
        used = cols < deg[:, None]
        assert np.all(adj[used] < n) and np.all(adj[~used] == nil)
    # every out-edge u->v appears as an in-edge of v, and the counts match
    src = np.repeat(np.arange(n, dtype=np.int64), 4).reshape(n, 4)
    mo = np.arange(4)[None, :] < g.out_deg[:, None]
    mi = np.arange(4)[None, :] < g.in_deg[:, None]
    e_out = np.stack([src[mo], g.out_succ[mo].astype(np.int64)], 1)
    e_in = np.stack([g.in_pred[mi].astype(np.int64), src[mi]], 1)
    a = e_out[np.lexsort((e_out[:, 1], e_out[:, 0]))]
    b = e_in[np.lexsort((e_in[:, 1], e_in[:, 0]))]
    assert np.array_equal(a, b), "toNodes and fromNodes disagree"
    # an edge u->v overlaps by k-1 bases
    if len(e_out):
        sel = e_out[:: max(1, len(e_out) // 200000)]
        cu = kmer_codes_at(primary, secondary, L, k, fp[sel[:, 0]])
        cv = kmer_codes_at(primary, secondary, L, k, fp[sel[:, 1]])
        assert np.all(cu[:, 1:] == cv[:, :-1]), "edge endpoints do not overlap by k-1"
    # node k-mers are distinct and N-free
    step = max(1, n // 500000)
    codes = kmer_codes_at(primary, secondary, L, k, fp[::step])
    assert np.all(codes < 4)
    if step == 1 and n:
        lo, hi = pack_codes(codes)
        assert len(np.unique(np.stack([lo, hi], 1), axis=0)) == n, "duplicate k-mers among nodes"
