"""Feasibility numbers for minimizer-partitioned runs ("super-k-mers"), CPU only:
how many bytes per window a run format ships, the run-length distribution, and how evenly the
minimizer buckets spread windows / distinct k-mers over partitions and GPUs, on a sample of a
BASELINE workload.   python profiles/experiments/superkmer_stats.py igh_2x50_5M 200000 [m]
"""
import sys, os, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from vdjer_b200 import synth

def mix32(x):
    x = x.astype(np.uint64)
    x = (x * np.uint64(0x9E3779B1)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x85EBCA77)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(13)
    return x

def main():
    name = sys.argv[1]; pairs = int(sys.argv[2]); m = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    wl = dict(synth.CONFIGS[name]); L, k = wl["read_length"], wl["k"]
    gen = {kk: v for kk, v in wl.items() if kk not in ("k", "mf", "mq", "n_pairs")}
    p, s = synth.generate(seed=12345, n_pairs=pairs, **gen)
    rb = 2 * L + 1
    both = np.concatenate([p[: (p.size // rb) * rb], s[: (s.size // rb) * rb]]).reshape(-1, rb)
    R = both.shape[0]; w = L - k + 1
    code = np.full(256, 4, np.uint8)
    for i, c in enumerate(b"ACGT"): code[c] = i
    b = code[both[:, 1:1 + L]]                       # [R, L]
    q = both[:, 1 + L:].astype(np.int16) - 33
    validb = b < 4
    b0 = np.where(validb, b, 0).astype(np.uint64)
    # m-mer values and hashes at every position
    nm = L - m + 1
    mv = np.zeros((R, nm), np.uint64)
    for j in range(m): mv |= b0[:, j:j + nm] << np.uint64(2 * j)
    mh = mix32(mv)
    # window minimizer hash = min over k-m+1 m-mers
    span = k - m + 1
    wmin = np.full((R, w), np.uint64(1) << np.uint64(40), np.uint64)
    for j in range(span): wmin = np.minimum(wmin, mh[:, j:j + w])
    bucket = (mix32(wmin ^ np.uint64(0x5bd1e995)) >> np.uint64(24)).astype(np.int64)   # re-mixed: the MIN of 26 hashes is not uniform
    cv = np.cumsum(np.concatenate([np.zeros((R, 1), np.int32), (~validb).astype(np.int32)], 1), 1)
    wvalid = (cv[:, k:k + w] - cv[:, :w]) == 0
    good = validb & (q >= 20)
    cg = np.cumsum(np.concatenate([np.zeros((R, 1), np.int32), (~good).astype(np.int32)], 1), 1)
    wgated = (cg[:, k:k + w] - cg[:, :w]) == 0
    # runs: maximal stretches of consecutive valid windows with the same bucket, length-capped
    cap = min(24, 64 - k)
    newrun = wvalid.copy()
    newrun[:, 1:] &= ~(wvalid[:, :-1] & (bucket[:, 1:] == bucket[:, :-1]))
    # apply the cap: walk
    runlen = []
    nr = 0
    W = int(wvalid.sum())
    # vectorised run lengths
    idx = np.flatnonzero(wvalid.ravel())
    starts = newrun.ravel()[idx]
    rid = np.cumsum(starts) - 1
    lens = np.bincount(rid)
    extra = (np.maximum(lens - 1, 0) // cap)        # splits by the cap
    n_runs = int(len(lens) + extra.sum())
    hist = np.bincount(np.minimum(lens, 40))
    out = {"workload": name, "pairs": pairs, "L": L, "k": k, "m": m, "records": int(R), "valid_windows": W,
           "gated_fraction": float(wgated.sum() / max(1, W)),
           "runs": n_runs, "windows_per_run": W / n_runs, "bytes_per_window_32B_runs": 32.0 * n_runs / W,
           "run_length_hist_1_to_40": hist[1:].tolist()}
    # balance: windows per 8-bit bucket, folded to P partitions, to G GPUs (p & (G-1))
    bw = np.bincount(bucket[wvalid], minlength=256).astype(np.float64)
    # k-mer hash comparison: hash of the window's k-mer (use 2 words)
    lo = np.zeros((R, w), np.uint64); hi = np.zeros((R, w), np.uint64)
    for j in range(k):
        if j < 32: lo |= b0[:, j:j + w] << np.uint64(2 * j)
        else: hi |= b0[:, j:j + w] << np.uint64(2 * (j - 32))
    with np.errstate(over="ignore"):
        h = lo ^ (hi * np.uint64(0x9E3779B97F4A7C15)); h ^= h >> np.uint64(32); h *= np.uint64(0xD6E8FEB86659FD93)
        h ^= h >> np.uint64(32); h *= np.uint64(0xD6E8FEB86659FD93); h ^= h >> np.uint64(32)
    kb = (h >> np.uint64(56)).astype(np.int64)
    kw = np.bincount(kb[wvalid], minlength=256).astype(np.float64)
    # distinct gated k-mers per bucket
    sel = wgated.ravel()
    keys = np.stack([lo.ravel()[sel], hi.ravel()[sel]], 1)
    uk, ui = np.unique(keys, axis=0, return_index=True)
    db_min = np.bincount(bucket.ravel()[sel][ui], minlength=256).astype(np.float64)
    db_k = np.bincount(kb.ravel()[sel][ui], minlength=256).astype(np.float64)
    out["distinct_gated_kmers"] = int(len(uk))
    def imb(v, groups):
        g = np.zeros(groups); 
        for b_ in range(256): g[b_ % groups] += v[b_]
        return float(g.max() / g.mean())
    def imbP(v, P):
        g = v.reshape(P, 256 // P).sum(1); return float(g.max() / g.mean())
    out["max_over_mean"] = {
        "windows_per_gpu8_minimizer": imb(bw, 8), "windows_per_gpu8_kmerhash": imb(kw, 8),
        "windows_per_gpu4_minimizer": imb(bw, 4), "windows_per_gpu2_minimizer": imb(bw, 2),
        "windows_per_P64_minimizer": imbP(bw, 64), "windows_per_P64_kmerhash": imbP(kw, 64),
        "distinct_per_P64_minimizer": imbP(db_min, 64), "distinct_per_P64_kmerhash": imbP(db_k, 64),
        "distinct_per_gpu8_minimizer": imb(db_min, 8), "distinct_per_P256_minimizer": imbP(db_min, 256)}
    print(json.dumps(out))

main()
