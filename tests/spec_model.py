"""Order-free restatement of the hot path in plain Python: the exact
reductions the CUDA kernels implement (counts, min-stamps, distinct-read test, first-NB occurrence
log, closed-form quality test), without any sequential state.  Small inputs only; used to check
the kernel *design* against the sequential oracle on CPU."""
from __future__ import annotations

import numpy as np


def order_free_build(primary, secondary, L, k, mf, mq):
    rec = 2 * L + 1
    w = L - k + 1
    buf = bytes(np.asarray(primary, np.uint8)[: (len(primary) // rec) * rec]) + \
        bytes(np.asarray(secondary, np.uint8)[: (len(secondary) // rec) * rec])
    R = len(buf) // rec
    seqs = [buf[r * rec + 1: r * rec + 1 + L] for r in range(R)]
    quals = [np.frombuffer(buf[r * rec + 1 + L: (r + 1) * rec], np.uint8).astype(np.int64) - 33 for r in range(R)]
    strand = [buf[r * rec] for r in range(R)]

    T = min(min(mq, 254), 214)
    NB = (T + 19) // 20 if T > 0 else 0
    p1 = {}
    n_gated = 0
    for r in range(R):
        s, q = seqs[r], quals[r]
        for i in range(w):
            km = s[i:i + k]
            if b"N" in km or (q[i:i + k] & 0xFF).min() < 20:
                continue
            n_gated += 1
            e = p1.setdefault(km, dict(cnt=0, reads=set(), occ=[]))
            e["cnt"] += 1
            e["reads"].add((s, strand[r]))
            e["occ"].append((r * w + i))
    surv = set()
    for km, e in p1.items():
        cnt = min(e["cnt"], 32765)
        if cnt < mf or len(e["reads"]) < 2:
            continue
        ok = True
        if 20 * (cnt - 1) < T:
            occ = sorted(e["occ"])[:NB]
            assert len(occ) == cnt
            S = np.zeros(k, np.int64)
            for n, st in enumerate(occ):
                r, i = divmod(st, w)
                S += quals[r][0:k] if n == 0 else quals[r][i:i + k]
            ok = bool((S >= T).all())
        if ok:
            surv.add(km)

    INF = 1 << 62
    p2 = {km: dict(cnt=0, first=INF, out={}) for km in surv}
    for r in range(R):
        s = seqs[r]
        for i in range(w):
            km = s[i:i + k]
            e = p2.get(km)
            if e is None:
                continue
            st = r * w + i
            e["cnt"] += 1
            e["first"] = min(e["first"], st)
            if i + 1 < w and s[i + k:i + k + 1] in (b"A", b"C", b"G", b"T"):
                c = s[i + k:i + k + 1]
                e["out"][c] = min(e["out"].get(c, INF), st)
    order = sorted(p2, key=lambda km: p2[km]["first"])
    rank = {km: n for n, km in enumerate(order)}
    n = len(order)
    out = dict(first_pos=np.zeros(n, np.uint64), frequency=np.zeros(n, np.uint16), out_deg=np.zeros(n, np.uint8),
               in_deg=np.zeros(n, np.uint8), out_succ=np.full((n, 4), 0xFFFFFFFF, np.uint32),
               in_pred=np.full((n, 4), 0xFFFFFFFF, np.uint32), n_pre_total=len(p1), n_gated=n_gated)
    for km in order:
        e, i = p2[km], rank[km]
        out["first_pos"][i] = e["first"]
        out["frequency"][i] = min(e["cnt"], 32765)
        succ = [(t, rank[km[1:] + c]) for c, t in e["out"].items() if km[1:] + c in p2]
        succ.sort(reverse=True)
        out["out_deg"][i] = len(succ)
        for d, (_, v) in enumerate(succ):
            out["out_succ"][i, d] = v
        pred = []
        for c in (b"A", b"C", b"G", b"T"):
            pk = c + km[:-1]
            if pk in p2 and km[-1:] in p2[pk]["out"]:
                pred.append((p2[pk]["out"][km[-1:]], rank[pk]))
        pred.sort(reverse=True)
        out["in_deg"][i] = len(pred)
        for d, (_, v) in enumerate(pred):
            out["in_pred"][i, d] = v
    return out
