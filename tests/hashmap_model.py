"""TEST INFRASTRUCTURE: a sequential model of the layout of the reference's `nodes` map
(dense_hash_map<const char*, node*, my_hash, eqstr>, assembler2_vdj.c:1368-1369) after build_graph2
has inserted the nodes in creation order (:305).  The library's hashmap-layout export
(VDJGRAPH_FLAG_HASHMAP_LAYOUT) must reproduce it bucket for bucket; this model is itself pinned
against the real container (oracle.loader.reference_iteration_order, tests/test_hashmap_layout.py).

  hash     MurmurHash64A(kmer, kmer_size, 97) (hash_utils.c:5-46, hash_utils.h:21-28), then
           hash / sizeof(void*) because the key type is a pointer (sparsehash
           internal/hashtable-common.h:352-361, hash_munger<HashKey*>)
  table    32 buckets (HT_DEFAULT_STARTING_BUCKETS, densehashtable.h:308); an insert that would make
           the element count exceed buckets / 2 first rebuilds the table at twice the size by
           re-inserting the old table's elements IN BUCKET ORDER (resize_delta :631-653, copy_from)
  probing  bucket = (hash + 1 + 2 + ... + i) & (buckets - 1) on the i-th collision (JUMP_ :119)
"""
from __future__ import annotations

import numpy as np

M64 = (1 << 64) - 1


def murmur64a(key: bytes, seed: int = 97) -> int:
    m, r = 0xC6A4A7935BD1E995, 47
    n = len(key)
    h = (seed ^ (n * m)) & M64
    for i in range(n // 8):
        k = int.from_bytes(key[8 * i:8 * i + 8], "little")
        k = (k * m) & M64
        k ^= k >> r
        k = (k * m) & M64
        h ^= k
        h = (h * m) & M64
    tail = key[8 * (n // 8):]
    if tail:
        h ^= int.from_bytes(tail, "little")
        h = (h * m) & M64
    h ^= h >> r
    h = (h * m) & M64
    h ^= h >> r
    return h


def layout(kmer_codes: np.ndarray):
    """kmer_codes: [n, k] 2-bit codes (A0 C1 G2 T3) of the nodes in creation order.
    Returns (slots, buckets): slots[b] = node in bucket b or -1."""
    asc = np.frombuffer(b"ACGT", np.uint8)[kmer_codes]
    hashes = [murmur64a(asc[i].tobytes()) >> 3 for i in range(len(asc))]
    nb, table, cnt = 32, [-1] * 32, 0

    def put(tab, mask, e):
        b, probes = hashes[e] & mask, 0
        while tab[b] != -1:
            probes += 1
            b = (b + probes) & mask
        tab[b] = e

    for e in range(len(hashes)):
        if cnt + 1 > nb // 2:
            nb *= 2
            fresh = [-1] * nb
            for x in table:
                if x != -1:
                    put(fresh, nb - 1, x)
            table = fresh
        put(table, nb - 1, e)
        cnt += 1
    return np.array(table, np.int64), nb
