"""CPU: workload generator and record-format helpers."""
import numpy as np

from tests.util import kmer_codes_at, pack_codes
from vdjer_b200 import synth


def test_generator_is_deterministic_and_thread_independent(built):
    a = synth.generate(n_pairs=500, read_length=50, seed=5, n_clones=20, threads=1)
    b = synth.generate(n_pairs=500, read_length=50, seed=5, n_clones=20, threads=4)
    c = synth.generate(n_pairs=500, read_length=50, seed=6, n_clones=20, threads=4)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert not np.array_equal(a[0], c[0])


def test_record_format(built):
    L = 75
    p, s = synth.generate(n_pairs=200, read_length=L, seed=1, n_clones=5, frac_secondary=0.25, threads=2)
    rec = 2 * L + 1
    assert p[-1] == 0 and s[-1] == 0
    assert (p.size - 1) % rec == 0 and (p.size - 1) // rec == 600 and (s.size - 1) // rec == 200
    r = p[:-1].reshape(-1, rec)
    assert np.all(r[:, 0] == ord("0"))
    assert set(np.unique(r[:, 1:1 + L])) <= set(b"ACGTN")
    assert r[:, 1 + L:].min() >= 33
    # record 2i+1 is the reverse complement of record 2i with reversed qualities (bam_read.c:206-244)
    comp = np.zeros(256, np.uint8)
    for a, b in zip(b"ACGTN", b"TGCAN"):
        comp[a] = b
    assert np.array_equal(comp[r[0::2, 1:1 + L]][:, ::-1], r[1::2, 1:1 + L])
    assert np.array_equal(r[0::2, 1 + L:][:, ::-1], r[1::2, 1 + L:])


def test_pack_helpers():
    buf = synth.records_from_reads(["ACGTACGTACGTACGTACGTACGTACGTACGTACGTA"], both_strands=False)
    codes = kmer_codes_at(buf, np.zeros(1, np.uint8), 37, 35, [0, 2])
    lo, hi = pack_codes(codes)
    assert codes.shape == (2, 35) and list(codes[0][:4]) == [0, 1, 2, 3] and list(codes[1][:2]) == [2, 3]
    assert int(lo[0]) & 0xFF == 0b11100100 and int(hi[0]) == (0 | (1 << 2) | (2 << 4))


def test_forward_reads_helper_and_oracle_equivalence(built):
    """forward_reads() keeps the even records; rebuilding the doubled buffer from them (what
    k_pack does on the device for vdjgraph_stage_forward: bam_read.c:130-145, 230-243) gives back the
    generator's buffer byte for byte, so the oracle sees the same input either way."""
    import pytest

    from vdjer_b200 import forward_reads
    L = 50
    rb = 2 * L + 1
    p, s = synth.generate(n_pairs=300, read_length=L, seed=9, n_clones=10, threads=2)
    for buf in (p, s):
        f = forward_reads(buf, L)
        assert f[-1] == 0 and (f.size - 1) * 2 == buf.size - 1
        fr = f[:-1].reshape(-1, rb)
        assert np.array_equal(fr, buf[:-1].reshape(-1, rb)[0::2])
        comp = np.arange(256, dtype=np.uint8)
        for a, b in zip(b"ACGT", b"TGCA"):
            comp[a] = b
        rc = np.concatenate([fr[:, :1], comp[fr[:, 1:1 + L]][:, ::-1], fr[:, 1 + L:][:, ::-1]], axis=1)
        doubled = np.stack([fr, rc], axis=1).reshape(-1)
        assert np.array_equal(doubled, buf[:-1])
    assert forward_reads(np.zeros(1, np.uint8), L).size == 1          # empty buffer: just the NUL
    with pytest.raises(ValueError):
        forward_reads(p[: 3 * rb], L)                                  # odd number of records


def test_generator_forward_only_is_the_even_records(built):
    """forward_only=True writes the reads once: exactly the even records of the default output, for
    any thread count and pair offset."""
    from vdjer_b200 import forward_reads
    for L, n, thr in [(50, 1500, 3), (100, 333, 5)]:
        p, s = synth.generate(n_pairs=n, read_length=L, seed=11, n_clones=40, threads=thr, pair_offset=77)
        fp, fs = synth.generate(n_pairs=n, read_length=L, seed=11, n_clones=40, threads=2, pair_offset=77, forward_only=True)
        assert np.array_equal(fp, forward_reads(p, L)) and np.array_equal(fs, forward_reads(s, L))
