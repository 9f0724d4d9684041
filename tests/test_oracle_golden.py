"""CPU: the oracle's C restatement (oracle/vdj_oracle.c) against the committed outputs of the
compiled reference (tests/golden/*.npz, made by tests/golden/make_golden.py), and -- when the
compiled reference itself is present (build container) -- against fresh runs of it."""
import os

import numpy as np
import pytest

from oracle import loader
from tests.cases import CASES, make_inputs
from tests.util import GOLDEN_DIR, sha

ARRAYS = ["pre_first_pos", "pre_freq", "pre_qual_sums", "first_pos", "frequency", "out_deg", "out_succ",
          "in_deg", "in_pred"]


@pytest.mark.parametrize("name", sorted(CASES))
def test_port_matches_reference_golden(built, name):
    case = CASES[name]
    gold = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    primary, secondary = make_inputs(case)
    assert sha(primary, secondary) == str(gold["input_sha256"]), "input generator drifted from the fixture"
    if "primary" in gold:
        assert np.array_equal(primary, gold["primary"]) and np.array_equal(secondary, gold["secondary"])
    got = loader.build(primary, secondary, case["L"], case["k"], case["mf"], case["mq"], kind="port")
    assert got["n_pre_total"] == int(gold["n_pre_total"])
    for a in ARRAYS:
        assert np.array_equal(got[a], gold[a]), f"{name}: {a} differs from the reference"


@pytest.mark.skipif(not loader.have_reference(), reason="compiled reference (oracle/_ref) not built here")
@pytest.mark.parametrize("seed,L,k,mf,mq", [(1, 50, 35, 3, 90), (2, 50, 25, 2, 60), (3, 75, 35, 3, 90),
                                             (4, 100, 35, 3, 90), (5, 50, 25, 1, 20), (6, 60, 41, 2, 254)])
def test_port_matches_live_reference(built, seed, L, k, mf, mq):
    from vdjer_b200 import synth
    primary, secondary = synth.generate(n_pairs=4000, read_length=L, seed=seed, n_clones=50 + 10 * seed, threads=2)
    a = loader.build(primary, secondary, L, k, mf, mq, kind="port")
    b = loader.build(primary, secondary, L, k, mf, mq, kind="reference")
    assert loader.diff(a, b) == []


def test_reference_flag_variants_agree(built):
    """The reference built with its own flags (-g) and with -O2 give identical tables."""
    if not (loader.have_reference() and os.path.exists(loader.REF_LIB_G)):
        pytest.skip("compiled reference not built here")
    from vdjer_b200 import synth
    primary, secondary = synth.generate(n_pairs=1500, read_length=50, seed=9, n_clones=30, threads=2)
    a = loader.build(primary, secondary, 50, 35, 3, 90, kind="reference", variant="O2")
    b = loader.build(primary, secondary, 50, 35, 3, 90, kind="reference", variant="g")
    assert loader.diff(a, b) == []


def test_order_free_spec_matches_oracle(built):
    """the commutative-reduction form the kernels implement, restated in
    numpy/python on a small case, equals the sequential oracle."""
    from tests.spec_model import order_free_build
    from vdjer_b200 import synth
    for seed, L, k, mf, mq in [(21, 40, 21, 2, 60), (22, 50, 35, 3, 90), (23, 40, 15, 1, 214)]:
        primary, secondary = synth.generate(n_pairs=600, read_length=L, seed=seed, n_clones=12, threads=1)
        want = loader.build(primary, secondary, L, k, mf, mq, kind="port")
        got = order_free_build(primary, secondary, L, k, mf, mq)
        for a in ["first_pos", "frequency", "out_deg", "out_succ", "in_deg", "in_pred"]:
            assert np.array_equal(got[a], want[a]), (seed, a)
        assert got["n_pre_total"] == want["n_pre_total"] and got["n_gated"] == want["n_gated"]
