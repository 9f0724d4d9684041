"""GPU: vdjgraph_multi_* -- the sharded build as ONE call of a single process (one host thread per device
inside the call), which is how a single-process caller like V'DJer would use several GPUs.  Here the
"devices" are the same GPU named several times (thread barriers between the phases and between the steps of
the finish); the kernels and the phase order are those of the multi-process build (tests/mgpu_parity.py)."""
import numpy as np
import pytest

from oracle import loader
from tests.util import assert_graph_equal
from vdjer_b200 import GraphBuilder, MultiBuilder, VdjGraphError, forward_reads, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("G", [2, 4])
def test_multi_build_in_one_process_equals_oracle(built, G):
    L, k, mf, mq = 50, 35, 3, 90
    primary, secondary = synth.generate(n_pairs=20000, read_length=L, seed=401, n_clones=300, threads=4)
    want = loader.build(primary, secondary, L, k, mf, mq, kind="port")
    with MultiBuilder(L, k, mf, mq, devices=[0] * G) as mb:
        got = mb.build(primary, secondary)
        assert_graph_equal(got, want, f"multi G={G}")
        stats = [mb.rank_stats(r) for r in range(G)]
        assert sum(s["n_pre_total"] for s in stats) == want["n_pre_total"]      # every k-mer has exactly one owner
        # the same contexts again, from forward reads only
        again = mb.build_forward(forward_reads(primary, L), forward_reads(secondary, L))
        assert_graph_equal(again, want, f"multi forward G={G}")
    with GraphBuilder(L, k, mf, mq) as gb:
        single = gb.build(primary, secondary)
    for name in ["first_pos", "frequency", "out_deg", "in_deg", "out_succ", "in_pred"]:
        assert np.array_equal(getattr(got, name), getattr(single, name)), name


def test_multi_build_reports_a_failing_rank(built):
    """A record with a bad strand byte lands in ONE rank's range: that rank fails, the others stop with it,
    the call returns that rank's status and message."""
    L = 50
    primary, secondary = synth.generate(n_pairs=4000, read_length=L, seed=402, n_clones=40, threads=2)
    bad = np.array(primary, copy=True)
    rb = 2 * L + 1
    bad[(bad.size // rb - 3) * rb] = ord("x")          # strand byte of a record near the end of the primary buffer
    with MultiBuilder(L, 35, 2, 60, devices=[0, 0]) as mb:
        with pytest.raises(VdjGraphError) as e:
            mb.build(bad, secondary)
        assert e.value.code == -2 and "rank" in str(e.value)
        ok = mb.build(primary, secondary)                 # the contexts are usable afterwards
        assert ok.n_nodes == loader.build(primary, secondary, L, 35, 2, 60, kind="port")["n_nodes"]


def test_multi_create_rejects_bad_device_lists(built):
    with pytest.raises(VdjGraphError) as e:
        MultiBuilder(50, devices=[0, 0, 0])
    assert e.value.code == -1
