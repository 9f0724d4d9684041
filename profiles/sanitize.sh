# compute-sanitizer over one small build (memcheck, then racecheck on shared memory)
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys
sys.path.insert(0, ".")
from vdjer_b200 import GraphBuilder, synth
from oracle import loader
from tests.util import assert_graph_equal
for (L, k, mf, mq) in [(50, 35, 3, 90), (100, 50, 2, 120)]:
    p, s = synth.generate(n_pairs=3000, read_length=L, seed=9, n_clones=60, threads=2)
    want = loader.build(p, s, L, k, mf, mq, kind="port")
    # one round with the map layout, and a multi-round build (merged finish)
    for kw in (dict(hashmap_layout=True), dict(rounds=2)):
        with GraphBuilder(L, k, mf, mq, **kw) as gb:
            g = gb.build(p, s)
        assert_graph_equal(g, want, "sanitizer run")
        print("ok", L, k, kw, g.n_nodes)
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python /tmp/san.py > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python /tmp/san.py > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/racecheck.log
