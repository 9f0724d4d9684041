mkdir -p gpurun_out
run() { echo "== $*"; env "$@" python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python profiles/bench_summary.py | grep -E "kernel_ms"; }
run VDJGRAPH_DBG=0
run VDJGRAPH_DBG=32
run VDJGRAPH_DBG=64
run VDJGRAPH_DBG=96
