mkdir -p gpurun_out
run() { echo "== $*"; env "$@" python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python profiles/bench_summary.py | grep -E "kernel_ms"; }
run VDJGRAPH_HOT_T=5 VDJGRAPH_HOT_FLUSH=1
run VDJGRAPH_HOT_T=5 VDJGRAPH_HOT_FLUSH=2
run VDJGRAPH_HOT_T=5 VDJGRAPH_HOT_FLUSH=3
run VDJGRAPH_HOT_T=64 VDJGRAPH_HOT_FLUSH=2
run VDJGRAPH_HOT_T=1024 VDJGRAPH_HOT_FLUSH=3
run VDJGRAPH_HOT_T=2000000000
