# bench under a few knob settings (env: VDJGRAPH_*)
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python profiles/bench_summary.py | grep -E "kernel_ms|slow"; }
run VDJGRAPH_L1_REFRESH=0
run VDJGRAPH_L1_REFRESH=1
run VDJGRAPH_L1_REFRESH=2
run VDJGRAPH_L1_REFRESH=2 VDJGRAPH_LOAD2=0.25
run VDJGRAPH_L1_REFRESH=2 VDJGRAPH_LOAD2=0.25 VDJGRAPH_LOAD1=0.33
run VDJGRAPH_L1_REFRESH=2 VDJGRAPH_LOAD2=0.25 VDJGRAPH_QFLUSH1=32
run VDJGRAPH_L1_REFRESH=0
