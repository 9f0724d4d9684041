# bench under a few knob settings (env: VDJGRAPH_*)
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" python bench.py --steps 4 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python profiles/bench_summary.py | grep -E "kernel_ms|slow"; }
run VDJGRAPH_QFLUSH1=1
run VDJGRAPH_QFLUSH1=8
run VDJGRAPH_QFLUSH1=16
run VDJGRAPH_QFLUSH1=24
run VDJGRAPH_QFLUSH1=32
run VDJGRAPH_QFLUSH1=16 VDJGRAPH_QFLUSH2=32
run VDJGRAPH_QFLUSH1=16 VDJGRAPH_QFLUSH2=64 VDJGRAPH_QDENSE2=16
