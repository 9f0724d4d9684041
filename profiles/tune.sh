# parity tests, then bench under a few knob settings (env: VDJGRAPH_LOAD1/LOAD2/SLICE_MB)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { echo "== $*"; env "$@" python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python profiles/bench_summary.py; }
run X=1
run VDJGRAPH_LOAD1=0.33 VDJGRAPH_LOAD2=0.33
run VDJGRAPH_LOAD1=0.25 VDJGRAPH_LOAD2=0.25
run VDJGRAPH_LOAD1=0.33 VDJGRAPH_LOAD2=0.33 VDJGRAPH_SLICE_MB=12
run VDJGRAPH_LOAD1=0.33 VDJGRAPH_LOAD2=0.33 VDJGRAPH_SLICE_MB=48
