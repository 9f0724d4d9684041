mkdir -p gpurun_out
run() { echo "== $*"; env "$@" python bench.py --steps 4 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python profiles/bench_summary.py | grep -E "value|kernel_ms"; }
run X=0
