# quick GPU check: parity tests + a short bench (no ncu)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bq.log 2>&1
tail -1 gpurun_out/bq.log | python profiles/bench_summary.py || tail -20 gpurun_out/bq.log
