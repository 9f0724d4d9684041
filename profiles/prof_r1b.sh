set -x
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b3.log 2>&1
tail -1 gpurun_out/b3.log | python profiles/bench_summary.py
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_pass2|k_pass1|k_scatter|k_count' -s 4 -c 4 -o gpurun_out/prof_r1b python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b5.log 2>&1
ls -la gpurun_out
