# Round-2 profile of record: launch list (gpu__time_duration) + one full capture of the main kernels
# + the bench lines of every single-GPU BASELINE config.  Run under gpurun on one B200:
#     bash profiles/prof_r2.sh        then here:  ncu -i gpurun_out/prof_r2.ncu-rep --page raw --csv > gpurun_out/prof_r2_raw.csv
#                                                  python profiles/summarize.py r2
mkdir -p gpurun_out
# (k_pack and k_count, the staging kernels, run once per chunk and are left out of the launch list)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_scatter|k_init|k_pass|k_prune|k_build|k_collect|k_assign|k_export|k_unpack|DeviceRadix|DeviceScan' -c 400 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-forward > gpurun_out/launches_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_scatter|k_pass1|k_pass2|k_export|k_prune' -s 5 -c 5 \
    -o gpurun_out/prof_r2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-forward > gpurun_out/prof_r2.log 2>&1
# k_count alone (it runs per staging chunk, a quarter of the machine each): the first launches of a run
ncu --set full --clock-control none -k regex:'k_count' -c 2 -o gpurun_out/prof_r2_count python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-forward > /dev/null 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
tail -1 gpurun_out/bench_r2.json | python profiles/bench_summary.py
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r2_reference.json 2>> gpurun_out/bench_r2.err
tail -1 gpurun_out/bench_r2_reference.json | python profiles/bench_summary.py
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload igh_sensitive_2x50_5M > gpurun_out/bench_r2_c3_sensitive.json 2>> gpurun_out/bench_r2.err
tail -1 gpurun_out/bench_r2_c3_sensitive.json | python profiles/bench_summary.py
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload igk_2x75_20M > gpurun_out/bench_r2_c4_igk_2x75_20M.json 2>> gpurun_out/bench_r2.err
tail -1 gpurun_out/bench_r2_c4_igk_2x75_20M.json | python profiles/bench_summary.py
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --pageable > gpurun_out/bench_r2_pageable.json 2>> gpurun_out/bench_r2.err
tail -1 gpurun_out/bench_r2_pageable.json | python profiles/bench_summary.py | head -2
