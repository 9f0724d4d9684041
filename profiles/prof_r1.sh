# Round-1 profile of record: launch list (gpu__time_duration) + one full capture of the main kernels.
# Run under gpurun on one B200:  bash profiles/prof_r1.sh
mkdir -p gpurun_out
# (k_pack, the staging kernel, runs once per 32768-record chunk and is left out: 611 launches per staging)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_count|k_scatter|k_init|k_pass|k_prune|k_build|k_collect|k_assign|k_export|DeviceRadix|DeviceScan' -c 400 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_r1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_scatter|k_count|k_pass1|k_pass2|k_export|k_prune' -s 6 -c 6 \
    -o gpurun_out/prof_r1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/prof_r1.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
tail -1 gpurun_out/bench_r1.json | python profiles/bench_summary.py
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r1_reference.json 2>> gpurun_out/bench_r1.err
tail -1 gpurun_out/bench_r1_reference.json | python profiles/bench_summary.py
# the other single-GPU BASELINE configs, for the record (parity for them: tests/test_parity_gpu.py full-size properties)
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload igh_sensitive_2x50_5M > gpurun_out/bench_r1_c3_sensitive.json 2>> gpurun_out/bench_r1.err
tail -1 gpurun_out/bench_r1_c3_sensitive.json | python profiles/bench_summary.py
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload igk_2x75_20M > gpurun_out/bench_r1_c4_igk_2x75_20M.json 2>> gpurun_out/bench_r1.err
tail -1 gpurun_out/bench_r1_c4_igk_2x75_20M.json | python profiles/bench_summary.py
# configs[4]'s per-GPU share at 8 GPUs (50 M pairs 2x100 = 13.2 G windows, 211 GB of tuples): super-partition rounds on one B200
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload pooled_2x100_50M_per_gpu > gpurun_out/bench_r1_c5_50M_per_gpu.json 2>> gpurun_out/bench_r1.err
tail -1 gpurun_out/bench_r1_c5_50M_per_gpu.json | python profiles/bench_summary.py
# the bounce path (pageable record buffers, as the unmodified bam_read.c allocates them)
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --pageable > gpurun_out/bench_r1_pageable.json 2>> gpurun_out/bench_r1.err
tail -1 gpurun_out/bench_r1_pageable.json | python profiles/bench_summary.py | head -1
