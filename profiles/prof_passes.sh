mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_scatter|k_count|k_pass1|k_pass2|k_export|k_prune' -s 6 -c 6 -o gpurun_out/prof_r1g python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b7.log 2>&1
