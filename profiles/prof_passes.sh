mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_pass1' -s 1 -c 1 -o gpurun_out/prof_r1e python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b7.log 2>&1
