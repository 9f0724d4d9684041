mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_pass1|k_pass2' -s 2 -c 2 -o gpurun_out/prof_r1f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b7.log 2>&1
