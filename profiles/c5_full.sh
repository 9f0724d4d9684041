# BASELINE configs[4] at FULL size (400 M pairs 2x100, 1.06e11 windows) on 8 B200s -- NOT RUN in round 1
# (it needs most of a round's GPU budget).  Every rank generates its 50 M pairs as forward reads only
# (20 GB of host text per rank instead of 40), stages them with vdjgraph_shard_stage_forward (the
# reverse-complement records are derived on the device) and the build walks the hash space in rounds
# (auto: 4 on a 180 GB device).  Run under `gpurun --gpus 8`.
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 \
    bench.py --gpus 8 --workload pooled_2x100_50M_per_gpu --forward-inputs --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/bench_c5_full_8gpu.json 2> gpurun_out/bench_c5_full_8gpu.err
tail -1 gpurun_out/bench_c5_full_8gpu.json | python profiles/bench_summary.py || tail -30 gpurun_out/bench_c5_full_8gpu.err
