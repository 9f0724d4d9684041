# multi-GPU parity + bench (argument: number of GPUs)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m | head -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_parity.py 2>&1 | grep -v "^W1\|^\*\*\*\|OMP_NUM" | tail -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bm$N.log 2>&1
tail -1 gpurun_out/bm$N.log | python profiles/bench_summary.py || tail -30 gpurun_out/bm$N.log
