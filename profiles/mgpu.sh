# multi-GPU parity + bench (argument: number of GPUs; second argument: "noparity" to skip the parity script;
# third argument "hostbar": the bench once more with host barriers in the finish)
N=${1:-2}
mkdir -p gpurun_out
if [ "$2" != "noparity" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_parity.py 2>&1 | grep -v "^W1\|^\*\*\*\|OMP_NUM" | tail -6 | tee gpurun_out/mgpu_parity_r2_${N}gpu.txt
fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bm$N.log 2>&1
tail -1 gpurun_out/bm$N.log | python profiles/bench_summary.py || tail -30 gpurun_out/bm$N.log
if [ "$3" == "hostbar" ]; then   # the same with host barriers between the steps of the finish
VDJGRAPH_HOST_BARRIERS=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bm${N}_hostbar.log 2>&1
tail -1 gpurun_out/bm${N}_hostbar.log | python profiles/bench_summary.py || tail -30 gpurun_out/bm${N}_hostbar.log
fi
