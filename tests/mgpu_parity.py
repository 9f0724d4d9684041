"""Real multi-GPU parity check, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_parity.py

Every rank generates the same seeded records, keeps its contiguous share, and the sharded build
(tuples written into the owners' buffers over NVLink through CUDA IPC mappings) must reproduce the
oracle's graph of the WHOLE input bit for bit on rank 0.  Not collected by pytest (needs torchrun);
the same phases are covered on one GPU by tests/test_shard_gpu.py."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from vdjer_b200 import GraphBuilder, shard, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    # last column: super-partition rounds (0 = auto = 1 here); with more than one round every rank
    # scatters and runs its passes once per round, a barrier apart
    for (L, k, mf, mq, pairs, clones, seed, rounds) in [(50, 35, 3, 90, 200000, 3000, 301, 0), (50, 25, 1, 20, 40000, 300, 302, 4),
                                                        (100, 50, 2, 120, 30000, 500, 303, 2)]:
        primary, secondary = synth.generate(n_pairs=pairs, read_length=L, seed=seed, n_clones=clones, threads=4)
        rb = 2 * L + 1
        total = primary.size // rb + secondary.size // rb
        lo, hi = shard.shard_ranges(total, world)[rank]
        p, s = shard.split_records(primary, secondary, L, lo, hi)
        gb = GraphBuilder(L, k, mf, mq, device=local, rounds=rounds)
        db = shard.DistributedBuilder(gb, dist, device=f"cuda:{local}")
        g = db.build(p, s)
        t0 = time.perf_counter()
        for _ in range(3):
            db.run()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        if rank == 0:
            from oracle import loader
            from tests.util import assert_graph_equal
            want = loader.build(primary, secondary, L, k, mf, mq, kind="port")
            try:
                assert_graph_equal(g, want, f"world={world} L={L} k={k}")
                print(f"PARITY OK world={world} L={L} k={k} mf={mf} mq={mq} rounds={g.stats['rounds']}: {g.n_nodes} nodes, run {dt * 1e3:.2f} ms", flush=True)
            except AssertionError as e:
                ok = False
                print("PARITY FAIL", e, flush=True)
        db.close()
        gb.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
