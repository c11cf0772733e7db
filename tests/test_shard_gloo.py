"""CPU, world_size=2 over gloo: the multi-GPU plumbing (contiguous sharding in input order, rebased read buffers,
max/sum reductions of timings and cell counts, in-order gather).  The per-shard compute is the CPU oracle standing in for a
GPU (there is none here); the merged result must equal the single-process result byte for byte."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from bbmap_b200 import shard, workloads as wl
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    genome = wl.random_genome(20000, seed=3)
    reads, tasks = wl.make_msa_tasks(genome, 301, seed=4, flags=wl.TF_SCORE)
    r_loc, t_loc = shard.shard_msa_tasks(reads, tasks, rank, world)
    outs, _, cells = orc.get().run_batch(r_loc, genome, t_loc)
    ms = shard.max_over_ranks([10.0 + rank, 5.0 - rank])
    tot = shard.sum_over_ranks([float(cells), float(len(t_loc))])
    merged = shard.gather_records(outs, dst=0)
    if rank == 0:
        q.put((merged.tobytes(), ms, tot))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding(oracle):
    from bbmap_b200 import shard, workloads as wl
    for n in (0, 1, 7, 301):
        for w in (1, 2, 3, 8):
            b = [shard.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
    genome = wl.random_genome(20000, seed=3)
    reads, tasks = wl.make_msa_tasks(genome, 301, seed=4, flags=wl.TF_SCORE)
    exp, _, cells = oracle.run_batch(reads, genome, tasks)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, ms, tot = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert merged == exp.tobytes()
    assert ms == [11.0, 5.0]
    assert tot == [float(cells), 301.0]
