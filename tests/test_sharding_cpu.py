"""N > 1 host logic on CPU (gloo, world_size 2): the path shards by stereo pair with no data-path
collective; ranks agree on the partition, the max-over-ranks timing reduction and the whole-job
throughput arithmetic that bench.py reports."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from stereoscene_b200 import sharding, synth
    pairs = sharding.pairs_for_rank(total_pairs=6, rank=rank, world=world)
    xl, _ = synth.stereo_features(len(pairs), (64, 128), 8, seed=100 + rank)
    # every rank times its own work; the job time is the max
    ms = torch.tensor([10.0 + 3.0 * rank], dtype=torch.float64)
    job_ms = sharding.max_over_ranks(ms)
    value = sharding.whole_job_voxels_per_s(voxels_per_pair=32 * 32 * 8, pairs_per_rank=len(pairs), world=world,
                                            steps=4, ms_total=float(job_ms))
    gathered = [None] * world
    dist.all_gather_object(gathered, (pairs, float(xl.sum())))
    q.put((rank, pairs, float(job_ms), value, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_pair_sharding_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, p0, t0, v0, g0), (r1, p1, t1, v1, g1) = out
    assert sorted(p0 + p1) == [0, 1, 2, 3, 4, 5] and not set(p0) & set(p1)       # disjoint cover
    assert t0 == t1 == 13.0                                                   # max over ranks
    assert v0 == v1
    assert g0 == g1 and g0[0][1] != g0[1][1]                                  # different data per rank


def test_sharding_arithmetic():
    sys.path.insert(0, ROOT)
    from stereoscene_b200 import sharding
    assert sharding.pairs_for_rank(8, 3, 8) == [3]
    assert sum(len(sharding.pairs_for_rank(10, r, 4)) for r in range(4)) == 10
    assert sharding.whole_job_voxels_per_s(2097152, 1, 8, 10, 1000.0) == 2097152 * 8 * 10 / 1.0
