"""Worker of tests/test_gpu_multi.py: run under torchrun with one rank per GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from strainflair_b200 import dist as sfdist   # noqa: E402
from strainflair_b200 import engine, synth    # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    graph, reads = synth.make_config("small", scale=0.4)
    g0, g1 = sfdist.shard_bounds(reads.n_groups, world, rank)
    eng = engine.AbundanceEngine(graph, dev, comm=dist.group.WORLD, deterministic=False)
    res = eng.run(reads.slice_groups(g0, g1))
    if rank == 0:
        single = engine.AbundanceEngine(graph, dev, deterministic=False).run(reads)
        for k in ("NB_READS", "PAIRS", "SINGLES", "FILTERED", "UNASSIGNED", "ASSIGNED_U", "ASSIGNED_M", "INCONSIST",
                  "MULTI_ALIGN", "SE", "SAME_CP", "DIFF_CP", "SE_M", "SECOND_PASS", "AMBIG_CP_RESOLVED"):
            assert res.C[k] == single.C[k], (k, res.C[k], single.C[k])
        assert np.array_equal(res.uniq_reads, single.uniq_reads)
        assert np.array_equal(res.hist, single.hist)
        assert np.array_equal(res.mult_hits > 0, single.mult_hits > 0)
        for name in ("uniq_norm", "mult_reads", "mult_norm", "uniq_abund", "mult_abund", "table"):
            a, b = getattr(res, name), getattr(single, name)
            ok = np.abs(a - b) <= 1e-9 * np.maximum(np.abs(a), np.abs(b))
            ok |= np.isnan(a) & np.isnan(b)
            assert ok.all(), name
        assert np.array_equal(res.uniq_abund > 0, single.uniq_abund > 0)
        print(f"MGPU_OK world={world} groups={reads.n_groups}")
    # order-independent mode: bit-identical whatever the number of GPUs
    eng_d = engine.AbundanceEngine(graph, dev, comm=dist.group.WORLD, deterministic=True)
    res_d = eng_d.run(reads.slice_groups(g0, g1))
    if rank == 0:
        single_d = engine.AbundanceEngine(graph, dev, deterministic=True).run(reads)
        for name in ("uniq_norm", "mult_reads", "mult_norm", "uniq_abund", "mult_abund", "table"):
            assert np.array_equal(getattr(res_d, name), getattr(single_d, name), equal_nan=True), name
        print("MGPU_DET_OK")
    # ... and whatever the kernels: the shared-memory tile kernels on the shards against the default on one GPU
    res_t = engine.AbundanceEngine(graph, dev, comm=dist.group.WORLD, tile_kernels=True).run(reads.slice_groups(g0, g1))
    if rank == 0:
        for name in ("uniq_reads", "hist", "uniq_norm", "mult_reads", "mult_norm", "uniq_abund", "mult_abund", "table"):
            assert np.array_equal(getattr(res_t, name), getattr(single_d, name), equal_nan=True), name
        print("MGPU_TILE_OK")
    # two pipelined jobs per rank (submit / collect), each rank's shard staged as two host shards: same bits again
    mine = reads.slice_groups(g0, g1)
    staged = [engine.HostReads(mine.slice_groups(*sfdist.shard_bounds(mine.n_groups, 2, i))) for i in range(2)]
    j1, j2 = eng_d.submit(staged, depth=2), eng_d.submit(staged, depth=2)
    outs = [eng_d.collect(j1)]
    j3 = eng_d.submit(staged, depth=2)
    outs += [eng_d.collect(j2), eng_d.collect(j3)]
    for out in outs:
        assert out.C["NB_READS"] == res_d.C["NB_READS"]
        for name in ("uniq_norm", "mult_reads", "mult_norm", "table"):
            assert np.array_equal(getattr(out, name), getattr(res_d, name), equal_nan=True), name
    if rank == 0:
        print("MGPU_PIPE_OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
