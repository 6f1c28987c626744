"""CPU (gloo, world size 2): the sharding and exchange protocol of the multi-GPU path
(strainflair_b200/dist.py), with the oracle standing in for the kernels on each rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import abundance as oa
from strainflair_b200 import dist as sfdist
from strainflair_b200 import synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    graph, reads = synth.make_config("tiny")
    g0, g1 = sfdist.shard_bounds(reads.n_groups, world, rank)
    shard = reads.slice_groups(g0, g1)
    orc = oa.AbundanceOracle(graph.as_dict(), shard.as_dict())
    orc.run_pass1()
    # exchange 1: exact integer sum of the per-path unique counts
    uniq = torch.tensor(orc.uniq_reads, dtype=torch.int64)
    sfdist.all_reduce_sum(uniq, None)
    orc.run_pass2(uniq_reads=[int(v) for v in uniq])
    # exchange 2: slot arrays onto the path-partitioned layout, table rows gathered
    path_off_slots = graph.path_off + np.arange(graph.n_paths + 1)
    pb = sfdist.path_partition(path_off_slots, world)
    plain_bounds = graph.path_off[pb]
    ua, ma = torch.from_numpy(orc.uniq_abund.copy()), torch.from_numpy(orc.mult_abund.copy())
    sfdist.reduce_scatter_ranges(ua, plain_bounds, None)
    sfdist.reduce_scatter_ranges(ma, plain_bounds, None)
    lo, hi = int(plain_bounds[rank]), int(plain_bounds[rank + 1])
    rows = torch.stack([torch.tensor([float(ua[graph.path_off[p]:graph.path_off[p + 1]].sum()),
                                      float(ma[graph.path_off[p]:graph.path_off[p + 1]].sum())], dtype=torch.float64)
                        for p in range(int(pb[rank]), int(pb[rank + 1]))]) if pb[rank + 1] > pb[rank] \
        else torch.zeros((0, 2), dtype=torch.float64)
    table = sfdist.all_gather_rows(rows, pb, None)
    counters = torch.tensor([orc.C[k] for k in oa.COUNTERS], dtype=torch.int64)
    sfdist.all_reduce_sum(counters, None)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), uniq=uniq.numpy(), table=table.numpy(), counters=counters.numpy(),
             ua=ua.numpy()[lo:hi], ma=ma.numpy()[lo:hi], lo=lo, hi=hi)
    dist.destroy_process_group()


def test_two_rank_protocol_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    graph, reads = synth.make_config("tiny")
    ref = oa.AbundanceOracle(graph.as_dict(), reads.as_dict()).run()
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    want_table = np.array([[ref.uniq_abund[graph.path_off[p]:graph.path_off[p + 1]].sum(),
                            ref.mult_abund[graph.path_off[p]:graph.path_off[p + 1]].sum()] for p in range(graph.n_paths)])
    for part in parts:
        assert np.array_equal(part["uniq"], np.asarray(ref.uniq_reads))                       # integers: exact
        assert np.array_equal(part["counters"], np.array([ref.C[k] for k in oa.COUNTERS]))
        assert np.allclose(part["table"], want_table, rtol=1e-9, atol=0)
        lo, hi = int(part["lo"]), int(part["hi"])
        assert np.allclose(part["ua"], ref.uniq_abund[lo:hi], rtol=1e-9, atol=0)
        assert np.allclose(part["ma"], ref.mult_abund[lo:hi], rtol=1e-9, atol=0)
    assert parts[0]["hi"] == parts[1]["lo"] and parts[1]["hi"] == graph.n_path_nodes


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partitions_cover_everything(world):
    graph, reads = synth.make_config("tiny")
    b = [sfdist.shard_bounds(reads.n_groups, world, r) for r in range(world)]
    assert b[0][0] == 0 and b[-1][1] == reads.n_groups and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
    pb = sfdist.path_partition(graph.path_off + np.arange(graph.n_paths + 1), world)
    assert pb[0] == 0 and pb[-1] == graph.n_paths and np.all(np.diff(pb) >= 0)
    # shards re-based correctly: concatenating them gives back the reads
    parts = [reads.slice_groups(*x) for x in b]
    assert sum(p.n_groups for p in parts) == reads.n_groups and sum(p.n_records for p in parts) == reads.n_records
    assert np.array_equal(np.concatenate([p.walk_node for p in parts]), reads.walk_node)
