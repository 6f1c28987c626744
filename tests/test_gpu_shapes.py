"""GPU parity at the shapes of BASELINE.json's configurations: the CUDA path against the oracle on samples of the C3,
C4 and C5 workloads (the full graphs, as many read pairs as the Python oracle finishes in well under a minute)."""
import numpy as np
import pytest
import torch

from oracle import abundance as oa
from strainflair_b200 import engine, synth
from test_gpu_parity import check_against_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,pairs", [("c3", 60_000), ("c4", 12_000), ("c5", 30_000)])
def test_config_shapes_match_oracle(name, pairs):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    gkw, rkw = synth.CONFIGS[name]
    idx = list(synth.CONFIGS).index(name)
    graph = synth.make_graph(synth.SEED0 + idx, **gkw)
    reads = synth.make_reads(graph, synth.SEED0 + idx + 1000, n_pairs=pairs, read_len=rkw["read_len"])
    orc = oa.AbundanceOracle(graph.as_dict(), reads.as_dict()).run()
    eng = engine.AbundanceEngine(graph)                   # default: path sets from the step index, shared-memory bins, tile order
    assert eng.step_index and eng.tile_bins and not eng.dg.steps["steps_partial"]     # every tile of these graphs is simple
    assert eng.tiling is not None and eng.tiling.n_tiles >= 3000          # every cluster is a tile (or shares one)
    host = eng.stage(reads)
    assert host.grp_begin > 0.9 * reads.n_groups                          # ... and holds most read groups
    res = eng.run(host)
    check_against_oracle(graph, reads, res, orc)
    if name == "c4":
        assert res.C["ST_TUPLES"] / reads.n_alns > 8 and res.n_deferred > 0.7 * reads.n_groups
    # the other kernel families on the same sample: path sets with every sum at the L2, tuple lists from the occurrence
    # index, the all-in-one shared-memory tile kernels, and the global-memory kernels on reads in file order
    check_against_oracle(graph, reads, engine.AbundanceEngine(graph, tile_bins=False).run(host), orc)
    check_against_oracle(graph, reads, engine.AbundanceEngine(graph, step_index=False).run(host), orc)
    if name == "c5":          # (every engine rebuilds the device layout of the 74 M-slot graph on the host: two families more
        return                #  would double the two minutes this case takes; they run on C3 and C4)
    check_against_oracle(graph, reads, engine.AbundanceEngine(graph, tile_kernels=True).run(host), orc)
    check_against_oracle(graph, reads, engine.AbundanceEngine(graph, tiles=False).run(reads), orc)
