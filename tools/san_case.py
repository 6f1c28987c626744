"""Small parity cases for compute-sanitizer runs (memcheck / racecheck / synccheck): the tile kernels and the
global-memory kernels on the `tiny` configuration, a high-multiplicity graph whose pairs overflow the tuple buffers
(hand-over to the generic kernels) and the reference's golden cases, each checked against the oracle.

    compute-sanitizer --tool memcheck  python tools/san_case.py
    compute-sanitizer --tool racecheck python tools/san_case.py
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np  # noqa: E402

from oracle import abundance as oa  # noqa: E402
from strainflair_b200 import engine, synth  # noqa: E402
import test_gpu_parity as tp  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
cases = []
g, r = synth.make_config("tiny", scale=scale)
cases.append(("tiny", g, r))
g = synth.make_graph(11, n_clusters=4, chain_len=40, hap_lo=24, hap_hi=24, n_strains=9, bubble_p=0.4, flip_p=0.03)
cases.append(("many-haplotypes", g, synth.make_reads(g, 12, n_pairs=int(1200 * scale), read_len=100, p_tie=0.3)))
for name, graph, reads in cases:
    orc = oa.AbundanceOracle(graph.as_dict(), reads.as_dict()).run()
    for mode in tp.MODES:
        res = engine.abundance_pass(graph, reads, **mode)
        tp.check_against_oracle(graph, reads, res, orc)
        print(name, mode, "ok", reads.n_groups, "groups", flush=True)
print("sanitizer cases passed")
