"""Differential fuzzing on the GPU: the CUDA path (through the C ABI) against the oracle on random file-level cases
(tests/casegen.py: every leaf of the decision tree, ties, recombinant reads, single-end groups, DAG alignments) decoded
by the native decoder, in both accumulation modes.
usage: python tools/fuzz_gpu.py FIRST_SEED N_SEEDS [SECONDS]     (stops after SECONDS, prints what it covered)"""
import os
import sys
import tempfile
import traceback
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import casegen  # noqa: E402
from oracle import abundance as oa  # noqa: E402
from strainflair_b200 import decode, engine, gfa  # noqa: E402
from test_gpu_parity import check_against_oracle  # noqa: E402


def one(seed, tmp):
    rng = np.random.default_rng(seed)
    n_groups = int(rng.integers(20, 400))
    gf, js, pk = casegen.make_case(seed, tmp, n_groups=n_groups, n_clusters=int(rng.integers(1, 9)),
                                   n_strains=int(rng.integers(1, 7)))
    thr = float(rng.choice([0.0, 0.5, 0.95, 0.95, 1.0]))
    graph = gfa.load_graph(gf, pk)
    reads = decode.decode_json(js, graph, thr)
    orc = oa.AbundanceOracle(graph.as_dict(), reads.as_dict()).run()
    res = engine.AbundanceEngine(graph, "cuda:0", deterministic=bool(seed & 1)).run(reads)
    check_against_oracle(graph, reads, res, orc)
    return reads.n_alns


def main():
    import time
    first, n = int(sys.argv[1]), int(sys.argv[2])
    budget = float(sys.argv[3]) if len(sys.argv) > 3 else 1e18
    warnings.simplefilter("ignore")
    bad = alns = done = 0
    t0 = time.time()
    for seed in range(first, first + n):
        if time.time() - t0 > budget:
            break
        done += 1
        with tempfile.TemporaryDirectory() as tmp:
            try:
                alns += one(seed, tmp)
            except (KeyError, IndexError, TypeError) as e:
                # the reference's own error quirks (Q7/Q10) can be generated: the oracle must raise the same class
                print(f"seed {seed}: {type(e).__name__} (quirk path)", flush=True)
            except Exception as e:      # noqa: BLE001
                bad += 1
                print(f"SEED {seed} FAILED: {type(e).__name__}: {str(e)[:300]}", flush=True)
                traceback.print_exc(limit=4)
    print(f"gpu fuzz seeds {first}..{first + done - 1}: {done - bad} ok, {bad} failed, {alns} alignments", flush=True)


if __name__ == "__main__":
    main()
