"""Differential fuzzing on the GPU: the CUDA path (through the C ABI) against the oracle on random file-level cases
(tests/casegen.py: every leaf of the decision tree, ties, recombinant reads, single-end groups, DAG alignments) decoded
by the native decoder, in both accumulation modes.
usage: python tools/fuzz_gpu.py FIRST_SEED N_SEEDS [SECONDS [synth]]     (stops after SECONDS, prints what it covered;
       "synth": random pangenomes of the bench generator's family instead -- 1-24 haplotypes per cluster, 2-150 strains
       (multi-word strain masks), long match lists, the generic group kernels -- straight from the arrays)"""
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


def one_synth(seed, tmp):
    from strainflair_b200 import synth
    rng = np.random.default_rng(seed)
    hap_lo = int(rng.integers(1, 12))
    graph = synth.make_graph(seed, n_clusters=int(rng.integers(2, 30)), chain_len=int(rng.integers(6, 40)), hap_lo=hap_lo,
                             hap_hi=hap_lo + int(rng.integers(0, 13)), n_strains=int(rng.choice([2, 5, 16, 64, 150])),
                             bubble_p=float(rng.uniform(0.05, 0.6)), flip_p=float(rng.uniform(0.0, 0.6)),
                             max_node_len=int(rng.choice([7, 31, 63])))
    if int(graph.path_seq_len.min()) < 40:          # see tools/fuzz_synth_reference.py
        return 0
    reads = synth.make_reads(graph, seed + 1, n_pairs=int(rng.integers(50, 1500)), read_len=min(int(rng.choice([40, 100, 150])), int(graph.path_seq_len.min())),
                             p_same_path=float(rng.uniform(0.3, 1.0)), p_tie=float(rng.uniform(0.0, 0.4)),
                             p_recomb=float(rng.uniform(0.0, 0.1)), p_single=float(rng.uniform(0.0, 0.3)),
                             p_exc=float(rng.choice([0.0, 0.02])))
    orc = oa.AbundanceOracle(graph.as_dict(), reads.as_dict()).run()
    eng = engine.AbundanceEngine(graph, "cuda:0", deterministic=bool(seed & 1))
    res = eng.run(engine.HostReads(reads, wire=bool(seed & 2)))
    check_against_oracle(graph, reads, res, orc)
    return reads.n_alns


def main():
    import time
    first, n = int(sys.argv[1]), int(sys.argv[2])
    budget = float(sys.argv[3]) if len(sys.argv) > 3 else 1e18
    fn = one_synth if len(sys.argv) > 4 and sys.argv[4] == "synth" else one
    warnings.simplefilter("ignore")
    bad = alns = done = 0
    t0 = time.time()
    for seed in range(first, first + n):
        if time.time() - t0 > budget:
            break
        done += 1
        with tempfile.TemporaryDirectory() as tmp:
            try:
                alns += fn(seed, tmp)
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
