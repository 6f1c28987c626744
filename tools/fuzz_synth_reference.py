"""Differential fuzzing on the CPU, bench-generator family: random bubble-chain pangenomes (strainflair_b200/synth.py:
1-24 haplotypes per cluster, 2-150 strains, ties, recombinant and single-end reads) exported to GFA/JSON/pickle; the
unmodified reference on the files against the oracle on the files and on the generator's arrays.
usage: python tools/fuzz_synth_reference.py FIRST_SEED N_SEEDS"""
import os
import sys
import tempfile
import traceback
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import compare  # noqa: E402
import refharness  # noqa: E402
from oracle import abundance as oa  # noqa: E402
from oracle import decode as od  # noqa: E402
from oracle import graph as og  # noqa: E402
from strainflair_b200 import synth  # noqa: E402


def one(seed, tmp):
    rng = np.random.default_rng(seed)
    hap_lo = int(rng.integers(1, 12))
    g = synth.make_graph(seed, n_clusters=int(rng.integers(2, 30)), chain_len=int(rng.integers(6, 40)), hap_lo=hap_lo,
                         hap_hi=hap_lo + int(rng.integers(0, 13)), n_strains=int(rng.choice([2, 5, 16, 64, 150])),
                         bubble_p=float(rng.uniform(0.05, 0.6)), flip_p=float(rng.uniform(0.0, 0.6)),
                         max_node_len=int(rng.choice([7, 31, 63])))
    # reads no longer than the shortest path (a tied alignment on a shorter sibling path would not be a tie) and not
    # shorter than 40 bases (the generator's one-substitution reads must stay above the 0.95 threshold)
    if int(g.path_seq_len.min()) < 40:
        return {"SKIPPED_SHORT_PATHS": 1}
    r = synth.make_reads(g, seed + 1, n_pairs=int(rng.integers(50, 700)), read_len=min(int(rng.choice([40, 100, 150])), int(g.path_seq_len.min())),
                         p_same_path=float(rng.uniform(0.3, 1.0)), p_tie=float(rng.uniform(0.0, 0.4)),
                         p_recomb=float(rng.uniform(0.0, 0.1)), p_single=float(rng.uniform(0.0, 0.3)))
    gf, js, pk = synth.export_files(g, r, tmp)
    go = og.load_graph(gf, pk)
    try:
        ref = refharness.run_json2csv(gf, js, pk, 0.95, outdir=tmp)
    except KeyError:
        # Q7: a ratio outside [0, 1] reached a histogram (the exporter does not keep every read shape inside);
        # the oracle must fail the same way on the same files
        try:
            oa.AbundanceOracle(go, od.decode_json(js, go, 0.95)).run()
        except KeyError:
            return {"Q7_KEYERROR_BOTH": 1}
        raise AssertionError("reference raised KeyError, the oracle did not")
    orc = oa.AbundanceOracle(go, od.decode_json(js, go, 0.95)).run()
    compare.check_state_vs_reference(orc, ref, exact=True)
    orc2 = oa.AbundanceOracle(g.as_dict(), r.as_dict()).run()          # straight from the arrays
    assert orc2.C == orc.C and orc2.uniq_reads == orc.uniq_reads
    assert np.array_equal(orc2.uniq_abund, orc.uniq_abund) and np.array_equal(orc2.mult_abund, orc.mult_abund)
    return orc.C


def main():
    first, n = int(sys.argv[1]), int(sys.argv[2])
    warnings.simplefilter("ignore")
    bad = 0
    tot = {}
    for seed in range(first, first + n):
        with tempfile.TemporaryDirectory() as tmp:
            try:
                for k, v in one(seed, tmp).items():
                    tot[k] = tot.get(k, 0) + v
            except Exception as e:      # noqa: BLE001
                bad += 1
                print(f"SEED {seed} FAILED: {type(e).__name__}: {str(e)[:300]}", flush=True)
                traceback.print_exc(limit=3)
    print(f"seeds {first}..{first + n - 1}: {n - bad} ok, {bad} failed; counters summed: "
          + ", ".join(f"{k}={v}" for k, v in tot.items() if v), flush=True)


if __name__ == "__main__":
    main()
