"""Differential fuzzing of the native unassigned-read analysis (csrc/unassigned.cpp, sfb_min_strains) against
oracle/report.py, which runs the real set / scipy.signal.find_peaks / networkx clique code of the reference's
Cluster.min_strains_inference (json2csv.py:149-201).
usage: python tools/fuzz_min_strains.py FIRST_SEED N_SEEDS [CLUSTERS_PER_SEED]"""
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from oracle import report as orp  # noqa: E402
from strainflair_b200 import report  # noqa: E402
from test_host import _random_walk_sets  # noqa: E402


def main():
    first, n = int(sys.argv[1]), int(sys.argv[2])
    per = int(sys.argv[3]) if len(sys.argv) > 3 else 400
    warnings.simplefilter("ignore")
    bad = total = kept = 0
    for seed in range(first, first + n):
        sets = _random_walk_sets(np.random.default_rng(seed), per)
        got = report.min_strains(sets, threads=2)
        for k, walks in enumerate(sets):
            if not walks:
                continue
            want = orp.min_strains([list(w) for w in walks])
            total += 1
            kept += want[1] > 0
            ok = (int(got[0][k]), int(got[1][k]), int(got[2][k])) == tuple(int(v) for v in want[:3]) and \
                ((np.isnan(got[3][k]) and np.isnan(want[3])) or got[3][k] == want[3])
            if not ok:
                bad += 1
                print(f"SEED {seed} cluster {k} FAILED: native {[g[k] for g in got]} reference code {want} walks {walks}", flush=True)
    print(f"seeds {first}..{first + n - 1}: {total} clusters ({kept} with reads kept by the support filter), {bad} differences", flush=True)


if __name__ == "__main__":
    main()
