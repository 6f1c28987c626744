"""Differential fuzzing on the CPU: the unmodified reference (imported from /root/reference), the oracle and the
native host code (GFA loader + JSON decoder) on random cases of varied shape.
usage: python tools/fuzz_reference.py FIRST_SEED N_SEEDS     (prints one line per failing seed, a summary at the end)"""
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
import compare  # noqa: E402
import refharness  # noqa: E402
from oracle import abundance as oa  # noqa: E402
from oracle import decode as od  # noqa: E402
from oracle import graph as og  # noqa: E402
from oracle import report as orp  # noqa: E402
from strainflair_b200 import decode, gfa  # noqa: E402
from strainflair_b200.schema import GRAPH_ARRAYS, READS_ARRAYS  # noqa: E402


def one(seed, tmp):
    rng = np.random.default_rng(seed)
    n_groups = int(rng.integers(20, 260))
    n_clusters = int(rng.integers(1, 9))
    n_strains = int(rng.integers(1, 7))
    thr = float(rng.choice([0.0, 0.5, 0.9, 0.95, 0.95, 0.95, 1.0]))
    gf, js, pk = casegen.make_case(seed, tmp, n_groups=n_groups, n_clusters=n_clusters, n_strains=n_strains)
    ref = refharness.run_json2csv(gf, js, pk, thr, outdir=tmp)
    g = og.load_graph(gf, pk)
    r = od.decode_json(js, g, thr)
    orc = oa.AbundanceOracle(g, r).run()
    compare.check_state_vs_reference(orc, ref, exact=True)
    order = [g["strain_names"].index(s) for s in ref["species"]]
    assert orp.gene_table_text(orc, order) == open(ref["prefix"] + "_gene_table.csv").read(), "gene table text"
    for k, hn in enumerate(orp.HIST_NAMES):
        assert orp.hist_text(orc, k) == open(ref["prefix"] + "_" + hn + ".csv").read(), hn
    cl = orp.cluster_unassigned(orc)
    assert orp.unassigned_text(orc, cl, order) == open(ref["prefix"] + "_unassigned.csv").read(), "unassigned text"
    # native loaders against the oracle's
    ng = gfa.load_graph(gf, pk)
    for k in GRAPH_ARRAYS:
        assert np.array_equal(getattr(ng, k), np.asarray(g[k])), ("graph", k)
    nr = decode.decode_json(js, ng, thr)
    for k in READS_ARRAYS:
        a, b = getattr(nr, k), np.asarray(r[k])
        assert np.array_equal(a.reshape(-1), b.reshape(-1)), ("reads", k)


def main():
    first, n = int(sys.argv[1]), int(sys.argv[2])
    warnings.simplefilter("ignore")
    bad = 0
    for seed in range(first, first + n):
        with tempfile.TemporaryDirectory() as tmp:
            try:
                one(seed, tmp)
            except Exception as e:      # noqa: BLE001
                bad += 1
                print(f"SEED {seed} FAILED: {type(e).__name__}: {str(e)[:300]}", flush=True)
                traceback.print_exc(limit=3)
    print(f"seeds {first}..{first + n - 1}: {n - bad} ok, {bad} failed", flush=True)


if __name__ == "__main__":
    main()
