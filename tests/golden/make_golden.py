"""Regenerates the golden fixtures under tests/golden/case_*/ by running the UNMODIFIED
reference (imported by file path from /root/reference) on seeded synthetic inputs.

    python tests/golden/make_golden.py

Run it in the build container (where /root/reference is mounted).  The inputs
(graph.gfa, reads.json, dict_clusters.pickle) and the reference's outputs
(expected.json: counters, per-path accumulators, histograms, unassigned-cluster
records, the CSV texts and the strain profile) are committed; the GPU box has
no reference and replays them.
"""
import json
import os
import pickle
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import casegen      # noqa: E402
import refharness   # noqa: E402

# (seed, groups, clusters, strains): chosen so that together they hit every leaf of the
# pass-1/pass-2 decision tree (asserted in tests/test_oracle_golden.py::test_golden_covers_all_leaves)
CASES = [(1, 80, 4, 4), (2, 80, 5, 4), (5, 120, 3, 2), (7, 100, 6, 5), (11, 150, 4, 3), (23, 200, 5, 6),
         (42, 60, 2, 2), (77, 120, 6, 3)]


def _jsonable(x):
    if isinstance(x, dict):
        return {str(k): _jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_jsonable(v) for v in x]
    if hasattr(x, "item"):
        return x.item()
    return x


def main():
    assert refharness.have_reference(), "needs /root/reference"
    for seed, n_groups, n_clu, n_str in CASES:
        out = os.path.join(HERE, f"case_{seed:03d}")
        shutil.rmtree(out, ignore_errors=True)
        gfa, js, pk = casegen.make_case(seed, out, n_groups=n_groups, n_clusters=n_clu, n_strains=n_str)
        tmp = tempfile.mkdtemp()
        ref = refharness.run_json2csv(gfa, js, pk, 0.95, outdir=tmp)
        prefix = ref.pop("prefix")
        texts = {}
        for suffix in ("gene_table", "mappingscore_freq", "identityscore_freq", "subst_freq", "indel_freq",
                       "errors_freq", "unassigned"):
            texts[suffix] = open(f"{prefix}_{suffix}.csv").read()
        try:
            prof = refharness.run_strains(prefix + "_gene_table.csv", os.path.join(tmp, "prof"))
            texts["strains_profile"] = open(prof).read()
        except Exception as e:     # ragged gene table (Q5) -> pandas ParserError
            texts["strains_profile_error"] = type(e).__name__
        # the CLI, for stdout + allclusters.pickle
        cli_dir = tempfile.mkdtemp()
        stdout = refharness.run_json2csv_cli(["-g", gfa, "-m", js, "-p", pk, "-o", os.path.join(cli_dir, "o")], cli_dir)
        with open(os.path.join(cli_dir, "allclusters.pickle"), "rb") as fh:
            allclusters = pickle.load(fh)
        texts["stdout_tail"] = stdout.split("\n")[-12:]
        ref["allclusters"] = allclusters
        ref["texts"] = texts
        ref["params"] = dict(seed=seed, n_groups=n_groups, n_clusters=n_clu, n_strains=n_str, thr=0.95)
        with open(os.path.join(out, "expected.json"), "w") as fh:
            json.dump(_jsonable(ref), fh)
        print(out, {k: v for k, v in ref["counters"].items() if v})


if __name__ == "__main__":
    main()
