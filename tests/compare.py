"""Comparison helpers shared by the parity tests: oracle state vs. a reference result
dict (tests/refharness.py or a golden fixture), and CSV-by-column-name comparison."""
import math

import numpy as np

REL = 1e-9   # BASELINE.json north_star: fp64 abundances within 1e-9 relative


def close(a, b, rel=REL):
    a, b = float(a), float(b)
    if math.isnan(a) or math.isnan(b):
        return math.isnan(a) and math.isnan(b)
    return abs(a - b) <= rel * max(abs(a), abs(b))


def check_state_vs_reference(orc, ref, exact=True):
    """``orc``: anything exposing the oracle's accumulator attributes (uniq_reads, uniq_norm,
    mult_reads, mult_norm, uniq_abund, mult_abund, hist, C, g).  ``exact``: floats bit-equal
    (oracle, which sums in the reference's order) or within REL (CUDA path)."""
    g = orc.g
    P = len(g["path_off"]) - 1
    assert P == len(ref["paths"]), (P, len(ref["paths"]))
    for k, v in ref["counters"].items():
        assert int(orc.C[k]) == int(v), (k, orc.C[k], v)
    names = g["strain_names"]
    eq = (lambda a, b: float(a) == float(b)) if exact else close
    for p, rp in enumerate(ref["paths"]):
        lo, hi = int(g["path_off"][p]), int(g["path_off"][p + 1])
        assert [int(g["node_ids"][k]) for k in g["path_nodes"][lo:hi]] == [int(x) for x in rp["node_ids"]], p
        o = g["pstr_off"]
        mine = {names[int(s)]: int(c) for s, c in zip(g["pstr_id"][o[p]:o[p + 1]], g["pstr_cnt"][o[p]:o[p + 1]])}
        assert mine == {k: int(v) for k, v in rp["strain_ids"].items()}, (p, mine, rp["strain_ids"])
        assert list(mine) == list(rp["strain_ids"]), "strain insertion order"
        c = int(g["path_cluster"][p])
        assert (None if c < 0 else c) == rp["cluster_id"], p
        assert int(g["path_seq_len"][p]) == int(rp["seq_len"])
        assert int(orc.uniq_reads[p]) == int(rp["uniq_reads"]), (p, orc.uniq_reads[p], rp["uniq_reads"])
        for mine_v, ref_v, what in ((orc.uniq_norm[p], rp["uniq_norm"], "uniq_norm"),
                                    (orc.mult_reads[p], rp["mult_reads"], "mult_reads"),
                                    (orc.mult_norm[p], rp["mult_norm"], "mult_norm")):
            assert eq(mine_v, ref_v), (p, what, mine_v, ref_v)
        for mine_a, ref_a, what in ((orc.uniq_abund[lo:hi], rp["uniq_abund"], "uniq_abund"),
                                    (orc.mult_abund[lo:hi], rp["mult_abund"], "mult_abund")):
            assert len(mine_a) == len(ref_a)
            for i, (x, y) in enumerate(zip(mine_a, ref_a)):
                assert eq(x, y), (p, what, i, x, y)
    # histograms: reference keys = strains that created a path
    hist_names = ("mappingscore_freq", "identityscore_freq", "subst_freq", "indel_freq", "errors_freq")
    for k, hn in enumerate(hist_names):
        assert [names[s] for s in g["header_strains"]] == list(ref["hist"][hn].keys()), "histogram row order"
        for s in g["header_strains"]:
            assert [int(v) for v in orc.hist[k, s]] == [int(v) for v in ref["hist"][hn][names[s]]], (hn, names[s])


def parse_csv(text, sep=";"):
    lines = [ln for ln in text.split("\n") if ln != ""]
    header = lines[0].split(sep)
    rows = [ln.split(sep) for ln in lines[1:]]
    return header, rows


def check_csv_by_column(mine_text, ref_text, sep=";", exact=True, key_col=None):
    """Same columns (any order, Q6), same rows, cells equal as text or (floats) within REL."""
    mh, mr = parse_csv(mine_text, sep)
    rh, rr = parse_csv(ref_text, sep)
    assert sorted(mh) == sorted(rh), (mh, rh)
    assert len(mr) == len(rr), (len(mr), len(rr))
    idx = [mh.index(c) for c in rh]
    if key_col is not None:          # rows may come in another order too (strain rows follow the column order, Q6)
        k_m, k_r = mh.index(key_col), rh.index(key_col)
        assert sorted(r[k_m] for r in mr) == sorted(r[k_r] for r in rr)
        by_key = {r[k_m]: r for r in mr}
        mr = [by_key[r[k_r]] for r in rr]
    for i, (a, b) in enumerate(zip(mr, rr)):
        assert len(a) == len(b) == len(rh), ("ragged row", i)
        for j, c in enumerate(rh):
            x, y = a[idx[j]], b[j]
            if x == y:
                continue
            assert not exact, (i, c, x, y)
            fx, fy = (0.0 if v == "" else float(v) for v in (x, y))
            assert close(fx, fy), (i, c, x, y)
