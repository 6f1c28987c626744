"""Strain-level profile from a gene table (the reference's ``compute_strains_abundance``,
compute_strains_abundance.py:48-72): CSV in, device aggregation, CSV out."""
import math

import numpy as np

N_FIXED = 11        # the last 11 columns of the gene table are not strains (compute_strains_abundance.py:52)
OUT_COLS = ("detected_genes", "mean_abund", "mean_abund_nz", "median_abund", "median_abund_nz")


def read_gene_table(path):
    """';'-separated gene table -> (strain names, counts float64[P, S], columns dict).  Empty / nan / None
    cells become 0 (pandas ``fillna(0)``, :50); a ragged file is an error, as it is for pandas (Q5)."""
    with open(path) as fh:
        header = fh.readline().rstrip("\n").split(";")
        rows = []
        for k, ln in enumerate(fh):
            if not ln.strip():
                continue
            cells = ln.rstrip("\n").split(";")
            if len(cells) != len(header):
                raise ValueError(f"Error tokenizing data. Expected {len(header)} fields in line {k + 2}, saw {len(cells)}")
            rows.append(cells)
    S = len(header) - N_FIXED
    if S < 0:
        raise ValueError("not a gene table: fewer than 11 columns")

    def num(x):
        return 0.0 if x in ("", "nan", "NaN", "None") else float(x)

    data = np.array([[num(x) for x in r] for r in rows], np.float64).reshape(len(rows), len(header))
    cols = {name: data[:, S + i] for i, name in enumerate(header[S:])}
    return header[:S], data[:, :S], cols


def profile_from_arrays(strain_names, counts, cols, thr=0.5, device="cuda:0"):
    from . import engine
    counts = np.asarray(counts, np.float64)
    P, S = counts.shape
    pos = counts > 0
    rows, strains = np.nonzero(pos)
    row_off = np.zeros(P + 1, np.int64)
    np.cumsum(pos.sum(axis=1), out=row_off[1:])
    out = engine.strain_aggregate(row_off, strains, counts[rows, strains], cols["mean_abund_multiple"],
                                  cols["mean_abund_multiple_nz"], cols["ratio_covered_nodes"], S, thr, device)
    return out


def profile_text(strain_names, out, thr):
    """pandas ``to_csv`` of the profile: NaN -> empty cell; a column whose values are all the int 0 the
    reference assigned (no strain above the threshold) prints ``0``, otherwise floats (:57-72)."""
    det = out[:, 0]
    any_float = bool(np.any(det > thr))          # some strain got a float mean/median -> float64 column

    def cell(v, integer):
        if isinstance(v, float) and math.isnan(v):
            return ""
        return "0" if integer else repr(float(v))

    lines = ["," + ",".join(OUT_COLS) + "\n"]
    for s, name in enumerate(strain_names):
        cells = [cell(float(det[s]), False)] + [cell(float(out[s, k]), not any_float) for k in range(1, 5)]
        lines.append(f"{name}," + ",".join(cells) + "\n")
    return "".join(lines)


def compute_strains_abundance(input_csv, out_stem, thr=0.5, device="cuda:0"):
    names, counts, cols = read_gene_table(input_csv)
    out = profile_from_arrays(names, counts, cols, thr, device)
    with open(f"{out_stem}.csv", "w") as fh:
        fh.write(profile_text(names, out, thr))
    return names, out
