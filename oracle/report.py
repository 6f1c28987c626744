"""Oracle (test infrastructure): outputs of the abundance path.

Restates ``Pangenome.print_gene_table`` (reference json2csv.py:399-474),
``print_error_distribution`` (:515-526), ``print_unassigned_info`` (:528-549),
the unassigned-read bookkeeping (:884-891) with
``Cluster.min_strains_inference`` (:149-201) and
``compute_strains_abundance_main`` (compute_strains_abundance.py:48-72).
networkx / scipy are called exactly where the reference calls them.
"""
import warnings

import numpy as np

GENE_COLS = ("cluster", "seq_len", "nb_uniq_mapped", "nb_uniq_mapped_normalized", "nb_multimapped",
             "nb_multimapped_normalized", "mean_abund_uniq", "mean_abund_uniq_nz", "mean_abund_multiple",
             "mean_abund_multiple_nz", "ratio_covered_nodes")
HIST_NAMES = ("mappingscore_freq", "identityscore_freq", "subst_freq", "indel_freq", "errors_freq")


def _mean(x):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return np.mean(x)


def gene_rows(orc):
    """Per-path values of the gene table (json2csv.py:434-469), as Python/numpy scalars."""
    g = orc.g
    rows = []
    for p in range(orc.P):
        u = orc.uniq_abund[g["path_off"][p]:g["path_off"][p + 1]]
        m = orc.mult_abund[g["path_off"][p]:g["path_off"][p + 1]]
        um = u + m
        c = int(g["path_cluster"][p])
        rows.append(dict(
            cluster=None if c < 0 else c,
            seq_len=int(g["path_seq_len"][p]),
            nb_uniq_mapped=orc.uniq_reads[p],
            nb_uniq_mapped_normalized=orc.uniq_norm[p],
            nb_multimapped=orc.uniq_reads[p] + orc.mult_reads[p],
            nb_multimapped_normalized=orc.uniq_norm[p] + orc.mult_norm[p],
            mean_abund_uniq=_mean(u),
            mean_abund_uniq_nz=_mean(u[u > 0]),
            mean_abund_multiple=_mean(um),
            mean_abund_multiple_nz=_mean(um[um > 0]),
            ratio_covered_nodes=int(np.count_nonzero((u > 0) | (m > 0))) / float(len(u)),
        ))
    return rows


def gene_table_text(orc, strain_order=None):
    """The gene table as the reference writes it.  ``strain_order`` is the list
    of strain indices to use as header (default: first-seen order); the
    reference's own order is a set iteration order (Q6)."""
    g = orc.g
    names = g["strain_names"]
    keys = list(g["header_strains"] if strain_order is None else strain_order)
    out = ["".join(f"{names[s]};" for s in keys) + ";".join(GENE_COLS) + "\n"]
    for p, row in enumerate(gene_rows(orc)):
        o = g["pstr_off"]
        cnt = {s: 0 for s in keys}
        for s, c in zip(g["pstr_id"][o[p]:o[p + 1]], g["pstr_cnt"][o[p]:o[p + 1]]):
            if int(s) not in cnt:
                keys.append(int(s))          # ragged row, Q5 (:413-418)
            cnt[int(s)] = int(c)
        out.append("".join(f"{cnt[s]};" for s in keys) + ";".join(f"{row[k]}" for k in GENE_COLS) + "\n")
    return "".join(out)


def hist_text(orc, which):
    g = orc.g
    out = ["Strain;" + ";".join(str(i / 100) for i in range(101)) + "\n"]
    for s in g["header_strains"]:
        out.append(f"{g['strain_names'][s]}" + "".join(f";{int(v)}" for v in orc.hist[which, s]) + "\n")
    return "".join(out)


# ---- unassigned reads ---------------------------------------------------
def orient_small(walk):
    """canonical_tiny (json2csv.py:63-71)."""
    if len(walk) == 1:
        return walk
    ups = sum(b - a > 0 for a, b in zip(walk, walk[1:]))
    return walk if ups / (len(walk) - 1) > 0.5 else walk[::-1]


def compat(c1, c2):
    """check_incomp_comp (json2csv.py:83-97): 1 compatible, -1 incompatible, 0 disjoint."""
    shared = list(set(c1) & set(c2))
    if not shared:
        return 0
    pivot = shared[0]
    for i1 in [i for i, v in enumerate(c1) if v == pivot]:
        for i2 in [i for i, v in enumerate(c2) if v == pivot]:
            left = min(i1, i2)
            right = min(len(c1) - i1, len(c2) - i2)
            if c1[i1 - left:i1 + right] == c2[i2 - left:i2 + right]:
                return 1
    return -1


def nested(a, b):
    """sublist (json2csv.py:99-100)."""
    return set(a) <= set(b) or set(b) <= set(a)


def min_strains(walks):
    """Cluster.min_strains_inference (json2csv.py:149-201).
    Returns (before, n_after_filter, after, mean_support)."""
    import networkx as nx
    import scipy.signal

    walks = [orient_small(w) for w in walks]
    n = len(walks)
    net = nx.Graph()
    clash = {k: [] for k in range(n)}
    agree = {k: [] for k in range(n)}
    for i in range(n):
        net.add_node(i)
        for j in range(i + 1, n):
            verdict = compat(walks[i], walks[j])
            if verdict == -1:
                clash[i].append(j)
                clash[j].append(i)
                net.add_edge(i, j)
            elif verdict == 1:
                agree[i].append(j)
                agree[j].append(i)
    sizes = [len(c) for c in nx.clique.find_cliques(net)]
    before = max(sizes) if sizes else 0
    kept = []
    for i in range(n):
        strict = [r for r in agree[i] if not nested(walks[i], walks[r])]
        ok = False
        if len(strict) >= 2:
            support = [1] * len(walks[i])
            for r in agree[i]:
                for node in walks[r]:
                    if node in walks[i]:
                        support[walks[i].index(node)] += 1
            ok = len(scipy.signal.find_peaks(-np.asarray(support))[0]) == 0
        if ok:
            kept.append(walks[i])
        else:
            net.remove_node(i)
    sizes = [len(c) for c in nx.clique.find_cliques(net)]
    after = max(sizes) if sizes else 0
    tally = {}
    for w in kept:
        for node in w:
            tally[node] = tally.get(node, 0) + 1
    return before, len(kept), after, _mean(list(tally.values()))


def cluster_unassigned(orc):
    """Replay the bookkeeping events (json2csv.py:884-891, 903-910, 1037-1044)
    into per-cluster records, in read order, with original GFA node ids."""
    g, r = orc.g, orc.r
    n_clu = len(g["clu_off"]) - 1
    clusters = [dict(reads_name=[], unassigned_paths=[], unassigned_abund={}, count=0, count_norm=0)
                for _ in range(n_clu)]
    for grp, m, cid in orc.book:
        a = int(r["grp_aln_off"][2 * grp + m])
        walk = [int(g["node_ids"][k]) for k in orc._walk(a)]
        cov = orc._cov(a)
        c = clusters[cid]
        c["reads_name"].append(r["mate_names"][2 * grp + m] if "mate_names" in r else "")
        c["unassigned_paths"].append(walk)
        c["count"] += 1
        c["count_norm"] += 1 / int(r["aln_read_len"][a])
        for node, v in zip(walk, cov):
            c["unassigned_abund"][node] = c["unassigned_abund"].get(node, 0) + v
    return clusters


def unassigned_text(orc, clusters=None, strain_order=None):
    g = orc.g
    names = g["strain_names"]
    keys = list(g["header_strains"] if strain_order is None else strain_order)
    clusters = cluster_unassigned(orc) if clusters is None else clusters
    out = ["Clusters;" + "".join(f"{names[s]};" for s in keys)
           + "nb_reads_unfilt;min_strains_unfilt;nb_reads_afterfilt;min_strains;mean_abund\n"]
    for cid, c in enumerate(clusters):
        before = n_after = after = support = 0
        if c["count"] != 0:
            before, n_after, after, support = min_strains(c["unassigned_paths"])
        cells = []
        for s in keys:
            tot = 0
            for p in g["clu_paths"][g["clu_off"][cid]:g["clu_off"][cid + 1]]:
                o = g["pstr_off"]
                for sid, cnt in zip(g["pstr_id"][o[p]:o[p + 1]], g["pstr_cnt"][o[p]:o[p + 1]]):
                    if int(sid) == s:
                        tot += int(cnt)
            cells.append(f"{tot};")
        out.append(f"Cluster_{cid};" + "".join(cells)
                   + f"{len(c['unassigned_paths'])};{before};{n_after};{after};{support}\n")
    return "".join(out)


def clusters_pickle_obj(clusters):
    """Content of allclusters.pickle (json2csv.py:1338-1344)."""
    return {f"cluster_{i}": dict(reads_name=c["reads_name"], unassigned_paths=c["unassigned_paths"],
                                 unassigned_abund=dict(c["unassigned_abund"]))
            for i, c in enumerate(clusters)}


def summary_lines(C, thr, prefix):
    """The stdout summary (json2csv.py:1346-1356)."""
    nb = C["NB_READS"]
    unused = C["FILTERED"] + C["UNASSIGNED"] + C["INCONSIST"] + C["INCONSIST_M"] + C["MULTI_ALIGN"]
    used = C["ASSIGNED_U"] + C["ASSIGNED_M"]
    return [
        f"Done, csv results are in {prefix}_gene_table.csv, and error distribution are in {prefix}_X_freq.csv",
        f"{nb} reads processed ({C['PAIRS']} pairs, {nb - C['PAIRS'] * 2} singles).",
        f"{C['FILTERED']} reads did not met the user-defined mapping score threshold of {thr}.",
        f"{C['UNASSIGNED']} reads could not be assigned to any colored path / gene.",
        f"{C['INCONSIST'] + C['INCONSIST_M']} reads have been dropped for inconsistancies of strain assignation "
        f"between two reads of the same pair (paired-end only).",
        f"{C['AMBIG_CP_RESOLVED']} colored path and {C['AMBIG_RESOLVED']} strain assignation ambiguities resolved, "
        f"and {C['AMBIG_REDUCED']} strain assignation ambiguities reduced (paired-end only).",
        f"{C['MULTI_ALIGN']} reads shown multiple alignments that could not be resolved.",
        f"In total, {unused} reads were not used.",
        f"{C['ASSIGNED_U']} reads have been assigned during the first step (unique mapping).",
        f"{C['ASSIGNED_M']} reads have been assigned during the second step (multiple mapping).",
        f"In total, {used} ({used / nb * 100}%) reads were used.",
    ]


# ---- strain level ---------------------------------------------------------
def strain_profile(strain_cols, counts, table, thr=0.5):
    """compute_strains_abundance (compute_strains_abundance.py:48-72) on arrays.

    strain_cols  list of column names
    counts       float64[P, S]  strain copy counts per gene row (NaN already -> 0)
    table        dict of float64[P] for 'mean_abund_multiple', 'mean_abund_multiple_nz',
                 'ratio_covered_nodes' (NaN already -> 0)
    Returns dict column -> list (Python scalars; ints where pandas keeps ints)."""
    counts = np.asarray(counts, np.float64)
    keep = (counts > 0).sum(axis=1) == 1
    counts = counts[keep]
    mam = np.asarray(table["mean_abund_multiple"], np.float64)[keep]
    mam_nz = np.asarray(table["mean_abund_multiple_nz"], np.float64)[keep]
    cov = np.asarray(table["ratio_covered_nodes"], np.float64)[keep]
    S = len(strain_cols)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        detected = [float(np.float64(counts[cov > 0, s].sum()) / np.float64(counts[:, s].sum())) for s in range(S)]
        out = {"detected_genes": detected}
        for name, src, fn in (("mean_abund", mam, np.mean), ("mean_abund_nz", mam_nz, np.mean),
                              ("median_abund", mam, np.median), ("median_abund_nz", mam_nz, np.median)):
            col = []
            for s in range(S):
                if detected[s] > thr:
                    sel = counts[:, s] > 0
                    col.append(float(fn(src[sel] / counts[sel, s])))
                else:
                    col.append(0)
            if any(isinstance(v, float) for v in col):       # pandas: one float makes the column float64
                col = [float(v) for v in col]
            tot = np.nansum(np.asarray(col, np.float64)) if col else 0
            if tot != 0:
                col = [float(v) for v in (np.asarray(col, np.float64) / (tot / 100))]
            out[name] = col
    return out


def read_gene_table(path):
    """Minimal ';' CSV reader with pandas' NaN -> 0 rule (compute_strains_abundance.py:48-50)."""
    with open(path) as fh:
        header = fh.readline().rstrip("\n").split(";")
        rows = [ln.rstrip("\n").split(";") for ln in fh if ln.strip()]
    S = len(header) - 11
    def num(x):
        if x in ("", "nan", "NaN", "None"):
            return 0.0
        return float(x)
    for i, r in enumerate(rows):
        if len(r) != len(header):        # pandas: ParserError "Expected N fields in line i, saw M" (Q5)
            raise ValueError(f"Error tokenizing data. Expected {len(header)} fields in line {i + 2}, saw {len(r)}")
    data = np.array([[num(x) for x in r] for r in rows], np.float64).reshape(len(rows), len(header))
    table = {name: data[:, S + i] for i, name in enumerate(header[S:])}
    return header[:S], data[:, :S], table


def strain_profile_text(strain_cols, prof):
    cols = ("detected_genes", "mean_abund", "mean_abund_nz", "median_abund", "median_abund_nz")
    def cell(v):
        if isinstance(v, float) and np.isnan(v):
            return ""
        return repr(v) if isinstance(v, float) else str(v)
    out = ["," + ",".join(cols) + "\n"]
    for i, s in enumerate(strain_cols):
        out.append(f"{s}," + ",".join(cell(prof[c][i]) for c in cols) + "\n")
    return "".join(out)
