"""Output files of ``json2csv`` from the engine's results (host side).

File formats and cell formatting follow the reference byte for byte where it is deterministic:
gene table (json2csv.py:399-474), five error-distribution tables (:515-526), per-cluster unassigned report
(:528-549, with ``Cluster.min_strains_inference`` :149-201), ``allclusters.pickle`` (:1338-1344) and the
stdout summary (:1346-1356).  The only free choice is the order of the strain columns, which in the
reference is a ``set`` iteration order: here it is first-seen order.
"""
import pickle

import numpy as np

from . import _lib
from .decode import mate_name

GENE_COLS = ("cluster", "seq_len", "nb_uniq_mapped", "nb_uniq_mapped_normalized", "nb_multimapped",
             "nb_multimapped_normalized", "mean_abund_uniq", "mean_abund_uniq_nz", "mean_abund_multiple",
             "mean_abund_multiple_nz", "ratio_covered_nodes")
HIST_FILES = ("mappingscore_freq", "identityscore_freq", "subst_freq", "indel_freq", "errors_freq")


def _f(x):
    """str() of a Python/numpy float as the reference's f-strings print it."""
    return repr(float(x))              # repr(nan) == "nan"


def gene_table_lines(graph, res):
    """Yields the gene table line by line.  Typing quirks kept: counters never touched print as the int 0
    (json2csv.py:114-118), touched ones as floats."""
    names = graph.strain_names
    keys = list(graph.header_strains)
    yield "".join(f"{names[s]};" for s in keys) + ";".join(GENE_COLS) + "\n"
    t = res.table
    po, pid, pcnt = graph.pstr_off, graph.pstr_id, graph.pstr_cnt
    P = graph.n_paths
    col = np.full(max(len(graph.strain_names), 1), -1, np.int64)
    col[np.asarray(keys, np.int64)] = np.arange(len(keys))
    if P and len(keys) and (len(pid) == 0 or (col[pid].min() >= 0 and pcnt.max() < 10 and pcnt.min() >= 0)):
        # the usual case (every strain in the header, single-digit copy counts): strain cells of all rows at once
        cells = np.full((P, 2 * len(keys)), ord(";"), np.uint8)
        cells[:, 0::2] = ord("0")
        row = np.repeat(np.arange(P), np.diff(po))
        cells[row, 2 * col[pid]] = ord("0") + pcnt.astype(np.uint8)
        strain_cells = cells.tobytes().decode("ascii")
        width = 2 * len(keys)
        uniq_all, multi_all = res.uniq_reads.astype(np.int64).tolist(), (np.asarray(res.mult_hits) > 0).tolist()
        cluster_all, seq_all, rows = graph.path_cluster.tolist(), graph.path_seq_len.tolist(), t.tolist()
        for p in range(P):
            uniq, multi_touched, c, v = uniq_all[p], multi_all[p], cluster_all[p], rows[p]
            yield (strain_cells[p * width:(p + 1) * width]
                   + f"{'None' if c < 0 else c};{seq_all[p]};{uniq};{_f(v[0]) if uniq else '0'};"
                     f"{_f(v[1]) if multi_touched else uniq};{_f(v[2]) if (uniq or multi_touched) else '0'};"
                     f"{_f(v[3])};{_f(v[4])};{_f(v[5])};{_f(v[6])};{_f(v[7])}\n")
        return
    for p in range(P):
        cnt = dict.fromkeys(keys, 0)
        for e in range(int(po[p]), int(po[p + 1])):
            s = int(pid[e])
            if s not in cnt:
                keys.append(s)              # strain seen only on duplicate P lines: ragged row (Q5, :413-418)
            cnt[s] = int(pcnt[e])
        uniq = int(res.uniq_reads[p])
        multi_touched = int(res.mult_hits[p]) > 0
        c = int(graph.path_cluster[p])
        cells = [
            "None" if c < 0 else str(c),
            str(int(graph.path_seq_len[p])),
            str(uniq),
            _f(t[p, 0]) if uniq else "0",
            _f(t[p, 1]) if multi_touched else str(uniq),
            _f(t[p, 2]) if (uniq or multi_touched) else "0",
            _f(t[p, 3]), _f(t[p, 4]), _f(t[p, 5]), _f(t[p, 6]), _f(t[p, 7]),
        ]
        yield "".join(f"{cnt[s]};" for s in keys) + ";".join(cells) + "\n"


def write_gene_table(path, graph, res):
    with open(path, "w") as fh:
        fh.writelines(gene_table_lines(graph, res))


def write_histogram(path, graph, res, which):
    with open(path, "w") as fh:
        fh.write("Strain;" + ";".join(str(i / 100) for i in range(101)) + "\n")
        for s in graph.header_strains:
            fh.write(f"{graph.strain_names[s]}" + "".join(f";{int(v)}" for v in res.hist[which, s]) + "\n")


# ---- unassigned reads ---------------------------------------------------------------------------------
def booked_events(reads, res):
    """(group, mate, cluster) of every read recorded as unassigned, in read order
    (json2csv.py:884-891, 903-910, 1037-1044)."""
    code = res.grp_code
    lo, hi = code & 15, code >> 4
    out = []
    for g in np.flatnonzero((lo == _lib.UNASSIGNED_BOOK) | (hi == _lib.UNASSIGNED_BOOK)):
        for m, c in ((0, lo[g]), (1, hi[g])):
            if c == _lib.UNASSIGNED_BOOK:
                a = int(reads.grp_aln_off[2 * g + m])
                out.append((int(g), m, -(int(res.aln_assign[a]) + 2)))
    return out


def cluster_records(graph, reads, res):
    """Per-cluster unassigned-read records; node ids are the GFA's (the reference keeps ``int(node_id)``)."""
    clusters = [dict(reads_name=[], unassigned_paths=[], unassigned_abund={}, count=0, count_norm=0)
                for _ in range(graph.n_clusters)]
    wo = reads.aln_walk_off
    for g, m, cid in booked_events(reads, res):
        a = int(reads.grp_aln_off[2 * g + m])
        lo, hi = int(wo[a]), int(wo[a + 1])
        dense = reads.walk_node[lo:hi]
        walk = [int(v) for v in graph.node_ids[dense]]
        packed = reads.walk_alen[lo:hi].astype(np.int64)
        neg = packed < 0
        cov = np.where(neg, 0, packed & 0xFFFF) / np.where(neg, 1, packed >> 16)     # aligned bases / node length
        if neg.any():
            cov[neg] = reads.exc_cov[-packed[neg] - 1]
        rec = clusters[cid]
        rec["reads_name"].append(mate_name(reads, g, m))
        rec["unassigned_paths"].append(walk)
        rec["count"] += 1
        rec["count_norm"] += 1 / int(reads.aln_read_len[a])
        ab = rec["unassigned_abund"]
        for node, v in zip(walk, cov):
            ab[node] = ab.get(node, 0) + float(v)
    return clusters


def min_strains(cluster_walks, threads=None):
    """Lower bounds on the number of strains explaining each cluster's unassigned reads (json2csv.py:149-201), native
    (csrc/unassigned.cpp).  ``cluster_walks``: per cluster, the list of walks (lists of node ids).  Returns four
    arrays over the clusters: largest clique of the incompatibility graph, reads kept by the support filter, largest
    clique among them, mean per-node support of the kept reads (NaN when none)."""
    import ctypes as C
    import os
    n = len(cluster_walks)
    counts = np.fromiter((len(w) for w in cluster_walks), np.int64, n)
    clu_off = np.concatenate([[0], np.cumsum(counts, dtype=np.int64)]).astype(np.int64)
    lens = np.fromiter((len(w) for ws in cluster_walks for w in ws), np.int64, int(clu_off[-1]))
    walk_off = np.concatenate([[0], np.cumsum(lens, dtype=np.int64)]).astype(np.int64)
    nodes = np.fromiter((v for ws in cluster_walks for w in ws for v in w), np.int64, int(walk_off[-1]))
    before, kept, after = (np.zeros(n, np.int64) for _ in range(3))
    support = np.zeros(n, np.float64)
    rc = _lib.load().sfb_min_strains(n, clu_off.ctypes.data, walk_off.ctypes.data, nodes.ctypes.data,
                                     int(threads or len(os.sched_getaffinity(0))), before.ctypes.data, kept.ctypes.data,
                                     after.ctypes.data, support.ctypes.data)
    if rc:
        raise _lib.SfbError(f"sfb_min_strains failed ({rc})")
    return before, kept, after, support


def write_unassigned(path, graph, clusters, strain_counts):
    """strain_counts: int64[C, S] from sfb_cluster_strain_counts."""
    keys = list(graph.header_strains)
    with open(path, "w") as fh:
        fh.write("Clusters;" + "".join(f"{graph.strain_names[s]};" for s in keys)
                 + "nb_reads_unfilt;min_strains_unfilt;nb_reads_afterfilt;min_strains;mean_abund\n")
        before, kept, after, support = min_strains([rec["unassigned_paths"] for rec in clusters])
        for cid, rec in enumerate(clusters):
            tail = f"{before[cid]};{kept[cid]};{after[cid]};{float(support[cid])}" if rec["count"] != 0 else "0;0;0;0"
            fh.write(f"Cluster_{cid};" + "".join(f"{int(strain_counts[cid, s])};" for s in keys)
                     + f"{len(rec['unassigned_paths'])};{tail}\n")


def write_clusters_pickle(path, clusters):
    obj = {f"cluster_{i}": dict(reads_name=c["reads_name"], unassigned_paths=c["unassigned_paths"],
                                unassigned_abund=dict(c["unassigned_abund"])) for i, c in enumerate(clusters)}
    with open(path, "wb") as fh:
        pickle.dump(obj, fh)


def summary_lines(C, thr, prefix):
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
        f"In total, {used} ({used / nb * 100}%) reads were used.",          # ZeroDivisionError on an empty file, as :1356
    ]
