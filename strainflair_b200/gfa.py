"""GFA + cluster pickle -> Graph arrays (host side, one pass per graph).

Reproduces what the reference's ``Pangenome.build_pangenome`` and ``fill_clusters_info`` build
(json2csv.py:223-319) in the array schema of ``strainflair_b200.schema.Graph``:

* ``S`` lines give node lengths (ids are arbitrary positive ints -> dense indices in first-seen order);
* ``P`` lines are canonicalised (json2csv.py:46-61), merged when their canonical node string is equal
  (:260-266, strain copy counts add up) and numbered in first-seen order;
* the strain of a path is the accession found in its name (:73-81); the header of the gene table lists
  only strains that created a path (``species_names``, :268), which is what makes Q5's ragged rows;
* the pickle maps ``Cluster_<k>`` -> gene names -> path ids (:310-318, last writer wins, duplicates kept).
"""
import pickle
import re

import numpy as np

from .schema import Graph

_ACCESSION = re.compile(r"[a-zA-Z]+_?[a-zA-Z]*[0-9]+")
_FLIP = {"+": "-", "-": "+"}


def strain_of_name(name):
    head = name.rsplit("_", 1)[0] if "_" in name else ""
    for part in head.split("|"):
        if _ACCESSION.match(part):
            return part
    return None


def _canonical(field):
    toks = field.split(",")
    if len(toks) == 1:
        return toks[0][:-1] + "+", toks
    if int(toks[0][:-1]) < int(toks[-1][:-1]):
        return field, toks
    toks = [t[:-1] + _FLIP.get(t[-1], "+") for t in reversed(toks)]
    return ",".join(toks), toks


def load_graph(gfa_path, pickle_path=None, progress=None) -> Graph:
    node_index = {}
    node_ids, node_len = [], []
    seen = {}
    path_names = {}
    path_chunks = []            # per path: int32 array of dense nodes
    path_strains = []           # per path: {strain index: copies}, insertion ordered
    strains = {}
    strain_names = []
    header = []
    in_header = set()
    with open(gfa_path, "r") as fh:
        for n_line, raw in enumerate(fh):
            if progress is not None and n_line % 10000 == 0:
                progress(fh)
            f = raw.strip().split("\t")
            kind = f[0]
            if kind == "S":
                nid = int(f[1])
                k = node_index.get(nid)
                if k is None:
                    node_index[nid] = len(node_ids)
                    node_ids.append(nid)
                    node_len.append(len(f[2]))
                else:
                    node_len[k] = len(f[2])
            elif kind == "P":
                key, toks = _canonical(f[2])
                acc = strain_of_name(f[1])
                sid = strains.get(acc)
                if sid is None:
                    sid = strains[acc] = len(strain_names)
                    strain_names.append(acc)
                pid = seen.get(key)
                if pid is not None:
                    d = path_strains[pid]
                    d[sid] = d.get(sid, 0) + 1
                    path_names[f[1]] = pid
                    continue
                if sid not in in_header:
                    in_header.add(sid)
                    header.append(sid)
                pid = len(path_chunks)
                path_chunks.append(np.fromiter((node_index[int(t[:-1])] for t in toks), np.int32, count=len(toks)))
                path_strains.append({sid: 1})
                seen[key] = pid
                path_names[f[1]] = pid
    P = len(path_chunks)
    path_off = np.zeros(P + 1, np.int64)
    if P:
        np.cumsum([len(c) for c in path_chunks], out=path_off[1:])
    path_nodes = np.concatenate(path_chunks) if P else np.zeros(0, np.int32)
    node_len = np.asarray(node_len, np.int32)
    csum = np.concatenate(([0], np.cumsum(node_len[path_nodes], dtype=np.int64)))
    seq_len = csum[path_off[1:]] - csum[path_off[:-1]]
    pstr_off = np.zeros(P + 1, np.int64)
    pstr_id, pstr_cnt = [], []
    for p, d in enumerate(path_strains):
        pstr_id.extend(d.keys())
        pstr_cnt.extend(d.values())
        pstr_off[p + 1] = len(pstr_id)
    g = Graph(
        node_ids=np.asarray(node_ids, np.int64), node_len=node_len, path_off=path_off, path_nodes=path_nodes,
        path_seq_len=seq_len, pstr_off=pstr_off, pstr_id=np.asarray(pstr_id, np.int32),
        pstr_cnt=np.asarray(pstr_cnt, np.int32), path_cluster=np.full(P, -1, np.int32),
        clu_off=np.zeros(1, np.int64), clu_paths=np.zeros(0, np.int32), strain_names=strain_names,
        header_strains=header, path_names=path_names, node_index=node_index,
    )
    if pickle_path is not None:
        attach_clusters(g, pickle_path)
    return g


def attach_clusters(g: Graph, pickle_path):
    with open(pickle_path, "rb") as fh:
        table = pickle.load(fh)
    offs, members = [0], []
    for k in range(len(table)):
        for gene in table["Cluster_" + str(k)]["genes_list"]:       # KeyError on a gap, like the reference (:312)
            pid = g.path_names.get(gene)
            if pid is not None:
                members.append(pid)
                g.path_cluster[pid] = k
        offs.append(len(members))
    g.clu_off = np.asarray(offs, np.int64)
    g.clu_paths = np.asarray(members, np.int32)
    return g
