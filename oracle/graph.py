"""Oracle (test infrastructure): GFA + cluster pickle -> graph arrays.

Restates ``Pangenome.build_pangenome`` (reference json2csv.py:223-298),
``canonical`` (:46-61), ``get_accession_number`` (:73-81) and
``Pangenome.fill_clusters_info`` (:300-319) and emits the interchange schema
described in DESIGN.md ("Graph arrays").

Interchange schema (dict, numpy arrays unless noted)
    node_ids      int64[Nn]   original GFA node id of dense node index k
    node_len      int32[Nn]   sequence length of node k
    path_off      int64[P+1]  CSR offsets into path_nodes
    path_nodes    int32[T]    dense node index per path position (signs dropped)
    path_seq_len  int64[P]    sum of node_len over the path (:220-221)
    pstr_off      int64[P+1]  CSR offsets of the path's strain list
    pstr_id       int32[]     strain index, in dict-insertion order of strain_ids
    pstr_cnt      int32[]     copy count of that strain on that path
    path_cluster  int32[P]    cluster id, -1 for None
    clu_off       int64[C+1]  CSR offsets of cluster -> colored path ids
    clu_paths     int32[]     path ids (duplicates kept, :317)
    strain_names  list        strain index -> accession (may contain None);
                              first-seen order over *all* P lines
    header_strains list[int]  strain indices in `species_names`, in first-seen
                              order (only sightings that created a new path, :268)
    path_names    dict        gene (P line) name -> path id (:265,295)
"""
import pickle
import re

import numpy as np

_ACC = re.compile(r"[a-zA-Z]+_?[a-zA-Z]*[0-9]+")


def accession_of(gene_name):
    """Strain id carried by a path name (json2csv.py:73-81)."""
    stem = "_".join(gene_name.split("_")[:-1])
    for piece in stem.split("|"):
        if _ACC.match(piece):
            return piece
    return None


def canonical_nodes(spec):
    """Orientation-canonical form of a P-line node list (json2csv.py:46-61).

    Returns the canonical *string* (the dedup key, signs included)."""
    items = spec.split(",")
    if len(items) == 1:
        return items[0][:-1] + "+"
    if int(items[0][:-1]) < int(items[-1][:-1]):
        return spec
    flip = {"+": "-"}
    return ",".join(tok[:-1] + flip.get(tok[-1], "+") for tok in reversed(items))


def load_graph(gfa_path, pickle_path=None):
    node_index = {}          # GFA id -> dense index
    node_ids, node_len = [], []
    by_content = {}          # canonical string -> pid
    path_names = {}
    paths = []               # list of dense-node lists
    path_strains = []        # list of {strain index: copies} (insertion ordered)
    strain_index = {}        # accession -> strain index
    strain_names = []
    header_strains = []

    def strain_of(name):
        acc = accession_of(name)
        if acc not in strain_index:
            strain_index[acc] = len(strain_names)
            strain_names.append(acc)
        return strain_index[acc]

    with open(gfa_path, "r") as fh:
        for raw in fh:
            cols = raw.strip().split("\t")
            tag = cols[0]
            if tag == "S":
                nid = int(cols[1])
                if nid in node_index:            # dict overwrite (:251)
                    node_len[node_index[nid]] = len(cols[2])
                else:
                    node_index[nid] = len(node_ids)
                    node_ids.append(nid)
                    node_len.append(len(cols[2]))
            elif tag == "P":
                key = canonical_nodes(cols[2])
                sid = strain_of(cols[1])
                pid = by_content.get(key)
                if pid is not None:              # duplicate content (:260-266)
                    d = path_strains[pid]
                    d[sid] = d.get(sid, 0) + 1
                    path_names[cols[1]] = pid
                    continue
                if sid not in header_strains:    # species_names.add (:268)
                    header_strains.append(sid)
                nodes = [node_index[int(tok[:-1])] for tok in key.split(",")]  # KeyError as :282
                pid = len(paths)
                paths.append(nodes)
                path_strains.append({sid: 1})
                path_names[cols[1]] = pid
                by_content[key] = pid

    P = len(paths)
    path_off = np.zeros(P + 1, np.int64)
    for i, p in enumerate(paths):
        path_off[i + 1] = path_off[i] + len(p)
    path_nodes = np.fromiter((n for p in paths for n in p), np.int32, count=int(path_off[-1]))
    node_len_a = np.asarray(node_len, np.int32)
    seq_len = np.array([int(node_len_a[p].sum()) if len(p) else 0 for p in paths], np.int64) \
        if P else np.zeros(0, np.int64)

    pstr_off = np.zeros(P + 1, np.int64)
    pstr_id, pstr_cnt = [], []
    for i, d in enumerate(path_strains):
        for s, c in d.items():
            pstr_id.append(s)
            pstr_cnt.append(c)
        pstr_off[i + 1] = len(pstr_id)

    g = dict(
        node_ids=np.asarray(node_ids, np.int64), node_len=node_len_a,
        path_off=path_off, path_nodes=path_nodes, path_seq_len=seq_len,
        pstr_off=pstr_off, pstr_id=np.asarray(pstr_id, np.int32),
        pstr_cnt=np.asarray(pstr_cnt, np.int32),
        path_cluster=np.full(P, -1, np.int32),
        clu_off=np.zeros(1, np.int64), clu_paths=np.zeros(0, np.int32),
        strain_names=strain_names, header_strains=header_strains,
        path_names=path_names, node_index=node_index,
    )
    if pickle_path is not None:
        attach_clusters(g, pickle_path)
    return g


def attach_clusters(g, pickle_path):
    """``fill_clusters_info`` (json2csv.py:300-319)."""
    with open(pickle_path, "rb") as fh:
        table = pickle.load(fh)
    clu_off = [0]
    clu_paths = []
    for cid in range(len(table)):
        for gene in table["Cluster_" + str(cid)]["genes_list"]:
            pid = g["path_names"].get(gene)
            if pid is None:
                continue
            clu_paths.append(pid)
            g["path_cluster"][pid] = cid         # last writer wins (:318)
        clu_off.append(len(clu_paths))
    g["clu_off"] = np.asarray(clu_off, np.int64)
    g["clu_paths"] = np.asarray(clu_paths, np.int32)
    return g
