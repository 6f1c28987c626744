"""GFA + cluster pickle -> Graph arrays (host side, one pass per graph).

The GFA is parsed by the native loader of libsfb200.so (csrc/gfa.cpp), which reproduces what the reference's
``Pangenome.build_pangenome`` builds (json2csv.py:223-298) in the array schema of ``strainflair_b200.schema.Graph``:

* ``S`` lines give node lengths (ids are arbitrary ints -> dense indices in first-seen order);
* ``P`` lines are canonicalised (json2csv.py:46-61), merged when their canonical node string is equal
  (:260-266, strain copy counts add up) and numbered in first-seen order;
* the strain of a path is the accession found in its name (:73-81); the header of the gene table lists
  only strains that created a path (``species_names``, :268), which is what makes Q5's ragged rows.

The cluster pickle (a Python pickle) is read here: ``Cluster_<k>`` -> gene names -> path ids
(``fill_clusters_info``, :300-319: last writer wins, duplicates kept).
"""
import ctypes as C
import os
import pickle

import numpy as np

from . import _lib
from .schema import Graph

_ERRORS = {2: KeyError, 4: IndexError, 5: OSError, 6: ValueError}


def load_graph(gfa_path, pickle_path=None) -> Graph:
    from . import cache
    g = cache.load_graph_entry(gfa_path)                  # only when SFB200_CACHE_DIR is set
    if g is None:
        g = _parse_gfa(gfa_path)
        cache.save_graph_entry(gfa_path, g)
    if pickle_path is not None:
        attach_clusters(g, pickle_path)
    return g


def _parse_gfa(gfa_path) -> Graph:
    lib = _lib.load()
    h = lib.sfb_load_gfa(os.fsencode(gfa_path))
    try:
        msg = C.c_char_p()
        kind = lib.sfb_gfa_error(h, C.byref(msg))
        if kind:
            raise _ERRORS.get(kind, RuntimeError)((msg.value or b"").decode(errors="replace"))
        size = [int(lib.sfb_gfa_size(h, k)) for k in range(10)]

        def take(which, n, dt):
            out = np.empty(n, dt)
            if n:
                lib.sfb_gfa_copy(h, which, out.ctypes.data)
            return out

        N, P, T, E, S, sbytes, n_genes, gbytes, n_header, none_idx = size
        node_ids, node_len = take(0, N, np.int64), take(1, N, np.int32)
        path_off, path_nodes = take(2, P + 1, np.int64), take(3, T, np.int32)
        pstr_off, pstr_id, pstr_cnt = take(4, P + 1, np.int64), take(5, E, np.int32), take(6, E, np.int32)
        header = [int(s) for s in take(7, n_header, np.int32)]
        s_off, s_blob = take(8, S + 1, np.int64), take(9, sbytes, np.uint8).tobytes()
        strain_names = [s_blob[int(s_off[i]):int(s_off[i + 1])].decode() for i in range(S)]
        if none_idx >= 0:
            strain_names[none_idx] = None                     # name without accession (Q14)
        g_off, g_blob, g_pid = take(10, n_genes + 1, np.int64), take(11, gbytes, np.uint8).tobytes(), take(12, n_genes, np.int32)
        path_names = {g_blob[int(g_off[i]):int(g_off[i + 1])].decode(): int(g_pid[i]) for i in range(n_genes)}
    finally:
        lib.sfb_gfa_free(h)
    csum = np.concatenate(([0], np.cumsum(node_len[path_nodes], dtype=np.int64)))
    g = Graph(
        node_ids=node_ids, node_len=node_len, path_off=path_off, path_nodes=path_nodes,
        path_seq_len=csum[path_off[1:]] - csum[path_off[:-1]], pstr_off=pstr_off, pstr_id=pstr_id, pstr_cnt=pstr_cnt,
        path_cluster=np.full(P, -1, np.int32), clu_off=np.zeros(1, np.int64), clu_paths=np.zeros(0, np.int32),
        strain_names=strain_names, header_strains=header, path_names=path_names, node_index=None,
    )
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
