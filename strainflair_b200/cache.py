"""On-disk caches of the two host-side loaders: the Graph arrays next to an
``all_graphs.gfa`` (an index-side artefact: parsed once, queried many times) and the decoded Reads CSR of an alignment
JSON (decoded once per (file, graph, threshold)).

Opt-in: nothing is written unless the environment variable ``SFB200_CACHE_DIR`` names a directory (``"."`` = next to
the input file).  The command lines stay drop-in: no new flag.  A cache entry is one uncompressed ``.npz``; its file
name carries a key made of the input's absolute path, size and mtime, the parameters that shape the result and the
format version, so a changed input or parameter simply misses.  Entries are written to a temporary name and renamed.
"""
import hashlib
import os
import json

import numpy as np

from .schema import GRAPH_ARRAYS, READS_ARRAYS, Graph, Reads

FORMAT = 1


def cache_dir(for_file):
    d = os.environ.get("SFB200_CACHE_DIR")
    if not d:
        return None
    return os.path.dirname(os.path.abspath(for_file)) if d == "." else d


def _key(path, *extra):
    st = os.stat(path)
    h = hashlib.sha1()
    h.update(repr((os.path.abspath(path), st.st_size, st.st_mtime_ns, FORMAT) + extra).encode())
    return h.hexdigest()[:20]


def _entry(path, kind, *extra):
    d = cache_dir(path)
    if d is None:
        return None
    return os.path.join(d, f"{os.path.basename(path)}.{kind}.{_key(path, *extra)}.sfb200.npz")


def _save(entry, arrays):
    os.makedirs(os.path.dirname(entry), exist_ok=True)
    tmp = f"{entry}.{os.getpid()}.tmp.npz"
    try:
        np.savez(tmp, **arrays)
        os.replace(tmp, entry)
    except OSError:
        if os.path.exists(tmp):
            os.remove(tmp)                   # a read-only or full cache directory is not an error of the job


def _blob(obj):
    return np.frombuffer(json.dumps(obj).encode("utf-8"), np.uint8)


def graph_fingerprint(graph: Graph):
    """What the decoder reads of a graph: node ids and lengths."""
    h = hashlib.sha1()
    h.update(np.ascontiguousarray(graph.node_ids, np.int64).tobytes())
    h.update(np.ascontiguousarray(graph.node_len, np.int32).tobytes())
    return h.hexdigest()[:16]


# ---- graph (before the cluster pickle is attached: that one is a small, separate input) ----------------------
def load_graph_entry(gfa_path):
    entry = _entry(gfa_path, "graph")
    if entry is None or not os.path.exists(entry):
        return None
    try:
        with np.load(entry, allow_pickle=False) as z:
            kw = {k: z[k] for k in GRAPH_ARRAYS}
            extras = json.loads(z["extras"].tobytes().decode("utf-8"))     # plain JSON: a cache entry is never executed
            extras["strain_names"] = [None if v is None else str(v) for v in extras["strain_names"]]
            extras["header_strains"] = [int(v) for v in extras["header_strains"]]
            if extras.get("path_names") is not None:
                extras["path_names"] = {str(k): int(v) for k, v in extras["path_names"].items()}
    except Exception:
        return None                          # unreadable or foreign entry: parse again
    P = len(kw["path_off"]) - 1
    kw["path_cluster"] = np.full(P, -1, np.int32)
    kw["clu_off"], kw["clu_paths"] = np.zeros(1, np.int64), np.zeros(0, np.int32)
    return Graph(**kw, **extras, node_index=None)


def save_graph_entry(gfa_path, graph: Graph):
    entry = _entry(gfa_path, "graph")
    if entry is None:
        return
    arrays = {k: getattr(graph, k) for k in GRAPH_ARRAYS}
    arrays["extras"] = _blob(dict(strain_names=[None if v is None else str(v) for v in graph.strain_names],
                                  header_strains=[int(v) for v in graph.header_strains],
                                  path_names=None if graph.path_names is None else
                                  {str(k): int(v) for k, v in graph.path_names.items()}))
    _save(entry, arrays)


# ---- decoded reads ---------------------------------------------------------------------------------------------
_READS_EXTRA = ("name_off", "aln_stats")


def load_reads_entry(json_path, graph: Graph, thr, want_cov):
    entry = _entry(json_path, "reads", float(thr), graph_fingerprint(graph))
    if entry is None or not os.path.exists(entry):
        return None
    try:
        with np.load(entry, allow_pickle=False) as z:
            if want_cov and "walk_cov" not in z.files:
                return None
            kw = {k: z[k] for k in READS_ARRAYS}
            reads = Reads(**kw, mate_names=None)
            reads.name_off, reads.aln_stats = z["name_off"], z["aln_stats"]
            reads.name_blob = z["name_blob"].tobytes()
            if want_cov:
                reads.walk_cov = z["walk_cov"]
    except Exception:
        return None                          # unreadable entry: decode again
    return reads


def save_reads_entry(json_path, graph: Graph, thr, reads: Reads):
    entry = _entry(json_path, "reads", float(thr), graph_fingerprint(graph))
    if entry is None:
        return
    arrays = {k: getattr(reads, k) for k in READS_ARRAYS}
    arrays["name_off"], arrays["aln_stats"] = reads.name_off, reads.aln_stats
    arrays["name_blob"] = np.frombuffer(reads.name_blob, np.uint8)
    if getattr(reads, "walk_cov", None) is not None:
        arrays["walk_cov"] = reads.walk_cov
    _save(entry, arrays)
