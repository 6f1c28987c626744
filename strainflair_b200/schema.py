"""Host-side array schema of the abundance engine ("Data layout").

``Graph`` is the reference's ``Pangenome`` after ``build_pangenome`` + ``fill_clusters_info``
(json2csv.py:203-319) flattened to arrays; ``Reads`` is the output of alignment decoding
(json2csv.py:551-739) as CSR.  Both are plain containers of numpy arrays.
"""
import numpy as np

GRAPH_ARRAYS = dict(
    node_ids=np.int64, node_len=np.int32, path_off=np.int64, path_nodes=np.int32, path_seq_len=np.int64,
    pstr_off=np.int64, pstr_id=np.int32, pstr_cnt=np.int32, path_cluster=np.int32, clu_off=np.int64,
    clu_paths=np.int32,
)
READS_ARRAYS = dict(
    grp_nmates=np.uint8, grp_aln_off=np.int64, mate_seq_len=np.int32, aln_walk_off=np.int64,
    aln_read_len=np.int32, aln_bins=np.uint8, walk_node=np.int32, walk_alen=np.int32, exc_cov=np.float64,
)


class _Bag:
    _arrays = {}
    _extras = ()

    def __init__(self, **kw):
        for k, dt in self._arrays.items():
            if k not in kw:
                raise KeyError(f"{type(self).__name__}: missing array {k!r}")
            setattr(self, k, np.ascontiguousarray(kw[k], dtype=dt))
        for k in self._extras:
            setattr(self, k, kw.get(k))

    @classmethod
    def from_dict(cls, d):
        return cls(**{k: d[k] for k in list(cls._arrays) + [e for e in cls._extras if e in d]})

    def as_dict(self):
        out = {k: getattr(self, k) for k in self._arrays}
        out.update({k: getattr(self, k) for k in self._extras if getattr(self, k) is not None})
        return out


class Graph(_Bag):
    """strain_names: index -> accession (may hold None, Q14); header_strains: indices in
    ``species_names`` (first-seen order); path_names: gene name -> path id; node_index: GFA id -> dense."""
    _arrays = GRAPH_ARRAYS
    _extras = ("strain_names", "header_strains", "path_names", "node_index")

    @property
    def n_paths(self):
        return len(self.path_off) - 1

    @property
    def n_nodes(self):
        return len(self.node_len)

    @property
    def n_strains(self):
        return len(self.strain_names)

    @property
    def n_clusters(self):
        return len(self.clu_off) - 1

    @property
    def n_path_nodes(self):
        return int(self.path_off[-1])


class Reads(_Bag):
    """mate_names: optional list of 2G read names (only used by the unassigned report)."""
    _arrays = READS_ARRAYS
    _extras = ("mate_names",)

    def __init__(self, **kw):
        super().__init__(**kw)
        self.aln_bins = self.aln_bins.reshape(-1, 5)

    @property
    def n_groups(self):
        return len(self.grp_nmates)

    @property
    def n_alns(self):
        return len(self.aln_walk_off) - 1

    @property
    def n_records(self):
        return len(self.walk_node)

    @property
    def n_mates(self):
        return int(self.grp_nmates.sum())

    def permute_groups(self, order, return_index=False):
        """The same read groups in another order (``order[k]`` = old index of new group k); mates, alignments and
        records follow their group.  Every sum of the abundance pass is independent of the group order.
        ``return_index``: also return the old index of every new alignment."""
        order = np.asarray(order, np.int64)
        ms = (2 * order[:, None] + np.arange(2)).reshape(-1)                      # mate slots in the new order
        n_aln = np.diff(self.grp_aln_off)[ms]
        grp_aln_off = np.concatenate(([0], np.cumsum(n_aln))).astype(np.int64)
        A = int(grp_aln_off[-1])
        aln = np.repeat(self.grp_aln_off[:-1][ms], n_aln) + (np.arange(A, dtype=np.int64) - np.repeat(grp_aln_off[:-1], n_aln))
        L = np.diff(self.aln_walk_off)[aln]
        aln_walk_off = np.concatenate(([0], np.cumsum(L))).astype(np.int64)
        R = int(aln_walk_off[-1])
        rec = np.repeat(self.aln_walk_off[:-1][aln], L) + (np.arange(R, dtype=np.int64) - np.repeat(aln_walk_off[:-1], L))
        names = None
        if self.mate_names is not None:
            names = [self.mate_names[int(k)] for k in ms]
        moved = Reads(
            grp_nmates=self.grp_nmates[order], grp_aln_off=grp_aln_off, mate_seq_len=self.mate_seq_len[ms],
            aln_walk_off=aln_walk_off, aln_read_len=self.aln_read_len[aln], aln_bins=self.aln_bins[aln],
            walk_node=self.walk_node[rec], walk_alen=self.walk_alen[rec], exc_cov=self.exc_cov, mate_names=names,
        )
        return (moved, aln) if return_index else moved

    def locality_order(self):
        """Stable order of the read groups by the first node of their first retained alignment (groups without one
        last): neighbours in that order start in the same region of the graph -- node ids of a pangenome built
        cluster by cluster are contiguous per cluster -- so the gathers of the pass hit the same index entries, paths
        and accumulator slots."""
        first_aln = self.grp_aln_off[0:-1:2]
        has = self.grp_aln_off[2::2] > first_aln
        key = np.full(self.n_groups, np.iinfo(np.int64).max, np.int64)
        if self.n_alns:
            a = np.minimum(first_aln, self.n_alns - 1)
            w = np.minimum(self.aln_walk_off[:-1][a], max(self.n_records - 1, 0))
            key[has] = self.walk_node[w][has] if self.n_records else 0
        return np.argsort(key, kind="stable")

    def slice_groups(self, g0, g1):
        """Contiguous shard [g0, g1) of read groups, re-based (pairs are never split)."""
        a0, a1 = int(self.grp_aln_off[2 * g0]), int(self.grp_aln_off[2 * g1])
        r0, r1 = int(self.aln_walk_off[a0]), int(self.aln_walk_off[a1])
        alen = self.walk_alen[r0:r1]
        return Reads(
            grp_nmates=self.grp_nmates[g0:g1], grp_aln_off=self.grp_aln_off[2 * g0:2 * g1 + 1] - a0,
            mate_seq_len=self.mate_seq_len[2 * g0:2 * g1], aln_walk_off=self.aln_walk_off[a0:a1 + 1] - r0,
            aln_read_len=self.aln_read_len[a0:a1], aln_bins=self.aln_bins[a0:a1],
            walk_node=self.walk_node[r0:r1], walk_alen=alen, exc_cov=self.exc_cov,
            mate_names=None if self.mate_names is None else self.mate_names[2 * g0:2 * g1],
        )
