"""Tiles of the graph for the shared-memory kernels (include/sfb200.h, ``sfb_tiles``).

A tile is a set of colored paths with consecutive ids whose nodes have consecutive dense indices, closed in the sense
that no path outside the tile crosses one of its nodes.  StrainFLAIR builds one variation graph per gene cluster and
concatenates their id spaces (graphs_construction.py:88, concat_graphs.py), so a cluster -- or a few small neighbouring
clusters -- is a tile; a graph whose clusters interleave their node ids simply has fewer (or no) tiles and its reads go
through the global-memory kernels.  Nothing here depends on ``dict_clusters.pickle``: the tiles come from the paths.
"""
import os

import numpy as np

SLOT_LIMIT, PATH_LIMIT, NODE_LIMIT = 32767, 255, 65520       # what the 15/8/16-bit local ids of the kernels can hold
ITEM_GROUPS = 4096                                            # read groups per work item


class Tiling:
    """node_tile int32[N] (-1 = none), tile_desc int32[NT, 8] = (p0, p1, n0, n1, flags, 0, 0, 0), sizes of the largest
    tile, and the edge index of the "simple" tiles (flags bit 0: at most 64 paths and no node twice in a path): for
    every node its distinct successors on the tile's paths with the set of paths taking that step as a 64-bit mask
    (bit = path id - p0), and the set of paths through the node.  A walk matches a path iff the path takes every
    step of the walk, so get_matching_path (json2csv.py:341-363) becomes an AND over the steps."""

    def __init__(self, graph, slot_cap=4096, node_cap=4096, path_cap=PATH_LIMIT, merge_target=640, tuple_cap=None):
        # (slot_cap / node_cap keep the shared-memory image of the largest tile, which sizes every launch, moderate)
        P, N = graph.n_paths, graph.n_nodes
        slot_cap, node_cap, path_cap = min(slot_cap, SLOT_LIMIT), min(node_cap, NODE_LIMIT), min(path_cap, PATH_LIMIT)
        self.node_tile = np.full(N, -1, np.int32)
        self.tile_desc = np.zeros((0, 8), np.int32)
        self.edge_off = np.zeros(N + 1, np.int32)
        self.edge_succ = np.zeros(0, np.int32)
        self.edge_mask = np.zeros(0, np.uint64)
        self.node_mask = np.zeros(N, np.uint64)
        self.max_edges = 0
        self.max_slots = self.max_nodes = self.max_paths = 0
        if P and N:
            self._build(graph, slot_cap, node_cap, path_cap, merge_target)
        k = os.environ.get("SFB200_TILE_K")
        if tuple_cap is None and k:
            tuple_cap = int(k)
        if tuple_cap is None:          # match tuples per alignment kept in shared memory: about one per path of the tile;
            # simple tiles keep path sets instead, so a graph of simple tiles takes the smallest store
            all_simple = self.n_tiles > 0 and bool(self.tile_desc[:, 4].all())
            tuple_cap = 8 if all_simple or self.max_paths <= 9 else 16 if self.max_paths <= 20 else 32
        self.tuple_cap = int(tuple_cap)

    @property
    def n_tiles(self):
        return len(self.tile_desc)

    def _build(self, graph, slot_cap, node_cap, path_cap, merge_target):
        P, N = graph.n_paths, graph.n_nodes
        off = graph.path_off.astype(np.int64)
        plen = np.diff(off)
        nodes = graph.path_nodes.astype(np.int64)
        has = plen > 0
        lo, hi = np.zeros(P, np.int64), np.zeros(P, np.int64)
        if has.any():
            starts = off[:-1][has]
            lo[has] = np.minimum.reduceat(nodes, starts)
            hi[has] = np.maximum.reduceat(nodes, starts)
            # a path without nodes joins the tile of the previous path with nodes (the next one for leading ones)
            idx = np.where(has, np.arange(P), -1)
            np.maximum.accumulate(idx, out=idx)
            first = int(np.argmax(has))
            idx[idx < 0] = first
            lo, hi = lo[idx], hi[idx]
        else:
            return
        # node ranges of the paths merge into segments of the node id space: a segment starts at a covered node that no
        # range reaches from the left, and ends at a covered node that no range leaves to the right
        starts_at = np.bincount(lo, minlength=N + 1)[:N]
        ends_at = np.bincount(hi, minlength=N + 1)[:N]
        inside = np.cumsum(starts_at - np.concatenate(([0], ends_at[:-1])))      # ranges containing node n
        covered = inside > 0
        start = covered & (inside - starts_at == 0)
        end = covered & (inside - ends_at == 0)
        seg_of_node = np.cumsum(start) - 1
        n_seg = int(start.sum())
        seg_lo = np.flatnonzero(start)
        seg_hi = np.flatnonzero(end) + 1                                  # exclusive
        seg_of_path = seg_of_node[lo]
        pmin = np.full(n_seg, P, np.int64)
        pmax = np.full(n_seg, -1, np.int64)
        ids = np.arange(P, dtype=np.int64)
        np.minimum.at(pmin, seg_of_path, ids)
        np.maximum.at(pmax, seg_of_path, ids)
        count = np.bincount(seg_of_path, minlength=n_seg)
        slots = np.bincount(seg_of_path, weights=plen + 1, minlength=n_seg).astype(np.int64)
        ok = (count == pmax - pmin + 1) & (slots <= slot_cap) & (seg_hi - seg_lo <= node_cap) & (count <= path_cap)
        # a segment is "simple" when it has at most 64 paths and none of them visits a node twice
        seg_dup = np.zeros(n_seg, bool)
        self._dup = self._dup_paths(graph)
        seg_dup[seg_of_path[self._dup]] = True
        seg_simple = (count <= 64) & ~seg_dup
        # neighbouring small segments (in node order AND path order) merge up to merge_target slots; a simple segment only
        # merges with simple ones, and only while the merged tile stays simple
        desc = []
        cur = None
        for s in range(n_seg):
            if not ok[s]:
                if cur is not None:
                    desc.append(cur)
                    cur = None
                continue
            item = [int(pmin[s]), int(pmax[s]) + 1, int(seg_lo[s]), int(seg_hi[s]), int(slots[s]), bool(seg_simple[s])]
            if cur is not None and cur[1] == item[0] and cur[4] + item[4] <= merge_target \
                    and item[3] - cur[2] <= node_cap and item[1] - cur[0] <= path_cap \
                    and cur[5] == item[5] and (not item[5] or item[1] - cur[0] <= 64):
                cur = [cur[0], item[1], cur[2], item[3], cur[4] + item[4], cur[5]]
            else:
                if cur is not None:
                    desc.append(cur)
                cur = item
        if cur is not None:
            desc.append(cur)
        if not desc:
            return
        desc = np.asarray([d[:5] for d in desc], np.int64)
        self.tile_desc = np.zeros((len(desc), 8), np.int32)
        self.tile_desc[:, :4] = desc[:, :4]
        self.max_slots = int(desc[:, 4].max())
        self.max_nodes = int((desc[:, 3] - desc[:, 2]).max())
        self.max_paths = int((desc[:, 1] - desc[:, 0]).max())
        span = desc[:, 3] - desc[:, 2]
        self.node_tile[np.repeat(desc[:, 2], span) + (np.arange(int(span.sum())) - np.repeat(np.cumsum(span) - span, span))] = \
            np.repeat(np.arange(len(desc), dtype=np.int32), span)
        self._edges(graph, desc)

    @staticmethod
    def _dup_paths(graph):
        """Ids of the paths that visit a node more than once."""
        P, N = graph.n_paths, graph.n_nodes
        path_of = np.repeat(np.arange(P, dtype=np.int64), np.diff(graph.path_off.astype(np.int64)))
        srt = np.sort(path_of * N + graph.path_nodes.astype(np.int64))
        if len(srt) < 2:
            return np.zeros(0, np.int64)
        return np.unique(srt[1:][srt[1:] == srt[:-1]] // N)

    def _edges(self, graph, desc):
        """Edge index of the simple tiles (see the class docstring)."""
        P, N = graph.n_paths, graph.n_nodes
        off = graph.path_off.astype(np.int64)
        plen = np.diff(off)
        nodes = graph.path_nodes.astype(np.int64)
        path_of = np.repeat(np.arange(P, dtype=np.int64), plen)
        tile_of_path = np.full(P, -1, np.int64)
        np_t = desc[:, 1] - desc[:, 0]
        tile_of_path[np.repeat(desc[:, 0], np_t) + (np.arange(int(np_t.sum())) - np.repeat(np.cumsum(np_t) - np_t, np_t))] = \
            np.repeat(np.arange(len(desc)), np_t)
        # a node twice in one path?  (path, node) pairs must be distinct
        dup_path = np.zeros(P, bool)
        dup_path[self._dup] = True
        bad_tile = np.zeros(len(desc), bool)
        tp = tile_of_path[dup_path]
        bad_tile[tp[tp >= 0]] = True
        simple = (np_t <= 64) & ~bad_tile
        self.tile_desc[:, 4] = simple.astype(np.int32)
        # occurrences on paths of simple tiles
        t_occ = tile_of_path[path_of]
        use = t_occ >= 0
        use[use] = simple[t_occ[use]]
        bit = np.left_shift(np.uint64(1), (path_of - desc[np.maximum(t_occ, 0), 0]).clip(0, 63).astype(np.uint64))
        nm = np.zeros(N, np.uint64)
        np.bitwise_or.at(nm, nodes[use], bit[use])
        self.node_mask = nm
        # steps: occurrence -> next occurrence of the same path
        last = np.zeros(len(nodes), bool)
        last[off[1:][plen > 0] - 1] = True
        step = use & ~last
        a, b, m = nodes[step], nodes[np.flatnonzero(step) + 1], bit[step]
        ekey = a * N + b
        order = np.argsort(ekey, kind="stable")
        ekey, m = ekey[order], m[order]
        first = np.ones(len(ekey), bool)
        first[1:] = ekey[1:] != ekey[:-1]
        starts = np.flatnonzero(first)
        self.edge_mask = np.bitwise_or.reduceat(m, starts) if len(starts) else np.zeros(0, np.uint64)
        ua = ekey[starts] // N
        self.edge_succ = (ekey[starts] % N).astype(np.int32)
        self.edge_off = np.concatenate(([0], np.cumsum(np.bincount(ua, minlength=N)))).astype(np.int32)
        per_tile = self.edge_off[desc[:, 3]] - self.edge_off[desc[:, 2]]
        self.max_edges = int(per_tile.max()) if len(per_tile) else 0

    STEP_DTYPE = np.dtype([("succ0", "<i4"), ("succ1", "<i4"), ("mask0", "<u8"), ("mask1", "<u8"), ("p0", "<i4"), ("n_succ", "<i4")])

    def step_records(self, graph):
        """The per-node records of ``sfb_steps`` (include/sfb200.h) and whether any node with occurrences lies outside
        the simple tiles (its walks then need the occurrence-index kernel)."""
        if getattr(self, "_steps", None) is not None:
            return self._steps
        N = graph.n_nodes
        rec = np.zeros(N, self.STEP_DTYPE)
        rec["succ0"] = rec["succ1"] = -1
        rec["p0"] = -1
        t = self.node_tile
        simple = np.zeros(N, bool)
        if self.n_tiles:
            inside = t >= 0
            simple[inside] = self.tile_desc[t[inside], 4] != 0
            rec["p0"][simple] = self.tile_desc[t[simple], 0]
        crossed = np.bincount(graph.path_nodes, minlength=N) > 0
        rec["p0"][~crossed] = 0                     # no path through the node: no match, whatever the rest of the walk
        n_succ = np.diff(self.edge_off).astype(np.int32)
        rec["n_succ"] = n_succ
        e0 = self.edge_off[:-1].astype(np.int64)
        for k, (sk, mk) in enumerate((("succ0", "mask0"), ("succ1", "mask1"))):
            has = n_succ > k
            rec[sk][has] = self.edge_succ[e0[has] + k]
            rec[mk][has] = self.edge_mask[e0[has] + k]
        self._steps = (rec, bool((crossed & ~simple).any()))
        return self._steps

    def items(self, tile_grp_off):
        """Work items int32[n, 4] = (tile, first group, end group, 0) of a read shard in tile order."""
        cnt = np.diff(tile_grp_off[:self.n_tiles + 1])
        n_chunk = (cnt + ITEM_GROUPS - 1) // ITEM_GROUPS
        total = int(n_chunk.sum())
        out = np.zeros((total, 4), np.int32)
        if total:
            tile = np.repeat(np.arange(self.n_tiles, dtype=np.int64), n_chunk)
            k = np.arange(total, dtype=np.int64) - np.repeat(np.cumsum(n_chunk) - n_chunk, n_chunk)
            g0 = tile_grp_off[tile] + k * ITEM_GROUPS
            out[:, 0] = tile
            out[:, 1] = g0
            out[:, 2] = np.minimum(g0 + ITEM_GROUPS, tile_grp_off[tile + 1])
        return out
