"""Host orchestration of the abundance pass on one B200 (one process per GPU).

PyTorch is used for device buffers, streams and (optionally) torch.distributed collectives only;
every computation is a hand-written kernel of libsfb200.so called through ctypes (include/sfb200.h).

Pipeline (reference json2csv.py:811-1279 + :399-474), per read shard:
    match -> classify -> pass-1 scatter -> [sum uniq_reads over shards] -> pass 2
    -> [sum accumulators over shards] -> path reduce
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .schema import Graph, Reads


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _to_dev(a, device, pin=False):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if pin:
        t = t.pin_memory()
    return t.to(device, non_blocking=True)


class DeviceGraph:
    """The graph in the device layout of include/sfb200.h: every path followed by one sentinel slot,
    plus the node -> slot occurrence index that replaces Node.traversed_path and the position scan
    (json2csv.py:352-357)."""

    def __init__(self, graph: Graph, device, tiles=True):
        self.graph = graph
        self.device = torch.device(device)
        P = graph.n_paths
        T = graph.n_path_nodes
        plen = np.diff(graph.path_off)
        n_slots = T + P
        if n_slots > _lib.MAX_SLOTS:
            raise ValueError(f"graph too large for 30-bit slot codes: {n_slots} slots")
        # plain position -> slot
        self.slot_of_plain = (np.arange(T, dtype=np.int64) + np.repeat(np.arange(P, dtype=np.int64), plen))
        slot_node = np.full(n_slots, -1, np.int32)
        slot_node[self.slot_of_plain] = graph.path_nodes
        slot_path = np.repeat(np.arange(P, dtype=np.int32), plen + 1)
        path_off = (graph.path_off + np.arange(P + 1, dtype=np.int64)).astype(np.int32)
        order = np.argsort(graph.path_nodes, kind="stable")
        occ_slot = self.slot_of_plain[order].astype(np.int32)
        occ_pair = np.empty((T, 2), np.int32)
        occ_pair[:, 0] = occ_slot
        occ_pair[:, 1] = slot_node[occ_slot.astype(np.int64) + 1]      # successor on the path (-1 at the path end)
        # slot -> path table per block of 64 slots: (path of the block's first slot, first slot inside the block where
        # the next path starts | INT_MAX when none | -1 when several paths start in the block: walk path_off then)
        n_blk = (n_slots + 63) // 64
        blk_lo = np.arange(n_blk, dtype=np.int64) * 64
        starts = path_off[:-1].astype(np.int64)
        first = np.searchsorted(starts, blk_lo, side="right")               # index of the first path start > block start
        last = np.searchsorted(starts, blk_lo + 63, side="right")           # starts <= block end
        nxt = np.full(n_blk, 0x7FFFFFFF, np.int64)
        one = (last - first) == 1
        nxt[one] = starts[first[one]]
        nxt[(last - first) > 1] = -1
        blk_path = np.stack([slot_path[::64][:n_blk].astype(np.int64), nxt], axis=1).astype(np.int32)
        occ_off = np.zeros(graph.n_nodes + 1, np.int64)
        np.cumsum(np.bincount(graph.path_nodes, minlength=graph.n_nodes), out=occ_off[1:])
        S = graph.n_strains
        words = max(1, (S + 63) // 64)
        smask = np.zeros((P, words), np.uint64)
        owner = np.repeat(np.arange(P, dtype=np.int64), np.diff(graph.pstr_off))
        sid = graph.pstr_id.astype(np.int64)
        np.bitwise_or.at(smask, (owner, sid >> 6), np.left_shift(np.uint64(1), (sid & 63).astype(np.uint64)))
        nstrain = np.diff(graph.pstr_off).astype(np.int32)
        strain0 = np.zeros(P, np.int32)
        has = nstrain > 0
        strain0[has] = graph.pstr_id[graph.pstr_off[:-1][has]]
        self.n_slots, self.n_paths, self.n_strains, self.mask_words = n_slots, P, S, words
        self.path_off_host = path_off
        d = self.device
        # tiles of the graph for the shared-memory kernels (tiles.py); none when the graph has no closed parts that fit
        from .tiles import Tiling
        import os
        self.tiling = None
        if tiles:
            # (the tiling only depends on the graph: engines on the same Graph object share it)
            slot_cap = int(os.environ.get("SFB200_TILE_SLOTS", "4096"))
            key = (slot_cap, os.environ.get("SFB200_TILE_K"))
            cache = graph.__dict__.setdefault("_tiling_cache", {})
            if key not in cache:
                cache[key] = Tiling(graph, slot_cap=slot_cap)
            self.tiling = cache[key]
        self.steps = None                        # pointers of the step index inside sfb_graph, when the graph has simple tiles
        if self.tiling is not None and self.tiling.n_tiles == 0:
            self.tiling = None
        if self.tiling is not None:
            tl = self.tiling
            self.tile_desc = _to_dev(tl.tile_desc.reshape(-1), d)
            self.tile_edges = dict(edge_off=_to_dev(tl.edge_off, d), edge_succ=_to_dev(tl.edge_succ, d),
                                   edge_mask=_to_dev(tl.edge_mask.view(np.int64), d),
                                   node_mask=_to_dev(tl.node_mask.view(np.int64), d))
            probe = _lib.SfbTiles(n_tiles=tl.n_tiles, n_items=0, max_slots=tl.max_slots, max_nodes=tl.max_nodes,
                                  max_paths=tl.max_paths, max_edges=tl.max_edges, tuple_cap=tl.tuple_cap)
            need = int(_lib.load().sfb_tile_smem_bytes(C.byref(probe)))
            # step index of the simple tiles for sfb_match_steps (needs no shared memory: kept whatever the tile kernels say)
            if tl.tile_desc[:, 4].any():
                rec, partial = tl.step_records(graph)
                self.node_step = _to_dev(rec.view(np.int64), d)
                self.steps = dict(node_step=self.node_step.data_ptr(), node_mask=self.tile_edges["node_mask"].data_ptr(),
                                  step_off=self.tile_edges["edge_off"].data_ptr(), step_succ=self.tile_edges["edge_succ"].data_ptr(),
                                  step_mask=self.tile_edges["edge_mask"].data_ptr(), steps_partial=int(partial))
            if need < 0 or need > 227 * 1024:
                self.tiling = None
        self.t = dict(
            slot_node=_to_dev(slot_node, d), path_off=_to_dev(path_off, d), occ_off=_to_dev(occ_off.astype(np.int32), d), occ_pair=_to_dev(occ_pair.reshape(-1), d), blk_path=_to_dev(blk_path.reshape(-1), d),
            path_seq_len=_to_dev(graph.path_seq_len, d), path_nstrain=_to_dev(nstrain, d),
            path_strain0=_to_dev(strain0, d), path_smask=_to_dev(smask.view(np.int64), d),
            path_cluster=_to_dev(graph.path_cluster, d),
        )
        self.c = _lib.SfbGraph(n_nodes=graph.n_nodes, n_paths=P, n_slots=n_slots, n_strains=S, mask_words=words,
                               **{k: v.data_ptr() for k, v in self.t.items()})
        # the same graph with its step index: sfb_match_steps and the group kernels on path sets
        self.c_steps = None if self.steps is None else _lib.SfbGraph(
            n_nodes=graph.n_nodes, n_paths=P, n_slots=n_slots, n_strains=S, mask_words=words,
            **{k: v.data_ptr() for k, v in self.t.items()}, **self.steps)

    def nbytes(self):
        return sum(v.numel() * v.element_size() for v in self.t.values())


READS_DEVICE_FIELDS = ("grp_nmates", "grp_aln_off", "mate_seq_len", "aln_walk_off", "aln_bins", "walk_node",
                       "walk_alen", "exc_cov")


WIRE_FIELDS = ("grp_nmates", "mate_naln", "mate_seq_len", "aln_len", "aln_bins", "walk_node", "walk_al", "exc_idx",
               "exc_cov")


class _Packed:
    """Handle on the native packer's result (csrc/wire.cpp): sizes first, then copies straight into the arena."""

    def __init__(self, reads: Reads, threads=None):
        import os
        self.lib = _lib.load()
        keep = {k: np.ascontiguousarray(getattr(reads, k), dt).reshape(-1) for k, dt in (
            ("grp_nmates", np.uint8), ("grp_aln_off", np.int64), ("mate_seq_len", np.int32), ("aln_walk_off", np.int64),
            ("aln_bins", np.uint8), ("walk_node", np.int32), ("walk_alen", np.int32), ("exc_cov", np.float64))}
        view = _lib.SfbReads(n_groups=reads.n_groups, n_alns=reads.n_alns, n_records=reads.n_records,
                             n_exc=len(keep["exc_cov"]), **{k: v.ctypes.data for k, v in keep.items()})
        self.h = self.lib.sfb_wire_pack(C.byref(view), int(threads or len(os.sched_getaffinity(0))))
        msg = C.c_char_p()
        self.status = self.lib.sfb_packed_status(self.h, C.byref(msg))
        if self.status > 1:
            text = (msg.value or b"").decode()
            self.close()
            raise MemoryError(text)
        self.nbytes = [int(self.lib.sfb_packed_size(self.h, k)) for k in range(len(_lib.WIRE_ARRAYS))] if self.status == 0 else None
        self.seq_len_common = int(self.lib.sfb_packed_size(self.h, 100))

    def copy_to(self, which, ptr):
        self.lib.sfb_packed_copy(self.h, which, C.c_void_p(ptr))

    def close(self):
        if self.h:
            self.lib.sfb_packed_free(self.h)
            self.h = None


def wire_arrays(reads: Reads, threads=None):
    """Compact transfer layout (include/sfb200.h, sfb_wire) of a read shard as numpy arrays: (arrays, seq_len_common),
    or None when it does not apply (a mate with > 255 alignments, a walk of > 255 nodes, a read of > 65535 bases, or
    so many records needing an explicit coverage / a wide node step that the plain layout is smaller)."""
    p = _Packed(reads, threads)
    try:
        if p.status:
            return None
        arrays = {}
        for k, (name, dt) in enumerate(zip(_lib.WIRE_ARRAYS, _lib.WIRE_DTYPES)):
            arrays[name] = np.empty(p.nbytes[k] // np.dtype(dt).itemsize, dt)
            if p.nbytes[k]:
                p.copy_to(k, arrays[name].ctypes.data)
        return arrays, p.seq_len_common
    finally:
        p.close()


def _layout(sizes):
    """Offsets of the arrays (name -> bytes) inside one arena (256-byte aligned), and the arena size."""
    offs, pos = {}, 0
    for k, n in sizes.items():
        offs[k] = (pos, int(n))
        pos += (int(n) + 255) & ~255
    return offs, max(pos, 256)


class HostReads:
    """Reads staged in ONE pinned host arena, so that the upload of a shard is a single DMA at full PCIe rate
    (the end-to-end path copies from here every step).  ``wire``: use the compact transfer layout when the shard
    allows it (offsets as 1-byte counts, 16-bit lengths; expanded on the device by sfb_expand_reads)."""

    def __init__(self, reads: Reads, pin=True, wire=True, tiling=None, tile_grp_off=None):
        # ``tiling`` (tiles.Tiling of the graph; AbundanceEngine.stage passes its own): the groups are staged in tile
        # order -- those whose alignments all start and end inside one tile first, tile by tile, the rest after them in
        # file order -- with the work items of the tile kernels.  The per-path results do not depend on the order;
        # per-group / per-alignment outputs are put back in file order by fetch.
        # ``tile_grp_off``: the groups are in tile order already (a slice of a permuted data set, stage_shards), this is
        # where each tile's groups start in it.
        self.order = self.aln_index = None
        self.tiling, self.items, self.grp_begin, self.aln_begin = tiling, None, 0, 0
        if tiling is not None and reads.n_groups > 0:
            if tile_grp_off is None:
                reads, self.order, self.aln_index, tile_grp_off = tile_permute(reads, tiling)
            self.items = tiling.items(tile_grp_off)
            self.grp_begin = int(tile_grp_off[tiling.n_tiles])
            self.aln_begin = int(reads.grp_aln_off[2 * self.grp_begin])
        self.reads = reads
        packed = _Packed(reads) if wire else None
        try:
            self.wire = packed is not None and packed.status == 0
            if self.wire:
                sizes = dict(zip(_lib.WIRE_ARRAYS, packed.nbytes))
                self.seq_len_common = packed.seq_len_common
                self.n_exc, self.n_wide, self.n_seq_exc, self.n_bins_exc = \
                    (sizes[k] // 8 for k in ("exc_cov", "wide_idx", "seq_exc_idx", "bins_exc_idx"))
            else:
                arrays = {k: np.ascontiguousarray(getattr(reads, k)).reshape(-1) for k in READS_DEVICE_FIELDS}
                sizes = {k: a.nbytes for k, a in arrays.items()}
                self.seq_len_common = 0
                self.n_exc, self.n_wide, self.n_seq_exc, self.n_bins_exc = len(reads.exc_cov), 0, 0, 0
            if self.items is not None:
                sizes["tile_items"] = self.items.nbytes
            self.offs, self.size = _layout(sizes)
            self.arena = torch.empty(self.size, dtype=torch.uint8, pin_memory=bool(pin))
            view = self.arena.numpy()
            for k, (name, (o, n)) in enumerate(self.offs.items()):
                if not n:
                    continue
                if name == "tile_items":
                    view[o:o + n] = self.items.reshape(-1).view(np.uint8)
                elif self.wire:
                    packed.copy_to(k, view.ctypes.data + o)
                else:
                    view[o:o + n] = arrays[name].view(np.uint8)
        finally:
            if packed is not None:
                packed.close()
        self.n_groups, self.n_alns, self.n_records = reads.n_groups, reads.n_alns, reads.n_records

    def nbytes(self):
        return sum(n for _, n in self.offs.values())


class DeviceReads:
    """Device copy of a HostReads arena (one DMA), expanded on the device when it is in the transfer layout.
    ``arena`` / ``wide``: existing device buffers to reuse."""

    def __init__(self, host: HostReads, device, arena=None, wide=None, node_len=None, expand=True, tile_desc=None,
                 tile_edges=None):
        self.host = host
        self.tile_desc, self.tile_edges = tile_desc, tile_edges
        if arena is None or arena.numel() < host.size:
            arena = torch.empty(host.size, dtype=torch.uint8, device=device)
        self.arena = arena
        self.arena[:host.size].copy_(host.arena, non_blocking=True)          # the one H2D copy of the shard
        self.device, self.wide, self.node_len = device, wide, node_len
        self.n_groups, self.n_alns, self.n_records, self.n_exc = host.n_groups, host.n_alns, host.n_records, host.n_exc
        self.c = None
        if expand:
            self.expand()

    def expand(self):
        """Build the sfb_reads view (device-side expansion of the transfer layout, on the current stream)."""
        host, device, wide, node_len = self.host, self.device, self.wide, self.node_len
        base = self.arena.data_ptr()
        G, A, R, E = host.n_groups, host.n_alns, host.n_records, host.n_exc
        self.n_groups, self.n_alns, self.n_records, self.n_exc = G, A, R, E
        ptr = {k: (base + o if n else 0) for k, (o, n) in host.offs.items()}
        items_ptr = ptr.pop("tile_items", 0)
        self.tiles_c = None
        if host.items is not None and len(host.items) and self.tile_desc is not None:
            tl = host.tiling
            self.tiles_c = _lib.SfbTiles(n_tiles=tl.n_tiles, n_items=len(host.items), max_slots=tl.max_slots,
                                         max_nodes=tl.max_nodes, max_paths=tl.max_paths, max_edges=tl.max_edges,
                                         tuple_cap=tl.tuple_cap, tile_desc=self.tile_desc.data_ptr(), items=items_ptr,
                                         **{k: v.data_ptr() for k, v in self.tile_edges.items()})
        self.grp_begin, self.aln_begin = (host.grp_begin, host.aln_begin) if self.tiles_c is not None else (0, 0)
        self.wide = wide
        if not host.wire:
            self.c = _lib.SfbReads(n_groups=G, n_alns=A, n_records=R, n_exc=E, **ptr)
            return
        # expanded arrays (one device allocation, reused across steps)
        need = _lib.workspace(_lib.WS_EXPANDED, n_groups=G, n_alns=A, n_records=R)
        if wide is None or wide.numel() < need:
            wide = torch.empty(need, dtype=torch.uint8, device=device)
        self.wide = wide
        p, out = wide.data_ptr(), {}
        for k, n in (("grp_aln_off", 8 * (2 * G + 1)), ("aln_walk_off", 8 * (A + 1)), ("scratch", 8 * (max(2 * G, A) // 4096 + 2)),
                     ("mate_seq_len", 8 * G), ("aln_bins", 5 * A), ("walk_node", 4 * R), ("walk_alen", 4 * R)):
            out[k] = p
            p += (n + 255) & ~255
        w = _lib.SfbWire(n_groups=G, n_alns=A, n_records=R, n_exc=E, n_wide=host.n_wide, n_seq_exc=host.n_seq_exc,
                         n_bins_exc=host.n_bins_exc, seq_len_common=host.seq_len_common, **ptr)
        self.c = _lib.SfbReads(n_groups=G, n_alns=A, n_records=R, n_exc=E, grp_nmates=ptr["grp_nmates"],
                               grp_aln_off=out["grp_aln_off"], mate_seq_len=out["mate_seq_len"] if G else 0,
                               aln_walk_off=out["aln_walk_off"], aln_bins=out["aln_bins"] if A else 0,
                               walk_node=out["walk_node"] if R else 0, walk_alen=out["walk_alen"] if R else 0,
                               exc_cov=ptr["exc_cov"])
        _lib.check(_lib.load().sfb_expand_reads(C.byref(w), C.c_void_p(node_len.data_ptr()), C.byref(self.c),
                                                C.c_void_p(out["scratch"]), _stream()))


class Work:
    """Per-shard work arrays (sfb_work)."""

    def __init__(self, n_groups, n_alns, match_cap, device, app_cap=0):
        i32, u8 = torch.int32, torch.uint8
        self.n_groups, self.n_alns = n_groups, n_alns
        self.match_cap = int(max(2, match_cap))
        # pass-2 work items: one per match that receives a share.  With tuple lists that is at most n_alns + match_cap;
        # with path sets (sfb_match_steps) the matches are not in match_buf, so the default is a guess and a pass that
        # needs more reports its exact need (cursor[2]) for the rerun
        self.app_cap = max(int(n_alns + self.match_cap), int(app_cap), _lib.workspace(_lib.WS_APP_ITEMS, n_alns=n_alns))
        self.t = dict(
            aln_match=torch.empty(max(1, n_alns), dtype=torch.int64, device=device),      # (x, y) pairs
            aln_sets=torch.empty(2 * max(1, n_alns), dtype=torch.int64, device=device),   # forward / reverse path sets
            match_buf=torch.empty(self.match_cap, dtype=torch.int64, device=device),       # (code, pid) pairs
            app_buf=torch.empty(2 * self.app_cap, dtype=torch.int64, device=device),       # 16-byte items
            cursor=torch.zeros(_lib.CURSOR_N, dtype=torch.int64, device=device),
            slow=torch.empty(max(1, n_groups), dtype=i32, device=device),
            grp_code=torch.zeros(max(1, n_groups), dtype=u8, device=device),
            grp_code2=torch.zeros(max(1, n_groups), dtype=u8, device=device),
            aln_assign=torch.empty(max(1, n_alns), dtype=i32, device=device),
            deferred=torch.empty(max(1, n_groups), dtype=i32, device=device),
        )
        self.c = _lib.SfbWork(match_cap=self.match_cap, app_cap=self.app_cap,
                              **{k: v.data_ptr() for k, v in self.t.items()})

    def reset(self):
        self.t["cursor"].zero_()
        self.t["grp_code2"].zero_()


class Accum:
    """Accumulators (sfb_accum), one flat fp64 and one flat int64 allocation so that a shard
    exchange is two collectives."""

    def __init__(self, dg: DeviceGraph, deterministic=False):
        P, S, n_slots = dg.n_paths, dg.n_strains, dg.n_slots
        dev = dg.device
        self.P, self.S, self.n_slots = P, S, n_slots
        # fp64 block: uniq_norm | mult_reads | mult_norm | uniq_abund | mult_abund; in the order-independent mode two
        # such planes of 64-bit INTEGERS (96-bit fixed point: units of 2^-32 and of 2^-64, see sfb200.h), collapsed
        # into the fp64 values of plane 0 before the path reduction
        self.F = 3 * P + 2 * n_slots
        self.planes = 2 if deterministic else 1
        dims = dict(n_paths=P, n_slots=n_slots, n_strains=S, deterministic=int(deterministic))
        assert _lib.workspace(_lib.WS_ACCUM_F64, **dims) == self.planes * self.F + 1
        self.f64_all = torch.zeros(self.planes * self.F + 1, dtype=torch.float64, device=dev)   # + one element that stays 0
        self.f64 = self.f64_all[:self.F]
        # int64 block: uniq_reads | hist | counters | mult_hits (uint32 pairs packed in the tail)
        self.n_hist = _lib.N_HIST * S * _lib.N_BINS
        self.i64 = torch.zeros(_lib.workspace(_lib.WS_ACCUM_I64, **dims), dtype=torch.int64, device=dev)
        f, i = self.f64, self.i64
        self.uniq_norm, self.mult_reads, self.mult_norm = f[0:P], f[P:2 * P], f[2 * P:3 * P]
        self.uniq_abund, self.mult_abund = f[3 * P:3 * P + n_slots], f[3 * P + n_slots:]
        self.uniq_reads = i[0:P]
        self.hist = i[P:P + self.n_hist]
        self.counters = i[P + self.n_hist:P + self.n_hist + _lib.N_COUNTERS]
        self.mult_hits = i[P + self.n_hist + _lib.N_COUNTERS:].view(torch.int32)
        self.c = _lib.SfbAccum(
            uniq_reads=self.uniq_reads.data_ptr(), uniq_norm=self.uniq_norm.data_ptr(),
            mult_reads=self.mult_reads.data_ptr(), mult_norm=self.mult_norm.data_ptr(),
            mult_hits=self.mult_hits.data_ptr(), uniq_abund=self.uniq_abund.data_ptr(),
            mult_abund=self.mult_abund.data_ptr(), hist=self.hist.data_ptr(), counters=self.counters.data_ptr(),
            det_stride=self.F if deterministic else 0)

    def zero(self):
        self.f64_all.zero_()
        self.i64.zero_()

    def flat(self):
        """The whole fp64 block as the tensor to sum over ranks: int64 in the order-independent mode (exact), else fp64."""
        return self.f64_all.view(torch.int64) if self.planes > 1 else self.f64_all

    def plane(self, k):
        return self.flat()[k * self.F:(k + 1) * self.F]

    def exchange(self, which, bounds, comm):
        """The reduce-scatter of one slot array (0 uniq_abund, 1 mult_abund; every plane) onto the path-partitioned layout,
        built once per accumulator set."""
        from . import dist as sfdist
        key = "_rs%d" % which
        if getattr(self, key, None) is None:
            flat, P, n = self.flat(), self.P, self.n_slots
            arrays = [self.plane(k)[3 * P + which * n:3 * P + (which + 1) * n] for k in range(self.planes)]
            setattr(self, key, sfdist.RangeReduceScatter(flat, arrays, bounds, comm))
        return getattr(self, key)


class Result:
    """Host copy of everything the reports need; attribute names follow the oracle's."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


class MatchOverflow(RuntimeError):
    pass


def locality_permute(reads: Reads, threads=None):
    """(reads in locality order, order, aln_index): schema.Reads.locality_order + permute_groups, done by the native
    multi-threaded sfb_locality_permute (csrc/locality.cpp)."""
    import os
    G, A, R = reads.n_groups, reads.n_alns, reads.n_records
    src = {k: np.ascontiguousarray(getattr(reads, k), dt).reshape(-1) for k, dt in (
        ("grp_nmates", np.uint8), ("grp_aln_off", np.int64), ("mate_seq_len", np.int32), ("aln_walk_off", np.int64),
        ("aln_bins", np.uint8), ("walk_node", np.int32), ("walk_alen", np.int32), ("exc_cov", np.float64))}
    dst = dict(grp_nmates=np.empty(G, np.uint8), grp_aln_off=np.empty(2 * G + 1, np.int64),
               mate_seq_len=np.empty(2 * G, np.int32), aln_walk_off=np.empty(A + 1, np.int64),
               aln_bins=np.empty(5 * A, np.uint8), walk_node=np.empty(R, np.int32), walk_alen=np.empty(R, np.int32),
               exc_cov=src["exc_cov"])
    order, aln_index = np.empty(G, np.int64), np.empty(A, np.int64)
    view = lambda d: _lib.SfbReads(n_groups=G, n_alns=A, n_records=R, n_exc=len(src["exc_cov"]),       # noqa: E731
                                   **{k: v.ctypes.data for k, v in d.items()})
    vs, vd = view(src), view(dst)
    _lib.check(_lib.load().sfb_locality_permute(C.byref(vs), int(threads or len(os.sched_getaffinity(0))),
                                                C.c_void_p(order.ctypes.data), C.c_void_p(aln_index.ctypes.data), C.byref(vd)))
    names = None
    if reads.mate_names is not None:
        names = [reads.mate_names[2 * int(g) + m] for g in order for m in (0, 1)]
    moved = Reads(grp_nmates=dst["grp_nmates"], grp_aln_off=dst["grp_aln_off"], mate_seq_len=dst["mate_seq_len"],
                  aln_walk_off=dst["aln_walk_off"], aln_read_len=reads.aln_read_len[aln_index],
                  aln_bins=dst["aln_bins"].reshape(A, 5), walk_node=dst["walk_node"], walk_alen=dst["walk_alen"],
                  exc_cov=reads.exc_cov, mate_names=names)
    return moved, order, aln_index


def tile_permute(reads: Reads, tiling, threads=None):
    """(reads in tile order, order, aln_index, tile_grp_off): the native sfb_tile_permute (csrc/locality.cpp)."""
    import os
    G, A, R = reads.n_groups, reads.n_alns, reads.n_records
    src = {k: np.ascontiguousarray(getattr(reads, k), dt).reshape(-1) for k, dt in (
        ("grp_nmates", np.uint8), ("grp_aln_off", np.int64), ("mate_seq_len", np.int32), ("aln_walk_off", np.int64),
        ("aln_bins", np.uint8), ("walk_node", np.int32), ("walk_alen", np.int32), ("exc_cov", np.float64))}
    dst = dict(grp_nmates=np.empty(G, np.uint8), grp_aln_off=np.empty(2 * G + 1, np.int64),
               mate_seq_len=np.empty(2 * G, np.int32), aln_walk_off=np.empty(A + 1, np.int64),
               aln_bins=np.empty(5 * A, np.uint8), walk_node=np.empty(R, np.int32), walk_alen=np.empty(R, np.int32),
               exc_cov=src["exc_cov"])
    order, aln_index = np.empty(G, np.int64), np.empty(A, np.int64)
    tile_grp_off = np.zeros(tiling.n_tiles + 2, np.int64)
    view = lambda d: _lib.SfbReads(n_groups=G, n_alns=A, n_records=R, n_exc=len(src["exc_cov"]),       # noqa: E731
                                   **{k: v.ctypes.data for k, v in d.items()})
    vs, vd = view(src), view(dst)
    node_tile = np.ascontiguousarray(tiling.node_tile, np.int32)
    _lib.check(_lib.load().sfb_tile_permute(C.byref(vs), C.c_void_p(node_tile.ctypes.data), tiling.n_tiles,
                                            int(threads or len(os.sched_getaffinity(0))), C.c_void_p(order.ctypes.data),
                                            C.c_void_p(aln_index.ctypes.data), C.byref(vd),
                                            C.c_void_p(tile_grp_off.ctypes.data)))
    names = None
    if reads.mate_names is not None:
        names = [reads.mate_names[2 * int(g) + m] for g in order for m in (0, 1)]
    moved = Reads(grp_nmates=dst["grp_nmates"], grp_aln_off=dst["grp_aln_off"], mate_seq_len=dst["mate_seq_len"],
                  aln_walk_off=dst["aln_walk_off"], aln_read_len=reads.aln_read_len[aln_index],
                  aln_bins=dst["aln_bins"].reshape(A, 5), walk_node=dst["walk_node"], walk_alen=dst["walk_alen"],
                  exc_cov=reads.exc_cov, mate_names=names)
    return moved, order, aln_index, tile_grp_off


def stage_shards(reads: Reads, tiling, n_shards, pin=True, wire=True):
    """A data set staged as ``n_shards`` HostReads for AbundanceEngine.submit / run_stream: the read groups are put in tile
    order FIRST and the shards cut at tile boundaries, so that every tile's groups sit in one shard (file-order slices
    staged one by one spread every tile over all shards: as many work items, each a fraction of the size, in every
    shard).  The groups outside the tile range follow in the last shards.  Without a tiling: file-order slices."""
    from . import dist as sfdist
    G = reads.n_groups
    if tiling is None or G == 0:
        return [HostReads(reads.slice_groups(*sfdist.shard_bounds(G, n_shards, k)), pin=pin, wire=wire) for k in range(n_shards)]
    moved, _, _, tile_grp_off = tile_permute(reads, tiling)
    edges = np.asarray(tile_grp_off[:tiling.n_tiles + 1], np.int64)          # group index where each tile starts; [-1] = end of the tile range
    cuts = [0]
    for k in range(1, n_shards):
        want = sfdist.shard_bounds(G, n_shards, k)[0]
        if want < edges[-1]:                                                 # inside the tile range: the nearest tile boundary
            j = int(np.searchsorted(edges, want))
            lo, hi = int(edges[max(j - 1, 0)]), int(edges[min(j, len(edges) - 1)])
            want = lo if want - lo <= hi - want else hi
        cuts.append(max(want, cuts[-1]))
    cuts.append(G)
    out = []
    for g0, g1 in zip(cuts[:-1], cuts[1:]):
        off = np.zeros(tiling.n_tiles + 2, np.int64)
        off[:tiling.n_tiles + 1] = np.clip(edges, g0, g1) - g0
        off[tiling.n_tiles + 1] = g1 - g0
        out.append(HostReads(moved.slice_groups(g0, g1), pin=pin, wire=wire, tiling=tiling, tile_grp_off=off))
    return out


def restore_file_order(res, order, aln_index):
    """Per-group and per-alignment outputs of a pass over permuted read groups (``order[k]`` / ``aln_index[k]`` = old
    index of new group / alignment k), back in the order of the input file."""
    for name in ("grp_code", "grp_code2"):
        v = getattr(res, name, None)
        if v is not None:
            out = np.empty_like(v)
            out[order] = v
            setattr(res, name, out)
    for name in ("aln_assign", "aln_match"):
        v = getattr(res, name, None)
        if v is not None:
            out = np.empty_like(v)
            out[aln_index] = v
            setattr(res, name, out)
    return res


class AbundanceEngine:
    """One engine per process/GPU.  ``comm`` is an optional torch.distributed process group: when
    given, ``run`` treats ``reads`` as this rank's shard of the read groups."""

    def __init__(self, graph: Graph, device="cuda:0", comm=None, deterministic=True, tiles=True, tile_kernels=None,
                 step_index=None, tile_bins=None):
        if not torch.cuda.is_available():
            raise _lib.SfbError("no CUDA device: the abundance engine has no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self.graph = graph
        import os
        if os.environ.get("SFB200_TILES", "1").strip().lower() in ("0", "false", "no", "off"):
            tiles = False
        self.dg = DeviceGraph(graph, self.device, tiles=tiles)
        self.tiling = self.dg.tiling             # None: no tile order, every read group through the global-memory kernels
        # tile order is a layout decision of the staging (neighbouring groups touch the same part of the graph); which
        # kernels then run is a second one: the shared-memory tile kernels, or the global-memory kernels on everything
        # (measured on C3, profiles/README.md r02: the global-memory kernels on tile-ordered reads are the faster ones,
        # so they are the default; SFB200_TILE_KERNELS=1 / tile_kernels=True switches the tile kernels on)
        if tile_kernels is None:
            tile_kernels = os.environ.get("SFB200_TILE_KERNELS", "0").strip().lower() in ("1", "true", "yes", "on")
        self.tile_kernels = bool(tile_kernels) and self.tiling is not None
        # matching through the step index of the simple tiles (sfb_match_steps) wherever the graph has such tiles;
        # step_index=False / SFB200_STEP_INDEX=0: the occurrence-index kernel (sfb_match) for every alignment
        if step_index is None:
            step_index = os.environ.get("SFB200_STEP_INDEX", "1").strip().lower() not in ("0", "false", "no", "off")
        self.step_index = bool(step_index) and self.dg.steps is not None
        self.gc = self.dg.c_steps if self.step_index else self.dg.c          # the sfb_graph every call of this engine gets
        # node-level sums of the tile-ordered groups in shared-memory bins (sfb_tile_scatter1 / sfb_tile_share), on top
        # of the path-set kernels; SFB200_TILE_BINS=0: every sum through the L2 (sfb_pass1_scatter / sfb_pass2_scatter)
        if tile_bins is None:
            tile_bins = os.environ.get("SFB200_TILE_BINS", "1").strip().lower() not in ("0", "false", "no", "off")
        self.tile_bins = bool(tile_bins) and self.step_index and self.tiling is not None and not self.tile_kernels
        self.node_len = _to_dev(graph.node_len, self.device)
        self.comm = comm
        # order-independent sums (96-bit fixed point, the default): bit-identical across runs, shard counts and GPU
        # counts; False = plain fp64 REDs, whose order (hence the last bits) varies from run to run
        self.deterministic = deterministic
        self.acc = Accum(self.dg, deterministic)
        self.table = torch.empty((max(1, graph.n_paths), _lib.TABLE_COLS), dtype=torch.float64, device=self.device)
        self.world = 1 if comm is None else torch.distributed.get_world_size(comm)
        self.rank = 0 if comm is None else torch.distributed.get_rank(comm)
        if comm is not None:
            from . import dist as sfdist
            self.path_bounds = sfdist.path_partition(self.dg.path_off_host, self.world)
            self.slot_bounds = self.dg.path_off_host.astype(np.int64)[self.path_bounds]
        self.work = None
        self.side = torch.cuda.Stream(device=self.device)     # pass-1 scatter runs beside the uniq exchange + pass-2 resolve
        self.overlap = True
        self.match_cap_hint = 0
        self.app_cap_hint = 0
        self.launches = 0

    # ---- stages (each is one kernel launch through the C ABI) -----------------------------
    def _work_for(self, dr):
        cap = max(self.match_cap_hint, _lib.workspace(_lib.WS_MATCH_PAIRS, n_alns=dr.n_alns))   # (code, path id) pairs
        if self.work is None or self.work.n_groups < dr.n_groups or self.work.n_alns < dr.n_alns \
                or self.work.match_cap < cap or self.work.app_cap < self.app_cap_hint:
            self.work = None
            self.work = Work(dr.n_groups, dr.n_alns, cap, self.device, self.app_cap_hint)
        return self.work

    def stage_shards(self, reads: Reads, n_shards, pin=True, wire=True):
        """``reads`` as ``n_shards`` staged shards for submit / run_stream (tile order first, cut at tile boundaries)."""
        return stage_shards(reads, self.tiling, n_shards, pin=pin, wire=wire)

    def stage(self, reads: Reads, pin=True, wire=True):
        """Host staging of a read shard for this engine's graph: tile order + work items, transfer layout, one pinned arena."""
        return HostReads(reads, pin=pin, wire=wire, tiling=self.tiling)

    def _bind(self, dr, work):
        """With the tile kernels, the global-memory kernels own the groups / alignments after the tile range of this
        shard; with the shared-memory bins they match and classify everything and only the node-level sums of the tile
        range go elsewhere (the deferred groups of that range are not listed: sfb_tile_share finds them)."""
        if self.tile_kernels:
            work.c.grp_begin, work.c.aln_begin, work.c.p2_begin = dr.grp_begin, dr.aln_begin, 0
        else:
            work.c.grp_begin, work.c.aln_begin = 0, 0
            work.c.p2_begin = dr.grp_begin if (self.tile_bins and dr.tiles_c is not None) else 0

    def tile_scatter1(self, dr, work):
        """Pass-1 node loops: shared-memory bins for the tile range, the L2 for the rest."""
        if self.tile_bins and dr.tiles_c is not None:
            _lib.check(self.lib.sfb_tile_scatter1(C.byref(self.gc), C.byref(dr.tiles_c), C.byref(dr.c), C.byref(work.c),
                                                  C.byref(self.acc.c), _stream()))
            self.launches += 1
            work.c.aln_begin = dr.aln_begin
            self.pass1_scatter(dr, work)
            work.c.aln_begin = 0
        else:
            self.pass1_scatter(dr, work)

    def tile_share(self, dr, work, uniq_global=None):
        if not (self.tile_bins and dr.tiles_c is not None):
            return
        u = self.acc.uniq_reads if uniq_global is None else uniq_global
        _lib.check(self.lib.sfb_tile_share(C.byref(self.gc), C.byref(dr.tiles_c), C.byref(dr.c), C.byref(work.c),
                                           C.byref(self.acc.c), C.c_void_p(u.data_ptr()), _stream()))
        self.launches += 1

    def tile_pass1(self, dr, work):
        if dr.tiles_c is None or not self.tile_kernels:
            return
        _lib.check(self.lib.sfb_tile_pass1(C.byref(self.gc), C.byref(dr.tiles_c), C.byref(dr.c), C.byref(work.c),
                                           C.byref(self.acc.c), _stream()))
        self.launches += 1

    def tile_pass2(self, dr, work, uniq_global=None):
        if dr.tiles_c is None or not self.tile_kernels:
            return self.tile_share(dr, work, uniq_global)
        u = self.acc.uniq_reads if uniq_global is None else uniq_global
        _lib.check(self.lib.sfb_tile_pass2(C.byref(self.gc), C.byref(dr.tiles_c), C.byref(dr.c), C.byref(work.c),
                                           C.byref(self.acc.c), C.c_void_p(u.data_ptr()), _stream()))
        self.launches += 1

    def match(self, dr, work):
        if self.step_index:
            _lib.check(self.lib.sfb_match_steps(C.byref(self.gc), C.byref(dr.c), C.byref(work.c), C.byref(self.acc.c), _stream()))
            self.launches += 2 if self.dg.steps["steps_partial"] else 1
            return
        _lib.check(self.lib.sfb_match(C.byref(self.gc), C.byref(dr.c), C.byref(work.c), C.byref(self.acc.c), _stream()))
        self.launches += 1

    def classify(self, dr, work):
        _lib.check(self.lib.sfb_classify(C.byref(self.gc), C.byref(dr.c), C.byref(work.c), C.byref(self.acc.c), _stream()))
        self.launches += 3       # register-path kernel, tuple lists for the generic kernel, generic kernel

    def pass1_scatter(self, dr, work):
        _lib.check(self.lib.sfb_pass1_scatter(C.byref(self.gc), C.byref(dr.c), C.byref(work.c), C.byref(self.acc.c), _stream()))
        self.launches += 1

    def pass2(self, dr, work, uniq_global=None):
        u = self.acc.uniq_reads if uniq_global is None else uniq_global
        _lib.check(self.lib.sfb_pass2(C.byref(self.gc), C.byref(dr.c), C.byref(work.c), C.byref(self.acc.c),
                                      C.c_void_p(u.data_ptr()), _stream()))
        self.launches += 4

    def pass2_resolve(self, dr, work, uniq_global=None):
        u = self.acc.uniq_reads if uniq_global is None else uniq_global
        _lib.check(self.lib.sfb_pass2_resolve(C.byref(self.gc), C.byref(dr.c), C.byref(work.c), C.byref(self.acc.c),
                                              C.c_void_p(u.data_ptr()), _stream()))
        self.launches += 3

    def pass2_scatter(self, dr, work):
        _lib.check(self.lib.sfb_pass2_scatter(C.byref(self.gc), C.byref(dr.c), C.byref(work.c), C.byref(self.acc.c), _stream()))
        self.launches += 1

    def collapse(self):
        acc = self.acc
        _lib.check(self.lib.sfb_accum_collapse(C.c_void_p(acc.f64_all.data_ptr()), acc.F, acc.F, _stream()))
        self.launches += 1

    def path_reduce(self, p_begin=0, p_end=None):
        p_end = self.graph.n_paths if p_end is None else p_end
        out = self.table[p_begin:p_end]
        _lib.check(self.lib.sfb_path_reduce(C.byref(self.gc), C.byref(self.acc.c), p_begin, p_end,
                                            C.c_void_p(out.data_ptr()), _stream()))
        self.launches += 1
        return out

    # ---- the abundance pass ----------------------------------------------------------------
    def run_device(self, dr: DeviceReads, marks=None):
        """Whole pass on reads already resident in HBM; leaves results on the device.
        ``marks``: optional list that receives (stage name, CUDA event recorded after the stage)."""
        import torch.distributed as dist
        from . import dist as sfdist

        def mark(name):
            if marks is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((name, ev))

        work = self._work_for(dr)
        mark("start")
        work.reset()
        self.acc.zero()
        mark("zero")
        self._bind(dr, work)
        self.tile_pass1(dr, work)
        mark("tile_pass1")
        self.match(dr, work)
        mark("match")
        self.classify(dr, work)
        mark("classify")
        # exchange 1 first: the node loops below fill every SM (persistent grids), a collective enqueued behind them waits
        main = torch.cuda.current_stream()
        uniq_global = None
        if self.comm is not None:
            uniq_global = self.acc.uniq_reads.clone()
            dist.all_reduce(uniq_global, group=self.comm)          # exact integer sum over shards
            mark("allreduce_uniq")
        # the node loop of pass 1 only writes uniq_abund, which nothing reads before the path reduction:
        # it runs beside pass 2 on a second stream
        if self.overlap:
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                self.tile_scatter1(dr, work)
        else:
            self.tile_scatter1(dr, work)
            mark("pass1_scatter")
        if self.comm is not None:
            self._exchange_uniq(main)
        self.tile_pass2(dr, work, uniq_global)
        mark("tile_share" if self.tile_bins else "tile_pass2")
        self.pass2_resolve(dr, work, uniq_global)
        mark("pass2_resolve")
        if self.overlap:
            main.wait_stream(self.side)
        self.pass2_scatter(dr, work)
        mark("pass2_scatter")
        self._finish_pass(mark)
        return work

    def _exchange_uniq(self, main):
        """uniq_abund is final once the pass-1 node loops are done: its reduce-scatter onto the path-partitioned layout is
        issued behind them on the side stream and runs under pass 2."""
        with torch.cuda.stream(self.side if self.overlap else main):
            self.acc.exchange(0, self.slot_bounds, self.comm)()

    def _finish_pass(self, mark=lambda name: None):
        """Exchange 2 (multi-GPU) and the per-path reduction."""
        import torch.distributed as dist
        from . import dist as sfdist
        acc = self.acc
        if self.comm is None:
            if acc.planes > 1:
                self.collapse()
            self.path_reduce()
            mark("path_reduce")
            return
        # small per-path state everywhere, slot arrays onto the path-partitioned layout (every plane: the partial
        # sums of the order-independent mode are exact, so the cross-rank sums are too)
        P, n_slots = acc.P, acc.n_slots
        dist.all_reduce(acc.i64, group=self.comm)
        for k in range(acc.planes):
            dist.all_reduce(acc.plane(k)[:3 * P], group=self.comm)
        acc.exchange(1, self.slot_bounds, self.comm)()             # mult_abund (uniq_abund went under pass 2)
        if acc.planes > 1:
            self.collapse()
        mark("reduce_accum")
        p0, p1 = int(self.path_bounds[self.rank]), int(self.path_bounds[self.rank + 1])
        mine = self.path_reduce(p0, p1)
        full = sfdist.all_gather_rows(mine, self.path_bounds, self.comm)
        self.table[:P].copy_(full)
        mark("path_reduce")

    def upload(self, host: HostReads):
        """H2D of a staged shard (+ device-side expansion of the transfer layout), reusing the device buffers."""
        dr = DeviceReads(host, self.device, getattr(self, "_arena", None), getattr(self, "_wide", None), self.node_len,
                         tile_desc=getattr(self.dg, "tile_desc", None) if (self.tile_kernels or self.tile_bins) else None,
                         tile_edges=getattr(self.dg, "tile_edges", None))
        self._arena, self._wide = dr.arena, dr.wide
        return dr

    # ---- end-to-end pass over staged shards ------------------------------------------------
    class _Job:
        """One run of the pass in flight (``submit`` ... ``collect``): its own accumulators, per-shard device buffers,
        result table and pinned read-back buffers, so that two jobs can overlap."""

        def __init__(self, eng):
            self.acc = Accum(eng.dg, eng.deterministic)
            self.table = torch.empty_like(eng.table)
            self.slots, self.stage, self.done, self.shards, self.live = {}, None, None, None, None
            self.pending = False                        # submitted, not collected yet

    def _job_context(self, depth):
        if getattr(self, "_copy", None) is None:
            self._copy = torch.cuda.Stream(device=self.device)
            self._jobs, self._next_job = [], 0
        while len(self._jobs) < depth:
            job = AbundanceEngine._Job(self)
            if not self._jobs:                          # the first context is the engine's own set of buffers
                job.acc, job.table = self.acc, self.table
            self._jobs.append(job)
        job = self._jobs[self._next_job % depth]
        self._next_job += 1
        return job

    def submit(self, shards, depth=2, _reuse=None):
        """Enqueue one whole pass over a data set staged as several HostReads shards (read groups split contiguously)
        and the read-back of its results; returns a job for ``collect``.  Nothing here waits for the GPU: the H2D copy
        of shard k+1 overlaps with match / classify / pass-1 scatter of shard k, pass 2 runs once every shard has
        contributed to ``uniq_reads``, and -- with ``depth`` jobs in flight, each on its own buffers -- the uploads
        of the next job overlap with pass 2 of this one."""
        import torch.distributed as dist
        job = _reuse if _reuse is not None else self._job_context(depth)
        if job.pending and _reuse is None:
            self._next_job -= 1
            raise RuntimeError(f"submit: {depth} jobs are in flight already; collect the oldest one first "
                               "(its buffers would be reused)")
        if job.done is not None:
            job.done.synchronize()                  # the job that used these buffers before must have finished
        main = torch.cuda.current_stream()
        self.acc, self.table = job.acc, job.table
        self.acc.zero()
        if job.done is not None:
            self._copy.wait_event(job.done)         # its buffers are free again
        live = []
        for k, host in enumerate(shards):
            slot = job.slots.setdefault(k, {})
            with torch.cuda.stream(self._copy):
                dr = DeviceReads(host, self.device, slot.get("arena"), slot.get("wide"), self.node_len, expand=False,
                                 tile_desc=getattr(self.dg, "tile_desc", None) if (self.tile_kernels or self.tile_bins) else None,
                         tile_edges=getattr(self.dg, "tile_edges", None))
                arrived = torch.cuda.Event()
                arrived.record(self._copy)
            slot["arena"] = dr.arena
            main.wait_event(arrived)
            dr.expand()
            slot["wide"] = dr.wide
            cap = max(slot.get("cap", 0), _lib.workspace(_lib.WS_MATCH_PAIRS, n_alns=dr.n_alns))
            work = slot.get("work")
            if work is None or work.n_groups < dr.n_groups or work.n_alns < dr.n_alns or work.match_cap < cap \
                    or work.app_cap < slot.get("app", 0):
                work = slot["work"] = Work(dr.n_groups, dr.n_alns, cap, self.device, slot.get("app", 0))
            work.reset()
            self._bind(dr, work)
            self.tile_pass1(dr, work)
            self.match(dr, work)
            self.classify(dr, work)
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                self.tile_scatter1(dr, work)
            live.append((dr, work))
        uniq_global = None
        if self.comm is not None:
            uniq_global = self.acc.uniq_reads.clone()
            dist.all_reduce(uniq_global, group=self.comm)
            self._exchange_uniq(main)
        for dr, work in live:
            self._bind(dr, work)
            self.tile_pass2(dr, work, uniq_global)
            self.pass2_resolve(dr, work, uniq_global)
        main.wait_stream(self.side)
        for dr, work in live:
            self._bind(dr, work)
            self.pass2_scatter(dr, work)
        self._finish_pass()
        # read-back (small, pinned, asynchronous): per-path state, table, every shard's cursors
        P, acc = self.acc.P, self.acc
        if job.stage is None or job.stage["c"].shape[0] < len(live):
            job.stage = dict(
                f=torch.empty(3 * P, dtype=torch.float64).pin_memory(),
                i=torch.empty(acc.i64.numel(), dtype=torch.int64).pin_memory(),
                t=torch.empty((max(1, P), _lib.TABLE_COLS), dtype=torch.float64).pin_memory(),
                c=torch.empty((max(1, len(live)), _lib.CURSOR_N), dtype=torch.int64).pin_memory())
        st = job.stage
        st["f"].copy_(acc.f64[:3 * P], non_blocking=True)
        st["i"].copy_(acc.i64, non_blocking=True)
        st["t"].copy_(self.table, non_blocking=True)
        for k, (_, work) in enumerate(live):
            st["c"][k].copy_(work.t["cursor"], non_blocking=True)
        job.done = torch.cuda.Event()
        job.done.record(main)
        job.shards, job.live, job.pending = shards, live, True
        return job

    def collect(self, job):
        """Wait for a submitted job and return its per-path results (no per-group codes).  A shard whose match buffer
        was too small (first use of a workload) is regrown to its exact need and the job is run again."""
        for attempt in range(3):
            job.done.synchronize()
            res = self._result_from(job.acc, job.stage["f"].numpy(), job.stage["i"].numpy(), job.stage["t"].numpy())
            cur = job.stage["c"].numpy()
            if res.C["MATCH_OVERFLOW"] == 0:
                job.pending = False
                res.n_deferred = int(cur[:len(job.live), 1].sum() + cur[:len(job.live), 7].sum())
                res.match_fill = int(cur[len(job.live) - 1, 0]) if job.live else 0
                res.app_fill = int(cur[:len(job.live), 2].max()) if job.live else 0
                check_errors(res.C)
                return res
            for k in range(len(job.live)):            # grow every shard's buffers to their exact need, rerun
                job.slots[k]["cap"] = int(cur[k, 0]) + 1024
                job.slots[k]["app"] = int(cur[k, 2]) + 1024
            job = self.submit(job.shards, _reuse=job)     # same buffers again
        job.pending = False
        raise MatchOverflow("match buffer still too small after regrowing")

    def run_stream(self, shards):
        """End-to-end pass over a data set staged as several HostReads shards: ``collect(submit(shards))``."""
        return self.collect(self.submit(shards, depth=max(1, len(getattr(self, "_jobs", [])))))

    def run(self, reads, keep_codes=True, keep_slots=True, keep_matches=None):
        """Public entry: host arrays in, host results out (H2D and D2H included).  The reports need the
        per-path table, counters and histograms; ``keep_slots`` also brings the two per-node accumulator
        arrays back, ``keep_codes`` the per-group leaf codes and assignments, ``keep_matches`` (default: as
        ``keep_codes``) the match lists of every alignment -- diagnostics, written out in Python."""
        host = reads if isinstance(reads, HostReads) else self.stage(reads)
        for _ in range(3):
            dr = self.upload(host)
            work = self.run_device(dr)
            res = self.fetch(work, dr, keep_codes, keep_slots, keep_matches)
            over = int(res.counters_raw[_lib.COUNTER_NAMES.index("MATCH_OVERFLOW")])
            if over == 0:
                return res
            self.match_cap_hint = res.match_fill + 1024                    # exact need, rerun
            self.app_cap_hint = res.app_fill + 1024
        raise MatchOverflow("match buffer still too small after regrowing")

    def _sets_to_tuples(self, res, sets, reads):
        """Diagnostics (``keep_codes``): the path sets sfb_match_steps left in ``aln_sets`` written out as the tuple lists
        sfb_match would have produced -- forward matches by ascending path, then reverse; the slot of a tuple is where its
        path has the walk's first (reversed: last) node."""
        if reads is None:
            raise ValueError("the staged reads were dropped: the path sets cannot be turned into match lists")
        g = self.graph
        starts = g.path_off[:-1].astype(np.int64) + np.arange(g.n_paths)                   # sentinel layout
        aln_match, extra = res.aln_match.copy(), []
        base = len(res.match_buf)
        for a in np.flatnonzero(aln_match[:, 0] == _lib.MATCH_SETS):
            p0 = int(aln_match[a, 1])
            w0, w1 = int(reads.aln_walk_off[a]), int(reads.aln_walk_off[a + 1])
            tuples = []
            for word, node, flag in ((int(sets[a, 0]), int(reads.walk_node[w0]), 0),
                                     (int(sets[a, 1]), int(reads.walk_node[w1 - 1]), _lib.MATCH_REV)):
                b = 0
                while word >> b:
                    if (word >> b) & 1:
                        pid = p0 + b
                        nodes = g.path_nodes[g.path_off[pid]:g.path_off[pid + 1]]
                        pos = int(np.flatnonzero(nodes == node)[0])
                        tuples.append((int(starts[pid]) + pos | flag, pid))
                    b += 1
            if len(tuples) == 1:
                aln_match[a] = tuples[0]
            else:
                aln_match[a] = (_lib.MATCH_MULTI | (base + len(extra)), len(tuples))
                extra.extend(tuples)
        res.aln_match = aln_match
        if extra:
            res.match_buf = np.concatenate([res.match_buf, np.asarray(extra, np.uint32).reshape(-1, 2)])

    def _result_from(self, acc, f64, i64, table):
        """Host copies of the read-back buffers as a Result."""
        P = acc.P
        if acc.planes > 1 and P:
            # order-independent mode: terms and per-accumulator totals must stay below 2^31 (include/sfb200.h,
            # sfb_accum.det_stride); the per-path totals bound the per-node ones
            top = max(float(i64[0:P].max()), float(np.nanmax(f64[0:3 * P])))
            if top >= 2.0 ** 31:
                raise OverflowError(f"order-independent accumulation: a per-path total reached {top:.3g} >= 2^31; "
                                    "shard the reads over more passes / GPUs or use deterministic=False")
        counters_raw = i64[P + acc.n_hist:P + acc.n_hist + _lib.N_COUNTERS].copy()
        return Result(
            g=self.graph.as_dict(),
            uniq_reads=i64[0:P].copy(),
            hist=i64[P:P + acc.n_hist].reshape(_lib.N_HIST, acc.S, _lib.N_BINS).copy(),
            counters_raw=counters_raw,
            C={k: int(v) for k, v in zip(_lib.COUNTER_NAMES, counters_raw)},
            mult_hits=i64[P + acc.n_hist + _lib.N_COUNTERS:].view(np.int32)[:P].copy(),
            uniq_norm=f64[0:P].copy(), mult_reads=f64[P:2 * P].copy(), mult_norm=f64[2 * P:3 * P].copy(),
            table=table[:P].copy() if P else np.zeros((0, _lib.TABLE_COLS)),
        )

    def fetch(self, work, dr, keep_codes=True, keep_slots=True, keep_matches=None):
        acc, dg = self.acc, self.dg
        keep_matches = keep_codes if keep_matches is None else (keep_matches and keep_codes)
        P, n_slots = acc.P, acc.n_slots
        if keep_slots and self.comm is not None:
            # whole slot arrays are wanted on the host: complete them from their owner ranks (not part of the pass)
            import torch.distributed as dist
            for arr in (acc.uniq_abund, acc.mult_abund):
                for k in range(self.world):
                    lo, hi = int(self.slot_bounds[k]), int(self.slot_bounds[k + 1])
                    if hi > lo:
                        dist.broadcast(arr[lo:hi], src=dist.get_global_rank(self.comm, k), group=self.comm)
        # small results through pinned staging buffers (one D2H each)
        if getattr(self, "_stage", None) is None:
            self._stage = dict(
                f=torch.empty(3 * P, dtype=torch.float64).pin_memory(),
                i=torch.empty(acc.i64.numel(), dtype=torch.int64).pin_memory(),
                t=torch.empty((max(1, P), _lib.TABLE_COLS), dtype=torch.float64).pin_memory(),
                c=torch.empty(_lib.CURSOR_N, dtype=torch.int64).pin_memory())
        st = self._stage
        st["f"].copy_(acc.f64[:3 * P], non_blocking=True)
        st["i"].copy_(acc.i64, non_blocking=True)
        st["t"].copy_(self.table, non_blocking=True)
        st["c"].copy_(work.t["cursor"], non_blocking=True)
        slots = None
        if keep_slots:
            plain = torch.empty((2, max(1, self.graph.n_path_nodes)), dtype=torch.float64, device=self.device)
            for k, arr in enumerate((acc.uniq_abund, acc.mult_abund)):                   # sentinels dropped on the device
                _lib.check(self.lib.sfb_drop_sentinels(C.byref(self.gc), C.c_void_p(arr.data_ptr()),
                                                       C.c_void_p(plain[k].data_ptr()), _stream()))
                self.launches += 1
            slots = (plain[0, :self.graph.n_path_nodes].cpu(), plain[1, :self.graph.n_path_nodes].cpu())
        torch.cuda.current_stream().synchronize()
        res = self._result_from(acc, st["f"].numpy(), st["i"].numpy(), st["t"].numpy())
        res.n_deferred, res.match_fill, res.app_fill = int(st["c"][1]) + int(st["c"][7]), int(st["c"][0]), int(st["c"][2])
        if slots is not None:
            res.uniq_abund, res.mult_abund = slots[0].numpy(), slots[1].numpy()
        if keep_codes:
            res.grp_code = work.t["grp_code"][:dr.n_groups].cpu().numpy()
            res.grp_code2 = work.t["grp_code2"][:dr.n_groups].cpu().numpy()
            res.aln_assign = work.t["aln_assign"][:dr.n_alns].cpu().numpy()
            if keep_matches and (not self.tile_kernels or getattr(dr, "tiles_c", None) is None):   # (the tile kernels keep the match lists in shared memory)
                res.aln_match = work.t["aln_match"][:dr.n_alns].cpu().numpy().view(np.uint32).reshape(-1, 2)
                res.match_buf = work.t["match_buf"][:min(work.match_cap, max(0, res.match_fill))].cpu().numpy() \
                    .view(np.uint32).reshape(-1, 2)
                if (res.aln_match[:, 0] == _lib.MATCH_SETS).any():
                    sets = work.t["aln_sets"][:2 * dr.n_alns].cpu().numpy().view(np.uint64).reshape(-1, 2)
                    self._sets_to_tuples(res, sets, getattr(getattr(dr, "host", None), "reads", None))
            host = getattr(dr, "host", None)
            if host is not None and getattr(host, "order", None) is not None:
                restore_file_order(res, host.order, host.aln_index)
        check_errors(res.C)
        return res

    def d2h_bytes(self, keep_slots=False):
        """Bytes read back per pass by ``fetch`` (without the optional per-group codes)."""
        n = (3 * self.acc.P + self.acc.i64.numel() + self.table.numel() + 8) * 8
        return n + (2 * 8 * self.graph.n_path_nodes if keep_slots else 0)


def check_errors(Cn):
    """Bug-compatible error behaviour: same exception classes as the reference."""
    if Cn["BAD_BIN"]:
        raise KeyError("alignment ratio outside [0, 1] used as histogram key (json2csv.py:335-339)")
    if Cn["NO_PATH"]:
        raise IndexError("list index out of range")          # nodes[...].traversed_path[0], json2csv.py:875
    if Cn["NO_CLUSTER"]:
        raise TypeError("list indices must be integers or slices, not NoneType")   # clusters[None], json2csv.py:884


def decode_matches(res, a):
    """Match list of alignment ``a`` as [(path id, position, reversed)] (test helper)."""
    x, y = (int(v) for v in res.aln_match[a])
    if x == _lib.MATCH_NONE:
        return []
    if x & _lib.MATCH_MULTI:
        base = x & 0x7FFFFFFF
        pairs = [(int(c), int(p)) for c, p in res.match_buf[base:base + y]]
    else:
        pairs = [(x, y)]
    starts = res.g["path_off"][:-1] + np.arange(len(res.g["path_off"]) - 1)   # sentinel layout: path p starts at path_off[p] + p
    return [(pid, int((c & _lib.SLOT_MASK) - starts[pid]), bool(c & _lib.MATCH_REV)) for c, pid in pairs]


def abundance_pass(graph: Graph, reads: Reads, device="cuda:0", deterministic=True, tiles=True, tile_kernels=None,
                   step_index=None, tile_bins=None) -> Result:
    """One-call public API: decoded alignments + graph in, per-path abundances and counters out."""
    return AbundanceEngine(graph, device, deterministic=deterministic, tiles=tiles, tile_kernels=tile_kernels,
                           step_index=step_index, tile_bins=tile_bins).run(reads)


def cluster_strain_counts(graph: Graph, device="cuda:0"):
    """Strain copy counts per cluster (print_unassigned_info, json2csv.py:544) -> int64[C, S] on the host."""
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise _lib.SfbError("no CUDA device: the abundance engine has no CPU fallback")
    dev = torch.device(device)
    C_, S = graph.n_clusters, graph.n_strains
    out = torch.zeros((max(1, C_), max(1, S)), dtype=torch.int64, device=dev)
    if C_ and S:
        t = [_to_dev(a.astype(np.int32), dev) for a in (graph.clu_off, graph.clu_paths, graph.pstr_off, graph.pstr_id,
                                                         graph.pstr_cnt)]
        _lib.check(lib.sfb_cluster_strain_counts(C_, S, *[C.c_void_p(x.data_ptr()) for x in t],
                                                 C.c_void_p(out.data_ptr()), _stream()))
        torch.cuda.current_stream().synchronize()
    return out[:C_, :S].cpu().numpy()


def strain_aggregate(row_off, row_strain, row_count, mam, mam_nz, ratio_cov, n_strains, thr, device="cuda:0"):
    """compute_strains_abundance.py:52-69 on the device.  Inputs: CSR of the positive strain counts of each
    gene row + three gene-table columns (NaN already replaced by 0).  Returns float64[S, 5]."""
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise _lib.SfbError("no CUDA device: the abundance engine has no CPU fallback")
    dev = torch.device(device)
    P, S = len(mam), int(n_strains)
    out = torch.zeros((max(1, S), 5), dtype=torch.float64, device=dev)
    if S:
        d = lambda a, dt: _to_dev(np.ascontiguousarray(a, dt), dev)    # noqa: E731
        t_off, t_str, t_cnt = d(row_off, np.int32), d(row_strain, np.int32), d(row_count, np.float64)
        t_mam, t_nz, t_cov = d(mam, np.float64), d(mam_nz, np.float64), d(ratio_cov, np.float64)
        ws_rows = torch.empty(max(1, P), dtype=torch.int32, device=dev)
        ws_vals = torch.empty(max(1, 3 * P), dtype=torch.float64, device=dev)
        ws_counts = torch.empty(4 * S, dtype=torch.float64, device=dev)
        p = lambda x: C.c_void_p(x.data_ptr())                        # noqa: E731
        _lib.check(lib.sfb_strain_aggregate(P, S, p(t_off), p(t_str), p(t_cnt), p(t_mam), p(t_nz), p(t_cov), float(thr),
                                            p(ws_rows), p(ws_vals), p(ws_counts), p(out), _stream()))
        torch.cuda.current_stream().synchronize()
    return out[:S].cpu().numpy()
