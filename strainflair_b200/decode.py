"""Alignment decoding: ``vg view -j -K`` JSON -> Reads, through the native streaming decoder of
libsfb200.so (csrc/decode.cpp; reference json2csv.py:551-739).  Decode time is reported separately from the
abundance pass (BASELINE.json north_star)."""
import ctypes as C
import json
import os
import time

import numpy as np

from . import _lib
from .schema import Graph, Reads

_ERRORS = {1: ValueError, 2: KeyError, 3: ZeroDivisionError, 4: IndexError, 5: OSError, 6: ValueError}


def is_gamp(path):
    """GAMP (what `vg mpmap` writes, StrainFLAIR.sh:582) or its `vg view -j -K` JSON rendering (:592-602)?  JSON lines
    start with '{'; a GAMP stream starts with a gzip / BGZF header or with a small binary varint."""
    with open(path, "rb") as fh:
        head = fh.read(64).lstrip()
    return bool(head) and not head.startswith(b"{")


def decode_alignments(path, graph: Graph, thr=0.95, **kw) -> Reads:
    """The alignment file of a query, whichever of the two forms it is in."""
    return decode_gamp(path, graph, thr, **kw) if is_gamp(path) else decode_json(path, graph, thr, **kw)


def decode_gamp(gamp_path, graph: Graph, thr=0.95, threads=None, block_records=0, want_cov=False) -> Reads:
    """Decode the GAMP file of `vg mpmap` directly (no JSON hop): same Reads as ``decode_json`` of its JSON rendering.
    The container and message layout are restated from vg's published format (parity unpinned: no vg at hand)."""
    return decode_json(gamp_path, graph, thr, threads, block_records, want_cov, _gamp=True)


def decode_json(json_path, graph: Graph, thr=0.95, threads=None, block_bytes=0, want_cov=False, _gamp=False) -> Reads:
    """Decode a whole alignment file.  Raises the exception class the reference would raise on bad input
    (unknown node id -> KeyError, zero-length read -> ZeroDivisionError, malformed line -> ValueError)."""
    from . import cache
    t0 = time.perf_counter()
    hit = cache.load_reads_entry(json_path, graph, thr, want_cov)     # only when SFB200_CACHE_DIR is set
    if hit is not None:
        hit.decode_seconds, hit.from_cache = time.perf_counter() - t0, True
        return hit
    lib = _lib.load()
    threads = threads or len(os.sched_getaffinity(0))
    ids = np.ascontiguousarray(graph.node_ids, np.int64)
    lens = np.ascontiguousarray(graph.node_len, np.int32)
    t0 = time.perf_counter()
    entry = lib.sfb_decode_gamp if _gamp else lib.sfb_decode_json
    h = entry(os.fsencode(json_path), ids.ctypes.data, lens.ctypes.data, len(ids), float(thr), int(threads), int(block_bytes))
    try:
        msg = C.c_char_p()
        kind = lib.sfb_decoded_error(h, C.byref(msg))
        if kind:
            text = (msg.value or b"").decode(errors="replace")
            if kind == 1:
                raise json.JSONDecodeError(text, "", 0)
            raise _ERRORS.get(kind, RuntimeError)(text)
        G, A, R, E, NB = (int(lib.sfb_decoded_size(h, k)) for k in range(5))

        def take(which, n, dt):
            out = np.empty(n, dt)
            if n:
                lib.sfb_decoded_copy(h, which, out.ctypes.data)
            return out

        arrays = dict(
            grp_nmates=take(0, G, np.uint8), grp_aln_off=take(1, 2 * G + 1, np.int64),
            mate_seq_len=take(2, 2 * G, np.int32), aln_walk_off=take(3, A + 1, np.int64),
            aln_read_len=take(4, A, np.int32), aln_bins=take(5, 5 * A, np.uint8).reshape(A, 5),
            walk_node=take(7, R, np.int32), walk_alen=take(8, R, np.int32), exc_cov=take(10, E, np.float64),
        )
        name_off = take(11, 2 * G + 1, np.int64)
        blob = take(12, NB, np.uint8).tobytes()
        reads = Reads(**arrays, mate_names=None)
        reads.name_off, reads.name_blob = name_off, blob
        reads.aln_stats = take(6, 5 * A, np.float64).reshape(A, 5)
        if want_cov:
            reads.walk_cov = take(9, R, np.float64)
        reads.decode_seconds, reads.from_cache = time.perf_counter() - t0, False
        cache.save_reads_entry(json_path, graph, thr, reads)
        return reads
    finally:
        lib.sfb_decoded_free(h)


def mate_name(reads: Reads, group, mate):
    """Read name of a mate (lazy: names are only needed for the unassigned-read report)."""
    if getattr(reads, "mate_names", None) is not None:
        return reads.mate_names[2 * group + mate]
    k = 2 * group + mate
    return reads.name_blob[int(reads.name_off[k]):int(reads.name_off[k + 1])].decode("utf-8", errors="replace")
