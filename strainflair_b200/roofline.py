"""Algorithmic (compulsory) bytes of each kernel of the abundance pass, from the realised workload.

Definitions ("Roofline arithmetic").  Caches are ignored: every byte a
kernel must read or write to do its job is counted once, with the array element sizes of include/sfb200.h.
The per-launch work statistics come from the ST_* counters the kernels maintain.

    G groups, A retained alignments, R alignment-node records,
    cand     candidate occurrences tested by match        (ST_CAND)
    cmp      path-node compares executed by match          (ST_COMPARES)
    tuples   matches found                                 (ST_TUPLES)
    fill     (code, path id) pairs written to match_buf (alignments with several matches)
    r1       records added in pass 1                       (ST_P1_RECORDS)
    n1       unique assignments (ASSIGN_SAME + ASSIGN_DIFF leaves)
    hist     histogram-updating reads                      (ST_HIST)
    n_def    deferred groups; a_def their alignments (upper bound: all alignments of deferred groups)
    app      pass-2 (read, path) redistributions           (ST_P2_APPLIED)
    r2       records added in pass 2                       (ST_P2_RECORDS)
"""


def kernel_bytes(n_groups, n_alns, n_records, n_paths, n_slots, C, match_fill, n_deferred, n_strains, tile_groups=0,
                 tile_alns=0):
    G, A, R = n_groups, n_alns, n_records
    cand, cmp_, tuples = C["ST_CAND"], C["ST_COMPARES"], C["ST_TUPLES"]
    r1, app, r2, hist = C["ST_P1_RECORDS"], C["ST_P2_APPLIED"], C["ST_P2_RECORDS"], C["ST_HIST"]
    n1 = C["SAME_CP"] + C["DIFF_CP"]
    out = {}
    # match: walk offsets (8) + walk nodes (4/record) + occurrence ranges (2 ints per orientation)
    #        + (slot, successor) pairs (8/candidate; the first compare is served by the pair)
    #        + further compared path nodes (4/compare) + slot->path (8/match) + result pair (8) + multi lists (8/pair)
    out["match"] = 8 * A + 4 * R + 16 * A + 8 * cand + 4 * max(0, cmp_ - cand) + 8 * tuples + 8 * A + 8 * match_fill
    # classify: per group nmates(1) + 2 alignment offsets(16) + 2 seq_len(8) + code(1); per alignment
    #           aln_match(8) + aln_assign(4); multi lists (8/pair); per unique assignment uniq_reads RMW(16)
    #           + uniq_norm RMW(16) + path_seq_len(8) + 5 bins; per histogram read 5 RMW(16); deferred list (4)
    out["classify"] = 26 * G + 12 * A + 8 * match_fill + 45 * n1 + 80 * hist + 4 * n_deferred
    # pass-1 scatter: walk offsets(8) + aln_assign(4) per alignment; per added record the packed
    #                 aligned/node-length word(4) + fp64 RMW(16)
    out["pass1_scatter"] = 12 * A + 20 * r1
    # pass 2: deferred id(4) + group header(26) per deferred group; per redistribution: match pair(8) +
    #         uniq_reads gather(8) + 3 RMW (mult_reads, mult_norm 16 each, mult_hits 8) + path_seq_len(8)
    #         + work item write+read(32) + walk offsets(16); per record alen(4) + node length(4) + RMW(16)
    out["pass2_resolve"] = 30 * n_deferred + 96 * app          # incl. the 16-byte work item write
    out["pass2_scatter"] = 16 * app + 16 * app + 20 * r2          # work item read + walk offsets; packed record(4) + RMW(16)
    out["pass2"] = out["pass2_resolve"] + out["pass2_scatter"]
    # path reduce: two fp64 slot arrays in, 8 fp64 + 3 counters per path
    out["path_reduce"] = 16 * n_slots + (8 * 8 + 4 * 8 + 8) * n_paths
    # Tile kernels: the same algorithmic work (the per-unit figures above are those of the algorithm, not of one
    # implementation), attributed by the share of alignments / groups that sit in tiles: tile_pass1 = match + classify +
    # pass-1 scatter of those reads, tile_pass2 = pass-2 resolve + scatter; the global-memory kernels keep the rest.
    ft = tile_alns / A if A else 0.0
    out["tile_pass1"] = ft * (out["match"] + out["classify"] + out["pass1_scatter"])
    out["tile_pass2"] = ft * out["pass2"]
    for k in ("match", "classify", "pass1_scatter", "pass2_resolve", "pass2_scatter", "pass2"):
        out[k] = (1.0 - ft) * out[k]
    out["total"] = sum(v for k, v in out.items() if k != "pass2")
    # What the tile design has to move through HBM at the least (no index gathers, no per-record accumulator RMWs):
    # every input array once in pass 1 (group header 25 B, alignment 8 + 5 B, record 8 B), grp_code 1 B + aln_assign 4 B
    # out, the deferred groups' inputs once more in pass 2 + grp_code 1 B in / grp_code2 1 B out, the tile images
    # (12 B per slot per work item ~ per tile) and the flushes (16 B per touched slot and pass), the path reduction.
    fd = n_deferred / G if G else 0.0
    inputs = 25 * G + 13 * A + 8 * R
    out["stream_total"] = inputs + G + 4 * A + fd * inputs + G + fd * G + 2 * (12 + 16) * n_slots + out["path_reduce"]
    # SURVEY.md section 8d's closed form, with the realised c (compares/record), m (slot RMWs/record) and
    # candidates+matches per alignment: R(8 + 4c + 16m) + A(24 + 8(cand+matches)/A) + 16G + 20T + 8(S+11)P
    out["survey_total"] = (8 * R + 4 * cmp_ + 16 * (r1 + r2)) + (24 * A + 8 * (cand + tuples)) + 16 * G \
        + 20 * n_slots + 8 * (n_strains + 11) * n_paths
    return out
