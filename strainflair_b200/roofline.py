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
                 tile_alns=0, step_index=False, bins_groups=0, bins_alns=0, n_nodes=0):
    if step_index:
        return path_set_bytes(n_groups, n_alns, n_records, n_paths, n_slots, C, match_fill, n_deferred, n_strains,
                              bins_groups, bins_alns, n_nodes)
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


def path_set_bytes(G, A, R, P, n_slots, C, match_fill, n_deferred, n_strains, bins_groups, bins_alns, n_nodes=0):
    """The same accounting for the path-set kernels (sfb_match_steps, the group kernels on sets, the shared-memory bins):
    ``steps`` = steps of the walks evaluated (ST_COMPARES), ``bins_groups`` / ``bins_alns`` = read groups / alignments of
    the tile range whose node-level sums go through shared memory."""
    steps, tuples = C["ST_COMPARES"], C["ST_TUPLES"]
    r1, app, r2, hist = C["ST_P1_RECORDS"], C["ST_P2_APPLIED"], C["ST_P2_RECORDS"], C["ST_HIST"]
    n1 = C["SAME_CP"] + C["DIFF_CP"]
    fb = bins_alns / A if A else 0.0            # share of the alignments / groups in the tile range
    fg = bins_groups / G if G else 0.0
    out = {}
    # match: walk offsets (8) + the id of every visited node (4; steps + one per walk) + the two path sets (16) and the match
    #        word (8) per alignment + the step index ONCE (32 bytes per node of the graph: reads in tile order keep the
    #        records of a tile in L1/L2, the 32 bytes gathered per visited node are reported beside, not as HBM bytes)
    out["match"] = 8 * A + 4 * (steps + A) + 24 * A + 8 * match_fill + 32 * n_nodes
    out["match_gathered_index_bytes"] = 32 * (steps + A)
    # classify: group header (26) + per alignment match word (8), sets (16), aln_assign (4); per unique assignment the slot
    #           look-up (walk offset 8, node 4, node_mask 8, occ_off 8, occurrence 8) + uniq_reads / uniq_norm RMW (32)
    #           + path_seq_len (8) + 5 bins; per histogram read 5 RMW (16); deferred list (4)
    out["classify"] = 26 * G + 28 * A + 8 * match_fill + 81 * n1 + 80 * hist + 4 * (1 - fg) * n_deferred
    # pass-1 node loops: aln_assign (4) + walk offsets (8) per alignment; per added record the packed word (4); the sums:
    #   tile range   shared-memory bins, flushed once per work item: two 8-byte RMWs per slot of the tile (32 per slot)
    #   the rest     an fp64 / two int64 RMWs per record at the L2 (16)
    out["tile_scatter1"] = fb * (12 * A + 4 * r1) + 32 * n_slots * (1 if fb else 0)
    out["pass1_scatter"] = (1 - fb) * (12 * A + 20 * r1)
    # pass 2 in the tile range: grp_code per group (1) + header (26), grp_code2 (1), match words and sets of the deferred
    #   groups' alignments (~24 each, 2.2 per group); per share: uniq_reads (8), path_seq_len (8), strain word (8), walk
    #   offsets (16), node (4), node_mask (8), occ_off (8), occurrence (8); per record the packed word (4); flush as above
    #   (+ 40 per path)
    d_in = fg * n_deferred
    out["tile_share"] = fg * G + 27 * d_in + 24 * 2.2 * d_in + fg * (68 * app + 4 * r2) + (32 * n_slots + 40 * P) * (1 if fg else 0)
    # pass 2 outside it: as before, with 16-byte work items that name their path
    out["pass2_resolve"] = (1 - fg) * (30 * n_deferred + 96 * app)
    out["pass2_scatter"] = (1 - fg) * (32 * app + 36 * app + 20 * r2)
    out["pass2"] = out["tile_share"] + out["pass2_resolve"] + out["pass2_scatter"]
    out["path_reduce"] = 16 * n_slots + (8 * 8 + 4 * 8 + 8) * P
    out["tile_pass1"] = out["tile_pass2"] = 0.0
    out["total"] = sum(v for k, v in out.items() if k not in ("pass2", "match_gathered_index_bytes"))
    fd = n_deferred / G if G else 0.0
    inputs = 25 * G + 13 * A + 8 * R
    out["stream_total"] = inputs + G + 4 * A + fd * inputs + G + fd * G + 2 * (12 + 16) * n_slots + out["path_reduce"]
    # SURVEY.md section 8d's closed form R(8 + 4c + 16m) + A(24 + 8 (cand + matches)/A) + 16G + 20T + 8(S+11)P with the
    # realised work of THIS algorithm: c = steps per record (the step index tests no candidate occurrences: cand = 0),
    # m = node-level additions per record (16 bytes each, wherever the accumulator lives)
    out["survey_total"] = (8 * R + 4 * steps + 16 * (r1 + r2)) + (24 * A + 8 * tuples) + 16 * G \
        + 20 * n_slots + 8 * (n_strains + 11) * P
    return out
