"""Seeded synthetic pangenomes and alignments, generated directly as arrays.

Graph family: ``n_clusters`` gene clusters, each a bubble chain of ``chain_len`` positions (a position is
one node, or two alleles with probability ``bubble_p``); each cluster carries a few haplotype paths that
differ from a founder at few bubbles, identical haplotypes are merged (as the reference merges P lines
with identical content, json2csv.py:260-266) and each path is carried by 1-3 of ``n_strains`` strains.

Reads: pairs of ``read_len`` bp placed on a haplotype path by coordinate, so that a walk is a contiguous
run of path nodes with partial coverage on its two end nodes; 50 % reversed; a share of mates below the
score threshold (no retained alignment), with a second tied alignment on a sibling haplotype, with one
node swapped to the other allele (recombinant), and a share of single-end groups.
"""
import numpy as np

from .schema import Graph, Reads

CONFIGS = {
    # name: (graph kwargs, reads kwargs) -- C3/C4/C5 of BASELINE.json, plus small test sizes
    "tiny": (dict(n_clusters=40, chain_len=12, hap_lo=1, hap_hi=5, n_strains=6, bubble_p=0.4, flip_p=0.3),
             dict(n_pairs=2000, read_len=60)),
    "small": (dict(n_clusters=400, chain_len=40, hap_lo=1, hap_hi=8, n_strains=16, bubble_p=0.3, flip_p=0.15),
              dict(n_pairs=50_000, read_len=150)),
    # C3: 2 M nodes, ~50 k paths, ~7.4 M path positions; independent alleles on every second position give
    # ~2 matching paths per alignment and ~30 % of the groups deferred to pass 2
    "c3": (dict(n_clusters=8_900, chain_len=150, hap_lo=2, hap_hi=9, n_strains=64, bubble_p=0.5, flip_p=0.5),
           dict(n_pairs=25_000_000, read_len=150)),
    # C4: 16 near-identical haplotypes per cluster: ~8+ candidate paths per read, most groups deferred
    "c4": (dict(n_clusters=3_125, chain_len=150, hap_lo=16, hap_hi=16, n_strains=64, bubble_p=0.33, flip_p=0.02),
           dict(n_pairs=25_000_000, read_len=150)),
    # C5: 20 M nodes, 1e9 alignment-node records over all ranks (sharded by read group)
    "c5": (dict(n_clusters=89_000, chain_len=150, hap_lo=2, hap_hi=9, n_strains=64, bubble_p=0.5, flip_p=0.5),
           dict(n_pairs=83_000_000, read_len=150)),
}
SEED0 = 20261019


def make_graph(seed, n_clusters, chain_len=150, hap_lo=1, hap_hi=8, n_strains=64, bubble_p=0.3, flip_p=0.1,
               max_node_len=63):
    rng = np.random.default_rng(seed)
    C, K, S = n_clusters, chain_len, n_strains
    bubble = rng.random((C, K)) < bubble_p
    width = 1 + bubble.astype(np.int64)
    base = np.cumsum(width.reshape(-1)) - width.reshape(-1)          # node index of allele 0 at (c, k)
    base = base.reshape(C, K)
    n_nodes = int(width.sum())
    node_len = rng.integers(1, max_node_len + 1, n_nodes).astype(np.int32)
    n_hap = rng.integers(hap_lo, hap_hi + 1, C)
    owner = np.repeat(np.arange(C), n_hap)                            # cluster of each raw haplotype
    founder = (rng.random((C, K)) < 0.5) & bubble
    flips = (rng.random((len(owner), K)) < flip_p) & bubble[owner]
    allele = founder[owner] ^ flips
    # merge identical haplotypes inside a cluster, keep first-seen order
    key = np.concatenate([owner[:, None].astype(np.uint32).view(np.uint8).reshape(len(owner), 4),
                          np.packbits(allele, axis=1)], axis=1)
    _, first = np.unique(key, axis=0, return_index=True)
    first.sort()
    owner, allele = owner[first], allele[first]
    P = len(owner)
    path_nodes = (base[owner] + allele).astype(np.int32).reshape(-1)
    path_off = np.arange(P + 1, dtype=np.int64) * K
    seq_len = node_len[path_nodes].reshape(P, K).sum(axis=1).astype(np.int64)
    # strains per path: 1..3 distinct ids, copy count 1 (2 with p = 0.02)
    n_s = 1 + (rng.random(P) < 0.3) + (rng.random(P) < 0.1)
    n_s = np.minimum(n_s, S)
    s0 = rng.integers(0, S, P)
    d1 = rng.integers(1, max(2, S // 2), P)
    d2 = rng.integers(1, max(2, S // 2), P)
    cand = np.stack([s0, (s0 + d1) % S, (s0 + d1 + d2) % S], axis=1)
    if S < 4:                                                         # tiny strain sets: dedupe by construction
        n_s = np.minimum(n_s, 1 + (S > 1))
        cand[:, 1] = (s0 + 1) % S
    take = np.arange(3)[None, :] < n_s[:, None]
    pstr_id = cand[take].astype(np.int32)
    pstr_off = np.concatenate(([0], np.cumsum(n_s))).astype(np.int64)
    pstr_cnt = (1 + (rng.random(len(pstr_id)) < 0.02)).astype(np.int32)
    # clusters: one gene per (path, strain, copy) -> the path id repeats in the cluster list (json2csv.py:317)
    genes_per_path = np.add.reduceat(pstr_cnt, pstr_off[:-1]) if P else np.zeros(0, np.int64)
    clu_paths = np.repeat(np.arange(P, dtype=np.int32), genes_per_path)
    per_cluster = np.bincount(owner, weights=genes_per_path, minlength=C).astype(np.int64)
    clu_off = np.concatenate(([0], np.cumsum(per_cluster))).astype(np.int64)
    # header = strains in order of first appearance as the FIRST strain of a path (species_names, :268)
    firsts = pstr_id[pstr_off[:-1]]
    _, idx = np.unique(firsts, return_index=True)
    header = [int(s) for s in firsts[np.sort(idx)]]
    g = Graph(
        node_ids=np.arange(1, n_nodes + 1, dtype=np.int64), node_len=node_len, path_off=path_off,
        path_nodes=path_nodes, path_seq_len=seq_len, pstr_off=pstr_off, pstr_id=pstr_id, pstr_cnt=pstr_cnt,
        path_cluster=owner.astype(np.int32), clu_off=clu_off, clu_paths=clu_paths,
        strain_names=[f"NC_{100000 + i}.1" for i in range(S)], header_strains=header, path_names=None,
        node_index=None,
    )
    g.chain_len = K
    g.base = base
    g.allele = allele
    g.bubble = bubble
    g.on_path = np.zeros(n_nodes, bool)
    g.on_path[path_nodes] = True
    return g


def _python_round_bins(values):
    """Histogram bin (Python round(x, 2) * 100) of each value, through the few distinct values."""
    uniq, inv = np.unique(values, return_inverse=True)
    table = np.array([int(round(round(float(v), 2) * 100)) if 0 <= round(float(v), 2) <= 1 else 255 for v in uniq],
                     np.uint8)
    return table[inv]


def make_reads(graph, seed, n_pairs, read_len=150, p_same_path=0.95, p_low_score=0.03, p_tie=0.10,
               p_recomb=0.01, p_single=0.02, p_exc=0.0, chunk=1_000_000, threads=8):
    """Returns Reads for ``n_pairs`` read groups (a ``p_single`` share of them single-end).  Keep ``read_len`` at or
    below the shortest path's sequence length: a "tied" second alignment placed on a shorter sibling path would be
    clipped to another length, hence another normalised score, and would not be a tie for the reference
    (json2csv.py:729-737) -- the arrays would then describe an input no JSON file decodes to."""
    rng = np.random.default_rng(seed)
    K = graph.chain_len
    P = graph.n_paths
    plen = graph.node_len[graph.path_nodes].reshape(P, K).astype(np.int64)
    cum = np.zeros((P, K + 1), np.int64)
    np.cumsum(plen, axis=1, out=cum[:, 1:])
    seq_len = cum[:, -1]
    BIG = int(seq_len.max()) + 1
    flat_cum = (cum[:, :-1] + np.arange(P)[:, None] * BIG).reshape(-1)   # start coordinate of every path node
    # siblings: paths of the same cluster are contiguous
    clu = graph.path_cluster.astype(np.int64)
    clu_first = np.searchsorted(clu, clu, side="left")
    clu_size = np.searchsorted(clu, clu, side="right") - clu_first
    # every chunk has its own generator (seed, chunk index): the result does not depend on the thread count
    starts = list(range(0, n_pairs, chunk))

    def one(ci):
        crng = np.random.default_rng([seed, ci])
        return _reads_chunk(graph, crng, min(chunk, n_pairs - starts[ci]), read_len, p_same_path, p_low_score,
                            p_tie, p_recomb, p_single, K, P, cum, seq_len, BIG, flat_cum, clu_first, clu_size)

    if len(starts) > 1 and threads > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=threads) as ex:
            parts = list(ex.map(one, range(len(starts))))
    else:
        parts = [one(ci) for ci in range(len(starts))]
    out = _concat(parts)
    if p_exc > 0 and len(out["walk_alen"]):
        # a few records whose coverage is given as an explicit fp64 (merged-node case, json2csv.py:576-577)
        n_exc = max(1, int(len(out["walk_alen"]) * p_exc))
        pick = rng.choice(len(out["walk_alen"]), n_exc, replace=False)
        out["exc_cov"] = rng.random(n_exc) * 1.5
        out["walk_alen"][pick] = -(np.arange(n_exc, dtype=np.int32) + 1)
    return Reads(**out, mate_names=None)


def _reads_chunk(graph, rng, G, read_len, p_same, p_low, p_tie, p_recomb, p_single, K, P, cum, seq_len, BIG,
                 flat_cum, clu_first, clu_size):
    nm = np.where(rng.random(G) < p_single, 1, 2).astype(np.uint8)
    # mate table: 2 per group, second masked out for singles
    p1 = rng.integers(0, P, G)
    same = rng.random(G) < p_same
    p2 = np.where(same, p1, rng.integers(0, P, G))
    path = np.stack([p1, p2], axis=1).reshape(-1)
    room = np.maximum(seq_len[path] - read_len, 0)
    x = (rng.random(2 * G) * (room + 1)).astype(np.int64)
    x = np.minimum(x, room)
    # mate 2 near mate 1 when on the same path (insert of 0..300 bp)
    x1 = x[0::2]
    near = np.minimum(x1 + rng.integers(0, 300, G), room[1::2])
    x[1::2] = np.where(same, near, x[1::2])
    present = np.ones(2 * G, bool)
    present[1::2] = nm == 2
    low = rng.random(2 * G) < p_low
    n_aln = np.where(present & ~low, 1 + (rng.random(2 * G) < p_tie), 0).astype(np.int64)
    # alignment table
    mate_of = np.repeat(np.arange(2 * G), n_aln)
    A = len(mate_of)
    is_tie = np.zeros(A, bool)
    starts = np.cumsum(n_aln) - n_aln
    second = starts[n_aln == 2] + 1
    is_tie[second] = True
    ap = path[mate_of]
    # tied alignment: same coordinates on a sibling haplotype of the cluster
    sib = clu_first[ap] + (rng.integers(0, 1 << 30, A) % clu_size[ap])
    ap = np.where(is_tie, sib, ap)
    ax = np.minimum(x[mate_of], np.maximum(seq_len[ap] - read_len, 0))
    rl = np.minimum(read_len, seq_len[ap])
    s = np.searchsorted(flat_cum, ax + ap * BIG, side="right") - 1 - ap * K
    e = np.searchsorted(flat_cum, ax + rl - 1 + ap * BIG, side="right") - 1 - ap * K
    L = (e - s + 1).astype(np.int64)
    woff = np.concatenate(([0], np.cumsum(L)))
    R = int(woff[-1])
    j = np.arange(R, dtype=np.int64) - np.repeat(woff[:-1], L)         # position inside the walk
    a_of = np.repeat(np.arange(A), L)
    k = s[a_of] + j                                                    # chain position
    node = graph.path_nodes[ap[a_of] * K + k].astype(np.int32)
    n_lo = cum[ap[a_of], k]
    n_hi = cum[ap[a_of], k + 1]
    alen = (np.minimum(n_hi, (ax + rl)[a_of]) - np.maximum(n_lo, ax[a_of])).astype(np.int32)
    # recombinant: swap one node to the other allele where the position is a bubble
    rec = rng.random(A) < p_recomb
    jr = (rng.integers(0, 1 << 30, A) % L)
    hit = rec[a_of] & (j == jr[a_of])
    cl = graph.path_cluster[ap[a_of]]
    is_b = graph.bubble[cl, k]
    alt = graph.base[cl, k] + (1 - graph.allele[ap[a_of], k].astype(np.int64))
    alt = np.where(is_b, alt, node)
    swap = hit & is_b & graph.on_path[alt]       # a node no path crosses is an IndexError in the reference (Q10)
    node = np.where(swap, alt, node).astype(np.int32)
    # orientation: reverse the record order of half of the walks
    rev = rng.random(A) < 0.5
    src = np.where(rev[a_of], woff[:-1][a_of] + (L[a_of] - 1 - j), np.arange(R))
    node, alen = node[src], alen[src]
    alen = (alen | (graph.node_len[node].astype(np.int32) << 16)).astype(np.int32)    # aligned bases | node length << 16
    # error ratios -> histogram bins
    n_sub = rng.choice(np.array([0, 0, 0, 0, 1], np.int64), A)
    n_sub[second] = n_sub[second - 1]            # retained alignments of a mate are tied: same score
    rlA = rl.astype(np.float64)
    pen = np.maximum(1, rl // 50)                # score penalty of a substitution; keeps the score above 0.95
    stats = np.stack([(rl - pen * n_sub) / rlA, (rlA - n_sub) / rlA, n_sub / rlA, np.zeros(A), n_sub / rlA], axis=1)
    bins = np.stack([_python_round_bins(stats[:, c]) for c in range(5)], axis=1)
    grp_aln_off = np.concatenate(([0], np.cumsum(n_aln))).astype(np.int64)
    return dict(
        grp_nmates=nm, grp_aln_off=grp_aln_off, mate_seq_len=np.where(present, read_len, 0).astype(np.int32),
        aln_walk_off=woff.astype(np.int64), aln_read_len=rl.astype(np.int32), aln_bins=bins,
        walk_node=node, walk_alen=alen, exc_cov=np.zeros(0, np.float64),
    )


def _concat(parts):
    if len(parts) == 1:
        return parts[0]
    out = {}
    a_base = r_base = 0
    offs_g, offs_a = [np.zeros(1, np.int64)], [np.zeros(1, np.int64)]
    for p in parts:
        offs_g.append(p["grp_aln_off"][1:] + a_base)
        offs_a.append(p["aln_walk_off"][1:] + r_base)
        a_base += int(p["grp_aln_off"][-1])
        r_base += int(p["aln_walk_off"][-1])
    out["grp_aln_off"] = np.concatenate(offs_g)
    out["aln_walk_off"] = np.concatenate(offs_a)
    for k in ("grp_nmates", "mate_seq_len", "aln_read_len", "aln_bins", "walk_node", "walk_alen", "exc_cov"):
        out[k] = np.concatenate([p[k] for p in parts])
    return out


def make_config(name, scale=1.0, seed=None):
    """Graph + reads of a named configuration; ``scale`` shrinks the number of pairs (bounded CPU samples)."""
    gkw, rkw = CONFIGS[name]
    idx = list(CONFIGS).index(name)
    seed = SEED0 + idx if seed is None else seed
    graph = make_graph(seed, **gkw)
    rkw = dict(rkw)
    rkw["n_pairs"] = max(1, int(rkw["n_pairs"] * scale))
    reads = make_reads(graph, seed + 1000, **rkw)
    return graph, reads


def export_files(graph, reads, outdir, prefix="synth"):
    """Writes the configuration as the files the command lines take: <prefix>.gfa, <prefix>.json
    (`vg view -j -K` style, one subpath per alignment) and <prefix>_clusters.pickle.  Decoding the JSON
    with thr = 0.95 gives back ``reads`` array for array (explicit-coverage records excepted)."""
    import json
    import os
    import pickle
    os.makedirs(outdir, exist_ok=True)
    rng = np.random.default_rng(12345)
    names = graph.strain_names
    gfa = os.path.join(outdir, prefix + ".gfa")
    lut = np.frombuffer(b"ACGT", np.uint8)
    with open(gfa, "w") as fh:
        fh.write("H\tVN:Z:1.0\n")
        seq = lut[rng.integers(0, 4, int(graph.node_len.sum()))].tobytes().decode()
        off = np.concatenate(([0], np.cumsum(graph.node_len)))
        for k in range(graph.n_nodes):
            fh.write(f"S\t{int(graph.node_ids[k])}\t{seq[off[k]:off[k + 1]]}\n")
        genes = [[] for _ in range(graph.n_clusters)]
        gene_no = 0
        for p in range(graph.n_paths):
            nodes = ",".join(f"{int(graph.node_ids[n])}+" for n in graph.path_nodes[graph.path_off[p]:graph.path_off[p + 1]])
            for e in range(int(graph.pstr_off[p]), int(graph.pstr_off[p + 1])):
                for _ in range(int(graph.pstr_cnt[e])):
                    gene_no += 1
                    name = f"{names[int(graph.pstr_id[e])]}_{gene_no}"
                    fh.write(f"P\t{name}\t{nodes}\t*\n")
                    if graph.path_cluster[p] >= 0:
                        genes[int(graph.path_cluster[p])].append(name)
    pk = os.path.join(outdir, prefix + "_clusters.pickle")
    with open(pk, "wb") as fh:
        pickle.dump({f"Cluster_{c}": {"genes_list": g, "len_rep": 0} for c, g in enumerate(genes)}, fh)
    js = os.path.join(outdir, prefix + ".json")
    ids = graph.node_ids
    with open(js, "w") as fh:
        for g in range(reads.n_groups):
            nm = int(reads.grp_nmates[g])
            for m in range(nm):
                name = f"r{g}/{m + 1}" if nm == 2 else f"r{g}"
                pair = f"r{g}/{2 - m}" if nm == 2 else None
                rl = int(reads.mate_seq_len[2 * g + m])
                seq = "AC"[m] * rl
                a0, a1 = int(reads.grp_aln_off[2 * g + m]), int(reads.grp_aln_off[2 * g + m + 1])
                lines = []
                for a in range(a0, a1):
                    w0, w1 = int(reads.aln_walk_off[a]), int(reads.aln_walk_off[a + 1])
                    sub = int(reads.aln_bins[a, 0]) < 100          # one substitution (see _reads_chunk)
                    maps = []
                    for r_ in range(w0, w1):
                        al = int(reads.walk_alen[r_]) & 0xFFFF
                        edits = [{"from_length": al, "to_length": al}]
                        if sub and r_ == w0 and al >= 1:
                            edits = ([{"from_length": al - 1, "to_length": al - 1}] if al > 1 else []) + \
                                    [{"from_length": 1, "to_length": 1, "sequence": "T"}]
                        maps.append({"position": {"node_id": str(int(ids[reads.walk_node[r_]]))}, "edit": edits})
                    lines.append({"name": name, "sequence": seq, "start": [0], "mapping_quality": 60,
                                  "subpath": [{"path": {"mapping": maps}, "score": rl - max(1, rl // 50) * int(sub)}]})
                if not lines:                                     # present, but below the score threshold
                    lines.append({"name": name, "sequence": seq, "start": [0], "subpath": [
                        {"path": {"mapping": [{"position": {"node_id": str(int(ids[0]))},
                                               "edit": [{"from_length": rl, "to_length": rl}]}]}, "score": rl // 2}]})
                for ln in lines:
                    if pair:
                        ln["paired_read_name"] = pair
                    fh.write(json.dumps(ln) + "\n")
    return gfa, js, pk
