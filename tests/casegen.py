"""Seeded generator of small reference-readable cases: GFA + vg-mpmap JSON + cluster pickle.

Used by the live-reference comparison, by ``tests/golden/make_golden.py`` and by
the decoder tests.  It aims at *branch coverage* of the reference's decision
tree, not at realism: tiny bubble-chain clusters,
paths shared between strains, repeated nodes, reversed walks, tied alignments,
multi-subpath DAGs, unmapped mates, recombinant walks.
"""
import json
import pickle
import random

import numpy as np


def make_graph(rng, n_clusters=4, n_strains=4, first_node=1):
    """Returns dict(nodes={id: len}, plines=[(name, [(id, sign)])], clusters={...}, strains=[...])."""
    strains = [f"NC_{100000 + 7 * i}.{1 + i % 3}" for i in range(n_strains)]
    nodes = {}
    plines = []
    clusters = {}
    nid = first_node
    gene_no = 0
    for c in range(n_clusters):
        chain = []
        for _ in range(rng.randint(2, 7)):
            alleles = [nid]
            nodes[nid] = rng.randint(1, 40)
            nid += 1
            if rng.random() < 0.45:
                alleles.append(nid)
                nodes[nid] = rng.randint(1, 40)
                nid += 1
            chain.append(alleles)
        haps = []
        for _ in range(rng.randint(1, 4)):
            hap = [rng.choice(al) for al in chain]
            if rng.random() < 0.15 and len(hap) > 2:      # repeated node inside a path
                hap.insert(rng.randrange(1, len(hap)), hap[0])
            haps.append(hap)
        genes = []
        for hap in haps:
            for s in rng.sample(strains, rng.randint(1, min(3, n_strains))):
                for _ in range(2 if rng.random() < 0.1 else 1):
                    gene_no += 1
                    name = f"gi|{1000 + gene_no}|ref|{s}|_{gene_no}" if rng.random() < 0.5 else f"{s}_{gene_no}"
                    if rng.random() < 0.3:
                        spec = [(n, "-") for n in reversed(hap)]
                    else:
                        spec = [(n, "+") for n in hap]
                    plines.append((name, spec))
                    genes.append(name)
        if rng.random() < 0.2:
            genes.append(f"ghost_{c}_1")                 # gene absent from the graph (:314)
        clusters[f"Cluster_{c}"] = {"genes_list": genes, "len_rep": 100}
    used = sorted({n for _, spec in plines for n, _ in spec})   # nodes on >= 1 path (Q10 otherwise)
    return dict(nodes=nodes, plines=plines, clusters=clusters, strains=strains, used=used)


def write_gfa(graph, path, rng=None):
    with open(path, "w") as fh:
        fh.write("H\tVN:Z:1.0\n")
        for n, ln in graph["nodes"].items():
            seq = "".join((rng or random).choice("ACGT") for _ in range(ln))
            fh.write(f"S\t{n}\t{seq}\n")
        for name, spec in graph["plines"]:
            fh.write(f"P\t{name}\t" + ",".join(f"{n}{s}" for n, s in spec) + "\t" +
                     ",".join("1M" for _ in spec) + "\n")
        fh.write("L\t1\t+\t2\t+\t0M\n")


def write_pickle(graph, path):
    with open(path, "wb") as fh:
        pickle.dump(graph["clusters"], fh)


def _mapping(rng, node, node_len, first, last, perfect):
    """One mapping on ``node``: returns (mapping dict, to_length)."""
    if first and last:
        span = rng.randint(1, node_len)
    elif first or last:
        span = rng.randint(1, node_len) if rng.random() < 0.8 else node_len
    else:
        span = node_len
    edits = []
    to_total = 0
    left = span
    while left > 0:
        take = left if perfect or rng.random() < 0.7 else rng.randint(1, left)
        kind = "m" if perfect else rng.choices("msid", weights=[80, 8, 6, 6])[0]
        if kind == "m":
            edits.append({"from_length": take, "to_length": take})
            to_total += take
        elif kind == "s":
            edits.append({"from_length": take, "to_length": take, "sequence": "A" * take})
            to_total += take
        elif kind == "i":
            k = rng.randint(1, 3)
            edits.append({"to_length": k, "sequence": "C" * k})
            to_total += k
            continue
        else:
            take = min(take, 2)
            edits.append({"from_length": take})
        left -= take
    if not edits:
        edits.append({"from_length": 1, "to_length": 1})
        to_total = 1
    pos = {"node_id": str(node) if rng.random() < 0.7 else node}
    if rng.random() < 0.5:
        pos["offset"] = "3"
    return {"position": pos, "edit": edits, "rank": "1"}, to_total


def _alignment_line(rng, graph, name, pair, walk, perfect=True, score_drop=0, dag=False):
    """One JSON line whose best route spells ``walk``.  Error ratios are kept inside
    [0, 1] (outside, the reference dies with a KeyError, Q7 -- tested separately)."""
    if not perfect:
        return _build_line(rng, graph, name, pair, walk, False, score_drop, dag, safe=True)
    return _build_line(rng, graph, name, pair, walk, True, score_drop, dag)


def _build_line(rng, graph, name, pair, walk, perfect, score_drop, dag, safe=False):
    maps = []
    tot = 0
    for i, n in enumerate(walk):
        m, t = _mapping(rng, n, graph["nodes"][n], i == 0, i == len(walk) - 1, perfect)
        # occasionally split a node over two mappings (merge rule, :576-577)
        if rng.random() < 0.12 and len(m["edit"]) == 1 and m["edit"][0].get("from_length", 0) >= 2 \
                and m["edit"][0].get("to_length", 0) == m["edit"][0].get("from_length", 0) \
                and "sequence" not in m["edit"][0]:
            f = m["edit"][0]["from_length"]
            a = rng.randint(1, f - 1)
            maps.append(({"position": m["position"], "edit": [{"from_length": a, "to_length": a}]}, a))
            maps.append(({"position": dict(m["position"]), "edit": [{"from_length": f - a, "to_length": f - a}]}, f - a))
        else:
            maps.append((m, t))
        tot += t
    if safe:
        dele = sum(int(e.get("from_length", 0)) for mm, _ in maps for e in mm["edit"] if int(e.get("to_length", 0)) == 0)
        match = sum(int(e.get("to_length", 0)) for mm, _ in maps for e in mm["edit"]
                    if int(e.get("to_length", 0)) == int(e.get("from_length", 0)) and "sequence" not in e)
        if tot <= 0 or dele > match or match <= 0:
            return _build_line(rng, graph, name, pair, walk, True, score_drop, dag)
    # cut the mapping list into 1..3 consecutive subpaths
    n_sub = min(len(maps), rng.choice([1, 1, 2, 3]))
    cuts = sorted(rng.sample(range(1, len(maps)), n_sub - 1)) if n_sub > 1 else []
    parts = [maps[a:b] for a, b in zip([0] + cuts, cuts + [len(maps)])]
    subs = []
    budget = max(0, tot - score_drop)
    for k, part in enumerate(parts):
        t = sum(x[1] for x in part)
        sc = min(t, budget) if k < len(parts) - 1 else budget
        budget -= sc
        sp = {"path": {"mapping": [x[0] for x in part]}}
        if sc != 0 or rng.random() < 0.3:
            sp["score"] = sc
        if k < len(parts) - 1:
            sp["next"] = [k + 1]
        subs.append(sp)
    starts = [0]
    if dag and len(subs) >= 1:
        # add an alternative first subpath (worse, equal or better score) that joins subpath 1 (or ends)
        alt_node = rng.choice(graph["used"])
        m, t = _mapping(rng, alt_node, graph["nodes"][alt_node], True, len(subs) == 1, True)
        first_to = sum(int(e.get("to_length", 0)) for mm in subs[0]["path"]["mapping"] for e in mm["edit"])
        m["edit"] = [{"from_length": first_to, "to_length": first_to}] if first_to else m["edit"]
        alt_sc = max(0, min(first_to, subs[0].get("score", 0) + rng.choice([-2, -1, 0, 0, 1])))
        alt = {"path": {"mapping": [m]}, "score": alt_sc}
        if "next" in subs[0]:
            alt["next"] = list(subs[0]["next"])
        subs.append(alt)
        starts = [0, len(subs) - 1] if rng.random() < 0.5 else [len(subs) - 1, 0]
    seq_len = tot + (rng.randint(1, 5) if rng.random() < 0.1 else 0)
    seq = "".join(rng.choice("ACGT") for _ in range(max(1, seq_len)))
    line = {"name": name, "sequence": seq, "start": starts, "subpath": subs, "mapping_quality": 60}
    if pair:
        line["paired_read_name"] = pair
    return line


def _pick_walk(rng, graph, recombinant=False):
    name, spec = rng.choice(graph["plines"])
    ids = [n for n, _ in spec]
    L = rng.randint(1, min(4, len(ids)))
    s = rng.randrange(0, len(ids) - L + 1)
    walk = ids[s:s + L]
    if recombinant:
        walk = list(walk)
        walk[rng.randrange(len(walk))] = rng.choice(graph["used"])
        if rng.random() < 0.5:
            walk.append(rng.choice(graph["used"]))
    if rng.random() < 0.5:
        walk = walk[::-1]
    return walk, name


def make_reads(rng, graph, n_groups=60, paired_frac=0.75):
    lines = []
    for r in range(n_groups):
        paired = rng.random() < paired_frac
        n1 = f"read{r}" + ("/1" if paired and rng.random() < 0.5 else "")
        n2 = (n1[:-2] + "/2" if n1.endswith("/1") else n1) if paired else ""
        base_walk, gene = _pick_walk(rng, graph, recombinant=rng.random() < 0.12)
        for mate, (nm, other) in enumerate(((n1, n2), (n2, n1)) if paired else ((n1, ""),)):
            roll = rng.random()
            if roll < 0.08:                               # unmapped mate: no subpath
                seq = "".join(rng.choice("ACGT") for _ in range(30))
                ln = {"name": nm, "sequence": seq}
                if other:
                    ln["paired_read_name"] = other
                lines.append(ln)
                continue
            if mate == 0 or rng.random() < 0.7:
                walk = base_walk if mate == 0 else _sibling(rng, graph, base_walk)
            else:
                walk, _ = _pick_walk(rng, graph, recombinant=rng.random() < 0.1)
            drop = rng.choice([0, 0, 0, 0, 1, 2, 12]) if roll > 0.2 else 0
            first = _alignment_line(rng, graph, nm, other, walk, perfect=rng.random() < 0.7,
                                    score_drop=drop, dag=rng.random() < 0.2)
            lines.append(first)
            extra = rng.choices([0, 1, 2], weights=[70, 22, 8])[0]
            for _ in range(extra):                        # more lines for the same mate
                w2 = _sibling(rng, graph, walk) if rng.random() < 0.7 else _pick_walk(rng, graph)[0]
                more = _alignment_line(rng, graph, nm, other, w2, perfect=True,
                                       score_drop=rng.choice([0, 0, 1, 3]), dag=False)
                # same mate <=> same sequence string (:717,728); scores are normalised by the
                # route's own read_len, so equal-looking lines may still differ in score
                more["sequence"] = first["sequence"]
                lines.append(more)
    return lines


def _sibling(rng, graph, walk):
    """A walk of similar position on some path sharing a node with ``walk`` (or the same walk)."""
    cands = [spec for _, spec in graph["plines"] if any(n == walk[0] or n == walk[-1] for n, _ in spec)]
    if not cands or rng.random() < 0.4:
        return list(walk)
    ids = [n for n, _ in rng.choice(cands)]
    anchor = walk[0] if walk[0] in ids else walk[-1]
    i = ids.index(anchor)
    L = rng.randint(1, min(4, len(ids)))
    s = max(0, min(i - rng.randint(0, L - 1), len(ids) - L))
    w = ids[s:s + L]
    return w[::-1] if rng.random() < 0.5 else w


def write_json(lines, path):
    with open(path, "w") as fh:
        for ln in lines:
            fh.write(json.dumps(ln) + "\n")


def make_case(seed, outdir, n_groups=60, n_clusters=4, n_strains=4):
    """Writes <outdir>/graph.gfa, reads.json, dict_clusters.pickle and returns their paths."""
    import os
    rng = random.Random(seed)
    graph = make_graph(rng, n_clusters=n_clusters, n_strains=n_strains)
    lines = make_reads(rng, graph, n_groups=n_groups)
    os.makedirs(outdir, exist_ok=True)
    gfa = os.path.join(outdir, "graph.gfa")
    js = os.path.join(outdir, "reads.json")
    pk = os.path.join(outdir, "dict_clusters.pickle")
    write_gfa(graph, gfa, rng)
    write_json(lines, js)
    write_pickle(graph, pk)
    return gfa, js, pk


def stress_transfer_layout(graph, reads, rng, variant):
    """Bend a synthetic shard (in place) so that it exercises the exception lists of the transfer layout.
    The walks may stop following the graph: the expansion does not care."""
    R, A = reads.n_records, reads.n_alns
    keep = reads.walk_alen >= 0
    if variant == "wide_steps":
        hit = rng.random(R) < 0.05
        reads.walk_node[hit] = rng.integers(0, graph.n_nodes, int(hit.sum()))          # far-away node ids
        edge = np.zeros(R, bool)
        edge[reads.aln_walk_off[:-1]] = True
        edge[reads.aln_walk_off[1:] - 1] = True
        nl = graph.node_len[reads.walk_node].astype(np.int64)
        al = np.where(edge, np.minimum(reads.walk_alen & 0xFFFF, nl), nl)
        reads.walk_alen[keep] = (al | (nl << 16))[keep].astype(np.int32)
    elif variant == "long_nodes":
        graph.node_len[:] = np.minimum(graph.node_len.astype(np.int64) * 40, 0x7FFF)   # aligned lengths beyond a byte
        edge = np.zeros(R, bool)
        edge[reads.aln_walk_off[:-1]] = True
        edge[reads.aln_walk_off[1:] - 1] = True
        nl = graph.node_len[reads.walk_node].astype(np.int64)
        al = np.where(edge, rng.integers(0, 255, R), nl)
        al = np.where(rng.random(R) < 0.05, rng.integers(255, 600, R), al)
        reads.walk_alen[keep] = (al | (nl << 16))[keep].astype(np.int32)
    elif variant == "ragged_lengths":
        present = reads.mate_seq_len > 0
        reads.mate_seq_len[present] = rng.integers(50, 151, int(present.sum()))
        some = rng.random(A) < 0.3                                                      # and > 255 distinct bin tuples
        reads.aln_bins[some] = rng.integers(0, 101, (int(some.sum()), 5))
        reads.aln_bins[rng.random(A) < 0.2, 2] = 255
