"""GPU parity: the CUDA path (through the C ABI) against the oracle on the same seeded inputs."""
import glob
import json
import os

import numpy as np
import pytest
import torch

import casegen
import compare
from oracle import abundance as oa
from oracle import decode as od
from oracle import graph as og
from oracle import report as orp
from strainflair_b200 import engine, synth
from strainflair_b200.schema import Graph, Reads

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


# the kernel families (global-memory kernels on tile-ordered reads = the default, shared-memory tile kernels,
# global-memory kernels on reads in file order) and the two accumulation modes
MODES = [dict(), dict(deterministic=False), dict(tile_kernels=True), dict(tile_kernels=True, deterministic=False),
         dict(tiles=False), dict(tiles=False, deterministic=False), dict(step_index=False), dict(tile_bins=False),
         dict(tile_bins=False, deterministic=False)]
MODE_IDS = ["default", "default-plain", "tilekernels-det", "tilekernels-plain", "fileorder-det", "fileorder-plain",
            "occurrence-index-match", "pathsets-l2sums-det", "pathsets-l2sums-plain"]
all_modes = pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)


def need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def slot_to_plain(res, code):
    """sentinel-layout match code -> plain slot (path_off[pid] + pos), like the oracle's aln_assign"""
    g = res.g
    slot = code & 0x3FFFFFFF
    starts = g["path_off"][:-1] + np.arange(len(g["path_off"]) - 1)
    pid = np.searchsorted(starts, slot, side="right") - 1
    return slot - pid


def check_against_oracle(graph, reads, res, orc, rel=compare.REL):
    for k in oa.COUNTERS:
        assert res.C[k] == orc.C[k], (k, res.C[k], orc.C[k])
    assert np.array_equal(res.grp_code, orc.grp_code)
    assert np.array_equal(res.grp_code2, orc.grp_code2)
    assert np.array_equal(res.uniq_reads, np.asarray(orc.uniq_reads, np.int64))
    assert np.array_equal(res.hist, orc.hist)
    assert res.n_deferred == len(orc.deferred)
    # unique assignments: same alignment, same start slot
    mine = np.where(res.aln_assign >= 0, slot_to_plain(res, res.aln_assign.astype(np.int64)), -1)
    assert np.array_equal(mine, orc.aln_assign)
    # recorded unassigned reads
    book = sorted((g, m) for g, m, _ in orc.book)
    got = []
    for g_ in np.flatnonzero(((res.grp_code & 15) == oa.UNASSIGNED_BOOK) | ((res.grp_code >> 4) == oa.UNASSIGNED_BOOK)):
        for m in range(2):
            if (int(res.grp_code[g_]) >> (4 * m)) & 15 == oa.UNASSIGNED_BOOK:
                got.append((int(g_), m))
    assert got == book
    for g_, m, c in orc.book:
        a = int(reads.grp_aln_off[2 * g_ + m])
        assert -(int(res.aln_assign[a]) + 2) == c
    for name in ("uniq_norm", "mult_reads", "mult_norm", "uniq_abund", "mult_abund"):
        a = np.asarray(getattr(res, name), np.float64)
        b = np.asarray(getattr(orc, name), np.float64)
        assert a.shape == b.shape, name
        err = np.abs(a - b)
        tol = rel * np.maximum(np.abs(a), np.abs(b))
        bad = np.flatnonzero(err > tol)
        assert bad.size == 0, (name, bad[:5], a[bad[:5]], b[bad[:5]])
    # zero pattern must be identical (decides the nz means and ratio_covered_nodes)
    assert np.array_equal(res.uniq_abund > 0, orc.uniq_abund > 0)
    assert np.array_equal(res.mult_abund > 0, orc.mult_abund > 0)
    assert np.array_equal(res.mult_hits > 0, np.array([not isinstance(v, int) for v in orc.mult_reads]))
    rows = orp.gene_rows(orc)
    cols = ("nb_uniq_mapped_normalized", "nb_multimapped", "nb_multimapped_normalized", "mean_abund_uniq",
            "mean_abund_uniq_nz", "mean_abund_multiple", "mean_abund_multiple_nz", "ratio_covered_nodes")
    for p, row in enumerate(rows):
        for k, c in enumerate(cols):
            assert compare.close(res.table[p, k], row[c], rel), (p, c, res.table[p, k], row[c])


@all_modes
@pytest.mark.parametrize("name,scale", [("tiny", 1.0), ("small", 0.2)])
def test_synthetic_matches_oracle(name, scale, mode):
    need_gpu()
    graph, reads = synth.make_config(name, scale=scale)
    res = engine.abundance_pass(graph, reads, **mode)
    orc = oa.AbundanceOracle(graph.as_dict(), reads.as_dict()).run()
    check_against_oracle(graph, reads, res, orc)


@all_modes
@pytest.mark.parametrize("n_hap", [16, 48])
def test_many_candidate_paths_take_the_generic_kernels(n_hap, mode):
    """~10 tuples per alignment (LocalView: > 8 per mate) and ~24, up to 60 with ties (GlobalView: > 32)."""
    need_gpu()
    graph = synth.make_graph(11, n_clusters=6, chain_len=80, hap_lo=n_hap, hap_hi=n_hap, n_strains=9, bubble_p=0.4,
                             flip_p=0.03)
    reads = synth.make_reads(graph, 12, n_pairs=1200, read_len=100, p_tie=0.3)
    res = engine.abundance_pass(graph, reads, **mode)
    orc = oa.AbundanceOracle(graph.as_dict(), reads.as_dict()).run()
    per_aln = res.C["ST_TUPLES"] / max(1, reads.n_alns)
    assert per_aln > (9 if n_hap == 16 else 20), per_aln
    assert orc.C["ASSIGNED_M"] > 1000 and orc.C["MULTI_ALIGN"] > 100
    check_against_oracle(graph, reads, res, orc)


@all_modes
def test_more_than_64_strains_uses_multiword_masks(mode):
    need_gpu()
    graph = synth.make_graph(21, n_clusters=30, chain_len=14, hap_lo=1, hap_hi=6, n_strains=150, bubble_p=0.4, flip_p=0.3)
    reads = synth.make_reads(graph, 22, n_pairs=2500, read_len=80, p_same_path=0.6)
    res = engine.abundance_pass(graph, reads, **mode)
    orc = oa.AbundanceOracle(graph.as_dict(), reads.as_dict()).run()
    assert orc.C["DIFF_CP"] > 0 and orc.C["INCONSIST"] > 0
    check_against_oracle(graph, reads, res, orc)


@all_modes
def test_explicit_coverage_records(mode):
    need_gpu()
    gkw, rkw = synth.CONFIGS["tiny"]
    graph = synth.make_graph(7, **gkw)
    reads = synth.make_reads(graph, 8, n_pairs=3000, read_len=60, p_exc=0.05)
    res = engine.abundance_pass(graph, reads, **mode)
    orc = oa.AbundanceOracle(graph.as_dict(), reads.as_dict()).run()
    check_against_oracle(graph, reads, res, orc)


@pytest.mark.parametrize("mode", [dict(tiles=False), dict(), dict(step_index=False)],
                         ids=["fileorder-occurrence-index", "tileorder-step-index", "tileorder-occurrence-index"])
def test_match_lists_equal_oracle(mode):
    need_gpu()
    graph, reads = synth.make_config("tiny")
    eng = engine.AbundanceEngine(graph, **mode)                 # (the tile kernels keep their match lists in shared memory)
    assert eng.step_index == (mode == dict())
    res = eng.run(reads)
    orc = oa.AbundanceOracle(graph.as_dict(), reads.as_dict())
    for a in range(reads.n_alns):
        mine = engine.decode_matches(res, a)
        assert [(p, pos) for p, pos, _ in mine] == orc.match(a), a
        # forward matches first, then reverse
        flags = [rv for _, _, rv in mine]
        assert flags == sorted(flags)


def _repetitive_case(seed, n_paths=40, alphabet=4, n_walks=4000):
    """Paths over a tiny node alphabet (repeated nodes, palindromes, many occurrences per node) and walks cut from
    them in both orientations, some run past either path end or altered: every corner of get_matching_path
    (json2csv.py:341-363) and of its reversed-walk twin (:861-866) -- several positions in one path, more than
    eight matches, overruns at slot 0 and at the last slot."""
    rng = np.random.default_rng(seed)
    paths, seen = [], set()
    while len(paths) < n_paths:
        n = int(rng.integers(1, 14))
        pth = tuple(int(x) for x in rng.integers(0, alphabet, n))
        if pth not in seen:                    # identical contents are one Path in the reference (:260)
            seen.add(pth)
            paths.append(list(pth))
    node_len = rng.integers(1, 40, alphabet).astype(np.int32)
    plen = np.array([len(x) for x in paths])
    path_off = np.concatenate(([0], np.cumsum(plen))).astype(np.int64)
    path_nodes = np.concatenate(paths).astype(np.int32)
    P = len(paths)
    graph = Graph(
        node_ids=np.arange(1, alphabet + 1), node_len=node_len, path_off=path_off, path_nodes=path_nodes,
        path_seq_len=np.array([int(node_len[x].sum()) for x in paths]), pstr_off=np.arange(P + 1),
        pstr_id=(np.arange(P) % 3).astype(np.int32), pstr_cnt=np.ones(P, np.int32),
        path_cluster=np.zeros(P, np.int32), clu_off=np.array([0, P]), clu_paths=np.arange(P, dtype=np.int32),
        strain_names=["NC_1.1", "NC_2.1", "NC_3.1"], header_strains=[0, 1, 2], path_names=None)
    walks = []
    for _ in range(n_walks):
        pth = paths[int(rng.integers(0, P))]
        i = int(rng.integers(0, len(pth)))
        j = int(rng.integers(i + 1, len(pth) + 1))
        wk = list(pth[i:j])
        kind = rng.random()
        if kind < 0.4:
            wk = wk[::-1]
        elif kind < 0.55:
            wk = wk + [int(rng.integers(0, alphabet))]                    # past the path end (or onto another path)
        elif kind < 0.7:
            wk = [int(rng.integers(0, alphabet))] + wk
        elif kind < 0.8:
            wk[int(rng.integers(0, len(wk)))] = int(rng.integers(0, alphabet))
        walks.append(wk)
    A = len(walks)
    wl = np.array([len(x) for x in walks])
    walk_node = np.concatenate(walks).astype(np.int32)
    reads = Reads(
        grp_nmates=np.ones(A, np.uint8), grp_aln_off=np.minimum(np.arange(2 * A + 1) + 1, 2 * A + 1) // 2,
        mate_seq_len=np.full(2 * A, 100, np.int32), aln_walk_off=np.concatenate(([0], np.cumsum(wl))),
        aln_read_len=np.full(A, 100, np.int32), aln_bins=np.zeros((A, 5), np.uint8), walk_node=walk_node,
        walk_alen=(1 | (node_len[walk_node].astype(np.int32) << 16)), exc_cov=np.zeros(0), mate_names=None)
    return graph, reads


@pytest.mark.parametrize("seed,alphabet", [(1, 2), (2, 3), (3, 9), (4, 20), (5, 40)])
def test_match_corner_cases_equal_oracle(seed, alphabet):
    need_gpu()
    graph, reads = _repetitive_case(seed, alphabet=alphabet)
    res = engine.abundance_pass(graph, reads, tiles=False)
    orc = oa.AbundanceOracle(graph.as_dict(), reads.as_dict())
    most = 0
    for a in range(reads.n_alns):
        mine = engine.decode_matches(res, a)
        assert [(p, pos) for p, pos, _ in mine] == orc.match(a), (a, reads.walk_node[reads.aln_walk_off[a]:reads.aln_walk_off[a + 1]])
        most = max(most, len(mine))
    if alphabet <= 3:
        assert most > 8                           # the serial path for long match lists was taken
    orc.run()
    check_against_oracle(graph, reads, res, orc)
    # the same corners against the shared-memory index: all 40 paths share the alphabet, so they are one tile
    eng = engine.AbundanceEngine(graph, tile_kernels=True)
    assert eng.tiling is not None and eng.tiling.n_tiles == 1 and eng.tile_kernels
    check_against_oracle(graph, reads, eng.run(reads), orc)


def _mixed_case(seed):
    """A graph whose first cluster is not a simple tile (repeated nodes inside paths) next to clusters that are: the
    step-index kernel matches the walks that start in the simple part and lists the others for the occurrence-index
    kernel; walks that cross from one part into the other match nothing either way."""
    rng = np.random.default_rng(seed)
    g0, r0 = _repetitive_case(seed, n_paths=12, alphabet=5, n_walks=600)
    base = g0.n_nodes
    paths = [g0.path_nodes[g0.path_off[p]:g0.path_off[p + 1]].tolist() for p in range(g0.n_paths)]
    n_rep = len(paths)
    clusters = [list(range(n_rep))]
    node_len = list(g0.node_len)
    for c in range(3):                                         # bubble chains: position i has one or two alleles
        alle = []
        for i in range(14):
            k = 2 if rng.random() < 0.5 else 1
            alle.append(list(range(base, base + k)))
            base += k
            node_len.extend(int(x) for x in rng.integers(1, 40, k))
        mine, seen = [], set()
        for _ in range(6):
            pth = tuple(int(a[int(rng.integers(0, len(a)))]) for a in alle)
            if rng.random() < 0.4:
                pth = pth[::-1]
            if pth not in seen:
                seen.add(pth)
                mine.append(len(paths))
                paths.append(list(pth))
        clusters.append(mine)
    P = len(paths)
    node_len = np.asarray(node_len, np.int32)
    plen = np.array([len(x) for x in paths])
    graph = Graph(
        node_ids=np.arange(1, base + 1), node_len=node_len, path_off=np.concatenate(([0], np.cumsum(plen))).astype(np.int64),
        path_nodes=np.concatenate(paths).astype(np.int32), path_seq_len=np.array([int(node_len[x].sum()) for x in paths]),
        pstr_off=np.arange(P + 1), pstr_id=(np.arange(P) % 3).astype(np.int32), pstr_cnt=np.ones(P, np.int32),
        path_cluster=np.concatenate([np.full(len(m), c, np.int32) for c, m in enumerate(clusters)]),
        clu_off=np.concatenate(([0], np.cumsum([len(m) for m in clusters]))), clu_paths=np.arange(P, dtype=np.int32),
        strain_names=["NC_1.1", "NC_2.1", "NC_3.1"], header_strains=[0, 1, 2], path_names=None)
    walks = [r0.walk_node[r0.aln_walk_off[a]:r0.aln_walk_off[a + 1]].tolist() for a in range(r0.n_alns)]
    crossed = np.unique(np.concatenate(paths))                 # (a walk that starts on a node no path crosses raises, Q10)

    def any_node():
        return int(crossed[int(rng.integers(0, len(crossed)))])
    for _ in range(2400):
        pth = paths[int(rng.integers(n_rep, P))]
        i = int(rng.integers(0, len(pth)))
        j = int(rng.integers(i + 1, min(len(pth), i + 9) + 1))
        wk = list(pth[i:j])
        kind = rng.random()
        if kind < 0.4:
            wk = wk[::-1]
        elif kind < 0.5:
            wk = wk + [any_node()]                             # onto any node, the non-simple part included
        elif kind < 0.6:
            wk = [any_node()] + wk
        elif kind < 0.7:
            wk[int(rng.integers(0, len(wk)))] = any_node()
        walks.append(wk)
    order = rng.permutation(len(walks))
    walks = [walks[k] for k in order]
    A = len(walks)
    wl = np.array([len(x) for x in walks])
    walk_node = np.concatenate(walks).astype(np.int32)
    reads = Reads(
        grp_nmates=np.ones(A, np.uint8), grp_aln_off=np.minimum(np.arange(2 * A + 1) + 1, 2 * A + 1) // 2,
        mate_seq_len=np.full(2 * A, 100, np.int32), aln_walk_off=np.concatenate(([0], np.cumsum(wl))),
        aln_read_len=np.full(A, 100, np.int32), aln_bins=np.zeros((A, 5), np.uint8), walk_node=walk_node,
        walk_alen=(1 | (node_len[walk_node].astype(np.int32) << 16)), exc_cov=np.zeros(0), mate_names=None)
    return graph, reads


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_step_index_with_occurrence_index_leftovers(seed):
    need_gpu()
    graph, reads = _mixed_case(seed)
    eng = engine.AbundanceEngine(graph)
    assert eng.step_index and eng.dg.steps["steps_partial"]
    res = eng.run(reads)
    orc = oa.AbundanceOracle(graph.as_dict(), reads.as_dict())
    n_match = 0
    for a in range(reads.n_alns):
        mine = engine.decode_matches(res, a)
        assert [(p, pos) for p, pos, _ in mine] == orc.match(a), (a, reads.walk_node[reads.aln_walk_off[a]:reads.aln_walk_off[a + 1]])
        n_match += len(mine)
    assert n_match > reads.n_alns // 2
    orc.run()
    check_against_oracle(graph, reads, res, orc)
    check_against_oracle(graph, reads, engine.abundance_pass(graph, reads, step_index=False), orc)


def test_match_buffer_regrows():
    need_gpu()
    graph, reads = synth.make_config("tiny")
    eng = engine.AbundanceEngine(graph)
    eng._work_for = lambda dr, _orig=eng._work_for: _small_work(eng, dr)
    res = eng.run(reads)
    orc = oa.AbundanceOracle(graph.as_dict(), reads.as_dict()).run()
    check_against_oracle(graph, reads, res, orc)


def _small_work(eng, dr):
    cap = max(eng.match_cap_hint, 16)
    if eng.work is None or eng.work.match_cap < cap or eng.work.n_alns < dr.n_alns:
        eng.work = engine.Work(dr.n_groups, dr.n_alns, cap, eng.device)
    return eng.work


@all_modes
@pytest.mark.parametrize("d", sorted(glob.glob(os.path.join(HERE, "golden", "case_*"))), ids=os.path.basename)
def test_golden_cases_through_cuda(d, mode):
    """The reference's own outputs (tests/golden) reproduced by the CUDA path: ints exact, fp64 <= 1e-9."""
    need_gpu()
    with open(os.path.join(d, "expected.json")) as fh:
        exp = json.load(fh)
    g = og.load_graph(os.path.join(d, "graph.gfa"), os.path.join(d, "dict_clusters.pickle"))
    r = od.decode_json(os.path.join(d, "reads.json"), g, exp["params"]["thr"])
    graph, reads = Graph.from_dict(g), Reads.from_dict(r)
    res = engine.abundance_pass(graph, reads, **mode)
    compare.check_state_vs_reference(res, exp, exact=False)


@all_modes
def test_empty_and_degenerate_inputs(mode):
    need_gpu()
    graph, reads = synth.make_config("tiny")
    empty = reads.slice_groups(0, 0)
    res = engine.abundance_pass(graph, empty, **mode)
    assert res.C["NB_READS"] == 0 and not res.uniq_abund.any()
    one = reads.slice_groups(5, 6)
    res = engine.abundance_pass(graph, one, **mode)
    orc = oa.AbundanceOracle(graph.as_dict(), one.as_dict()).run()
    check_against_oracle(graph, one, res, orc)


def test_error_quirks_raise_like_reference():
    """Q7 / Q10: same exception classes as the reference (and as the oracle on the same arrays)."""
    need_gpu()
    graph, reads = synth.make_config("tiny")
    bad = Reads.from_dict(reads.as_dict())
    bad.aln_bins = bad.aln_bins.copy()
    bad.aln_bins[:, 0] = 255                       # score outside [0, 1] used as a histogram key
    with pytest.raises(KeyError):
        engine.abundance_pass(graph, bad)
    with pytest.raises(KeyError):
        oa.AbundanceOracle(graph.as_dict(), bad.as_dict()).run()
    # an unassigned read is recorded in the cluster of the first path through its first node: cluster None -> TypeError
    g2 = Graph.from_dict(graph.as_dict())
    g2.path_cluster = np.full_like(graph.path_cluster, -1)
    orc = oa.AbundanceOracle(graph.as_dict(), reads.as_dict()).run()
    assert orc.book, "the configuration must record some unassigned reads"
    with pytest.raises(TypeError):
        engine.abundance_pass(g2, reads)
    with pytest.raises(TypeError):
        oa.AbundanceOracle(g2.as_dict(), reads.as_dict()).run()
    # ... and a first node that no path crosses -> IndexError
    g3 = synth.make_graph(5, n_clusters=6, chain_len=10, hap_lo=1, hap_hi=2, n_strains=3, bubble_p=0.6, flip_p=0.0)
    off_path = np.flatnonzero(~g3.on_path)
    assert off_path.size, "needs a node outside every path"
    r3 = synth.make_reads(g3, 6, n_pairs=50, read_len=40)
    r3 = Reads.from_dict(r3.as_dict())
    r3.walk_node = r3.walk_node.copy()
    single = np.flatnonzero(np.diff(r3.aln_walk_off) >= 1)
    r3.walk_node[r3.aln_walk_off[single]] = off_path[0]          # every walk now starts on that node: nothing matches
    with pytest.raises(IndexError):
        engine.abundance_pass(g3, r3)
    with pytest.raises(IndexError):
        oa.AbundanceOracle(g3.as_dict(), r3.as_dict()).run()


@pytest.mark.parametrize("variant", ["plain", "wide_steps", "long_nodes", "ragged_lengths"])
def test_transfer_layout_expands_to_the_same_arrays(variant):
    """sfb_expand_reads(wire(reads)) == the wide arrays, bit for bit (offset scans across tile boundaries, node ids
    from steps incl. wide ones, bins from the dictionary, explicit coverages, read lengths)."""
    need_gpu()
    gkw, _ = synth.CONFIGS["small"]
    graph = synth.make_graph(3, **gkw)
    reads = synth.make_reads(graph, 4, n_pairs=30_000, read_len=150, p_exc=0.01)
    casegen.stress_transfer_layout(graph, reads, np.random.default_rng(5), variant)
    eng = engine.AbundanceEngine(graph)
    host = engine.HostReads(reads, wire=True)
    assert host.wire and host.nbytes() < (0.25 if variant == "plain" else 0.6) * engine.HostReads(reads, wire=False).nbytes()
    assert (host.n_wide > 0) == (variant == "wide_steps") and (host.n_bins_exc > 0) == (variant == "ragged_lengths")
    dr = eng.upload(host)
    torch.cuda.synchronize()

    # read the expanded arrays back through the pointers of the sfb_reads struct
    def view(ptr, n, dt):
        nbytes = n * torch.empty(0, dtype=dt).element_size()
        buf = dr.wide if dr.wide.data_ptr() <= ptr < dr.wide.data_ptr() + dr.wide.numel() else dr.arena
        off = ptr - buf.data_ptr()
        return buf[off:off + nbytes].view(dt).cpu().numpy()

    G, A, R = reads.n_groups, reads.n_alns, reads.n_records
    assert np.array_equal(view(dr.c.grp_aln_off, 2 * G + 1, torch.int64), reads.grp_aln_off)
    assert np.array_equal(view(dr.c.aln_walk_off, A + 1, torch.int64), reads.aln_walk_off)
    assert np.array_equal(view(dr.c.mate_seq_len, 2 * G, torch.int32), reads.mate_seq_len)
    assert np.array_equal(view(dr.c.walk_node, R, torch.int32), reads.walk_node)
    assert np.array_equal(view(dr.c.aln_bins, 5 * A, torch.uint8).reshape(A, 5), reads.aln_bins)

    def coverage(word, exc):
        word = word.astype(np.int64)
        neg = word < 0
        cov = np.where(neg, 0, word & 0xFFFF) / np.where(neg, 1, word >> 16)
        cov[neg] = exc[-word[neg] - 1]
        return cov

    exc_dev = view(dr.c.exc_cov, host.n_exc, torch.float64) if host.n_exc else np.zeros(0)
    got = view(dr.c.walk_alen, R, torch.int32)
    assert np.array_equal(coverage(got, exc_dev), coverage(reads.walk_alen, reads.exc_cov))
    assert np.array_equal(got[got >= 0], reads.walk_alen[got >= 0])
    if variant != "plain":
        return
    # and the pass gives the same result from either layout
    a = eng.run(engine.HostReads(reads, wire=True))
    b = eng.run(engine.HostReads(reads, wire=False))
    assert a.C == b.C and np.array_equal(a.uniq_reads, b.uniq_reads) and np.array_equal(a.grp_code, b.grp_code)
    assert np.allclose(a.mult_abund, b.mult_abund, rtol=1e-12, atol=0)


def test_transfer_layout_long_walks_take_the_direct_path():
    """Walks of ~35 nodes: 256 of them do not fit the shared-memory image of k_expand_walks, the block writes them
    straight to global memory; mixed with a stretch of short walks that is staged."""
    need_gpu()
    graph = synth.make_graph(11, n_clusters=60, chain_len=150, hap_lo=1, hap_hi=4, n_strains=8, bubble_p=0.3, flip_p=0.2,
                             max_node_len=7)
    long_reads = synth.make_reads(graph, 12, n_pairs=4000, read_len=150)
    short_reads = synth.make_reads(graph, 13, n_pairs=4000, read_len=12)
    assert long_reads.n_records > 20 * long_reads.n_alns and short_reads.n_records < 8 * short_reads.n_alns
    eng = engine.AbundanceEngine(graph)
    for reads in (long_reads, short_reads):
        host = engine.HostReads(reads, wire=True)
        assert host.wire
        a = eng.run(host)
        b = eng.run(engine.HostReads(reads, wire=False))
        assert a.C == b.C and np.array_equal(a.uniq_reads, b.uniq_reads) and np.array_equal(a.grp_code, b.grp_code)
        assert np.array_equal(a.aln_assign, b.aln_assign)
        assert np.allclose(a.uniq_abund, b.uniq_abund, rtol=1e-12, atol=0)
        assert np.allclose(a.mult_abund, b.mult_abund, rtol=1e-12, atol=0)
        orc = oa.AbundanceOracle(graph.as_dict(), reads.as_dict()).run()
        check_against_oracle(graph, reads, a, orc)


def test_streamed_shards_equal_single_pass():
    """run_stream (uploads overlapped with kernels, shards accumulated into the same arrays) == run."""
    need_gpu()
    from strainflair_b200 import dist as sfdist
    graph, reads = synth.make_config("small", scale=0.5)
    eng = engine.AbundanceEngine(graph)
    one = eng.run(reads, keep_codes=False)
    for k in (1, 3, 7):
        shards = [engine.HostReads(reads.slice_groups(*sfdist.shard_bounds(reads.n_groups, k, i))) for i in range(k)]
        for _ in range(2):                                   # second call reuses the device buffers
            many = eng.run_stream(shards)
            for name in oa.COUNTERS:
                assert many.C[name] == one.C[name], (k, name)
            assert np.array_equal(many.uniq_reads, one.uniq_reads) and np.array_equal(many.hist, one.hist)
            assert many.n_deferred == one.n_deferred
            ok = np.abs(many.table - one.table) <= 1e-9 * np.maximum(np.abs(many.table), np.abs(one.table))
            assert (ok | (np.isnan(many.table) & np.isnan(one.table))).all()
            assert np.allclose(many.mult_reads, one.mult_reads, rtol=1e-9, atol=0)


def test_shards_cut_at_tile_boundaries_equal_single_pass():
    """run_stream over engine.stage_shards (tile order first, shards cut at tile boundaries) == run: integers exact, the
    order-independent fp64 sums bit for bit."""
    need_gpu()
    graph, reads = synth.make_config("small", scale=0.5)
    eng = engine.AbundanceEngine(graph)
    one = eng.run(reads, keep_codes=False, keep_slots=False)
    for k in (1, 2, 5):
        many = eng.run_stream(eng.stage_shards(reads, k))
        for name in oa.COUNTERS:
            assert many.C[name] == one.C[name], (k, name)
        assert np.array_equal(many.uniq_reads, one.uniq_reads) and np.array_equal(many.hist, one.hist)
        assert many.n_deferred == one.n_deferred
        assert np.array_equal(many.table, one.table, equal_nan=True)
        for name in ("uniq_norm", "mult_reads", "mult_norm"):
            assert np.array_equal(getattr(many, name), getattr(one, name)), (k, name)


def test_pipelined_jobs_equal_single_passes():
    """submit/collect with two jobs in flight (uploads of job k+1 under pass 2 of job k, each job on its own
    accumulators and device buffers): every job returns what a plain run of its reads returns, in any interleaving,
    including the first jobs whose match buffers are still too small and are run again."""
    need_gpu()
    from collections import deque
    from strainflair_b200 import dist as sfdist
    graph, reads_a = synth.make_config("small", scale=0.5)
    reads_b = synth.make_reads(graph, 4242, n_pairs=9000, read_len=150, p_same_path=0.7)
    eng = engine.AbundanceEngine(graph, deterministic=True)
    ref = {"a": eng.run(reads_a, keep_codes=False), "b": eng.run(reads_b, keep_codes=False)}
    sets = {}
    for name, rd, k in (("a", reads_a, 4), ("b", reads_b, 3)):
        sets[name] = [engine.HostReads(rd.slice_groups(*sfdist.shard_bounds(rd.n_groups, k, i))) for i in range(k)]
    order = ["a", "b", "b", "a", "a", "b", "a"]
    pending, got = deque(), []
    for name in order:
        pending.append((name, eng.submit(sets[name], depth=2)))
        if len(pending) == 2:
            nm, job = pending.popleft()
            got.append((nm, eng.collect(job)))
    while pending:
        nm, job = pending.popleft()
        got.append((nm, eng.collect(job)))
    assert [nm for nm, _ in got] == order
    for nm, res in got:
        one = ref[nm]
        for name in oa.COUNTERS:
            assert res.C[name] == one.C[name], (nm, name)
        assert np.array_equal(res.uniq_reads, one.uniq_reads) and np.array_equal(res.hist, one.hist)
        assert res.n_deferred == one.n_deferred
        # order-independent accumulation: bit-identical whatever the sharding and the interleaving
        assert np.array_equal(res.table, one.table, equal_nan=True)
        assert np.array_equal(res.mult_reads, one.mult_reads) and np.array_equal(res.uniq_norm, one.uniq_norm)
    # a third job without collecting the oldest would reuse its buffers: refused
    j1, j2 = eng.submit(sets["a"], depth=2), eng.submit(sets["b"], depth=2)
    with pytest.raises(RuntimeError):
        eng.submit(sets["a"], depth=2)
    assert np.array_equal(eng.collect(j1).table, ref["a"].table, equal_nan=True)
    j3 = eng.submit(sets["a"], depth=2)
    assert np.array_equal(eng.collect(j2).table, ref["b"].table, equal_nan=True)
    assert np.array_equal(eng.collect(j3).table, ref["a"].table, equal_nan=True)
    # the plain entry still works on the engine afterwards
    again = eng.run(reads_b, keep_codes=False)
    assert np.array_equal(again.table, ref["b"].table, equal_nan=True)


def test_order_independent_mode_is_bit_reproducible():
    """deterministic=True: fp64 sums do not depend on thread order or on how the reads are sharded (bit-identical
    across runs and across run / run_stream with any number of shards), and still match the oracle to 1e-9."""
    need_gpu()
    from strainflair_b200 import dist as sfdist
    graph, reads = synth.make_config("small", scale=0.5)
    eng = engine.AbundanceEngine(graph, deterministic=True)
    a = eng.run(reads)
    b = eng.run(reads)
    for name in ("uniq_norm", "mult_reads", "mult_norm", "uniq_abund", "mult_abund", "table"):
        x, y = getattr(a, name), getattr(b, name)
        assert np.array_equal(x, y, equal_nan=True), name
    for k in (2, 5):
        shards = [engine.HostReads(reads.slice_groups(*sfdist.shard_bounds(reads.n_groups, k, i))) for i in range(k)]
        c = eng.run_stream(shards)
        for name in ("uniq_norm", "mult_reads", "mult_norm", "table"):
            assert np.array_equal(getattr(a, name), getattr(c, name), equal_nan=True), (k, name)
    orc = oa.AbundanceOracle(graph.as_dict(), reads.as_dict()).run()
    check_against_oracle(graph, reads, a, orc)
    # the default mode gives the same numbers up to the order of the additions
    fast = engine.AbundanceEngine(graph).run(reads)
    assert np.allclose(fast.mult_abund, a.mult_abund, rtol=1e-12, atol=0)
