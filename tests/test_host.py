"""CPU-only tests of the host side: GFA loader and native JSON decoder against the oracle's, on the golden
fixtures (which hold the reference's outputs) and on fresh random cases; CLI argument handling."""
import glob
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import casegen
from oracle import decode as od
from oracle import graph as og
from strainflair_b200 import decode, gfa

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CASES = sorted(glob.glob(os.path.join(HERE, "golden", "case_*")))
GRAPH_ARRAYS = ("node_ids", "node_len", "path_off", "path_nodes", "path_seq_len", "pstr_off", "pstr_id", "pstr_cnt",
                "path_cluster", "clu_off", "clu_paths")
READS_ARRAYS = ("grp_nmates", "grp_aln_off", "mate_seq_len", "aln_walk_off", "aln_read_len", "aln_bins", "walk_node",
                "walk_alen", "exc_cov")


def same_graph(a, b):
    for k in GRAPH_ARRAYS:
        assert np.array_equal(getattr(a, k), b[k]), k
    assert a.strain_names == b["strain_names"] and a.header_strains == b["header_strains"]
    assert a.path_names == b["path_names"]


def same_reads(r, ro):
    for k in READS_ARRAYS:
        assert np.array_equal(getattr(r, k), ro[k]), k
    assert np.array_equal(r.walk_cov, ro["walk_cov"])          # bit-exact fp64 coverages
    assert np.array_equal(r.aln_stats, ro["aln_stats"])
    assert [decode.mate_name(r, i // 2, i % 2) for i in range(2 * r.n_groups)] == ro["mate_names"]


@pytest.mark.parametrize("d", CASES, ids=os.path.basename)
def test_loader_and_decoder_match_oracle_on_golden(d):
    files = [os.path.join(d, f) for f in ("graph.gfa", "reads.json", "dict_clusters.pickle")]
    g = gfa.load_graph(files[0], files[2])
    go = og.load_graph(files[0], files[2])
    same_graph(g, go)
    for thr, threads, block in ((0.95, 0, 0), (0.5, 3, 700), (0.0, 1, 4096), (1.0, 2, 100_000)):
        same_reads(decode.decode_json(files[1], g, thr, threads=threads or None, block_bytes=block, want_cov=True),
                   od.decode_json(files[1], go, thr))


@pytest.mark.parametrize("seed", range(500, 530))
def test_decoder_matches_oracle_on_random_cases(seed, tmp_path):
    a, b, c = casegen.make_case(seed, str(tmp_path), n_groups=70, n_clusters=2 + seed % 4, n_strains=2 + seed % 3)
    g = gfa.load_graph(a, c)
    go = og.load_graph(a, c)
    same_graph(g, go)
    same_reads(decode.decode_json(b, g, 0.95, threads=1 + seed % 4, block_bytes=[0, 300, 5000][seed % 3], want_cov=True),
               od.decode_json(b, go, 0.95))


def _write(tmp_path, lines):
    p = tmp_path / "r.json"
    p.write_text("".join(json.dumps(x) + "\n" if not isinstance(x, str) else x for x in lines))
    return str(p)


GFA = "S\t1\tACGTACGTAC\nS\t2\tACGTA\nS\t3\tAAAA\nP\tNC_1.1_1\t1+,2+\t*\n"


def _line(name, node, n, score=None, **extra):
    ln = {"name": name, "sequence": "A" * n, "start": [0], "subpath": [
        {"path": {"mapping": [{"position": {"node_id": str(node)}, "edit": [{"from_length": n, "to_length": n}]}]}}]}
    if score is not None:
        ln["subpath"][0]["score"] = score
    ln.update(extra)
    return ln


@pytest.mark.parametrize("which,exc", [("unknown_node", KeyError), ("bad_json", ValueError), ("zero_len", ZeroDivisionError),
                                        ("no_sequence", KeyError), ("next_out_of_range", IndexError)])
def test_decoder_errors_match_oracle(which, exc, tmp_path):
    (tmp_path / "g.gfa").write_text(GFA)
    g = gfa.load_graph(str(tmp_path / "g.gfa"))
    go = og.load_graph(str(tmp_path / "g.gfa"))
    if which == "unknown_node":
        lines = [_line("a", 99, 4, 4)]
    elif which == "bad_json":
        lines = [_line("a", 1, 4, 4), "{not json\n"]
    elif which == "zero_len":
        ln = _line("a", 1, 4, 4)
        ln["subpath"][0]["path"]["mapping"][0]["edit"] = [{"from_length": 3}]
        lines = [ln]
    elif which == "no_sequence":
        ln = _line("a", 1, 4, 4)
        del ln["sequence"]
        lines = [ln]
    else:
        ln = _line("a", 1, 4, 4)
        ln["subpath"][0]["next"] = [7]
        ln["sequence"] = "A" * 9
        lines = [ln]
    p = _write(tmp_path, lines)
    with pytest.raises(exc):
        od.decode_json(p, go, 0.95)
    with pytest.raises(exc):
        decode.decode_json(p, g, 0.95)


def test_decoder_handles_int_and_string_fields(tmp_path):
    """vg writes 64-bit ints as strings and 32-bit ints as numbers; the reference int()s both (json2csv.py:562-567)."""
    (tmp_path / "g.gfa").write_text(GFA)
    g = gfa.load_graph(str(tmp_path / "g.gfa"))
    go = og.load_graph(str(tmp_path / "g.gfa"))
    ln = _line("r", 1, 6, 6)
    ln["subpath"][0]["path"]["mapping"][0]["position"] = {"node_id": 1, "offset": "2", "is_reverse": True}
    ln["subpath"][0]["path"]["mapping"][0]["edit"] = [{"from_length": "3", "to_length": "3"},
                                                      {"from_length": 3, "to_length": 3, "sequence": "TTT"}]
    ln["mapping_quality"] = 60
    ln["quality"] = "ISEhISEh"
    ln["annotation"] = {"nested": [1, {"a": [True, None, "x\\\"y"]}]}
    p = _write(tmp_path, [ln])
    same_reads(decode.decode_json(p, g, 0.5, want_cov=True), od.decode_json(p, go, 0.5))


def _cli(args, cwd):
    return subprocess.run([sys.executable, "-m", "strainflair_b200"] + args, cwd=cwd, capture_output=True, text=True,
                          env=dict(os.environ, PYTHONPATH=ROOT))


def test_cli_flag_handling_matches_reference(tmp_path):
    """-f is rejected with exit code 2 (Q2); missing inputs print the usage and exit 0 (json2csv.py:1295-1323)."""
    r = _cli(["json2csv", "-f"], str(tmp_path))
    assert r.returncode == 2 and "option -f not recognized" in r.stdout and "Usage:" in r.stdout
    r = _cli(["json2csv", "-g", "x.gfa"], str(tmp_path))
    assert r.returncode == 0 and r.stdout.startswith("Usage:")
    r = _cli(["compute_strains_abundance", "-x"], str(tmp_path))
    assert r.returncode == 2 and "option -x not recognized" in r.stdout
    r = _cli(["compute_strains_abundance", "-i", "a.csv"], str(tmp_path))
    assert r.returncode == 0 and r.stdout.startswith("Usage:")


def test_synthetic_workload_round_trips_through_files(tmp_path):
    """The bench workload is defined on arrays; exported as GFA/JSON/pickle and decoded again it must be the
    same arrays (so the arrays mean what the reference would read from those files)."""
    from strainflair_b200 import synth
    for name, scale in (("tiny", 1.0), ("small", 0.04)):
        g, r = synth.make_config(name, scale=scale)
        gf, js, pk = synth.export_files(g, r, str(tmp_path / name))
        g2 = gfa.load_graph(gf, pk)
        for k in ("node_len", "path_off", "path_nodes", "path_seq_len", "pstr_off", "pstr_cnt", "path_cluster",
                  "clu_off", "clu_paths"):
            assert np.array_equal(getattr(g, k), getattr(g2, k)), k
        assert [g.strain_names[s] for s in g.pstr_id] == [g2.strain_names[s] for s in g2.pstr_id]
        r2 = decode.decode_json(js, g2, 0.95)
        for k in ("grp_nmates", "grp_aln_off", "mate_seq_len", "aln_walk_off", "aln_read_len", "aln_bins", "walk_node",
                  "walk_alen"):
            assert np.array_equal(getattr(r, k), getattr(r2, k)), (name, k)


@pytest.mark.parametrize("which,exc", [("path_before_segment", KeyError), ("short_line", IndexError), ("bad_id", ValueError)])
def test_gfa_errors_match_oracle(which, exc, tmp_path):
    text = {"path_before_segment": "P\tNC_1.1_1\t1+,2+\t*\nS\t1\tAC\nS\t2\tA\n",
            "short_line": "S\t1\n", "bad_id": "S\tx1\tACGT\n"}[which]
    p = tmp_path / "g.gfa"
    p.write_text(text)
    with pytest.raises(exc):
        og.load_graph(str(p))
    with pytest.raises(exc):
        gfa.load_graph(str(p))


def test_gfa_corner_cases_match_oracle(tmp_path):
    """Orientation flips, single-node paths, names without accession (None strain, Q14), duplicate content from
    several strains, re-declared segments, CRLF and trailing blanks, W/L/H lines ignored."""
    text = ("H\tVN:Z:1.0\r\n"
            "S\t10\tACGTAC\nS\t7\tAC\nS\t3\tA\nS\t10\tACG\n"          # node 10 re-declared: length 3
            "P\tNC_000913.3_12\t10+,7+,3+\t*\n"                               # reversed into 3-,7-,10-
            "P\tgi|55|ref|NZ_CP1.1|_9\t3-,7-,10-\t*\n"                        # same content, another strain
            "P\tnoaccession\t7+\t*\n"                                         # single node, no '_' -> None strain
            "P\tplasmid_x_1\t7-\t*\n"                                         # single node forced '+': same path, None again
            "P\tNC_000913.3_13\t3+,7+\t*  \n"
            "P\tAB12|zz_4\t3+,7-\t*\n"                                        # same nodes, other signs: a distinct path
            "L\t3\t+\t7\t+\t0M\nW\tx\t1\tc\t0\t5\t>3>7\n\n")
    p = tmp_path / "g.gfa"
    p.write_text(text)
    same_graph(gfa.load_graph(str(p)), og.load_graph(str(p)))


def _expand_on_host(w, common, n_nodes_len):
    """numpy restatement of sfb_expand_reads (csrc/expand.cu)."""
    G, A = len(w["grp_nmates"]), len(w["aln_len"])
    grp_aln_off = np.concatenate([[0], np.cumsum(w["mate_naln"], dtype=np.int64)])
    off = np.concatenate([[0], np.cumsum(w["aln_len"], dtype=np.int64)])
    if len(w["mate_seq_len"]):
        seq = w["mate_seq_len"].astype(np.int32)
    else:
        seq = np.where((np.arange(2 * G) & 1) < np.repeat(w["grp_nmates"], 2), common, 0).astype(np.int32)
        seq[w["seq_exc_idx"]] = w["seq_exc_val"]
    bins = np.full((A, 5), 255, np.uint8)
    known = w["aln_code"] < 255
    bins[known] = w["bins_dict"].reshape(255, 5)[w["aln_code"][known]]
    bins[w["bins_exc_idx"]] = w["bins_exc_val"].reshape(-1, 5)
    R = int(off[-1])
    node, word = np.empty(R, np.int32), np.empty(R, np.int32)
    wide = dict(zip(w["wide_idx"].tolist(), w["wide_node"].tolist()))
    for a in range(A):
        anchor = int(w["node0_anchor"][a >> 8])           # first node: 16 bits above the block's smallest one, or listed
        cur = anchor + int(w["aln_node0"][a]) if anchor >= 0 else int(w["node0_wide"][256 * (-1 - anchor) + (a & 255)])
        for r in range(off[a], off[a + 1]):
            if r > off[a]:
                nib = (int(w["walk_step"][r >> 1]) >> (4 * (r & 1))) & 15          # step + 8, 0 = listed
                cur = wide[r] if nib == 0 else cur + nib - 8
            nl = int(n_nodes_len[cur])
            al = int(w["aln_al_first"][a]) if r == off[a] else int(w["aln_al_last"][a]) if r == off[a + 1] - 1 else nl & 0xFFFF
            node[r], word[r] = cur, np.int64(al | (nl << 16)).astype(np.int32)
    word[w["exc_idx"]] = -(np.arange(len(w["exc_idx"]), dtype=np.int32) + 1)
    return grp_aln_off, off, seq, bins, node, word


def _coverage(word, exc):
    word = word.astype(np.int64)
    neg = word < 0
    cov = np.where(neg, 0, word & 0xFFFF) / np.where(neg, 1, word >> 16)
    cov[neg] = exc[-word[neg] - 1]
    return cov


@pytest.mark.parametrize("variant", ["plain", "wide_steps", "long_nodes", "ragged_lengths"])
def test_transfer_layout_round_trip_on_host(variant):
    """engine.wire_arrays holds everything needed to rebuild the shard (what sfb_expand_reads does on the device)."""
    from strainflair_b200 import engine, synth
    gkw, _ = synth.CONFIGS["tiny"]
    graph = synth.make_graph(11, **gkw)
    reads = synth.make_reads(graph, 12, n_pairs=3000, read_len=150, p_exc=0.02)
    rng = np.random.default_rng(1)
    casegen.stress_transfer_layout(graph, reads, rng, variant)
    w, common = engine.wire_arrays(reads)
    assert (len(w["wide_idx"]) > 0) == (variant == "wide_steps")
    assert (len(w["mate_seq_len"]) > 0) == (variant == "ragged_lengths") and (len(w["bins_exc_idx"]) > 0) == (variant == "ragged_lengths")
    assert np.all(np.diff(w["exc_idx"]) > 0) and np.all(np.diff(w["wide_idx"]) > 0)
    grp_aln_off, off, seq, bins, node, word = _expand_on_host(w, common, graph.node_len)
    assert np.array_equal(grp_aln_off, reads.grp_aln_off) and np.array_equal(off, reads.aln_walk_off)
    assert np.array_equal(seq, reads.mate_seq_len) and np.array_equal(bins, reads.aln_bins)
    assert np.array_equal(node, reads.walk_node)
    assert np.array_equal(_coverage(word, w["exc_cov"]), _coverage(reads.walk_alen, reads.exc_cov))
    unlisted = np.setdiff1d(np.arange(reads.n_records), w["exc_idx"])
    assert np.array_equal(word[unlisted], reads.walk_alen[unlisted])
    if variant == "plain":
        assert len(w["exc_idx"]) < len(reads.exc_cov) + 0.01 * reads.n_records
        # a shard whose exceptions would outweigh the saving falls back to the plain layout
        reads.walk_node[:] = rng.integers(0, graph.n_nodes, reads.n_records)
        assert engine.wire_arrays(reads) is None


def _random_walk_sets(rng, n_clusters):
    """Per cluster a handful of walks over a small node alphabet: sub-walks of a few 'haplotypes' (so that many agree),
    mutated, reversed, with repeated nodes and cyclic shifts (so that the choice of the anchor node matters), node ids
    from ranges that collide in CPython's set table (multiples of 8, huge ids, mixed)."""
    out = []
    for c in range(n_clusters):
        style = c % 6
        base = [1, 1000, 8, 2 ** 20, 2 ** 40, 7][style]
        stride = [1, 1, 8, 64, 2 ** 31, 3][style]
        alphabet = [base + stride * k for k in range(int(rng.integers(4, 40)))]
        haps = []
        for _ in range(int(rng.integers(1, 4))):
            h = list(alphabet)
            for _ in range(int(rng.integers(0, 4))):
                h[int(rng.integers(len(h)))] = int(rng.choice(alphabet))           # repeated nodes
            if rng.random() < 0.3:
                k = int(rng.integers(1, len(h)))
                h = h[k:] + h[:k]                                                  # cyclic shift
            haps.append(h)
        walks = []
        for _ in range(int(rng.integers(0, 14))):
            h = haps[int(rng.integers(len(haps)))]
            i = int(rng.integers(len(h)))
            w = h[i:i + int(rng.integers(1, 12))]
            if rng.random() < 0.25:
                w[int(rng.integers(len(w)))] = int(rng.choice(alphabet))
            if rng.random() < 0.4:
                w = w[::-1]
            walks.append([int(v) for v in w])
        out.append(walks)
    return out


def test_native_min_strains_matches_the_python_restatement():
    """csrc/unassigned.cpp (CPython set order replayed, find_peaks and clique sizes restated) == oracle/report.py,
    which runs the real set / scipy / networkx code of the reference's Cluster.min_strains_inference."""
    from oracle import report as orp
    from strainflair_b200 import report
    rng = np.random.default_rng(77)
    sets = _random_walk_sets(rng, 1500)
    sets.append([])                                     # a cluster without unassigned reads
    sets.append([[5]])                                  # one single-node walk
    sets.append([[3, 2, 1], [1, 2, 3], [2, 3, 4], [9, 3, 4, 1]])
    got = report.min_strains(sets, threads=3)
    n_filtered = 0
    for k, walks in enumerate(sets):
        if not walks:
            assert (got[0][k], got[1][k], got[2][k], got[3][k]) == (0, 0, 0, 0.0)
            continue
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = orp.min_strains([list(w) for w in walks])
        assert (int(got[0][k]), int(got[1][k]), int(got[2][k])) == tuple(int(v) for v in want[:3]), (k, walks)
        assert (np.isnan(got[3][k]) and np.isnan(want[3])) or got[3][k] == want[3], (k, walks)
        n_filtered += want[1] > 0
    assert n_filtered > 50                               # the support filter keeps something often enough to be tested
    one = report.min_strains(sets, threads=1)
    assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(got, one))


def test_relation_pivot_follows_cpython_set_order():
    """Pairs of walks whose verdict depends on which shared node anchors the comparison: the native code must pick
    CPython's (first element of the set intersection in hash-table order)."""
    from oracle import report as orp
    from strainflair_b200 import report
    rng = np.random.default_rng(3)
    n_dep = 0
    for _ in range(3000):
        k = int(rng.integers(3, 30))
        ids = rng.choice(np.arange(1, 400) * int(rng.choice([1, 8, 16, 1024])), k, replace=False).tolist()
        a = ids[:int(rng.integers(2, k + 1))]
        b = list(a)
        rng.shuffle(b)
        b = b[:int(rng.integers(1, len(b) + 1))] + [int(v) + 100000 for v in ids[:int(rng.integers(0, 3))]]
        verdicts = set()
        shared = set(a) & set(b)
        for pivot in shared:                             # what each possible anchor would say
            i, j = a.index(pivot), b.index(pivot)
            left, right = min(i, j), min(len(a) - i, len(b) - j)
            verdicts.add(a[i - left:i + right] == b[j - left:j + right])
        n_dep += len(verdicts) > 1
        want = orp.compat(orp.orient_small(a), orp.orient_small(b))
        before = int(report.min_strains([[a, b]], threads=1)[0][0])
        assert before == (2 if want == -1 else 1), (a, b)
    assert n_dep > 100


@pytest.mark.parametrize("d", CASES[:4], ids=os.path.basename)
def test_gene_table_writer_matches_oracle_text(d):
    """report.gene_table_lines (vectorised strain cells, and the row-by-row path for ragged rows / large copy counts)
    prints what the oracle prints, given the oracle's numbers."""
    import types
    from oracle import abundance as oa
    from oracle import report as orp
    from strainflair_b200 import report
    files = [os.path.join(d, f) for f in ("graph.gfa", "reads.json", "dict_clusters.pickle")]
    g = gfa.load_graph(files[0], files[2])
    go = og.load_graph(files[0], files[2])
    orc = oa.AbundanceOracle(go, od.decode_json(files[1], go, 0.95))
    orc.run()
    rows = orp.gene_rows(orc)
    cols = ("nb_uniq_mapped_normalized", "nb_multimapped", "nb_multimapped_normalized", "mean_abund_uniq",
            "mean_abund_uniq_nz", "mean_abund_multiple", "mean_abund_multiple_nz", "ratio_covered_nodes")
    table = np.array([[float(r[k]) for k in cols] for r in rows], np.float64).reshape(len(rows), 8)
    touched = np.array([not isinstance(m, int) for m in orc.mult_reads], np.uint8)      # the reference's int 0 until touched
    res = types.SimpleNamespace(table=table, uniq_reads=np.array([int(u) for u in orc.uniq_reads], np.int64), mult_hits=touched)
    assert "".join(report.gene_table_lines(g, res)) == orp.gene_table_text(orc)
    # the same rows through the row-by-row writer (a copy count of 12 switches the vectorised cells off)
    big = int(g.pstr_cnt[0])
    g.pstr_cnt[0] = go["pstr_cnt"][0] = 12
    assert "".join(report.gene_table_lines(g, res)) == orp.gene_table_text(orc)
    g.pstr_cnt[0] = go["pstr_cnt"][0] = big


def test_on_disk_caches_hit_miss_and_invalidate(tmp_path, monkeypatch):
    """SFB200_CACHE_DIR: the parsed graph and the decoded reads come back array for array from their cache entries;
    a changed input, threshold or graph misses; without the variable nothing is written."""
    import shutil
    from strainflair_b200 import cache
    from strainflair_b200.schema import GRAPH_ARRAYS, READS_ARRAYS
    src = os.path.join(HERE, "golden", "case_011")
    work = tmp_path / "in"
    shutil.copytree(src, work)
    gf, js, pk = (str(work / n) for n in ("graph.gfa", "reads.json", "dict_clusters.pickle"))

    monkeypatch.delenv("SFB200_CACHE_DIR", raising=False)
    g0 = gfa.load_graph(gf, pk)
    r0 = decode.decode_json(js, g0, 0.95, want_cov=True)
    assert not r0.from_cache and sorted(os.listdir(work)) == ["dict_clusters.pickle", "expected.json", "graph.gfa", "reads.json"]

    cdir = tmp_path / "cache"
    monkeypatch.setenv("SFB200_CACHE_DIR", str(cdir))
    g1 = gfa.load_graph(gf, pk)                       # miss: parsed and stored
    r1 = decode.decode_json(js, g1, 0.95, want_cov=True)
    assert not r1.from_cache and len(os.listdir(cdir)) == 2
    g2 = gfa.load_graph(gf, pk)                       # hit
    r2 = decode.decode_json(js, g2, 0.95, want_cov=True)
    assert r2.from_cache
    for k in GRAPH_ARRAYS:
        assert np.array_equal(getattr(g0, k), getattr(g2, k)), k
    assert g2.strain_names == g0.strain_names and g2.header_strains == g0.header_strains and g2.path_names == g0.path_names
    for k in READS_ARRAYS:
        assert np.array_equal(getattr(r0, k), getattr(r2, k)), k
    assert np.array_equal(r0.aln_stats, r2.aln_stats) and np.array_equal(r0.walk_cov, r2.walk_cov)
    assert np.array_equal(r0.name_off, r2.name_off) and r0.name_blob == r2.name_blob
    assert decode.mate_name(r2, 0, 0) == decode.mate_name(r0, 0, 0)

    assert not decode.decode_json(js, g2, 0.5).from_cache          # another threshold: its own entry
    assert decode.decode_json(js, g2, 0.5).from_cache
    st = os.stat(js)
    os.utime(js, ns=(st.st_atime_ns, st.st_mtime_ns + 10**9))      # touched input: miss
    assert not decode.decode_json(js, g2, 0.95).from_cache
    g3 = gfa.load_graph(gf, pk)
    g3.node_len = g3.node_len + 1                                  # another graph: miss (the decoder reads node lengths)
    assert not decode.decode_json(js, g3, 0.95).from_cache
    # an unreadable entry is ignored, not fatal
    entry = cache._entry(gf, "graph")
    with open(entry, "wb") as fh:
        fh.write(b"not an npz")
    g4 = gfa.load_graph(gf, pk)
    assert np.array_equal(g4.path_nodes, g0.path_nodes)
    monkeypatch.setenv("SFB200_CACHE_DIR", ".")                    # "." = next to the input
    gfa.load_graph(gf, pk)
    assert any(n.endswith(".sfb200.npz") for n in os.listdir(work))


def test_group_permutation_keeps_every_result():
    """Reads.permute_groups / locality_order: the same groups in another order give the same counters, integer
    arrays and histograms, the same leaf code per group and the same fp64 sums (to rounding) in the oracle."""
    from oracle import abundance as oa
    from strainflair_b200 import synth
    graph, reads = synth.make_config("tiny")
    base = oa.AbundanceOracle(graph.as_dict(), reads.as_dict()).run()
    ident = reads.permute_groups(np.arange(reads.n_groups))
    for k in ("grp_nmates", "grp_aln_off", "mate_seq_len", "aln_walk_off", "aln_read_len", "aln_bins", "walk_node", "walk_alen"):
        assert np.array_equal(getattr(ident, k), getattr(reads, k)), k
    local = reads.locality_order()
    firsts = []
    for g in local:
        a0, a1 = int(reads.grp_aln_off[2 * g]), int(reads.grp_aln_off[2 * g + 2])
        firsts.append(int(reads.walk_node[reads.aln_walk_off[a0]]) if a1 > a0 else 1 << 62)
    assert firsts == sorted(firsts) and sorted(local.tolist()) == list(range(reads.n_groups))
    for order in (local, np.random.default_rng(3).permutation(reads.n_groups)):
        moved = reads.permute_groups(order)
        assert (moved.n_groups, moved.n_alns, moved.n_records) == (reads.n_groups, reads.n_alns, reads.n_records)
        got = oa.AbundanceOracle(graph.as_dict(), moved.as_dict()).run()
        assert got.C == base.C and got.uniq_reads == base.uniq_reads and np.array_equal(got.hist, base.hist)
        assert np.array_equal(base.grp_code[order], got.grp_code) and np.array_equal(base.grp_code2[order], got.grp_code2)
        for name in ("uniq_norm", "mult_reads", "mult_norm", "uniq_abund", "mult_abund"):
            a, b = np.asarray(getattr(base, name), np.float64), np.asarray(getattr(got, name), np.float64)
            assert np.allclose(a, b, rtol=1e-12, atol=0) and np.array_equal(a > 0, b > 0), name


def _tile_keys(reads, tiling):
    """numpy statement of sfb_tile_permute's keys: the tile holding the first and last node of every alignment, and (inside
    a tile) the first node of the group's first alignment with a walk."""
    keys = np.full(reads.n_groups, tiling.n_tiles, np.int64)
    first = np.zeros(reads.n_groups, np.int64)
    for g in range(reads.n_groups):
        a0, a1 = int(reads.grp_aln_off[2 * g]), int(reads.grp_aln_off[2 * g + 2])
        tiles, node = set(), None
        for a in range(a0, a1):
            w0, w1 = int(reads.aln_walk_off[a]), int(reads.aln_walk_off[a + 1])
            if w1 > w0:
                tiles.add(int(tiling.node_tile[reads.walk_node[w0]]))
                tiles.add(int(tiling.node_tile[reads.walk_node[w1 - 1]]))
                if node is None:
                    node = int(reads.walk_node[w0])
        if len(tiles) == 1 and -1 not in tiles:
            keys[g] = tiles.pop()
            first[g] = node
    return keys, first


def test_tiling_of_the_synthetic_graph():
    """tiles.Tiling: every tile is a closed part of the graph (consecutive paths, consecutive nodes, no path outside
    crosses one of its nodes) within the limits of the kernels; the bubble-chain generator gives one tile per cluster
    or per few small clusters."""
    from strainflair_b200 import synth
    from strainflair_b200.tiles import Tiling
    for name in ("tiny", "small"):
        graph, _ = synth.make_config(name, scale=0.01)
        tl = Tiling(graph)
        assert tl.n_tiles > 0 and tl.tuple_cap in (8, 16, 32)
        path_of_node = [set() for _ in range(graph.n_nodes)]
        for p in range(graph.n_paths):
            for n in graph.path_nodes[graph.path_off[p]:graph.path_off[p + 1]]:
                path_of_node[int(n)].add(p)
        seen_paths = 0
        for t, (p0, p1, n0, n1) in enumerate(tl.tile_desc[:, :4]):
            assert p0 < p1 and n0 < n1
            slots = int(graph.path_off[p1] - graph.path_off[p0]) + (p1 - p0)
            assert slots <= tl.max_slots <= 32767 and p1 - p0 <= tl.max_paths <= 255 and n1 - n0 <= tl.max_nodes
            for p in range(p0, p1):
                nodes = graph.path_nodes[graph.path_off[p]:graph.path_off[p + 1]]
                assert nodes.min() >= n0 and nodes.max() < n1
            for n in range(n0, n1):
                assert all(p0 <= p < p1 for p in path_of_node[n]) and tl.node_tile[n] == t
            seen_paths += p1 - p0
        assert seen_paths == graph.n_paths                      # the generator's clusters are all tileable
    # a graph whose two paths interleave their node ids is one closed part; with a small cap it has no tile at all
    from strainflair_b200.schema import Graph
    g = Graph(node_ids=np.arange(1, 9), node_len=np.full(8, 5, np.int32), path_off=np.array([0, 4, 8]),
              path_nodes=np.array([0, 2, 4, 6, 1, 3, 5, 7], np.int32), path_seq_len=np.array([20, 20]),
              pstr_off=np.array([0, 1, 2]), pstr_id=np.array([0, 1], np.int32), pstr_cnt=np.array([1, 1], np.int32),
              path_cluster=np.array([0, 1], np.int32), clu_off=np.array([0, 1, 2]), clu_paths=np.array([0, 1], np.int32),
              strain_names=["a", "b"], header_strains=[0, 1])
    one = Tiling(g)
    assert one.n_tiles == 1 and tuple(one.tile_desc[0]) == (0, 2, 0, 8, 1, 0, 0, 0)
    assert Tiling(g, slot_cap=6).n_tiles == 0 and (Tiling(g, slot_cap=6).node_tile == -1).all()


def test_tile_staging_and_file_order_restoration():
    """HostReads(tiling=...) stages the groups in tile order (native sfb_tile_permute == the numpy statement of its key
    + a stable sort), with the work items of the tile kernels, and engine.restore_file_order puts per-group /
    per-alignment outputs back in file order (the oracle stands in for the kernels here)."""
    from oracle import abundance as oa
    from strainflair_b200 import engine, synth
    from strainflair_b200.tiles import ITEM_GROUPS, Tiling
    graph, reads = synth.make_config("tiny")
    tl = Tiling(graph)
    host = engine.HostReads(reads, pin=False, wire=True, tiling=tl)
    plain = engine.HostReads(reads, pin=False, wire=True)
    assert plain.order is None and plain.items is None and host.order is not None and host.wire
    keys, first = _tile_keys(reads, tl)
    want = np.lexsort((first, keys))                 # by tile, inside a tile by first node, stable
    assert np.array_equal(host.order, want)
    moved, aln_index = reads.permute_groups(host.order, return_index=True)
    assert np.array_equal(host.reads.walk_node, moved.walk_node) and np.array_equal(host.aln_index, aln_index)
    assert host.grp_begin == int((keys < tl.n_tiles).sum()) and 0 < host.grp_begin < reads.n_groups
    assert host.aln_begin == int(moved.grp_aln_off[2 * host.grp_begin])
    # the work items cover the tile range exactly once, tile by tile, in pieces of at most ITEM_GROUPS groups
    covered = np.zeros(reads.n_groups, np.int32)
    for tile, g0, g1, _ in host.items:
        assert 0 < g1 - g0 <= ITEM_GROUPS and (keys[host.order[g0:g1]] == tile).all()
        covered[g0:g1] += 1
    assert (covered[:host.grp_begin] == 1).all() and (covered[host.grp_begin:] == 0).all()
    assert "tile_items" in host.offs
    base = oa.AbundanceOracle(graph.as_dict(), reads.as_dict()).run()
    got = oa.AbundanceOracle(graph.as_dict(), moved.as_dict()).run()
    res = engine.Result(grp_code=np.asarray(got.grp_code).copy(), grp_code2=np.asarray(got.grp_code2).copy(),
                        aln_assign=np.asarray(got.aln_assign).copy())
    assert not np.array_equal(res.grp_code, base.grp_code)
    engine.restore_file_order(res, host.order, host.aln_index)
    assert np.array_equal(res.grp_code, base.grp_code) and np.array_equal(res.grp_code2, base.grp_code2)
    assert np.array_equal(res.aln_assign, base.aln_assign)
    # empty and tile-free shards
    empty = engine.HostReads(reads.slice_groups(0, 0), pin=False, tiling=tl)
    assert empty.items is None and empty.grp_begin == 0


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_native_locality_permutation_equals_numpy(threads):
    """csrc/locality.cpp (stable radix sort of the groups + parallel copy) == Reads.locality_order + permute_groups,
    whatever the number of threads; groups without alignments, empty and one-group inputs included."""
    from strainflair_b200 import engine, synth
    graph, reads = synth.make_config("small", scale=0.3)
    cases = [reads, reads.slice_groups(0, 0), reads.slice_groups(7, 8), reads.slice_groups(100, 1500),
             synth.make_reads(graph, 5, n_pairs=3000, read_len=150, p_low_score=0.6, p_single=0.5, p_exc=0.05)]
    fields = ("grp_nmates", "grp_aln_off", "mate_seq_len", "aln_walk_off", "aln_read_len", "aln_bins", "walk_node",
              "walk_alen", "exc_cov")
    for rd in cases:
        moved, order, aln_index = engine.locality_permute(rd, threads=threads)
        want_order = rd.locality_order()
        want, want_index = rd.permute_groups(want_order, return_index=True)
        assert np.array_equal(order, want_order) and np.array_equal(aln_index, want_index)
        for k in fields:
            assert np.array_equal(getattr(moved, k), getattr(want, k)), k


@pytest.mark.parametrize("n_shards", [1, 3, 4, 9])
def test_stage_shards_cuts_tile_ordered_groups_at_tile_boundaries(n_shards):
    """engine.stage_shards: the data set in tile order first, then cut into shards at tile boundaries -- the shards
    together hold every group once (the same permuted arrays, end to end), a tile's groups sit in one shard, and each
    shard's work items and tile range describe its own slice."""
    from strainflair_b200 import engine, synth
    from strainflair_b200.tiles import Tiling
    graph, reads = synth.make_config("small", scale=0.3)
    tl = Tiling(graph)
    whole = engine.HostReads(reads, pin=False, wire=True, tiling=tl)
    shards = engine.stage_shards(reads, tl, n_shards, pin=False)
    assert len(shards) == n_shards and sum(s.n_groups for s in shards) == reads.n_groups
    assert np.array_equal(np.concatenate([s.reads.walk_node for s in shards]), whole.reads.walk_node)
    assert np.array_equal(np.concatenate([s.reads.grp_nmates for s in shards]), whole.reads.grp_nmates)
    owner, g_base = {}, 0
    for k, s in enumerate(shards):
        assert s.order is None                                  # (per-group outputs are not restored for staged shards)
        covered = np.zeros(s.n_groups, np.int32)
        for tile, g0, g1, _ in (s.items if s.items is not None else []):
            assert owner.setdefault(int(tile), k) == k, "a tile in two shards"
            covered[g0:g1] += 1
        assert (covered[:s.grp_begin] == 1).all() and (covered[s.grp_begin:] == 0).all()
        assert s.aln_begin == int(s.reads.grp_aln_off[2 * s.grp_begin])
        g_base += s.n_groups
    assert sum(s.grp_begin for s in shards) == whole.grp_begin
    # without a tiling: plain file-order slices
    flat = engine.stage_shards(reads, None, n_shards, pin=False)
    assert np.array_equal(np.concatenate([s.reads.walk_node for s in flat]), reads.walk_node)


def test_transfer_layout_first_nodes_wide_and_narrow_blocks():
    """First nodes travel as 16 bits above the smallest one of their block of 256 alignments; a block that spans more than
    65535 node ids lists its first nodes instead.  Node ids spread over millions, read groups in file order and sorted by
    first node: both kinds of block occur and expand to the same arrays."""
    from strainflair_b200 import engine, synth
    gkw, _ = synth.CONFIGS["tiny"]
    graph = synth.make_graph(21, **gkw)
    reads = synth.make_reads(graph, 22, n_pairs=4000, read_len=150)
    # the upper half of the node ids moved 70000 up: steps stay small, blocks of 256 alignments that hold first nodes of
    # both halves span more than 16 bits
    new_id = np.arange(graph.n_nodes, dtype=np.int64) + 70000 * (np.arange(graph.n_nodes, dtype=np.int64) >= graph.n_nodes // 2)
    node_len = np.ones(int(new_id[-1]) + 1, graph.node_len.dtype)
    node_len[new_id] = graph.node_len
    reads.walk_node[:] = new_id[reads.walk_node]
    seen = set()
    for rd in (reads, reads.permute_groups(reads.locality_order())):
        w, common = engine.wire_arrays(rd)
        wide_blocks = int((w["node0_anchor"] < 0).sum())
        seen.add("wide" if wide_blocks else "none")
        seen.add("narrow" if wide_blocks < len(w["node0_anchor"]) else "all")
        assert len(w["node0_wide"]) == 256 * wide_blocks
        grp_aln_off, off, seq, bins, node, word = _expand_on_host(w, common, node_len)
        assert np.array_equal(node, rd.walk_node) and np.array_equal(off, rd.aln_walk_off)
    assert {"wide", "narrow"} <= seen


@pytest.mark.parametrize("seed", range(12))
def test_transfer_layout_round_trip_random_shards(seed):
    """The packer / expansion pair on random shards: random group order (file order, locality order or shuffled), shard
    boundaries that fall anywhere (odd record offsets: the step nibbles of two shards' chunks share bytes), node ids
    spread so that some blocks of first nodes are narrow and some wide, steps beyond the nibble range."""
    from strainflair_b200 import engine, synth
    rng = np.random.default_rng(1000 + seed)
    gkw, _ = synth.CONFIGS["tiny"]
    graph = synth.make_graph(31 + seed, **gkw)
    reads = synth.make_reads(graph, 41 + seed, n_pairs=int(rng.integers(1, 1500)), read_len=150, p_exc=0.03)
    gap = int(rng.choice([0, 9, 70000]))                       # ids of the upper part of the graph moved up by `gap`
    cut = int(rng.integers(1, graph.n_nodes))
    new_id = np.arange(graph.n_nodes, dtype=np.int64) + gap * (np.arange(graph.n_nodes) >= cut)
    node_len = np.ones(int(new_id[-1]) + 1, graph.node_len.dtype)
    node_len[new_id] = graph.node_len
    reads.walk_node[:] = new_id[reads.walk_node]
    order = [np.arange(reads.n_groups), reads.locality_order(), rng.permutation(reads.n_groups)][seed % 3]
    rd = reads.permute_groups(order)
    g0 = int(rng.integers(0, rd.n_groups))
    g1 = int(rng.integers(g0, rd.n_groups + 1))
    for part in (rd, rd.slice_groups(g0, g1), rd.slice_groups(0, g0)):
        out = engine.wire_arrays(part)
        if out is None:                                        # (the layout does not pay for this shard: plain arrays)
            continue
        w, common = out
        grp_aln_off, off, seq, bins, node, word = _expand_on_host(w, common, node_len)
        assert np.array_equal(grp_aln_off, part.grp_aln_off) and np.array_equal(off, part.aln_walk_off)
        assert np.array_equal(seq, part.mate_seq_len) and np.array_equal(bins, part.aln_bins)
        assert np.array_equal(node, part.walk_node)
        assert np.array_equal(_coverage(word, w["exc_cov"]), _coverage(part.walk_alen, part.exc_cov))


def test_transfer_layout_is_independent_of_the_thread_count():
    """The packer works on chunks of 65536 alignments in parallel; step nibbles of neighbouring chunks share a byte.  More
    than one chunk: the same bytes whatever the number of threads, and the round trip holds."""
    from strainflair_b200 import engine, synth
    graph, reads = synth.make_config("small", scale=1.0)
    assert reads.n_alns > 65536
    ref = None
    for threads in (1, 3, 16):
        w, common = engine.wire_arrays(reads, threads=threads)
        if ref is None:
            ref = w
        else:
            assert all(np.array_equal(ref[k], w[k]) for k in ref), threads
    grp_aln_off, off, seq, bins, node, word = _expand_on_host(ref, common, graph.node_len)
    assert np.array_equal(node, reads.walk_node) and np.array_equal(off, reads.aln_walk_off)
