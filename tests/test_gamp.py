"""Direct GAMP input (StrainFLAIR.sh:582 writes the file, :592-602 is the `vg view -j -K`
hop this removes).  PARITY UNPINNED: no vg binary, schema or GAMP file is available here, so the container and the
message layout are restated from vg's published format and the tests can only close the loop on themselves -- GAMP
files are written from the golden JSON inputs by tests/gamp_writer.py (Google's protobuf library, a schema restated
independently of the decoder) and decode(GAMP) must equal decode(JSON) array for array."""
import glob
import json
import os

import numpy as np
import pytest

import gamp_writer
from oracle import graph as og
from strainflair_b200 import decode
from strainflair_b200.schema import Graph

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = sorted(glob.glob(os.path.join(HERE, "golden", "case_*")))
FIELDS = ("grp_nmates", "grp_aln_off", "mate_seq_len", "aln_walk_off", "aln_read_len", "aln_bins", "walk_node", "walk_alen",
          "exc_cov", "name_off", "aln_stats")


def same_reads(a, b):
    for k in FIELDS:
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert a.name_blob == b.name_blob


def load(d):
    with open(os.path.join(d, "expected.json")) as fh:
        thr = json.load(fh)["params"]["thr"]
    g = Graph.from_dict(og.load_graph(os.path.join(d, "graph.gfa"), os.path.join(d, "dict_clusters.pickle")))
    return g, thr, os.path.join(d, "reads.json")


@pytest.mark.parametrize("d", CASES, ids=os.path.basename)
def test_gamp_decodes_like_its_json_rendering(d, tmp_path):
    g, thr, js = load(d)
    want = decode.decode_json(js, g, thr, want_cov=True)
    path = str(tmp_path / "reads.gamp")
    for kw in (dict(), dict(tagged=False, compress=None), dict(group_size=1, compress="gzip"), dict(group_size=7)):
        n = gamp_writer.write_gamp(js, path, **kw)
        assert n > 0 and decode.is_gamp(path) and not decode.is_gamp(js)
        for threads, block in ((1, 0), (4, 0), (3, 5)):          # tiny blocks: read groups cut by a block, blocks widened
            got = decode.decode_gamp(path, g, thr, threads=threads, block_records=block, want_cov=True)
            same_reads(got, want)
            assert np.array_equal(got.walk_cov, want.walk_cov)
    same_reads(decode.decode_alignments(path, g, thr), want)


def test_gamp_names_errors_and_foreign_streams(tmp_path):
    g, thr, js = load(CASES[0])
    lines = [json.loads(ln) for ln in open(js)]
    # a backslash and a quote in read names: JSON escapes them, protobuf carries them raw
    odd = {}
    for ln in lines:
        for k in ("name", "paired_read_name"):
            if k in ln:
                ln[k] = odd.setdefault(ln[k], ln[k].replace("r", 'r\\"x', 1))
    js2 = str(tmp_path / "odd.json")
    with open(js2, "w") as fh:
        for ln in lines:
            fh.write(json.dumps(ln) + "\n")
    path = str(tmp_path / "odd.gamp")
    gamp_writer.write_gamp(js2, path)
    a, b = decode.decode_json(js2, g, thr), decode.decode_gamp(path, g, thr)
    same_reads(a, b)
    assert b'\\"x' in b.name_blob
    # the reference's exceptions, from the binary form too: unknown node -> KeyError, empty read -> ZeroDivisionError
    bad = [dict(ln) for ln in lines]
    k = next(i for i, ln in enumerate(bad) if '"node_id"' in json.dumps(ln))
    text = json.dumps(bad[k])
    first = text.index('"node_id": ') + len('"node_id": ')
    bad[k] = json.loads(text[:first] + '"99999"' + text[text.index(",", first) if "," in text[first:first + 12] else text.index("}", first):])
    with open(js2, "w") as fh:
        for ln in bad:
            fh.write(json.dumps(ln) + "\n")
    gamp_writer.write_gamp(js2, path)
    with pytest.raises(KeyError):
        decode.decode_json(js2, g, thr)
    with pytest.raises(KeyError):
        decode.decode_gamp(path, g, thr)
    # a stream of another type, and a truncated file
    raw = gamp_writer._varint(2) + gamp_writer._varint(3) + b"GAM" + gamp_writer._varint(1) + b"\x00"
    open(path, "wb").write(raw)
    with pytest.raises(ValueError, match="type tag 'GAM'"):
        decode.decode_gamp(path, g, thr)
    gamp_writer.write_gamp(js, path, compress=None)
    blob = open(path, "rb").read()
    open(path, "wb").write(blob[:len(blob) // 2])
    with pytest.raises(ValueError):
        decode.decode_gamp(path, g, thr)
    open(path, "wb").write(b"")
    assert decode.decode_gamp(path, g, thr).n_groups == 0


def test_json2csv_command_line_takes_the_gamp_file(tmp_path, monkeypatch):
    """`json2csv -m reads.gamp`: the decoded arrays (all the command line feeds to the engine) equal those of the JSON."""
    g, thr, js = load(CASES[-1])
    path = str(tmp_path / "reads.gamp")
    gamp_writer.write_gamp(js, path)
    same_reads(decode.decode_alignments(path, g, thr), decode.decode_alignments(js, g, thr))
