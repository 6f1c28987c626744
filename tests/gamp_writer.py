"""Test infrastructure: writes GAMP files (vg's stream framing around MultipathAlignment protobuf messages) from the
`vg view -j -K` JSON rendering, with Google's protobuf library and a schema restated from vg.proto -- an encoder that
shares no code with the decoder under test (csrc/decode.cpp, parse_pb / sfb_decode_gamp)."""
import gzip
import json

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_T = descriptor_pb2.FieldDescriptorProto


def _schema():
    f = descriptor_pb2.FileDescriptorProto(name="vg_subset.proto", package="vg", syntax="proto3")

    def msg(name, fields):
        m = f.message_type.add(name=name)
        for fname, num, typ, rep, tname in fields:
            fd = m.field.add(name=fname, number=num, type=typ, label=_T.LABEL_REPEATED if rep else _T.LABEL_OPTIONAL)
            if tname:
                fd.type_name = ".vg." + tname
    msg("Position", [("node_id", 1, _T.TYPE_INT64, False, None), ("offset", 2, _T.TYPE_INT64, False, None),
                     ("is_reverse", 4, _T.TYPE_BOOL, False, None)])
    msg("Edit", [("from_length", 1, _T.TYPE_INT32, False, None), ("to_length", 2, _T.TYPE_INT32, False, None),
                 ("sequence", 3, _T.TYPE_STRING, False, None)])
    msg("Mapping", [("position", 1, _T.TYPE_MESSAGE, False, "Position"), ("edit", 2, _T.TYPE_MESSAGE, True, "Edit"),
                    ("rank", 5, _T.TYPE_INT64, False, None)])
    msg("Path", [("name", 1, _T.TYPE_STRING, False, None), ("mapping", 2, _T.TYPE_MESSAGE, True, "Mapping")])
    msg("Subpath", [("path", 1, _T.TYPE_MESSAGE, False, "Path"), ("next", 2, _T.TYPE_UINT32, True, None),
                    ("score", 3, _T.TYPE_INT32, False, None)])
    msg("MultipathAlignment", [("sequence", 1, _T.TYPE_STRING, False, None), ("quality", 2, _T.TYPE_BYTES, False, None),
                               ("name", 3, _T.TYPE_STRING, False, None), ("subpath", 6, _T.TYPE_MESSAGE, True, "Subpath"),
                               ("mapping_quality", 7, _T.TYPE_INT32, False, None), ("start", 8, _T.TYPE_UINT32, True, None),
                               ("paired_read_name", 9, _T.TYPE_STRING, False, None)])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(f)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("vg.MultipathAlignment"))


MultipathAlignment = _schema()


def _varint(n):
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def message_of(line: dict):
    m = MultipathAlignment()
    for k in ("sequence", "name", "paired_read_name"):
        if k in line:
            setattr(m, k, line[k])
    if "mapping_quality" in line:
        m.mapping_quality = int(line["mapping_quality"])
    m.start.extend(int(x) for x in line.get("start", []))
    for sp in line.get("subpath", []):
        s = m.subpath.add()
        if "score" in sp:
            s.score = int(sp["score"])
        s.next.extend(int(x) for x in sp.get("next", []))
        if "path" in sp:
            s.path.SetInParent()
            for mp in sp["path"].get("mapping", []):
                mm = s.path.mapping.add()
                if "position" in mp:
                    mm.position.SetInParent()
                    if "node_id" in mp["position"]:
                        mm.position.node_id = int(mp["position"]["node_id"])
                    if mp["position"].get("is_reverse"):
                        mm.position.is_reverse = True
                    if "offset" in mp["position"]:
                        mm.position.offset = int(mp["position"]["offset"])
                for ed in mp.get("edit", []):
                    e = mm.edit.add()
                    e.from_length, e.to_length = int(ed.get("from_length", 0)), int(ed.get("to_length", 0))
                    if ed.get("sequence"):
                        e.sequence = ed["sequence"]
    return m


def write_gamp(json_path, gamp_path, group_size=1000, tagged=True, compress="bgzf"):
    """group_size messages per group; ``tagged``: every group starts with the "MGAM" type tag; ``compress``: None,
    "gzip" (one member) or "bgzf" (many members, like vg's BGZF output)."""
    msgs = [message_of(json.loads(ln)).SerializeToString() for ln in open(json_path) if ln.strip()]
    groups = []
    for i in range(0, len(msgs), group_size):
        part = ([b"MGAM"] if tagged else []) + msgs[i:i + group_size]
        groups.append(_varint(len(part)) + b"".join(_varint(len(x)) + x for x in part))
    raw = b"".join(groups)
    if compress is None:
        blob = raw
    elif compress == "gzip":
        blob = gzip.compress(raw)
    else:
        blob = b"".join(gzip.compress(raw[i:i + 65280]) for i in range(0, len(raw), 65280)) + gzip.compress(b"")
    with open(gamp_path, "wb") as fh:
        fh.write(blob)
    return len(msgs)
