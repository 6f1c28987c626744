"""CPU-only: the C-ABI library builds/loads, exports every symbol the header declares, and the Python
mirrors of the header's enums agree with the header and with the oracle."""
import ctypes
import os
import re

import pytest

from oracle import abundance as oa
from strainflair_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "sfb200.h")).read()


def declared_functions():
    return sorted(set(re.findall(r"^\s*(?:int|const char\*|sfb_decoded\*|sfb_gfa\*|sfb_packed\*|int64_t|void)\s+(sfb_[a-z0-9_]+)\s*\(", HEADER, re.M)))


def test_header_declares_expected_entry_points():
    assert declared_functions() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert lib.sfb_abi_version() == _lib.ABI_VERSION
    assert b"sm_100a" in lib.sfb_version()


def _enum_block(first):
    m = re.search(r"enum\s*\{([^}]*" + first + r"[^}]*)\}", HEADER)
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    names, val = {}, -1
    for item in body.split(","):
        item = item.strip()
        if not item:
            continue
        if "=" in item:
            k, v = item.split("=")
            val = int(v)
            names[k.strip()] = val
        else:
            val += 1
            names[item] = val
    return names


def test_counter_enum_matches_python_and_oracle():
    e = _enum_block("SFB_C_NB_READS")
    assert e["SFB_N_COUNTERS"] == _lib.N_COUNTERS
    for i, name in enumerate(_lib.COUNTER_NAMES):
        assert e["SFB_C_" + name] == i
    assert tuple(_lib.COUNTER_NAMES[:len(oa.COUNTERS)]) == oa.COUNTERS


def test_leaf_codes_match_oracle():
    e = _enum_block("SFB_ABSENT")
    for name in ("ABSENT", "FILTERED", "UNASSIGNED", "UNASSIGNED_BOOK", "DEFER", "ASSIGN_SAME", "ASSIGN_DIFF",
                 "MULTI_ALIGN", "INCONSIST", "SE_COUNTED"):
        assert e["SFB_" + name] == getattr(oa, name) == getattr(_lib, name)
    e2 = _enum_block("SFB_P2_NONE")
    for name in ("NONE", "ASSIGNED", "MULTI_ALIGN", "SE_ASSIGNED", "UNASSIGNED", "INCONSIST"):
        assert e2["SFB_P2_" + name] == getattr(oa, "P2_" + name)


def test_struct_sizes_are_pointer_and_int64_multiples():
    for st in (_lib.SfbGraph, _lib.SfbReads, _lib.SfbWork, _lib.SfbAccum, _lib.SfbTiles):
        assert ctypes.sizeof(st) % 8 == 0
    assert ctypes.sizeof(_lib.SfbGraph) == 8 * 21
    assert ctypes.sizeof(_lib.SfbReads) == 8 * 12
    assert ctypes.sizeof(_lib.SfbWire) == 8 * 28
    assert ctypes.sizeof(_lib.SfbWork) == 8 * 15
    assert ctypes.sizeof(_lib.SfbTiles) == 8 * 13
    assert ctypes.sizeof(_lib.SfbAccum) == 8 * 10


def test_argument_validation_needs_no_gpu():
    """Null/negative arguments are rejected before any CUDA call."""
    lib = _lib.load()
    g = _lib.SfbGraph()
    r = _lib.SfbReads()
    w = _lib.SfbWork()
    a = _lib.SfbAccum()
    rc = lib.sfb_match(ctypes.byref(g), ctypes.byref(r), ctypes.byref(w), ctypes.byref(a), None)
    assert rc == -1
    assert lib.sfb_last_error().startswith(b"check_graph")
    g.mask_words = 1
    rc = lib.sfb_match(ctypes.byref(g), ctypes.byref(r), ctypes.byref(w), ctypes.byref(a), None)
    assert rc == -1 and b"null" in lib.sfb_last_error()
    with pytest.raises(_lib.SfbError):
        _lib.check(rc)
    # tiles: limits of the local ids, tuple capacity, shared-memory size
    t = _lib.SfbTiles(n_tiles=1, n_items=0, max_slots=1359, max_nodes=472, max_paths=9, max_edges=451, tuple_cap=8)
    assert 40_000 < lib.sfb_tile_smem_bytes(ctypes.byref(t)) < 100_000
    t.tuple_cap = 12
    assert lib.sfb_tile_smem_bytes(ctypes.byref(t)) == -1 and b"tuple_cap" in lib.sfb_last_error()
    t.tuple_cap, t.max_slots = 8, 40_000
    assert lib.sfb_tile_smem_bytes(ctypes.byref(t)) == -1 and b"limits" in lib.sfb_last_error()


def test_engine_refuses_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from strainflair_b200 import engine, synth
    g, _ = synth.make_config("tiny")
    with pytest.raises(_lib.SfbError):
        engine.AbundanceEngine(g)
