"""ctypes binding of libsfb200.so (include/sfb200.h).  There is no CPU fallback: if the
library is missing or cannot be built, importing this module's users fails loudly."""
import ctypes as C
import os

from . import build as _build

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_u8p = C.POINTER(C.c_uint8)
c_u32p = C.POINTER(C.c_uint32)
c_u64p = C.POINTER(C.c_uint64)
c_f64p = C.POINTER(C.c_double)


# arrays of the transfer layout, in struct order
WIRE_DTYPES = ("u1", "u1", "u2", "i8", "i4", "u1", "u1", "u1", "i8", "u1", "u2", "u1", "u1", "u1", "i8", "f8", "i8", "i4", "i4", "i4")
WIRE_ARRAYS = ("grp_nmates", "mate_naln", "mate_seq_len", "seq_exc_idx", "seq_exc_val", "aln_len", "aln_code", "bins_dict",
               "bins_exc_idx", "bins_exc_val", "aln_node0", "aln_al_first", "aln_al_last", "walk_step", "exc_idx", "exc_cov",
               "wide_idx", "wide_node", "node0_anchor", "node0_wide")


class SfbGraph(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int64), ("n_paths", C.c_int64), ("n_slots", C.c_int64), ("n_strains", C.c_int64),
        ("mask_words", C.c_int64),
        ("slot_node", C.c_void_p), ("path_off", C.c_void_p),
        ("occ_off", C.c_void_p), ("occ_pair", C.c_void_p), ("blk_path", C.c_void_p), ("path_seq_len", C.c_void_p),
        ("path_nstrain", C.c_void_p), ("path_strain0", C.c_void_p), ("path_smask", C.c_void_p),
        ("path_cluster", C.c_void_p),
        ("node_step", C.c_void_p), ("node_mask", C.c_void_p), ("step_off", C.c_void_p), ("step_succ", C.c_void_p),
        ("step_mask", C.c_void_p), ("steps_partial", C.c_int64),
    ]


class SfbReads(C.Structure):
    _fields_ = [
        ("n_groups", C.c_int64), ("n_alns", C.c_int64), ("n_records", C.c_int64), ("n_exc", C.c_int64),
        ("grp_nmates", C.c_void_p), ("grp_aln_off", C.c_void_p), ("mate_seq_len", C.c_void_p),
        ("aln_walk_off", C.c_void_p), ("aln_bins", C.c_void_p), ("walk_node", C.c_void_p),
        ("walk_alen", C.c_void_p), ("exc_cov", C.c_void_p),
    ]


class SfbWire(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("n_groups", "n_alns", "n_records", "n_exc", "n_wide", "n_seq_exc", "n_bins_exc",
                                         "seq_len_common")] + \
               [(k, C.c_void_p) for k in WIRE_ARRAYS]


class SfbWork(C.Structure):
    _fields_ = [
        ("aln_match", C.c_void_p), ("aln_sets", C.c_void_p), ("match_buf", C.c_void_p), ("match_cap", C.c_int64), ("cursor", C.c_void_p),
        ("grp_code", C.c_void_p), ("grp_code2", C.c_void_p), ("aln_assign", C.c_void_p), ("deferred", C.c_void_p),
        ("slow", C.c_void_p), ("app_buf", C.c_void_p), ("app_cap", C.c_int64),
        ("p2_begin", C.c_int64), ("grp_begin", C.c_int64), ("aln_begin", C.c_int64),
    ]


class SfbDims(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("n_groups", "n_alns", "n_records", "n_paths", "n_slots", "n_strains", "deterministic")]


WS_EXPANDED, WS_MATCH_PAIRS, WS_APP_ITEMS, WS_ACCUM_F64, WS_ACCUM_I64 = range(5)


def workspace(what, **dims):
    """sfb_workspace_bytes: the buffer sizes come from the library, not from formulas restated here."""
    d = SfbDims(**{k: int(v) for k, v in dims.items()})
    n = int(load().sfb_workspace_bytes(int(what), C.byref(d)))
    if n < 0:
        raise SfbError(f"sfb_workspace_bytes({what}): {load().sfb_last_error().decode()}")
    return n


class SfbTiles(C.Structure):
    _fields_ = [
        ("n_tiles", C.c_int64), ("n_items", C.c_int64), ("max_slots", C.c_int64), ("max_nodes", C.c_int64),
        ("max_paths", C.c_int64), ("max_edges", C.c_int64), ("tuple_cap", C.c_int64), ("tile_desc", C.c_void_p),
        ("items", C.c_void_p), ("edge_off", C.c_void_p), ("edge_succ", C.c_void_p), ("edge_mask", C.c_void_p),
        ("node_mask", C.c_void_p),
    ]


CURSOR_N = 16          # entries of sfb_work.cursor


class SfbAccum(C.Structure):
    _fields_ = [
        ("uniq_reads", C.c_void_p), ("uniq_norm", C.c_void_p), ("mult_reads", C.c_void_p), ("mult_norm", C.c_void_p),
        ("mult_hits", C.c_void_p), ("uniq_abund", C.c_void_p), ("mult_abund", C.c_void_p), ("hist", C.c_void_p),
        ("counters", C.c_void_p), ("det_stride", C.c_int64),
    ]


# mirrors of the enums in include/sfb200.h (checked against the header in tests/test_abi.py)
COUNTER_NAMES = (
    "NB_READS", "PAIRS", "SINGLES", "FILTERED", "UNASSIGNED", "ASSIGNED_U", "ASSIGNED_M", "INCONSIST", "INCONSIST_M",
    "MULTI_ALIGN", "SE", "SAME_CP", "DIFF_CP", "ONE_NOALIGN", "ONE_NOCP", "AMBIG_CP_RESOLVED", "AMBIG_RESOLVED",
    "AMBIG_REDUCED", "SE_M", "SECOND_PASS", "BAD_BIN", "MATCH_OVERFLOW", "NO_PATH", "NO_CLUSTER",
    "ST_CAND", "ST_COMPARES", "ST_TUPLES", "ST_P1_RECORDS", "ST_P2_APPLIED", "ST_P2_RECORDS", "ST_HIST",
)
N_COUNTERS = 32
N_HIST, N_BINS, TABLE_COLS = 5, 101, 8
SLOT_MASK, MATCH_REV, MATCH_MULTI, MATCH_NONE, MATCH_SETS = 0x3FFFFFFF, 0x40000000, 0x80000000, 0xFFFFFFFF, 0xFFFFFFFE
MAX_SLOTS = 0x3FFFFFFF
(ABSENT, FILTERED, UNASSIGNED, UNASSIGNED_BOOK, DEFER, ASSIGN_SAME, ASSIGN_DIFF, MULTI_ALIGN, INCONSIST,
 SE_COUNTED) = range(10)
ABI_VERSION = 14

EXPORTS = ("sfb_abi_version", "sfb_version", "sfb_last_error", "sfb_expand_reads", "sfb_match", "sfb_match_steps", "sfb_classify", "sfb_pass1_scatter",
           "sfb_pass2", "sfb_pass2_resolve", "sfb_pass2_scatter", "sfb_accum_collapse", "sfb_path_reduce", "sfb_cluster_strain_counts", "sfb_strain_aggregate",
           "sfb_tile_pass1", "sfb_tile_pass2", "sfb_tile_scatter1", "sfb_tile_share", "sfb_tile_smem_bytes", "sfb_tile_permute", "sfb_workspace_bytes",
           "sfb_decode_json", "sfb_decode_gamp", "sfb_decoded_error", "sfb_decoded_size", "sfb_decoded_copy", "sfb_decoded_free",
           "sfb_min_strains", "sfb_wire_pack", "sfb_packed_status", "sfb_packed_size", "sfb_packed_copy", "sfb_packed_free",
           "sfb_load_gfa", "sfb_gfa_error", "sfb_gfa_size", "sfb_gfa_copy", "sfb_gfa_free", "sfb_locality_permute", "sfb_drop_sentinels")


class SfbError(RuntimeError):
    pass


_lib = None


def lib_path():
    return os.environ.get("SFB200_LIB", _build.LIB)


def load(build_if_missing=True):
    """Returns the loaded library.  Builds it in-tree with nvcc when absent/stale and a compiler exists."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if build_if_missing and "SFB200_LIB" not in os.environ:
        try:
            _build.build()
        except (FileNotFoundError, PermissionError) as e:     # no nvcc on this machine: use the prebuilt .so if there is one
            if not os.path.exists(path):
                raise SfbError(f"libsfb200.so is missing and could not be built: {e}") from e
        # (a compile error is not caught: a stale library must never stand in for sources that do not build)
    if not os.path.exists(path):
        raise SfbError(f"{path} not found; run `python -m strainflair_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(path)
    lib.sfb_abi_version.restype = C.c_int
    lib.sfb_version.restype = C.c_char_p
    lib.sfb_last_error.restype = C.c_char_p
    G, R, W, A = C.POINTER(SfbGraph), C.POINTER(SfbReads), C.POINTER(SfbWork), C.POINTER(SfbAccum)
    vp, i64, f64 = C.c_void_p, C.c_int64, C.c_double
    lib.sfb_expand_reads.argtypes = [C.POINTER(SfbWire), vp, R, vp, vp]
    lib.sfb_match.argtypes = [G, R, W, A, vp]
    lib.sfb_match_steps.argtypes = [G, R, W, A, vp]
    lib.sfb_classify.argtypes = [G, R, W, A, vp]
    lib.sfb_pass1_scatter.argtypes = [G, R, W, A, vp]
    lib.sfb_pass2.argtypes = [G, R, W, A, vp, vp]
    lib.sfb_pass2_resolve.argtypes = [G, R, W, A, vp, vp]
    lib.sfb_pass2_scatter.argtypes = [G, R, W, A, vp]
    lib.sfb_accum_collapse.argtypes = [vp, i64, i64, vp]
    lib.sfb_path_reduce.argtypes = [G, A, i64, i64, vp, vp]
    lib.sfb_drop_sentinels.argtypes = [G, vp, vp, vp]
    lib.sfb_drop_sentinels.restype = C.c_int
    lib.sfb_cluster_strain_counts.argtypes = [i64, i64, vp, vp, vp, vp, vp, vp, vp]
    lib.sfb_strain_aggregate.argtypes = [i64, i64, vp, vp, vp, vp, vp, vp, f64, vp, vp, vp, vp, vp]
    for name in EXPORTS[3:15]:
        getattr(lib, name).restype = C.c_int
    TL = C.POINTER(SfbTiles)
    lib.sfb_tile_pass1.argtypes = [G, TL, R, W, A, vp]
    lib.sfb_tile_pass1.restype = C.c_int
    lib.sfb_tile_pass2.argtypes = [G, TL, R, W, A, vp, vp]
    lib.sfb_tile_pass2.restype = C.c_int
    lib.sfb_tile_scatter1.argtypes = [G, TL, R, W, A, vp]
    lib.sfb_tile_scatter1.restype = C.c_int
    lib.sfb_tile_share.argtypes = [G, TL, R, W, A, vp, vp]
    lib.sfb_tile_share.restype = C.c_int
    lib.sfb_tile_smem_bytes.argtypes = [TL]
    lib.sfb_tile_smem_bytes.restype = C.c_int64
    lib.sfb_tile_permute.argtypes = [R, vp, i64, C.c_int, vp, vp, R, vp]
    lib.sfb_tile_permute.restype = C.c_int
    lib.sfb_workspace_bytes.argtypes = [C.c_int, C.POINTER(SfbDims)]
    lib.sfb_workspace_bytes.restype = C.c_int64
    lib.sfb_min_strains.argtypes = [i64, vp, vp, vp, C.c_int, vp, vp, vp, vp]
    lib.sfb_min_strains.restype = C.c_int
    lib.sfb_locality_permute.argtypes = [R, C.c_int, vp, vp, R]
    lib.sfb_locality_permute.restype = C.c_int
    lib.sfb_wire_pack.argtypes = [R, C.c_int]
    lib.sfb_wire_pack.restype = C.c_void_p
    lib.sfb_packed_status.argtypes = [vp, C.POINTER(C.c_char_p)]
    lib.sfb_packed_status.restype = C.c_int
    lib.sfb_packed_size.argtypes = [vp, C.c_int]
    lib.sfb_packed_size.restype = C.c_int64
    lib.sfb_packed_copy.argtypes = [vp, C.c_int, vp]
    lib.sfb_packed_copy.restype = C.c_int
    lib.sfb_packed_free.argtypes = [vp]
    lib.sfb_packed_free.restype = None
    lib.sfb_decode_json.argtypes = [C.c_char_p, vp, vp, i64, f64, C.c_int, i64]
    lib.sfb_decode_json.restype = C.c_void_p
    lib.sfb_decode_gamp.argtypes = [C.c_char_p, vp, vp, i64, f64, C.c_int, i64]
    lib.sfb_decode_gamp.restype = C.c_void_p
    lib.sfb_decoded_error.argtypes = [vp, C.POINTER(C.c_char_p)]
    lib.sfb_decoded_error.restype = C.c_int
    lib.sfb_decoded_size.argtypes = [vp, C.c_int]
    lib.sfb_decoded_size.restype = C.c_int64
    lib.sfb_decoded_copy.argtypes = [vp, C.c_int, vp]
    lib.sfb_decoded_copy.restype = C.c_int
    lib.sfb_decoded_free.argtypes = [vp]
    lib.sfb_decoded_free.restype = None
    lib.sfb_load_gfa.argtypes = [C.c_char_p]
    lib.sfb_load_gfa.restype = C.c_void_p
    lib.sfb_gfa_error.argtypes = [vp, C.POINTER(C.c_char_p)]
    lib.sfb_gfa_error.restype = C.c_int
    lib.sfb_gfa_size.argtypes = [vp, C.c_int]
    lib.sfb_gfa_size.restype = C.c_int64
    lib.sfb_gfa_copy.argtypes = [vp, C.c_int, vp]
    lib.sfb_gfa_copy.restype = C.c_int
    lib.sfb_gfa_free.argtypes = [vp]
    lib.sfb_gfa_free.restype = None
    if lib.sfb_abi_version() != ABI_VERSION:
        raise SfbError(f"ABI mismatch: library {lib.sfb_abi_version()}, binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise SfbError(f"libsfb200 error {rc}: {load().sfb_last_error().decode()}")
