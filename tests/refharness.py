"""Runs the UNMODIFIED reference (imported by file path from /root/reference) and
collects everything the parity tests compare.  Only usable where the reference is
mounted (the build container); the GPU box replays the committed golden fixtures.
"""
import contextlib
import importlib.util
import io
import os
import pickle
import sys

REF_DIR = "/root/reference/strainflair"


def have_reference():
    return os.path.isfile(os.path.join(REF_DIR, "json2csv.py"))


_mods = {}


def _load(name):
    if name not in _mods:
        spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF_DIR, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _mods[name] = mod
    return _mods[name]


COUNTERS = ("NB_READS", "PAIRS", "SINGLES", "FILTERED", "UNASSIGNED", "ASSIGNED_U", "ASSIGNED_M",
            "INCONSIST", "INCONSIST_M", "MULTI_ALIGN", "SE", "SAME_CP", "DIFF_CP", "ONE_NOALIGN",
            "ONE_NOCP", "AMBIG_CP_RESOLVED", "AMBIG_RESOLVED", "AMBIG_REDUCED", "SE_M", "SECOND_PASS")


def run_json2csv(gfa, js, pk, thr=0.95, outdir=None):
    """Drive the reference's library-level entry points in the order of json2csv_main
    (json2csv.py:1326-1344) and return a dict of raw results (+ the files it wrote)."""
    ref = _load("json2csv")
    ref.update_progress = lambda p: None
    pg = ref.Pangenome()
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        pg.build_pangenome(gfa)
        pg.fill_clusters_info(pk)
        ref.parse_vgmpmap(js, pg, thr, False)
    res = dict(
        counters={k: getattr(ref, k) for k in COUNTERS},
        paths=[dict(node_ids=list(p.node_ids), cluster_id=p.cluster_id, strain_ids=dict(p.strain_ids),
                    uniq_reads=p.total_mapped_unique_reads, uniq_norm=p.total_mapped_unique_reads_normalized,
                    mult_reads=p.total_mapped_mult_reads, mult_norm=p.total_mapped_mult_reads_normalized,
                    uniq_abund=list(p.unique_mapped_abundances), mult_abund=list(p.multiple_mapped_abundances),
                    seq_len=pg.get_sequence_length(p)) for p in pg.paths],
        species=list(pg.species_names),
        hist={n: {s: [d[i / 100] for i in range(101)] for s, d in getattr(pg, n).items()}
              for n in ("mappingscore_freq", "identityscore_freq", "subst_freq", "indel_freq", "errors_freq")},
        clusters=[dict(colored_paths_id=list(c.colored_paths_id), reads_name=list(c.reads_name),
                       unassigned_paths=[list(map(int, w)) for w in c.unassigned_paths],
                       unassigned_abund=dict(c.unassigned_abund), count=c.unassigned_count,
                       count_norm=c.unassigned_count_norm, afterfilt=c.unassigned_count_afterfilt,
                       before=c.estimated_min_strains_before, after=c.estimated_min_strains_after,
                       abund=c.estimated_abund) for c in pg.clusters],
    )
    if outdir is not None:
        os.makedirs(outdir, exist_ok=True)
        prefix = os.path.join(outdir, "ref")
        with contextlib.redirect_stdout(sink):
            pg.print_gene_table(prefix + "_gene_table.csv")
            for n in ("mappingscore_freq", "identityscore_freq", "subst_freq", "indel_freq", "errors_freq"):
                pg.print_error_distribution(prefix, n)
            pg.print_unassigned_info(prefix + "_unassigned.csv")
        res["prefix"] = prefix
    return res


def run_json2csv_cli(argv, cwd):
    """Call the reference's json2csv_main() with sys.argv = argv inside ``cwd``."""
    ref = _load("json2csv")
    ref.update_progress = lambda p: None
    old_argv, old_cwd = sys.argv, os.getcwd()
    sink = io.StringIO()
    try:
        sys.argv = ["json2csv"] + list(argv)
        os.chdir(cwd)
        with contextlib.redirect_stdout(sink):
            ref.json2csv_main()
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)
    return sink.getvalue()


def run_strains(gene_table, out_stem, thr=None):
    """Reference compute_strains_abundance_main() with the pandas>=2 shim (Q13)."""
    import pandas as pd
    if not hasattr(pd.DataFrame, "iteritems"):
        pd.DataFrame.iteritems = pd.DataFrame.items
    ref = _load("compute_strains_abundance")
    old = sys.argv
    sink = io.StringIO()
    try:
        sys.argv = ["compute_strains_abundance", "-i", gene_table, "-o", out_stem] + \
                   (["-t", str(thr)] if thr is not None else [])
        import warnings
        with contextlib.redirect_stdout(sink), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref.compute_strains_abundance_main()
    finally:
        sys.argv = old
    return out_stem + ".csv"
