"""Drop-in command lines: ``json2csv`` and ``compute_strains_abundance`` with the reference's flags, file names,
stdout lines and exit codes (json2csv.py:1281-1356, compute_strains_abundance.py:9-72)."""
import getopt
import sys
import time


def _usage_json2csv():
    print(f"Usage: python {sys.argv[0]} -g graph_file_name (gfa) -m mapped_file_name (json) -p dictionary_file_name "
          f"(pickle) -t alignment_score_threshold -o prefix_output_files_name -f")


def update_progress(progress):
    """The reference's console progress bar on stderr (json2csv.py:16-35), same text: its callers tick it through the
    GFA (:245,298), the two passes over the alignments (:819,1055 / :1074,1275) and the gene table (:411,474)."""
    bar, status = 50, ""
    if progress < 0:
        progress, status = 0, "Halt...\r\n"
    if progress >= 1:
        progress, status = 1, "Done...\r\n"
    block = int(round(bar * progress))
    sys.stderr.write("\rPercent: [{0}] {1}% {2}".format("#" * block + "-" * (bar - block), round(progress * 100, 2), status))
    sys.stderr.flush()


def run_json2csv(graph_file, mapping_file, pickle_file, prefix, thr=0.95, device="cuda:0", timings=None):
    """The whole json2csv job; returns (graph, reads, result)."""
    from . import decode, engine, gfa, report
    t = {}
    t0 = time.perf_counter()
    print("Load the pangenome graph")
    update_progress(0)
    graph = gfa.load_graph(graph_file, pickle_file)
    update_progress(1)
    t["graph_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    print("Parsing Alignment: first pass")
    update_progress(0)
    reads = decode.decode_alignments(mapping_file, graph, thr)      # `vg view -j -K` JSON, or the GAMP file itself
    t["decode_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    eng = engine.AbundanceEngine(graph, device, deterministic=True)     # files out: bit-reproducible from run to run
    t["abundance_engine_s"] = time.perf_counter() - t0
    update_progress(1)                                                  # (the first pass is the decode + the kernels' pass 1)
    print("Parsing Alignment: second pass")
    update_progress(0)
    res = eng.run(reads, keep_codes=True, keep_slots=False, keep_matches=False)    # the reports need the codes, not the match lists
    update_progress(1)
    t["abundance_s"] = time.perf_counter() - t0
    t0 = t1 = time.perf_counter()

    def lap(key):
        nonlocal t1
        t[key] = time.perf_counter() - t1
        t1 = time.perf_counter()

    clusters = report.cluster_records(graph, reads, res)
    lap("report_cluster_records_s")
    gene_csv = prefix + "_gene_table.csv"
    print(f"Print gene-level results to file {gene_csv}")
    update_progress(0)
    report.write_gene_table(gene_csv, graph, res)
    update_progress(1)
    lap("report_gene_table_s")
    for k, name in enumerate(report.HIST_FILES):
        print(f"Print {name} distribution results to file {prefix + '_' + name + '.csv'}")
        report.write_histogram(prefix + "_" + name + ".csv", graph, res, k)
    print(f"Print cluster-level unassigned reads distribution results to file {prefix + '_unassigned.csv'}")
    lap("report_histograms_s")
    report.write_unassigned(prefix + "_unassigned.csv", graph, clusters, engine.cluster_strain_counts(graph, device))
    lap("report_unassigned_s")
    report.write_clusters_pickle("allclusters.pickle", clusters)        # current directory, as json2csv.py:1344 (Q11)
    for line in report.summary_lines(res.C, thr, prefix):
        print(line)
    lap("report_pickle_s")
    t["report_s"] = time.perf_counter() - t0
    if timings is not None:
        timings.update(t)
    return graph, reads, res


def json2csv_main():
    graph_file = pickle_file = mapping_file = prefix = None
    thr = 0.95
    try:
        opts, _ = getopt.getopt(sys.argv[1:], "hg:p:o:m:t:")     # no "f": -f is rejected, as in the reference (Q2)
    except getopt.GetoptError as err:
        print(err)
        _usage_json2csv()
        sys.exit(2)
    for o, a in opts:
        if o == "-h":
            _usage_json2csv()
        elif o == "-g":
            graph_file = a
        elif o == "-m":
            mapping_file = a
        elif o == "-p":
            pickle_file = a
        elif o == "-o":
            prefix = a
        elif o == "-t":
            thr = float(a)
    if not graph_file or not pickle_file or not mapping_file:
        _usage_json2csv()
        sys.exit()
    if prefix is None:
        raise TypeError("unsupported operand type(s) for +: 'NoneType' and 'str'")     # json2csv.py:1325
    run_json2csv(graph_file, mapping_file, pickle_file, prefix, thr)


def _usage_strains():
    print(f"Usage: python {sys.argv[0]} -i input_file (csv) -o out_file -t thr")


def compute_strains_abundance_main():
    input_file = out_file = None
    thr = 0.5
    try:
        opts, _ = getopt.getopt(sys.argv[1:], "hi:o:t:")
    except getopt.GetoptError as err:
        print(err)
        _usage_strains()
        sys.exit(2)
    for o, a in opts:
        if o == "-h":
            _usage_strains()
        elif o == "-i":
            input_file = a
        elif o == "-o":
            out_file = a
        elif o == "-t":
            thr = float(a)
    if not input_file or not out_file:
        _usage_strains()
        sys.exit()
    from . import strains
    strains.compute_strains_abundance(input_file, out_file, thr)
