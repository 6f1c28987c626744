"""Whole json2csv job (files in, files out) on a synthetic sample, with the time of every phase.

    python tools/job_bench.py [n_pairs] [copies]
"""
import contextlib
import io
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from strainflair_b200 import cli, synth  # noqa: E402


def main():
    n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 40_000
    copies = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    out = os.environ.get("SFB_DECODE_DIR", "gpurun_out/dec")
    os.makedirs(out, exist_ok=True)
    gkw, rkw = synth.CONFIGS["c3"]
    graph = synth.make_graph(synth.SEED0 + 2, **gkw)
    sample = synth.make_reads(graph, 4242, n_pairs=n_pairs, read_len=rkw["read_len"])
    gf, js, pk = synth.export_files(graph, sample, out)
    big = os.path.join(out, "job.json")
    blob = open(js, "rb").read()
    with open(big, "wb") as dst:
        for _ in range(copies):
            dst.write(blob)
    cwd = os.getcwd()
    os.chdir(out)                      # allclusters.pickle goes to the current directory
    try:
        for rep in range(2):
            t = {}
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                g, reads, res = cli.run_json2csv(os.path.basename(gf), "job.json", os.path.basename(pk), "job_out", 0.95, timings=t)
            t["total_s"] = time.perf_counter() - t0
            t = {k: round(v, 3) for k, v in t.items()}
            print(json.dumps({"pairs": n_pairs * copies, "json_MB": round(os.path.getsize("job.json") / 1e6, 1),
                              "groups": reads.n_groups, "alns": reads.n_alns, **t}), flush=True)
    finally:
        os.chdir(cwd)


if __name__ == "__main__":
    main()
