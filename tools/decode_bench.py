"""Decoder throughput on a synthetic mapping file (phase split on stderr with SFB_DECODE_TRACE=1).

    python tools/decode_bench.py [n_pairs] [threads ...]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from strainflair_b200 import decode, synth  # noqa: E402


def main():
    n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 40_000
    threads = [int(x) for x in sys.argv[2:]] or [1, len(os.sched_getaffinity(0))]
    out = os.environ.get("SFB_DECODE_DIR", "gpurun_out/dec")
    os.makedirs(out, exist_ok=True)
    big = os.path.join(out, f"big_{n_pairs}.json")
    gkw, rkw = synth.CONFIGS["c3"]
    graph = synth.make_graph(synth.SEED0 + 2, **gkw)
    if not os.path.exists(big):
        sample = synth.make_reads(graph, 4242, n_pairs=n_pairs, read_len=rkw["read_len"])
        gf, js, pk = synth.export_files(graph, sample, out)
        blob = open(js, "rb").read()
        with open(big, "wb") as dst:
            for _ in range(8):
                dst.write(blob)
    size = os.path.getsize(big)
    n_lines = sum(chunk.count(b"\n") for chunk in iter(lambda f=open(big, "rb"): f.read(1 << 24), b""))
    for t in threads:
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            got = decode.decode_json(big, graph, 0.95, threads=t)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        print(f"threads {t}: {n_lines / best:.3e} lines/s, {size / best / 1e6:.1f} MB/s, {best:.3f} s, "
              f"alns {got.n_alns} recs {got.n_records}", flush=True)


if __name__ == "__main__":
    main()
