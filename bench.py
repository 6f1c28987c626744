#!/usr/bin/env python
"""Benchmark of the abundance pass (BASELINE.json metric: aligned reads/s through the abundance pass,
plus % of the HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3] [--impl ours|reference]

One process per GPU (torchrun for N > 1).  A step is one full abundance pass
(match -> classify -> pass-1 scatter -> [allreduce] -> pass 2 -> [allreduce] -> path reduce) over this rank's
shard of the synthetic workload; every rank holds a full-size shard (weak scaling).  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "aligned_reads_per_s_abundance_pass"
UNIT = "reads/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="c3", help="synthetic workload (strainflair_b200/synth.py CONFIGS)")
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the config's read pairs per rank")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample-pairs", type=int, default=160_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-shards", type=int, default=2)
    ap.add_argument("--no-det", action="store_true", help="skip the measurements of the other modes (plain fp64 atomics, global-memory kernels)")
    ap.add_argument("--no-extras", action="store_true", help="skip the C4 / C5 sub-measurements")
    return ap.parse_args()


def workload_name(args):
    from strainflair_b200 import synth
    gkw, rkw = synth.CONFIGS[args.config]
    pairs = int(rkw["n_pairs"] * args.scale)
    return (f"{args.config}: synthetic {pairs} read pairs/GPU ({rkw['read_len']} bp) over a "
            f"{gkw['n_clusters']}-cluster x {gkw['chain_len']}-position bubble-chain pangenome, {gkw['n_strains']} strains")


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (the reference is Python and cannot travel to the GPU box)
# ------------------------------------------------------------------------------------------------
def cpu_port_rate(graph_d, reads_d, cores, repeats=1):
    """Times the oracle's abundance pass (decode excluded, like the GPU arm) on ``cores`` processes.
    Returns (reads/s, records/s, seconds per pass, result of the last pass as a dict of arrays)."""
    from oracle import abundance as oa
    from oracle import parallel as op
    n_reads = int(reads_d["grp_nmates"].sum())
    n_rec = len(reads_d["walk_node"])
    best = None
    if cores <= 1:
        idx = oa.AbundanceOracle.build_index(graph_d)
        for _ in range(repeats):
            t0 = time.perf_counter()
            orc = oa.AbundanceOracle(graph_d, reads_d, index=idx).run()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        import numpy as np
        out = dict(C=orc.C, uniq_reads=np.asarray(orc.uniq_reads, np.int64), hist=orc.hist, n_deferred=len(orc.deferred),
                   mult_hits=np.array([not isinstance(v, int) for v in orc.mult_reads]),
                   **{k: np.asarray(getattr(orc, k), np.float64) for k in ("uniq_norm", "mult_reads", "mult_norm", "uniq_abund", "mult_abund")})
    else:
        po = op.ParallelOracle(graph_d, reads_d, cores=cores)
        try:
            for _ in range(repeats):
                t0 = time.perf_counter()
                out = po.run()
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
        finally:
            po.close()
    return n_reads / best, n_rec / best, best, out


def decode_rates(graph, synth, idx, rkw, n_pairs):
    """JSON -> CSR decode throughput on a sample written to a temporary directory."""
    import shutil
    import tempfile
    from oracle import decode as od
    from strainflair_b200 import decode
    tmp = tempfile.mkdtemp(prefix="sfb_decode_")
    try:
        sample = synth.make_reads(graph, synth.SEED0 + idx + 1000, n_pairs=n_pairs, read_len=rkw["read_len"])
        gf, js, pk = synth.export_files(graph, sample, tmp)
        # the sample repeated 8 times (every copy starts a new read group): a file big enough to amortise start-up
        big = os.path.join(tmp, "big.json")
        with open(big, "wb") as dst:
            blob = open(js, "rb").read()
            for _ in range(8):
                dst.write(blob)
        n_lines = 8 * blob.count(b"\n")
        size = os.path.getsize(big)
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            got = decode.decode_json(big, graph, 0.95)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        assert got.n_alns == 8 * sample.n_alns
        # Python decoder (the reference's json.loads + BFS + metaalign2align, restated) on the first lines, one core
        head = os.path.join(tmp, "head.json")
        n_head = 0
        with open(js) as src, open(head, "w") as dst:
            for ln in src:
                if n_head >= 20_000 and ln.startswith('{"name": "r') and "/1" in ln[:24]:
                    break
                dst.write(ln)
                n_head += 1
        gd = graph.as_dict()
        gd["node_index"] = {int(v): k for k, v in enumerate(graph.node_ids)}
        t0 = time.perf_counter()
        od.decode_json(head, gd, 0.95)
        dt_py = time.perf_counter() - t0
        out = {"native_lines_per_s": n_lines / best, "native_MBps": size / best / 1e6,
               "native_threads": len(os.sched_getaffinity(0)), "lines": n_lines,
               "python_port_lines_per_s": n_head / dt_py, "python_port_cores": 1}
        # the whole json2csv job through the command-line entry (files in, files out), second run
        import contextlib
        import io
        from strainflair_b200 import cli
        cwd = os.getcwd()
        os.chdir(tmp)                                   # allclusters.pickle goes to the current directory
        try:
            for _ in range(2):
                t, t0 = {}, time.perf_counter()
                with contextlib.redirect_stdout(io.StringIO()):
                    _, job_reads, _ = cli.run_json2csv(gf, big, pk, os.path.join(tmp, "job"), 0.95, timings=t)
                t["total_s"] = time.perf_counter() - t0
        finally:
            os.chdir(cwd)
        out["job"] = {"what": "json2csv command line on the same file: GFA + JSON + pickle in, seven CSV files + pickle out",
                      "read_pairs": 8 * n_pairs, "json_MB": round(size / 1e6, 1),
                      "reads_per_s": job_reads.n_mates / t["total_s"],
                      "seconds": {k[:-2]: round(v, 3) for k, v in t.items()}}
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from strainflair_b200 import synth
    gkw, rkw = synth.CONFIGS[args.config]
    idx = list(synth.CONFIGS).index(args.config)
    graph = synth.make_graph(synth.SEED0 + idx, **gkw)
    cores = len(os.sched_getaffinity(0))
    pairs = min(args.cpu_sample_pairs, int(rkw["n_pairs"] * args.scale))
    reads = synth.make_reads(graph, synth.SEED0 + idx + 1000, n_pairs=pairs, read_len=rkw["read_len"])
    gd, rd = graph.as_dict(), reads.as_dict()
    from oracle import abundance as oa
    from oracle import parallel as op
    n_reads = int(rd["grp_nmates"].sum())
    po = op.ParallelOracle(gd, rd, cores=cores) if cores > 1 else None
    index = oa.AbundanceOracle.build_index(gd)

    def step():
        if po is not None:
            po.run()
        else:
            oa.AbundanceOracle(gd, rd, index=index).run()

    try:
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = time.perf_counter() - t0
    finally:
        if po is not None:
            po.close()
    value = n_reads * args.steps / dt
    sample = f"first {pairs} read pairs of the workload per step, oracle port (Python), decode excluded"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32/f64", "data": "synthetic",
        "config": config_dict(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "records_per_s": len(rd["walk_node"]) * args.steps / dt,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock, power and throttle reasons of one GPU, sampled DURING the timed region: NVML in-process every 2 ms
    (a timed region of a few steps is tens of milliseconds), `nvidia-smi` every 100 ms when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _nvml_row(self):
        n, h = self.nvml, self.handle
        sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
        try:
            power = n.nvmlDeviceGetPowerUsage(h) / 1000.0
        except Exception:
            power = 0.0
        try:
            bits = n.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            bits = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        flag = lambda name, old: bool(bits & getattr(n, name, getattr(n, old, 0)))   # noqa: E731
        act = lambda v: "Active" if v else "Not Active"                              # noqa: E731
        return [str(sm), str(self.max_sm), str(power),
                act(flag("nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown")),
                act(flag("nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown")),
                act(flag("nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown")),
                act(flag("nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap"))]

    def run(self):
        while not self.stop_flag:
            if self.nvml is not None:
                try:
                    self.rows.append(self._nvml_row())
                except Exception:
                    self.nvml = None
                    continue
                time.sleep(0.002)
                continue
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
KERNEL_STAGES = ("tile_pass1", "match", "classify", "tile_pass2", "tile_share", "pass2_resolve", "pass2_scatter", "path_reduce")


def log_mem(tag):
    """Host memory of this process to stderr (a rank killed for memory leaves no other trace)."""
    try:
        import psutil
        rss = psutil.Process().memory_info().rss / 2 ** 30
        avail = psutil.virtual_memory().available / 2 ** 30
        print(f"[bench rank {os.environ.get('RANK', '0')}] {tag}: rss {rss:.1f} GiB, available {avail:.1f} GiB", file=sys.stderr, flush=True)
    except Exception:
        pass


def check_parity(res, ref, what):
    """GPU result against the CPU port's on the same read sample: integers exact, fp64 within 1e-9 relative with the
    same zero pattern (north_star).  Raises on a difference."""
    import numpy as np
    for k, v in ref["C"].items():
        assert int(res.C[k]) == int(v), (what, k, res.C[k], v)
    assert np.array_equal(res.uniq_reads, ref["uniq_reads"]), what
    assert np.array_equal(res.hist, ref["hist"]), what
    assert res.n_deferred == ref["n_deferred"], (what, res.n_deferred, ref["n_deferred"])
    assert np.array_equal(res.mult_hits > 0, ref["mult_hits"]), what
    for name in ("uniq_norm", "mult_reads", "mult_norm", "uniq_abund", "mult_abund"):
        a, b = np.asarray(getattr(res, name), np.float64), np.asarray(ref[name], np.float64)
        assert a.shape == b.shape, (what, name)
        bad = np.flatnonzero(np.abs(a - b) > 1e-9 * np.maximum(np.abs(a), np.abs(b)))
        assert bad.size == 0, (what, name, bad[:5], a[bad[:5]], b[bad[:5]])
        assert np.array_equal(a > 0, b > 0), (what, name, "zero pattern")
    return {"status": "ok", "against": what,
            "checked": "20 counters, uniq_reads, histograms, deferred groups: exact; 5 fp64 accumulators: <= 1e-9 relative, "
                       "identical zero pattern"}


def device_pass(eng, reads, steps, warmup, world, dist, device, want_marks=True):
    """Stages ``reads`` for ``eng``, uploads them once and times ``steps`` passes with the inputs resident in HBM.
    Returns (ms for all steps [max over ranks], per-stage ms, stats of one pass, launches per step, host staging)."""
    import torch
    from strainflair_b200 import _lib
    t0 = time.perf_counter()
    host = eng.stage(reads, pin=True)
    stage_s = time.perf_counter() - t0
    host.reads = None                     # the staged arena is all the passes need
    dr = eng.upload(host)
    torch.cuda.synchronize()
    done = tries = 0
    while done < max(warmup, 1):          # warm-up (also settles the capacity of the match / work-item buffers)
        work = eng.run_device(dr)
        torch.cuda.synchronize()
        over = int(eng.acc.counters[_lib.COUNTER_NAMES.index("MATCH_OVERFLOW")].item())
        if over:                          # grow the buffers to the exact need; not a warm-up step
            tries += 1
            assert tries <= 4, "match / work-item buffers still too small after regrowing"
            eng.match_cap_hint = int(work.t["cursor"][0].item()) + 1024
            eng.app_cap_hint = int(work.t["cursor"][2].item()) + 1024
            continue
        done += 1
    stats = eng.fetch(work, dr, keep_codes=False, keep_slots=False)
    assert stats.C["MATCH_OVERFLOW"] == 0
    launches0 = eng.launches
    marks_all = []
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        marks = [] if want_marks else None
        eng.run_device(dr, marks)
        if want_marks:
            marks_all.append(marks)
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    stage_ms = {}
    for marks in marks_all:
        for (_, e_prev), (name, e) in zip(marks[:-1], marks[1:]):
            stage_ms[name] = stage_ms.get(name, 0.0) + e_prev.elapsed_time(e) / steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, stage_ms, stats, (eng.launches - launches0) // max(1, steps), host, stage_s, dr


def roofline_block(args, graph, eng, reads, host, stats, stage_ms, world):
    from strainflair_b200 import roofline
    Cr = dict(stats.C)
    if world > 1:                 # the counters were summed over ranks; the roofline line is per GPU
        Cr = {k: v // world for k, v in Cr.items()}
    by = roofline.kernel_bytes(reads.n_groups, reads.n_alns, reads.n_records, graph.n_paths, eng.dg.n_slots,
                               Cr, stats.match_fill, stats.n_deferred, graph.n_strains,
                               tile_groups=host.grp_begin if eng.tile_kernels else 0,
                               tile_alns=host.aln_begin if eng.tile_kernels else 0, step_index=eng.step_index,
                               bins_groups=host.grp_begin if eng.tile_bins else 0,
                               bins_alns=host.aln_begin if eng.tile_bins else 0, n_nodes=graph.n_nodes)
    kernels = {k: v for k, v in stage_ms.items() if k in KERNEL_STAGES and by.get(k, 0) > 0}
    r01_work = 19_772_806_676          # SURVEY 8d's closed form on the work the round-1 kernels did for this workload (BENCH_r01)
    top = max(kernels, key=kernels.get)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = by[top] / (kernels[top] * 1e-3) / 1e9
    traffic = traffic_src = None
    try:                      # DRAM bytes per launch of that kernel from the committed ncu capture of this workload
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tj = json.load(fh)
        traffic = tj.get(f"{args.config}@{args.scale}", {}).get(top)
        traffic_src = tj.get("source")
    except Exception:
        pass
    t_all = sum(kernels.values()) * 1e-3
    whole, survey, stream = by["total"] / t_all / 1e9, by["survey_total"] / t_all / 1e9, by["stream_total"] / t_all / 1e9
    return {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic,
            "traffic_source": ("profile constant: " + traffic_src) if traffic_src else None,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s",
            "algorithmic_bytes": by[top], "kernel_ms": kernels[top],
            "whole_pass": {"achieved": whole, "frac": whole / peak, "algorithmic_bytes": by["total"]},
            "whole_pass_survey_formula": {"achieved": survey, "frac": survey / peak,
                                          "algorithmic_bytes": by["survey_total"],
                                          "bytes_per_record": by["survey_total"] / max(1, reads.n_records)},
            "whole_pass_survey_formula_r01_work": None if not (args.config == "c3" and args.scale == 1.0) else {
                "what": "the same closed form with the candidates / compares the occurrence-index kernels of round 1 needed "
                        "for this workload (19.77 GB, BENCH_r01): comparable across rounds, the step index needs less",
                "achieved": r01_work / t_all / 1e9, "frac": r01_work / t_all / 1e9 / peak, "algorithmic_bytes": r01_work},
            "whole_pass_streamed_bytes": {"what": "bytes the tile design must move through HBM: every input array once in "
                                                  "pass 1, the deferred groups again in pass 2, outputs, accumulator flushes",
                                          "achieved": stream, "frac": stream / peak, "bytes": by["stream_total"]},
            "per_kernel": {k: {"ms": round(kernels[k], 4), "GBps": by[k] / (kernels[k] * 1e-3) / 1e9,
                               "frac": by[k] / (kernels[k] * 1e-3) / 1e9 / peak} for k in kernels}}


def extra_config(name, scale, steps, device, comm, world, rank, dist, strong=False):
    """Device-resident pass on another BASELINE.json configuration (C4: ~8+ candidate paths per read; C5: the 20 M-node
    graph), a few steps, reported beside the headline workload.  ``strong``: the configuration's reads are split over
    the ranks (strong scaling) instead of one ``scale`` share per rank."""
    import torch
    from strainflair_b200 import engine, synth
    gkw, rkw = synth.CONFIGS[name]
    idx = list(synth.CONFIGS).index(name)
    graph = synth.make_graph(synth.SEED0 + idx, **gkw)
    pairs = max(1, int(rkw["n_pairs"] * scale / (world if strong else 1)))
    reads = synth.make_reads(graph, synth.SEED0 + idx + 1000 + 7919 * rank, n_pairs=pairs, read_len=rkw["read_len"])
    eng = engine.AbundanceEngine(graph, device, comm=comm)
    ms, stage_ms, stats, launches, host, stage_s, dr = device_pass(eng, reads, steps, 2, world, dist, device)
    tot = torch.tensor([float(reads.n_mates), float(reads.n_records)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tot)
    out = {"workload": f"{name}: {pairs} read pairs/GPU, {graph.n_nodes} nodes, {graph.n_paths} paths, {eng.dg.n_slots} slots",
           "scaling": "strong" if strong else "weak", "ms_per_step": ms / steps, "steps": steps,
           "reads_per_s": float(tot[0].item()) * steps / (ms * 1e-3), "records_per_s": float(tot[1].item()) * steps / (ms * 1e-3),
           "stage_ms": {k: round(v, 4) for k, v in stage_ms.items()},
           "groups_in_tiles": host.grp_begin / max(1, reads.n_groups), "deferred_groups": stats.n_deferred,
           "mean_matches_per_alignment": (stats.C["ST_TUPLES"] / max(1, world)) / max(1, reads.n_alns),
           "host_staging_s": round(stage_s, 2)}
    del eng, dr, host, reads, graph
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from strainflair_b200 import synth
    gkw, rkw = synth.CONFIGS[args.config]
    idx = list(synth.CONFIGS).index(args.config)
    graph = synth.make_graph(synth.SEED0 + idx, **gkw)
    pairs = max(1, int(rkw["n_pairs"] * args.scale))

    # CPU baseline first (forks workers; do it before CUDA is initialised); its result is kept for the parity check
    cpu = cpu_out = cpu_sample = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = len(os.sched_getaffinity(0))
        sp = min(int(2.5 * args.cpu_sample_pairs), pairs)
        cpu_sample = synth.make_reads(graph, synth.SEED0 + idx + 1000, n_pairs=sp, read_len=rkw["read_len"])
        rate, rec_rate, secs, cpu_out = cpu_port_rate(graph.as_dict(), cpu_sample.as_dict(), cores)
        one = min(sp, 40_000)
        rate1, _, secs1, _ = cpu_port_rate(graph.as_dict(), cpu_sample.slice_groups(0, one).as_dict(), 1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "records_per_s": rec_rate,
               "single_process": {"value": rate1, "unit": UNIT, "sample_pairs": one, "seconds": round(secs1, 1)},
               "speedup_over_single_process": rate / rate1,
               "sample": f"first {sp} read pairs of the workload, oracle port (Python) sharded over {cores} processes, "
                         f"decode excluded, {secs:.1f} s"}

    dec = None          # (decode and the command-line job are timed after the GPU measurements, see below)

    import numpy as np
    import torch
    import torch.distributed as dist
    from strainflair_b200 import engine
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    comm = None
    # from here on this process stays on the host cores of its own GPU's NUMA node (the pinned staging arenas are
    # allocated and filled there, the packer threads run there); the CPU baseline above had every core
    numa_cores = None
    if not os.environ.get("SFB200_NO_NUMA"):
        from strainflair_b200 import dist as sfdist_
        numa_cores = sfdist_.bind_to_gpu_numa(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
        comm = dist.group.WORLD

    t_gen = time.perf_counter()
    reads = synth.make_reads(graph, synth.SEED0 + idx + 1000 + 7919 * rank, n_pairs=pairs, read_len=rkw["read_len"])
    t_gen = time.perf_counter() - t_gen
    log_mem("reads generated")
    # headline: the default engine -- read groups in tile order, order-independent (deterministic) accumulation
    eng = engine.AbundanceEngine(graph, device, comm=comm)
    sampler = ClockSampler(local)
    sampler.start()
    ms, stage_ms, stats, launches, host, stage_s, dr = device_pass(eng, reads, args.steps, args.warmup, world, dist, device)
    sampler.stop_flag = True

    local_reads = int(reads.grp_nmates.sum())
    tot = torch.tensor([float(local_reads), float(reads.n_records), float(reads.n_alns)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    total_reads, total_records, total_alns = float(tot[0].item()), float(tot[1].item()), float(tot[2].item())
    value = total_reads * args.steps / (ms * 1e-3)

    # parity: the same engine on the CPU arm's sample, against the CPU port's result of that run
    parity = None
    if cpu_out is not None:
        parity = check_parity(eng.run(cpu_sample), cpu_out, f"oracle port on the first {cpu_sample.n_groups} read pairs of the workload")
        for key, kw in (("path_sets_sums_at_l2", dict(tile_bins=False)), ("occurrence_index_tuple_lists", dict(step_index=False)),
                        ("all_in_one_tile_kernels", dict(tile_kernels=True))):
            parity[key] = check_parity(engine.AbundanceEngine(graph, device, **kw).run(cpu_sample), cpu_out, "same sample, " + key)["status"]
    elif world > 1:
        # rank-count independence: a fixed sample split over the ranks against the same sample on one rank
        sample = synth.make_reads(graph, synth.SEED0 + idx + 555, n_pairs=min(200_000, pairs), read_len=rkw["read_len"])
        from strainflair_b200 import dist as sfdist
        a, b = sfdist.shard_bounds(sample.n_groups, world, rank)
        many = eng.run(sample.slice_groups(a, b), keep_codes=False, keep_slots=False)
        if rank == 0:
            one = engine.AbundanceEngine(graph, device).run(sample, keep_codes=False, keep_slots=False)
            for k in one.C:
                if not k.startswith("ST_") and k != "MATCH_OVERFLOW":
                    assert one.C[k] == many.C[k], (k, one.C[k], many.C[k])
            assert np.array_equal(one.uniq_reads, many.uniq_reads) and np.array_equal(one.hist, many.hist)
            for name in ("uniq_norm", "mult_reads", "mult_norm", "table"):
                assert np.array_equal(getattr(one, name), getattr(many, name), equal_nan=True), name
            parity = {"status": "ok", "against": f"one rank on the same {sample.n_groups} read pairs",
                      "checked": f"counters, uniq_reads, histograms exact; per-path fp64 sums and the gene table bit-identical "
                                 f"between 1 and {world} ranks (order-independent accumulation)"}

    # for comparison: the plain accumulation mode (fp64 sums, not reproducible); every node-level sum through the L2
    # instead of the shared-memory bins; matches from the occurrence index as tuple lists (the kernels of round 1 on
    # tile-ordered reads); the same on reads in file order; the all-in-one shared-memory tile kernels
    alt = {}
    if not args.no_det:
        for key, kw in (("plain_fp64_sums", dict(deterministic=False)), ("path_sets_sums_at_l2", dict(tile_bins=False)),
                        ("occurrence_index_tuple_lists", dict(step_index=False)), ("file_order", dict(tiles=False)),
                        ("all_in_one_tile_kernels", dict(tile_kernels=True))):
            e2 = engine.AbundanceEngine(graph, device, comm=comm, **kw)
            e2.match_cap_hint, e2.app_cap_hint = eng.match_cap_hint, eng.app_cap_hint
            ms2, st2, _, _, _, _, dr2 = device_pass(e2, reads, max(3, args.steps // 4), 2, world, dist, device)
            n2 = max(3, args.steps // 4)
            alt[key] = {"ms_per_step": ms2 / n2, "value": total_reads * n2 / (ms2 * 1e-3), "unit": UNIT,
                        "stage_ms": {k: round(v, 4) for k, v in st2.items()}}
            del e2, dr2
            torch.cuda.empty_cache()

    # end to end: pinned host buffers -> device -> kernels -> results on the host, every step
    e2e = None
    if not args.no_e2e:
        from strainflair_b200 import dist as sfdist
        del dr
        cuts = [sfdist.shard_bounds(reads.n_groups, args.e2e_shards, k) for k in range(args.e2e_shards)]
        t0 = time.perf_counter()
        shards = eng.stage_shards(reads, len(cuts), pin=True)      # tile order first, shards cut at tile boundaries
        for s_ in shards:
            s_.reads = None
        pack_s = time.perf_counter() - t0
        log_mem("e2e shards staged")
        h2d = sum(s_.nbytes() for s_ in shards)
        from collections import deque
        eng.run_stream(shards)
        eng.run_stream(shards)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        # one step alone: upload, kernels, read-back, nothing else in flight (latency of a single data set)
        t0 = time.perf_counter()
        for _ in range(min(args.steps, 5)):
            res = eng.run_stream(shards)
        torch.cuda.synchronize()
        t_single = torch.tensor([(time.perf_counter() - t0) / min(args.steps, 5)], dtype=torch.float64, device=device)
        # steady state, two steps in flight (submit / collect): the uploads of step k+1 run under pass 2 of step k,
        # every step on its own accumulators and device buffers; every step's inputs are copied from pinned host
        # memory and its results read back inside the timed region
        pending = deque()
        for _ in range(2):
            pending.append(eng.submit(shards, depth=2))
        while pending:
            eng.collect(pending.popleft())
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pending.append(eng.submit(shards, depth=2))
            if len(pending) == 2:
                res = eng.collect(pending.popleft())   # what the reports need: table, per-path counters, histograms
        while pending:
            res = eng.collect(pending.popleft())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(t_single, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        assert res.C["NB_READS"] == stats.C["NB_READS"]
        d2h = eng.d2h_bytes(keep_slots=False) * world
        e2e = {"value": total_reads * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "shards": args.e2e_shards,
               "d2h_bytes_per_step": d2h, "ms_per_step": dt / args.steps * 1e3, "steps_in_flight": 2,
               "single_step_ms": float(t_single.item()) * 1e3,
               "pack_ms": pack_s * 1e3,
               "pack_note": "host staging of one step's decoded reads (tile order + transfer layout + copy into pinned "
                            "arenas, native, all host threads), done once outside the timed region: in a job it belongs to "
                            "the decode stage, which is reported separately"}
        del shards

    extras = {}
    if not args.no_extras and args.config == "c3" and args.scale == 1.0:
        del reads
        torch.cuda.empty_cache()
        if world == 1:
            extras["c4"] = extra_config("c4", 1.0, 5, device, comm, world, rank, dist)
            extras["c5q"] = extra_config("c5", 0.25, 5, device, comm, world, rank, dist)
        else:
            extras["c5"] = extra_config("c5", 1.0, 5, device, comm, world, rank, dist, strong=True)

    # decode, reported separately (north_star): native streaming decoder on all cores vs the oracle's Python decoder, and
    # the whole json2csv job through the command line.  Last: the job runs its own engine in this process, and an engine
    # that ran before the measured one made k_classify 35 % slower (2.51 against 1.85 ms, profiles/README.md r03e).
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.cuda.empty_cache()
        dec = decode_rates(graph, synth, idx, rkw, min(40_000, pairs))

    if rank == 0:
        reads_n = {"groups": host.n_groups, "alns": host.n_alns, "records": host.n_records}

        class _R:            # sizes of this rank's shard for the roofline arithmetic
            n_groups, n_alns, n_records = reads_n["groups"], reads_n["alns"], reads_n["records"]
        Cr = {k: v // world for k, v in stats.C.items()} if world > 1 else dict(stats.C)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32/f64", "data": "synthetic",
            "config": config_dict(args, world),
            "workload_stats": workload_stats(args, world, _R, graph, eng, stats, Cr, host),
            "accumulation": "order-independent (96-bit fixed point): results bit-identical across runs, shard counts and GPU counts",
            "kernels": ("sfb_match_steps (step index, path sets) -> sfb_classify on sets -> sfb_tile_scatter1 || sfb_tile_share "
                        "(shared-memory bins) + global-memory kernels for the read groups outside the tile range")
                       if eng.tile_bins else ("path sets, sums at the L2" if eng.step_index else "occurrence index, tuple lists"),
            "parity": parity,
            "records_per_s": total_records * args.steps / (ms * 1e-3),
            "alignments_per_s": total_alns * args.steps / (ms * 1e-3),
            "gpu_launches": launches * args.steps,
            "stage_ms": {k: round(v, 4) for k, v in stage_ms.items()},
            "roofline": roofline_block(args, graph, eng, _R, host, stats, stage_ms, world),
            "clocks": sampler.summary(),
            "e2e": e2e,
            "cpu_baseline": cpu,
            "decode": dec,
            "other_modes": alt,
            "extra": extras,
            "host_staging_s": round(stage_s, 2),
            "host_cores_bound": None if numa_cores is None else len(numa_cores),
            "gen_seconds": round(t_gen, 1),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def config_dict(args, world):
    """The same dictionary from both arms (the driver compares them)."""
    return {"workload": workload_name(args), "l2": "inputs (GBs of walk arrays) exceed the 126 MB L2",
            "parallelism": f"read groups sharded over {world} GPU(s), graph replicated"}


def workload_stats(args, world, R, graph, eng, stats, Cr, host):
    return {"groups_per_gpu": R.n_groups, "alignments_per_gpu": R.n_alns,
            "records_per_gpu": R.n_records, "paths": graph.n_paths, "nodes": graph.n_nodes,
            "slots": eng.dg.n_slots, "deferred_groups": stats.n_deferred,
            "group_order": "tile order (host staging: read groups by the graph tile that holds them, inside a tile by first node, the rest last)",
            "tiles": 0 if eng.tiling is None else eng.tiling.n_tiles,
            "groups_in_tiles": host.grp_begin / max(1, R.n_groups),
            "tile_tuple_cap": None if eng.tiling is None else eng.tiling.tuple_cap,
            "mean_candidates_per_alignment": Cr["ST_CAND"] / max(1, R.n_alns),
            "mean_steps_per_alignment": Cr["ST_COMPARES"] / max(1, R.n_alns) if eng.step_index else None,
            "mean_matches_per_alignment": Cr["ST_TUPLES"] / max(1, R.n_alns),
            "match_buf_pairs": stats.match_fill, "pass2_work_items": stats.app_fill,
            "pass1_records": Cr["ST_P1_RECORDS"], "pass2_shares": Cr["ST_P2_APPLIED"], "pass2_records": Cr["ST_P2_RECORDS"],
            "unique_assignments": Cr["SAME_CP"] + Cr["DIFF_CP"], "histogram_reads": Cr["ST_HIST"]}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
