"""Experiment harness: time single kernels of the pass on the C3 workload (scale given), for kernel variants
built with -DSFB_EXP=n into separate libraries (SFB200_LIB selects one)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from strainflair_b200 import engine, synth

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.2
cfg = sys.argv[2] if len(sys.argv) > 2 else "c3"
gkw, rkw = synth.CONFIGS[cfg]
idx = list(synth.CONFIGS).index(cfg)
graph = synth.make_graph(synth.SEED0 + idx, **gkw)
reads = synth.make_reads(graph, synth.SEED0 + 1000 + idx, n_pairs=int(rkw["n_pairs"] * scale), read_len=150)
if os.environ.get("SFB_SORT"):      # read groups in locality order (schema.Reads.locality_order)
    reads = reads.permute_groups(reads.locality_order())
eng = engine.AbundanceEngine(graph, "cuda:0")
dr = eng.upload(engine.HostReads(reads))
eng.match_cap_hint = int(os.environ.get("SFB_CAP_MULT", "2")) * reads.n_alns
work = eng._work_for(dr)
stages = ("match", "classify", "pass1_scatter", "pass2_resolve", "pass2_scatter")
tot = {k: 0.0 for k in stages}
for it in range(8):
    work.reset(); eng.acc.zero()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)]
    evs[0].record()
    for k, st in enumerate(stages):
        if st == "pass2_resolve" and os.environ.get("SFB_EXP_SORT"):
            n_def = int(work.t["cursor"][1].item())
            work.t["deferred"][:n_def] = torch.sort(work.t["deferred"][:n_def]).values
            evs[k].record()
        getattr(eng, st)(dr, work)
        evs[k + 1].record()
    torch.cuda.synchronize()
    if it >= 3:
        for k, st in enumerate(stages):
            tot[st] += evs[k].elapsed_time(evs[k + 1]) / 5
from strainflair_b200 import _lib
cn = dict(zip(_lib.COUNTER_NAMES, eng.acc.counters.cpu().tolist()))
chk = (cn["MATCH_OVERFLOW"], cn["ST_TUPLES"], cn["ST_CAND"], cn["ASSIGNED_U"], int(eng.acc.uniq_reads.sum().item()), int(work.t["aln_match"][:reads.n_alns].sum().item()) & 0xffffffff)
print(os.path.basename(os.environ.get("SFB200_LIB", "default")), "sorted" if os.environ.get("SFB_SORT") else "file-order", {k: round(v, 4) for k, v in tot.items()}, cfg, "alns", reads.n_alns, "check", chk, flush=True)
