"""Experiment harness: time single kernels of the pass on the C3 workload (scale given), for kernel variants
built with -DSFB_EXP=n into separate libraries (SFB200_LIB selects one)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from strainflair_b200 import engine, synth

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.2
gkw, rkw = synth.CONFIGS["c3"]
graph = synth.make_graph(synth.SEED0 + 2, **gkw)
reads = synth.make_reads(graph, synth.SEED0 + 1002, n_pairs=int(rkw["n_pairs"] * scale), read_len=150)
eng = engine.AbundanceEngine(graph, "cuda:0")
dr = engine.DeviceReads(engine.HostReads(reads), eng.device)
work = eng._work_for(dr)
for stage in ("match",):
    fn = getattr(eng, stage)
    for _ in range(3):
        work.reset(); eng.acc.zero(); fn(dr, work)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    t = 0.0
    for _ in range(5):
        work.reset(); eng.acc.zero()
        ev[0].record(); fn(dr, work); ev[1].record(); torch.cuda.synchronize()
        t += ev[0].elapsed_time(ev[1])
    print(os.environ.get("SFB200_LIB", "default"), stage, "ms", round(t / 5, 4), "alns", reads.n_alns, flush=True)
