"""Times sfb_expand_reads (device-side expansion of the transfer layout) alone on a C3-shaped shard."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from strainflair_b200 import engine, synth

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.125
graph, reads = synth.make_config("c3", scale=scale)
eng = engine.AbundanceEngine(graph, "cuda:0")
host = eng.stage(reads, pin=False) if not os.environ.get("SFB_FILE_ORDER") else engine.HostReads(reads)   # tile order, as the engine stages it
dr = engine.DeviceReads(host, eng.device, None, None, eng.node_len, expand=False)
torch.cuda.synchronize()
ts = []
for it in range(8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dr.expand()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
best = min(ts[2:])
print(os.path.basename(os.environ.get("SFB200_LIB", "default")), f"expand {best:.3f} ms for {reads.n_records} records "
      f"({reads.n_records / best / 1e6:.1f} G records/s, {8 * reads.n_records / best / 1e6:.0f} GB/s written)", flush=True)
