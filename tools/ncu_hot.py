"""Top SASS instructions of a kernel by stall samples, from `ncu --page source --csv` output (stdin)."""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[1]
i_src, i_samp, i_exec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
body = [r for r in rows[2:] if len(r) > i_exec]
tot = sum(int(r[i_samp] or 0) for r in body)
texec = sum(int(r[i_exec] or 0) for r in body)
print("total samples", tot, "warp instructions", texec)
top = sorted(body, key=lambda r: -int(r[i_samp] or 0))[: int(sys.argv[1]) if len(sys.argv) > 1 else 25]
for r in top:
    print(f"{100*int(r[i_samp])/tot:5.1f}%  exec {int(r[i_exec]):>10}  {r[i_src].strip()[:110]}")
