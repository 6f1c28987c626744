"""Top SASS instructions of a kernel by stall samples, from `ncu --page source --csv` output (stdin).
usage: ncu -i rep --page source --csv --kernel-name NAME | python tools/ncu_hot.py [N] [table_index]"""
import csv, sys
rows = list(csv.reader(sys.stdin))
n_top = int(sys.argv[1]) if len(sys.argv) > 1 else 25
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
a = starts[which]
b = starts[which + 1] if which + 1 < len(starts) else len(rows)
print(rows[a][1][:80])
hdr = rows[a + 1]
i_src, i_samp, i_exec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
body = [r for r in rows[a + 2:b] if len(r) > i_exec]
tot = sum(int(r[i_samp] or 0) for r in body)
texec = sum(int(r[i_exec] or 0) for r in body)
print("total samples", tot, "warp instructions", texec, "static instructions", len(body))
for r in sorted(body, key=lambda r: -int(r[i_samp] or 0))[:n_top]:
    print(f"{100*int(r[i_samp])/tot:5.1f}%  exec {int(r[i_exec]):>10}  {r[i_src].strip()[:110]}")
