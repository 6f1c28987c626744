"""Top source lines of an `ncu --page source --csv --print-source cuda,sass` export: samples, instructions, active lanes."""
import csv
import sys

path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur, out, hdr = None, [], None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":        # per-line summary rows have "-" as the SASS address
        continue
    g = lambda name: r[hdr.index(name)]                         # noqa: E731
    try:
        ie = int(g("Instructions Executed"))
        out.append((int(g("# Samples")), ie, int(g("Thread Instructions Executed")) / max(1, ie), cur, r[0], r[1].strip()[:110]))
    except ValueError:
        pass
tot_s, tot_i = sum(o[0] for o in out), sum(o[1] for o in out)
print(f"total samples {tot_s}, warp instructions {tot_i}")
for s, i, act, f, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{100 * s / max(1, tot_s):5.1f}% smp {100 * i / max(1, tot_i):5.1f}% ins  lanes {act:5.1f}  {f}:{ln}  {src}")
