"""Stall samples and instructions of an `ncu --page source --csv --print-source cuda,sass` export, grouped by the
function (of the given source file) that contains each line."""
import csv
import re
import sys

path, src = sys.argv[1], sys.argv[2]
lines = open(src).read().split("\n")
# function starts: a line at column 0 that looks like a definition
starts = []
for i, ln in enumerate(lines, 1):
    m = re.match(r"^(?:template.*\n)?(?:__device__|__global__|static|template|struct)\b.*?(\w+)\s*\(", ln)
    if ln.startswith(("__device__", "__global__", "static ", "struct ")):
        m2 = re.search(r"(\w+)\s*(\(|\{)", ln.split("__forceinline__")[-1].split("__noinline__")[-1])
        starts.append((i, m2.group(1) if m2 else ln[:30]))
def func_of(n):
    name = "?"
    for i, nm in starts:
        if i <= n:
            name = nm
        else:
            break
    return name
rows = list(csv.reader(open(path)))
cur = hdr = None
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":
        continue
    g = lambda n: r[hdr.index(n)]                       # noqa: E731
    try:
        smp, ie, te = int(g("# Samples")), int(g("Instructions Executed")), int(g("Thread Instructions Executed"))
        lsb, bar = int(g("stall_long_sb")), int(g("stall_barrier"))
    except ValueError:
        continue
    key = func_of(int(r[0])) if cur == src.split("/")[-1] else cur
    a = agg.setdefault(key, [0, 0, 0, 0, 0])
    for k, v in enumerate((smp, ie, te, lsb, bar)):
        a[k] += v
ts, ti = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print(f"samples {ts}  warp instructions {ti}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if a[0] * 200 < ts and a[1] * 200 < ti:
        continue
    print(f"{100 * a[0] / ts:5.1f}% smp (long_sb {100 * a[3] / max(1, a[0]):3.0f}% barrier {100 * a[4] / max(1, a[0]):3.0f}%)  "
          f"{100 * a[1] / ti:5.1f}% ins  lanes {a[2] / max(1, a[1]):5.1f}  {k}")
