"""Instruction count of one kernel per source line, from `nvdisasm -g` of a cubin built with -lineinfo (no GPU needed):
which lines the code size of a kernel comes from (instruction-cache pressure shows up as "no instruction" stalls).
usage: python tools/sass_lines.py <file.dis> <kernel substring> [top n]"""
import collections
import re
import sys

path, kern, top = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40
cur, on, cnt = None, False, collections.Counter()
for l in open(path):
    if l.startswith(".text."):
        on = kern in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l):
        cnt[cur] += 1
print("total", sum(cnt.values()))
byfile = collections.Counter()
for (f, ln), c in cnt.items():
    byfile[f] += c
print(byfile.most_common())
for (f, ln), c in cnt.most_common(top):
    print(c, f, ln)
