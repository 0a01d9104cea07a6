#!/usr/bin/env python
"""Static SASS instruction count per source line of one kernel. usage: sass_by_line.py all.sass kernel_substr"""
import re, sys, collections
lines = open(sys.argv[1]).read().split("\n")
key = sys.argv[2]
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and key in l)
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith(".text.")), len(lines))
cur = None
cnt = collections.Counter()
order = []
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        cnt[cur] += 1
tot = sum(cnt.values())
print("total", tot)
for k in sorted(cnt, key=lambda k: (k[0], k[1]) if k else ("", 0)):
    print(k, cnt[k])
