#!/usr/bin/env python3
"""Developer tool: SASS instruction count of one kernel attributed to the source functions of csrc/*.cuh (via -lineinfo).
usage: python tools/sass_by_function.py lib/libcpdp_quadrotor.so k_riccati_bdf"""
import collections
import os
import re
import subprocess
import sys
import tempfile

so, kern = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
inside = False
cnt = collections.Counter()
cur = None
for line in txt:
    if line.startswith(".text."):
        inside = kern in line
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line) and cur:
        cnt[cur] += 1
cache = {}


def func(path, ln):
    if path not in cache:
        starts = []
        try:
            for i, l in enumerate(open(path), 1):
                m = re.match(r"\s*(?:template <[^>]*>\s*)?(?:CPDP_(?:HD|D|D_NOINLINE|GLOBAL)\b|static|inline|__device__).*?\b(\w+)\s*\(", l)
                if m and not l.strip().startswith("//") and "=" not in l.split("(")[0]:
                    starts.append((i, m.group(1)))
        except OSError:
            pass
        cache[path] = starts
    name = os.path.basename(path)
    for i, n in cache[path]:
        if i <= ln:
            name = n
    return name


tot = collections.Counter()
for (f, l), c in cnt.items():
    tot[func(f, l)] += c
for k, v in tot.most_common(40):
    print("%6d  %s" % (v, k))
print("%6d  total" % sum(tot.values()))
