#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of a .so / .cubin (cuobjdump -sass).  Usage: sass_hist.py lib.so [kernel-substring]"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur = None
hist = collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        hist[cur][m.group(2).split(".")[0]] += 1
for k, c in hist.items():
    if want in k:
        print(k, sum(c.values()))
        print("   ", ", ".join("%s:%d" % kv for kv in c.most_common(24)))
