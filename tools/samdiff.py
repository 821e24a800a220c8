#!/usr/bin/env python3
"""Compare two SAM files record-by-record ignoring @PG; print a compact summary."""
import sys
def load(p):
    return [l.rstrip("\n") for l in open(p) if not l.startswith("@PG")]
a, b = load(sys.argv[1]), load(sys.argv[2])
maxshow = int(sys.argv[3]) if len(sys.argv) > 3 else 10
print("lines", len(a), len(b))
def key(l): return l.split("\t")[0]
from collections import OrderedDict
def group(ls):
    d = OrderedDict()
    for l in ls:
        if l.startswith("@"): d.setdefault("@hdr", []).append(l)
        else: d.setdefault(key(l), []).append(l)
    return d
ga, gb = group(a), group(b)
nd = 0
for k in ga:
    if ga[k] != gb.get(k):
        nd += 1
        if nd <= maxshow:
            print("== read", k)
            for l in ga[k]:
                f = l.split("\t"); print("  ref:", f[1], f[3], f[4], f[5], " ".join(f[11:]))
            for l in gb.get(k, []):
                f = l.split("\t"); print("  got:", f[1], f[3], f[4], f[5], " ".join(f[11:]))
print("differing reads:", nd, "of", len(ga))
sys.exit(1 if (nd or len(a) != len(b)) else 0)
