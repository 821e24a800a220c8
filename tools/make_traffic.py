#!/usr/bin/env python3
"""profiles/traffic.json from an ncu summary (tools/ncu_summary.py) of the pool kernel on a bench-shaped launch:
dram__bytes_read.sum + dram__bytes_write.sum of ONE launch, tied to the library sources by build.source_hash()
(bench.py reports it as roofline.traffic only when the hash matches the sources it runs).
usage: make_traffic.py <summary.json> <out traffic.json> <note>"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hisat2_b200 import build
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def parse(v):
    x, u = str(v).split()
    return float(x) * UNIT[u]


d = json.load(open(sys.argv[1]))
out = {"source_hash": build.source_hash(), "dram_bytes_per_launch": parse(d["dram_read"]) + parse(d["dram_write"]),
       "dram_read_bytes": parse(d["dram_read"]), "dram_write_bytes": parse(d["dram_write"]), "kernel": d.get("kernel"),
       "launch_ms_under_ncu": d.get("time"), "note": sys.argv[3] if len(sys.argv) > 3 else ""}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out))
