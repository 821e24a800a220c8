#!/usr/bin/env python3
"""Short run for ncu: align <reads.fa> (default data/sim200k_1.fa) `iters` times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hisat2_b200 as h2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
fa = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "data", "sim200k_1.fa")
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
opts = dict(kv.split("=") for kv in sys.argv[3:])
opts = {k: int(v) for k, v in opts.items()}
idx = h2.Index(os.path.join(ROOT, "data", "22_20-21M"), **opts)
print("opts", opts)
batch = h2.ReadBatch.from_fasta(fa)
for i in range(iters):
    r = idx.align(batch)
    print("iter", i, "kernel ms", r.ms_kernel, "reads/s", batch.n / (r.ms_kernel / 1e3))
    r.close()
