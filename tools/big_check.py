#!/usr/bin/env python3
"""HBM-resident regime: 200 Mbp synthetic reference (4 chromosomes, repeat families, N gaps; index image
~0.4 GB, 3.5k local indexes).  Parity against the reference binary run on this box, then throughput."""
import os, sys, time, subprocess, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hisat2_b200 as h2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = os.path.join(ROOT, "data_big")
REF = os.path.join(ROOT, "oracle", "_ref", "hisat2-align-s")
t0 = time.time()
idx = h2.Index(os.path.join(B, "synth200"))
print("open %.1f s, image %.1f MB" % (time.time() - t0, idx.image().nbytes / 1e6), flush=True)

def lines(b):
    return [l for l in b.splitlines(True) if not l.startswith(b"@PG")]

for paired in (False, True):
    f1, f2 = os.path.join(B, "big200k_1.fa"), os.path.join(B, "big200k_2.fa")
    batch = h2.ReadBatch.from_fasta(f1, path2=f2 if paired else None)
    res = idx.align(batch)
    assert int((res.reads["err"] != 0).sum()) == 0
    sam = idx.sam_header() + idx.format_sam(batch, res)
    out = "/tmp/ref_big.sam"
    best = None
    for p in (8, 16, 32, 64):
        if p > (os.cpu_count() or 1): break
        t = time.time()
        subprocess.run([REF, "--no-spliced-alignment", "-f", "-x", os.path.join(B, "synth200")] +
                       (["-1", f1, "-2", f2] if paired else ["-U", f1]) + ["-S", out, "-p", str(p), "--reorder"],
                       check=True, stderr=subprocess.DEVNULL)
        dt = time.time() - t
        if best is None or dt < best[0]: best = (dt, p)
    same = lines(sam) == lines(open(out, "rb").read())
    print("paired=%d reads=%d parity_with_reference=%s  gpu kernel %.2f ms (%.2f M reads/s)  reference best %.2f s at -p %d (%.3f M reads/s incl. index load)"
          % (paired, batch.n, same, res.ms_kernel, batch.n / res.ms_kernel / 1e3, best[0], best[1], batch.n / best[0] / 1e6), flush=True)
    if not paired:
        # throughput on 1M reads (the 200k batch tiled 5x) and the seed kernel alone
        k = 5
        big = h2.ReadBatch(np.tile(batch.seq, k), np.arange(0, k * batch.n + 1, dtype=np.uint64) * 101, np.tile(batch.seeds, k),
                           [b"t%d" % i for i in range(k * batch.n)])
        for it in range(3):
            r = idx.align(big, resident_iters=1)
            print("  1M reads: kernel %.2f ms  %.2f M reads/s  alg %.1f GB/s" % (r.ms_kernel, big.n / r.ms_kernel / 1e3,
                  float(r.reads["alg_bytes"].astype(np.float64).sum()) / r.ms_kernel / 1e6), flush=True)
            r.close()
        s = idx.seed_search(big, max_range=4)
        print("  seed kernel (2 passes): %.2f ms  %.1f M reads/s  %.2f G LF/s  alg %.1f GB/s" %
              (s.ms_kernel, big.n / s.ms_kernel / 1e3, s.n_lf / s.ms_kernel / 1e6, s.alg_bytes / s.ms_kernel / 1e6), flush=True)
