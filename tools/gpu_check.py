#!/usr/bin/env python3
"""GPU parity check: align FASTA reads on the device through the C ABI, format
SAM, and compare byte-for-byte with the unmodified reference binary run on the
same box (oracle/_ref).  usage: gpu_check.py <index> <reads.fa> [reads2.fa]"""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hisat2_b200 as h2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFBIN = os.path.join(ROOT, "oracle", "_ref", "hisat2-align-s")

def main():
    idx_base, r1 = sys.argv[1], sys.argv[2]
    r2 = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("-") else None
    t0 = time.time()
    idx = h2.Index(idx_base)
    t1 = time.time()
    batch = h2.ReadBatch.from_fasta(r1, path2=r2)
    t2 = time.time()
    res = idx.align(batch, allow_capacity=True)
    t3 = time.time()
    sam = idx.sam_header() + idx.format_sam(batch, res)
    t4 = time.time()
    nerr = int((res.reads["err"] != 0).sum())
    print("open %.2fs parse %.2fs align %.3fs (h2d %.2f ms kernel %.2f ms d2h %.2f ms) sam %.2fs reads %d alns %d err-reads %d LF/read %.1f"
          % (t1 - t0, t2 - t1, t3 - t2, res.ms_h2d, res.ms_kernel, res.ms_d2h, t4 - t3, batch.n, len(res.alns), nerr,
             res.reads["n_lf"].mean()))
    out = "/tmp/gpu_%s.sam" % os.path.basename(r1)
    open(out, "wb").write(sam)
    refout = "/tmp/ref_%s.sam" % os.path.basename(r1)
    cmd = [REFBIN, "--no-spliced-alignment", "-f", "-x", idx_base, "-S", refout, "-p", str(os.cpu_count() or 1), "--reorder"]
    cmd += (["-1", r1, "-2", r2] if r2 else ["-U", r1])
    t5 = time.time()
    subprocess.run(cmd, check=True, stderr=subprocess.DEVNULL)
    t6 = time.time()
    a = [l for l in open(refout, "rb") if not l.startswith(b"@PG")]
    b = sam.splitlines(keepends=True)
    same = a == b
    nd = sum(1 for x, y in zip(a, b) if x != y) + abs(len(a) - len(b))
    print("reference (-p %d) %.2fs ; SAM identical: %s (differing lines %d of %d)" % (os.cpu_count(), t6 - t5, same, nd, len(a)))
    return 0 if same else 1

if __name__ == "__main__":
    sys.exit(main())
