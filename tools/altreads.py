#!/usr/bin/env python3
"""Reads that carry ALT alleles: apply a seeded random, non-overlapping subset of a HISAT2 SNP list
(single / deletion / insertion, 0-based positions) to the reference and sample reads from the result.

usage: altreads.py <ref.fa> <snps.snp> <n_pairs> <out_prefix> [--seed S] [--frac F] [--sub RATE] [--paired]
"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import simreads


def load_fasta(path):
    names, seqs = [], []
    for l in open(path):
        if l.startswith(">"):
            names.append(l[1:].split()[0]); seqs.append([])
        else:
            seqs[-1].append(l.strip())
    return names, ["".join(x) for x in seqs]


def apply_alts(names, seqs, snp_path, frac, rng):
    by_chr = {n: [] for n in names}
    for l in open(snp_path):
        f = l.rstrip("\n").split("\t")
        if len(f) < 5 or f[2] not in by_chr:
            continue
        by_chr[f[2]].append((int(f[3]), f[1], f[4]))
    out = []
    for n, sq in zip(names, seqs):
        alts = sorted(by_chr[n])
        pieces, cur, last_end = [], 0, -1
        for pos, kind, val in alts:
            if pos <= last_end + 1 or rng.random() > frac:
                continue
            if kind == "single":
                pieces.append(sq[cur:pos]); pieces.append(val); cur = pos + 1; last_end = pos
            elif kind == "deletion":
                ln = int(val)
                pieces.append(sq[cur:pos]); cur = pos + ln; last_end = pos + ln - 1
            elif kind == "insertion":
                pieces.append(sq[cur:pos]); pieces.append(val); cur = pos; last_end = pos
        pieces.append(sq[cur:])
        out.append("".join(pieces))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("ref"); ap.add_argument("snps"); ap.add_argument("n", type=int); ap.add_argument("out_prefix")
    ap.add_argument("--seed", type=int, default=1); ap.add_argument("--frac", type=float, default=0.5)
    ap.add_argument("--sub", type=float, default=0.003); ap.add_argument("--paired", action="store_true")
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    names, seqs = load_fasta(a.ref)
    alt = apply_alts(names, seqs, a.snps, a.frac, rng)
    cat = ("N" * 500).join(alt)
    arr = np.frombuffer(cat.upper().encode(), dtype=np.uint8)
    m1, m2 = simreads.simulate(arr, a.n, a.seed, sub=a.sub)
    simreads.write_fasta(a.out_prefix + "_1.fa", m1, prefix="a")
    if a.paired:
        simreads.write_fasta(a.out_prefix + "_2.fa", m2, prefix="a")


if __name__ == "__main__":
    main()
