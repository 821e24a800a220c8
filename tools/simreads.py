#!/usr/bin/env python3
"""Seeded synthetic read generator.

The reference's own simulator (hisat2_simulate_reads.py:101-107) fails on
Python 3 and filters out the example chromosome, so the benchmark reads come
from this repo-local generator instead: fragment length U[200,400], 101-bp
mates from both fragment ends (mate 2 reverse-complemented), random strand
swap, per-base substitution rate 0.5 %, no Ns, FASTA.  Deterministic for a
given (reference, n, seed).  numpy only.
"""
import argparse
import numpy as np

COMP = np.frombuffer(bytes.maketrans(b"ACGTN", b"TGCAN"), dtype=np.uint8)


def load_fasta_codes(path):
    """Return (name, uint8 ASCII upper-case array) of the first sequence."""
    name = None
    chunks = []
    with open(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                if name is not None:
                    break
                name = line[1:].split()[0].decode()
            else:
                chunks.append(line.strip().upper())
    seq = np.frombuffer(b"".join(chunks), dtype=np.uint8)
    return name, seq


def simulate(seq, n, seed=1, rdlen=101, fmin=200, fmax=400, sub=0.005):
    """Return (mate1, mate2) uint8 ASCII arrays of shape (n, rdlen).

    Fragments containing N are resampled; mate1 is the fragment's left end
    on the chosen strand, mate2 the reverse complement of its right end
    (--fr orientation)."""
    rng = np.random.default_rng(seed)
    L = len(seq)
    isn = (seq == ord("N")).astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(isn)])
    m1 = np.empty((n, rdlen), dtype=np.uint8)
    m2 = np.empty((n, rdlen), dtype=np.uint8)
    filled = 0
    ar = np.arange(rdlen)
    while filled < n:
        want = n - filled
        flen = rng.integers(fmin, fmax + 1, size=want)
        pos = rng.integers(0, L - fmax, size=want)
        ok = (csum[pos + flen] - csum[pos]) == 0
        pos, flen = pos[ok], flen[ok]
        k = len(pos)
        left = seq[pos[:, None] + ar[None, :]]
        right = seq[(pos + flen - rdlen)[:, None] + ar[None, :]]
        right_rc = COMP[right[:, ::-1]]
        left_rc = COMP[left[:, ::-1]]
        swap = rng.integers(0, 2, size=k).astype(bool)
        a = np.where(swap[:, None], right_rc, left)
        b = np.where(swap[:, None], left, right_rc)
        # wait: on the minus strand mate1 = revcomp(right end), mate2 = left end
        for arr in (a, b):
            mut = rng.random(arr.shape) < sub
            shift = rng.integers(1, 4, size=arr.shape)
            codes = np.searchsorted(np.frombuffer(b"ACGT", dtype=np.uint8), arr)
            newc = np.frombuffer(b"ACGT", dtype=np.uint8)[(codes + shift) & 3]
            arr[mut] = newc[mut]
        m1[filled:filled + k] = a
        m2[filled:filled + k] = b
        filled += k
    return m1, m2


def add_noise(reads, seed, indel=0.0, nrate=0.0, ragged=False):
    """Return a list of byte strings: reads with random short indels, Ns and
    (optionally) ragged lengths -- edge cases for the parity tests."""
    rng = np.random.default_rng(seed + 7919)
    out = []
    for r in reads:
        b = bytearray(r.tobytes())
        if indel > 0 and rng.random() < indel * len(b):
            pos = int(rng.integers(5, len(b) - 5))
            ln = int(rng.integers(1, 4))
            if rng.random() < 0.5:
                del b[pos:pos + ln]
            else:
                b[pos:pos] = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), ln))
        if nrate > 0:
            for i in np.nonzero(rng.random(len(b)) < nrate)[0]:
                b[int(i)] = ord("N")
        if ragged:
            cut = int(rng.integers(0, 70))
            if cut and rng.random() < 0.5:
                b = b[:len(b) - cut]
        out.append(bytes(b))
    return out


def write_fasta_list(path, reads, prefix="r"):
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            f.write(b">%s%d\n%s\n" % (prefix.encode(), i, r))


def write_fasta(path, reads, prefix="r"):
    n, rdlen = reads.shape
    with open(path, "wb") as f:
        step = 100000
        for s in range(0, n, step):
            e = min(n, s + step)
            out = bytearray()
            blk = reads[s:e]
            for i in range(e - s):
                out += b">%s%d\n" % (prefix.encode(), s + i)
                out += blk[i].tobytes()
                out += b"\n"
            f.write(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reference")
    ap.add_argument("n", type=int)
    ap.add_argument("out_prefix")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--paired", action="store_true")
    ap.add_argument("--indel", type=float, default=0.0, help="per-base chance of one short indel per read")
    ap.add_argument("--nrate", type=float, default=0.0, help="per-base N rate")
    ap.add_argument("--sub", type=float, default=0.005)
    ap.add_argument("--ragged", action="store_true", help="randomly truncate half of the reads")
    ap.add_argument("--rdlen", type=int, default=101)
    a = ap.parse_args()
    _, seq = load_fasta_codes(a.reference)
    if a.rdlen != 101:
        m1, m2 = simulate(seq, a.n, a.seed, rdlen=a.rdlen, fmin=max(2 * a.rdlen, 100), fmax=max(3 * a.rdlen, 200), sub=a.sub)
    else:
        m1, m2 = simulate(seq, a.n, a.seed, sub=a.sub)
    if a.indel > 0 or a.nrate > 0 or a.ragged:
        write_fasta_list(a.out_prefix + "_1.fa", add_noise(m1, a.seed, a.indel, a.nrate, a.ragged))
        if a.paired:
            write_fasta_list(a.out_prefix + "_2.fa", add_noise(m2, a.seed + 1, a.indel, a.nrate, a.ragged))
        return
    write_fasta(a.out_prefix + "_1.fa", m1)
    if a.paired:
        write_fasta(a.out_prefix + "_2.fa", m2)


if __name__ == "__main__":
    main()
