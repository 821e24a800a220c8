#!/usr/bin/env python3
"""RNA-like reads with their splice sites, seeded (test data for the populated splice-site DB).

A "transcript" is 2-3 exons cut from one reference sequence: introns start at a GT and end at an AG four times out
of five (else anywhere).  Single-end reads are 101-base windows of a transcript, biased so that many start or end
only 1-11 bases beyond a junction (the case only a splice-site DB resolves); pairs are fragments of 180-420 bases whose
mates often lie in different exons (template-length adjustment).  0.5 % substitutions, half of the reads reverse
complemented.  The site list holds 85 % of the introns used (+ / - strand by motif), decoys, duplicates and an unknown
sequence name -- everything SpliceSiteDB::read has to cope with (splice_site.cpp:727-775).

  python tools/sim_rna.py ref.fa out_prefix [n_se n_pairs seed]   ->  <prefix>.fa, <prefix>_1.fa, <prefix>_2.fa, <prefix>_ss.txt
"""
import random
import sys

COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


def load_fasta(path):
    refs, name = {}, None
    for l in open(path):
        if l.startswith(">"):
            name = l[1:].split()[0]
            refs[name] = []
        else:
            refs[name].append(l.strip().upper())
    return {k: "".join(v) for k, v in refs.items()}


def revcomp(s):
    return "".join(COMP[c] for c in reversed(s))


def transcript(rng, chrom, seq, nexon, first_len, exon_len=(140, 420)):
    """-> (spliced sequence, [(chrom, left, right, strand)], exon lengths) or None"""
    pos = rng.randrange(0, max(1, len(seq) - 16000))
    parts, sites, lens = [], [], []
    for k in range(nexon):
        ln = first_len if (k == 0 and first_len) else rng.randrange(*exon_len)
        parts.append(seq[pos:pos + ln]); lens.append(ln); pos += ln
        if k + 1 < nexon:
            if rng.random() < 0.8:
                if seq[pos:pos + 2] != "GT":
                    return None
                e = -1
                for _ in range(50):
                    cand = pos + rng.randrange(60, 4000)
                    a = seq.find("AG", cand, cand + 200)
                    if a > 0:
                        e = a + 2
                        break
                if e < 0:
                    return None
                sites.append((chrom, pos - 1, e, "+"))
                pos = e
            else:
                e = pos + rng.randrange(30, 3000)
                sites.append((chrom, pos - 1, e, rng.choice("+-")))
                pos = e
    t = "".join(parts)
    if "N" in t or len(t) != sum(lens):
        return None
    return t, sites, lens


def mutate(rng, rd, rate=0.005):
    return "".join(rng.choice("ACGT".replace(c, "")) if rng.random() < rate else c for c in rd)


def sim(refs, n_se, n_pairs, seed, rdlen=101):
    rng = random.Random(seed)
    names = [k for k in refs if len(refs[k]) > 20000]
    se, pairs, used = [], [], set()
    while len(se) < n_se:
        chrom = rng.choice(names)
        tr = transcript(rng, chrom, refs[chrom], rng.choice([2, 2, 2, 3]), 0)
        if tr is None:
            continue
        t, sites, lens = tr
        j = lens[0]                                     # first junction, transcript coordinates
        r = rng.random()
        if r < 0.3:
            start = j - rng.randrange(1, 12)             # 1-11 bases before the junction, rest beyond
        elif r < 0.6:
            start = j - rdlen + rng.randrange(1, 12)     # all but 1-11 bases before the junction
        else:
            start = rng.randrange(0, len(t) - rdlen)
        if start < 0 or start + rdlen > len(t):
            continue
        rd = mutate(rng, t[start:start + rdlen])
        se.append(rd if rng.random() < 0.5 else revcomp(rd))
        if rng.random() < 0.85:
            used.update(sites)
    while len(pairs) < n_pairs:
        chrom = rng.choice(names)
        tr = transcript(rng, chrom, refs[chrom], rng.choice([2, 2, 3]), 0)
        if tr is None:
            continue
        t, sites, lens = tr
        fl = rng.randrange(180, 420)
        if fl > len(t):
            continue
        start = rng.randrange(0, len(t) - fl + 1)
        frag = t[start:start + fl]
        m1, m2 = mutate(rng, frag[:rdlen]), mutate(rng, revcomp(frag[-rdlen:]))
        if rng.random() < 0.5:
            m1, m2 = m2, m1
        pairs.append((m1, m2))
        if rng.random() < 0.85:
            used.update(sites)
    sl = sorted(used)
    for _ in range(max(20, len(sl) // 4)):              # decoys
        chrom = rng.choice(names)
        p = rng.randrange(100, len(refs[chrom]) - 5000)
        sl.append((chrom, p, p + rng.randrange(50, 3000), rng.choice("+-")))
    rng.shuffle(sl)
    lines = ["%s\t%d\t%d\t%s" % s for s in sl[:20]] + ["chrNOPE\t10\t200\t+"] + ["%s\t%d\t%d\t%s" % s for s in sl]
    return se, pairs, lines


def write(prefix, se, pairs, lines, tag="k"):
    with open(prefix + ".fa", "w") as f:
        for i, r in enumerate(se):
            f.write(">%s%d\n%s\n" % (tag, i, r))
    with open(prefix + "_1.fa", "w") as f1, open(prefix + "_2.fa", "w") as f2:
        for i, (a, b) in enumerate(pairs):
            f1.write(">%sp%d/1\n%s\n" % (tag, i, a))
            f2.write(">%sp%d/2\n%s\n" % (tag, i, b))
    with open(prefix + "_ss.txt", "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    refs = load_fasta(sys.argv[1])
    n_se = int(sys.argv[3]) if len(sys.argv) > 3 else 600
    n_pairs = int(sys.argv[4]) if len(sys.argv) > 4 else 300
    seed = int(sys.argv[5]) if len(sys.argv) > 5 else 23
    se, pairs, lines = sim(refs, n_se, n_pairs, seed)
    write(sys.argv[2], se, pairs, lines)
    print(len(se), "reads,", len(pairs), "pairs,", len(lines), "site lines")
