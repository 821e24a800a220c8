/* simreads_fast.c -- seeded synthetic read generator for bench.py / the full-size tests:
 * fragment length U[fmin, fmax], rdlen-bp mates from both fragment ends (mate 2 reverse-complemented, --fr),
 * random strand swap, per-base substitution rate `sub`, fragments containing N are resampled, FASTA output
 * with fixed-width names (">r<digits>\n<bases>\n").  Fragment i depends only on (seed, i), so any shard of
 * the read set can be generated on its own.  Test / benchmark tooling, not part of the product. */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static inline uint64_t sm64(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline char comp(char c) { switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return 'N'; } }

/* ref: upper-case ASCII reference of length L.  Writes n records per mate starting at fragment `first`.
 * out1 / out2 must hold n * recsz bytes, recsz = 2 + digits + 1 + rdlen + 1.  out2 may be NULL. */
uint64_t ht2_simreads(const char* ref, uint64_t L, uint64_t first, uint64_t n, uint64_t seed, int rdlen, int fmin, int fmax,
                      double sub, int digits, char* out1, char* out2)
{
    const uint64_t recsz = (uint64_t)(2 + digits + 1 + rdlen + 1);
    const uint64_t subThr = (uint64_t)(sub * 16777216.0);
    for (uint64_t k = 0; k < n; k++) {
        const uint64_t id = first + k;
        uint64_t st = seed * 0xD1342543DE82EF95ull + id * 0x2545F4914F6CDD1Dull + 0x1234567ull;
        uint64_t pos, flen;
        for (;;) {
            flen = (uint64_t)fmin + sm64(&st) % (uint64_t)(fmax - fmin + 1);
            pos = sm64(&st) % (L - (uint64_t)fmax);
            int bad = 0;
            for (uint64_t i = 0; i < flen; i++) if (ref[pos + i] == 'N') { bad = 1; break; }
            if (!bad) break;
        }
        const int swap = (int)(sm64(&st) & 1);
        char* o[2] = {out1 + k * recsz, out2 ? out2 + k * recsz : NULL};
        for (int m = 0; m < 2; m++) {
            if (!o[m]) continue;
            char* p = o[m];
            *p++ = '>'; *p++ = 'r';
            uint64_t v = id;
            for (int d = digits - 1; d >= 0; d--) { p[d] = (char)('0' + v % 10); v /= 10; }
            p += digits; *p++ = '\n';
            /* mate 1: left end (plus strand) or revcomp of the right end (minus strand); mate 2 the other one */
            const int leftEnd = (m == 0) ? !swap : swap;
            const int rc = (m == 0) ? swap : !swap;
            const char* src = ref + (leftEnd ? pos : pos + flen - (uint64_t)rdlen);
            for (int i = 0; i < rdlen; i++) {
                char c = rc ? comp(src[rdlen - 1 - i]) : src[i];
                const uint64_t r = sm64(&st);
                if ((r & 0xFFFFFF) < subThr) {
                    const int code = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3;
                    c = "ACGT"[(code + 1 + (int)((r >> 24) % 3)) & 3];
                }
                p[i] = c;
            }
            p[rdlen] = '\n';
        }
    }
    return n * recsz;
}
