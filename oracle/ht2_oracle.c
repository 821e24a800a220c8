/* oracle/ht2_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of the index-search primitives of
 * HISAT2's alignment hot path, written independently of hisat2_b200/csrc so
 * that the CUDA kernels can be checked against something that is neither
 * themselves nor a black box.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load this file's shared library.
 *
 * Parity status: PINNED (linear and graph indexes).  `ht2_oracle dump`
 * reproduces, byte for byte, the output of oracle/_ref/ref_dump (the unmodified reference's own
 * HI_Aligner::partialSearch / GFM::getOffset / joinedToTextOff driven over the
 * same reads); tests/test_oracle.py checks the committed golden dump
 * tests/golden/tiny_dump.txt.  The full per-read policy (go(), SAM) is pinned
 * against the reference binary itself (oracle/_ref/hisat2-align-s), not here.
 *
 * Each function cites the reference lines it restates.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    /* header (gfm.h:5917-5979; GFMParams::init gfm.h:134-176) */
    uint32_t len, gbwtLen, numNodes, eftabLen;
    int32_t lineRate, offRate, ftabChars;
    int linear;
    uint32_t sideSz, sideBwtSz, sideBwtLen, numSides, ftabLen, offsLen;
    uint32_t nPat, nFrag, nz;
    uint32_t *plen, *rstarts, *zoffs, *ftab, *eftab, *offs;
    uint32_t fchr[5];
    uint8_t* bwt;
    uint32_t minK;
    /* graph indexes only: prefix tables built once from the F and M bit arrays */
    uint32_t* mrank;   /* mrank[r] = # set M bits in rows [0,r)          (rank_M, gfm.h:4100) */
    uint32_t* fsel;    /* fsel[n]  = row of the n-th (1-based) set F bit (select_F, gfm.h:4113) */
    uint32_t nF;
} ht2o_index;

static uint8_t* slurp(const char* path, size_t* n) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t* b = (uint8_t*)malloc((size_t)sz + 16);
    if (fread(b, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(b); return NULL; }
    fclose(f);
    *n = (size_t)sz;
    return b;
}
static uint32_t rd32(const uint8_t** p) { uint32_t v; memcpy(&v, *p, 4); *p += 4; return v; }

/* GFM::readIntoMemory, global index (gfm.h:5917-6458) */
ht2o_index* ht2o_open(const char* base) {
    char path[4096];
    size_t n1 = 0, n2 = 0;
    snprintf(path, sizeof(path), "%s.1.ht2", base);
    uint8_t* d1 = slurp(path, &n1);
    snprintf(path, sizeof(path), "%s.2.ht2", base);
    uint8_t* d2 = slurp(path, &n2);
    if (!d1 || !d2) return NULL;
    ht2o_index* ix = (ht2o_index*)calloc(1, sizeof(ht2o_index));
    const uint8_t* p = d1;
    if (rd32(&p) != 1) return NULL;
    rd32(&p);
    ix->len = rd32(&p); ix->gbwtLen = rd32(&p); ix->numNodes = rd32(&p);
    ix->lineRate = (int32_t)rd32(&p); rd32(&p);
    ix->offRate = (int32_t)rd32(&p); ix->ftabChars = (int32_t)rd32(&p);
    ix->eftabLen = rd32(&p); rd32(&p);
    ix->linear = (ix->len + 1 == ix->gbwtLen || ix->gbwtLen == 0);
    if (ix->gbwtLen == 0) ix->gbwtLen = ix->len + 1;
    if (ix->numNodes == 0) ix->numNodes = ix->len + 1;
    ix->sideSz = 1u << ix->lineRate;
    uint32_t gbwtSz;
    if (ix->linear) { ix->sideBwtSz = ix->sideSz - 16; ix->sideBwtLen = ix->sideBwtSz * 4; gbwtSz = ix->gbwtLen / 4 + 1; }
    else { ix->sideBwtSz = ix->sideSz - 24; ix->sideBwtLen = ix->sideBwtSz * 2; gbwtSz = ix->gbwtLen / 2 + 1; }
    ix->numSides = (gbwtSz + ix->sideBwtSz - 1) / ix->sideBwtSz;
    ix->ftabLen = (1u << (2 * ix->ftabChars)) + 1;
    ix->offsLen = (ix->numNodes + (1u << ix->offRate) - 1) >> ix->offRate;
    ix->nPat = rd32(&p); ix->plen = (uint32_t*)p; p += 4 * (size_t)ix->nPat;
    ix->nFrag = rd32(&p); ix->rstarts = (uint32_t*)p; p += 12 * (size_t)ix->nFrag;
    ix->bwt = (uint8_t*)p; p += (size_t)ix->numSides * ix->sideSz;
    ix->nz = rd32(&p); ix->zoffs = (uint32_t*)p; p += 4 * (size_t)ix->nz;
    for (int i = 0; i < 5; i++) ix->fchr[i] = rd32(&p);
    ix->ftab = (uint32_t*)p; p += 4 * (size_t)ix->ftabLen;
    ix->eftab = (uint32_t*)p;
    ix->offs = (uint32_t*)(d2 + 4);
    if (!ix->linear) {
        /* Side layout of a graph index (gfm.h:3394-3398, 3146-3154): F bits at side + sideBwtSz/2,
         * M bits at side + sideBwtSz*3/4, one bit per BW row.  The tables below are the plain
         * definitions of rank1(M) and select1(F); the reference (and the CUDA path) get the same
         * values from the per-side F_loc / M_occ trailer entries. */
        ix->mrank = (uint32_t*)calloc((size_t)ix->gbwtLen + 2, 4);
        ix->fsel = (uint32_t*)calloc((size_t)ix->gbwtLen + 3, 4);
        uint32_t m = 0, nf = 0;
        for (uint32_t r = 0; r < ix->gbwtLen; r++) {
            const uint8_t* side = ix->bwt + (size_t)(r / ix->sideBwtLen) * ix->sideSz;
            uint32_t k = r % ix->sideBwtLen;
            const uint8_t* fb = side + (ix->sideBwtSz >> 1);
            const uint8_t* mb = side + (ix->sideBwtSz - (ix->sideBwtSz >> 2));
            ix->mrank[r] = m;
            if ((mb[k >> 3] >> (k & 7)) & 1) m++;
            if ((fb[k >> 3] >> (k & 7)) & 1) ix->fsel[++nf] = r;
        }
        ix->mrank[ix->gbwtLen] = m; ix->mrank[ix->gbwtLen + 1] = m;
        ix->nF = nf;
        for (uint32_t n = nf + 1; n < ix->gbwtLen + 3; n++) ix->fsel[n] = ix->gbwtLen;
    }
    /* HI_Aligner ctor (hi_aligner.h:3979-3984) */
    uint32_t g = ix->len;
    ix->minK = 0;
    while (g > 0) { g >>= 2; ix->minK++; }
    return ix;
}

/* BW char of a row (GFM::rowL) */
static int bw_char(const ht2o_index* ix, uint32_t row) {
    const uint8_t* side = ix->bwt + (size_t)(row / ix->sideBwtLen) * ix->sideSz;
    uint32_t k = row % ix->sideBwtLen;
    return (side[k >> 2] >> ((k & 3) * 2)) & 3;
}
static int is_z(const ht2o_index* ix, uint32_t row) {
    for (uint32_t i = 0; i < ix->nz; i++) if (ix->zoffs[i] == row) return 1;
    return 0;
}
/* LF(row,c): char-by-char count inside the side, the slow obvious way
 * (GFM::countBt2Side gfm.h:2958-2999 / countUpTo :3166-3226). */
uint32_t ht2o_lf(const ht2o_index* ix, uint32_t row, int c) {
    uint32_t s = row / ix->sideBwtLen, k = row % ix->sideBwtLen;
    const uint8_t* side = ix->bwt + (size_t)s * ix->sideSz;
    uint32_t cnt = 0;
    for (uint32_t i = 0; i < k; i++) {
        int ch = (side[i >> 2] >> ((i & 3) * 2)) & 3;
        if (ch == c) {
            if (c == 0 && is_z(ix, s * ix->sideBwtLen + i)) continue; /* '$' stored as 'A' */
            cnt++;
        }
    }
    const uint32_t* acgt = (const uint32_t*)(side + ix->sideBwtSz + (ix->linear ? 0 : 8));
    return acgt[c] + cnt + ix->fchr[c];
}

/* GFM::ftabLoHi (gfm.h:2569-2715) */
static uint32_t ftab_entry(const ht2o_index* ix, uint32_t i, int hi) {
    uint32_t cmp = ix->linear ? ix->len : ix->gbwtLen;
    if (ix->ftab[i] <= cmp) return ix->ftab[i];
    uint32_t e = ix->ftab[i] ^ 0xffffffffu;
    return ix->eftab[e * 2 + (hi ? 1 : 0)];
}
void ht2o_ftab(const ht2o_index* ix, const uint8_t* seq, uint32_t off, uint32_t* top, uint32_t* bot) {
    uint32_t fi = 0;
    for (int i = 0; i < ix->ftabChars; i++) fi = (fi << 2) | seq[off + i];
    *top = ftab_entry(ix, fi, 1);
    *bot = ftab_entry(ix, fi + 1, 0);
}

/* one backward-extension step on a linear index (mapLF gfm.h:3739-3752 for
 * ranges, mapGLF1/mapLF1 gfm.h:3957-3975, 3889-3912 for single rows) */
static void step(const ht2o_index* ix, uint32_t top, uint32_t bot, int c, uint32_t* nt, uint32_t* nb) {
    if (bot - top != 1) { *nt = ht2o_lf(ix, top, c); *nb = ht2o_lf(ix, bot, c); return; }
    if (bw_char(ix, top) != c || is_z(ix, top)) { *nt = *nb = 0; return; }
    *nt = ht2o_lf(ix, top, c);
    *nb = *nt + 1;
}

#define HT2O_MAX_IE 64
typedef struct { uint32_t first, second; } ht2o_ie;

/* GFM::getInEdgeCount (gfm.h:4172-4210): nodes of [top,bot) with more than one incoming edge */
static uint32_t in_edge_count(const ht2o_index* ix, uint32_t top, uint32_t bot, ht2o_ie* out) {
    uint32_t n = 0, curr = 0, num0s = 0;
    for (uint32_t r = top + 1; r < bot; r++) {
        const uint8_t* side = ix->bwt + (size_t)(r / ix->sideBwtLen) * ix->sideSz;
        uint32_t k = r % ix->sideBwtLen;
        int bit = ((side + (ix->sideBwtSz >> 1))[k >> 3] >> (k & 7)) & 1;
        if (bit) { curr++; num0s = 0; }
        else {
            num0s++;
            if (num0s == 1 && n < HT2O_MAX_IE) { out[n].first = curr; n++; }
            if (n > 0) out[n - 1].second = num0s;
        }
    }
    return n;
}
/* GFM::mapGLF (gfm.h:3759-3837) and mapGLF1 with a base (gfm.h:3957-4020), from the definitions:
 * node = rank1(M, r'+1) - 1, first row of node k = select1(F, k+1). */
static void gstep(const ht2o_index* ix, uint32_t top, uint32_t bot, int c, uint32_t k, uint32_t* nt, uint32_t* nb,
                  uint32_t* nnt, uint32_t* nnb, ht2o_ie* ie, uint32_t* nie) {
    *nie = 0;
    if (bot - top != 1) {
        uint32_t t = ht2o_lf(ix, top, c), b = ht2o_lf(ix, bot, c);
        if (t + 1 >= ix->gbwtLen || t >= b) { *nt = *nb = *nnt = *nnb = 0; return; }
        *nnt = ix->mrank[t + 1] - 1;
        *nt = ix->fsel[*nnt + 1];
        *nnb = ix->mrank[b];
        *nb = ix->fsel[*nnb + 1];
        if (*nnb - *nnt <= k && *nnb - *nnt < *nb - *nt) *nie = in_edge_count(ix, *nt, *nb, ie);
        return;
    }
    if (bw_char(ix, top) != c || is_z(ix, top)) { *nt = *nb = *nnt = *nnb = 0; return; }
    uint32_t t = ht2o_lf(ix, top, c);
    *nnt = ix->mrank[t + 1] - 1;
    *nt = ix->fsel[*nnt + 1];
    *nnb = *nnt + 1;
    *nb = ix->fsel[*nnb + 1];
    if (*nt + 1 < *nb) { ie[0].first = 0; ie[0].second = *nb - *nt - 1; *nie = 1; }   /* hi_aligner.h:6476-6482 */
}

typedef struct { uint32_t bwoff, len, top, bot, type, pseudo, anchor, node_top, node_bot, nie; ht2o_ie ie[HT2O_MAX_IE]; } ht2o_hit;

/* HI_Aligner::partialSearch (hi_aligner.h:6361-6600), linear and graph indexes.
 * Advances *cur / *done like the reference; returns the hit pushed. */
ht2o_hit ht2o_partial_search(const ht2o_index* ix, const uint8_t* seq, uint32_t len, uint32_t* cur, int* done,
                             int pseudo_in, int anchor_in, uint32_t khits) {
    ht2o_hit h;
    const uint32_t ftabLen = (uint32_t)ix->ftabChars, minK = ix->minK;
    const uint32_t kseeds = khits * 2 > 5 ? khits * 2 : 5;
    int pseudo_ = pseudo_in, anchor_ = anchor_in;
    uint32_t offset = *cur, dep = offset, left = len - dep;
    h.bwoff = offset; h.top = h.bot = h.node_top = h.node_bot = 0xffffffffu; h.type = 1; h.pseudo = h.anchor = 0; h.nie = 0;
    if (left < ftabLen + 1) { *cur = len; h.len = *cur - offset; *done = 1; return h; }
    for (uint32_t i = 0; i < ftabLen; i++) {
        if (seq[len - dep - 1 - i] > 3) {
            *cur += i + 1; h.len = *cur - offset;
            if (*cur >= len) *done = 1;
            return h;
        }
    }
    uint32_t top, bot;
    ht2o_ftab(ix, seq, len - dep - ftabLen, &top, &bot);
    dep += ftabLen;
    if (top >= bot) { *cur = dep; h.len = *cur - offset; if (*cur >= len) *done = 1; return h; }
    uint32_t ow_top = 0, ow_bot = 0; /* node_range starts (0,0) */
    uint32_t same_range = 0, similar_range = 0;
    uint32_t k5 = khits < 5 ? khits : 5;
    ht2o_ie ie[HT2O_MAX_IE], tie[HT2O_MAX_IE];
    uint32_t nie = 0, ntie = 0;
    while (dep < len) {
        int c = seq[len - dep - 1];
        uint32_t nt = 0, nb = 0, nnt = 0, nnb = 0;
        ntie = 0;
        if (c <= 3) {
            if (ix->linear) { step(ix, top, bot, c, &nt, &nb); nnt = nt; nnb = nb; }
            else gstep(ix, top, bot, c, kseeds, &nt, &nb, &nnt, &nnb, tie, &ntie);
        }
        if (nt >= nb) break;
        uint32_t nw = nnb - nnt, ow = ow_bot - ow_top;
        if (pseudo_) {
            if (nw < ow && ow <= k5) {
                if (dep - offset >= minK + 6 && similar_range >= 5) { h.pseudo = 1; break; }
            }
            if (nw != 1) {
                if (nw + 2 >= ow) similar_range++;
                else if (nw + 4 < ow) similar_range = 0;
            } else pseudo_ = 0;
        }
        if (anchor_) {
            if (nw != 1 && ow == nw) { same_range++; if (same_range >= 5) anchor_ = 0; }
            else same_range = 0;
            if (dep - offset >= minK + 8 && nw >= 4) anchor_ = 0;
        }
        top = nt; bot = nb; ow_top = nnt; ow_bot = nnb;
        nie = ntie; memcpy(ie, tie, sizeof(ht2o_ie) * ntie);
        dep++;
        if (anchor_ && dep - offset >= minK + 12 && bot - top == 1) { h.anchor = 1; break; }
    }
    /* a hit made of the ftab lookup alone keeps node_range == (0,0) and is pushed with blank
     * coordinates; so is a hit whose node range is narrower than its row range without an
     * in-edge list (hi_aligner.h:6550-6590) */
    int report = ow_top < ow_bot;
    if (ow_bot - ow_top < bot - top && nie == 0) report = 0;
    if (report) {
        h.top = top; h.bot = bot; h.node_top = ow_top; h.node_bot = ow_bot;
        h.nie = nie; memcpy(h.ie, ie, sizeof(ht2o_ie) * nie);
    }
    h.len = dep - offset;
    h.type = h.anchor ? 3 : (h.pseudo ? 2 : 1);
    *cur = dep;
    if (*cur >= len) *done = 1;
    return h;
}

/* GFM::getOffset(row, node) on a graph (gfm.h:5682-5716): follow each row's own edge label
 * (mapGLF1 without a base, gfm.h:4030-4095) until a sampled node or a '$' row is met. */
uint32_t ht2o_resolve_graph(const ht2o_index* ix, uint32_t row, uint32_t node) {
    const uint32_t mask = 0xffffffffu << ix->offRate;
    if (is_z(ix, row)) return 0;
    if ((node & mask) == node && ix->offs[node >> ix->offRate] != 0xffffffffu) return ix->offs[node >> ix->offRate];
    uint32_t jumps = 0;
    for (;;) {
        uint32_t t = ht2o_lf(ix, row, bw_char(ix, row));
        node = ix->mrank[t + 1] - 1;
        row = ix->fsel[node + 1];
        jumps++;
        if (is_z(ix, row)) return jumps;
        if ((node & mask) == node && ix->offs[node >> ix->offRate] != 0xffffffffu) return jumps + ix->offs[node >> ix->offRate];
    }
}

/* GFM::getOffset / tryOffset (gfm.h:5682-5716, 2719-2734) */
uint32_t ht2o_resolve(const ht2o_index* ix, uint32_t row) {
    uint32_t steps = 0;
    for (;;) {
        if (is_z(ix, row)) return steps;
        if ((row & (0xffffffffu << ix->offRate)) == row) return ix->offs[row >> ix->offRate] + steps;
        row = ht2o_lf(ix, row, bw_char(ix, row));
        steps++;
    }
}
/* GFM::joinedToTextOff (gfm.h:5527-5600), rejectStraddle = false */
int ht2o_joined_to_text(const ht2o_index* ix, uint32_t off, uint32_t* tidx, uint32_t* toff) {
    uint32_t top = 0, bot = ix->nFrag, elt = 0xffffffffu;
    for (;;) {
        uint32_t old = elt;
        elt = top + ((bot - top) >> 1);
        if (old == elt) { *tidx = 0xffffffffu; return 0; }
        uint32_t lower = ix->rstarts[elt * 3];
        uint32_t upper = (elt == ix->nFrag - 1) ? ix->len : ix->rstarts[(elt + 1) * 3];
        if (lower <= off) {
            if (upper > off) {
                *tidx = ix->rstarts[elt * 3 + 1];
                *toff = off - lower + ix->rstarts[elt * 3 + 2];
                return 1;
            }
            top = elt;
        } else bot = elt;
    }
}

/* ---- `ht2_oracle dump`: same record format as oracle/ref_dump.cpp ---------- */
#ifdef HT2_ORACLE_MAIN
static int dnacat(int c) { return strchr("ACGTacgtBDHKMNRSVWXYbdhkmnrsvwxy-", c) != NULL && c != 0; }
static uint8_t asc2dna(int c) {
    switch (c) { case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; case 'N': case 'n': return 4; default: return 0; }
}
static void dump_read(const ht2o_index* ix, const uint8_t* fwseq, uint32_t len, size_t rdid, int no_spliced) {
    uint8_t* rc = (uint8_t*)malloc(len + 1);
    for (uint32_t i = 0; i < len; i++) { uint8_t c = fwseq[len - i - 1]; rc[i] = c < 4 ? (uint8_t)(3 - c) : 4; }
    uint32_t khits = ix->linear ? 5 : 10;
    for (int fwi = 0; fwi < 2; fwi++) {
        const uint8_t* seq = fwi == 0 ? fwseq : rc;
        uint32_t cur = 0; int done = 0; unsigned nh = 0;
        while (!done) {
            ht2o_hit h = ht2o_partial_search(ix, seq, len, &cur, &done, ix->linear && !no_spliced, 1, khits);
            printf("H %zu %d %u %u %u %u %u %u %u\n", rdid, fwi == 0, h.bwoff, h.len, h.top, h.bot, h.type, h.pseudo, h.anchor);
            const int blank = (h.top == 0xffffffffu);
            if (!ix->linear && !blank) {
                printf("G %zu %d %u %u %u %u", rdid, fwi == 0, nh, h.node_top, h.node_bot, h.nie);
                for (uint32_t e = 0; e < h.nie; e++) printf(" %u:%u", h.ie[e].first, h.ie[e].second);
                printf("\n");
            }
            if (!blank && h.node_bot - h.node_top <= 4) {
                /* element i of the node range = first BW row of node i (group_walk.h:545-560) */
                uint32_t num_iedges = 0, e = 0;
                for (uint32_t i = 0; i < h.node_bot - h.node_top; i++) {
                    while (e < h.nie) { if (i <= h.ie[e].first) break; num_iedges += h.ie[e].second; e++; }
                    uint32_t r = h.top + i + num_iedges;
                    uint32_t j = ix->linear ? ht2o_resolve(ix, r) : ht2o_resolve_graph(ix, r, h.node_top + i), tidx = 0, toff = 0;
                    ht2o_joined_to_text(ix, j, &tidx, &toff);
                    printf("C %zu %d %u %u %u %u %u\n", rdid, fwi == 0, nh, r, j, tidx, toff);
                }
            }
            nh++;
            if (done) break;
            if (!h.pseudo) { if (cur + 1 < len) cur++; }
        }
    }
    free(rc);
}
int main(int argc, char** argv) {
    if (argc < 5 || strcmp(argv[1], "dump")) { fprintf(stderr, "usage: ht2_oracle dump <index> <reads.fa> <no_spliced>\n"); return 2; }
    ht2o_index* ix = ht2o_open(argv[2]);
    if (!ix) { fprintf(stderr, "cannot open index\n"); return 1; }
    FILE* f = fopen(argv[3], "r");
    if (!f) return 1;
    int no_spliced = atoi(argv[4]);
    char* line = NULL; size_t cap = 0; ssize_t n;
    uint8_t* seq = (uint8_t*)malloc(1 << 20); uint32_t len = 0; size_t rdid = 0; int have = 0;
    while ((n = getline(&line, &cap, f)) >= 0) {
        if (line[0] == '>') { if (have) { if (len) dump_read(ix, seq, len, rdid, no_spliced); rdid++; } have = 1; len = 0; }
        else for (ssize_t i = 0; i < n; i++) if (dnacat((unsigned char)line[i])) seq[len++] = asc2dna(line[i]);
    }
    if (have) { if (len) dump_read(ix, seq, len, rdid, no_spliced); }
    return 0;
}
#endif
