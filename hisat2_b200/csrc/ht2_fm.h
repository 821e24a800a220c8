// ht2_fm.h -- FM-index search primitives over the packed image (host+device).
//
// One set of templates serves the global index (IT = uint32_t) and the local
// 56-kbp indexes (IT = uint16_t).  Semantics follow the reference functions
// cited at each definition; the code is written for the GPU (whole-side loads,
// branch-free popcount rank) rather than translated.
#ifndef HT2_FM_H_
#define HT2_FM_H_

#include "ht2_image.h"

#if defined(__CUDA_ARCH__)
#define HT2_POPC64(x) __popcll(x)
#define HT2_FFS64(x) __ffsll((long long)(x))
#else
#define HT2_POPC64(x) __builtin_popcountll(x)
#define HT2_FFS64(x) __builtin_ffsll((long long)(x))
#endif

#define HT2_IDX_MAX32 0xffffffffu

#if defined(__CUDACC__)
#define HT2_ALIGN16 __align__(16)
#else
#define HT2_ALIGN16 __attribute__((aligned(16)))
#endif

// View of one FM index inside the blob.
template <typename IT>
struct Ht2Fm {
    const Ht2Gfm* g;
    const uint8_t* gfm;
    const IT* ftab;
    const IT* eftab;
    const IT* offs;
    const IT* zoffs;
    const IT* rstarts;
    const IT* plen;
    uint32_t z0;          // the '$' row of a linear index
    uint32_t offMask, offRate;
    HT2_HD void init(const uint8_t* blob, const Ht2Gfm* geom) {
        g = geom;
        z0 = geom->zOff0;
        offMask = geom->offMask;
        offRate = geom->offRate;
        gfm = blob + geom->o_gfm;
        ftab = (const IT*)(blob + geom->o_ftab);
        eftab = (const IT*)(blob + geom->o_eftab);
        offs = (const IT*)(blob + geom->o_offs);
        zoffs = (const IT*)(blob + geom->o_zoffs);
        rstarts = (const IT*)(blob + geom->o_rstarts);
        plen = (const IT*)(blob + geom->o_plen);
    }
    static HT2_HD IT imax() { return (IT)~(IT)0; }
};

// Replicates the 2-bit code c across a 64-bit word.
HT2_HD uint64_t ht2_rep2(int c) {
    return 0x5555555555555555ull * (uint64_t)c;
}

// ---- linear indexes: 32-byte rank sides (ht2_image.h) ----------------------
struct HT2_ALIGN16 Ht2SideBwt { uint64_t lo, hi; };   // 64 BW chars

template <typename IT>
HT2_HD const uint8_t* ht2_side(const Ht2Fm<IT>& fm, uint32_t row) {
    return fm.gfm + ((uint64_t)(row >> HT2_SIDE_SHIFT) << 5);
}

// # occurrences of c among the first n (<64) chars of a side.
HT2_HD uint32_t ht2_count_side(const Ht2SideBwt& w, int c, uint32_t n) {
    const uint64_t p = ht2_rep2(c);
    uint64_t x0 = ~(w.lo ^ p), x1 = ~(w.hi ^ p);
    x0 = x0 & (x0 >> 1) & 0x5555555555555555ull;
    x1 = x1 & (x1 >> 1) & 0x5555555555555555ull;
    const uint64_t low = (1ull << ((n & 31) << 1)) - 1;
    const bool hi = n >= 32;
    const uint64_t m0 = hi ? ~0ull : low, m1 = hi ? low : 0ull;
    return (uint32_t)(HT2_POPC64(x0 & m0) + HT2_POPC64(x1 & m1));
}

HT2_HD int ht2_side_char(const Ht2SideBwt& w, uint32_t charOff) {
    const uint64_t v = (charOff & 32) ? w.hi : w.lo;
    return (int)((v >> ((charOff & 31) << 1)) & 3);
}

// BW character at 'row' (GFM::rowL).
template <typename IT>
HT2_HD int ht2_rowL(const Ht2Fm<IT>& fm, uint32_t row) {
    const uint32_t charOff = row & (HT2_SIDE_CHARS - 1);
    return (ht2_side(fm, row)[charOff >> 2] >> ((charOff & 3) << 1)) & 3;
}

// LF(row, c) = fchr[c] + occ(c, row)   (GFM::countBt2Side gfm.h:2958-2999,
// mapLF(l,c) gfm.h:3712-3732).  The '$' is stored as an 'A' and must not be
// counted (gfm.h:2967-2979); fchr[c] is folded into the side's occ entries.
template <typename IT>
HT2_HD uint32_t ht2_lf_side(const Ht2Fm<IT>& fm, const uint8_t* side, const Ht2SideBwt& w, uint32_t row, int c) {
    const uint32_t charOff = row & (HT2_SIDE_CHARS - 1);
    uint32_t cnt = ht2_count_side(w, c, charOff);
    // '$' inside [sideStart, row): unsigned compare also rejects z < sideStart
    if (c == 0 && (uint32_t)(fm.z0 - (row - charOff)) < charOff) cnt--;
    return (uint32_t)(IT)(((const uint32_t*)(side + 16))[c] + cnt);
}
template <typename IT>
HT2_HD uint32_t ht2_lf(const Ht2Fm<IT>& fm, uint32_t row, int c) {
    const uint8_t* side = ht2_side(fm, row);
    const Ht2SideBwt w = *(const Ht2SideBwt*)side;
    return ht2_lf_side(fm, side, w, row, c);
}
// Both boundaries of a range at once: the two side loads are issued together
// (one L2 round trip instead of two dependent ones) and the result is the plain
// two-boundary formula also for one-row ranges -- LF(top+1,c) - LF(top,c) is 1
// exactly when the row holds c (and is not '$'), which is what mapLF1's
// rowL test computes (gfm.h:3889-3911) -- so all lanes run the same code.
template <typename IT>
HT2_HD void ht2_lf2(const Ht2Fm<IT>& fm, uint32_t top, uint32_t bot, int c, uint32_t& ntop, uint32_t& nbot) {
    const uint8_t* st = ht2_side(fm, top);
    const uint8_t* sb = ht2_side(fm, bot);
    const Ht2SideBwt wt = *(const Ht2SideBwt*)st;
    const Ht2SideBwt wb = *(const Ht2SideBwt*)sb;
    const uint32_t ot = ((const uint32_t*)(st + 16))[c];
    const uint32_t ob = ((const uint32_t*)(sb + 16))[c];
    const uint32_t ct = top & (HT2_SIDE_CHARS - 1), cb = bot & (HT2_SIDE_CHARS - 1);
    uint32_t nt = ot + ht2_count_side(wt, c, ct);
    uint32_t nb = ob + ht2_count_side(wb, c, cb);
    if (c == 0) {
        nt -= ((uint32_t)(fm.z0 - (top - ct)) < ct) ? 1u : 0u;
        nb -= ((uint32_t)(fm.z0 - (bot - cb)) < cb) ? 1u : 0u;
    }
    ntop = (uint32_t)(IT)nt;
    nbot = (uint32_t)(IT)nb;
}

// c = rowL(row); returns LF(row, c)  (one side load; GFM::mapLF1 gfm.h:3889-3911).
template <typename IT>
HT2_HD uint32_t ht2_lf_own(const Ht2Fm<IT>& fm, uint32_t row, int& c) {
    const uint8_t* side = ht2_side(fm, row);
    const Ht2SideBwt w = *(const Ht2SideBwt*)side;
    c = ht2_side_char(w, row & (HT2_SIDE_CHARS - 1));
    return ht2_lf_side(fm, side, w, row, c);
}

template <typename IT>
HT2_HD bool ht2_is_zoff(const Ht2Fm<IT>& fm, uint32_t row) {
    return row == fm.z0;
}

// ftab lookup for the ftabChars bases seq[off..off+ftabChars) (GFM::ftabSeqToInt/ftabHi/ftabLo/ftabLoHi, gfm.h:2569-2715).
// Caller guarantees no Ns in the window.
template <typename IT>
HT2_HD void ht2_ftab_lohi(const Ht2Fm<IT>& fm, const uint8_t* seq, uint32_t off,
                          uint32_t& top, uint32_t& bot) {
    const uint32_t fc = fm.g->ftabChars;
    uint32_t fi = 0;
    for (uint32_t i = 0; i < fc; i++) fi = (fi << 2) | seq[off + i]; // fw index: left-to-right (gfm.h:2578-2587)
    IT hi = fm.ftab[fi];
    if (hi > fm.g->ftabCmp) hi = fm.eftab[(uint32_t)((IT)(hi ^ Ht2Fm<IT>::imax())) * 2 + 1];
    IT lo = fm.ftab[fi + 1];
    if (lo > fm.g->ftabCmp) lo = fm.eftab[(uint32_t)((IT)(lo ^ Ht2Fm<IT>::imax())) * 2];
    top = hi;
    bot = lo;
}

#endif // HT2_FM_H_
