// ht2_graph.h -- graph-FM (GBWT) search primitives over the packed image (host+device).
//
// A graph index (hisat2-build --snp/--ss/--exon) keeps one BW row per INCOMING
// EDGE, listed node by node in node-rank order; F marks the first incoming edge
// of every node and M is the unary out-degree code (gbwt_graph.h:2621-2649).
// One search step on base c is
//     r'   = fchr[c] + rank_c(bwt, r)                (character rank)
//     node = rank1(M, r')                            (edge -> node id)
//     row  = select1(F, node + 1)                    (node -> its first BW row)
// and ranges carry both [top,bot) rows and [node_top,node_bot) nodes.
//
// Inside the image graph indexes are re-laid as 64-byte rank sides of 64 rows (ht2_image.h,
// ht2_index.cpp:relayGraph) -- the .ht2 format keeps 208 rows per 128-byte side (a3), i.e. a
// div/mod per rank and byte loops over 26-byte bit arrays:
//   [16 B BW chars][u64 F][u64 M][u32 occ[4] incl. fchr][u32 M_occ][u32 F_loc][pad]
// M_occ = rank1(M, sideStart); F_loc = select1(F, M_occ).  One search boundary touches the
// character side of the row, the M side of the LF result and the F side(s) at F_loc.
// Reference functions restated here: GFM::mapGLF (gfm.h:3759-3837), mapGLF1
// (:3957-4095), mapLF1 (:3889-3950), countBt2Side (:2958-2999), rank_M (:4100),
// countMSide (:3146), select_F (:4113-4168), getInEdgeCount (:4172-4210),
// tryOffset/getOffset (:2719, :5682-5716).
#ifndef HT2_GRAPH_H_
#define HT2_GRAPH_H_

#include "ht2_fm.h"

struct HT2_ALIGN16 Ht2GSide {
    Ht2SideBwt bwt;        // 64 BW chars
    uint64_t   F, M;       // one bit per row
    uint32_t   occ[4];     // fchr[c] + #c in rows [0, sideStart), '$' rows not counted
    uint32_t   M_occ;      // rank1(M, sideStart)
    uint32_t   F_loc;      // select1(F, M_occ): row of the M_occ-th set F bit
    uint32_t   pad[2];
};

template <typename IT>
HT2_HD const Ht2GSide* ht2g_side(const Ht2Fm<IT>& fm, uint32_t row) {
    return (const Ht2GSide*)(fm.gfm + ((uint64_t)(row >> HT2_SIDE_SHIFT) << 6));
}

template <typename IT>
HT2_HD int ht2g_rowL(const Ht2Fm<IT>& fm, uint32_t row) {
    return ht2_side_char(ht2g_side(fm, row)->bwt, row & (HT2_SIDE_CHARS - 1));
}

template <typename IT>
HT2_HD bool ht2g_is_zoff(const Ht2Fm<IT>& fm, uint32_t row) {
    if (row == fm.z0) return true;
    for (uint32_t i = 1; i < fm.g->nzOffs; i++) if (row == fm.zoffs[i]) return true;
    return false;
}

// fchr[c] + occ(c, row): countBt2Side with the '$' rows (several on a graph) not counted as 'A'.
template <typename IT>
HT2_HD uint32_t ht2g_lf(const Ht2Fm<IT>& fm, uint32_t row, int c) {
    const Ht2GSide* sd = ht2g_side(fm, row);
    const uint32_t charOff = row & (HT2_SIDE_CHARS - 1);
    const Ht2SideBwt w = sd->bwt;
    uint32_t cnt = sd->occ[c] + ht2_count_side(w, c, charOff);
    if (c == 0) {
        const uint32_t sideStart = row - charOff;
        if ((uint32_t)(fm.z0 - sideStart) < charOff) cnt--;
        for (uint32_t i = 1; i < fm.g->nzOffs; i++) if ((uint32_t)((uint32_t)fm.zoffs[i] - sideStart) < charOff) cnt--;
    }
    return (uint32_t)(IT)cnt;
}

// rank1(M, row) = # set M bits in rows [0,row).
template <typename IT>
HT2_HD uint32_t ht2g_rank_M(const Ht2Fm<IT>& fm, uint32_t row) {
    const Ht2GSide* sd = ht2g_side(fm, row);
    const uint32_t k = row & (HT2_SIDE_CHARS - 1);
    return (uint32_t)(IT)(sd->M_occ + (uint32_t)HT2_POPC64(sd->M & ((1ull << k) - 1)));
}

// position (0-based) of the n-th (1-based, n <= popcount) set bit of x
HT2_HD uint32_t ht2g_select64(uint64_t x, uint32_t n) {
    uint32_t pos = 0;
    uint32_t lo = (uint32_t)x, c = (uint32_t)HT2_POPC64((uint64_t)lo);
    if (n > c) { n -= c; pos = 32; lo = (uint32_t)(x >> 32); }
    c = (uint32_t)HT2_POPC64((uint64_t)(lo & 0xffffu)); if (n > c) { n -= c; pos += 16; lo >>= 16; } lo &= 0xffffu;
    c = (uint32_t)HT2_POPC64((uint64_t)(lo & 0xffu));   if (n > c) { n -= c; pos += 8;  lo >>= 8; }  lo &= 0xffu;
    c = (uint32_t)HT2_POPC64((uint64_t)(lo & 0xfu));    if (n > c) { n -= c; pos += 4;  lo >>= 4; }  lo &= 0xfu;
    c = (uint32_t)HT2_POPC64((uint64_t)(lo & 0x3u));    if (n > c) { n -= c; pos += 2;  lo >>= 2; }  lo &= 0x3u;
    if (n > (lo & 1u)) pos += 1;
    return pos;
}

// Row of the count-th (>= 1) set F bit at or after 'row' (select_F walks forward across sides).
template <typename IT>
HT2_NI uint32_t ht2g_select_F(const Ht2Fm<IT>& fm, uint32_t row, uint32_t count) {
    uint32_t s = row >> HT2_SIDE_SHIFT;
    const uint32_t lastSide = fm.g->numSides;    // one zeroed slack side follows the last side
    uint64_t bits = ((const Ht2GSide*)(fm.gfm + ((uint64_t)s << 6)))->F >> (row & (HT2_SIDE_CHARS - 1));
    uint32_t base = row;
    while (true) {
        const uint32_t pc = (uint32_t)HT2_POPC64(bits);
        if (pc >= count) return base + ht2g_select64(bits, count);
        count -= pc;
        s++;
        if (s > lastSide) return fm.g->gbwtLen;   // ran off the end (cannot happen on a well-formed index)
        base = s << HT2_SIDE_SHIFT;
        bits = ((const Ht2GSide*)(fm.gfm + ((uint64_t)s << 6)))->F;
    }
}

// Edge row r (an LF result) -> (node id, first BW row of that node); the tail shared by mapGLF's top
// boundary and mapGLF1 (gfm.h:3786-3807, 3978-4000).  The reference steps back a side when its
// trailer's M_occ exceeds the node; with F_loc = select1(F, M_occ) that case is F_loc itself.
// Also returns where a further select can start and the M_occ it is relative to, so that mapGLF1
// selects the NEXT node's first row from the same place.
template <typename IT>
HT2_HD void ht2g_edge_to_node(const Ht2Fm<IT>& fm, uint32_t r, uint32_t& node, uint32_t& firstRow,
                              uint32_t& F_loc, uint32_t& M_occ) {
    const Ht2GSide* sd = ht2g_side(fm, r + 1);
    const uint32_t k = (r + 1) & (HT2_SIDE_CHARS - 1);
    M_occ = sd->M_occ;
    F_loc = sd->F_loc;
    node = (uint32_t)(IT)(M_occ + (uint32_t)HT2_POPC64(sd->M & ((1ull << k) - 1)) - 1);
    if (node + 1 == M_occ) { firstRow = F_loc; return; }                     // the M_occ-th node itself
    const uint32_t start = M_occ > 0 ? F_loc + 1 : 0;
    firstRow = ht2g_select_F(fm, start, node + 1 - M_occ);
}

// GFM::getInEdgeCount: for the rows [top,bot) (top is a node's first row) list
// (node index within the range, # extra incoming edges) for nodes with > 1 row.
// Returns the number of entries written (cap entries at most; excess sets overflow).
template <typename IT>
HT2_HD uint32_t ht2g_in_edge_count(const Ht2Fm<IT>& fm, uint32_t top, uint32_t bot, uint16_t (*out)[2], uint32_t cap, bool& overflow) {
    uint32_t n = 0, curr = 0, num0s = 0;
    bool curOk = false;
    for (uint32_t r = top + 1; r < bot; r++) {
        const int bit = (int)((ht2g_side(fm, r)->F >> (r & (HT2_SIDE_CHARS - 1))) & 1);
        if (bit) { curr++; num0s = 0; }
        else {
            num0s++;
            if (num0s == 1) {
                curOk = n < cap;
                if (curOk) { out[n][0] = (uint16_t)curr; out[n][1] = 0; n++; } else overflow = true;
            }
            if (curOk) out[n - 1][1] = (uint16_t)num0s;
        }
    }
    return n;
}

#define HT2G_MAX_IEDGES 24

// GFM::mapGLF.  k = kseeds: the in-edge list is only built for node ranges <= k.
template <typename IT>
HT2_NI void ht2g_mapGLF(const Ht2Fm<IT>& fm, uint32_t top, uint32_t bot, int c, uint32_t k,
                        uint32_t& ntop, uint32_t& nbot, uint32_t& node_top, uint32_t& node_bot,
                        uint16_t (*iedges)[2], uint32_t& niedges, bool& overflow) {
    niedges = 0;
    uint32_t t = ht2g_lf(fm, top, c);
    uint32_t b = ht2g_lf(fm, bot, c);
    if (t + 1 >= fm.g->gbwtLen || t >= b) { ntop = nbot = node_top = node_bot = 0; return; }
    uint32_t F_loc, M_occ;
    ht2g_edge_to_node(fm, t, node_top, ntop, F_loc, M_occ);
    {
        const Ht2GSide* sd = ht2g_side(fm, b);
        const uint32_t kk = b & (HT2_SIDE_CHARS - 1);
        const uint32_t bM = sd->M_occ;
        node_bot = (uint32_t)(IT)(bM + (uint32_t)HT2_POPC64(sd->M & ((1ull << kk) - 1)));
        const uint32_t start = bM > 0 ? sd->F_loc + 1 : 0;
        nbot = ht2g_select_F(fm, start, node_bot + 1 - bM);     // node_bot + 1 > bM always (gfm.h:3816)
    }
    if (node_bot - node_top <= k && node_bot - node_top < nbot - ntop)
        niedges = ht2g_in_edge_count(fm, ntop, nbot, iedges, HT2G_MAX_IEDGES, overflow);
}

// GFM::mapGLF1(row, l, c): one-row range extended with base c.
template <typename IT>
HT2_NI void ht2g_mapGLF1c(const Ht2Fm<IT>& fm, uint32_t row, int c,
                          uint32_t& ntop, uint32_t& nbot, uint32_t& node_top, uint32_t& node_bot) {
    if (ht2g_rowL(fm, row) != c || ht2g_is_zoff(fm, row)) { ntop = nbot = node_top = node_bot = 0; return; }
    const uint32_t t = ht2g_lf(fm, row, c);
    uint32_t F_loc, M_occ;
    ht2g_edge_to_node(fm, t, node_top, ntop, F_loc, M_occ);
    node_bot = node_top + 1;
    nbot = ht2g_select_F(fm, M_occ > 0 ? F_loc + 1 : 0, node_bot + 1 - M_occ);   // node_bot + 1 > M_occ always
}

// GFM::mapGLF1(row, l): follow the row's own character.  Returns false on a '$' row.
template <typename IT>
HT2_NI bool ht2g_mapGLF1(const Ht2Fm<IT>& fm, uint32_t row, uint32_t& ntop, uint32_t& node_top) {
    if (ht2g_is_zoff(fm, row)) return false;
    const int c = ht2g_rowL(fm, row);
    const uint32_t t = ht2g_lf(fm, row, c);
    uint32_t F_loc, M_occ;
    ht2g_edge_to_node(fm, t, node_top, ntop, F_loc, M_occ);
    return true;
}

// GFM::getOffset(row, node): joined-text offset of a node reached through BW row 'row'.
template <typename IT>
HT2_NI uint32_t ht2g_get_offset(const Ht2Fm<IT>& fm, uint32_t row, uint32_t node, uint32_t& nsteps) {
    nsteps = 0;
    if (ht2g_is_zoff(fm, row)) return 0;
    if ((node & fm.offMask) == node) {
        const IT off = fm.offs[node >> fm.offRate];
        if (off != Ht2Fm<IT>::imax()) return off;
    }
    uint32_t jumps = 0;
    while (true) {
        uint32_t nrow, nnode;
        ht2g_mapGLF1(fm, row, nrow, nnode);
        jumps++;
        row = nrow;
        if (ht2g_is_zoff(fm, row)) { nsteps = jumps; return (uint32_t)(IT)jumps; }
        if ((nnode & fm.offMask) == nnode) {
            const IT off = fm.offs[nnode >> fm.offRate];
            if (off != Ht2Fm<IT>::imax()) { nsteps = jumps; return (uint32_t)(IT)(jumps + off); }
        }
    }
}

#endif // HT2_GRAPH_H_
