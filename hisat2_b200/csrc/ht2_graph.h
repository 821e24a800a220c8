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
// Graph indexes stay in the .ht2 side format inside the image (a3):
//   [sideGbwtSz/2 B 2-bit BW chars][sideGbwtSz/4 B F bits][sideGbwtSz/4 B M bits]
//   [F_loc][M_occ][A][C][G][T]                        (entries are IT wide)
// Reference functions restated here: GFM::mapGLF (gfm.h:3759-3837), mapGLF1
// (:3957-4095), mapLF1 (:3889-3950), countBt2Side (:2958-2999), rank_M (:4100),
// countMSide (:3146), select_F (:4113-4168), getInEdgeCount (:4172-4210),
// tryOffset/getOffset (:2719, :5682-5716).
#ifndef HT2_GRAPH_H_
#define HT2_GRAPH_H_

#include "ht2_fm.h"

template <typename IT>
struct Ht2GLoc {
    const uint8_t* side;
    uint32_t sideNum, charOff;
};

template <typename IT>
HT2_HD Ht2GLoc<IT> ht2g_locate(const Ht2Fm<IT>& fm, uint32_t row) {
    Ht2GLoc<IT> l;
    l.sideNum = row / fm.g->sideGbwtLen;
    l.charOff = row - l.sideNum * fm.g->sideGbwtLen;
    l.side = fm.gfm + (uint64_t)l.sideNum * fm.g->sideSz;
    return l;
}

template <typename IT>
HT2_HD int ht2g_rowL(const Ht2Fm<IT>& fm, uint32_t row) {
    const Ht2GLoc<IT> l = ht2g_locate(fm, row);
    return (l.side[l.charOff >> 2] >> ((l.charOff & 3) << 1)) & 3;
}

template <typename IT>
HT2_HD bool ht2g_is_zoff(const Ht2Fm<IT>& fm, uint32_t row) {
    for (uint32_t i = 0; i < fm.g->nzOffs; i++) if (row == fm.zoffs[i]) return true;
    return false;
}

// fchr[c] + occ(c, row): countBt2Side with the '$' rows (several on a graph) not counted as 'A'.
template <typename IT>
HT2_HD uint32_t ht2g_lf(const Ht2Fm<IT>& fm, uint32_t row, int c) {
    const Ht2Gfm* g = fm.g;
    const Ht2GLoc<IT> l = ht2g_locate(fm, row);
    uint32_t cnt = ht2_count_upto(l.side, l.charOff, c);
    if (c == 0) {
        const uint32_t sideStart = row - l.charOff;
        for (uint32_t i = 0; i < g->nzOffs; i++) {
            const uint32_t z = fm.zoffs[i];
            if (z >= sideStart && z < row) cnt--;
        }
    }
    const IT* acgt = (const IT*)(l.side + g->sideGbwtSz + 2 * sizeof(IT));
    return (uint32_t)(IT)(acgt[c] + cnt + g->fchr[c]);
}

// rank1(M, row) = # set M bits in rows [0,row).
template <typename IT>
HT2_HD uint32_t ht2g_rank_M(const Ht2Fm<IT>& fm, uint32_t row) {
    const Ht2Gfm* g = fm.g;
    const Ht2GLoc<IT> l = ht2g_locate(fm, row);
    const uint32_t cnt = ht2_count_bits(l.side + (g->sideGbwtSz - (g->sideGbwtSz >> 2)), l.charOff);
    const IT* tr = (const IT*)(l.side + g->sideGbwtSz);
    return (uint32_t)(IT)(tr[1] + cnt);
}

HT2_HD int ht2g_bit(const uint8_t* bits, uint32_t i) { return (bits[i >> 3] >> (i & 7)) & 1; }

// Row of the count-th (>= 1) set F bit at or after 'row' (select_F walks forward across sides).
template <typename IT>
HT2_HD uint32_t ht2g_select_F(const Ht2Fm<IT>& fm, uint32_t row, uint32_t count) {
    const Ht2Gfm* g = fm.g;
    Ht2GLoc<IT> l = ht2g_locate(fm, row);
    while (true) {
        const uint8_t* fbits = l.side + (g->sideGbwtSz >> 1);
        // whole bytes first, then bit by bit inside the byte that holds the answer
        while (l.charOff < g->sideGbwtLen) {
            if ((l.charOff & 7) == 0 && l.charOff + 8 <= g->sideGbwtLen) {
                const uint32_t pc = (uint32_t)HT2_POPC64((uint64_t)fbits[l.charOff >> 3]);
                if (pc < count) { count -= pc; l.charOff += 8; continue; }
            }
            if (ht2g_bit(fbits, l.charOff)) {
                if (--count == 0) return l.sideNum * g->sideGbwtLen + l.charOff;
            }
            l.charOff++;
        }
        l.sideNum++;
        l.charOff = 0;
        l.side += g->sideSz;
    }
}

// Edge row r (an LF result) -> (node id, first BW row of that node); the tail
// shared by mapGLF's top boundary and mapGLF1 (gfm.h:3786-3807, 3978-4000).
// Also returns where select started and the M_occ it was relative to, so that
// mapGLF1 can select the NEXT node's first row from the same start.
template <typename IT>
HT2_HD void ht2g_edge_to_node(const Ht2Fm<IT>& fm, uint32_t r, uint32_t& node, uint32_t& firstRow,
                              uint32_t& F_loc, uint32_t& M_occ) {
    const Ht2Gfm* g = fm.g;
    node = (uint32_t)(IT)(ht2g_rank_M(fm, r + 1) - 1);
    Ht2GLoc<IT> l = ht2g_locate(fm, r + 1);
    while (true) {
        const IT* tr = (const IT*)(l.side + g->sideGbwtSz);
        F_loc = tr[0]; M_occ = tr[1];
        if (M_occ <= node) break;
        l.side -= g->sideSz;       // the node's first edge lies before this side's M prefix
    }
    if (M_occ > 0) F_loc = (uint32_t)(IT)(F_loc + 1);
    firstRow = (node + 1 > M_occ) ? ht2g_select_F(fm, F_loc, node + 1 - M_occ) : F_loc;
}

// GFM::getInEdgeCount: for the rows [top,bot) (top is a node's first row) list
// (node index within the range, # extra incoming edges) for nodes with > 1 row.
// Returns the number of entries written (cap entries at most; excess sets overflow).
template <typename IT>
HT2_HD uint32_t ht2g_in_edge_count(const Ht2Fm<IT>& fm, uint32_t top, uint32_t bot, uint16_t (*out)[2], uint32_t cap, bool& overflow) {
    const Ht2Gfm* g = fm.g;
    Ht2GLoc<IT> l = ht2g_locate(fm, top);
    uint32_t n = 0, curr = 0, num0s = 0;
    bool first = true, curOk = false;
    while (top < bot) {
        const uint8_t* fbits = l.side + (g->sideGbwtSz >> 1);
        if (first) first = false;
        else if (ht2g_bit(fbits, l.charOff)) { curr++; num0s = 0; }
        else {
            num0s++;
            if (num0s == 1) {
                curOk = n < cap;
                if (curOk) { out[n][0] = (uint16_t)curr; out[n][1] = 0; n++; } else overflow = true;
            }
            if (curOk) out[n - 1][1] = (uint16_t)num0s;
        }
        if (l.charOff + 1 == g->sideGbwtLen) { l.sideNum++; l.charOff = 0; l.side += g->sideSz; }
        else l.charOff++;
        top++;
    }
    return n;
}

#define HT2G_MAX_IEDGES 24

// GFM::mapGLF.  k = kseeds: the in-edge list is only built for node ranges <= k.
template <typename IT>
HT2_HD void ht2g_mapGLF(const Ht2Fm<IT>& fm, uint32_t top, uint32_t bot, int c, uint32_t k,
                        uint32_t& ntop, uint32_t& nbot, uint32_t& node_top, uint32_t& node_bot,
                        uint16_t (*iedges)[2], uint32_t& niedges, bool& overflow) {
    const Ht2Gfm* g = fm.g;
    niedges = 0;
    uint32_t t = ht2g_lf(fm, top, c);
    uint32_t b = ht2g_lf(fm, bot, c);
    if (t + 1 >= g->gbwtLen || t >= b) { ntop = nbot = node_top = node_bot = 0; return; }
    uint32_t F_loc, M_occ;
    ht2g_edge_to_node(fm, t, node_top, ntop, F_loc, M_occ);
    {
        node_bot = ht2g_rank_M(fm, b);
        const Ht2GLoc<IT> l = ht2g_locate(fm, b);
        const IT* tr = (const IT*)(l.side + g->sideGbwtSz);
        uint32_t bF = tr[0], bM = tr[1];
        if (bM > 0) bF = (uint32_t)(IT)(bF + 1);
        nbot = (node_bot + 1 > bM) ? ht2g_select_F(fm, bF, node_bot + 1 - bM) : bF;
    }
    if (node_bot - node_top <= k && node_bot - node_top < nbot - ntop)
        niedges = ht2g_in_edge_count(fm, ntop, nbot, iedges, HT2G_MAX_IEDGES, overflow);
}

// GFM::mapGLF1(row, l, c): one-row range extended with base c.
template <typename IT>
HT2_HD void ht2g_mapGLF1c(const Ht2Fm<IT>& fm, uint32_t row, int c,
                          uint32_t& ntop, uint32_t& nbot, uint32_t& node_top, uint32_t& node_bot) {
    if (ht2g_rowL(fm, row) != c || ht2g_is_zoff(fm, row)) { ntop = nbot = node_top = node_bot = 0; return; }
    const uint32_t t = ht2g_lf(fm, row, c);
    uint32_t F_loc, M_occ;
    ht2g_edge_to_node(fm, t, node_top, ntop, F_loc, M_occ);
    node_bot = node_top + 1;
    nbot = (node_bot + 1 > M_occ) ? ht2g_select_F(fm, F_loc, node_bot + 1 - M_occ) : F_loc;
}

// GFM::mapGLF1(row, l): follow the row's own character.  Returns false on a '$' row.
template <typename IT>
HT2_HD bool ht2g_mapGLF1(const Ht2Fm<IT>& fm, uint32_t row, uint32_t& ntop, uint32_t& node_top) {
    if (ht2g_is_zoff(fm, row)) return false;
    const int c = ht2g_rowL(fm, row);
    const uint32_t t = ht2g_lf(fm, row, c);
    uint32_t F_loc, M_occ;
    ht2g_edge_to_node(fm, t, node_top, ntop, F_loc, M_occ);
    return true;
}

// GFM::getOffset(row, node): joined-text offset of a node reached through BW row 'row'.
template <typename IT>
HT2_HD uint32_t ht2g_get_offset(const Ht2Fm<IT>& fm, uint32_t row, uint32_t node, uint32_t& nsteps) {
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
