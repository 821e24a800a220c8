// ht2_gwalk.h -- the group walk over a graph FM index (host+device).
//
// Restates GWState::init / GWState::advance / GroupWalk2S::init / advanceElement
// (group_walk.h:509-870, 1023-1336, 1404-1495) for graph (GBWT) indexes: the rows of a hit range are walked
// left TOGETHER, the range splitting by BW character (mapLFRange, gfm.h:3636) and at '$' rows, until every
// element (node) of the original range has met a sampled node.  On a linear index this gives exactly what
// walking every row on its own gives; on a graph it does not: rows that reach the same node merge, and the
// reference then resolves the merged element to its own index inside the range ("sa.offs[jmap] =
// gws.map[j]", group_walk.h:1189, 1238) -- a behaviour the alignments depend on and that is reproduced here.
#ifndef HT2_GWALK_H_
#define HT2_GWALK_H_

#include "ht2_graph.h"

#ifndef HT2_GW_MAXELT
#define HT2_GW_MAXELT 64      /* max(khits, kseeds) elements: --very-sensitive runs -k 30, i.e. 60 seeds */
#endif
#ifndef HT2_GW_MAXROWS
#define HT2_GW_MAXROWS 128
#endif
#ifndef HT2_GW_MAXST
#define HT2_GW_MAXST 96
#endif
#define HT2_GW_MASK 0xffffffffu

struct Ht2GwState {
    uint32_t top, bot, node_top, node_bot, step, mapi, nmap, nie;
    uint16_t map[HT2_GW_MAXELT];
    uint16_t ie[HT2G_MAX_IEDGES][2];
};

struct Ht2GWalk {
    Ht2GwState st[HT2_GW_MAXST];
    uint32_t nst;
    uint32_t offs[HT2_GW_MAXELT];       // resolved joined offsets (HT2_GW_MASK = not yet)
    uint16_t fmapRange[HT2_GW_MAXELT];  // which state tracks the element (GWHit::fmap[].first)
    uint32_t nelt;
    uint32_t err;
    uint32_t nLF;
};

template <typename IT>
struct Ht2GroupWalk {
    const Ht2Fm<IT>& fm;
    Ht2GWalk& G;
    HT2_HD Ht2GroupWalk(const Ht2Fm<IT>& f, Ht2GWalk& g) : fm(f), G(g) {}

    // GFM::tryOffset (gfm.h:2719-2734): unsampled nodes and unsampled entries both read as "not yet"
    HT2_HD uint32_t tryOffset(uint32_t row, uint32_t node) const {
        if (ht2g_is_zoff(fm, row)) return 0;
        if ((node & fm.offMask) == node) {
            const IT off = fm.offs[node >> fm.offRate];
            return off == Ht2Fm<IT>::imax() ? HT2_GW_MASK : (uint32_t)off;
        }
        return HT2_GW_MASK;
    }

    // one mapGLF call on rows [top,bot) with base c; optional in-edge list
    HT2_HD void mapGLF(uint32_t top, uint32_t bot, int c, uint32_t k, uint32_t& nt, uint32_t& nb, uint32_t& nnt, uint32_t& nnb,
                       uint16_t (*ie)[2], uint32_t& nie, bool wantIe) {
        bool overflow = false;
        G.nLF += 2;
        uint16_t dummy[1][2];
        uint32_t dn = 0;
        if (wantIe) ht2g_mapGLF(fm, top, bot, c, k, nt, nb, nnt, nnb, ie, nie, overflow);
        else ht2g_mapGLF(fm, top, bot, c, 0u, nt, nb, nnt, nnb, dummy, dn, overflow);
        if (overflow) G.err |= 1;
    }

    HT2_HD uint32_t newState() {
        if (G.nst >= HT2_GW_MAXST) { G.err |= 2; return HT2_GW_MAXST - 1; }
        Ht2GwState& S = G.st[G.nst];
        S.top = S.bot = S.node_top = S.node_bot = S.step = S.mapi = S.nmap = S.nie = 0;
        return G.nst++;
    }

    // GWState::init (group_walk.h:509-870) for state r; top/bot/node range/iedges/step/map are already set.
    HT2_HDN void stateInit(uint32_t r) {
        uint32_t trimBegin = 0, trimEnd = 0;
        bool empty = true;
        {
            Ht2GwState& S = G.st[r];
            uint32_t num_iedges = 0, e = 0;
            for (uint32_t i = S.mapi; i < S.nmap; i++) {
                if (G.offs[S.map[i]] == HT2_GW_MASK) {
                    while (e < S.nie) { if (i <= S.ie[e][0]) break; num_iedges += S.ie[e][1]; e++; }
                    const uint32_t bwrow = S.top + i + num_iedges, node = S.node_top + i;
                    uint32_t toff = tryOffset(bwrow, node);
                    if (toff != HT2_GW_MASK) {
                        toff = (uint32_t)(IT)(toff + S.step);
                        G.offs[S.map[i + S.mapi]] = toff;
                    }
                }
                if (G.offs[S.map[i]] != HT2_GW_MASK) { if (empty) trimBegin++; else trimEnd++; }
                else { trimEnd = 0; empty = false; G.fmapRange[S.map[i]] = (uint16_t)r; }
            }
            S.mapi += trimBegin;
            if (trimBegin > 0) {
                S.top += trimBegin;
                uint32_t e2 = 0;
                for (; e2 < S.nie; e2++) { if (S.ie[e2][0] >= trimBegin) break; S.top += S.ie[e2][1]; }
                if (e2 > 0) { for (uint32_t q = 0; q + e2 < S.nie; q++) { S.ie[q][0] = S.ie[q + e2][0]; S.ie[q][1] = S.ie[q + e2][1]; } S.nie -= e2; }
                for (uint32_t q = 0; q < S.nie; q++) S.ie[q][0] = (uint16_t)(S.ie[q][0] - trimBegin);
            }
            S.node_top += trimBegin;
            if (trimEnd > 0) {
                S.nmap -= trimEnd;
                S.bot -= trimEnd;
                const uint32_t node_range = S.node_bot - S.node_top;
                while (S.nie > 0) {
                    if (S.ie[S.nie - 1][0] < node_range - trimEnd) break;
                    S.bot -= S.ie[S.nie - 1][1];
                    S.nie--;
                }
            }
            S.node_bot -= trimEnd;
        }
        if (empty) return;
        // '$' rows strictly inside the range split it (group_walk.h:742-868)
        uint32_t tz[4]; uint32_t ntz = 0;
        {
            const Ht2GwState& S = G.st[r];
            for (uint32_t i = 0; i < fm.g->nzOffs; i++) {
                const uint32_t z = (i == 0) ? fm.z0 : (uint32_t)fm.zoffs[i];
                if (z > S.top && z < S.bot) { if (ntz < 4) tz[ntz++] = z; else G.err |= 4; }
            }
        }
        if (ntz > 0) {
            uint16_t g2n[HT2_GW_MAXROWS]; uint32_t ng = 0;
            const uint32_t top = G.st[r].top, bot = G.st[r].bot, node_top = G.st[r].node_top, node_bot = G.st[r].node_bot, mapi = G.st[r].mapi;
            {
                const Ht2GwState& S = G.st[r];
                uint32_t n = 0, e = 0;
                for (uint32_t row = 0; row < bot - top; row++) {
                    if (ng < HT2_GW_MAXROWS) g2n[ng++] = (uint16_t)n; else G.err |= 8;
                    if (e < S.nie) {
                        if (n == S.ie[e][0]) {
                            for (uint32_t a = 0; a < S.ie[e][1]; a++) { if (ng < HT2_GW_MAXROWS) g2n[ng++] = (uint16_t)n; else G.err |= 8; row++; }
                            e++;
                        }
                    }
                    n++;
                }
            }
            for (uint32_t i = 0; i < ntz; i++) {
                const uint32_t new_top = tz[i] + 1;
                if (i + 1 < ntz && new_top == tz[i + 1]) continue;
                if (new_top - top == ng) break;
                const uint32_t new_node_top = g2n[new_top - top] + node_top;
                const uint32_t new_bot = (i + 1 < ntz) ? tz[i + 1] : bot;
                uint32_t new_node_bot = node_bot;
                if (new_bot - top < ng) {
                    new_node_bot = node_top + g2n[new_bot - top];
                    if (new_bot - top > 0 && g2n[new_bot - top] == g2n[new_bot - top - 1]) new_node_bot++;
                }
                if (new_top >= new_bot) continue;
                const uint32_t r2 = newState();
                Ht2GwState& S2 = G.st[r2];
                for (uint32_t j = new_top - top; j + 1 < new_bot - top;) {
                    const uint32_t n = g2n[j];
                    uint32_t j2 = j + 1;
                    while (j2 < new_bot - top) { if (n != g2n[j2]) break; j2++; }
                    if (j + 1 < j2) {
                        if (S2.nie < HT2G_MAX_IEDGES) { S2.ie[S2.nie][0] = (uint16_t)(n - (new_node_top - node_top)); S2.ie[S2.nie][1] = (uint16_t)(j2 - j - 1); S2.nie++; }
                        else G.err |= 1;
                    }
                    j = j2;
                }
                S2.nmap = new_node_bot - new_node_top; S2.mapi = 0;
                if (S2.nmap > HT2_GW_MAXELT) { G.err |= 16; S2.nmap = HT2_GW_MAXELT; }
                for (uint32_t j = new_node_top; j < new_node_top + S2.nmap; j++) S2.map[j - new_node_top] = G.st[r].map[j - node_top + mapi];
                S2.top = new_top; S2.bot = new_bot; S2.node_top = new_node_top; S2.node_bot = new_node_bot; S2.step = G.st[r].step;
                stateInit(r2);
            }
            Ht2GwState& S = G.st[r];
            S.bot = tz[0];
            S.node_bot = g2n[S.bot - top - 1] + node_top + 1;
            S.nmap = S.node_bot - node_top + mapi;
            uint32_t width = S.node_bot - node_top;
            for (uint32_t e = 0; e < S.nie; e++) {
                if (S.ie[e][0] >= S.node_bot - node_top) { S.nie = e; break; }
                width += S.ie[e][1];
            }
            if (width != S.bot - top && S.nie > 0) {
                S.ie[S.nie - 1][1] -= 1;
                if (S.ie[S.nie - 1][1] == 0) S.nie--;
            }
        }
    }

    // the "range narrowed" merge step shared by the two branches of advance (group_walk.h:1166-1200, 1211-1251)
    HT2_HD void mergeStep(uint16_t* map, uint32_t& nmap, const uint8_t* ch, uint32_t nrows, int c, uint32_t curtop) {
        uint32_t j1 = 0, j2 = 0;
        for (uint32_t q = 0; q < nrows; q++) if (ch[q] == c) { j1 = q; break; }
        for (uint32_t j = 0; j + 1 < nmap; j++) {
            for (uint32_t q = j1 + 1; q < nrows; q++) if (ch[q] == c) { j2 = q; break; }
            uint32_t nt, nb, nnt, nnb, dn = 0;
            mapGLF(curtop + j1, curtop + j2 + 1, c, 0, nt, nb, nnt, nnb, NULL, dn, false);
            if (nnb - nnt == 1) {
                const uint32_t jmap = map[j];
                G.offs[jmap] = map[j];            // sic: the element's own index becomes its offset
                map[j] = 0xffffu;
            }
            j1 = j2; j2 = 0;
        }
        uint32_t m = 0;
        for (uint32_t j = 0; j < nmap; j++) if (map[j] != 0xffffu) map[m++] = map[j];
        nmap = m;
    }

    // GWState::advance (group_walk.h:1023-1336)
    HT2_HDN void advance(uint32_t r) {
        if (G.st[r].bot - G.st[r].top > 1) {
            bool first = true;
            uint32_t newtop = 0, newbot = 0, new_node_top = 0, new_node_bot = 0;
            uint16_t gmap[HT2_GW_MAXELT]; uint32_t ngmap = 0;
            uint16_t backup_ie[HT2G_MAX_IEDGES][2]; uint32_t nbackup = 0;
            const uint32_t top = G.st[r].top, bot = G.st[r].bot, node_top = G.st[r].node_top, node_bot = G.st[r].node_bot;
            const uint32_t mapi = G.st[r].mapi, step = G.st[r].step, nie = G.st[r].nie;
            uint16_t ie[HT2G_MAX_IEDGES][2];
            for (uint32_t e = 0; e < nie; e++) { ie[e][0] = G.st[r].ie[e][0]; ie[e][1] = G.st[r].ie[e][1]; }
            uint16_t omap[HT2_GW_MAXELT];
            for (uint32_t i = 0; i < G.st[r].nmap && i < HT2_GW_MAXELT; i++) omap[i] = G.st[r].map[i];
            uint32_t curtop = top, curbot = bot, cur_node_top = node_top, cur_node_bot = node_bot;
            for (uint32_t e = 0; e < nie + 1; e++) {
                if (e >= nie) {
                    if (e > 0) {
                        curtop = curbot + ie[e - 1][1];
                        curbot = bot;
                        if (curtop >= curbot) break;
                        cur_node_top = cur_node_bot;
                        cur_node_bot = node_bot;
                    }
                } else {
                    if (e > 0) {
                        curtop = curbot + ie[e - 1][1];
                        curbot = curtop + (ie[e][0] - ie[e - 1][0]);
                        cur_node_top = cur_node_bot;
                    } else curbot = curtop + ie[e][0] + 1;
                    cur_node_bot = node_top + ie[e][0] + 1;
                }
                // mapLFRange: which rows hold which character (gfm.h:3636; countBt2SideRange)
                const uint32_t nrows = curbot - curtop;
                uint8_t ch[HT2_GW_MAXROWS];
                uint32_t in[4] = {0, 0, 0, 0};
                if (nrows > HT2_GW_MAXROWS) { G.err |= 8; return; }
                for (uint32_t j = 0; j < nrows; j++) { ch[j] = (uint8_t)ht2g_rowL(fm, curtop + j); in[ch[j]]++; }
                G.nLF += 1;
                for (int c = 0; c < 4; c++) {
                    if (in[c] == 0) continue;
                    if (first) {
                        first = false;
                        mapGLF(curtop, curbot, c, cur_node_bot - cur_node_top, newtop, newbot, new_node_top, new_node_bot, backup_ie, nbackup, true);
                        for (uint32_t j = 0; j < nrows; j++) if (ch[j] == c) { if (ngmap < HT2_GW_MAXELT) gmap[ngmap++] = omap[j + mapi + (cur_node_top - node_top)]; else G.err |= 16; }
                        if (new_node_bot - new_node_top < ngmap) mergeStep(gmap, ngmap, ch, nrows, c, curtop);
                    } else {
                        const uint32_t r2 = newState();
                        uint32_t ntop, nbot, nnt, nnb, ntie = 0;
                        uint16_t tie[HT2G_MAX_IEDGES][2];
                        mapGLF(curtop, curbot, c, cur_node_bot - cur_node_top, ntop, nbot, nnt, nnb, tie, ntie, true);
                        Ht2GwState& S2 = G.st[r2];
                        S2.mapi = 0; S2.nmap = 0;
                        for (uint32_t j = 0; j < nrows; j++) if (ch[j] == c) { if (S2.nmap < HT2_GW_MAXELT) S2.map[S2.nmap++] = omap[j + mapi + (cur_node_top - node_top)]; else G.err |= 16; }
                        if (nnb - nnt < S2.nmap) { uint32_t nm = S2.nmap; mergeStep(S2.map, nm, ch, nrows, c, curtop); S2.nmap = nm; }
                        S2.top = ntop; S2.bot = nbot; S2.node_top = nnt; S2.node_bot = nnb; S2.step = step + 1;
                        S2.nie = ntie;
                        for (uint32_t q = 0; q < ntie; q++) { S2.ie[q][0] = tie[q][0]; S2.ie[q][1] = tie[q][1]; }
                        stateInit(r2);
                    }
                }
            }
            Ht2GwState& S = G.st[r];
            S.mapi = 0;
            S.top = newtop; S.bot = newbot; S.node_top = new_node_top; S.node_bot = new_node_bot;
            S.nie = nbackup;
            for (uint32_t q = 0; q < nbackup; q++) { S.ie[q][0] = backup_ie[q][0]; S.ie[q][1] = backup_ie[q][1]; }
            if (ngmap > 0) { S.nmap = ngmap; for (uint32_t q = 0; q < ngmap; q++) S.map[q] = gmap[q]; }
        } else {
            Ht2GwState& S = G.st[r];
            uint32_t nt = 0, nn = 0;
            ht2g_mapGLF1(fm, S.top, nt, nn);
            G.nLF += 1;
            S.top = nt; S.bot = nt + 1; S.node_top = nn; S.node_bot = nn + 1;
            if (S.mapi > 0) { S.map[0] = S.map[S.mapi]; S.mapi = 0; }
            S.nmap = 1;
            S.nie = 0;
        }
        G.st[r].step++;
        stateInit(r);
    }

    // GroupWalk2S::init (group_walk.h:1404-1445)
    HT2_HDN void init(uint32_t top, uint32_t bot, uint32_t node_top, uint32_t nelt, const uint16_t (*ie)[2], uint32_t nie) {
        G.nst = 0; G.err = 0; G.nLF = 0;
        if (nelt > HT2_GW_MAXELT) { G.err |= 16; nelt = HT2_GW_MAXELT; }
        G.nelt = nelt;
        if (nelt == 1) {
            // one node: the walk follows its first row only (extra incoming edges are skipped by advance(),
            // group_walk.h:1065-1090), i.e. it is GFM::getOffset(row, node) -- no state machinery needed
            uint32_t steps = 0;
            G.offs[0] = ht2g_get_offset(fm, top, node_top, steps);
            G.nLF += steps;
            return;
        }
        for (uint32_t i = 0; i < nelt; i++) { G.offs[i] = HT2_GW_MASK; G.fmapRange[i] = 0; }
        const uint32_t r = newState();
        Ht2GwState& S = G.st[r];
        S.top = top; S.bot = bot; S.node_top = node_top; S.node_bot = node_top + nelt; S.step = 0;
        S.mapi = 0; S.nmap = nelt;
        for (uint32_t i = 0; i < nelt; i++) S.map[i] = (uint16_t)i;
        S.nie = 0;
        for (uint32_t e = 0; e < nie && e < HT2G_MAX_IEDGES; e++) { S.ie[e][0] = ie[e][0]; S.ie[e][1] = ie[e][1]; S.nie++; }
        stateInit(r);
    }

    // GroupWalk2S::advanceElement (group_walk.h:1461-1495)
    HT2_HDN uint32_t resolve(uint32_t elt) {
        uint32_t guard = 0;
        while (G.offs[elt] == HT2_GW_MASK) {
            advance(G.fmapRange[elt]);
            if (G.err || ++guard > 100000u) { G.err |= 32; return 0; }
        }
        return G.offs[elt];
    }
};

#endif // HT2_GWALK_H_
