// ht2_seed.h -- the seed search on its own: the chain of partial searches of one
// read strand (HI_Aligner::partialSearch, hi_aligner.h:6361-6601, driven like
// nextBWT does, :4720-4751) and the resolution of small hit ranges to joined /
// text coordinates (GFM::getOffset gfm.h:5682, joinedToTextOff :5527), for
// linear AND graph (SNP) global indexes.  Host+device; used by the
// ht2gpu_seed_search kernel (ht2_gpu.cu) and by tests/hostsim.
#ifndef HT2_SEED_H_
#define HT2_SEED_H_

#include "ht2_graph.h"
#include "ht2_params.h"

struct Ht2SeedHit {            // one BWTHit (hi_aligner.h:108-208)
    uint32_t top, bot, node_top, node_bot;   // HT2_IDX_MAX32 when blank
    uint32_t bwoff, len;
    uint8_t  hit_type, pseudogeneStop, anchorStop, niedges;
    uint16_t iedges[HT2G_MAX_IEDGES][2];     // (node index in range, # extra incoming edges)
};

struct Ht2SeedState {          // ReadBWTHit cursor (hi_aligner.h:216-391)
    uint32_t len, cur, done, numPartialSearch, numUniqueSearch;
    uint32_t err;
    uint32_t nLF, algBytes;
};

// One step of the search on base c from [top,bot): linear -> ht2_lf2, graph -> mapGLF / mapGLF1.
template <bool GRAPH, typename IT>
HT2_HD void ht2_seed_step(const Ht2Fm<IT>& fm, uint32_t top, uint32_t bot, int c, uint32_t kseeds,
                          uint32_t& ntop, uint32_t& nbot, uint32_t& nntop, uint32_t& nnbot,
                          uint16_t (*ie)[2], uint32_t& nie, Ht2SeedState& st) {
    nie = 0;
    if (!GRAPH) {
        st.nLF += (bot - top == 1) ? 1u : 2u;
        st.algBytes += ((bot - top == 1) || (top >> HT2_SIDE_SHIFT) == (bot >> HT2_SIDE_SHIFT)) ? HT2_SIDE_BYTES : 2u * HT2_SIDE_BYTES;
        ht2_lf2(fm, top, bot, c, ntop, nbot);
        nntop = ntop; nnbot = nbot;
        return;
    }
    bool overflow = false;
    if (bot - top != 1) {
        st.nLF += 2;
        st.algBytes += 6u * fm.g->sideSz;          // per boundary: char side + M side + F side
        ht2g_mapGLF(fm, top, bot, c, kseeds, ntop, nbot, nntop, nnbot, ie, nie, overflow);
    } else {
        st.nLF += 1;
        st.algBytes += 3u * fm.g->sideSz;
        ht2g_mapGLF1c(fm, top, c, ntop, nbot, nntop, nnbot);
        if (ntop + 1 < nbot) { ie[0][0] = 0; ie[0][1] = (uint16_t)(nbot - ntop - 1); nie = 1; }
    }
    if (overflow) st.err |= 1;
}

// A partial search parked after a slice of LF steps (time slicing inside the alignment kernel: the slot stays in
// TS_PS and is regrouped, so that a round is never as long as its slowest lane's search).
struct Ht2SeedResume {
    uint32_t active, top, bot, ntop, nbot, dep, same, similar, nie;
    uint8_t  pseudo, anchor, pad[2];
    uint16_t ie[HT2G_MAX_IEDGES][2];
};

// HI_Aligner::partialSearch.  pseudogeneStop / anchorStop are in/out like the reference's.  With rs != NULL the
// search runs at most `budget` steps per call: it returns false after parking its state in *rs (call again with
// the same arguments to continue) and true when it has finished.
template <bool GRAPH>
HT2_HD bool ht2_seed_partial(const Ht2Fm<uint32_t>& fm, const Ht2ParamsCore& P, const uint8_t* seq, Ht2SeedState& st,
                             Ht2SeedHit& ph, bool& pseudogeneStop, bool& anchorStop, Ht2SeedResume* rs = NULL, uint32_t budget = 0xffffffffu) {
    bool pseudogeneStop_ = pseudogeneStop, anchorStop_ = anchorStop;
    pseudogeneStop = anchorStop = false;
    const uint32_t ftabLen = fm.g->ftabChars, len = st.len, minK = P.minK;
    const bool resume = rs != NULL && rs->active != 0;
    if (!resume) st.numPartialSearch++;
    const uint32_t offset = st.cur;
    uint32_t dep = offset;
    ph.top = ph.bot = ph.node_top = ph.node_bot = HT2_IDX_MAX32;
    ph.bwoff = offset; ph.hit_type = 1 /*CANDIDATE_HIT*/; ph.niedges = 0; ph.pseudogeneStop = ph.anchorStop = 0;
    uint32_t top = 0, bot = 0, ntop = 0, nbot = 0;
    uint32_t same_range = 0, similar_range = 0;
    uint16_t ie[HT2G_MAX_IEDGES][2], tie[HT2G_MAX_IEDGES][2];
    uint32_t nie = 0, ntie = 0;
    if (resume) {
        top = rs->top; bot = rs->bot; ntop = rs->ntop; nbot = rs->nbot; dep = rs->dep; same_range = rs->same; similar_range = rs->similar;
        pseudogeneStop_ = rs->pseudo != 0; anchorStop_ = rs->anchor != 0; nie = rs->nie;
        for (uint32_t e = 0; e < nie; e++) { ie[e][0] = rs->ie[e][0]; ie[e][1] = rs->ie[e][1]; }
        rs->active = 0;
    } else {
    const uint32_t left = len - dep;
    if (left < ftabLen + 1) { st.cur = len; ph.len = st.cur - offset; st.done = 1; return true; }
    for (uint32_t i = 0; i < ftabLen; i++) {
        if (seq[len - dep - 1 - i] > 3) {
            st.cur += (i + 1);
            ph.len = st.cur - offset;
            if (st.cur >= len) st.done = 1;
            return true;
        }
    }
    ht2_ftab_lohi(fm, seq, len - dep - ftabLen, top, bot);
    st.algBytes += 8;
    dep += ftabLen;
    if (top >= bot) { st.cur = dep; ph.len = st.cur - offset; if (st.cur >= len) st.done = 1; return true; }
    }
    const uint32_t khits5 = P.khits < 5 ? P.khits : 5;
    while (dep < len) {
        if (rs != NULL && budget-- == 0) {
            rs->active = 1; rs->top = top; rs->bot = bot; rs->ntop = ntop; rs->nbot = nbot; rs->dep = dep; rs->same = same_range; rs->similar = similar_range;
            rs->pseudo = pseudogeneStop_ ? 1 : 0; rs->anchor = anchorStop_ ? 1 : 0; rs->nie = nie;
            for (uint32_t e = 0; e < nie; e++) { rs->ie[e][0] = ie[e][0]; rs->ie[e][1] = ie[e][1]; }
            return false;
        }
        const int c = seq[len - dep - 1];
        uint32_t ttop = 0, tbot = 0, tntop = 0, tnbot = 0;
        ntie = 0;
        if (c <= 3) ht2_seed_step<GRAPH, uint32_t>(fm, top, bot, c, P.kseeds, ttop, tbot, tntop, tnbot, tie, ntie, st);
        if (ttop >= tbot) break;
        const uint32_t nw = tnbot - tntop, ow = nbot - ntop;
        if (pseudogeneStop_) {
            if (nw < ow && ow <= khits5) {
                if (dep - offset >= minK + 6 && similar_range >= 5) { st.numUniqueSearch++; pseudogeneStop = true; break; }
            }
            if (nw != 1) {
                if (nw + 2 >= ow) similar_range++;
                else if (nw + 4 < ow) similar_range = 0;
            } else pseudogeneStop_ = false;
        }
        if (anchorStop_) {
            if (nw != 1 && ow == nw) { same_range++; if (same_range >= 5) anchorStop_ = false; }
            else same_range = 0;
            if (dep - offset >= minK + 8 && nw >= 4) anchorStop_ = false;
        }
        top = ttop; bot = tbot; ntop = tntop; nbot = tnbot;
        nie = ntie;
        for (uint32_t e = 0; e < ntie; e++) { ie[e][0] = tie[e][0]; ie[e][1] = tie[e][1]; }
        dep++;
        if (anchorStop_) {
            if (dep - offset >= minK + 12 && bot - top == 1) { st.numUniqueSearch++; anchorStop = true; break; }
        }
    }
    if (top < bot) {
        uint8_t hit_type = 1;
        if (anchorStop) hit_type = 3; else if (pseudogeneStop) hit_type = 2;
        bool report = ntop < nbot;
        if (nbot - ntop < bot - top && nie == 0) report = false;
        if (report) {
            ph.top = top; ph.bot = bot; ph.node_top = ntop; ph.node_bot = nbot;
            ph.niedges = (uint8_t)nie;
            for (uint32_t e = 0; e < nie; e++) { ph.iedges[e][0] = ie[e][0]; ph.iedges[e][1] = ie[e][1]; }
        }
        ph.len = dep - offset;
        ph.hit_type = hit_type;
        st.cur = dep;
        if (st.cur >= len) { if (hit_type == 1) st.numUniqueSearch++; st.done = 1; }
    } else {
        // the last accepted range is never empty once the ftab range was not (loop breaks before committing)
        ph.len = dep - offset;
    }
    ph.pseudogeneStop = pseudogeneStop ? 1 : 0;
    ph.anchorStop = anchorStop ? 1 : 0;
    return true;
}

// Joined offset of element i of a hit's node range (group_walk.h:545-560: first BW row of the node).
template <bool GRAPH>
HT2_HD uint32_t ht2_seed_elt_offset(const Ht2Fm<uint32_t>& fm, const Ht2SeedHit& ph, uint32_t i, uint32_t& row, Ht2SeedState& st) {
    uint32_t num_iedges = 0;
    for (uint32_t e = 0; e < ph.niedges; e++) { if (i <= ph.iedges[e][0]) break; num_iedges += ph.iedges[e][1]; }
    row = ph.top + i + num_iedges;
    if (!GRAPH) {
        uint32_t r = row, steps = 0;
        while (true) {
            if (r == fm.z0) { st.nLF += steps; st.algBytes += steps * HT2_SIDE_BYTES; return steps; }
            if ((r & fm.offMask) == r) { st.nLF += steps; st.algBytes += steps * HT2_SIDE_BYTES + 4; return fm.offs[r >> fm.offRate] + steps; }
            int c;
            r = ht2_lf_own(fm, r, c);
            steps++;
        }
    }
    uint32_t steps = 0;
    const uint32_t off = ht2g_get_offset(fm, row, ph.node_top + i, steps);
    st.nLF += steps; st.algBytes += steps * 3u * fm.g->sideSz + 4;
    return off;
}

#endif // HT2_SEED_H_
