// ht2_core_impl.h -- the recursive part of the per-read state machine.
// Included from ht2_core.h.
#ifndef HT2_CORE_IMPL_H_
#define HT2_CORE_IMPL_H_

// SplicedAligner::hybridSearch_recur (spliced_aligner.h:331-2052).
// Branches guarded by !ssdb.empty() are omitted (the splice-site DB is empty
// when no splice sites are known and --no-spliced-alignment is on).
template <bool GRAPH> HT2_NI int64_t Ht2AlignerT<GRAPH>::hybridSearchRecur(uint32_t rdi, const Ht2Hit& hit, uint32_t hitoff, uint32_t hitlen,
                                              bool alignMate, uint32_t dep)
{
    int64_t maxsc = HT2_MIN_I64;
    const uint32_t rdlen = W->rd[rdi].len;
    const uint32_t minK = P->minK, minKL = P->minKLocal;
    const int64_t mmpMax = P->mmpMax;
    int64_t cushion = 0;
    if (P->noSplicedAlignment) {
        // alignMate ? rdlen * 0.03 * sc.mm(255) : 0   (double -> TAlScore)
        cushion = alignMate ? (int64_t)((double)rdlen * 0.03 * (double)ht2_mmpen(*P, 255)) : 0;
    }
    if (hit.score + cushion < minsc[rdi]) return maxsc;
    if (dep >= HT2_MAX_DEPTH) return maxsc;
    if (dep > W->maxDepth) W->maxDepth = dep;
#ifdef HT2_TRACE
    fprintf(stderr, "  recur rdi %u dep %u hitoff %u hitlen %u hit[rdoff %u len %u toff %u score %ld trim %u,%u] mate %d atts %u/%u\n", rdi, dep, hitoff, hitlen, hit.rdoff, hit.len, hit.toff, (long)hit.score, hit.trim5, hit.trim3, (int)alignMate, W->localindexatts, W->maxLocalindexatts);
#endif
    if (dep >= HT2_DEPTH_CAP) { W->err |= HT2_ERR_DEPTH; return maxsc; }
    if (hitoff == hit.rdoff - hit.trim5 && hitlen == hit.len + hit.trim5 + hit.trim3) {
        if (isSearched(hit, rdi)) return maxsc;
        addSearched(hit, rdi);
    }
    if (W->err) return maxsc;
    const uint32_t poolMark = W->poolTop;
    Ht2Coord coords[8];
    uint32_t ncoords = 0;
    uint16_t localHits[16];   // _local_genomeHits[dep] as pool indexes
    uint32_t nLocalHits = 0;

    if (hitoff == 0 && hitlen == rdlen) {
        // full-length: report (spliced_aligner.h:406-682 with ssdb.empty())
        if (!redundant(rdi, hit)) {
            reportHit(rdi, hit);
            if (hit.score > maxsc) maxsc = hit.score;
            W->poolTop = poolMark;
            return maxsc;
        }
    } else if (hitoff > 0 && (hitoff + hitlen == rdlen || hitoff + hitoff < rdlen - hitlen)) {
        // ---------------- extend to the left (spliced_aligner.h:683-1361) ----
        bool use_localindex = true;
        if (hitoff == hit.rdoff && hitoff <= minK) {
            uint32_t leftext = HT2_IDX_MAX32, rightext = 0;
            Ht2Hit* t = poolAlloc();
            copyHit(*t, hit);
            extend(*t, rdi, leftext, rightext, 1);
            if (t->rdoff == 0) use_localindex = false;
        }
        int lid = localIndexId(hit.tidx, hit.toff);
        bool success = false, first = true;
        uint32_t count = 0;
        const uint32_t max_count = 2;
        int64_t prev_score = hit.score;
        while (!success && count++ < max_count && use_localindex) {
            if (W->localindexatts >= W->maxLocalindexatts) break;
            if (first) first = false;
            else {
                lid = lid >= 0 ? prevLocal(lid) : -1;
                if (lid < 0 || localGeom(lid)->len == 0) break;
            }
            if (lid < 0) break;
            Ht2Fm<uint16_t> lfm; lfm.init(blob, localGeom(lid));
            uint32_t extlen = 0;
            uint32_t top = 0xffff, bot = 0xffff, node_top = 0xffff, node_bot = 0xffff;
            uint32_t extoff = hitoff - 1;
            if (extoff > 0) extoff -= 1;
            if (extoff < P->minAnchorLen) extoff = P->minAnchorLen;
            uint32_t nelt = HT2_IDX_MAX32;
            uint32_t max_nelt = 5; // std::max<index_t>(5, extlen) with extlen == 0
            bool no_extension = false;
            bool uniqueStop = false;
            uint32_t minUniqueLen = minKL;
            for (; extoff < rdlen; extoff++) {
                extlen = 0;
                uniqueStop = true;
                W->localindexatts++;
                if (lfm.g->len == 0) { nelt = 0; top = bot = node_top = node_bot = 0; }
                else nelt = gfmSearch(lfm, rdi, hit.fw != 0, extoff, extlen, top, bot, node_top, node_bot,
                                      uniqueStop, minUniqueLen, 0xffffu, P->kseeds, true);
                if (extoff + 1 - extlen >= hitoff) { no_extension = true; break; }
                if (nelt <= max_nelt) break;
            }
            if (nelt > 0 && nelt <= max_nelt && extlen >= P->minAnchorLen && !no_extension) {
                ncoords = 0;
                getGenomeCoordsLocal(lfm, top, bot, node_top, node_bot, hit.fw != 0, extoff + 1 - extlen, extlen, coords, ncoords, 8);
                sortCoords(coords, ncoords);
                for (int ri = (int)ncoords - 1; ri >= 0; ri--) {
                    const Ht2Coord& coord = coords[ri];
                    Ht2Hit* tp = poolAlloc();
                    Ht2Hit& tempHit = *tp;
                    initHit(tempHit, coord.fw != 0, extoff + 1 - extlen, extlen, 0, 0, coord.ref, coord.off, coord.joinedOff);
                    // adjustWithALT: identity on linear indexes (hi_aligner.h:2402)
                    if (!compatibleWith(tempHit, hit, rdi)) {
                        W->poolTop--;
                        if (count == 1) continue; else break;
                    }
                    if (uniqueStop) {
                        uint32_t leftext = HT2_IDX_MAX32, rightext = 0;
                        extend(tempHit, rdi, leftext, rightext, 0);
                    }
                    int64_t msc = minsc[rdi];
                    bool combined = combineWith(tempHit, hit, rdi, msc);
                    msc = sinkFloor(rdi, cushion);
                    bool keep = false;
                    if (combined && tempHit.score >= msc) {
                        if (tempHit.score >= prev_score - mmpMax) {
                            int64_t tmp_maxsc = hybridSearchRecur(rdi, tempHit, tempHit.rdoff, tempHit.len + tempHit.trim3, alignMate, dep + 1);
                            if (tmp_maxsc > maxsc) maxsc = tmp_maxsc;
                        } else {
                            if (nLocalHits < 16) { localHits[nLocalHits++] = (uint16_t)(tp - W->pool); keep = true; }
                            else W->err |= HT2_ERR_POOL;
                        }
                    }
                    if (!keep) W->poolTop--; // release tempHit (deeper frames already unwound)
                }
            }
            if (maxsc >= prev_score - mmpMax) success = true;
            if (!success && (W->localindexatts >= W->maxLocalindexatts || count == max_count || prevLocal(lid) < 0)) {
                for (uint32_t ti = 0; ti < nLocalHits; ti++) {
                    Ht2Hit& tempHit = W->pool[localHits[ti]];
                    int64_t msc = sinkFloor(rdi, cushion);
                    if (tempHit.score >= msc) {
                        int64_t tmp_maxsc = hybridSearchRecur(rdi, tempHit, tempHit.rdoff, tempHit.len + tempHit.trim3, alignMate, dep + 1);
                        if (tmp_maxsc > maxsc) maxsc = tmp_maxsc;
                    }
                }
            }
        }
        if (!success) {
            if (hitoff > minK && W->localindexatts < W->maxLocalindexatts) {
                uint32_t extlen = 0;
                uint32_t top = HT2_IDX_MAX32, bot = HT2_IDX_MAX32, node_top = HT2_IDX_MAX32, node_bot = HT2_IDX_MAX32;
                uint32_t extoff = hitoff - 1;
                bool uniqueStop = true;
                uint32_t nelt = gfmSearch(gfm, rdi, hit.fw != 0, extoff, extlen, top, bot, node_top, node_bot,
                                          uniqueStop, minK, HT2_IDX_MAX32, P->kseeds, false);
                if (nelt > 0 && nelt <= 5 && extlen >= minK) {
                    W->nCoords = 0;
                    bool straddled = false;
                    getGenomeCoords(top, bot, node_top, node_bot, hit.fw != 0, bot - top, extlen, true, straddled);
                    ncoords = W->nCoords < 8 ? W->nCoords : 8;
                    for (uint32_t i = 0; i < ncoords; i++) coords[i] = W->coords[i];
                    if (ncoords > 1) sortCoords(coords, ncoords);
                    for (int ri = (int)ncoords - 1; ri >= 0; ri--) {
                        const Ht2Coord& coord = coords[ri];
                        Ht2Hit* tp = poolAlloc();
                        Ht2Hit& tempHit = *tp;
                        initHit(tempHit, coord.fw != 0, extoff + 1 - extlen, extlen, 0, 0, coord.ref, coord.off, coord.joinedOff);
                        if (!compatibleWith(tempHit, hit, rdi)) { W->poolTop--; continue; }
                        if (uniqueStop) {
                            uint32_t leftext = HT2_IDX_MAX32, rightext = 0;
                            extend(tempHit, rdi, leftext, rightext, 0);
                        }
                        int64_t msc = minsc[rdi];
                        bool combined = combineWith(tempHit, hit, rdi, msc);
                        msc = sinkFloor(rdi, cushion);
                        if (combined && tempHit.score >= msc) {
                            int64_t tmp_maxsc = hybridSearchRecur(rdi, tempHit, tempHit.rdoff, tempHit.len + tempHit.trim3, alignMate, dep + 1);
                            if (tmp_maxsc > maxsc) maxsc = tmp_maxsc;
                        }
                        W->poolTop--;
                    }
                }
            }
            Ht2Hit* tp = poolAlloc();
            Ht2Hit& tempHit = *tp;
            copyHit(tempHit, hit);
            {
                int64_t floor_ = maxsc > minsc[rdi] ? maxsc : minsc[rdi];
                uint32_t trimMax = (uint32_t)((tempHit.score - floor_) / ht2_scpen(*P, 0));
                if (tempHit.rdoff < trimMax) {
                    Ht2Hit* trp = poolAlloc();
                    Ht2Hit& trimedHit = *trp;
                    copyHit(trimedHit, tempHit);
                    trimedHit.trim5 = tempHit.rdoff; // GenomeHit::trim5 (hi_aligner.h:831-854)
                    calculateScore(trimedHit, rdi);
                    int64_t tmp_score = trimedHit.score;
                    if (tmp_score > maxsc && tmp_score >= minsc[rdi]) {
                        int64_t tmp_maxsc = hybridSearchRecur(rdi, trimedHit, 0, trimedHit.len + trimedHit.trim5 + trimedHit.trim3, alignMate, dep + 1);
                        if (tmp_maxsc > maxsc) maxsc = tmp_maxsc;
                    }
                    W->poolTop--;
                }
            }
            int64_t msc = minsc[rdi];
            uint32_t mm = (uint32_t)((tempHit.score - msc) / mmpMax);
            uint32_t leftext = HT2_IDX_MAX32, rightext = 0;
            uint32_t num_mismatch_allowed = 1;
            if (hitoff <= minKL) num_mismatch_allowed = tempHit.rdoff < mm ? tempHit.rdoff : mm;
            extend(tempHit, rdi, leftext, rightext, num_mismatch_allowed);
            msc = sinkFloor(rdi, cushion);
            uint32_t need = minKL < hit.rdoff ? minKL : hit.rdoff;
            if (tempHit.score >= msc && leftext >= need) {
                int64_t tmp_maxsc = hybridSearchRecur(rdi, tempHit, tempHit.rdoff, tempHit.len + tempHit.trim3, alignMate, dep + 1);
                if (tmp_maxsc > maxsc) maxsc = tmp_maxsc;
            } else if (hitoff > minKL) {
                uint32_t jumplen = hitoff > minK ? minK : minKL;
                int64_t expected_score = hit.score - (int64_t)((hit.rdoff - hitoff) / jumplen) * mmpMax - mmpMax;
                if (expected_score >= msc) {
                    int64_t tmp_maxsc = hybridSearchRecur(rdi, hit, hitoff - jumplen, hitlen + jumplen, alignMate, dep + 1);
                    if (tmp_maxsc > maxsc) maxsc = tmp_maxsc;
                }
            }
        }
    } else {
        // ---------------- extend to the right (spliced_aligner.h:1362-2049) ---
        bool use_localindex = true;
        if (hit.len == hitlen && hitoff + hitlen + minK > rdlen) {
            uint32_t leftext = 0, rightext = HT2_IDX_MAX32;
            Ht2Hit* t = poolAlloc();
            copyHit(*t, hit);
            extend(*t, rdi, leftext, rightext, 1);
            if (t->rdoff + t->len == rdlen) use_localindex = false;
        }
        int lid = localIndexId(hit.tidx, hit.toff);
        bool success = false, first = true;
        uint32_t count = 0;
        const uint32_t max_count = 2;
        int64_t prev_score = hit.score;
        while (!success && count++ < max_count && use_localindex) {
            if (W->localindexatts >= W->maxLocalindexatts) break;
            if (first) first = false;
            else {
                lid = lid >= 0 ? nextLocal(lid) : -1;
                if (lid < 0 || localGeom(lid)->len == 0) break;
            }
            if (lid < 0) break;
            Ht2Fm<uint16_t> lfm; lfm.init(blob, localGeom(lid));
            uint32_t extlen = 0;
            uint32_t top = 0xffff, bot = 0xffff, node_top = 0xffff, node_bot = 0xffff;
            uint32_t extoff = hitoff + hitlen + minKL;
            if (extoff + 1 < rdlen) extoff += 1;
            if (extoff >= rdlen) extoff = rdlen - 1;
            uint32_t nelt = HT2_IDX_MAX32;
            uint32_t max_nelt = 5;
            bool no_extension = false;
            bool uniqueStop = false;
            uint32_t minUniqueLen = minKL;
            uint32_t maxHitLen = extoff - hitoff - hitlen;
            if (maxHitLen < minKL) maxHitLen = minKL;
            for (; maxHitLen < extoff + 1 && extoff < rdlen;) {
                extlen = 0;
                uniqueStop = false;
                W->localindexatts++;
                if (lfm.g->len == 0) { nelt = 0; top = bot = node_top = node_bot = 0; }
                else nelt = gfmSearch(lfm, rdi, hit.fw != 0, extoff, extlen, top, bot, node_top, node_bot,
                                      uniqueStop, minUniqueLen, maxHitLen & 0xffffu, P->kseeds, true);
                if (extoff < hitoff + hitlen) { no_extension = true; break; }
                if (nelt <= max_nelt) break;
                if (extoff + 1 < rdlen) extoff++;
                else {
                    if (extlen < maxHitLen) break;
                    else maxHitLen++;
                }
            }
            if (nelt > 0 && nelt <= max_nelt && extlen >= P->minAnchorLen && !no_extension) {
                ncoords = 0;
                getGenomeCoordsLocal(lfm, top, bot, node_top, node_bot, hit.fw != 0, extoff + 1 - extlen, extlen, coords, ncoords, 8);
                if (ncoords > 1) sortCoords(coords, ncoords);
                for (uint32_t ri = 0; ri < ncoords; ri++) {
                    const Ht2Coord& coord = coords[ri];
                    Ht2Hit* tp = poolAlloc();
                    Ht2Hit& tempHit = *tp;
                    initHit(tempHit, coord.fw != 0, extoff + 1 - extlen, extlen, 0, 0, coord.ref, coord.off, coord.joinedOff);
                    if (!compatibleWith(hit, tempHit, rdi)) {
                        W->poolTop--;
                        if (count == 1) continue; else break;
                    }
                    uint32_t leftext = 0, rightext = HT2_IDX_MAX32;
                    extend(tempHit, rdi, leftext, rightext, 0);
                    Ht2Hit* cp = poolAlloc();
                    Ht2Hit& combinedHit = *cp;
                    copyHit(combinedHit, hit);
                    int64_t msc = minsc[rdi];
                    bool combined = combineWith(combinedHit, tempHit, rdi, msc);
                    msc = sinkFloor(rdi, cushion);
                    bool keep = false;
                    if (combined && combinedHit.score >= msc) {
                        if (combinedHit.score >= prev_score - mmpMax) {
                            int64_t tmp_maxsc = hybridSearchRecur(rdi, combinedHit, combinedHit.rdoff - combinedHit.trim5,
                                                                  combinedHit.len + combinedHit.trim5, alignMate, dep + 1);
                            if (tmp_maxsc > maxsc) maxsc = tmp_maxsc;
                        } else {
                            if (nLocalHits < 16) {
                                // keep combinedHit: move it into tempHit's slot so the stack stays compact
                                copyHit(tempHit, combinedHit);
                                tempHit.hitcount = combinedHit.hitcount;
                                localHits[nLocalHits++] = (uint16_t)(tp - W->pool);
                                keep = true;
                            } else W->err |= HT2_ERR_POOL;
                        }
                    }
                    W->poolTop--;            // combinedHit
                    if (!keep) W->poolTop--; // tempHit
                }
            }
            if (maxsc >= prev_score - mmpMax) success = true;
            if (!success && (W->localindexatts >= W->maxLocalindexatts || count == max_count || nextLocal(lid) < 0)) {
                for (uint32_t ti = 0; ti < nLocalHits; ti++) {
                    Ht2Hit& tempHit = W->pool[localHits[ti]];
                    int64_t msc = sinkFloor(rdi, cushion);
                    if (tempHit.score >= msc) {
                        int64_t tmp_maxsc = hybridSearchRecur(rdi, tempHit, tempHit.rdoff - tempHit.trim5, tempHit.len + tempHit.trim5, alignMate, dep + 1);
                        if (tmp_maxsc > maxsc) maxsc = tmp_maxsc;
                    }
                }
            }
        }
        if (!success) {
            if (hitoff + hitlen + minK + 1 < rdlen && W->localindexatts < W->maxLocalindexatts) {
                uint32_t extlen = 0;
                uint32_t top = HT2_IDX_MAX32, bot = HT2_IDX_MAX32, node_top = HT2_IDX_MAX32, node_bot = HT2_IDX_MAX32;
                uint32_t extoff = hitoff + hitlen + minK + 1;
                bool uniqueStop = true;
                uint32_t nelt = gfmSearch(gfm, rdi, hit.fw != 0, extoff, extlen, top, bot, node_top, node_bot,
                                          uniqueStop, minK, HT2_IDX_MAX32, P->kseeds, false);
                if (nelt > 0 && nelt <= 5 && extlen >= minK) {
                    W->nCoords = 0;
                    bool straddled = false;
                    getGenomeCoords(top, bot, node_top, node_bot, hit.fw != 0, bot - top, extlen, true, straddled);
                    ncoords = W->nCoords < 8 ? W->nCoords : 8;
                    for (uint32_t i = 0; i < ncoords; i++) coords[i] = W->coords[i];
                    sortCoords(coords, ncoords);
                    for (uint32_t ri = 0; ri < ncoords; ri++) {
                        const Ht2Coord& coord = coords[ri];
                        Ht2Hit* tp = poolAlloc();
                        Ht2Hit& tempHit = *tp;
                        initHit(tempHit, coord.fw != 0, extoff + 1 - extlen, extlen, 0, 0, coord.ref, coord.off, coord.joinedOff);
                        if (!compatibleWith(hit, tempHit, rdi)) { W->poolTop--; continue; }
                        uint32_t leftext = 0, rightext = HT2_IDX_MAX32;
                        extend(tempHit, rdi, leftext, rightext, 0);
                        Ht2Hit* cp = poolAlloc();
                        Ht2Hit& combinedHit = *cp;
                        copyHit(combinedHit, hit);
                        int64_t msc = minsc[rdi];
                        bool combined = combineWith(combinedHit, tempHit, rdi, msc);
                        msc = sinkFloor(rdi, cushion);
                        if (combined && combinedHit.score >= msc) {
                            int64_t tmp_maxsc = hybridSearchRecur(rdi, combinedHit, combinedHit.rdoff - combinedHit.trim5,
                                                                  combinedHit.len + combinedHit.trim5, alignMate, dep + 1);
                            if (tmp_maxsc > maxsc) maxsc = tmp_maxsc;
                        }
                        W->poolTop -= 2;
                    }
                }
            }
            Ht2Hit* tp = poolAlloc();
            Ht2Hit& tempHit = *tp;
            copyHit(tempHit, hit);
            {
                uint32_t trimLen = rdlen - hitoff - tempHit.len - tempHit.trim5;
                int64_t floor_ = maxsc > minsc[rdi] ? maxsc : minsc[rdi];
                uint32_t trimMax = (uint32_t)((tempHit.score - floor_) / ht2_scpen(*P, 0));
                if (trimLen < trimMax) {
                    uint32_t trim3 = rdlen - hitoff - tempHit.len - tempHit.trim5;
                    Ht2Hit* trp = poolAlloc();
                    Ht2Hit& trimedHit = *trp;
                    copyHit(trimedHit, tempHit);
                    trimedHit.trim3 = trim3; // GenomeHit::trim3 (hi_aligner.h:855-876)
                    calculateScore(trimedHit, rdi);
                    int64_t tmp_score = trimedHit.score;
                    if (tmp_score > maxsc && tmp_score >= minsc[rdi]) {
                        int64_t tmp_maxsc = hybridSearchRecur(rdi, trimedHit, trimedHit.rdoff - trimedHit.trim5,
                                                              trimedHit.len + trimedHit.trim5 + trimedHit.trim3, alignMate, dep + 1);
                        if (tmp_maxsc > maxsc) maxsc = tmp_maxsc;
                    }
                    W->poolTop--;
                }
            }
            int64_t msc = minsc[rdi];
            uint32_t leftext = 0, rightext = HT2_IDX_MAX32;
            uint32_t mm = (uint32_t)((tempHit.score - msc) / mmpMax);
            uint32_t num_mismatch_allowed = 1;
            if (rdlen - hitoff - hitlen <= minKL) {
                uint32_t r = rdlen - tempHit.rdoff - tempHit.len;
                num_mismatch_allowed = r < mm ? r : mm;
            }
            extend(tempHit, rdi, leftext, rightext, num_mismatch_allowed);
            msc = sinkFloor(rdi, cushion);
            uint32_t need = rdlen - hit.len - hit.rdoff;
            if (minKL < need) need = minKL;
            if (tempHit.score >= msc && rightext >= need) {
                int64_t tmp_maxsc = hybridSearchRecur(rdi, tempHit, tempHit.rdoff - tempHit.trim5, tempHit.len + tempHit.trim5, alignMate, dep + 1);
                if (tmp_maxsc > maxsc) maxsc = tmp_maxsc;
            } else if (hitoff + hitlen + minKL < rdlen) {
                uint32_t jumplen = hitoff + hitlen + minK < rdlen ? minK : minKL;
                int64_t expected_score = hit.score - (int64_t)((hitlen - hit.len) / jumplen) * mmpMax - mmpMax;
                if (expected_score >= msc) {
                    int64_t tmp_maxsc = hybridSearchRecur(rdi, hit, hitoff, hitlen + jumplen, alignMate, dep + 1);
                    if (tmp_maxsc > maxsc) maxsc = tmp_maxsc;
                }
            }
        }
    }
    W->poolTop = poolMark;
    return maxsc;
}

// PairedEndPolicy::peClassifyPair (pe.cpp:38-132) with the hisat2 defaults
// olapOk = containOk = expandToFit = true, dovetailOk = false (hisat2.cpp:348-352).
// Returns true iff the pair is NOT discordant.
template <bool GRAPH> HT2_NI bool Ht2AlignerT<GRAPH>::peConcordant(int64_t off1, uint32_t len1, bool fw1, int64_t off2, uint32_t len2, bool fw2) const
{
    uint32_t maxfrag = P->maxFrag;
    if (len1 > maxfrag) maxfrag = len1;
    if (len2 > maxfrag) maxfrag = len2;
    uint32_t minfrag = P->minFrag;
    if (minfrag < 1) minfrag = 1;
    bool oneLeft = false;
    switch (P->pePolicy) {
        case 0: if (fw1 != fw2) return false; oneLeft = fw1; break;   // FF
        case 1: if (fw1 != fw2) return false; oneLeft = !fw1; break;  // RR
        case 2: if (fw1 == fw2) return false; oneLeft = fw1; break;   // FR
        default: if (fw1 == fw2) return false; oneLeft = !fw1; break; // RF
    }
    int64_t fraglo = off1 < off2 ? off1 : off2;
    int64_t h1 = off1 + len1, h2 = off2 + len2;
    int64_t fraghi = h1 > h2 ? h1 : h2;
    uint64_t frag = (uint64_t)(fraghi - fraglo);
    if (frag > maxfrag || frag < minfrag) return false;
    int64_t lo1 = off1, hi1 = off1 + len1 - 1, lo2 = off2, hi2 = off2 + len2 - 1;
    bool containment = (lo1 >= lo2 && hi1 <= hi2) || (lo2 >= lo1 && hi2 <= hi1);
    bool olap = (lo1 <= lo2 && hi1 >= lo2) || (lo1 <= hi2 && hi1 >= hi2) || containment;
    if (!olap) {
        if ((oneLeft && lo2 < lo1) || (!oneLeft && lo1 < lo2)) return false;
    }
    if ((oneLeft && (hi1 > hi2 || lo2 < lo1)) || (!oneLeft && (hi2 > hi1 || lo1 < lo2))) return false; // dovetail
    return true;
}

// HI_Aligner::pairReads (hi_aligner.h:5948-6057) + AlnSinkWrap::report for a
// pair (aln_sink.h:2565-2611) + ReportingState::foundConcordant (aln_sink.cpp:72-92).
template <bool GRAPH> HT2_NI void Ht2AlignerT<GRAPH>::pairReads()
{
    const uint32_t n1 = W->nRes[0], n2 = W->nRes[1];
    uint32_t start_i = W->concordInspected[0], start_j = W->concordInspected[1];
    W->concordInspected[0] = n1;
    W->concordInspected[1] = n2;
    for (uint32_t i = 0; i < n1; i++) {
        for (uint32_t j = (i >= start_i ? 0 : start_j); j < n2; j++) {
            const Ht2Res& r1 = W->res[0][i];
            const Ht2Res& r2 = W->res[1][j];
            if (r1.tidx != r2.tidx) continue;
            // Coord left/right of each mate; orient() == fw
            int64_t left = r1.toff, right = (int64_t)r1.toff + r1.rfextent - 1;
            int64_t left2 = r2.toff, right2 = (int64_t)r2.toff + r2.rfextent - 1;
#ifdef HT2_ENABLE_SPLICED   // AlnRes::refcoord_right adds the introns (aligner_result.h:1255-1268)
            for (uint32_t e = 0; e < r1.nedits; e++) if (r1.edits[e].type == HT2_EDIT_SPL) right += ht2_spl_len(r1.edits[e]);
            for (uint32_t e = 0; e < r2.nedits; e++) if (r2.edits[e].type == HT2_EDIT_SPL) right2 += ht2_spl_len(r2.edits[e]);
#endif
            if ((r1.fw != 0) == (P->gMate1fw != 0)) {
                if ((r2.fw != 0) != (P->gMate2fw != 0)) continue;
            } else {
                if ((r2.fw != 0) == (P->gMate2fw != 0)) continue;
                int64_t t = left; left = left2; left2 = t;
                t = right; right = right2; right2 = t;
            }
            if (left > left2) continue;
            if (right > right2) continue;
            if (right + (int64_t)(int)P->maxIntronLen < left2) continue;
            bool dna_frag_pass = true;
            if (P->noSplicedAlignment) {
                if (r1.toff < r2.toff) dna_frag_pass = peConcordant(r1.toff, r1.rfextent, r1.fw != 0, r2.toff, r2.rfextent, r2.fw != 0);
                else dna_frag_pass = peConcordant(r2.toff, r2.rfextent, r2.fw != 0, r1.toff, r1.rfextent, r1.fw != 0);
            }
            if (!P->noSplicedAlignment || dna_frag_pass) {
                int64_t threshold = W->bestPair;
                if (W->bestUnp[0] >= minsc[0] && W->bestUnp[1] >= minsc[1]) {
                    double t = (double)(W->bestUnp[0] + W->bestUnp[1]) -
                               (double)(uint64_t)((uint64_t)r1.rdlen + (uint64_t)r2.rdlen) * 0.03 * (double)ht2_mmpen(*P, 255);
                    int64_t tmp = (int64_t)t;
                    if (tmp > threshold) threshold = tmp;
                }
                int64_t score = r1.score + r2.score;
                if (score >= threshold || P->secondary) {
                    // sink.report(0, &r1, &r2)
                    if (score > W->concordBest) { W->concordBest = score; W->nconcord = 0; }
                    W->nconcord++;
                    if (W->nPairs >= HT2_MAX_PAIRS) { W->err |= HT2_ERR_PAIRS; return; }
                    W->pairs[W->nPairs][0] = (uint16_t)i;
                    W->pairs[W->nPairs][1] = (uint16_t)j;
                    W->nPairs++;
                    if (score > W->bestPair) { W->best2Pair = W->bestPair; W->bestPair = score; }
                    else if (score > W->best2Pair) W->best2Pair = score;
                }
            }
        }
    }
}

// HI_Aligner::alignMate (hi_aligner.h:5579-5767): use the alignment of one mate
// (rdi, at tidx:toff) as an anchor and search the local index(es) around it for
// the other mate.
template <bool GRAPH> HT2_NI bool Ht2AlignerT<GRAPH>::alignMateFn(uint32_t rdi, bool fw, uint32_t tidx, uint32_t toff)
{
    const uint32_t ordi = 1 - rdi;
    alignMateAnchors(rdi, fw, tidx, toff);
    for (uint32_t hi = 0; hi < W->nGenomeHits; hi++) {
        Ht2Hit& gh = W->genomeHits[hi];
        uint32_t leftext = HT2_IDX_MAX32, rightext = HT2_IDX_MAX32;
        extend(gh, ordi, leftext, rightext, 0);
        hybridSearchRecur(ordi, gh, gh.rdoff, gh.len, true, 0);
    }
    return true;
}

// anchor search part of alignMate (hi_aligner.h:5600-5717): fills W->genomeHits
template <bool GRAPH> HT2_NI void Ht2AlignerT<GRAPH>::alignMateAnchors(uint32_t rdi, bool fw, uint32_t tidx, uint32_t toff)
{
    const uint32_t ordi = 1 - rdi;
    const bool ofw = (fw == (P->gMate2fw != 0)) ? (P->gMate1fw != 0) : (P->gMate2fw != 0);
    const uint32_t rdlen = W->rd[ordi].len;
    const uint32_t minKL = P->minKLocal;
    W->nGenomeHits = 0;
    Ht2Coord coords[HT2_MAX_COORDS];
    int lid = localIndexId(tidx, toff);
    bool first = true;
    uint32_t count = 0;
    uint32_t max_hitlen = 0;
    while (count++ < 2) {
        if (first) first = false;
        else {
            if (W->nGenomeHits > 0) break;
            lid = lid >= 0 ? (fw ? nextLocal(lid) : prevLocal(lid)) : -1;
            if (lid < 0 || localGeom(lid)->len == 0) break;
        }
        if (lid < 0 || localGeom(lid)->len == 0) break;
        Ht2Fm<uint16_t> lfm; lfm.init(blob, localGeom(lid));
        uint32_t hitoff = rdlen - 1;
        while (hitoff >= minKL - 1) {
            uint32_t hitlen = 0;
            uint32_t top = 0xffff, bot = 0xffff, node_top = 0xffff, node_bot = 0xffff;
            bool uniqueStop = false;
            uint32_t nelt = gfmSearch(lfm, ordi, ofw, hitoff, hitlen, top, bot, node_top, node_bot,
                                      uniqueStop, minKL, 0xffffu, P->kseeds, true);
            if (nelt > 0 && nelt <= P->kseeds && hitlen > max_hitlen) {
                uint32_t ncoords = 0;
                getGenomeCoordsLocal(lfm, top, bot, node_top, node_bot, ofw, hitoff - hitlen + 1, hitlen, coords, ncoords, HT2_MAX_COORDS);
                W->nGenomeHits = 0;
                for (uint32_t ri = 0; ri < ncoords; ri++) {
                    const Ht2Coord& coord = coords[ri];
                    if (P->noSplicedAlignment) {
                        if (coord.off + P->maxFrag * 2 < toff || toff + P->maxFrag * 2 < coord.off) continue;
                    }
                    if (W->nGenomeHits >= HT2_MAX_GHITS) { W->err |= HT2_ERR_GHITS; break; }
                    adjustWithALTCoord(hitoff - hitlen + 1, hitlen, coord, ordi);   // plain init on linear indexes (hi_aligner.h:5692)
                }
                max_hitlen = hitlen;
            }
            if (hitlen > 0) hitoff -= (hitlen - 1);
            if (hitoff > 0) hitoff -= 1;
        }
    }
    const uint32_t maxsize = P->kseeds;
    if (W->nGenomeHits > maxsize) {
        uint32_t left = W->nGenomeHits;
        for (uint32_t i = 0; i + 1 < W->nGenomeHits; i++) {
            uint32_t rndi = W->rnd.nextU32() % left;
            if (rndi > 0) {
                Ht2Hit* t = poolAlloc();
                copyHit(*t, W->genomeHits[i]); uint32_t hc = W->genomeHits[i].hitcount;
                copyHit(W->genomeHits[i], W->genomeHits[i + rndi]); W->genomeHits[i].hitcount = W->genomeHits[i + rndi].hitcount;
                copyHit(W->genomeHits[i + rndi], *t); W->genomeHits[i + rndi].hitcount = hc;
                W->poolTop--;
            }
            left--;
        }
        W->nGenomeHits = maxsize;
    }
}

// HI_Aligner::go (hi_aligner.h:4048-4149); the repeat-index block
// (:4151-4636) runs only when a .rep index is loaded.
template <bool GRAPH> HT2_NI void Ht2AlignerT<GRAPH>::go()
{
    for (uint32_t rdi = 0; rdi < 2; rdi++) {
        for (uint32_t fwi = 0; fwi < 2; fwi++) {
            Ht2ReadHits& h = W->hits[rdi][fwi];
            h.len = W->rd[rdi].len; h.cur = 0; h.done = 0; h.numPartialSearch = 0; h.numUniqueSearch = 0; h.nhits = 0; h.nie = 0;
        }
        W->nSearched[rdi] = 0;
    }
    W->searchedTop = 0;
    W->nGenomeHits = 0;
    W->poolTop = 0;
    W->concordInspected[0] = W->concordInspected[1] = 0;
    uint32_t rdi; bool fw;
    bool found[2][2] = {{true, true}, {paired, paired}};
    while (nextBWT(rdi, fw)) {
        uint32_t fwi = fw ? 0 : 1;
        found[rdi][fwi] = align(rdi, fw);
        if (!found[0][0] && !found[0][1] && !found[1][0] && !found[1][1]) break;
        if (paired) pairReads();
        if (W->err) break;
    }
    if (paired && !W->err) {
        if (W->nPairs == 0 && (W->bestUnp[0] >= minsc[0] || W->bestUnp[1] >= minsc[1])) {
            bool mate_found = false;
            uint32_t rs_size[2] = {W->nRes[0], W->nRes[1]};
            for (uint32_t i = 0; i < 2; i++) {
                for (uint32_t j = 0; j < rs_size[i]; j++) {
                    const Ht2Res& res = W->res[i][j];
                    mate_found |= alignMateFn(i, res.fw != 0, res.tidx, res.toff);
                }
            }
            if (mate_found) pairReads();
        }
    }
}

#endif // HT2_CORE_IMPL_H_
