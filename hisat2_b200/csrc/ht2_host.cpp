// ht2_host.cpp -- see ht2_host.h.
#include "ht2_host.h"

#include <stdio.h>
#include <string.h>

#include <algorithm>

// ---------------------------------------------------------------------------
// parameters
// ---------------------------------------------------------------------------
void ht2_default_params(Ht2Params& P, const Ht2Image& img, bool noSplicedAlignment)
{
    memset(&P, 0, sizeof(P));
    const Ht2ImageHeader* H = img.header();
    P.mmpMax = 6; P.mmpMin = 2; P.scpMax = 2; P.scpMin = 1; P.npen = 1;
    P.rdGapConst = 5; P.rdGapLinear = 3; P.rfGapConst = 5; P.rfGapLinear = 3;
    P.mmcostConstant = 0;
    P.canSplPen = 0;
    P.khits = H->global.linearFM ? 5 : 10;              // hisat2.cpp:3903-3906
    P.kseeds = std::max<uint32_t>(5, P.khits * 2);      // hisat2.cpp:3174-3176
    P.secondary = 0;
    P.minIntronLen = 20; P.maxIntronLen = 500000;
    P.minAnchorLen = 7; P.minAnchorLenNoncan = 14;
    P.noSplicedAlignment = noSplicedAlignment ? 1 : 0;
    P.maxAltsTried = 16;
    P.anchorStop = 1;
    uint32_t genomeLen = H->global.len;                 // hi_aligner.h:3979-3985
    P.minK = 0;
    while (genomeLen > 0) { genomeLen >>= 2; P.minK++; }
    P.minKLocal = 8;
    P.pePolicy = 2; P.minFrag = 0; P.maxFrag = 1000;
    P.gMate1fw = 1; P.gMate2fw = 0;
    P.nofw = 0; P.norc = 0; P.mixed = 1; P.discord = 1;
}

// SimpleFunc::f<T> for the linear functions the defaults use (simple_func.h:86-108)
static double simpleLinear(double I, double X, double C, double L, double x) {
    return std::max(I, std::min(X, C + L * x));
}

int64_t ht2_minsc(uint32_t rdlen)
{
    // scoreMin = L,0,-0.2 (hisat2.cpp:441); clamped to <= 0 in end-to-end mode (:3395-3402)
    double v = simpleLinear(-1.7976931348623157e308, 1.7976931348623157e308, (double)0.0f, (double)-0.2f, (double)rdlen);
    int64_t m = (int64_t)v;
    if (m > 0) m = 0;
    return m;
}

Ht2ReadFilters ht2_filters(const Ht2HostRead& rd, int64_t minsc)
{
    Ht2ReadFilters f;
    size_t rdlen = rd.seq.size();
    // Scoring::nFilter (scoring.cpp:104-117); nCeil = L,0,0.15?? no: hisat2.cpp:443
    // nCeil.init(SIMPLE_FUNC_LINEAR, 0.0f, DMAX, 2.0f, 0.1f)
    size_t maxns = (size_t)simpleLinear(0.0, 1.7976931348623157e308, (double)2.0f, (double)0.1f, (double)rdlen);
    size_t ns = 0;
    f.nfilt = true;
    for (size_t i = 0; i < rdlen; i++) {
        if (rd.seq[i] == 4) { ns++; if (ns > maxns) { f.nfilt = false; break; } }
    }
    f.scfilt = (0 >= minsc);                 // Scoring::scoreFilter with match bonus 0
    f.lenfilt = true;
    if (rdlen <= 0 /*multiseedMms*/ || rdlen < 2) f.lenfilt = false;
    f.qcfilt = true;
    return f;
}

uint32_t ht2_gen_rand_seed(const Ht2HostRead& rd, uint32_t seed)
{
    uint32_t rseed = (seed + 101) * 59 * 61 * 67 * 71 * 73 * 79 * 83;
    size_t qlen = rd.seq.size();
    for (size_t i = 0; i < qlen; i++) {
        int p = (int)rd.seq[i];
        size_t off = ((i & 15) << 1);
        rseed ^= ((uint32_t)p << off);
    }
    for (size_t i = 0; i < qlen; i++) {
        int p = (int)rd.qual[i];
        size_t off = ((i & 3) << 3);
        rseed ^= ((uint32_t)p << off);
    }
    for (size_t i = 0; i < rd.name.size(); i++) {
        int p = (int)rd.name[i];
        if (p == '/') break;
        size_t off = ((i & 3) << 3);
        rseed ^= ((uint32_t)p << off);
    }
    return rseed;
}

void ht2_fill_read(Ht2Read& dst, const Ht2HostRead& src)
{
    uint32_t n = (uint32_t)src.seq.size();
    if (n > HT2_MAX_RDLEN) n = HT2_MAX_RDLEN;
    dst.len = n;
    for (uint32_t i = 0; i < n; i++) {
        dst.seq[0][i] = src.seq[i];
        dst.qual[0][i] = src.qual[i];
        uint8_t c = src.seq[n - i - 1];
        dst.seq[1][i] = c < 4 ? (uint8_t)(c ^ 3) : (uint8_t)4;
        dst.qual[1][i] = src.qual[n - i - 1];
    }
}

// ---------------------------------------------------------------------------
// FASTA (pat.cpp:725-849; alphabet.cpp asc2dnacat / asc2dna)
// ---------------------------------------------------------------------------
static int dnacat(int c) {
    switch (c) {
        case 'A': case 'C': case 'G': case 'T': case 'a': case 'c': case 'g': case 't': return 1;
        case 'B': case 'D': case 'H': case 'K': case 'M': case 'N': case 'R': case 'S': case 'V': case 'W': case 'X': case 'Y':
        case 'b': case 'd': case 'h': case 'k': case 'm': case 'n': case 'r': case 's': case 'v': case 'w': case 'x': case 'y': return 2;
        case '-': return 3;
        default: return 0;
    }
}
static uint8_t asc2dna(int c) {
    switch (c) {
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        case 'N': case 'n': return 4;
        default: return 0;
    }
}

bool ht2_read_fasta(const char* path, std::vector<Ht2HostRead>& out, int mate, std::string& err)
{
    FILE* f = fopen(path, "rb");
    if (!f) { err = std::string("could not open reads file ") + path; return false; }
    std::vector<char> buf;
    {
        fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
        buf.resize((size_t)sz);
        if (sz > 0 && fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) { fclose(f); err = "short read"; return false; }
        fclose(f);
    }
    size_t p = 0, n = buf.size();
    uint64_t readCnt = 0;
    while (p < n) {
        // skip comment / blank lines
        while (p < n && (buf[p] == '#' || buf[p] == ';' || buf[p] == '\r' || buf[p] == '\n')) {
            if (buf[p] == '#' || buf[p] == ';') { while (p < n && buf[p] != '\n') p++; }
            else p++;
        }
        if (p >= n) break;
        if (buf[p] != '>') { err = "reads file does not look like a FASTA file"; return false; }
        p++;
        Ht2HostRead r;
        r.mate = mate;
        while (p < n && buf[p] != '\n' && buf[p] != '\r') r.name.push_back(buf[p++]);
        while (p < n && (buf[p] == '\n' || buf[p] == '\r')) p++;
        while (p < n && buf[p] != '>') {
            int c = (unsigned char)buf[p++];
            if (dnacat(c) > 0) { r.seq.push_back(asc2dna(c)); r.qual.push_back('I'); }
        }
        if (r.name.empty()) r.name = std::to_string(readCnt);
        readCnt++;
        if (r.seq.empty()) continue; // "skipping empty FASTA read"
        out.push_back(r);
    }
    return true;
}

// ---------------------------------------------------------------------------
// SAM back end
// ---------------------------------------------------------------------------
void ht2_sam_header(std::string& o, const Ht2Image& img)
{
    o += "@HD\tVN:1.0\tSO:unsorted\n";
    const Ht2ImageHeader* H = img.header();
    for (uint32_t i = 0; i < H->nRefs; i++) {
        o += "@SQ\tSN:";
        const char* nm = img.refName(i);
        for (const char* c = nm; *c && !isspace((unsigned char)*c); c++) o.push_back(*c);
        o += "\tLN:";
        o += std::to_string(img.refPlen(i));
        o += "\n";
    }
}

namespace {

// AlnScore::calculate_hisat2_score (aligner_result.h:322-348) with repeat=0,
// no transcripts, splicescore 0.
int64_t hisat2Score(const Ht2Res& r)
{
    int64_t score = r.score;
    if (score > 0x7fffffffll) score = 0x7fffffffll;
    else if (score < -0x80000000ll) score = -0x80000000ll;
    int64_t splicescore = 255;
    int64_t trim = (int64_t)r.trim5p + (int64_t)r.trim3p; // leftTrim+rightTrim (hit.trim5+hit.trim3)
    if (trim > 65535) trim = 0; else trim = 65535 - trim;
    return (int64_t)((uint64_t)score << 32) | (splicescore << 16) | trim;
}

struct ScoreKey { int64_t score; int64_t h2; bool valid; };
bool keyGt(const ScoreKey& a, const ScoreKey& b) { // AlnScore::operator> (aligner_result.h:143-157)
    if (!b.valid) return a.valid;
    if (!a.valid) return false;
    return a.score > b.score || (a.score == b.score && a.h2 > b.h2);
}
bool keyEq(const ScoreKey& a, const ScoreKey& b) {
    return a.valid && b.valid && a.score == b.score && a.h2 == b.h2;
}

// AlnSinkWrap::selectByScore (aln_sink.h:2680-2755)
void selectByScore(const std::vector<Ht2Res>& rs1, const std::vector<Ht2Res>* rs2,
                   const std::vector<std::pair<uint16_t, uint16_t> >* pairs,
                   uint64_t num, std::vector<size_t>& select, Ht2Rng& rnd, bool secondary)
{
    size_t sz = pairs ? pairs->size() : rs1.size();
    if (sz < num) num = sz;
    select.clear();
    if (sz < 1) return;
    std::vector<std::pair<int64_t, size_t> > buf(sz);
    for (size_t i = 0; i < sz; i++) {
        if (pairs) buf[i].first = hisat2Score(rs1[(*pairs)[i].first]) + hisat2Score((*rs2)[(*pairs)[i].second]);
        else buf[i].first = hisat2Score(rs1[i]);
        buf[i].second = i;
    }
    std::sort(buf.begin(), buf.end());
    std::reverse(buf.begin(), buf.end());
    auto shufflePortion = [&](size_t begin, size_t cnt) {
        if (cnt < 2) return;
        size_t left = cnt;
        for (size_t i = begin; i < begin + cnt - 1; i++) {
            uint32_t rndi = rnd.nextU32() % (uint32_t)left;
            if (rndi > 0) std::swap(buf[i], buf[i + rndi]);
            left--;
        }
    };
    size_t streak = 0;
    for (size_t i = 1; i < buf.size(); i++) {
        if (buf[i].first == buf[i - 1].first) {
            if (streak == 0) streak = 1;
            streak++;
        } else {
            if (streak > 1) shufflePortion(i - streak, streak);
            streak = 0;
        }
    }
    if (streak > 1) shufflePortion(buf.size() - streak, streak);
    for (size_t i = 0; i < buf.size(); i++) {
        if (i >= num) break; // no repeat alignments in this build
        select.push_back(buf[i].second);
    }
    if (!secondary) {
        for (size_t i = 0; i + 1 < select.size(); i++) {
            if (buf[i].first != buf[i + 1].first) { select.resize(i + 1); break; }
        }
    }
}

struct Summ { // AlnSetSumm (aligner_result.cpp:1167-1260)
    ScoreKey best[2], secbest[2], bestPaired, secbestPaired;
    bool paired;
    size_t numAlns[2], numAlnsPaired;
    int64_t orefid, orefoff;
    void reset() {
        best[0].valid = best[1].valid = secbest[0].valid = secbest[1].valid = false;
        bestPaired.valid = secbestPaired.valid = false;
        paired = false; numAlns[0] = numAlns[1] = numAlnsPaired = 0; orefid = -1; orefoff = -1;
    }
    void addUnp(int j, const std::vector<Ht2Res>& rs) {
        for (size_t i = 0; i < rs.size(); i++) {
            ScoreKey sc = {rs[i].score, hisat2Score(rs[i]), true};
            if (keyGt(sc, best[j])) { secbest[j] = best[j]; best[j] = sc; }
            else if (keyGt(sc, secbest[j])) secbest[j] = sc;
        }
        numAlns[j] = rs.size();
    }
};

// BowtieMapq2::mapq (unique.h:170-400) for monotone scoring, canMax=false,
// exhausted=false.
int mapqV2(const Summ& s, bool mate1, size_t rdlen, size_t ordlen)
{
    const ScoreKey& bst = s.paired ? s.bestPaired : s.best[mate1 ? 0 : 1];
    const ScoreKey& sec = s.paired ? s.secbestPaired : s.secbest[mate1 ? 0 : 1];
    bool hasSecbest = sec.valid;
    bool equalSecbest = hasSecbest && keyEq(bst, sec);
    if (!hasSecbest || !equalSecbest) return 60;
    int64_t scPer = 0;
    int64_t scMin = (int64_t)((double)0.0f + (double)-0.2f * (double)(float)rdlen);
    if (s.paired) scMin += (int64_t)((double)0.0f + (double)-0.2f * (double)(float)ordlen);
    int64_t diff = scPer - scMin;
    int64_t best = bst.score;
    int64_t bestOver = best - scMin;
    int64_t secbest = sec.score;
    long a = labs((long)best), b = labs((long)secbest);
    int64_t bestdiff = labs(a - b);
    int ret;
    if (bestdiff >= diff * (double)0.9f) ret = (bestOver == diff) ? 39 : 33;
    else if (bestdiff >= diff * (double)0.8f) ret = (bestOver == diff) ? 38 : 27;
    else if (bestdiff >= diff * (double)0.7f) ret = (bestOver == diff) ? 37 : 26;
    else if (bestdiff >= diff * (double)0.6f) ret = (bestOver == diff) ? 36 : 22;
    else if (bestdiff >= diff * (double)0.5f) {
        if (bestOver == diff) ret = 35;
        else if (bestOver >= diff * (double)0.84f) ret = 25;
        else if (bestOver >= diff * (double)0.68f) ret = 16;
        else ret = 5;
    } else if (bestdiff >= diff * (double)0.4f) {
        if (bestOver == diff) ret = 34;
        else if (bestOver >= diff * (double)0.84f) ret = 21;
        else if (bestOver >= diff * (double)0.68f) ret = 14;
        else ret = 4;
    } else if (bestdiff >= diff * (double)0.3f) {
        if (bestOver == diff) ret = 32;
        else if (bestOver >= diff * (double)0.88f) ret = 18;
        else if (bestOver >= diff * (double)0.67f) ret = 15;
        else ret = 3;
    } else if (bestdiff >= diff * (double)0.2f) {
        if (bestOver == diff) ret = 31;
        else if (bestOver >= diff * (double)0.88f) ret = 17;
        else if (bestOver >= diff * (double)0.67f) ret = 11;
        else ret = 0;
    } else if (bestdiff >= diff * (double)0.1f) {
        if (bestOver == diff) ret = 30;
        else if (bestOver >= diff * (double)0.88f) ret = 12;
        else if (bestOver >= diff * (double)0.67f) ret = 7;
        else ret = 0;
    } else if (bestdiff > 0) {
        ret = (bestOver >= diff * (double)0.67f) ? 6 : 2;
    } else {
        ret = (bestOver >= diff * (double)0.67f) ? 1 : 0;
    }
    return ret;
}

struct Stacked { // StackedAln (aligner_result.h:723-895, aligner_result.cpp:660-1000)
    std::string ref, rel, read;
    std::vector<bool> snp;
    size_t trimLS, trimRS;
    void init(const std::vector<uint8_t>& s, const Ht2Edit* ed, size_t ned, size_t tLS, size_t tRS) {
        ref.clear(); rel.clear(); read.clear(); snp.clear();
        trimLS = tLS; trimRS = tRS;
        size_t rdoff = tLS;
        for (size_t i = 0; i < ned; i++) {
            size_t pos = ed[i].pos + tLS;
            while (rdoff < pos) {
                int c = s[rdoff++];
                ref.push_back("ACGTN"[c]); rel.push_back('='); snp.push_back(false); read.push_back("ACGTN"[c]);
            }
            bool isSnp = ed[i].snpID != HT2_IDX_MAX32;
            if (ed[i].type == HT2_EDIT_MM) {
                int c = s[rdoff++];
                ref.push_back((char)ed[i].chr); rel.push_back('X'); snp.push_back(isSnp); read.push_back("ACGTN"[c]);
            } else if (ed[i].type == HT2_EDIT_REF_GAP) {
                int c = s[rdoff++];
                ref.push_back('-'); rel.push_back('I'); snp.push_back(isSnp); read.push_back("ACGTN"[c]);
            } else if (ed[i].type == HT2_EDIT_READ_GAP) {
                ref.push_back((char)ed[i].chr); rel.push_back('D'); snp.push_back(isSnp); read.push_back('-');
            }
        }
        while (rdoff < s.size() - tRS) {
            int c = s[rdoff++];
            ref.push_back("ACGTN"[c]); rel.push_back('='); snp.push_back(false); read.push_back("ACGTN"[c]);
        }
    }
    void leftAlign(bool pastMms) {
        size_t ln = ref.size();
        for (size_t i = 0; i < ln; i++) {
            int r = rel[i];
            if (r != '=' && r != 'X' && r != 'N') {
                if (snp[i]) continue;
                size_t glen = 1;
                for (size_t j = i + 1; j < ln; j++) { if (r != (int)rel[j]) break; glen++; }
                size_t l = i - 1;
                size_t rr = l + glen;
                std::string& gp = (r == 'I') ? ref : read;
                const std::string& ngp = (r == 'I') ? read : ref;
                while (l > 0 && l < ln && ngp[l] == ngp[rr]) {
                    if (rel[l] == 'I' || rel[l] == 'D') break;
                    if (!pastMms && (rel[l] == 'X' || rel[l] == 'N')) break;
                    std::swap(gp[l], gp[rr]);
                    std::swap(rel[l], rel[rr]);
                    l--; rr--;
                }
                i += (glen - 1);
            }
        }
    }
    void cigar(std::string& o) const {
        if (trimLS > 0) { o += std::to_string(trimLS); o.push_back('S'); }
        size_t ln = ref.size();
        for (size_t i = 0; i < ln; i++) {
            char op = rel[i];
            if (op == 'X' || op == '=') op = 'M';
            size_t run = 1;
            for (; i + run < ln; run++) {
                char op2 = rel[i + run];
                if (op2 == 'X' || op2 == '=') op2 = 'M';
                if (op2 != op) break;
            }
            i += (run - 1);
            o += std::to_string(run); o.push_back(op);
        }
        if (trimRS > 0) { o += std::to_string(trimRS); o.push_back('S'); }
    }
    void mdz(std::string& o) const {
        bool mm_last = false, rdgap_last = false, first_print = true;
        size_t ln = ref.size();
        for (size_t i = 0; i < ln; i++) {
            char op = rel[i];
            if (op == '=') {
                size_t run = 1, ninserts = 0;
                for (; i + run < ln; run++) {
                    if (rel[i + run] == '=') {}
                    else if (rel[i + run] == 'I') ninserts++;
                    else break;
                }
                i += (run - 1);
                size_t r = run - ninserts;
                if (r > 0) { o += std::to_string(r); first_print = false; mm_last = false; rdgap_last = false; }
            } else if (op == 'X') {
                if (rdgap_last || mm_last || first_print) o.push_back('0');
                o.push_back(ref[i]);
                first_print = false; mm_last = true; rdgap_last = false;
            } else if (op == 'D') {
                if (mm_last || first_print) o.push_back('0');
                if (!rdgap_last) o.push_back('^');
                o.push_back(ref[i]);
                first_print = false; mm_last = false; rdgap_last = true;
            }
        }
        if (mm_last || rdgap_last) o.push_back('0');
    }
};

void invertPossHost(std::vector<Ht2Edit>& ed, size_t sz) {
    std::reverse(ed.begin(), ed.end());
    for (size_t i = 0; i < ed.size(); i++) {
        if (ed[i].type == HT2_EDIT_READ_GAP || ed[i].type == HT2_EDIT_SPL) ed[i].pos = (uint32_t)(sz - ed[i].pos);
        else ed[i].pos = (uint32_t)(sz - ed[i].pos - 1);
    }
}

void appendName(std::string& o, const std::string& name, bool omitSlashMate) {
    size_t namelen = name.size();
    if (omitSlashMate && namelen >= 2 && name[namelen - 2] == '/' &&
        (name[namelen - 1] == '1' || name[namelen - 1] == '2' || name[namelen - 1] == '3')) namelen -= 2;
    if (namelen > 255) namelen = 255;
    for (size_t i = 0; i < namelen; i++) {
        if (isspace((unsigned char)name[i])) return;
        o.push_back(name[i]);
    }
}
void appendRefName(std::string& o, const Ht2Image& img, uint32_t tidx) {
    for (const char* c = img.refName(tidx); *c && !isspace((unsigned char)*c); c++) o.push_back(*c);
}
void appendSeqQual(std::string& o, const Ht2HostRead& rd, bool fw) {
    size_t n = rd.seq.size();
    if (fw) for (size_t i = 0; i < n; i++) o.push_back("ACGTN"[rd.seq[i]]);
    else for (size_t i = 0; i < n; i++) { uint8_t c = rd.seq[n - i - 1]; o.push_back("ACGTN"[c < 4 ? (c ^ 3) : 4]); }
    o.push_back('\t');
    if (fw) for (size_t i = 0; i < n; i++) o.push_back((char)rd.qual[i]);
    else for (size_t i = 0; i < n; i++) o.push_back((char)rd.qual[n - i - 1]);
}
void appendYF(std::string& o, const Ht2ReadFilters& f) {
    const char* flag = "";
    if (!f.lenfilt) flag = "LN";
    else if (!f.nfilt) flag = "NS";
    else if (!f.scfilt) flag = "SC";
    else if (!f.qcfilt) flag = "QC";
    if (*flag) { o += "\tYF:Z:"; o += flag; }
}

} // namespace

// AlnSinkSam::appendMate (aln_sink.h:3024-3250) for an unpaired read
static void appendMateUnpaired(std::string& o, const Ht2Image& img, const Ht2HostRead& rd, const Ht2ReadFilters& f,
                               const Ht2Res* rs, const Summ& summ, bool primary)
{
    appendName(o, rd.name, false);
    o.push_back('\t');
    int fl = 0;
    if (!primary) fl |= 256;
    if (rs != NULL && !rs->fw) fl |= 16;
    if (rs == NULL) fl |= 4;
    o += std::to_string(fl);
    o.push_back('\t');
    if (rs == NULL) {
        o += "*\t0\t0\t*\t*\t0\t0\t";
        appendSeqQual(o, rd, true);
        o += "\tYT:Z:UU";
        appendYF(o, f);
        o.push_back('\n');
        return;
    }
    Stacked st;
    {
        // AlnRes::initStacked (aligner_result.h:1856-1873)
        size_t trimLS = rs->trim5p, trimRS = rs->trim3p;
        size_t len_trimmed = rd.seq.size() - trimLS - trimRS;
        std::vector<Ht2Edit> ned(rs->edits, rs->edits + rs->nedits);
        std::vector<uint8_t> s(rd.seq);
        if (!rs->fw) {
            invertPossHost(ned, len_trimmed);
            std::swap(trimLS, trimRS);
            size_t n = s.size();
            for (size_t i = 0; i < n; i++) { uint8_t c = rd.seq[n - i - 1]; s[i] = c < 4 ? (uint8_t)(c ^ 3) : (uint8_t)4; }
        }
        st.init(s, ned.data(), ned.size(), trimLS, trimRS);
        st.leftAlign(false);
    }
    appendRefName(o, img, rs->tidx);
    o.push_back('\t');
    o += std::to_string((int64_t)rs->toff + 1);
    o.push_back('\t');
    o += std::to_string(mapqV2(summ, true, rd.seq.size(), 0));
    o.push_back('\t');
    st.cigar(o);
    o += "\t*\t0\t0\t";
    appendSeqQual(o, rd, rs->fw != 0);
    // optional flags (sam.h:525-1010)
    o += "\tAS:i:"; o += std::to_string(rs->score);
    if (summ.secbest[0].valid) { o += "\tZS:i:"; o += std::to_string(summ.secbest[0].score); }
    o += "\tXN:i:0";
    size_t num_mm = 0, num_go = 0, num_gx = 0, NM = 0;
    for (size_t i = 0; i < rs->nedits; i++) {
        const Ht2Edit& e = rs->edits[i];
        NM++;
        if (e.type == HT2_EDIT_MM) num_mm++;
        else if (e.type == HT2_EDIT_READ_GAP) {
            num_go++; num_gx++;
            while (i < (size_t)rs->nedits - 1 && rs->edits[i + 1].pos == rs->edits[i].pos && rs->edits[i + 1].type == HT2_EDIT_READ_GAP) { i++; num_gx++; NM++; }
        } else if (e.type == HT2_EDIT_REF_GAP) {
            num_go++; num_gx++;
            while (i < (size_t)rs->nedits - 1 && rs->edits[i + 1].pos == rs->edits[i].pos + 1 && rs->edits[i + 1].type == HT2_EDIT_REF_GAP) { i++; num_gx++; NM++; }
        }
    }
    o += "\tXM:i:"; o += std::to_string(num_mm);
    o += "\tXO:i:"; o += std::to_string(num_go);
    o += "\tXG:i:"; o += std::to_string(num_gx);
    o += "\tNM:i:"; o += std::to_string(NM);
    o += "\tMD:Z:"; st.mdz(o);
    o += "\tYT:Z:UU";
    appendYF(o, f);
    o += "\tNH:i:"; o += std::to_string(summ.numAlns[0]);
    o.push_back('\n');
}

void ht2_finish_unpaired(std::string& o, const Ht2Image& img, const Ht2Params& P,
                         const Ht2HostRead& rd, const Ht2ReadFilters& f, Ht2ReadOut& out)
{
    Ht2Rng rnd; rnd.last = out.rngLast;
    const std::vector<Ht2Res>& rs = out.res[0];
    uint64_t nunpair1 = std::min<uint64_t>(rs.size(), P.khits); // ReportingState::getReport
    Summ summ; summ.reset();
    if (nunpair1 > 0) {
        summ.addUnp(0, rs);
        std::vector<size_t> select;
        selectByScore(rs, NULL, NULL, nunpair1, select, rnd, P.secondary != 0);
        summ.numAlns[0] = select.size();
        for (size_t i = 0; i < select.size(); i++) {
            appendMateUnpaired(o, img, rd, f, &rs[select[i]], summ, i == 0);
        }
    } else {
        appendMateUnpaired(o, img, rd, f, NULL, summ, true);
    }
    out.rngLast = rnd.last;
}
