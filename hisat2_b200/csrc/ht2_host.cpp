// ht2_host.cpp -- see ht2_host.h.
#include "ht2_host.h"

#include <stdio.h>
#include <string.h>
#include <math.h>

#include <algorithm>
#include <thread>

// decimal append without a temporary std::string
static inline void appendInt(std::string& o, int64_t v) {
    char buf[24]; int n = 0;
    uint64_t u = v < 0 ? (uint64_t)0 - (uint64_t)v : (uint64_t)v;
    do { buf[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) o.push_back('-');
    while (n) o.push_back(buf[--n]);
}

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
    P.bowtie2Dp = 0; P.gapbar = 4;                      // hisat2.cpp:529, 419
#ifdef HT2_ENABLE_SPLICED
    P.noncanSplPen = 12;                                // hisat2.cpp:494
#endif
    ht2_set_score_min(P, 'L', (double)0.0f, (double)-0.2f);   // hisat2.cpp:441
}

// SimpleFunc::f<TAlScore> (simple_func.h:86-108) with min = -DBL_MAX, max = DBL_MAX
// (SimpleFunc::init(type, C, L)), tabulated for every supported read length.
bool ht2_set_score_min(Ht2Params& P, char type, double C, double L)
{
    if (type != 'C' && type != 'L' && type != 'S' && type != 'G') return false;
    for (uint32_t len = 0; len <= HT2_PARAMS_MAX_RDLEN; len++) {
        double x = (double)len, X;
        if (type == 'C') X = 0.0; else if (type == 'L') X = x; else if (type == 'S') X = sqrt(x); else X = log(x);
        double v = std::max(-1.7976931348623157e308, std::min(1.7976931348623157e308, C + L * X));
        if (!(v == v)) v = 0.0;
        if (v > 2147483647.0) v = 2147483647.0;
        if (v < -2147483648.0) v = -2147483648.0;
        P.minscTab[len] = (int32_t)(int64_t)v;
    }
    return true;
}

// SimpleFunc::f<T> for the linear functions the defaults use (simple_func.h:86-108)
static double simpleLinear(double I, double X, double C, double L, double x) {
    return std::max(I, std::min(X, C + L * x));
}

int64_t ht2_minsc(const Ht2Params& P, uint32_t rdlen)
{
    // scoreMin.f(rdlen) (default L,0,-0.2, hisat2.cpp:441); clamped to <= 0 in end-to-end mode (:3395-3402)
    int64_t m = P.minscTab[rdlen <= HT2_PARAMS_MAX_RDLEN ? rdlen : HT2_PARAMS_MAX_RDLEN];
    if (m > 0) m = 0;
    return m;
}

Ht2ReadFilters ht2_filters(const Ht2HostRead& rd, int64_t minsc)
{
    Ht2ReadFilters f;
    size_t rdlen = rd.seq.size();
    // Scoring::nFilter (scoring.cpp:104-117).  The effective default is nCeil = L,0,0.15: hisat2.cpp:443
    // initialises 2 + 0.1 len, but SeedAlignmentPolicy::parseString -- always called, hisat2.cpp:1849 --
    // resets it to DEFAULT_N_CEIL_CONST/LINEAR = 0.0 / 0.15 (aligner_seed_policy.cpp:293-296, scoring.h:65-67).
    size_t maxns = (size_t)simpleLinear(0.0, 1.7976931348623157e308, (double)0.0f, (double)0.15f, (double)rdlen);
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
        // an empty record stays in the stream: the reference reports it as unaligned with YF:Z:LN
        out.push_back(r);
    }
    return true;
}

// FastqPatternSource::parse (pat.cpp:1030-1290) for well-formed 4-line Phred+33 records:
// '@name', bases ('.' reads as N, letters only), '+...', qualities kept as raw ASCII.
bool ht2_read_fastq(const char* path, std::vector<Ht2HostRead>& out, int mate, std::string& err)
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
    auto line = [&](std::string& o) {
        o.clear();
        while (p < n && buf[p] != '\n') { if (buf[p] != '\r') o.push_back(buf[p]); p++; }
        if (p < n) p++;
    };
    std::string l1, l2, l3, l4;
    while (p < n) {
        while (p < n && (buf[p] == '\n' || buf[p] == '\r')) p++;
        if (p >= n) break;
        line(l1); line(l2); line(l3); line(l4);
        if (l1.empty() || l1[0] != '@' || l3.empty() || l3[0] != '+') { err = "reads file does not look like a FASTQ file"; return false; }
        Ht2HostRead r;
        r.mate = mate;
        r.name = l1.substr(1);
        for (size_t i = 0; i < l2.size(); i++) {
            int c = (unsigned char)l2[i];
            if (c == '.') c = 'N';
            if (isalpha(c)) r.seq.push_back(asc2dna(c));
        }
        if (l4.size() < r.seq.size()) { err = "fewer quality values than bases for read " + r.name; return false; }
        if (l4.size() > r.seq.size()) { err = "more quality values than bases for read " + r.name; return false; }
        for (size_t i = 0; i < r.seq.size(); i++) {
            if ((unsigned char)l4[i] < 33) { err = "quality value below Phred+33 range in read " + r.name; return false; }
            r.qual.push_back((uint8_t)l4[i]);
        }
        if (r.name.empty()) r.name = std::to_string(readCnt);
        readCnt++;
        out.push_back(r);
    }
    return true;
}

// FASTA or FASTQ by the first record character (the CLI's -f / -q select explicitly).
bool ht2_read_reads(const char* path, std::vector<Ht2HostRead>& out, int mate, std::string& err)
{
    FILE* f = fopen(path, "rb");
    if (!f) { err = std::string("could not open reads file ") + path; return false; }
    int c;
    while ((c = fgetc(f)) != EOF && (c == '\n' || c == '\r')) {}
    fclose(f);
    return c == '@' ? ht2_read_fastq(path, out, mate, err) : ht2_read_fasta(path, out, mate, err);
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
        appendInt(o, img.refPlen(i));
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
    int64_t splicescore = 255, transcript_score = 0;
#ifdef HT2_ENABLE_SPLICED
    splicescore = (int64_t)(r.splicescore / 100);     // TAlScore splicescore = splicescore_ / 100 (aligner_result.h:339)
    if (splicescore > 255) splicescore = 0; else splicescore = 255 - splicescore;
    transcript_score = r.knownTranscripts ? 2 : (r.spliced ? 1 : 0);
#endif
    int64_t trim = (int64_t)r.trim5p + (int64_t)r.trim3p; // leftTrim+rightTrim (hit.trim5+hit.trim3)
    if (trim > 65535) trim = 0; else trim = 65535 - trim;
    return (int64_t)((uint64_t)score << 32) | (transcript_score << 24) | (splicescore << 16) | trim;
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
int mapqV2(const Ht2Params& P, const Summ& s, bool mate1, size_t rdlen, size_t ordlen)
{
    const ScoreKey& bst = s.paired ? s.bestPaired : s.best[mate1 ? 0 : 1];
    const ScoreKey& sec = s.paired ? s.secbestPaired : s.secbest[mate1 ? 0 : 1];
    bool hasSecbest = sec.valid;
    bool equalSecbest = hasSecbest && keyEq(bst, sec);
    if (!hasSecbest || !equalSecbest) return 60;
    int64_t scPer = 0;
    // scoreMin_.f<TAlScore>((float)rdlen) (unique.h:200-203), from the per-length table
    int64_t scMin = P.minscTab[rdlen <= HT2_PARAMS_MAX_RDLEN ? rdlen : HT2_PARAMS_MAX_RDLEN];
    if (s.paired) scMin += P.minscTab[ordlen <= HT2_PARAMS_MAX_RDLEN ? ordlen : HT2_PARAMS_MAX_RDLEN];
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
    std::vector<uint32_t> skip;   // intron lengths of the 'N' columns, in order
    size_t trimLS, trimRS;
    void init(const std::vector<uint8_t>& s, const Ht2Edit* ed, size_t ned, size_t tLS, size_t tRS) {
        ref.clear(); rel.clear(); read.clear(); snp.clear(); skip.clear();
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
#ifdef HT2_ENABLE_SPLICED
            else if (ed[i].type == HT2_EDIT_SPL) {   // aligner_result.cpp:711-718
                ref.push_back('N'); rel.push_back('N'); snp.push_back(false); read.push_back('N'); skip.push_back(ht2_spl_len(ed[i]));
            }
#endif
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
        if (trimLS > 0) { appendInt(o, trimLS); o.push_back('S'); }
        size_t ln = ref.size(), numSkips = 0;
        for (size_t i = 0; i < ln; i++) {
            char op = rel[i];
            if (op == 'X' || op == '=') op = 'M';
            size_t run = 1;
            if (op != 'N') {
                for (; i + run < ln; run++) {
                    char op2 = rel[i + run];
                    if (op2 == 'X' || op2 == '=') op2 = 'M';
                    if (op2 != op) break;
                }
                i += (run - 1);
            } else run = skip[numSkips++];   // aligner_result.cpp:815-833
            appendInt(o, run); o.push_back(op);
        }
        if (trimRS > 0) { appendInt(o, trimRS); o.push_back('S'); }
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
                    else if (rel[i + run] == 'I' || rel[i + run] == 'N') ninserts++;   // insertions and introns do not count (aligner_result.cpp:862-879)
                    else break;
                }
                i += (run - 1);
                size_t r = run - ninserts;
                if (r > 0) { appendInt(o, r); first_print = false; mm_last = false; rdgap_last = false; }
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
    if (n == 0) { o += "*\t*"; return; }   // aln_sink.h:3194, 3210
    const size_t at = o.size();
    o.resize(at + 2 * n + 1);
    char* d = &o[at];
    const uint8_t* sq = rd.seq.data(); const uint8_t* ql = rd.qual.data();
    if (fw) { for (size_t i = 0; i < n; i++) d[i] = "ACGTN"[sq[i]]; }
    else for (size_t i = 0; i < n; i++) { uint8_t c = sq[n - i - 1]; d[i] = "TGCAN"[c < 4 ? c : 4]; }
    d[n] = '\t';
    d += n + 1;
    if (fw) memcpy(d, ql, n);
    else for (size_t i = 0; i < n; i++) d[i] = (char)ql[n - i - 1];
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

// AlnFlags::pairing (aligner_result.h:383-398)
enum { PAIR_CONCORD_MATE1 = 1, PAIR_CONCORD_MATE2, PAIR_DISCORD_MATE1, PAIR_DISCORD_MATE2,
       PAIR_UNPAIRED_MATE1, PAIR_UNPAIRED_MATE2, PAIR_UNPAIRED };

struct MateFlags {
    int pairing; bool primary; bool oppAligned;
    bool partOfPair() const { return pairing < PAIR_UNPAIRED; }
    bool readMate1() const { return pairing == PAIR_CONCORD_MATE1 || pairing == PAIR_DISCORD_MATE1 || pairing == PAIR_UNPAIRED_MATE1; }
    bool concordant() const { return pairing == PAIR_CONCORD_MATE1 || pairing == PAIR_CONCORD_MATE2; }
    bool discordant() const { return pairing == PAIR_DISCORD_MATE1 || pairing == PAIR_DISCORD_MATE2; }
    bool unpairedMate() const { return pairing == PAIR_UNPAIRED_MATE1 || pairing == PAIR_UNPAIRED_MATE2; }
};

// AlnRes::setFragmentLength (aligner_result.h:1631-1694) without splice sites
static int64_t fragmentLength(const Ht2Res& me, const Ht2Res& o, bool meMate1)
{
    // AlnRes::setFragmentLength (aligner_result.h:1631-1697) with an empty splice-site DB; st2/en2 are the
    // extents shifted right by the alignment's own introns (getCoords, :1132-1147)
    auto ext = [](const Ht2Res& r, int64_t& st, int64_t& en, int64_t& st2, int64_t& en2) {
        int64_t trim_st = r.fw ? r.trim5p : r.trim3p, trim_en = r.fw ? r.trim3p : r.trim5p;
        int64_t introns = 0;
#ifdef HT2_ENABLE_SPLICED
        for (uint32_t e = 0; e < r.nedits; e++) if (r.edits[e].type == HT2_EDIT_SPL) introns += ht2_spl_len(r.edits[e]);
#endif
        st = (int64_t)r.toff - trim_st;
        en = (int64_t)r.toff + r.rfextent - 1 + trim_en;
        st2 = st + introns; en2 = en + introns;
    };
    int64_t st, en, st2, en2, ost, oen, ost2, oen2;
    ext(me, st, en, st2, en2); ext(o, ost, oen, ost2, oen2);
    bool imUpstream;
    if (st < ost) imUpstream = true;
    else if (st == ost) {
        if (me.fw && o.fw && meMate1) imUpstream = true;
        else if (me.fw && !o.fw) imUpstream = true;
        else imUpstream = false;
    } else imUpstream = false;
    int64_t up, dn;
    if (imUpstream) { up = std::min(st2, ost); dn = std::max(en2, oen); }
    else { up = std::min(st, ost2); dn = std::max(en, oen2); }
    int64_t fraglen = 1 + dn - up;
    if (!imUpstream) fraglen = -fraglen;
    return fraglen;
}

// AlnSinkSam::appendMate (aln_sink.h:3024-3250)
static void appendMate(std::string& o, const Ht2Image& img, const Ht2Params& P, const Ht2HostRead& rd, size_t ordlen, const Ht2ReadFilters& f,
                       const Ht2Res* rs, const Ht2Res* rso, const Summ& summ, const MateFlags& fl,
                       bool fraglenSet, int64_t fraglen, bool haveOscore)
{
    appendName(o, rd.name, fl.partOfPair());
    o.push_back('\t');
    int flag = 0;
    if (fl.partOfPair()) {
        flag |= 1;
        if (fl.concordant()) flag |= 2;
        if (!fl.oppAligned) flag |= 8;
        flag |= fl.readMate1() ? 64 : 128;
        if (fl.oppAligned && rso != NULL && !rso->fw) flag |= 32;
    }
    if (!fl.primary) flag |= 256;
    if (rs != NULL && !rs->fw) flag |= 16;
    if (rs == NULL) flag |= 4;
    appendInt(o, flag);
    o.push_back('\t');
    const char* ytz = fl.concordant() ? "CP" : fl.discordant() ? "DP" : fl.unpairedMate() ? "UP" : "UU";
    if (rs == NULL) {
        if (summ.orefid != -1) { appendRefName(o, img, (uint32_t)summ.orefid); o.push_back('\t'); appendInt(o, summ.orefoff + 1); o += "\t0\t*\t=\t"; appendInt(o, summ.orefoff + 1); o += "\t0\t"; }
        else o += "*\t0\t0\t*\t*\t0\t0\t";
        appendSeqQual(o, rd, true);
        o += "\tYT:Z:"; o += ytz;
        appendYF(o, f);
        o.push_back('\n');
        return;
    }
    // Gapless alignments (mismatches only; the vast majority): CIGAR and MD:Z follow from the edit
    // positions directly -- what StackedAln::buildCigar / buildMdz (aligner_result.cpp:793-1000) print for a
    // stack without I/D columns, with nothing for leftAlign to move.  Everything else goes through the
    // stacked form.
    bool gapless = true;
    for (uint32_t i = 0; i < rs->nedits; i++) if (rs->edits[i].type != HT2_EDIT_MM) { gapless = false; break; }
    static thread_local Stacked st;
    static thread_local std::vector<Ht2Edit> nedBuf;
    static thread_local std::vector<uint8_t> seqBuf;
    size_t trimLS = rs->trim5p, trimRS = rs->trim3p;
    const size_t len_trimmed = rd.seq.size() - trimLS - trimRS;
    if (!rs->fw) std::swap(trimLS, trimRS);
    if (!gapless) {
        // AlnRes::initStacked (aligner_result.h:1856-1873)
        nedBuf.assign(rs->edits, rs->edits + rs->nedits);
        seqBuf.assign(rd.seq.begin(), rd.seq.end());
        if (!rs->fw) {
            invertPossHost(nedBuf, len_trimmed);
            size_t n = seqBuf.size();
            for (size_t i = 0; i < n; i++) { uint8_t c = rd.seq[n - i - 1]; seqBuf[i] = c < 4 ? (uint8_t)(c ^ 3) : (uint8_t)4; }
        }
        st.init(seqBuf, nedBuf.data(), nedBuf.size(), trimLS, trimRS);
        st.leftAlign(false);
    }
    appendRefName(o, img, rs->tidx);
    o.push_back('\t');
    appendInt(o, (int64_t)rs->toff + 1);
    o.push_back('\t');
    appendInt(o, mapqV2(P, summ, rd.mate < 2, rd.seq.size(), ordlen));
    o.push_back('\t');
    if (gapless) {
        if (trimLS > 0) { appendInt(o, trimLS); o.push_back('S'); }
        if (len_trimmed > 0) { appendInt(o, len_trimmed); o.push_back('M'); }
        if (trimRS > 0) { appendInt(o, trimRS); o.push_back('S'); }
    } else st.cigar(o);
    o.push_back('\t');
    if (fl.partOfPair()) {
        if (rso != NULL && rs->tidx != rso->tidx) { appendRefName(o, img, rso->tidx); o.push_back('\t'); }
        else o += "=\t";
        appendInt(o, (int64_t)(rso ? rso->toff : rs->toff) + 1);
        o.push_back('\t');
    } else o += "*\t0\t";
    appendInt(o, fraglenSet ? fraglen : 0);
    o.push_back('\t');
    appendSeqQual(o, rd, rs->fw != 0);
    // optional flags (sam.h:525-1010)
    o += "\tAS:i:"; appendInt(o, rs->score);
    const ScoreKey& sb = summ.secbest[rd.mate < 2 ? 0 : 1];
    if (sb.valid) { o += "\tZS:i:"; appendInt(o, sb.score); }
    o += "\tXN:i:0";
    // counts exclude edits that are known ALTs (snpID < #alts, sam.h:574-647)
    const Ht2ImageHeader* IH = img.header();
    const uint32_t nAlts = IH->nAlts;
    const Ht2Alt* altTab = (const Ht2Alt*)(img.blob.data() + IH->o_alts);
    size_t num_mm = 0, num_go = 0, num_gx = 0, NM = 0;
    for (size_t i = 0; i < rs->nedits; i++) if (rs->edits[i].type != HT2_EDIT_SPL && rs->edits[i].snpID >= nAlts) NM++;
    for (size_t i = 0; i < rs->nedits; i++) {
        const Ht2Edit& e = rs->edits[i];
        if (e.type == HT2_EDIT_MM) { if (e.snpID >= nAlts) num_mm++; }
        else if (e.type == HT2_EDIT_READ_GAP) {
            if (e.snpID >= nAlts) { num_go++; num_gx++; }
            while (i < (size_t)rs->nedits - 1 && rs->edits[i + 1].pos == rs->edits[i].pos && rs->edits[i + 1].type == HT2_EDIT_READ_GAP) {
                i++; if (rs->edits[i].snpID >= nAlts) num_gx++;
            }
        } else if (e.type == HT2_EDIT_REF_GAP) {
            if (e.snpID >= nAlts) { num_go++; num_gx++; }
            while (i < (size_t)rs->nedits - 1 && rs->edits[i + 1].pos == rs->edits[i].pos + 1 && rs->edits[i + 1].type == HT2_EDIT_REF_GAP) {
                i++; if (rs->edits[i].snpID >= nAlts) num_gx++;
            }
        }
    }
    o += "\tXM:i:"; appendInt(o, num_mm);
    o += "\tXO:i:"; appendInt(o, num_go);
    o += "\tXG:i:"; appendInt(o, num_gx);
    o += "\tNM:i:"; appendInt(o, NM);
    o += "\tMD:Z:";
    if (gapless) {   // <matches>[<ref char><matches>]..., a 0 between adjacent mismatches and at either end
        size_t prevEnd = 0;
        for (uint32_t k = 0; k < rs->nedits; k++) {
            const Ht2Edit& e = rs->fw ? rs->edits[k] : rs->edits[rs->nedits - 1 - k];
            const size_t p = rs->fw ? (size_t)e.pos : len_trimmed - (size_t)e.pos - 1;
            appendInt(o, (int64_t)(p - prevEnd));
            o.push_back((char)e.chr);
            prevEnd = p + 1;
        }
        appendInt(o, (int64_t)(len_trimmed - prevEnd));
    } else st.mdz(o);
    if (summ.paired && haveOscore && rso) { o += "\tYS:i:"; appendInt(o, rso->score); }
    o += "\tYT:Z:"; o += ytz;
    appendYF(o, f);
#ifdef HT2_ENABLE_SPLICED
    {   // XS:A: AlnRes::spliced_whichsense_transcript (aligner_result.h:1288-1318, sam.h:925-937)
        uint8_t whichsense = HT2_SPL_UNKNOWN; bool any = false;
        for (uint32_t i = 0; i < rs->nedits; i++) {
            const Ht2Edit& e = rs->edits[i];
            if (e.type != HT2_EDIT_SPL) continue;
            any = true;
            const uint8_t d = (uint8_t)ht2_spl_dir(e);
            if (whichsense == HT2_SPL_UNKNOWN) whichsense = d;
            else if (d != HT2_SPL_UNKNOWN) {
                if ((whichsense == HT2_SPL_FW || whichsense == HT2_SPL_SEMI_FW) && d != HT2_SPL_FW && d != HT2_SPL_SEMI_FW) { whichsense = HT2_SPL_UNKNOWN; break; }
                if ((whichsense == HT2_SPL_RC || whichsense == HT2_SPL_SEMI_RC) && d != HT2_SPL_RC && d != HT2_SPL_SEMI_RC) { whichsense = HT2_SPL_UNKNOWN; break; }
            }
        }
        if (any && whichsense != HT2_SPL_UNKNOWN) { o += "\tXS:A:"; o.push_back((whichsense == HT2_SPL_FW || whichsense == HT2_SPL_SEMI_FW) ? '+' : '-'); }
    }
#endif
    if (fl.concordant() || fl.discordant()) { o += "\tNH:i:"; appendInt(o, summ.numAlnsPaired); }
    else { o += "\tNH:i:"; appendInt(o, (fl.pairing == PAIR_UNPAIRED || fl.readMate1()) ? summ.numAlns[0] : summ.numAlns[1]); }
    // Zs:Z: the known ALTs the alignment went through (sam.h:983-1032)
    if (nAlts > 0) {
        std::vector<Ht2Edit> ned(rs->edits, rs->edits + rs->nedits);
        const size_t len_trimmed = rd.seq.size() - rs->trim5p - rs->trim3p;
        if (!rs->fw) {   // Edit::invertPoss (edit.cpp:70-111), sort = false
            std::reverse(ned.begin(), ned.end());
            for (Ht2Edit& e : ned) e.pos = (e.type == HT2_EDIT_READ_GAP || e.type == HT2_EDIT_SPL) ? (uint32_t)(len_trimmed - e.pos) : (uint32_t)(len_trimmed - e.pos - 1);
        }
        bool first = true;
        uint32_t prev = 0xffffffffu;
        const char* names = (const char*)img.blob.data() + IH->o_altNames;
        for (size_t i = 0; i < ned.size(); i++) {
            if (ned[i].type == HT2_EDIT_SPL || ned[i].snpID >= nAlts) continue;   // a splice edit keeps its site probability in the snpID word
            const uint32_t si = ned[i].snpID;
            const Ht2Alt& snp = altTab[si];
            if (si == prev) continue;
            o += first ? "\tZs:Z:" : ",";
            uint64_t pos = ned[i].pos;
            size_t j = i;
            while (j > 0) {
                if (ned[j - 1].snpID < nAlts) {
                    const Ht2Alt& snp2 = altTab[ned[j - 1].snpID];
                    if (snp2.type == HT2_ALT_SNP_SGL) pos -= (ned[j - 1].pos + 1);
                    else if (snp2.type == HT2_ALT_SNP_DEL) pos -= ned[j - 1].pos;
                    else if (snp2.type == HT2_ALT_SNP_INS) pos -= (ned[j - 1].pos + snp.len);
                    break;
                }
                j--;
            }
            appendInt(o, pos);
            o += (snp.type == HT2_ALT_SNP_SGL) ? "|S|" : (snp.type == HT2_ALT_SNP_DEL ? "|D|" : "|I|");
            const char* nm = names;
            for (uint32_t k = 0; k < si; k++) nm += strlen(nm) + 1;
            o += nm;
            first = false;
            prev = si;
        }
    }
    o.push_back('\n');
}

void ht2_finish_unpaired(std::string& o, const Ht2Image& img, const Ht2Params& P,
                         const Ht2HostRead& rd, const Ht2ReadFilters& f, Ht2ReadOut& out)
{
    Ht2Rng rnd; rnd.last = out.rngLast;
    const std::vector<Ht2Res>& rs = out.res[0];
    uint64_t nunpair1 = std::min<uint64_t>(rs.size(), P.khits); // ReportingState::getReport
    Summ summ; summ.reset();
    MateFlags fl = {PAIR_UNPAIRED, true, false};
    if (nunpair1 > 0) {
        summ.addUnp(0, rs);
        std::vector<size_t> select;
        selectByScore(rs, NULL, NULL, nunpair1, select, rnd, P.secondary != 0);
        summ.numAlns[0] = select.size();
        for (size_t i = 0; i < select.size(); i++) {
            fl.primary = (i == 0);
            appendMate(o, img, P, rd, 0, f, &rs[select[i]], NULL, summ, fl, false, 0, false);
        }
    } else {
        appendMate(o, img, P, rd, 0, f, NULL, NULL, summ, fl, false, 0, false);
    }
    out.rngLast = rnd.last;
}

// AlnSinkWrap::finishRead for a pair (aln_sink.h:1939-2557)
void ht2_finish_paired(std::string& o, const Ht2Image& img, const Ht2Params& P,
                       const Ht2HostRead& rd1, const Ht2HostRead& rd2,
                       const Ht2ReadFilters& f1, const Ht2ReadFilters& f2, Ht2ReadOut& out)
{
    Ht2Rng rnd; rnd.last = out.rngLast;
    const std::vector<Ht2Res>& rs1u = out.res[0];
    const std::vector<Ht2Res>& rs2u = out.res[1];
    // ReportingState replay (aln_sink.cpp:72-131, 139-170)
    uint64_t nconcord_ = 0;
    {
        int64_t best = HT2_MIN_SCORE;
        for (size_t i = 0; i < out.pairs.size(); i++) {
            int64_t sc = rs1u[out.pairs[i].first].score + rs2u[out.pairs[i].second].score;
            if (sc > best) { best = sc; nconcord_ = 0; }
            nconcord_++;
        }
    }
    uint64_t nunpair1_ = rs1u.size(), nunpair2_ = rs2u.size();
    bool discordant = P.discord && out.pairs.empty() && nunpair1_ == 1 && nunpair2_ == 1;
    if (nconcord_ > 0) {
        uint64_t nconcord = std::min<uint64_t>(P.khits, nconcord_);
        Summ summ; summ.reset();
        summ.paired = true;
        for (size_t i = 0; i < out.pairs.size(); i++) {
            const Ht2Res& a = rs1u[out.pairs[i].first]; const Ht2Res& b = rs2u[out.pairs[i].second];
            ScoreKey sc = {a.score + b.score, hisat2Score(a) + hisat2Score(b), true};
            if (keyGt(sc, summ.bestPaired)) { summ.secbestPaired = summ.bestPaired; summ.bestPaired = sc; }
            else if (keyGt(sc, summ.secbestPaired)) summ.secbestPaired = sc;
        }
        summ.addUnp(0, rs1u); summ.addUnp(1, rs2u);
        std::vector<size_t> select;
        selectByScore(rs1u, &rs2u, &out.pairs, nconcord, select, rnd, P.secondary != 0);
        summ.numAlnsPaired = select.size();
        MateFlags fl1 = {PAIR_CONCORD_MATE1, true, true}, fl2 = {PAIR_CONCORD_MATE2, true, true};
        for (size_t i = 0; i < select.size(); i++) {
            const Ht2Res& a = rs1u[out.pairs[select[i]].first]; const Ht2Res& b = rs2u[out.pairs[select[i]].second];
            fl1.primary = fl2.primary = (i == 0);
            appendMate(o, img, P, rd1, rd2.seq.size(), f1, &a, &b, summ, fl1, true, fragmentLength(a, b, true), true);
            appendMate(o, img, P, rd2, rd1.seq.size(), f2, &b, &a, summ, fl2, true, fragmentLength(b, a, false), true);
        }
        out.rngLast = rnd.last;
        return;
    } else if (discordant) {
        Summ summ; summ.reset();
        summ.paired = true;
        {
            const Ht2Res& a = rs1u[0]; const Ht2Res& b = rs2u[0];
            ScoreKey sc = {a.score + b.score, hisat2Score(a) + hisat2Score(b), true};
            summ.bestPaired = sc;
        }
        summ.addUnp(0, rs1u); summ.addUnp(1, rs2u);
        summ.numAlnsPaired = 1; // AlnSetSumm::init counts rs1->size()
        std::vector<std::pair<uint16_t, uint16_t> > dp(1, std::make_pair((uint16_t)0, (uint16_t)0));
        std::vector<size_t> select;
        selectByScore(rs1u, &rs2u, &dp, 1, select, rnd, P.secondary != 0);
        MateFlags fl1 = {PAIR_DISCORD_MATE1, true, true}, fl2 = {PAIR_DISCORD_MATE2, true, true};
        const Ht2Res& a = rs1u[0]; const Ht2Res& b = rs2u[0];
        bool sameRef = a.tidx == b.tidx; // setMateParams (aligner_result.h:1594-1618)
        appendMate(o, img, P, rd1, rd2.seq.size(), f1, &a, &b, summ, fl1, sameRef, sameRef ? fragmentLength(a, b, true) : 0, true);
        appendMate(o, img, P, rd2, rd1.seq.size(), f2, &b, &a, summ, fl2, sameRef, sameRef ? fragmentLength(b, a, false) : 0, true);
        out.rngLast = rnd.last;
        return;
    }
    uint64_t nunpair1 = 0, nunpair2 = 0;
    if (P.mixed && nunpair1_ + nunpair2_ > 0) {
        nunpair1 = std::min<uint64_t>(nunpair1_, P.khits);
        nunpair2 = std::min<uint64_t>(nunpair2_, P.khits);
    }
    bool rep1 = nunpair1 > 0, rep2 = nunpair2 > 0;
    Summ summ1, summ2; summ1.reset(); summ2.reset();
    std::vector<size_t> select1, select2;
    const Ht2Res *repRs1 = NULL, *repRs2 = NULL;
    if (rep1) {
        summ1.addUnp(0, rs1u);
        if (rep2) summ1.addUnp(1, rs2u);
        selectByScore(rs1u, NULL, NULL, nunpair1, select1, rnd, P.secondary != 0);
        repRs1 = &rs1u[select1[0]];
    }
    if (rep2) {
        summ2.addUnp(1, rs2u);
        if (rep1) summ2.addUnp(0, rs1u);
        selectByScore(rs2u, NULL, NULL, nunpair2, select2, rnd, P.secondary != 0);
        repRs2 = &rs2u[select2[0]];
    }
    // numAlns1/2 setters are applied to both summaries (aln_sink.h:2238-2239, 2263-2264)
    if (rep1) { summ1.numAlns[0] = select1.size(); summ2.numAlns[0] = select1.size(); }
    if (rep2) { summ1.numAlns[1] = select2.size(); summ2.numAlns[1] = select2.size(); }
    MateFlags fl1 = {PAIR_UNPAIRED_MATE1, true, repRs2 != NULL}, fl2 = {PAIR_UNPAIRED_MATE2, true, repRs1 != NULL};
    int64_t refid = -1, refoff = -1;
    if (rep1) {
        // AlnSink::reportHits (aln_sink.h:730-790)
        if (repRs2 != NULL) {
            const Ht2Res* r1pri = &rs1u[select1[0]]; const Ht2Res* r2pri = &rs2u[select2[0]];
            appendMate(o, img, P, rd1, rd2.seq.size(), f1, r1pri, r2pri, summ1, fl1, false, 0, false);
            appendMate(o, img, P, rd2, rd1.seq.size(), f2, r2pri, r1pri, summ1, fl2, false, 0, false);
            fl1.primary = fl2.primary = false;
            for (size_t i = 1; i < select1.size(); i++)
                appendMate(o, img, P, rd1, rd2.seq.size(), f1, &rs1u[select1[i]], r2pri, summ1, fl1, false, 0, false);
            for (size_t i = 1; i < select2.size(); i++)
                appendMate(o, img, P, rd2, rd1.seq.size(), f2, &rs2u[select2[i]], r1pri, summ1, fl2, false, 0, false);
            fl1.primary = fl2.primary = true;
        } else {
            for (size_t i = 0; i < select1.size(); i++) {
                fl1.primary = (i == 0);
                appendMate(o, img, P, rd1, 0, f1, &rs1u[select1[i]], NULL, summ1, fl1, false, 0, false);
            }
            fl1.primary = true;
        }
        refid = rs1u[select1[0]].tidx; refoff = rs1u[select1[0]].toff;
    }
    if (rep2 && !rep1) {
        for (size_t i = 0; i < select2.size(); i++) {
            fl2.primary = (i == 0);
            appendMate(o, img, P, rd2, 0, f2, &rs2u[select2[i]], NULL, summ2, fl2, false, 0, false);
        }
        fl2.primary = true;
        refid = rs2u[select2[0]].tidx; refoff = rs2u[select2[0]].toff;
    }
    if (nunpair1 == 0) {
        Summ s; s.reset();
        if (nunpair2 > 0) { s.orefid = refid; s.orefoff = refoff; }
        MateFlags fl = {PAIR_UNPAIRED_MATE1, true, repRs2 != NULL};
        appendMate(o, img, P, rd1, 0, f1, NULL, NULL, s, fl, false, 0, false);
    }
    if (nunpair2 == 0) {
        Summ s; s.reset();
        if (nunpair1 > 0) { s.orefid = refid; s.orefoff = refoff; }
        MateFlags fl = {PAIR_UNPAIRED_MATE2, true, repRs1 != NULL};
        appendMate(o, img, P, rd2, 0, f2, NULL, NULL, s, fl, false, 0, false);
    }
    out.rngLast = rnd.last;
}

// The SAM back end of one batch (ht2gpu_format_sam): structured results -> SAM records, on host
// threads over contiguous unit ranges, chunks placed in read order (= --reorder).
bool ht2_format_batch(const Ht2Image& img, const Ht2Params& P, const ht2gpu_read_batch_t* b, const char* names,
                      const ht2gpu_result_batch_t* res, char** out, size_t* out_len, unsigned nth)
{
    const uint32_t units = b->paired ? b->n_reads / 2 : b->n_reads;
    // read-name table (names are '\0'-terminated, concatenated)
    std::vector<const char*> nameOf((size_t)b->n_reads + 1);
    {
        const char* nm = names;
        for (uint32_t i = 0; i < b->n_reads; i++) { nameOf[i] = nm; nm += strlen(nm) + 1; }
    }
    auto mkRead = [&](uint32_t i, int mate, Ht2HostRead& rd) {
        rd.name = nameOf[i];
        rd.mate = mate;
        const uint8_t* s = b->seq + b->offs[i];
        uint32_t len = (uint32_t)(b->offs[i + 1] - b->offs[i]);
        rd.seq.assign(s, s + len);
        if (b->qual) rd.qual.assign(b->qual + b->offs[i], b->qual + b->offs[i] + len); else rd.qual.assign(len, (uint8_t)'I');
    };
    // Reads are independent (each carries its own RNG state), so the back end runs on
    // host threads over contiguous unit ranges and the chunks are concatenated in order.
    auto doRange = [&](uint32_t u0, uint32_t u1, std::string& sam) {
      sam.reserve((size_t)(u1 - u0) * (b->paired ? 800 : 400));
      Ht2HostRead rd1, rd2;      // reused across the range: no per-read heap traffic once the buffers have grown
      Ht2ReadOut o;
      for (uint32_t u = u0; u < u1; u++) {
        if (b->paired) { mkRead(2 * u, 1, rd1); mkRead(2 * u + 1, 2, rd2); }
        else mkRead(u, 0, rd1);
        Ht2ReadFilters f1 = ht2_filters(rd1, ht2_minsc(P, (uint32_t)rd1.seq.size()));
        Ht2ReadFilters f2 = f1;
        if (b->paired) f2 = ht2_filters(rd2, ht2_minsc(P, (uint32_t)rd2.seq.size()));
        const ht2gpu_read_result_t& rr = res->reads[u];
        o.rngLast = rr.rng_state; o.err = rr.err;
        o.pairs.clear();
        uint32_t a = rr.aln_off;
        for (uint32_t m = 0; m < 2; m++) {
            o.res[m].resize(rr.n_aln[m]);
            for (uint32_t k = 0; k < rr.n_aln[m]; k++, a++) {
                const ht2gpu_aln_t& al = res->alns[a];
                Ht2Res& r = o.res[m][k];
                r.tidx = al.tidx; r.toff = al.toff; r.fw = al.fw; r.score = al.score;
                r.rdlen = (uint32_t)((m == 0 || !b->paired) ? rd1.seq.size() : rd2.seq.size());
                r.trim5p = al.trim5; r.trim3p = al.trim3; r.rfextent = al.ref_extent; r.nedits = al.n_edits;
                for (uint32_t e = 0; e < al.n_edits && e < HT2_MAX_EDITS; e++) {
                    const ht2gpu_edit_t& se = res->edits[al.edit_off + e];
                    r.edits[e].pos = se.pos; r.edits[e].chr = se.chr; r.edits[e].qchr = se.qchr; r.edits[e].type = se.type;
                    r.edits[e].pad = se.pad; r.edits[e].snpID = se.snp_id;
                }
#ifdef HT2_ENABLE_SPLICED
                {   // the splice part of the HISAT2 score key, from the edits alone: GenomeHit::spliced() and the
                    // splicescore rule of calculateScore (hi_aligner.h:3745-3817) in the hit's (reference-forward) orientation
                    const uint32_t rdlen = r.rdlen, n = r.nedits;
                    bool spl = false, known = true; double ss = 0; uint32_t nss = 0;
                    for (uint32_t e = 0; e < n; e++) {
                        const Ht2Edit& ed = r.edits[e];
                        if (ed.type != HT2_EDIT_SPL) continue;
                        spl = true; known = known && ht2_spl_known(ed);
                        if (ht2_spl_known(ed)) continue;
                        uint32_t before = 0, after = 0;   // in hit order: plain mismatches before, mismatches + gaps after
                        for (uint32_t k = 0; k < n; k++) {
                            if (k == e) continue;
                            const Ht2Edit& o2 = r.edits[k];
                            const bool isBefore = r.fw ? (k < e) : (k > e);
                            if (isBefore) { if (o2.type == HT2_EDIT_MM && o2.snpID == HT2_IDX_MAX32) before++; }
                            else if (o2.type == HT2_EDIT_MM || o2.type == HT2_EDIT_READ_GAP || o2.type == HT2_EDIT_REF_GAP) after++;
                        }
                        const uint32_t q = ed.pos + r.trim5p;                  // 5'->3' offset of the splice in the read
                        int left_anchor = (int)(r.fw ? q : rdlen - q), right_anchor = (int)rdlen - left_anchor;
                        left_anchor -= (int)(before * 2); right_anchor -= (int)(after * 2);
                        int shorter = left_anchor < right_anchor ? left_anchor : right_anchor;
                        if (shorter <= 0) shorter = 1;
                        if (shorter <= 15) { nss++; ss += (double)ht2_spl_len(ed); }
                    }
                    if (nss > 1) ss /= (double)nss;
                    r.spliced = spl ? 1 : 0; r.knownTranscripts = (spl && known) ? 1 : 0; r.splicescore = ss;
                }
#endif
            }
        }
        for (uint32_t k = 0; k < rr.n_pairs; k++)
            o.pairs.push_back(std::make_pair(res->pairs[2 * (rr.pair_off + k)], res->pairs[2 * (rr.pair_off + k) + 1]));
        if (b->paired) ht2_finish_paired(sam, img, P, rd1, rd2, f1, f2, o);
        else ht2_finish_unpaired(sam, img, P, rd1, f1, o);
      }
    };
    if (nth > 64) nth = 64;
    if (nth < 1 || units < 4096) nth = 1;
    std::vector<std::string> parts(nth);
    if (nth == 1) doRange(0, units, parts[0]);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nth; t++) {
            const uint32_t u0 = (uint32_t)((uint64_t)units * t / nth), u1 = (uint32_t)((uint64_t)units * (t + 1) / nth);
            th.emplace_back([&, t, u0, u1]() { doRange(u0, u1, parts[t]); });
        }
        for (auto& x : th) x.join();
    }
    size_t total = 0;
    std::vector<size_t> at(parts.size());
    for (size_t t = 0; t < parts.size(); t++) { at[t] = total; total += parts[t].size(); }
    char* p = (char*)malloc(total + 1);
    if (!p) return false;
    if (nth == 1) memcpy(p, parts[0].data(), parts[0].size());
    else {   // every thread places its own chunk (first touch of the output pages is spread out too)
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nth; t++) th.emplace_back([&, t]() { memcpy(p + at[t], parts[t].data(), parts[t].size()); });
        for (auto& x : th) x.join();
    }
    p[total] = 0;
    *out = p; if (out_len) *out_len = total;
    return true;
}
