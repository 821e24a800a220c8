// oracle/ref_dump.cpp -- TEST INFRASTRUCTURE.  A small driver that links the
// UNMODIFIED reference (headers/objects from /root/reference, built by
// oracle/Makefile) and dumps kernel-level golden vectors the end-to-end SAM
// cannot expose: for each read and strand, the chain of BWTHits that
// HI_Aligner::partialSearch (hi_aligner.h:6361) produces when called until
// the strand is exhausted, and for small ranges the joined/text coordinates
// GFM::getOffset + joinedToTextOff resolve (gfm.h:5682, 5527).
// Output (one line per record):
//   H <read#> <fw> <bwoff> <len> <top> <bot> <type> <pseudogeneStop> <anchorStop>
//   C <read#> <fw> <hit#> <row> <joinedOff> <tidx> <toff>
// usage: ref_dump <index_base> <reads.fa> <no_spliced 0|1>
#include <stdlib.h>
#include <stdint.h>
#include <iostream>
#include <fstream>
#include <string>
#include <cassert>
#include <stdexcept>
#include <math.h>
#include <utility>
#include <limits>
#include "alphabet.h"
#include "assert_helpers.h"
#include "endian_swap.h"
#include "hgfm.h"
#include "rfm.h"
#include "formats.h"
#include "sequence_io.h"
#include "tokenize.h"
#include "aln_sink.h"
#include "pat.h"
#include "threading.h"
#include "ds.h"
#include "aligner_metrics.h"
#include "sam.h"
#include "aligner_seed.h"
#include "splice_site.h"
#include "spliced_aligner.h"
#include "aligner_seed_policy.h"
#include "aligner_sw.h"
#include "aligner_sw_driver.h"
#include "aligner_cache.h"
#include "util.h"
#include "pe.h"
#include "tp.h"
#include "gp.h"
#include "simple_func.h"

using namespace std;

bool gColor = false, gColorExEnds = false, gReportOverhangs = false, gColorSeq = false, gColorEdit = false,
     gColorQual = false, gNoMaqRound = false, gStrandFix = false, gRangeMode = false;
int gVerbose = 0, gQuiet = 0;
bool gNofw = false, gNorc = false, gMate1fw = true, gMate2fw = false;
int gMinInsert = 0, gMaxInsert = 1000, gTrim5 = 0, gTrim3 = 0, gGapBarrier = 4, gAllowRedundant = 0;

extern void initializeCntLut();
extern void initializeCntBit();
MemoryTally gMemTally;
typedef uint32_t index_t;
typedef uint16_t local_index_t;

struct Dumper : public HI_Aligner<index_t, local_index_t> {
    Dumper(const GFM<index_t>& gfm) : HI_Aligner<index_t, local_index_t>(gfm, true, 0) {}
    void run(const GFM<index_t>& gfm, Read& rd, size_t rdid, const Scoring& sc, const ReportingParams& rp,
             bool linearPseudogene) {
        RandomSource rnd; rnd.init(0);
        for(int fwi = 0; fwi < 2; fwi++) {
            bool fw = (fwi == 0);
            ReadBWTHit<index_t> hit;
            hit.init(fw, (index_t)rd.length());
            size_t mineFw = 0, mineRc = 0;
            while(!hit.done()) {
                bool pseudogeneStop = linearPseudogene, anchorStop = true;
                partialSearch(gfm, rd, sc, rp, fw, 0, mineFw, mineRc, hit, rnd, pseudogeneStop, anchorStop);
                BWTHit<index_t>& ph = hit._partialHits.back();
                cout << "H " << rdid << " " << fw << " " << ph._bwoff << " " << ph._len << " " << ph._top << " "
                     << ph._bot << " " << ph._hit_type << " " << pseudogeneStop << " " << anchorStop << "\n";
                const bool graph = !gfm.gh().linearFM();
                if(graph && !ph.empty()) {
                    // graph indexes: node range and in-edge list of the hit (gfm.h:3759-3837)
                    cout << "G " << rdid << " " << fw << " " << (hit._partialHits.size() - 1) << " " << ph._node_top << " "
                         << ph._node_bot << " " << ph._node_iedge_count.size();
                    for(size_t e = 0; e < ph._node_iedge_count.size(); e++)
                        cout << " " << ph._node_iedge_count[e].first << ":" << ph._node_iedge_count[e].second;
                    cout << "\n";
                }
                if(!ph.empty() && ph._node_bot - ph._node_top <= 4) {
                    // element i of the node range: first BW row of the node (group_walk.h:545-560) and its offset
                    index_t num_iedges = 0; size_t e = 0;
                    for(index_t i = 0; i < ph._node_bot - ph._node_top; i++) {
                        while(e < ph._node_iedge_count.size()) {
                            if(i <= ph._node_iedge_count[e].first) break;
                            num_iedges += ph._node_iedge_count[e].second;
                            e++;
                        }
                        index_t r = ph._top + i + num_iedges;
                        index_t joff = gfm.getOffset(r, ph._node_top + i);
                        index_t tidx = 0, toff = 0, tlen = 0; bool straddled = false;
                        gfm.joinedToTextOff(ph._len, joff, tidx, toff, tlen, false, straddled);
                        cout << "C " << rdid << " " << fw << " " << (hit._partialHits.size() - 1) << " " << r << " "
                             << joff << " " << tidx << " " << toff << "\n";
                    }
                }
                if(hit.done()) break;
                if(!pseudogeneStop) { if(hit._cur + 1 < hit._len) hit._cur++; }
            }
        }
    }
};

int main(int argc, char** argv) {
    if(argc < 4) { cerr << "usage: ref_dump index reads.fa no_spliced" << endl; return 2; }
    string base = argv[1];
    bool no_spliced = atoi(argv[3]) != 0;
    initializeCntLut();
    initializeCntBit();
    ALTDB<index_t>* altdb = new ALTDB<index_t>();
    HGFM<index_t, local_index_t> gfm(base, altdb, NULL, NULL, -1, true, -1, 0, false, false, false, true, true, true, true,
                                     !no_spliced, false, false, false, false, false);
    gfm.loadIntoMemory(-1, true, true, true, true, false);
    SimpleFunc scoreMin; scoreMin.init(SIMPLE_FUNC_LINEAR, 0.0f, -0.2f);
    SimpleFunc nCeil; nCeil.init(SIMPLE_FUNC_LINEAR, 0.0f, std::numeric_limits<double>::max(), 2.0f, 0.1f);
    SimpleFunc icp, incp; icp.init(SIMPLE_FUNC_LOG, -8, 1); incp.init(SIMPLE_FUNC_LOG, -8, 1);
    Scoring sc(0, DEFAULT_MM_PENALTY_TYPE, 6, 2, 2, 1, scoreMin, nCeil, DEFAULT_N_PENALTY_TYPE, 1, false,
               5, 5, 3, 3, 4, 0, 12, 24, &icp, &incp);
    uint32_t khits = gfm.gh().linearFM() ? 5 : 10;
    ReportingParams rp(khits, max<uint32_t>(5, khits * 2), 0, 0, true, true, true, false, false, 0, false, false);
    Dumper d(gfm);
    ifstream in(argv[2]);
    string line, name, seq;
    size_t rdid = 0;
    auto flush = [&]() {
        if(name.empty() && seq.empty()) return;
        Read rd;
        rd.name.install(name.c_str());
        for(size_t i = 0; i < seq.size(); i++) {
            int c = seq[i];
            if(asc2dnacat[c] > 0) { rd.patFw.append(asc2dna[c]); rd.qual.append('I'); }
        }
        rd.finalize();
        rd.constructRevComps();
        rd.constructReverses();
        if(rd.length() > 0) d.run(gfm, rd, rdid, sc, rp, gfm.gh().linearFM() && !no_spliced);
        rdid++;
        name.clear(); seq.clear();
    };
    while(getline(in, line)) {
        if(!line.empty() && line[0] == '>') { flush(); name = line.substr(1); }
        else seq += line;
    }
    flush();
    return 0;
}
