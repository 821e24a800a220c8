// ht2_params.h -- alignment parameters shared by host and device code.
//
// Mirrors the option statics the reference hands to the hot path
// (hisat2.cpp:305-544 resetOptions; Scoring scoring.h:96-226; ReportingParams
// hisat2.cpp:3177-3189; TranscriptomePolicy hisat2.cpp:4076-4084; GraphPolicy
// :4086-4089).  Plain data so it can sit in __constant__ memory.
#ifndef HT2_PARAMS_H_
#define HT2_PARAMS_H_

#include <stdint.h>

#define HT2_PARAMS_MAX_RDLEN 1024

// The part the kernels receive BY VALUE (kernel parameter space, read through the constant cache).
struct Ht2ParamsCore {
    // Scoring
    int32_t mmpMax;        // --mp max (6)
    int32_t mmpMin;        // --mp min (2)
    int32_t scpMax;        // --sp max (2)
    int32_t scpMin;        // --sp min (1)
    int32_t npen;          // --np (1)
    int32_t rdGapConst;    // --rdg (5)
    int32_t rdGapLinear;   // --rdg (3)
    int32_t rfGapConst;    // --rfg (5)
    int32_t rfGapLinear;   // --rfg (3)
    int32_t mmcostConstant;// 1 = --ignore-quals (COST_MODEL_CONSTANT)
    int32_t canSplPen;     // --pen-cansplice (0)
    // Reporting
    uint32_t khits;        // -k (5 linear / 10 graph unless given)
    uint32_t kseeds;       // --max-seeds (max(5, 2k))
    uint32_t secondary;    // --secondary
    // Transcriptome policy
    uint32_t minIntronLen; // 20
    uint32_t maxIntronLen; // 500000
    uint32_t minAnchorLen; // 7
    uint32_t minAnchorLenNoncan; // 14
    uint32_t noSplicedAlignment; // --no-spliced-alignment
    // Graph policy
    uint32_t maxAltsTried; // 16
    // Search
    uint32_t anchorStop;   // 1
    uint32_t minK;         // ceil(log4(genome len)) (hi_aligner.h:3979-3984)
    uint32_t minKLocal;    // 8
    // Paired-end policy (pe.h): --fr, -I, -X
    int32_t  pePolicy;     // 0=FF 1=RR 2=FR 3=RF
    uint32_t minFrag;      // -I (0)
    uint32_t maxFrag;      // -X (1000)
    uint32_t gMate1fw;     // 1
    uint32_t gMate2fw;     // 0
    uint32_t nofw;         // --nofw
    uint32_t norc;         // --norc
    uint32_t mixed;        // !--no-mixed
    uint32_t discord;      // !--no-discordant
    // --bowtie2-dp (hisat2.cpp:293, 1770): 0 off, 1 when the anchor search found nothing >= minsc, 2 always
    uint32_t bowtie2Dp;
    int32_t  gapbar;       // --gbar (4): no gaps within this many rows of either read end (DP only)
#ifdef HT2_ENABLE_SPLICED
    int32_t  noncanSplPen; // --pen-noncansplice (12); --pen-cansplice is canSplPen above
#endif
};

// Host-side parameters = the core + the --score-min table.  The table reaches the device as a
// separate 1 KB buffer (DevBatch::minscTab): keeping it out of the by-value kernel parameter keeps
// every Ht2ParamsCore field a constant-bank read (a 1.2 KB parameter indexed dynamically, or the
// whole struct behind a global pointer, cost 6 % of the aligner's throughput).
struct Ht2Params : Ht2ParamsCore {
    // --score-min as a table: SimpleFunc::f<TAlScore>(len) for every read length (hisat2.cpp:3380,
    // simple_func.h:86-108), computed once on the host so that host and device agree for every
    // function type (C, L, S = sqrt, G = log).  NOT yet clamped to <= 0.
    int32_t  minscTab[HT2_PARAMS_MAX_RDLEN + 1];
};

#endif
