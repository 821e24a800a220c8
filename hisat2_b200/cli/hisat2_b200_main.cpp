// hisat2-b200 -- command-line front end with the hisat2-align-s option surface
// (hisat2.cpp:541-764) for the subset that reaches the GPU alignment path.
// Host C++ over the C ABI only (include/ht2gpu.h): ht2gpu_run_reads parses the reads on all host threads,
// aligns and formats them on the device in overlapped batches, and hands back SAM text in read order.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "ht2gpu.h"

static void usage() {
    fprintf(stderr, "hisat2-b200 -x <index> {-U <r> | -1 <m1> -2 <m2>} [-S out.sam] [-f|-q] --no-spliced-alignment [-k N]\n"
                    "            [--mp MX,MN] [--sp MX,MN] [--np N] [--rdg C,L] [--rfg C,L] [--ignore-quals] [--nofw] [--norc]\n"
                    "            [-I N] [-X N] [--no-mixed] [--no-discordant] [--seed N] [--batch N] [--device N] [--gpus N] [-p N]\n"
                    "            [-5 N] [-3 N] [-s N] [-u N] [--phred33|--phred64] [--no-temp-splicesite]\n"
                    "            [--known-splicesite-infile F] [--novel-splicesite-infile F] [--novel-splicesite-outfile F]\n"
                    "            [--bowtie2-dp 0|1|2] [--score-min F,C,L] [--gbar N] [--sensitive] [--very-sensitive] [--fast]\n");
}
static void two(const char* a, int32_t& x, int32_t& y) { sscanf(a, "%d,%d", &x, &y); }

int main(int argc, char** argv)
{
    ht2gpu_options_t o; ht2gpu_default_options(&o);
    o.no_spliced_alignment = 0; // must be requested explicitly, like the reference's default is spliced
    const char *idx = NULL, *u = NULL, *m1 = NULL, *m2 = NULL, *out = NULL;
    bool fastq = true; size_t batchSz = 0 /* library default */; uint32_t gseed = 0;
    int trim5 = 0, trim3 = 0, threads = 0, gpus = 1; bool phred64 = false; uint64_t skip = 0, upto = 0;
    bool sensitive = false, verySensitive = false, fast = false, noTempSpliceSite = false;
    const char *knownSs = NULL, *novelSs = NULL, *novelOut = NULL;   // --known-splicesite-infile / --novel-splicesite-infile (hisat2.cpp:1689-1691)
    bool mpGiven = false;   // "--mp a,b" becomes MMP=Q,a,b, which switches the cost model back to quality-aware even
                            // under --ignore-quals (aligner_seed_policy.cpp:396-418)
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) { usage(); exit(1); } return argv[++i]; };
        if (a == "-x") idx = next(); else if (a == "-U") u = next(); else if (a == "-1") m1 = next(); else if (a == "-2") m2 = next();
        else if (a == "-S") out = next(); else if (a == "-f") fastq = false; else if (a == "-q") fastq = true;
        else if (a == "--no-spliced-alignment") o.no_spliced_alignment = 1;
        else if (a == "--no-temp-splicesite") noTempSpliceSite = true;   // spliced mode with an empty splice-site DB: the only spliced form a (HT2_SPLICED=1) library build honours
        else if (a == "--known-splicesite-infile") knownSs = next(); else if (a == "--novel-splicesite-infile") novelSs = next();
        else if (a == "--novel-splicesite-outfile") novelOut = next();
        else if (a == "-k") o.khits = atoi(next()); else if (a == "--max-seeds") o.max_seeds = atoi(next());
        else if (a == "--secondary") o.secondary = 1;
        else if (a == "--mp") { two(next(), o.mp_max, o.mp_min); mpGiven = true; } else if (a == "--sp") { two(next(), o.sp_max, o.sp_min); o.sp_min = o.sp_max; /* the reference reads BOTH values from the first number, aligner_seed_policy.cpp:438-441 */ }
        else if (a == "--np") o.np = atoi(next()); else if (a == "--rdg") two(next(), o.rdg_const, o.rdg_linear);
        else if (a == "--rfg") two(next(), o.rfg_const, o.rfg_linear); else if (a == "--ignore-quals") o.ignore_quals = 1;
        else if (a == "--nofw") o.nofw = 1; else if (a == "--norc") o.norc = 1;
        else if (a == "-I" || a == "--minins") o.min_frag = atoi(next()); else if (a == "-X" || a == "--maxins") o.max_frag = atoi(next());
        else if (a == "--no-mixed") o.no_mixed = 1; else if (a == "--no-discordant") o.no_discordant = 1;
        else if (a == "--seed") gseed = (uint32_t)atoi(next()); else if (a == "--batch") batchSz = (size_t)atol(next());
        else if (a == "--device") o.device = atoi(next());
        else if (a == "--gpus") gpus = atoi(next());   // devices device .. device + gpus - 1: batches round-robin, output in input order
        else if (a == "--bowtie2-dp") { o.bowtie2_dp = atoi(next()); if (o.bowtie2_dp < 0 || o.bowtie2_dp > 2) { fprintf(stderr, "Error: --bowtie2-dp arg must be 0, 1, or 2\n"); return 1; } }
        else if (a == "--gbar") o.gbar = atoi(next());
        else if (a == "--score-min" || a == "--min-score") {   // <type>,<const>,<coeff>; missing tokens keep the default (PARSE_FUNC, aligner_seed_policy.cpp:47-62)
            std::string v = next(); size_t c1 = v.find(','), c2 = c1 == std::string::npos ? c1 : v.find(',', c1 + 1);
            if (!v.empty()) o.score_min_type = v[0];
            if (c1 != std::string::npos) o.score_min_const = atof(v.c_str() + c1 + 1);
            if (c2 != std::string::npos) o.score_min_coeff = atof(v.c_str() + c2 + 1);
            if (o.score_min_type != 'C' && o.score_min_type != 'L' && o.score_min_type != 'S' && o.score_min_type != 'G') { fprintf(stderr, "Error: bad function type in --score-min\n"); return 1; }
        }
        else if (a == "--sensitive") sensitive = true; else if (a == "--very-sensitive") verySensitive = true; else if (a == "--fast") fast = true;
        else if (a == "-p" || a == "--threads") threads = atoi(next());   // host parser threads; the alignment itself runs on the GPU
        else if (a == "-5" || a == "--trim5") trim5 = atoi(next()); else if (a == "-3" || a == "--trim3") trim3 = atoi(next());
        else if (a == "-s" || a == "--skip") skip = (uint64_t)atoll(next()); else if (a == "-u" || a == "--upto" || a == "--qupto") upto = (uint64_t)atoll(next());
        else if (a == "--phred64" || a == "--phred64-quals" || a == "--solexa1.3-quals") phred64 = true; else if (a == "--phred33" || a == "--phred33-quals") phred64 = false;
        else if (a == "--reorder" || a == "-t" || a == "--time" || a == "--fr" || a == "--quiet" || a == "--no-unal-ignored") {}
        else if (a == "-h" || a == "--help") { usage(); return 0; }
        else { fprintf(stderr, "Error: option %s is not supported by hisat2-b200\n", a.c_str()); return 1; }
    }
    if (!idx || (!u && !(m1 && m2))) { usage(); return 1; }
    if (!o.no_spliced_alignment && !noTempSpliceSite)
        fprintf(stderr, "Warning: temporary splice sites (the reference's default) make results depend on the order reads are processed in;\n"
                        "         hisat2-b200 aligns every read independently, i.e. as with --no-temp-splicesite\n");
    if (mpGiven) o.ignore_quals = 0;
    // presets, spelled out the way hisat2.cpp:1889-1909 applies them after option parsing
    if (fast) {}
    else if (sensitive) {
        if (o.bowtie2_dp == 0) o.bowtie2_dp = 1;
        if (o.khits > 0 && o.khits < 10) o.khits = 10;   // an omitted -k stays "index default": the parse-time default is 10 (hisat2.cpp:336, 3903-3906)
        o.score_min_type = 'L'; o.score_min_const = (double)0.0f; o.score_min_coeff = (double)-0.5f;
    } else if (verySensitive) {
        o.bowtie2_dp = 2;
        if (o.khits < 30) o.khits = 30;
        o.score_min_type = 'L'; o.score_min_const = (double)0.0f; o.score_min_coeff = (double)-1.0f;
    }
    o.seed = gseed;
    ht2gpu_handle_t* h = NULL;
    if (ht2gpu_open(idx, &o, &h) != HT2GPU_OK) { fprintf(stderr, "Error: %s\n", ht2gpu_last_error(h)); return 1; }
    const bool paired = m1 != NULL;
    FILE* fo = out ? fopen(out, "wb") : stdout;
    if (!fo) { fprintf(stderr, "Error: cannot write %s\n", out); return 1; }
    { char* hd; size_t hl; ht2gpu_sam_header(h, &hd, &hl); fwrite(hd, 1, hl, fo); ht2gpu_free_text(hd);
      fprintf(fo, "@PG\tID:hisat2\tPN:hisat2-b200\tVN:0.2\n"); }
    ht2gpu_reads_input_t in; memset(&in, 0, sizeof(in));
    in.path1 = paired ? m1 : u; in.path2 = paired ? m2 : NULL;
    in.format = fastq ? 1 : 0; in.trim5 = trim5; in.trim3 = trim3; in.phred64 = phred64 ? 1 : 0; in.seed = gseed;
    in.skip = skip; in.upto = upto; in.batch_reads = (uint32_t)batchSz; in.threads = threads;
    ht2gpu_run_stats_t st;
    auto sink = [](void* ctx, const char* sam, size_t len) -> int { return fwrite(sam, 1, len, (FILE*)ctx) == len ? 0 : 1; };
    std::vector<ht2gpu_handle_t*> hs(1, h);
    for (int g = 1; g < gpus; g++) {   // replicate the index device to device
        ht2gpu_options_t og = o; og.device = o.device + g;
        ht2gpu_handle_t* hg = NULL;
        if (ht2gpu_open_peer(h, &og, &hg) != HT2GPU_OK) { fprintf(stderr, "Error: device %d: %s\n", og.device, ht2gpu_last_error(hg)); return 1; }
        hs.push_back(hg);
    }
    if (knownSs || novelSs) {
        uint32_t ns = 0;
        for (ht2gpu_handle_t* hh : hs)
            if (ht2gpu_load_splicesites(hh, knownSs, novelSs, &ns) != HT2GPU_OK) { fprintf(stderr, "Error: %s\n", ht2gpu_last_error(hh)); return 1; }
    }
    if (novelOut)
        for (ht2gpu_handle_t* hh : hs)
            if (ht2gpu_collect_splicesites(hh, 1) != HT2GPU_OK) { fprintf(stderr, "Error: %s\n", ht2gpu_last_error(hh)); return 1; }
    if (ht2gpu_run_reads_multi(hs.data(), (int)hs.size(), &in, sink, fo, &st) != HT2GPU_OK) { fprintf(stderr, "Error: %s\n", ht2gpu_last_error(h)); return 1; }
    if (novelOut) {
        uint64_t nw = 0;
        if (ht2gpu_write_novel_splicesites(hs.data(), (int)hs.size(), novelOut, &nw) != HT2GPU_OK) { fprintf(stderr, "Error: %s\n", ht2gpu_last_error(h)); return 1; }
    }
    for (size_t g = 1; g < hs.size(); g++) ht2gpu_close(hs[g]);
    if (st.n_err_reads)
        fprintf(stderr, "Warning: %llu read(s) exceeded a device-side capacity; their records may differ from hisat2's\n", (unsigned long long)st.n_err_reads);
    if (fo != stdout) fclose(fo);
    ht2gpu_close(h);
    return 0;
}
