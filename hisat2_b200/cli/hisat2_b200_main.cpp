// hisat2-b200 -- command-line front end with the hisat2-align-s option surface
// (hisat2.cpp:541-764) for the subset that reaches the GPU alignment path.
// Host C++ over the C ABI only (include/ht2gpu.h); reads are parsed, packed into
// SoA batches, aligned on the device and printed as SAM.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "ht2gpu.h"

struct Reads { std::vector<uint8_t> seq, qual; std::vector<uint64_t> offs; std::vector<uint32_t> seeds; std::string names; size_t n; bool haveQual; };

static int dnacat(int c) { return c && strchr("ACGTacgtBDHKMNRSVWXYbdhkmnrsvwxy-", c) != NULL; }
static uint8_t asc2dna(int c) { switch (c) { case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; case 'N': case 'n': return 4; default: return 0; } }

// FASTA (pat.cpp:725-849) or FASTQ (pat.cpp:852-1100, 4-line records) into per-file vectors
static bool parseFile(const char* path, bool fastq, std::vector<std::string>& names, std::vector<std::string>& seqs, std::vector<std::string>& quals)
{
    FILE* f = strcmp(path, "-") ? fopen(path, "rb") : stdin;
    if (!f) { fprintf(stderr, "Error: could not open %s\n", path); return false; }
    std::string data; char buf[1 << 16]; size_t k;
    while ((k = fread(buf, 1, sizeof(buf), f)) > 0) data.append(buf, k);
    if (f != stdin) fclose(f);
    size_t p = 0, n = data.size(), cnt = 0;
    auto line = [&](std::string& out) { out.clear(); while (p < n && data[p] != '\n') { if (data[p] != '\r') out.push_back(data[p]); p++; } if (p < n) p++; };
    std::string l;
    if (!fastq) {
        std::string name, seq; bool have = false;
        while (p < n) {
            line(l);
            if (!l.empty() && l[0] == '>') {
                if (have) { names.push_back(name); seqs.push_back(seq); quals.push_back(std::string(seq.size(), 'I')); }
                name = l.substr(1); if (name.empty()) name = std::to_string(cnt); cnt++; seq.clear(); have = true;
            } else if (!l.empty() && (l[0] == '#' || l[0] == ';')) continue;
            else for (char c : l) if (dnacat((unsigned char)c)) seq.push_back((char)asc2dna(c));
        }
        if (have) { names.push_back(name); seqs.push_back(seq); quals.push_back(std::string(seq.size(), 'I')); }
    } else {
        while (p < n) {
            line(l); if (l.empty()) continue;
            if (l[0] != '@') { fprintf(stderr, "Error: reads file does not look like a FASTQ file\n"); return false; }
            std::string name = l.substr(1), s, plus, q; line(s); line(plus); line(q);
            std::string codes; std::string qq;
            for (size_t i = 0; i < s.size(); i++) if (dnacat((unsigned char)s[i])) { codes.push_back((char)asc2dna(s[i])); qq.push_back(i < q.size() ? q[i] : 'I'); }
            if (name.empty()) name = std::to_string(cnt); cnt++;
            names.push_back(name); seqs.push_back(codes); quals.push_back(qq);
        }
    }
    return true;
}

static void usage() {
    fprintf(stderr, "hisat2-b200 -x <index> {-U <r> | -1 <m1> -2 <m2>} [-S out.sam] [-f|-q] --no-spliced-alignment [-k N]\n"
                    "            [--mp MX,MN] [--sp MX,MN] [--np N] [--rdg C,L] [--rfg C,L] [--ignore-quals] [--nofw] [--norc]\n"
                    "            [-I N] [-X N] [--no-mixed] [--no-discordant] [--seed N] [--batch N] [--device N]\n"
                    "            [--bowtie2-dp 0|1|2] [--score-min F,C,L] [--gbar N] [--sensitive] [--very-sensitive] [--fast]\n");
}
static void two(const char* a, int32_t& x, int32_t& y) { sscanf(a, "%d,%d", &x, &y); }

int main(int argc, char** argv)
{
    ht2gpu_options_t o; ht2gpu_default_options(&o);
    o.no_spliced_alignment = 0; // must be requested explicitly, like the reference's default is spliced
    const char *idx = NULL, *u = NULL, *m1 = NULL, *m2 = NULL, *out = NULL;
    bool fastq = true; size_t batchSz = 1000000; uint32_t gseed = 0;
    bool sensitive = false, verySensitive = false, fast = false, noTempSpliceSite = false;
    bool mpGiven = false;   // "--mp a,b" becomes MMP=Q,a,b, which switches the cost model back to quality-aware even
                            // under --ignore-quals (aligner_seed_policy.cpp:396-418)
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) { usage(); exit(1); } return argv[++i]; };
        if (a == "-x") idx = next(); else if (a == "-U") u = next(); else if (a == "-1") m1 = next(); else if (a == "-2") m2 = next();
        else if (a == "-S") out = next(); else if (a == "-f") fastq = false; else if (a == "-q") fastq = true;
        else if (a == "--no-spliced-alignment") o.no_spliced_alignment = 1;
        else if (a == "--no-temp-splicesite") noTempSpliceSite = true;   // spliced mode with an empty splice-site DB: the only spliced form a (HT2_SPLICED=1) library build honours
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
        else if (a == "-p" || a == "--threads") next(); else if (a == "--reorder" || a == "-t" || a == "--fr") {}
        else if (a == "-h" || a == "--help") { usage(); return 0; }
        else { fprintf(stderr, "Error: option %s is not supported by hisat2-b200\n", a.c_str()); return 1; }
    }
    if (!idx || (!u && !(m1 && m2))) { usage(); return 1; }
    if (!o.no_spliced_alignment && !noTempSpliceSite) {
        fprintf(stderr, "Error: spliced alignment with temporary splice sites is order-dependent and not implemented; pass --no-spliced-alignment\n"
                        "       (--no-temp-splicesite is honoured by an experimental HT2_SPLICED=1 library build only)\n");
        return 1;
    }
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
    std::vector<std::string> n1, s1, q1, n2, s2, q2;
    bool paired = m1 != NULL;
    if (!parseFile(paired ? m1 : u, fastq, n1, s1, q1)) return 1;
    if (paired && !parseFile(m2, fastq, n2, s2, q2)) return 1;
    if (paired && n1.size() != n2.size()) { fprintf(stderr, "Error: mate files have different numbers of reads\n"); return 1; }
    FILE* fo = out ? fopen(out, "wb") : stdout;
    if (!fo) { fprintf(stderr, "Error: cannot write %s\n", out); return 1; }
    { char* hd; size_t hl; ht2gpu_sam_header(h, &hd, &hl); fwrite(hd, 1, hl, fo); ht2gpu_free_text(hd);
      fprintf(fo, "@PG\tID:hisat2\tPN:hisat2-b200\tVN:0.1\n"); }
    size_t total = n1.size();
    for (size_t s = 0; s < total; s += batchSz) {
        size_t e = s + batchSz < total ? s + batchSz : total;
        Reads R; R.n = 0; R.offs.push_back(0);
        auto add = [&](const std::string& nm, const std::string& sq, const std::string& ql, const char* suffix) {
            std::string name = nm;
            if (suffix && !(name.size() >= 2 && name[name.size() - 2] == '/')) name += suffix; // pat.cpp fixName
            R.seq.insert(R.seq.end(), sq.begin(), sq.end()); R.qual.insert(R.qual.end(), ql.begin(), ql.end());
            R.offs.push_back(R.seq.size());
            R.seeds.push_back(ht2gpu_read_seed((const uint8_t*)sq.data(), (const uint8_t*)ql.data(), (uint32_t)sq.size(), name.c_str(), gseed));
            R.names += name; R.names.push_back('\0'); R.n++;
        };
        for (size_t i = s; i < e; i++) { add(n1[i], s1[i], q1[i], paired ? "/1" : NULL); if (paired) add(n2[i], s2[i], q2[i], "/2"); }
        ht2gpu_read_batch_t b = {(uint32_t)R.n, paired ? 1 : 0, R.seq.data(), R.qual.data(), R.offs.data(), R.seeds.data()};
        ht2gpu_result_batch_t r;
        int rc = ht2gpu_align_batch(h, &b, &r);
        if (rc != HT2GPU_OK) { fprintf(stderr, "Error: %s\n", ht2gpu_last_error(h)); return 1; }
        char* sam; size_t len;
        if (ht2gpu_format_sam(h, &b, R.names.c_str(), &r, &sam, &len) != HT2GPU_OK) { fprintf(stderr, "Error: %s\n", ht2gpu_last_error(h)); return 1; }
        fwrite(sam, 1, len, fo);
        ht2gpu_free_text(sam); ht2gpu_free_results(&r);
    }
    if (fo != stdout) fclose(fo);
    ht2gpu_close(h);
    return 0;
}
