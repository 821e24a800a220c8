// ht2_index.h -- host-side owner of a packed index image (see ht2_image.h).
#ifndef HT2_INDEX_H_
#define HT2_INDEX_H_

#include <string>
#include <vector>
#include <map>
#include <tuple>

#include "ht2_image.h"

struct Ht2Image {
    std::vector<uint8_t> blob;
    const Ht2ImageHeader* header() const { return (const Ht2ImageHeader*)blob.data(); }
    const char* refName(uint32_t i) const {
        return (const char*)blob.data() + header()->o_names + ((const uint32_t*)(blob.data() + header()->o_nameOffs))[i];
    }
    uint32_t refPlen(uint32_t i) const {
        return ((const uint32_t*)(blob.data() + header()->global.o_plen))[i];
    }
};

// Parse <base>.[1-8].ht2 into an image.  Returns NULL and sets 'err' on failure.
Ht2Image* ht2_image_load(const char* base, std::string& err);

// The read-only splice-site DB of a run (ht2_ssdb.h) from the text files of --known-splicesite-infile ('known' = true)
// and --novel-splicesite-infile (false): "<chr> <left> <right> <+|->" per site, 0-based, as hisat2_extract_splice_sites.py
// writes them (SpliceSiteDB::read, splice_site.cpp:727-775: names compared up to the first white space, unknown
// names skipped, a site already present is dropped).  Files are read in the order given.  An unreadable file is
// skipped like the reference does (hisat2.cpp:4101-4116); a malformed one is an error.
struct Ht2SsFile { std::string path; bool known; };
// --novel-splicesite-outfile: junction records of printed alignments (Ht2SsRec, ht2_ssdb.h) aggregated like
// SpliceSiteDB::addSpliceSite does (number of alignments per site, smallest edit distance), and written like
// SpliceSiteDB::print (splice_site.cpp:565-651): read-count cutoffs from the distribution over all sites, sites
// within 10 bases of a better-supported neighbour suppressed.  The result does not depend on the order of 'add'.
struct Ht2NovelSites {
    std::map<std::tuple<uint32_t, uint32_t, uint32_t, uint32_t>, std::pair<uint32_t, uint32_t>> sites;   // (ref, left, right, dir) -> (numreads, editdist)
    void add(const struct Ht2SsRec* recs, size_t n);
    void merge(const Ht2NovelSites& o);
    bool write(const Ht2Image& img, const char* path, uint64_t* nWritten, std::string& err) const;
};
bool ht2_ssdb_build(const Ht2Image& img, const std::vector<Ht2SsFile>& files, std::vector<uint8_t>& blob, uint32_t& nSites, std::string& err);

#endif
