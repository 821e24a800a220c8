// ht2_index.h -- host-side owner of a packed index image (see ht2_image.h).
#ifndef HT2_INDEX_H_
#define HT2_INDEX_H_

#include <string>
#include <vector>

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

#endif
