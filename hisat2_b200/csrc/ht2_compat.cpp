// ht2_compat.cpp -- the hisat2lib C API (hisat2lib/ht2.h:67-150) exported from libht2gpu.so under the reference's
// own symbol names and struct layouts, so that existing ht2.h clients (the JNI module hisat2lib/java_jni, the
// Python module hisat2lib/pymodule, C callers) link against this library unchanged:
//   ht2_init / ht2_close / ht2_init_options                 ht2.h:67-70,  ht2_init.cpp:162-203
//   ht2_index_getrefnamebyid / ht2_index_getrefnames         ht2.h:79-94,  ht2_index.cpp:31-83
//   ht2_repeat_expand                                        ht2.h:123-127 (needs a repeat index: HT2_ERR_NOT_REPEAT here,
//                                                            the repeat path is not built -- DESIGN.md)
//   ht2_test_1 / ht2_repeat_dump_repeatmap                   ht2.h:144-145 (diagnostics)
// Host only: like the reference's ht2_init these never touch a device; alignment goes through ht2gpu.h.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "ht2_index.h"

extern "C" {

typedef int ht2_error_t;
enum { HT2_OK = 0, HT2_ERR = -1, HT2_ERR_NOT_REPEAT = -2 };
typedef void* ht2_handle_t;

struct ht2_options {            // ht2.h:44-58
    int offRate;
    int useMm, useShmem, mmSweep, noRefNames, noSplicedAlignment, gVerbose, startVerbose, sanityCheck;
    int useHaplotype;
};
typedef struct ht2_options ht2_option_t;

struct ht2_index_getrefnames_result { int count; char* names[0]; };          // ht2.h:81-84
struct ht2_position { uint32_t chr_id; int direction; uint64_t pos; };       // ht2.h:103-107
struct ht2_repeat_expand_result { int count; struct ht2_position positions[0]; };

}

namespace {
struct CompatHandle {
    Ht2Image* img;
    ht2_options opt;
    std::string name;
};
const ht2_options kDefaults = {-1, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // ht2_init.cpp:31-42
}

extern "C" {

ht2_error_t ht2_init_options(ht2_option_t* options)
{
    if (options == NULL) return HT2_ERR;
    memcpy(options, &kDefaults, sizeof(kDefaults));
    return HT2_OK;
}

ht2_handle_t ht2_init(const char* name, ht2_option_t* options)
{
    if (name == NULL) return NULL;
    std::string err;
    Ht2Image* img = ht2_image_load(name, err);
    if (!img) { fprintf(stderr, "ht2_init: %s\n", err.c_str()); return NULL; }   // the reference throws out of the HGFM constructor
    CompatHandle* h = new CompatHandle();
    h->img = img; h->name = name;
    h->opt = options ? *options : kDefaults;
    return (ht2_handle_t)h;
}

void ht2_close(ht2_handle_t handle)
{
    CompatHandle* h = (CompatHandle*)handle;
    if (!h) return;
    delete h->img;
    delete h;
}

const char* ht2_index_getrefnamebyid(ht2_handle_t handle, uint32_t chr_id)
{
    CompatHandle* h = (CompatHandle*)handle;
    if (!h || h->opt.noRefNames) return NULL;
    if (chr_id < h->img->header()->nRefs) return h->img->refName(chr_id);
    return NULL;
}

ht2_error_t ht2_index_getrefnames(ht2_handle_t handle, struct ht2_index_getrefnames_result** result_ptr)
{
    CompatHandle* h = (CompatHandle*)handle;
    if (!h || !result_ptr || h->opt.noRefNames) return HT2_ERR;
    const uint32_t n = h->img->header()->nRefs;
    // one malloc block the caller frees: the header, n + 1 name pointers, then the strings (ht2_index.cpp:49-80)
    const size_t hdr = sizeof(struct ht2_index_getrefnames_result) + sizeof(char*) * ((size_t)n + 1);
    size_t buf = 0;
    for (uint32_t i = 0; i < n; i++) buf += strlen(h->img->refName(i)) + 1;
    void* ptr = calloc(1, hdr + buf);
    if (!ptr) return HT2_ERR;
    struct ht2_index_getrefnames_result* r = (struct ht2_index_getrefnames_result*)ptr;
    r->count = (int)n;
    r->names[0] = (char*)ptr + hdr;
    for (uint32_t i = 0; i < n; i++) {
        const char* nm = h->img->refName(i);
        strcpy(r->names[i], nm);
        r->names[i + 1] = r->names[i] + strlen(nm) + 1;
    }
    *result_ptr = r;
    return HT2_OK;
}

ht2_error_t ht2_repeat_expand(ht2_handle_t, const char*, uint64_t, uint64_t, struct ht2_repeat_expand_result** result_ptr)
{
    if (result_ptr) *result_ptr = NULL;
    return HT2_ERR_NOT_REPEAT;   // no repeat index is loaded by this library
}

void ht2_test_1(ht2_handle_t handle)
{
    CompatHandle* h = (CompatHandle*)handle;
    if (!h) return;
    const uint32_t n = h->img->header()->nRefs;
    fprintf(stderr, "ht2lib: gfm refnames: %u\n", n + 1);
    for (uint32_t i = 0; i < n; i++) fprintf(stderr, "ht2lib:  %u -> %s\n", i, h->img->refName(i));
}

void ht2_repeat_dump_repeatmap(ht2_handle_t) {}

}
