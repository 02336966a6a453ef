// issl_internal.h -- structures shared by the host-side and device-side halves of libissl_cuda.
#ifndef ISSL_INTERNAL_H
#define ISSL_INTERNAL_H

#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <vector>

#include "issl_cuda.h"

// A validated view of an .issl image (file layout: isslCreateIndex.cpp:256-289).
struct issl_index {
    issl_info info{};
    uint64_t sliceLimit = 0;          // 2^sliceWidth lists per slice
    const uint8_t *base = nullptr;    // image
    size_t bytes = 0;
    bool mapped = false;              // true: we own an mmap of `bytes` at `base`
    const uint64_t *scorePairs = nullptr;   // scoresInFile x (mask, score bits)
    uint64_t scoresInFile = 0;        // entries physically present (std::map size; <= info.scoresCount)
    const uint64_t *offtargets = nullptr;   // offtargetsCount signatures
    const uint64_t *sizes = nullptr;        // sliceCount * sliceLimit list lengths
    const uint64_t *entries = nullptr;      // sliceCount * offtargetsCount x (occ << 32 | id)
};

int issl_set_error(int code, const char *fmt, ...);

// Sorted, de-duplicated (first occurrence wins, like phmap insert at isslScoreOfftargets.cpp:196)
// copy of a (mask, score) pair list.
void issl_sorted_score_table(const uint64_t *pairs, uint64_t n, std::vector<uint64_t> &masks,
                             std::vector<double> &scores);

// issl_device.cu: builds a resident index from unsorted site sort keys that already lie on `cuda_device`.
int issl_internal_device_from_keys(int cuda_device, int layout, uint64_t *dKeys, uint64_t *dKeysAlt, uint64_t nRaw,
                                   uint32_t seqLength, uint32_t sliceWidth, issl_device **out);

#endif
