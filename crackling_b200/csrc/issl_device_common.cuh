// issl_device_common.cuh -- small helpers shared by the device-side translation units of libissl_cuda.
#ifndef ISSL_DEVICE_COMMON_CUH
#define ISSL_DEVICE_COMMON_CUH

#include <algorithm>
#include <cstdint>
#include <cstring>

#include <cuda_runtime.h>

#include "issl_internal.h"

#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return issl_set_error(ISSL_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define CKR(call)                       \
    do {                                \
        int r_ = (call);                \
        if (r_ != ISSL_OK) return r_;   \
    } while (0)

namespace issl {

// growable device buffer
struct DBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return ISSL_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            e = cudaMalloc(&p, bytes);
            if (e != cudaSuccess) { p = nullptr; return issl_set_error(ISSL_ERR_NOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); }
            cap = bytes;
        } else cap = want;
        return ISSL_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

inline unsigned blocks_for(uint64_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

// host copy into a pinned staging buffer with all cores (a single memcpy stream tops out near 10 GB/s,
// well below what PCIe Gen5 can take)
inline void parallel_copy(void *dst, const void *src, size_t bytes)
{
    constexpr size_t kSlice = 1u << 20;
    const long slices = (long)((bytes + kSlice - 1) / kSlice);
#pragma omp parallel for schedule(static) if (slices > 4)
    for (long i = 0; i < slices; i++) {
        const size_t o = (size_t)i * kSlice, n = std::min(kSlice, bytes - o);
        memcpy(static_cast<uint8_t *>(dst) + o, static_cast<const uint8_t *>(src) + o, n);
    }
}

}  // namespace issl

#endif
