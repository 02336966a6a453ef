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

// growable device buffer; owns its allocation (freed on destruction) unless made a view()
struct DBuf {
    void *p = nullptr;
    size_t cap = 0;       // bytes allocated
    size_t used = 0;      // bytes last asked for
    bool owner = true;
    DBuf() = default;
    DBuf(const DBuf &) = delete;
    DBuf &operator=(const DBuf &) = delete;
    ~DBuf() { release(); }
    // scratch: over-allocates by a quarter so that slowly growing batches do not reallocate every call
    int ensure(size_t bytes) { return grow(bytes, bytes + bytes / 4 + 256); }
    // index storage: exactly what is asked for (a quarter of slack on an 87 GB index is 22 GB of HBM)
    int exact(size_t bytes) { return grow(bytes, bytes); }
    // a non-owning window onto memory the caller keeps
    void view(void *ptr, size_t bytes) { release(); p = ptr; cap = used = bytes; owner = false; }
    void release()
    {
        if (p && owner) cudaFree(p);
        p = nullptr; cap = used = 0; owner = true;
    }
    void swap(DBuf &o) { std::swap(p, o.p); std::swap(cap, o.cap); std::swap(used, o.used); std::swap(owner, o.owner); }
    template <class T> T *as() const { return static_cast<T *>(p); }

private:
    int grow(size_t bytes, size_t want)
    {
        used = bytes;
        if (bytes <= cap) return ISSL_OK;
        if (!owner) return issl_set_error(ISSL_ERR_NOMEM, "buffer view too small (%zu > %zu bytes)", bytes, cap);
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess && want != bytes) { cudaGetLastError(); want = bytes; e = cudaMalloc(&p, want); }
        if (e != cudaSuccess) {
            cudaGetLastError();
            p = nullptr; used = 0;
            return issl_set_error(ISSL_ERR_NOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
        }
        cap = want;
        return ISSL_OK;
    }
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
