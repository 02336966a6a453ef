// issl_sites.cu -- off-target site extraction on the device: the step before the index
// (SURVEY.md §8f rank 3).  Replaces /root/reference/src/crackling/utils/extractOfftargets.py:
//   * FASTA reading: explodeMultiFastaFile :26-62 and processingNode :73-90 (header lines dropped, every other
//     line stripped of surrounding whitespace, upper-cased and concatenated per record);
//   * the two look-ahead regexes :23-24 and the slicing :97-106 -- a forward site is the first 20 characters of a
//     match of [ACG][ACGT]{19}[ACGT][AG]G, a reverse site is rc() of the FIRST 20 characters of a match of
//     C[CT][ACGT][ACGT]{19}[TGC] (this snapshot's slicing, kept as is);
//   * the global sort :112-191 (Python string order = A < C < G < T per base, first base most significant).
// Sites are held as 40-bit sort keys in HBM; they can be written out as the text file the reference tool writes
// or handed straight to the index builder (issl_device.cu) without touching the disk.

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include "issl_device_common.cuh"
#include "issl_internal.h"

using namespace issl;

namespace {

constexpr int kSiteLen = 20;       // characters kept per site
constexpr int kWindow = 23;        // characters a pattern spans (site + N + 2 PAM characters)
constexpr int kCarry = kWindow - 1;
constexpr uint8_t kOther = 4;      // a kept character that is not A/C/G/T: breaks every window that covers it
constexpr uint8_t kDrop = 255;     // newline, stripped whitespace, header text
constexpr int kStartsPerThread = 32;

__host__ __device__ __forceinline__ bool is_eol(unsigned char c) { return c == '\n' || c == '\r'; }   // universal newlines
// ASCII subset of str.strip()'s whitespace
__host__ __device__ __forceinline__ bool is_space(unsigned char c) { return (c >= 0x09 && c <= 0x0d) || (c >= 0x1c && c <= 0x20); }

// raw FASTA bytes -> one code per byte: 0..3 = A C G T (either case), kOther, or kDrop.
// atLineStart: the chunk begins at the start of a line.  stripLeading: leading whitespace of a line is dropped
// too (the single-input path strips both sides, :34 and :59; the multi-input path only the right side, :89).
__global__ void k_fasta_codes(const unsigned char *raw, uint64_t n, int atLineStart, int stripLeading, uint8_t *codes)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned char c = raw[i];
    uint8_t code;
    if (is_eol(c)) code = kDrop;
    else if (is_space(c)) {
        uint64_t j = i + 1;
        while (j < n && is_space(raw[j]) && !is_eol(raw[j])) j++;
        bool drop = (j == n) || is_eol(raw[j]);                  // trailing
        if (!drop && stripLeading) {
            uint64_t k = i;
            while (k > 0 && is_space(raw[k - 1]) && !is_eol(raw[k - 1])) k--;
            drop = (k == 0) ? (atLineStart != 0) : is_eol(raw[k - 1]);   // leading
        }
        code = drop ? kDrop : kOther;
    } else {
        switch (c & 0xDF) {   // upper()
            case 'A': code = 0; break;
            case 'C': code = 1; break;
            case 'G': code = 2; break;
            case 'T': code = 3; break;
            default: code = kOther; break;
        }
        if (c < 'A') code = kOther;   // digits and punctuation share low bits with letters once masked
    }
    codes[i] = code;
}

// header lines and skipped records: the first byte stays as a separator, the rest is dropped
__global__ void k_mask_ranges(const uint64_t *ranges, uint32_t nRanges, uint64_t chunkBegin, uint64_t chunkEnd, uint8_t *codes)
{
    const uint64_t b = ranges[2 * blockIdx.x], e = ranges[2 * blockIdx.x + 1];
    const uint64_t lo = max(b, chunkBegin), hi = min(e, chunkEnd);
    for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) codes[i - chunkBegin] = (i == b) ? kOther : kDrop;
}

struct NotDropped {
    __host__ __device__ __forceinline__ bool operator()(const uint8_t &c) const { return c != kDrop; }
};

// codes[0..total): carry (kCarry codes of the previous chunk) followed by this chunk's kept characters.
// Each thread walks kStartsPerThread window starts with two rolling registers: w (newest code in the low bits:
// the window's first character ends up at bits 44..45) and r (newest code in the high bits: window character k
// at bits 2k).  onMatch(forward?, key) is called for every match, forward before reverse.
template <class F>
__device__ __forceinline__ void walk_windows(const uint8_t *codes, uint64_t s0, uint64_t s1, F &&onMatch)
{
    constexpr uint64_t kMask46 = (1ull << (2 * kWindow)) - 1, kMask40 = (1ull << (2 * kSiteLen)) - 1;
    uint64_t w = 0, r = 0;
    uint32_t run = 0;
    for (uint64_t p = s0; p < s1 + kCarry; p++) {
        const uint32_t c = codes[p];
        run = c < 4 ? run + 1 : 0;
        w = ((w << 2) | (c & 3)) & kMask46;
        r = (r >> 2) | ((uint64_t)(c & 3) << (2 * (kWindow - 1)));
        if (p < s0 + kCarry || run < (uint32_t)kWindow) continue;
        const uint32_t first = (uint32_t)(w >> 44) & 3, second = (uint32_t)(w >> 42) & 3, c21 = (uint32_t)(w >> 2) & 3, last = (uint32_t)w & 3;
        // [ACG][ACGT]{19}[ACGT][AG]G -> match[0:20]
        if (first != 3 && (c21 == 0 || c21 == 2) && last == 2) onMatch((w >> (2 * (kWindow - kSiteLen))) & kMask40);
        // C[CT][ACGT][ACGT]{19}[TGC] -> rc(match[0:20]): site base j = complement of window base 19-j, i.e. window
        // base k lands, complemented, at key bits 2k
        if (first == 1 && (second == 1 || second == 3) && last != 0) onMatch((~r) & kMask40);
    }
}

// EMIT = false: count matches into *counter.  EMIT = true: reserve room with one atomicAdd per thread, write keys.
template <bool EMIT>
__global__ void k_site_match(const uint8_t *codes, uint64_t total, unsigned long long *counter, uint64_t *keys, unsigned long long keyCap)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t s0 = t * kStartsPerThread;
    const bool active = total >= (uint64_t)kWindow && s0 <= total - kWindow;
    const uint64_t s1 = active ? min(s0 + (uint64_t)kStartsPerThread, total - kWindow + 1) : s0;   // starts [s0, s1)
    uint32_t found = 0;
    if (active) walk_windows(codes, s0, s1, [&](uint64_t) { found++; });
    if (!EMIT) {
        const unsigned int sum = __reduce_add_sync(0xffffffffu, found);   // whole warps: nobody has returned yet
        if ((threadIdx.x & 31) == 0 && sum) atomicAdd(counter, (unsigned long long)sum);
        return;
    }
    if (!found) return;
    unsigned long long out = atomicAdd(counter, (unsigned long long)found);
    walk_windows(codes, s0, s1, [&](uint64_t key) { if (out < keyCap) keys[out] = key; out++; });
}

// the last kCarry codes of [0, total) move to the front (total >= kCarry always holds: the buffer starts with a carry)
__global__ void k_keep_carry(uint8_t *codes, uint64_t total)
{
    const uint8_t v = codes[total - kCarry + threadIdx.x];
    __syncthreads();
    codes[threadIdx.x] = v;
}

__global__ void k_fill(uint8_t *p, uint32_t n, uint8_t v)
{
    if (threadIdx.x < n) p[threadIdx.x] = v;
}

// sorted keys -> text lines ("ACGT..."+LF), first base = most significant pair
__global__ void k_keys_to_text(const uint64_t *keys, uint64_t n, char *text)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint64_t k = keys[t];
    char *o = text + t * (kSiteLen + 1);
#pragma unroll
    for (int j = 0; j < kSiteLen; j++) o[j] = "ACGT"[(k >> (2 * (kSiteLen - 1 - j))) & 3];
    o[kSiteLen] = '\n';
}

}  // namespace

struct issl_sites {
    int dev = -1;
    cudaStream_t stream = nullptr;
    DBuf keys, keysAlt, raw, codes, compact, selTemp, ranges;
    uint8_t *stage[2] = {nullptr, nullptr};
    cudaEvent_t stageFree[2] = {nullptr, nullptr};
    unsigned long long *hCount = nullptr;   // pinned: [0] matches of the current chunk, [1] kept characters
    unsigned long long *dCount = nullptr;
    uint64_t n = 0;
    uint64_t characters = 0;   // sequence characters seen
    bool sorted = true;
    size_t chunkBytes = 128ull << 20;
};

extern "C" void issl_sites_destroy(issl_sites *s)
{
    if (!s) return;
    cudaSetDevice(s->dev);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (DBuf *b : {&s->keys, &s->keysAlt, &s->raw, &s->codes, &s->compact, &s->selTemp, &s->ranges}) b->release();
    for (int b = 0; b < 2; b++) {
        if (s->stage[b]) cudaFreeHost(s->stage[b]);
        if (s->stageFree[b]) cudaEventDestroy(s->stageFree[b]);
    }
    if (s->hCount) cudaFreeHost(s->hCount);
    if (s->dCount) cudaFree(s->dCount);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

extern "C" int issl_sites_create(int cuda_device, issl_sites **out)
{
    if (!out) return issl_set_error(ISSL_ERR_ARG, "issl_sites_create: null argument");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return issl_set_error(ISSL_ERR_NO_DEVICE, "no CUDA device available (%s); libissl_cuda has no CPU fallback",
                              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (cuda_device < 0 || cuda_device >= n) return issl_set_error(ISSL_ERR_NO_DEVICE, "CUDA device %d does not exist (%d present)", cuda_device, n);
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, cuda_device));
    if (p.major != 10) return issl_set_error(ISSL_ERR_NO_DEVICE, "CUDA device %d (%s) is not an sm_100 part", cuda_device, p.name);
    CK(cudaSetDevice(cuda_device));
    issl_sites *s = new issl_sites();
    s->dev = cuda_device;
    if (const char *env = getenv("ISSL_EXTRACT_CHUNK")) {   // bytes of FASTA per device pass (tests shrink it)
        const long long v = atoll(env);
        if (v >= 64 && v <= (1ll << 30)) s->chunkBytes = (size_t)v;
    }
    e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMallocHost(&s->hCount, 2 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&s->dCount, 2 * sizeof(unsigned long long));
    for (int b = 0; b < 2 && e == cudaSuccess; b++) {
        e = cudaMallocHost(&s->stage[b], s->chunkBytes);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->stageFree[b], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        issl_sites_destroy(s);
        return issl_set_error(ISSL_ERR_CUDA, "issl_sites_create: %s", cudaGetErrorString(e));
    }
    *out = s;
    return ISSL_OK;
}

namespace {

// Header lines of a FASTA buffer, [begin, end) without the line terminator.  A header is a line whose first
// character is '>' (:35, :82); on the single-input path the line has been stripped first (:34).
void find_headers(const unsigned char *text, size_t bytes, bool stripLeading, std::vector<std::pair<uint64_t, uint64_t>> &out)
{
    constexpr size_t kSlice = 4u << 20;
    const long slices = (long)((bytes + kSlice - 1) / kSlice);
    std::vector<std::vector<std::pair<uint64_t, uint64_t>>> per((size_t)slices);
#pragma omp parallel for schedule(dynamic, 4) if (slices > 8)
    for (long k = 0; k < slices; k++) {
        const size_t b = (size_t)k * kSlice, e = std::min(bytes, b + kSlice);
        for (const unsigned char *p = (const unsigned char *)memchr(text + b, '>', e - b); p;
             p = (p + 1 < text + e) ? (const unsigned char *)memchr(p + 1, '>', (size_t)(text + e - p - 1)) : nullptr) {
            size_t i = (size_t)(p - text), j = i;
            if (stripLeading) while (j > 0 && is_space(text[j - 1]) && !is_eol(text[j - 1])) j--;
            if (j != 0 && !is_eol(text[j - 1])) continue;   // '>' in the middle of a line is just a character
            size_t end = i;
            while (end < bytes && !is_eol(text[end])) end++;
            per[(size_t)k].emplace_back((uint64_t)i, (uint64_t)end);
        }
    }
    for (auto &v : per) out.insert(out.end(), v.begin(), v.end());
}

int grow_keys(issl_sites *s, uint64_t need)
{
    if (need * 8 <= s->keys.cap) return ISSL_OK;
    DBuf bigger;
    CKR(bigger.ensure(std::max<uint64_t>(need, s->n + s->n / 2) * 8));
    if (s->n) CK(cudaMemcpyAsync(bigger.p, s->keys.p, s->n * 8, cudaMemcpyDeviceToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    s->keys.swap(bigger);   // the old allocation leaves with `bigger`
    return ISSL_OK;
}

}  // namespace

extern "C" int issl_sites_add_fasta(issl_sites *s, const char *text_, size_t bytes, int single_input)
{
    if (!s || (bytes && !text_)) return issl_set_error(ISSL_ERR_ARG, "issl_sites_add_fasta: null argument");
    CK(cudaSetDevice(s->dev));
    if (bytes == 0) return ISSL_OK;
    const unsigned char *text = reinterpret_cast<const unsigned char *>(text_);
    const bool stripLeading = single_input != 0;
    cudaStream_t st = s->stream;

    // records: header lines delimit them.  On the multi-input path records are keyed by header text inside one
    // file and a repeated header starts the record afresh (:83: seqsByHeader[header] = []), so only the LAST
    // record of each header text contributes; the single-input path writes every record to its own file.
    std::vector<std::pair<uint64_t, uint64_t>> headers;
    find_headers(text, bytes, stripLeading, headers);
    std::vector<uint64_t> mask;   // [begin, end) pairs: first byte becomes a separator, the rest is dropped
    mask.reserve(headers.size() * 2);
    if (!stripLeading && headers.size() > 1) {
        std::unordered_map<std::string, size_t> last;
        for (size_t k = 0; k < headers.size(); k++) {
            // :82 keys on line[1:], which still holds the line terminator -- identical for every header but possibly
            // the file's last line; comparing the text without it is the same except for that corner
            std::string key(reinterpret_cast<const char *>(text) + headers[k].first + 1, headers[k].second - headers[k].first - 1);
            auto it = last.find(key);
            if (it != last.end()) {   // the earlier record of this name is discarded: extend its mask over its sequence
                const size_t prev = it->second;
                const uint64_t seqEnd = prev + 1 < headers.size() ? headers[prev + 1].first : bytes;
                mask[2 * prev + 1] = seqEnd;
            }
            last[key] = k;
            mask.push_back(headers[k].first);
            mask.push_back(headers[k].second);
        }
    } else {
        for (auto &h : headers) { mask.push_back(h.first); mask.push_back(h.second); }
    }
    if (!mask.empty()) {
        CKR(s->ranges.ensure(mask.size() * 8));
        CK(cudaMemcpyAsync(s->ranges.p, mask.data(), mask.size() * 8, cudaMemcpyHostToDevice, st));
    }

    const size_t chunk = s->chunkBytes;
    CKR(s->raw.ensure(chunk));
    CKR(s->codes.ensure(chunk));
    CKR(s->compact.ensure(chunk + kCarry + 64));
    size_t selBytes = 0;
    CK(cub::DeviceSelect::If(nullptr, selBytes, s->codes.as<uint8_t>(), s->compact.as<uint8_t>() + kCarry, s->dCount + 1, (int)chunk, NotDropped(), st));
    CKR(s->selTemp.ensure(selBytes));
    // a new file never continues the previous one's last window
    k_fill<<<1, 32, 0, st>>>(s->compact.as<uint8_t>(), kCarry, kOther);

    size_t maskCursor = 0;   // first mask range that may still intersect the chunks to come
    int buf = 0;
    bool atLineStart = true;
    for (size_t off = 0; off < bytes;) {
        // cut after a line terminator; a line longer than the chunk is cut after a non-blank character
        size_t end = std::min(bytes, off + chunk);
        if (end < bytes) {
            size_t cut = end;
            while (cut > off && !is_eol(text[cut - 1])) cut--;
            if (cut == off) {
                cut = end;
                while (cut > off && is_space(text[cut - 1])) cut--;
                if (cut == off) return issl_set_error(ISSL_ERR_UNSUPPORTED, "a run of more than %zu blank characters inside one line", chunk);
            }
            end = cut;
        }
        const size_t n = end - off;
        CK(cudaEventSynchronize(s->stageFree[buf]));
        parallel_copy(s->stage[buf], text + off, n);
        CK(cudaMemcpyAsync(s->raw.p, s->stage[buf], n, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(s->stageFree[buf], st));
        buf ^= 1;
        k_fasta_codes<<<blocks_for(n, 256), 256, 0, st>>>(s->raw.as<unsigned char>(), n, atLineStart ? 1 : 0, stripLeading ? 1 : 0, s->codes.as<uint8_t>());
        // mask ranges intersecting [off, end)
        while (maskCursor < mask.size() / 2 && mask[2 * maskCursor + 1] <= off) maskCursor++;
        size_t last = maskCursor;
        while (last < mask.size() / 2 && mask[2 * last] < end) last++;
        if (last > maskCursor)
            k_mask_ranges<<<(unsigned)(last - maskCursor), 256, 0, st>>>(s->ranges.as<uint64_t>() + 2 * maskCursor, (uint32_t)(last - maskCursor),
                                                                          (uint64_t)off, (uint64_t)end, s->codes.as<uint8_t>());
        CK(cub::DeviceSelect::If(s->selTemp.p, selBytes, s->codes.as<uint8_t>(), s->compact.as<uint8_t>() + kCarry, s->dCount + 1, (int)n, NotDropped(), st));
        CK(cudaMemsetAsync(s->dCount, 0, sizeof(unsigned long long), st));
        CK(cudaMemcpyAsync(s->hCount + 1, s->dCount + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const uint64_t kept = s->hCount[1], total = kept + kCarry;
        s->characters += kept;
        if (total >= (uint64_t)kWindow) {
            const uint64_t starts = total - kWindow + 1;
            const unsigned blocks = blocks_for((starts + kStartsPerThread - 1) / kStartsPerThread, 256);
            k_site_match<false><<<blocks, 256, 0, st>>>(s->compact.as<uint8_t>(), total, s->dCount, nullptr, 0);
            CK(cudaMemcpyAsync(s->hCount, s->dCount, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            const uint64_t found = s->hCount[0];
            if (found) {
                CKR(grow_keys(s, s->n + found));
                CK(cudaMemsetAsync(s->dCount, 0, sizeof(unsigned long long), st));
                k_site_match<true><<<blocks, 256, 0, st>>>(s->compact.as<uint8_t>(), total, s->dCount, s->keys.as<uint64_t>() + s->n, found);
                s->n += found;
                s->sorted = false;
            }
        }
        k_keep_carry<<<1, kCarry, 0, st>>>(s->compact.as<uint8_t>(), total);
        CK(cudaGetLastError());
        atLineStart = is_eol(text[end - 1]);
        off = end;
    }
    CK(cudaStreamSynchronize(st));
    s->characters -= mask.size() / 2;   // every header line left one separator among the kept characters
    return ISSL_OK;
}

extern "C" int issl_sites_count(const issl_sites *s, uint64_t *sites, uint64_t *characters)
{
    if (!s) return issl_set_error(ISSL_ERR_ARG, "issl_sites_count: null argument");
    if (sites) *sites = s->n;
    if (characters) *characters = s->characters;
    return ISSL_OK;
}

static int sort_sites(issl_sites *s)
{
    if (s->sorted || s->n < 2) { s->sorted = true; return ISSL_OK; }
    CKR(s->keysAlt.ensure(s->n * 8));
    cub::DoubleBuffer<uint64_t> db(s->keys.as<uint64_t>(), s->keysAlt.as<uint64_t>());
    size_t tb = 0;
    CK(cub::DeviceRadixSort::SortKeys(nullptr, tb, db, s->n, 0, 2 * kSiteLen, s->stream));
    DBuf tmp;
    CKR(tmp.ensure(tb));
    CK(cub::DeviceRadixSort::SortKeys(tmp.p, tb, db, s->n, 0, 2 * kSiteLen, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    tmp.release();
    if (db.Current() != s->keys.as<uint64_t>()) s->keys.swap(s->keysAlt);
    s->sorted = true;
    return ISSL_OK;
}

extern "C" int issl_sites_write_text(issl_sites *s, const char *path)
{
    if (!s || !path) return issl_set_error(ISSL_ERR_ARG, "issl_sites_write_text: null argument");
    CK(cudaSetDevice(s->dev));
    CKR(sort_sites(s));
    FILE *fp = fopen(path, "wb");
    if (!fp) return issl_set_error(ISSL_ERR_IO, "cannot create %s", path);
    const uint64_t perPass = s->chunkBytes / (kSiteLen + 1);
    int rc = ISSL_OK;
    DBuf dtext;
    if ((rc = dtext.ensure(perPass * (kSiteLen + 1))) != ISSL_OK) { fclose(fp); return rc; }
    bool ok = true;
    int buf = 0;
    uint64_t pendingBytes[2] = {0, 0};
    // double-buffered: the D2H of pass k overlaps the fwrite of pass k-1
    auto flush = [&](int b) {
        if (!pendingBytes[b]) return;
        if (cudaEventSynchronize(s->stageFree[b]) != cudaSuccess) { ok = false; return; }
        if (fwrite(s->stage[b], 1, pendingBytes[b], fp) != pendingBytes[b]) ok = false;
        pendingBytes[b] = 0;
    };
    for (uint64_t o = 0; o < s->n && ok; o += perPass) {
        const uint64_t n = std::min(perPass, s->n - o);
        k_keys_to_text<<<blocks_for(n, 256), 256, 0, s->stream>>>(s->keys.as<uint64_t>() + o, n, dtext.as<char>());
        if (cudaMemcpyAsync(s->stage[buf], dtext.p, n * (kSiteLen + 1), cudaMemcpyDeviceToHost, s->stream) != cudaSuccess ||
            cudaEventRecord(s->stageFree[buf], s->stream) != cudaSuccess) { rc = issl_set_error(ISSL_ERR_CUDA, "D2H of site text failed"); break; }
        pendingBytes[buf] = n * (kSiteLen + 1);
        flush(buf ^ 1);   // the previous pass goes to disk while this one crosses the bus
        buf ^= 1;
        // dtext is rewritten by the next pass: this pass's copy must have left it
        if (cudaStreamSynchronize(s->stream) != cudaSuccess) { rc = issl_set_error(ISSL_ERR_CUDA, "site text pass failed"); break; }
    }
    flush(0); flush(1);
    dtext.release();
    if (fclose(fp) != 0) ok = false;
    if (rc != ISSL_OK) return rc;
    if (!ok) return issl_set_error(ISSL_ERR_IO, "short write to %s", path);
    return ISSL_OK;
}

extern "C" int issl_sites_read_keys(issl_sites *s, uint64_t first, uint64_t n, uint64_t *out)
{
    if (!s || (n && !out)) return issl_set_error(ISSL_ERR_ARG, "issl_sites_read_keys: null argument");
    if (first > s->n || n > s->n - first) return issl_set_error(ISSL_ERR_ARG, "issl_sites_read_keys: range outside the %llu sites held", (unsigned long long)s->n);
    CK(cudaSetDevice(s->dev));
    CKR(sort_sites(s));
    if (n) CK(cudaMemcpy(out, s->keys.as<uint64_t>() + first, n * 8, cudaMemcpyDeviceToHost));
    return ISSL_OK;
}

extern "C" int issl_device_create_from_sites(issl_sites *s, uint32_t sliceWidth, int layout, issl_device **out)
{
    if (!s || !out) return issl_set_error(ISSL_ERR_ARG, "issl_device_create_from_sites: null argument");
    *out = nullptr;
    if (s->n == 0) return issl_set_error(ISSL_ERR_ARG, "issl_device_create_from_sites: no sites were extracted");
    CK(cudaSetDevice(s->dev));
    CKR(sort_sites(s));
    // the builder consumes its key buffer: give it a copy so the site list stays usable
    DBuf work;
    CKR(work.ensure(s->n * 8));
    CKR(s->keysAlt.ensure(s->n * 8));
    CK(cudaMemcpyAsync(work.p, s->keys.p, s->n * 8, cudaMemcpyDeviceToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    const int rc = issl_internal_device_from_keys(s->dev, layout, work.as<uint64_t>(), s->keysAlt.as<uint64_t>(), s->n, kSiteLen, sliceWidth, out);
    work.release();
    return rc;
}
