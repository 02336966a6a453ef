// issl_triple.cuh -- ISSL_LAYOUT_TRIPLE: every slice list sub-divided by two further slices.
//
// seqLength 20, sliceWidth 8: a site is five bytes (slices 0..4, four bases each).  For each of the
// ten slice triples T = (a<b<c) the index keeps one more copy of the sites, bucketed by the 24-bit
// key  slice a | slice b << 8 | slice c << 16  (ascending site id inside a bucket, as inside the
// reference's lists, isslCreateIndex.cpp:225-233), storing per entry only the 16 bits the key does not
// determine (slices p<q, the complement of T).  The reference's list (slice i, value v) is the union
// of the 65 536 buckets of any triple containing i whose key has v at slice i.
//
// A guide does not stream its five lists (ref isslScoreOfftargets.cpp:344-382: ~12 M entries at human
// scale); it reads only the buckets that can contain a site the reference would score (see
// issl_triple_visits in issl_host.cpp: 1 390 buckets of ~35 entries at maxDist 4), and every entry
// read costs 2 bytes.  The hit set, the order hits are accumulated in, and so every printed digit are
// unchanged.  "ref:" = /root/reference/src/ISSL/.
#pragma once

#include "issl_kernels.cuh"

namespace issl {

constexpr uint32_t kTripleCount = 10;
constexpr uint32_t kTripleBuckets = 1u << 24;
constexpr int kTripleThreads = 128;              // 16 octets; one octet (8 lanes x 16 B) reads one bucket
constexpr int kTripleKeyBits = 35;               // survivor key = guide << 35 | lowest exact slice << 32 | site id

// slices of triple t: a, b, c (key bytes 0..2) then p, q (residual bytes 0..1)
__constant__ uint8_t c_tripleSlices[kTripleCount][5] = {
    {0, 1, 2, 3, 4}, {0, 1, 3, 2, 4}, {0, 1, 4, 2, 3}, {0, 2, 3, 1, 4}, {0, 2, 4, 1, 3},
    {0, 3, 4, 1, 2}, {1, 2, 3, 0, 4}, {1, 2, 4, 0, 3}, {1, 3, 4, 0, 2}, {2, 3, 4, 0, 1}};
// resp(E): the triple responsible for a site whose set of exactly matching slices is E (bit s = slice s)
__constant__ uint8_t c_tripleResp[32];

struct TripleView {
    const uint16_t *res;    // [10][stride] residual bits (slice p | slice q << 8) per bucket entry
    const uint32_t *ids;    // [10][stride] site id per bucket entry
    const uint32_t *offs;   // [10][2^24 + 1] first entry of every bucket
    uint64_t stride;        // entries reserved per triple (multiple of 8, >= N + 64)
};

__host__ __device__ __forceinline__ uint32_t triple_key(uint64_t sig, uint32_t a, uint32_t b, uint32_t c)
{
    return (uint32_t)((sig >> (8 * a)) & 0xFFull) | ((uint32_t)((sig >> (8 * b)) & 0xFFull) << 8) |
           ((uint32_t)((sig >> (8 * c)) & 0xFFull) << 16);
}
__host__ __device__ __forceinline__ uint32_t triple_res(uint64_t sig, uint32_t p, uint32_t q)
{
    return (uint32_t)((sig >> (8 * p)) & 0xFFull) | ((uint32_t)((sig >> (8 * q)) & 0xFFull) << 8);
}

// ------------------------------------------------------------------------------------------------
// construction (once per index): sort (key, id) per triple, then residuals + bucket offsets
// ------------------------------------------------------------------------------------------------
__global__ void k_triple_keys(const uint64_t *sig, uint64_t n, uint32_t t, uint32_t *keys, uint32_t *ids)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = triple_key(sig[i], c_tripleSlices[t][0], c_tripleSlices[t][1], c_tripleSlices[t][2]);
    ids[i] = (uint32_t)i;
}

__global__ void k_triple_residuals(const uint64_t *sig, const uint32_t *sortedIds, uint64_t n, uint32_t t, uint16_t *res)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    res[i] = (uint16_t)triple_res(sig[sortedIds[i]], c_tripleSlices[t][3], c_tripleSlices[t][4]);
}

// offs[k] = number of entries with key < k, for k in [0, 2^24]
__global__ void k_triple_offsets(const uint32_t *sortedKeys, uint64_t n, uint32_t *offs)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > kTripleBuckets) return;
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (sortedKeys[mid] < k) lo = mid + 1; else hi = mid;
    }
    offs[k] = (uint32_t)lo;
}

// ------------------------------------------------------------------------------------------------
// K1t: bucket scan.  ref isslScoreOfftargets.cpp:344-390 for one guide, restricted to the buckets
// that can hold a site within maxDist.
//
// grid = (guides, visit chunks); one CTA = one guide x a range of the visit table; an octet of lanes
// takes one bucket at a time: two 4-byte offsets (one sector), then the bucket's residuals as 16-byte
// vectors (8 entries per lane, 64 per octet step; a bucket holds ~35 at human scale).  Per 32-bit word
// (two entries): XOR with the guide's residual, fold to per-base flags, two POPC; the minimum of the
// vector is compared with the bucket's budget.  The offsets of the next visit are requested before the
// current bucket is processed, so that two dependent round trips per octet are in flight.
// Survivors are re-tested exactly out of line: the site is rebuilt from bucket key + residual, E is
// recomputed, and the hit is kept only if this triple is resp(E) -- the stateless replacement for the
// reference's toggle bitset (:385-390, :463): each hit is produced exactly once, in no particular
// order, tagged with min(E) and its id, which sort back into the reference's visiting order.
// ------------------------------------------------------------------------------------------------
struct TripleArgs {
    TripleView tv;
    const uint64_t *guides;
    const uint8_t *done;          // optional: guides that already left through the early exit
    const uint32_t *visits;       // issl_triple_visits entries of this wave
    uint32_t nVisits, visitsPerCta;
    uint64_t *hitKeys;
    unsigned long long *hitCount;
    uint64_t hitCap;
    unsigned long long *streamed; // [0] entries of visited buckets, [1] bucket visits
    int maxDist;
};

__device__ __noinline__ void triple_slow(const TripleArgs &a, uint32_t guide, uint64_t g, uint32_t t, uint32_t key,
                                         uint32_t first, uint32_t start, uint32_t end, uint4 r)
{
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    const uint32_t sa = c_tripleSlices[t][0], sb = c_tripleSlices[t][1], sc = c_tripleSlices[t][2],
                   sp = c_tripleSlices[t][3], sq = c_tripleSlices[t][4];
    const uint64_t keyBits = ((uint64_t)(key & 0xFFu) << (8 * sa)) | ((uint64_t)((key >> 8) & 0xFFu) << (8 * sb)) |
                             ((uint64_t)((key >> 16) & 0xFFu) << (8 * sc));
    for (uint32_t i = 0; i < 8; i++) {
        const uint32_t pos = first + i;
        if (pos < start || pos >= end) continue;
        const uint32_t res = (w[i >> 1] >> (16 * (i & 1u))) & 0xFFFFu;
        const uint64_t site = keyBits | ((uint64_t)(res & 0xFFu) << (8 * sp)) | ((uint64_t)(res >> 8) << (8 * sq));
        const uint64_t x = site ^ g;
        if (distance64(x) > a.maxDist) continue;
        uint32_t E = 0;
        for (uint32_t s = 0; s < 5; s++) E |= (uint32_t)(((x >> (8 * s)) & 0xFFull) == 0) << s;
        if (E == 0 || c_tripleResp[E] != t) continue;
        const uint32_t id = a.tv.ids[(uint64_t)t * a.tv.stride + pos];
        const unsigned long long slot = atomicAdd(a.hitCount, 1ull);
        if (slot < a.hitCap)
            a.hitKeys[slot] = ((uint64_t)guide << kTripleKeyBits) | ((uint64_t)(__ffs(E) - 1) << 32) | id;
    }
}

__global__ void __launch_bounds__(kTripleThreads) k_scan_triple(const TripleArgs a)
{
    const uint32_t guide = blockIdx.x;
    if (a.done && a.done[guide]) return;
    __shared__ uint32_t sKey[kTripleCount], sRes[kTripleCount];
    __shared__ unsigned long long sCount[2];
    const uint64_t g = a.guides[guide];
    if (threadIdx.x < kTripleCount) {
        const uint32_t t = threadIdx.x;
        sKey[t] = triple_key(g, c_tripleSlices[t][0], c_tripleSlices[t][1], c_tripleSlices[t][2]);
        sRes[t] = triple_res(g, c_tripleSlices[t][3], c_tripleSlices[t][4]) * 0x10001u;
    }
    if (threadIdx.x < 2) sCount[threadIdx.x] = 0;
    __syncthreads();

    const uint32_t octet = threadIdx.x >> 3, lane8 = threadIdx.x & 7u;
    constexpr uint32_t kOctets = kTripleThreads / 8;
    const uint32_t v0 = blockIdx.y * a.visitsPerCta, v1 = min(a.nVisits, v0 + a.visitsPerCta);
    unsigned long long entries = 0, visited = 0;

    uint32_t e = v0 + octet;
    uint32_t t = 0, key = 0, budget = 0, start = 0, end = 0;
    if (e < v1) {
        const uint32_t v = __ldg(a.visits + e);
        t = (v >> 24) & 15u; budget = v >> 28; key = sKey[t] ^ (v & 0xFFFFFFu);
        const uint32_t *o = a.tv.offs + (uint64_t)t * (kTripleBuckets + 1) + key;
        start = __ldg(o); end = __ldg(o + 1);
    }
    while (e < v1) {
        // request the next visit's offsets first
        const uint32_t en = e + kOctets;
        uint32_t tn = 0, keyn = 0, budgetn = 0, startn = 0, endn = 0;
        if (en < v1) {
            const uint32_t v = __ldg(a.visits + en);
            tn = (v >> 24) & 15u; budgetn = v >> 28; keyn = sKey[tn] ^ (v & 0xFFFFFFu);
            const uint32_t *o = a.tv.offs + (uint64_t)tn * (kTripleBuckets + 1) + keyn;
            startn = __ldg(o); endn = __ldg(o + 1);
        }
        if (start < end) {
            const uint32_t gg = sRes[t];
            const uint4 *__restrict__ base = reinterpret_cast<const uint4 *>(a.tv.res + (uint64_t)t * a.tv.stride);
            const uint32_t lastVec = (end - 1) >> 3;
            for (uint32_t vi = (start >> 3) + lane8; vi <= lastVec; vi += 8) {
                const uint4 r = __ldg(base + vi);
                const uint32_t w[4] = {r.x ^ gg, r.y ^ gg, r.z ^ gg, r.w ^ gg};
                int m = 16;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t f = (w[k] | (w[k] >> 1)) & 0x55555555u;
                    m = min(m, min(__popc(f & 0xFFFFu), __popc(f >> 16)));
                }
                if (m <= (int)budget) triple_slow(a, guide, g, t, key, vi << 3, start, end, r);
            }
            if (lane8 == 0) { entries += end - start; }
        }
        if (lane8 == 0) visited++;
        e = en; t = tn; key = keyn; budget = budgetn; start = startn; end = endn;
    }
    if (a.streamed) {
        if (entries) atomicAdd(&sCount[0], entries);
        if (visited) atomicAdd(&sCount[1], visited);
        __syncthreads();
        if (threadIdx.x < 2 && sCount[threadIdx.x]) atomicAdd(a.streamed + threadIdx.x, sCount[threadIdx.x]);
    }
}

// candidates of one wave = list entries the reference would visit (ref :330-344): the unit of work
__global__ void k_wave_candidates(IndexView iv, const uint64_t *guides, const uint8_t *done, uint32_t nGuides,
                                  uint32_t slice0, uint32_t nSlices, unsigned long long *total)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long len = 0;
    if (i < (uint64_t)nGuides * nSlices) {
        const uint32_t gi = (uint32_t)(i / nSlices), s = slice0 + (uint32_t)(i % nSlices);
        if (!done || !done[gi]) len = iv.listLen[pair_list(iv, guides[gi], s)];
    }
    for (int o = 16; o > 0; o >>= 1) len += __shfl_down_sync(0xffffffffu, len, o);
    if ((threadIdx.x & 31) == 0 && len) atomicAdd(total, len);
}

}  // namespace issl
