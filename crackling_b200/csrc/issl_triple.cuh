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
// (two entries): XOR with the guide's residual, fold to per-base flags, two POPC, two compares with the
// bucket's budget -> an 8-bit pass mask per vector.  The offsets of the next visit are requested before
// the current bucket is processed, so that two dependent round trips per octet are in flight.
//
// A passing entry is within maxDist of the guide (bucket mismatches + residual mismatches); what is
// left is the de-duplication -- the stateless replacement for the reference's toggle bitset (:385-390,
// :463): E = slices matching exactly (from the visit's pattern and the residual's two bytes), and the
// hit is kept only if this triple is resp(E), so each hit is produced exactly once.  Kept hits go to a
// shared-memory list (position, triple, min(E)); when the CTA is done it reserves a range of the global
// key buffer with one atomic, resolves the site ids with all threads and writes
// key = guide << 35 | min(E) << 32 | id, which sorts back into the reference's visiting order.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kTripleHitCap = 768;   // per CTA; further hits are emitted straight to the global buffer

struct TripleVisit {
    uint32_t x;   // pattern24 | triple << 24 | budget << 28   (issl_triple_visits)
    uint32_t y;   // exact slices inside the triple (5-bit set) | slice p << 8 | slice q << 12
};

struct TripleArgs {
    TripleView tv;
    const uint64_t *guides;
    const uint8_t *done;          // optional: guides that already left through the early exit
    const TripleVisit *visits;    // this wave's part of the table
    uint32_t nVisits, visitsPerCta;
    uint64_t *hitKeys;
    unsigned long long *hitCount;
    uint64_t hitCap;
    unsigned long long *streamed; // [0] entries of visited buckets, [1] bucket visits
    int maxDist;
};

// resp(E) packed 4 bits per E (E = 0 never occurs: every visit has an exact slice)
__host__ __device__ constexpr uint32_t triple_resp_of(uint32_t E)
{
    int pick[3] = {0, 0, 0}, np = 0;
    for (int s = 0; s < 5 && np < 3; s++) if (E & (1u << s)) pick[np++] = s;
    for (int s = 0; s < 5 && np < 3; s++) if (!(E & (1u << s))) pick[np++] = s;
    for (int a = 0; a < 2; a++)
        for (int b = 0; b < 2 - a; b++)
            if (pick[b] > pick[b + 1]) { const int t = pick[b]; pick[b] = pick[b + 1]; pick[b + 1] = t; }
    const int T[10][3] = {{0, 1, 2}, {0, 1, 3}, {0, 1, 4}, {0, 2, 3}, {0, 2, 4}, {0, 3, 4}, {1, 2, 3}, {1, 2, 4}, {1, 3, 4}, {2, 3, 4}};
    for (int t = 0; t < 10; t++)
        if (T[t][0] == pick[0] && T[t][1] == pick[1] && T[t][2] == pick[2]) return (uint32_t)t;
    return 15u;
}
__host__ __device__ constexpr uint64_t triple_resp_pack(uint32_t e0)
{
    uint64_t v = 0;
    for (uint32_t e = 0; e < 16; e++) v |= (uint64_t)triple_resp_of(e0 + e) << (4 * e);
    return v;
}
constexpr uint64_t kRespLo = triple_resp_pack(0), kRespHi = triple_resp_pack(16);

__global__ void __launch_bounds__(kTripleThreads) k_scan_triple(const TripleArgs a)
{
    const uint32_t guide = blockIdx.x;
    if (a.done && a.done[guide]) return;
    __shared__ uint32_t sKey[kTripleCount], sRes[kTripleCount];
    __shared__ unsigned long long sCount[2];
    __shared__ uint2 sHits[kTripleHitCap];
    __shared__ uint32_t sNHits;
    __shared__ unsigned long long sBase;
    const uint64_t g = a.guides[guide];
    if (threadIdx.x < kTripleCount) {
        const uint32_t t = threadIdx.x;
        sKey[t] = triple_key(g, c_tripleSlices[t][0], c_tripleSlices[t][1], c_tripleSlices[t][2]);
        sRes[t] = triple_res(g, c_tripleSlices[t][3], c_tripleSlices[t][4]) * 0x10001u;
    }
    if (threadIdx.x < 2) sCount[threadIdx.x] = 0;
    if (threadIdx.x == 0) sNHits = 0;
    __syncthreads();

    const uint32_t octet = threadIdx.x >> 3, lane8 = threadIdx.x & 7u;
    constexpr uint32_t kOctets = kTripleThreads / 8;
    const uint32_t v0 = blockIdx.y * a.visitsPerCta, v1 = min(a.nVisits, v0 + a.visitsPerCta);
    unsigned long long entries = 0, visited = 0;
    const uint2 *__restrict__ visits = reinterpret_cast<const uint2 *>(a.visits);

    uint32_t e = v0 + octet;
    uint2 v = make_uint2(0, 0);
    uint32_t start = 0, end = 0;
    if (e < v1) {
        v = __ldg(visits + e);
        const uint32_t t = (v.x >> 24) & 15u;
        const uint32_t *o = a.tv.offs + (uint64_t)t * (kTripleBuckets + 1) + (sKey[t] ^ (v.x & 0xFFFFFFu));
        start = __ldg(o); end = __ldg(o + 1);
    }
    while (e < v1) {
        // request the next visit's offsets first
        const uint32_t en = e + kOctets;
        uint2 vn = make_uint2(0, 0);
        uint32_t startn = 0, endn = 0;
        if (en < v1) {
            vn = __ldg(visits + en);
            const uint32_t tn = (vn.x >> 24) & 15u;
            const uint32_t *o = a.tv.offs + (uint64_t)tn * (kTripleBuckets + 1) + (sKey[tn] ^ (vn.x & 0xFFFFFFu));
            startn = __ldg(o); endn = __ldg(o + 1);
        }
        if (start < end) {
            const uint32_t t = (v.x >> 24) & 15u, budget = v.x >> 28;
            const uint32_t gg = sRes[t];
            const uint4 *__restrict__ base = reinterpret_cast<const uint4 *>(a.tv.res + (uint64_t)t * a.tv.stride);
            const uint32_t lastVec = (end - 1) >> 3;
            for (uint32_t vi = (start >> 3) + lane8; vi <= lastVec; vi += 8) {
                const uint4 r = __ldg(base + vi);
                const uint32_t w0 = r.x ^ gg, w1 = r.y ^ gg, w2 = r.z ^ gg, w3 = r.w ^ gg;
                const uint32_t f0 = w0 | (w0 >> 1), f1 = w1 | (w1 >> 1), f2 = w2 | (w2 >> 1), f3 = w3 | (w3 >> 1);
                uint32_t pass = 0;
                pass |= (uint32_t)(__popc(f0 & 0x5555u) <= (int)budget) << 0;
                pass |= (uint32_t)(__popc(f0 & 0x55550000u) <= (int)budget) << 1;
                pass |= (uint32_t)(__popc(f1 & 0x5555u) <= (int)budget) << 2;
                pass |= (uint32_t)(__popc(f1 & 0x55550000u) <= (int)budget) << 3;
                pass |= (uint32_t)(__popc(f2 & 0x5555u) <= (int)budget) << 4;
                pass |= (uint32_t)(__popc(f2 & 0x55550000u) <= (int)budget) << 5;
                pass |= (uint32_t)(__popc(f3 & 0x5555u) <= (int)budget) << 6;
                pass |= (uint32_t)(__popc(f3 & 0x55550000u) <= (int)budget) << 7;
                const uint32_t first = vi << 3;
                const uint32_t lo = start > first ? start - first : 0u, hi = min(end - first, 8u);
                pass &= ((1u << hi) - 1u) & ~((1u << lo) - 1u);
                while (pass) {
                    const uint32_t i = __ffs(pass) - 1;
                    pass &= pass - 1;
                    const uint32_t ws = (i & 4u) ? ((i & 2u) ? w3 : w2) : ((i & 2u) ? w1 : w0);
                    const uint32_t x16 = (ws >> ((i & 1u) * 16u)) & 0xFFFFu;
                    const uint32_t E = (v.y & 31u) | ((uint32_t)((x16 & 0xFFu) == 0) << ((v.y >> 8) & 7u)) |
                                       ((uint32_t)((x16 >> 8) == 0) << ((v.y >> 12) & 7u));
                    const uint32_t resp = (uint32_t)(((E & 16u) ? kRespHi : kRespLo) >> (4 * (E & 15u))) & 15u;
                    if (resp != t) continue;
                    const uint32_t minE = __ffs(E) - 1;
                    const uint32_t slot = atomicAdd(&sNHits, 1u);
                    if (slot < kTripleHitCap) {
                        sHits[slot] = make_uint2(first + i, t | (minE << 4));
                    } else {   // rare (dense repeat families): straight to the global buffer
                        const uint32_t id = a.tv.ids[(uint64_t)t * a.tv.stride + first + i];
                        const unsigned long long gs = atomicAdd(a.hitCount, 1ull);
                        if (gs < a.hitCap) a.hitKeys[gs] = ((uint64_t)guide << kTripleKeyBits) | ((uint64_t)minE << 32) | id;
                    }
                }
            }
            if (lane8 == 0) entries += end - start;
        }
        if (lane8 == 0) visited++;
        e = en; v = vn; start = startn; end = endn;
    }
    if (a.streamed) {
        if (entries) atomicAdd(&sCount[0], entries);
        if (visited) atomicAdd(&sCount[1], visited);
    }
    __syncthreads();
    const uint32_t nLocal = min(sNHits, kTripleHitCap);
    if (threadIdx.x == 0 && nLocal) sBase = atomicAdd(a.hitCount, (unsigned long long)nLocal);
    if (a.streamed && threadIdx.x < 2 && sCount[threadIdx.x]) atomicAdd(a.streamed + threadIdx.x, sCount[threadIdx.x]);
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < nLocal; j += kTripleThreads) {
        const uint2 h = sHits[j];
        const uint32_t t = h.y & 15u;
        const uint32_t id = __ldg(a.tv.ids + (uint64_t)t * a.tv.stride + h.x);
        const unsigned long long slot = sBase + j;
        if (slot < a.hitCap) a.hitKeys[slot] = ((uint64_t)guide << kTripleKeyBits) | ((uint64_t)(h.y >> 4) << 32) | id;
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
