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

#include <type_traits>

#include "issl_kernels.cuh"
#include "issl_triple_tables.h"

namespace issl {

constexpr uint32_t kTripleCount = 10;
constexpr uint32_t kTripleBuckets = 1u << 24;
constexpr int kTripleThreads = 128;              // per CTA = per guide
constexpr int kTripleKeyBits = 36;               // survivor key = guide << 36 | ordering slice << 32 | site id
constexpr uint32_t kOrderSlices = 10;            // ordering slices: 5 (sliceWidth 8) or 10 (sliceWidth 4)
constexpr uint32_t kNoSlice = 15;                // a record that is not a hit in its triple

// slices of triple t: key bytes 0..2 (L, M, H), then p < q (residual bytes 0..1) -- issl_triple_tables.h
__constant__ uint8_t c_tripleSlices[kTripleCount][5] = ISSL_TRIPLE_LAYOUT_INIT;

struct TripleView {
    const uint16_t *res;    // [10][stride] residual bits (slice p | slice q << 8) per bucket entry
    const uint32_t *ids;    // [10][stride] site id per bucket entry; with occFlag, bit 31 = "occurs more than once"
    uint32_t occFlag;       // 1 when offtargetsCount < 2^31, so that bit 31 of an id is free for that flag
    const uint32_t *offs;   // [10][2^24 + 1] first entry of every bucket
    uint64_t stride;        // entries reserved per triple (multiple of 8, >= N + 64)
    // blocked, bit-sliced copy of res (optional): bucket k of triple t owns pitch/32 sub-blocks of 64 bytes at
    // ((t << 24) | k) * pitch * 2 bytes.  A sub-block holds up to 31 residuals TRANSPOSED: word p (p = 0..15) is
    // bit p of the residuals of its 32 slots; slot 0 is not a residual: its column holds the number of
    // residuals in the sub-block (bits 0..4) and the flag "the bucket has more entries than its block holds: the
    // rest is in res/offs" (bit 5).  A bucket fills sub-block 0 first.  One aligned read per visit, no offset
    // lookup in front of it, and 31 residuals are tested with ~30 bitwise instructions.
    const uint4 *blk;
    uint32_t pitch;         // 16-bit slots per bucket: 0 (no blocked copy), 32, 64 or 128
    // small indexes (a bacterial genome: one site per twenty buckets): one bit per bucket, "holds at least one entry" --
    // 2 MB per triple, resident in L2 -- so that the contiguous scan only follows the offsets of the ~5 % of its visits
    // that can find anything.  nullptr on indexes with a blocked copy or with mostly occupied buckets.
    const uint32_t *nonEmpty;   // [10][2^24 / 32]
    // sliceWidth 4 (ten 2-base slices): with maxDist <= 4 every site within maxDist agrees with the guide on a whole
    // byte, so the same buckets are read; only the order differs -- the reference meets a hit first in the lowest
    // 2-base slice that matches exactly, which may lie below the lowest exact byte
    uint32_t nibbleOrder;
    // sliceWidth 10 (four 5-base slices): the builder truncates slice values to 8 bits (ref isslCreateIndex.cpp:228), so
    // list (i, v) holds the sites that agree with v on the slice's FIRST FOUR bases, and a guide only ever looks such a list
    // up when the fifth base of its slice is A (its 10-bit slice value is then below 256; the lists above are empty).  The
    // copies are therefore built from PERMUTED signatures -- bytes 0..3 = the four-base units of slices 0..3, byte 4 = the
    // four fifth bases (sig_perm10) -- and a guide keeps a hit only if site and guide agree exactly on a unit whose gate
    // (fifth base of the guide = A) is open; the lowest such unit is the slice the reference meets the hit in.
    uint32_t perm10;
    // Site ids are ranks in the text order of the sites (extractOfftargets sorts its output and isslCreateIndex numbers
    // distinct lines as it reads them, ref isslCreateIndex.cpp:184-207): when the index really is sorted (checked once
    // at load, k_check_site_order), "ascending id" -- the order the reference accumulates the hits of one slice in,
    // ref isslScoreOfftargets.cpp:330-344 -- IS ascending site text, which a hit's own 40 bits give: the fused tail then
    // orders hits by sig_to_sortkey(site) and never looks an id up (two dependent 64-byte lines per hit saved).
    // A site that occurs more than once still needs its count: the blocks keep such entries FIRST in their bucket and
    // note how many there are (see k_triple_blocks), so only those hits pay for the lookup.
    uint32_t siteOrdered;
};
constexpr uint32_t kSubEntries = 31;

__host__ __device__ __forceinline__ uint32_t triple_key(uint64_t sig, uint32_t a, uint32_t b, uint32_t c)
{
    return (uint32_t)((sig >> (8 * a)) & 0xFFull) | ((uint32_t)((sig >> (8 * b)) & 0xFFull) << 8) |
           ((uint32_t)((sig >> (8 * c)) & 0xFFull) << 16);
}
__host__ __device__ __forceinline__ uint32_t triple_res(uint64_t sig, uint32_t p, uint32_t q)
{
    return (uint32_t)((sig >> (8 * p)) & 0xFFull) | ((uint32_t)((sig >> (8 * q)) & 0xFFull) << 8);
}

// sliceWidth 10: signature -> five bytes (units of slices 0..3, then the four fifth bases) and back
__host__ __device__ __forceinline__ uint64_t sig_perm10(uint64_t sig)
{
    uint64_t r = 0;
    for (int i = 0; i < 4; i++) {
        r |= ((sig >> (10 * i)) & 0xFFull) << (8 * i);
        r |= ((sig >> (10 * i + 8)) & 3ull) << (32 + 2 * i);
    }
    return r;
}
__host__ __device__ __forceinline__ uint64_t sig_unperm10(uint64_t p)
{
    uint64_t r = 0;
    for (int i = 0; i < 4; i++) {
        r |= ((p >> (8 * i)) & 0xFFull) << (10 * i);
        r |= ((p >> (32 + 2 * i)) & 3ull) << (10 * i + 8);
    }
    return r;
}
// the units of a guide whose gate is open: fifth base of slice i = A
__host__ __device__ __forceinline__ uint32_t gates10(uint64_t g)
{
    uint32_t m = 0;
    for (int i = 0; i < 4; i++) m |= (uint32_t)(((g >> (10 * i + 8)) & 3ull) == 0) << i;
    return m;
}

// ------------------------------------------------------------------------------------------------
// construction (once per index): sort (key, id) per triple, then residuals + bucket offsets
// ------------------------------------------------------------------------------------------------
// sort key = bucket << 1 | "occurs once": inside a bucket the sites that occur more than once come first, ascending
// id within either class (the sort is stable)
__global__ void k_triple_keys(const uint64_t *sig, const uint32_t *occ, uint64_t n, uint32_t t, uint32_t perm10, uint32_t *keys, uint32_t *ids)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t s = perm10 ? sig_perm10(sig[i]) : sig[i];
    keys[i] = (triple_key(s, c_tripleSlices[t][0], c_tripleSlices[t][1], c_tripleSlices[t][2]) << 1) | (occ[i] > 1 ? 0u : 1u);
    ids[i] = (uint32_t)i;
}

// A site's rank key in text order (base 0 most significant, A < C < G < T): sig_to_sortkey without its loop -- reverse all
// bits, then put the two bits of every base back in order
__device__ __forceinline__ uint64_t site_text_key(uint64_t sig, uint32_t L)
{
    uint64_t r = __brevll(sig);
    r = ((r & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((r & 0x5555555555555555ull) << 1);
    return r >> (64 - 2 * L);
}

// violations of "sites are in ascending text order" (then ids are text ranks, see TripleView::siteOrdered)
__global__ void k_check_site_order(const uint64_t *sig, uint64_t n, uint32_t L, unsigned long long *violations)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= n) return;
    if (site_text_key(sig[i], L) >= site_text_key(sig[i + 1], L)) atomicAdd(violations, 1ull);
}

__global__ void k_triple_residuals(const uint64_t *sig, const uint32_t *sortedIds, uint64_t n, uint32_t t, uint32_t perm10, uint16_t *res)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t s = perm10 ? sig_perm10(sig[sortedIds[i]]) : sig[sortedIds[i]];
    res[i] = (uint16_t)triple_res(s, c_tripleSlices[t][3], c_tripleSlices[t][4]);
}

// bit 31 of every stored id = the site occurs more than once (saves the occurrence lookup for all other hits)
__global__ void k_triple_flag_ids(const uint32_t *occ, uint64_t n, uint32_t *ids)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t id = ids[i];
    if (occ[id] > 1) ids[i] = id | 0x80000000u;
}

// offs[k] = number of entries with key < k, for k in [0, 2^24]
__global__ void k_triple_offsets(const uint32_t *sortedKeys, uint64_t n, uint32_t *offs)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > kTripleBuckets) return;
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if ((sortedKeys[mid] >> 1) < k) lo = mid + 1; else hi = mid;
    }
    offs[k] = (uint32_t)lo;
}

// one bit per bucket: it holds at least one entry (TripleView::nonEmpty)
__global__ void k_triple_nonempty(const uint32_t *offs, uint32_t *bits)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;   // < 2^24
    const uint32_t b = __ballot_sync(0xffffffffu, offs[k + 1] > offs[k]);
    if ((threadIdx.x & 31u) == 0) bits[k >> 5] = b;
}

// blocked copy of one triple: one thread transposes one sub-block.  Column 0 of a sub-block (bit 0 of its 16 words) is
// not a residual: words 0..4 = number of entries, word 5 = "the bucket has more entries than its block holds",
// words 6..10 = how many of this sub-block's entries occur more than once -- they are its first ones.
__global__ void k_triple_blocks(const uint16_t *res, const uint32_t *offs, const uint32_t *sortedKeys, uint32_t subs,
                                uint4 *blk /* this triple's blocks */)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)kTripleBuckets * subs) return;
    const uint32_t key = (uint32_t)(i / subs), sub = (uint32_t)(i % subs);
    const uint32_t start = offs[key], count = offs[key + 1] - start;
    uint32_t w[16];
#pragma unroll
    for (int p = 0; p < 16; p++) w[p] = 0;
    // the block holds the bucket's first subs*31 entries; the flag says that more follow in the contiguous copy
    const uint32_t first = sub * kSubEntries;
    const uint32_t n = count > first ? min(count - first, kSubEntries) : 0u;
    uint32_t multi = 0;
    for (uint32_t e = 0; e < n; e++) {
        const uint32_t r = res[start + first + e];
#pragma unroll
        for (int p = 0; p < 16; p++) w[p] |= ((r >> p) & 1u) << (e + 1);
        multi += 1u - (sortedKeys[start + first + e] & 1u);
    }
#pragma unroll
    for (int p = 0; p < 5; p++) w[p] |= (n >> p) & 1u;
    if (count > subs * kSubEntries) w[5] |= 1u;
#pragma unroll
    for (int p = 0; p < 5; p++) w[6 + p] |= (multi >> p) & 1u;
    uint4 *o = blk + i * 4;
    o[0] = make_uint4(w[0], w[1], w[2], w[3]);
    o[1] = make_uint4(w[4], w[5], w[6], w[7]);
    o[2] = make_uint4(w[8], w[9], w[10], w[11]);
    o[3] = make_uint4(w[12], w[13], w[14], w[15]);
}

// ------------------------------------------------------------------------------------------------
// K1t: bucket scan.  ref isslScoreOfftargets.cpp:344-390 (+ :392-502 in the fused tail) for one guide,
// restricted to the buckets that can hold a site within maxDist.
//
// grid = (guides, visit chunks); one CTA = one guide x a range of the visit table.  Two kernels share the
// helpers below:
//   k_scan_triple_blocked  reads the blocked, bit-sliced copy: one aligned read per visit, 31 residuals
//                          per 64-byte sub-block compared with ~30 LOP3 (the default);
//   k_scan_triple          reads the contiguous copy through the bucket offsets: an octet of lanes per
//                          bucket, 16-byte vectors, XOR / fold / POPC per entry (indexes too small for
//                          blocks, or when the blocked copy does not fit the HBM).
// An entry within its bucket's budget is within maxDist of the guide (bucket mismatches + residual
// mismatches); what is left is the de-duplication -- the stateless replacement for the reference's toggle
// bitset (:385-390, :463): E = slices matching exactly (from the visit's pattern and the residual's two
// bytes), and the hit is kept only if this triple is resp(E), so each hit is produced exactly once.
// Kept hits go to a shared-memory list of 8-byte records; when the scan is done the CTA either finishes
// the guide itself (fused tail, score_guide), or hands the hits on: as a per-guide segment for
// k_score_segments, or as keys  guide << 35 | min(E) << 32 | id  for the general pipeline (radix sort ->
// k_contrib -> k_accumulate), which sort back into the reference's visiting order.
// ------------------------------------------------------------------------------------------------
// How the scan reads a 64-byte sub-block (ISSL_BLOCK_POLICY):
//   0  four 16-byte streaming loads (__ldcs: LDG.E.EF.128)
//   1  four 16-byte loads with L1::no_allocate
//   3..6  two 32-byte loads (sm_100 has LDG.256: ld.global.v8.b32) -- a lane then asks for each of its two sectors once, where
//      four 16-byte loads ask for each twice (the second request waits on the first: ncu's 50 % L1 "hit" rate of the scan)
//      3 plain, 4 L1::no_allocate + L2::evict_first, 5 L1::evict_first, 6 L1::no_allocate
// ISSL_VISIT_POLICY: 1 = the visit table is read with L1::evict_last
// Measured per 100 000 guides (profiles/r02_ab_ld256.jsonl): 0: 3.660 ms, 1: 3.955, 3: 3.579, 5: 3.541, 6: 3.473, 4: 3.457 (default)
#ifndef ISSL_BLOCK_POLICY
#define ISSL_BLOCK_POLICY 4
#endif
#ifndef ISSL_VISIT_POLICY
#define ISSL_VISIT_POLICY 0
#endif
__device__ __forceinline__ uint4 ld_block_na(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
#if ISSL_BLOCK_POLICY == 3
#define ISSL_LD256 "ld.global.v8.b32"
#elif ISSL_BLOCK_POLICY == 4
#define ISSL_LD256 "ld.global.L1::no_allocate.L2::evict_first.v8.b32"
#elif ISSL_BLOCK_POLICY == 5
#define ISSL_LD256 "ld.global.L1::evict_first.v8.b32"
#else
#define ISSL_LD256 "ld.global.L1::no_allocate.v8.b32"
#endif
__device__ __forceinline__ void ld_block_256(const uint4 *p, uint4 &a, uint4 &b)
{
    asm volatile(ISSL_LD256 " {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
// the four 16-byte words of the sub-block at p
__device__ __forceinline__ void load_sub_block(const uint4 *__restrict__ p, uint4 &q0, uint4 &q1, uint4 &q2, uint4 &q3)
{
#if ISSL_BLOCK_POLICY >= 3
    ld_block_256(p, q0, q1);
    ld_block_256(p + 2, q2, q3);
#elif ISSL_BLOCK_POLICY == 1
    q0 = ld_block_na(p); q1 = ld_block_na(p + 1); q2 = ld_block_na(p + 2); q3 = ld_block_na(p + 3);
#else
    q0 = __ldcs(p); q1 = __ldcs(p + 1); q2 = __ldcs(p + 2); q3 = __ldcs(p + 3);
#endif
}
__device__ __forceinline__ uint2 ld_visit(const uint2 *p)
{
#if ISSL_VISIT_POLICY == 1
    uint2 r;
    asm volatile("ld.global.L1::evict_last.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
#else
    return __ldg(p);
#endif
}
#ifndef ISSL_TRIPLE_MIN_CTAS
#define ISSL_TRIPLE_MIN_CTAS 10   // resident CTAs per SM the scan kernel is compiled for (48 registers); 8, 9, 12 measured slower
#endif
#ifndef ISSL_TRIPLE_MIN_CTAS_FLUSH
#define ISSL_TRIPLE_MIN_CTAS_FLUSH ISSL_TRIPLE_MIN_CTAS   // the flush variant (maxDist 5-6, repeat families) spills 40 bytes at 48 registers
#endif
#ifndef ISSL_SUB_NOEARLY
#define ISSL_SUB_NOEARLY 1        // 1: an empty sub-block is compared like any other (no branch; its validity mask is empty)
#endif
#ifndef ISSL_TRIPLE_PIPE
#define ISSL_TRIPLE_PIPE 0        // 1: the next round's visit-table entry is loaded one round ahead
#endif
#ifndef ISSL_TRIPLE_PREFETCH
#define ISSL_TRIPLE_PREFETCH 0
#endif
#ifndef ISSL_TRIPLE_HIT_CAP
#define ISSL_TRIPLE_HIT_CAP 512
#endif
constexpr uint32_t kTripleHitCap = ISSL_TRIPLE_HIT_CAP;   // per CTA; a guide with more hits goes through the general pipeline

// ------------------------------------------------------------------------------------------------
// Finishing one guide inside a CTA (used by the fused tail of the bucket scan and by k_score_segments):
// every hit is scored where it lies (ref :392-461); the accumulation order wanted is (slice, id) -- the
// reference's visiting order, ref :330-344 -- so hits are split by slice (5 or 10 groups, a counting pass), every
// thread counts the rank of its own hits inside their group (~55 ids per group on a uniform genome), the
// contributions pass through a window of shared memory in that order, and one thread adds them up one rounded
// sum at a time with the reference's early exit (ref :394, :460, :466-502).
// ------------------------------------------------------------------------------------------------
constexpr uint64_t kSiteUnknown = ~0ull;   // a hit record without the site's signature: look it up in sig[]

struct ScoreParams {
    const uint64_t *sig;          // [N] site signatures
    const uint32_t *occ;          // [N] occurrences
    uint64_t nSites;              // N
    uint32_t occFlag;             // ids carry "occurs more than once" in bit 31
    uint32_t keyShift;            // order keys that are ids: id >> keyShift lies in [0, kKeyBuckets]
    uint32_t byFirstMismatch;     // order keys are site text ranks: groups by the first mismatch with the guide (first_mismatch_group)
    ScoreTables tb;
    int calcMit, calcCfd, method, checkExit;
    double maximumSum;
    double *totMit, *totCfd;      // running sums, carried across slice waves
    uint8_t *done;
};

// Hits are put into the reference's accumulation order -- slice, then ascending site id inside the slice (ref :330-344) --
// by a counting pass over (slice, key range) groups followed by a rank inside the group.  The groups have to be small for
// the rank to be cheap, and leading key bits do not make them small: a guide's hits are its near neighbours, half of them
// share its first three bases (with 16 or 64 ranges of leading text bits this loop was 11-12 % of the kernel's instructions,
// profiles/r02_ncu_source_*).  On an index in text order the groups are therefore cut where the hits differ FROM THE GUIDE:
// f = position of a site's first mismatch, and whether its base there is below or above the guide's.  Sites below the guide
// sort before it, by ascending f (one that leaves the guide earlier, downwards, precedes one that still follows it); sites
// above it sort after it, by descending f: group = f | 20 (the guide itself) | 40 - f, monotone in text order, and the
// largest group holds a fifth of a slice's hits.  Indexes that are not in text order fall back to ranges of the id.
constexpr uint32_t kKeyBuckets = 41;
constexpr uint32_t kOrderGroups = kOrderSlices * kKeyBuckets;

__device__ __forceinline__ uint32_t first_mismatch_group(uint64_t site, uint64_t g)
{
    const uint64_t x = site ^ g;
    if (x == 0) return 20u;
    const uint32_t f = (uint32_t)(__ffsll((long long)x) - 1) >> 1;
    const uint32_t sb = (uint32_t)(site >> (2 * f)) & 3u, gb = (uint32_t)(g >> (2 * f)) & 3u;
    return sb < gb ? f : 40u - f;
}

struct ScoreShared {
    union {   // (shared memory per SM is L1 the scan cannot use: the two never live at the same time)
        struct { double mit[kTripleThreads], cfd[kTripleThreads]; };   // a window of contributions in accumulation order
        uint32_t grp[(kOrderGroups + 31) / 32 * 32];                   // per group: count -> first position -> end position
    };
    uint32_t kept;
};
constexpr uint32_t kScoreGroupWords = kTripleHitCap;   // one 64-bit order key per hit

// load(j, slice, site, occ, orderKey): hit j of the guide -- the slice through which the reference meets it first
// (kNoSlice: not a hit here), its signature, its occurrences, and a key that ascends as the site id does (the id itself,
// or the site's text rank).  group[] may alias whatever load() reads: it is
// first written after a barrier that follows the last load.  Results go to totMitOut/totCfdOut/doneOut[guide].
template <class Load>
__device__ __forceinline__ void score_guide(ScoreShared &ss, uint64_t *group, uint32_t n, uint32_t guide, uint64_t g,
                                            const ScoreParams &sp, double *totMitOut, double *totCfdOut, uint8_t *doneOut, Load load)
{
    constexpr uint32_t kPerThread = kTripleHitCap / kTripleThreads;
    for (uint32_t i = threadIdx.x; i < (kOrderGroups + 31) / 32 * 32; i += kTripleThreads) ss.grp[i] = 0;   // (padded for the prefix sum)
    __syncthreads();
    uint32_t myGroup[kPerThread];
    uint64_t myKey[kPerThread];
    double myMit[kPerThread], myCfd[kPerThread];   // shared memory per SM is L1 the scan cannot use: contributions stay in registers
#pragma unroll
    for (uint32_t k = 0; k < kPerThread; k++) {
        const uint32_t j = threadIdx.x + k * kTripleThreads;
        myGroup[k] = kOrderGroups;
        if (j < n) {
            uint32_t slice = kNoSlice, occ = 1;
            uint64_t site = 0, key = 0;
            load(j, slice, site, occ, key);
            if (slice >= kOrderSlices) continue;   // not a hit in this triple (the triple responsible for it reports it)
            myKey[k] = key;
            myGroup[k] = slice * kKeyBuckets + (sp.byFirstMismatch ? first_mismatch_group(site, g) : min((uint32_t)(key >> sp.keyShift), kKeyBuckets - 1));
            atomicAdd(&ss.grp[myGroup[k]], 1u);
            double cm, cc;
            int dist;
            hit_contrib(sp.tb, g, site, occ, sp.calcMit, sp.calcCfd, cm, cc, dist);
            myMit[k] = cm; myCfd[k] = cc;
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {   // counts -> first positions (exclusive prefix sum by one warp)
        constexpr uint32_t kPer = (kOrderGroups + 31) / 32;   // (the counts are read twice rather than kept in kPer registers)
        static_assert(kPer * 32 <= sizeof(ss.mit) / 4 + sizeof(ss.cfd) / 4, "grp[] is padded to a multiple of 32 inside the window's memory");
        uint32_t sum = 0;
#pragma unroll 4
        for (uint32_t i = 0; i < kPer; i++) sum += ss.grp[threadIdx.x * kPer + i];
        uint32_t incl = sum;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)threadIdx.x >= o) incl += up;
        }
        uint32_t run = incl - sum;
#pragma unroll 4
        for (uint32_t i = 0; i < kPer; i++) { const uint32_t c = ss.grp[threadIdx.x * kPer + i]; ss.grp[threadIdx.x * kPer + i] = run; run += c; }
        if (threadIdx.x == 31) ss.kept = incl;
    }
    __syncthreads();
#pragma unroll
    for (uint32_t k = 0; k < kPerThread; k++)
        if (myGroup[k] < kOrderGroups) group[atomicAdd(&ss.grp[myGroup[k]], 1u)] = myKey[k];   // keys, grouped; grp[] ends up as end positions
    __syncthreads();
    // rank of every hit inside its group = number of smaller keys (the sites of one guide's hits are distinct)
    uint32_t myRank[kPerThread];
#pragma unroll
    for (uint32_t k = 0; k < kPerThread; k++) {
        myRank[k] = 0xFFFFFFFFu;
        if (myGroup[k] < kOrderGroups) {
            const uint32_t base = myGroup[k] ? ss.grp[myGroup[k] - 1] : 0u, end = ss.grp[myGroup[k]];
            const uint64_t mine = myKey[k];
            uint32_t r = 0;
            for (uint32_t q = base; q < end; q++) r += (uint32_t)(group[q] < mine);
            myRank[k] = base + r;
        }
    }
    const uint32_t kept = ss.kept;
    __syncthreads();   // grp[] is done with: its memory becomes the window
    // ordered accumulation with the reference's early exit (ref :394, :460, :466-502): the contributions pass through
    // a window of shared memory in rank order, kTripleThreads at a time, and one thread adds them up
    double mit = 0.0, cfd = 0.0;
    bool stop = false;
    if (threadIdx.x == 0) { mit = sp.totMit[guide]; cfd = sp.totCfd[guide]; }
    for (uint32_t w0 = 0; w0 < kept; w0 += kTripleThreads) {
#pragma unroll
        for (uint32_t k = 0; k < kPerThread; k++)
            if (myRank[k] - w0 < (uint32_t)kTripleThreads) { ss.mit[myRank[k] - w0] = myMit[k]; ss.cfd[myRank[k] - w0] = myCfd[k]; }
        __syncthreads();
        if (threadIdx.x == 0 && !stop) {
            const uint32_t m = min((uint32_t)kTripleThreads, kept - w0);
            if (!sp.checkExit) {
#pragma unroll 8
                for (uint32_t i = 0; i < m; i++) { mit = __dadd_rn(mit, ss.mit[i]); cfd = __dadd_rn(cfd, ss.cfd[i]); }
            } else {
                for (uint32_t i = 0; i < m && !stop; i++) {
                    mit = __dadd_rn(mit, ss.mit[i]);
                    cfd = __dadd_rn(cfd, ss.cfd[i]);
                    stop = exit_predicate(sp.method, mit, cfd, sp.maximumSum);
                }
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        totMitOut[guide] = mit; totCfdOut[guide] = cfd;
        if (stop) doneOut[guide] = 1;
    }
}

// occurrences of a site whose id is known: 1 when the stored id says so, else from occ[]
__device__ __forceinline__ uint32_t occ_of(const ScoreParams &sp, uint32_t idRaw)
{
    return (sp.occFlag && !(idRaw & 0x80000000u)) ? 1u : __ldg(sp.occ + (idRaw & (sp.occFlag ? 0x7FFFFFFFu : ~0u)));
}

struct TripleVisit {
    uint32_t x;   // pattern24 | triple << 24 | budget << 28   (issl_triple_visits)
    uint32_t y;   // exact slices inside the triple (5-bit set) | slice p << 8 | slice q << 12 | keep table << 16:
                  // bit (pExact | qExact << 1) of the table = "an entry of this bucket whose residual matches exactly on
                  // slice p / q (or not) is this triple's to report" (resp(E) == triple), worked out once on the host
};

struct HeavyDesc;

struct TripleArgs {
    TripleView tv;
    const uint64_t *guides;
    const uint8_t *done;          // optional: guides that already left through the early exit
    const TripleVisit *visits;    // this wave's part of the table
    uint32_t nVisits, visitsPerCta;
    // survivors for the general pipeline (radix sort -> k_contrib -> k_accumulate): key = guide << 35 | slice << 32 | id
    uint64_t *hitKeys;
    unsigned long long *hitCount;
    uint64_t hitCap;
    unsigned long long *streamed; // [0] entries of visited buckets, [1] bucket visits
    int maxDist;
    // per-guide segments (one CTA per guide only, else nullptr): a guide with at most kTripleHitCap survivors
    // gets a contiguous range of segKeys/segSites and is finished there by k_score_segments; only the others
    // go through the general pipeline
    uint64_t *segKeys;            // slice << 32 | id (bit 31 of the id: occurrence flag when tv.occFlag)
    uint64_t *segSites;           // the site's signature; kSiteUnknown: look it up in sig[]
    unsigned long long *segCount;
    uint64_t segCap;
    uint64_t *segOff;             // [guides] first entry of the guide's segment
    uint32_t *segCnt;             // [guides] entries (pre-zeroed; stays 0 for guides sent to the general pipeline)
    // fused scoring (one CTA per guide only): a guide with at most kTripleHitCap survivors is sorted, scored and
    // accumulated by the CTA that found them, straight from shared memory; nothing of it reaches global memory
    int fuse;
    ScoreParams sp;               // sp.totMit/totCfd/done: state before this wave (read only)
    double *totMitOut, *totCfdOut;   // state after it, written for EVERY guide (a re-launch after a buffer overflow
    uint8_t *doneOut;                // must start from the same state)
    unsigned long long *fusedHits;
    unsigned long long *maxRecords;   // largest number of candidate records of one guide (diagnostics)
    // guides with more hits than a CTA's record list holds (maxDist 5-6, repeat families), fused mode on an index in text
    // order: the CTA keeps such a guide's hits as 64-bit sort keys in its own chunks of this buffer, sorts them there and
    // finishes the guide itself (heavy_finish) -- nullptr: such guides go through the general pipeline instead
    uint64_t *heavyKeys;
    unsigned long long *heavyCount;   // keys handed out so far (the launch is repeated with a larger buffer if > heavyCap)
    uint64_t heavyCap;
    unsigned long long *heavyHits;    // hits finished this way (sum of the guides' key counts: the size of k_heavy_finish's second buffer)
    HeavyDesc *heavyDesc;             // [guides] one per heavy guide, in order of completion
    unsigned long long *heavyGuides;
    // warp-per-guide kernel (k_scan_triple_small): guides it cannot finish (more hits than a warp's list holds, a long
    // bucket remainder: repeat families) are listed in redo[]; the CTA-per-guide kernel then runs for exactly those (guideList)
    // visits whose bucket overflows its block, beyond the kTripleOvfCap a CTA notes in shared memory: one bit per visit, one
    // row of ovfWords words per CTA (zeroed before the launch); nullptr: such visits are finished by the lane that met them
    uint32_t *ovfBits;
    uint32_t ovfWords;
    uint32_t nGuides;
    uint32_t *redo;
    unsigned long long *redoCount;
    const uint32_t *guideList;
};

// resp(E) packed 4 bits per E (E = 0 never occurs: every visit has an exact slice)
__host__ __device__ constexpr uint64_t triple_resp_pack(uint32_t e0)
{
    uint64_t v = 0;
    for (uint32_t e = 0; e < 16; e++) v |= (uint64_t)issl_triple_resp(e0 + e) << (4 * e);
    return v;
}
constexpr uint64_t kRespLo = triple_resp_pack(0), kRespHi = triple_resp_pack(16);

// buckets with more entries than their block holds (~1 % of the visits at human scale) are noted during the scan and
// finished by the whole CTA afterwards: two dependent loads in the middle of a warp's round stall the other 31 lanes
constexpr uint32_t kTripleOvfCap = 64;
constexpr uint32_t kTripleLongCap = 16;      // of those, buckets whose remainder is long enough for the whole CTA to share
constexpr uint32_t kTripleLongBucket = 512;  // entries
constexpr uint32_t kHeavyChunks = 23;        // 512 * (2^23 - 1) keys: more than any buffer holds

// CTA-wide state of one guide's scan
struct TripleShared {
    uint32_t key[kTripleCount], res[kTripleCount];   // the guide's bucket key / residual (both halves) per triple
    uint4 mask[kTripleCount][4];  // bit-sliced scan: word p = all ones when bit p of the guide's residual is set
    uint32_t count[2];            // entries / visits of this CTA (native 32-bit shared-memory atomics)
    union {
        uint2 hits[kTripleHitCap];    // candidate records (record_y)
        uint32_t radix[4][256];       // heavy_finish: per-warp digit counters of the radix sort (the records are keys by then)
    };
    // a heavy guide's sort keys live in chunks of TripleArgs::heavyKeys: chunk k holds 512 << k keys
    uint64_t chunkBase[kHeavyChunks];
    uint32_t heavyChunks, heavyLen, heavyBroken;
    uint32_t nLong;
    uint16_t longList[kTripleLongCap];
    uint32_t nHits, nKept;
    uint32_t guide;               // the guide's index (kept here, not in a register of every thread)
    uint32_t gated;               // slices a hit may count as exactly matching: all five, or (sliceWidth 10) the guide's open gates
    uint32_t flushed;             // the guide's records were flushed to the general pipeline's buffer at least once
    uint32_t nOvf;                // visits whose bucket has more entries than its block holds: noted, finished after the loop
    uint16_t ovf[kTripleOvfCap];
    unsigned long long base;
};

__device__ __forceinline__ void triple_prologue(const TripleArgs &a, TripleShared &sh, uint64_t gTrue, uint32_t guide)
{
    if (threadIdx.x == 33) sh.guide = guide;
    const uint64_t g = a.tv.perm10 ? sig_perm10(gTrue) : gTrue;
    if (threadIdx.x == 32) sh.gated = a.tv.perm10 ? gates10(gTrue) : 31u;
    if (threadIdx.x < kTripleCount) {
        const uint32_t t = threadIdx.x;
        sh.key[t] = triple_key(g, c_tripleSlices[t][0], c_tripleSlices[t][1], c_tripleSlices[t][2]);
        const uint32_t r = triple_res(g, c_tripleSlices[t][3], c_tripleSlices[t][4]);
        sh.res[t] = r * 0x10001u;
        uint32_t *m = reinterpret_cast<uint32_t *>(sh.mask[t]);
        for (int p = 0; p < 16; p++) m[p] = 0u - ((r >> p) & 1u);
    }
    if (threadIdx.x < 2) sh.count[threadIdx.x] = 0;
    if (threadIdx.x == 0) { sh.nHits = 0; sh.nKept = 0; sh.flushed = 0; sh.nOvf = 0; sh.nLong = 0; sh.heavyChunks = 0; sh.heavyLen = 0; sh.heavyBroken = 0; }
    __syncthreads();
}

// A hit record: an entry within its bucket's budget, i.e. a site within maxDist of the guide, found in the triple
// responsible for it (record_keep).  The scan only notes where it is and which of the two residual slices match
// exactly; id, site and scores are worked out later, for all records of the guide in parallel.
//   x: blocked scan: bucket key | (entry + 1) << 24;  contiguous scan: position in the triple's copy
//   y: triple (0..3) | exact slices of the visit's pattern (4..8) | residual matches on slice p (9), on q (10) |
//      blocked scan (11) | occurs more than once (12) | blocked scan: the entry's residual (16..31) -- with the bucket key,
//      the whole site.  (Leaving the residual out of the record and reading it back from the block in the tail, all lanes
//      busy, was measured: the scan loop loses 16 % of its instructions, yet 3.91 -> 4.12 ms per 100 000 guides at maxDist 4
//      and 28.5 -> 40.4 ms at maxDist 5 -- the tail's dependent loads stall the CTA while it requests no blocks;
//      profiles/r02_ab_block_load.jsonl.)
constexpr uint32_t kRecBlocked = 1u << 11;
constexpr uint32_t kRecMulti = 1u << 12;     // blocked scan: the site occurs more than once (its count has to be looked up)

__device__ __forceinline__ uint32_t record_y(uint2 v, uint32_t pExact, uint32_t qExact, uint32_t flag)
{
    return ((v.x >> 24) & 15u) | ((v.y & 31u) << 4) | (pExact << 9) | (qExact << 10) | flag;
}

// E = slices on which site and guide agree exactly; the record is a hit only in the triple responsible for E
// (gated: the slices that count -- sliceWidth 10: units whose gate is open; the hit is then met in the lowest of THOSE)
__device__ __forceinline__ bool record_keep(uint2 h, uint32_t &minE, uint32_t gated = 31u)
{
    const uint32_t t = h.y & 15u;
    const uint32_t E = ((h.y >> 4) & 31u) | (((h.y >> 9) & 1u) << c_tripleSlices[t][3]) | (((h.y >> 10) & 1u) << c_tripleSlices[t][4]);
    const uint32_t resp = (uint32_t)(((E & 16u) ? kRespHi : kRespLo) >> (4 * (E & 15u))) & 15u;
    minE = __ffs(E & gated) - 1;
    // E == 0: sliceWidth 4 only -- no byte matches exactly (order_slice finds the 2-base slice); such entries are triple 0's
    if (E == 0) { minE = 5; return t == 0; }
    return resp == t && (E & gated) != 0;
}

// key of the general pipeline: guide << 36 | slice << 32 | id, plus -- outside the bits that are sorted -- "occurs
// once" when the stored id says so, which saves k_contrib the occurrence lookup.  The slice is the lowest exact BYTE;
// for sliceWidth 4 k_fix_order_slices turns it into the lowest exact 2-base slice before the keys are sorted.
__device__ __forceinline__ uint64_t general_key(const TripleView &tv, uint32_t guide, uint32_t minE, uint32_t idRaw)
{
    const uint64_t once = (tv.occFlag && !(idRaw & 0x80000000u)) ? kKeyOccursOnce : 0ull;
    return once | ((uint64_t)guide << kTripleKeyBits) | ((uint64_t)minE << 32) | (idRaw & (tv.occFlag ? 0x7FFFFFFFu : ~0u));
}

__device__ __forceinline__ uint32_t hit_position(const TripleView &tv, uint2 h)
{
    if (!(h.y & kRecBlocked)) return h.x;
    return __ldg(tv.offs + (uint64_t)(h.y & 15u) * (kTripleBuckets + 1) + (h.x & 0xFFFFFFu)) + (h.x >> 24) - 1u;
}

// site of a record of the blocked scan: bucket key + the entry's residual
__device__ __forceinline__ uint64_t hit_site(const TripleView &tv, uint2 h)
{
    if (!(h.y & kRecBlocked)) return kSiteUnknown;
    const uint32_t t = h.y & 15u, key = h.x & 0xFFFFFFu, r = h.y >> 16;
    const uint64_t s = ((uint64_t)(key & 0xFFu) << (8 * c_tripleSlices[t][0])) | ((uint64_t)((key >> 8) & 0xFFu) << (8 * c_tripleSlices[t][1])) |
                       ((uint64_t)(key >> 16) << (8 * c_tripleSlices[t][2])) | ((uint64_t)(r & 0xFFu) << (8 * c_tripleSlices[t][3])) |
                       ((uint64_t)(r >> 8) << (8 * c_tripleSlices[t][4]));
    return tv.perm10 ? sig_unperm10(s) : s;
}

// the slice through which the reference meets a hit first (ref isslScoreOfftargets.cpp:330-390): the lowest exactly
// matching slice -- a byte for sliceWidth 8, a 2-base nibble for sliceWidth 4
__device__ __forceinline__ uint32_t order_slice(const TripleView &tv, uint64_t siteXorGuide, uint32_t minExactByte)
{
    if (!tv.nibbleOrder) return minExactByte;
    uint32_t s = 0;
    while (s < 2 * minExactByte && ((siteXorGuide >> (4 * s)) & 15ull) != 0) s++;   // the exact byte is two exact nibbles
    return s;
}

// ... for a record whose id is known (general-pipeline keys, segments)
__device__ __forceinline__ uint32_t record_order_slice(const TripleView &tv, const uint64_t *sig, uint2 h, uint64_t g, uint32_t idRaw,
                                                       uint32_t minExactByte)
{
    if (!tv.nibbleOrder) return minExactByte;
    uint64_t site = hit_site(tv, h);
    if (site == kSiteUnknown) site = __ldg(sig + (idRaw & (tv.occFlag ? 0x7FFFFFFFu : ~0u)));
    return order_slice(tv, site ^ g, minExactByte);
}

template <bool CHECKED = false>
__device__ __forceinline__ void triple_push(const TripleArgs &a, TripleShared &sh, uint32_t guide, uint2 h)
{
    uint32_t slice;
    if (!CHECKED && !record_keep(h, slice, sh.gated)) return;   // found again, and reported, through the triple responsible for it
    const uint32_t slot = atomicAdd(&sh.nHits, 1u);
    if (slot < kTripleHitCap) {
        sh.hits[slot] = h;
    } else {   // rare (dense repeat families): straight to the general pipeline's buffer
        uint32_t minE;
        if (!record_keep(h, minE, sh.gated)) return;
        const uint32_t id = a.tv.ids[(uint64_t)(h.y & 15u) * a.tv.stride + hit_position(a.tv, h)];
        const unsigned long long gs = atomicAdd(a.hitCount, 1ull);
        if (gs < a.hitCap) a.hitKeys[gs] = general_key(a.tv, guide, minE, id);
    }
}

// a record that found the CTA's list full (rare: dense repeat families): straight to the general pipeline's buffer
__device__ __forceinline__ void triple_spill(const TripleArgs &a, uint32_t guide, uint2 h, uint32_t gated)
{
    uint32_t minE;
    if (!record_keep(h, minE, gated)) return;
    const uint32_t id = a.tv.ids[(uint64_t)(h.y & 15u) * a.tv.stride + hit_position(a.tv, h)];
    const unsigned long long gs = atomicAdd(a.hitCount, 1ull);
    if (gs < a.hitCap) a.hitKeys[gs] = general_key(a.tv, guide, minE, id);
}

// Guides with thousands of hits (maxDist 5 and 6, dense repeat families): when the CTA's record list is nearly full,
// all threads empty it into the general pipeline's key buffer (one reservation, ids resolved in parallel) and the
// scan goes on; such a guide is finished by the general pipeline.  Called by all threads of the CTA.
constexpr uint32_t kTripleFlushAt = kTripleHitCap - 128;

__device__ __forceinline__ void triple_flush(const TripleArgs &a, TripleShared &sh, uint32_t guide)
{
    const uint32_t n = min(sh.nHits, kTripleHitCap);
    if (threadIdx.x == 0) sh.base = atomicAdd(a.hitCount, (unsigned long long)n);
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < n; j += kTripleThreads) {
        const uint2 h = sh.hits[j];
        uint32_t minE;
        record_keep(h, minE, sh.gated);
        const uint32_t id = __ldg(a.tv.ids + (uint64_t)(h.y & 15u) * a.tv.stride + hit_position(a.tv, h));
        const unsigned long long slot = sh.base + j;
        if (slot < a.hitCap) a.hitKeys[slot] = general_key(a.tv, guide, minE, id);
    }
    __syncthreads();
    if (threadIdx.x == 0) { sh.nHits = 0; sh.flushed = 1; }
    __syncthreads();
}

// Shared memory of the scan kernels: the scan state; the grouping arrays of the fused tail reuse it once every
// thread has taken its hits out.
struct TripleSmem {
    union {
        TripleShared scan;
        uint64_t group[kScoreGroupWords];
    };
    ScoreShared score;
};
// ... without the fused tail: the scan state alone.  Less shared memory per SM is more L1, and the L1 holds the lines
// of the loads in flight: the scan is measurably faster with it.
struct TripleSmemScan {
    TripleShared scan;
};

// ------------------------------------------------------------------------------------------------
// Heavy guides (more hits than the CTA's record list holds: maxDist 5-6, dense repeat families), finished
// inside the scan kernel.  ref isslScoreOfftargets.cpp:330-344 fixes the accumulation order -- slice, then
// ascending site id, which on an index in text order is ascending site text -- and :466-502 makes the printed
// value depend on it, so the guide's hits have to be sorted before they are added up.  Whenever the record list
// fills, all threads turn the records into 64-bit keys
//     ordering slice << 60 | site text rank (40 bits) << 20 | occurrences (20 bits, saturating)
// and append them to the guide's own chunks of a global buffer (chunk k holds 512 << k keys, handed out by one
// atomic per chunk); when the scan is done the CTA sorts them with a least-significant-digit radix sort over the
// 44 ordering bits (six stable passes, ping-pong between the chunks and one contiguous range; each warp owns a
// quarter of the keys and its own digit counters, so a pass costs five barriers whatever the number of keys;
// passes in which all keys share the digit are skipped) and then walks them in order: 128 keys at a time are
// turned back into sites and scored by all threads (ref :392-461) and one thread adds the contributions up one
// rounded sum at a time with the reference's early exit.  No id is ever looked up (except for the occurrence
// count of the few sites that occur more than once) and there is no device-wide sort: the scan kernel leaves one
// descriptor per heavy guide, and k_heavy_finish gives each of them a CTA.
// ------------------------------------------------------------------------------------------------
// what the scan kernel leaves behind for a heavy guide, and what k_heavy_finish works from
struct HeavyDesc {
    uint32_t guide, len, pad0, pad1;
    uint64_t chunkBase[kHeavyChunks + 1];
};
struct HeavyArgs {
    const HeavyDesc *desc;
    uint64_t *keys;                   // the scan's key buffer (chunks)
    uint64_t *flat;                   // one more range per guide, the other half of the sort's ping-pong: sum(len) keys in all
    unsigned long long *flatCount;
    const uint64_t *guides;
    ScoreParams sp;                   // sp.totMit / totCfd: state before this wave
    double *totMitOut, *totCfdOut;    // state after it
    uint8_t *doneOut;
    uint32_t fineGroups;              // 161 counting-sort groups per ordering slice (at most six slices) instead of 41
};
struct HeavySmem {
    uint32_t radix[4][256];
    uint64_t chunkBase[kHeavyChunks + 1];
    double mit[kTripleThreads], cfd[kTripleThreads];
    unsigned long long base;
    uint32_t warpTot[4];
    uint32_t stop;
};

constexpr uint32_t kHeavyOccMax = 0xFFFFFu;   // occurrence counts from here on are looked up again when the hit is scored

__device__ __forceinline__ uint64_t heavy_key(uint32_t slice, uint64_t site, uint32_t occ)
{
    return ((uint64_t)slice << 60) | (site_text_key(site, 20) << 20) | (uint64_t)min(occ, kHeavyOccMax);
}

// inverse of site_text_key for 20-base sites
__device__ __forceinline__ uint64_t text_key_site(uint64_t textKey)
{
    uint64_t r = textKey << 24;
    r = ((r & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((r & 0x5555555555555555ull) << 1);
    return __brevll(r);
}

// where key number L of the guide lives
__device__ __forceinline__ uint64_t heavy_slot(const uint64_t *chunkBase, uint32_t L)
{
    const uint32_t unit = (L >> 9) + 1u, k = 31u - (uint32_t)__clz(unit);
    return chunkBase[k] + (L - (((1u << k) - 1u) << 9));
}

// room for the guide's keys [0, newLen): called by all threads
__device__ __forceinline__ void heavy_reserve(const TripleArgs &a, TripleShared &sh, uint32_t newLen)
{
    if (threadIdx.x == 0) {
        while ((((1u << sh.heavyChunks) - 1u) << 9) < newLen && sh.heavyChunks < kHeavyChunks) {
            const uint32_t k = sh.heavyChunks;
            const unsigned long long base = atomicAdd(a.heavyCount, 512ull << k);
            if (base + (512ull << k) > a.heavyCap) sh.heavyBroken = 1;   // the launch is repeated with a larger buffer
            sh.chunkBase[k] = base;
            sh.heavyChunks = k + 1;
        }
    }
    __syncthreads();
}

// site, occurrences and ordering slice of a kept record
__device__ __forceinline__ void record_resolve(const TripleArgs &a, uint2 h, uint64_t g, uint32_t gated, uint32_t &slice, uint64_t &site,
                                               uint32_t &occ)
{
    record_keep(h, slice, gated);
    site = hit_site(a.tv, h);
    occ = 1;
    if (site == kSiteUnknown) {
        const uint32_t idRaw = __ldg(a.tv.ids + (uint64_t)(h.y & 15u) * a.tv.stride + hit_position(a.tv, h));
        site = __ldg(a.sp.sig + (idRaw & (a.tv.occFlag ? 0x7FFFFFFFu : ~0u)));
        occ = occ_of(a.sp, idRaw);
    } else if (h.y & kRecMulti) {
        occ = occ_of(a.sp, __ldg(a.tv.ids + (uint64_t)(h.y & 15u) * a.tv.stride + hit_position(a.tv, h)));
    }
    if (a.tv.nibbleOrder) slice = order_slice(a.tv, site ^ g, slice);
}

// the record list becomes keys at the end of the guide's chunks.  Called by all threads.
__device__ __forceinline__ void heavy_flush(const TripleArgs &a, TripleShared &sh, uint64_t g)
{
    const uint32_t n = min(sh.nHits, kTripleHitCap), len = sh.heavyLen;
    heavy_reserve(a, sh, len + n);
    if (!sh.heavyBroken)
        for (uint32_t j = threadIdx.x; j < n; j += kTripleThreads) {
            uint32_t slice, occ;
            uint64_t site;
            record_resolve(a, sh.hits[j], g, sh.gated, slice, site, occ);
            a.heavyKeys[heavy_slot(sh.chunkBase, len + j)] = heavy_key(slice, site, occ);
        }
    __syncthreads();
    if (threadIdx.x == 0) { sh.heavyLen = len + n; sh.nHits = 0; sh.flushed = 1; }
    __syncthreads();
}

// occurrences of a site known only by its text: its id is its rank in sig[] (the index is in text order)
__device__ __forceinline__ uint32_t occ_by_text(const ScoreParams &sp, uint64_t textKey)
{
    uint64_t lo = 0, hi = sp.nSites;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (site_text_key(__ldg(sp.sig + mid), 20) < textKey) lo = mid + 1; else hi = mid;
    }
    return lo < sp.nSites ? __ldg(sp.occ + lo) : 1u;
}

// K2h: sort a heavy guide's keys and finish the guide (see above): one CTA per heavy guide, launched after the scan.  (Doing
// this at the end of the scan CTA itself was measured: 40.3 ms per 100 000 guides at maxDist 5 instead of 30.6 ms for the
// scan alone -- a chain of dependent L2 round trips during which the CTA requests no blocks; on its own, with a dozen
// guides per SM in flight, the same work hides behind itself.)
//   Two sorts.  Keys spread over many (slice, first mismatch with the guide) groups take a counting sort over up to 1 024
// groups followed by a rank inside the group (a handful of comparisons per key): three sweeps over the keys.  When a group is
// large (exact copies aside, a family whose members all leave the guide at the same base) the rank would be quadratic, and
// the keys take the radix sort: six stable passes, data-independent.  Either way every sweep requests four
// keys per thread before it uses the first: the keys sit in L2, and a sweep is bound by that round trip, not by arithmetic.
constexpr uint32_t kHeavyRankMax = 48;   // largest group the counting sort finishes by comparisons

__global__ void __launch_bounds__(kTripleThreads, 12) k_heavy_finish(const HeavyArgs a)
{
    __shared__ HeavySmem sh;
    const HeavyDesc &hd = a.desc[blockIdx.x];
    const uint32_t guide = hd.guide, n = hd.len;
    const uint64_t g = a.guides[guide];
    if (threadIdx.x <= kHeavyChunks) sh.chunkBase[threadIdx.x] = hd.chunkBase[threadIdx.x];
    if (threadIdx.x == 0) { sh.base = atomicAdd(a.flatCount, (unsigned long long)n); sh.stop = 0; }
    __syncthreads();
    uint64_t *const flat = a.flat + sh.base;
    uint64_t *const chunks = a.keys;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    uint32_t *const cnt = &sh.radix[0][0];
    auto load = [&](uint32_t c, uint32_t L) { return c ? flat[L] : chunks[heavy_slot(sh.chunkBase, L)]; };
    // groups of the counting sort: (slice, where and how the site first leaves the guide) -- monotone in (slice, site text),
    // like first_mismatch_group but split once more by the site's base at that position: 161 groups per slice, the largest
    // holding ~7 % of a slice's hits.  (Leading text bits do not spread a guide's hits: they are its near neighbours.)
    // sliceWidth 4 has ten ordering slices: 41 groups per slice there.
    const uint32_t fine = a.fineGroups;
    auto group_of = [&](uint64_t key) -> uint32_t {
        const uint64_t site = text_key_site((key >> 20) & 0xFFFFFFFFFFull), x = site ^ g;
        const uint32_t slice = (uint32_t)(key >> 60);
        if (!fine) return slice * kKeyBuckets + first_mismatch_group(site, g);
        if (x == 0) return slice * 161u + 80u;
        const uint32_t f = (uint32_t)(__ffsll((long long)x) - 1) >> 1;
        const uint32_t sb = (uint32_t)(site >> (2 * f)) & 3u, gb = (uint32_t)(g >> (2 * f)) & 3u;
        return slice * 161u + (sb < gb ? f * 4u + sb : 81u + (19u - f) * 4u + sb);
    };
    uint32_t cur = 0;                                                    // 0: the keys are in the chunks, 1: in `flat`

    // ---- counting sort over (slice, six leading text bits), if no group is large
    for (uint32_t i = threadIdx.x; i < 1024; i += kTripleThreads) cnt[i] = 0;
    __syncthreads();
    for (uint32_t j0 = threadIdx.x; j0 < n; j0 += 4 * kTripleThreads) {
        uint64_t k[4];
#pragma unroll
        for (int u = 0; u < 4; u++) k[u] = j0 + u * kTripleThreads < n ? load(0, j0 + u * kTripleThreads) : 0ull;
#pragma unroll
        for (int u = 0; u < 4; u++) if (j0 + u * kTripleThreads < n) atomicAdd(&cnt[group_of(k[u])], 1u);
    }
    __syncthreads();
    bool ranked;
    {
        uint32_t c[8], tot = 0, mx = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) { c[q] = cnt[8 * threadIdx.x + q]; tot += c[q]; mx = max(mx, c[q]); }
        uint32_t incl = tot;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += up;
        }
        if (lane == 31) sh.warpTot[warp] = incl;
        ranked = !__syncthreads_or(mx > kHeavyRankMax);
        if (ranked) {
            uint32_t run = incl - tot;
            for (uint32_t w = 0; w < warp; w++) run += sh.warpTot[w];
#pragma unroll
            for (int q = 0; q < 8; q++) { cnt[8 * threadIdx.x + q] = run; run += c[q]; }   // first position of every group
        }
        __syncthreads();
    }
    if (ranked) {
        for (uint32_t j0 = threadIdx.x; j0 < n; j0 += 4 * kTripleThreads) {   // chunks -> flat, grouped; cnt[] ends up as end positions
            uint64_t k[4];
#pragma unroll
            for (int u = 0; u < 4; u++) k[u] = j0 + u * kTripleThreads < n ? load(0, j0 + u * kTripleThreads) : 0ull;
#pragma unroll
            for (int u = 0; u < 4; u++) if (j0 + u * kTripleThreads < n) flat[atomicAdd(&cnt[group_of(k[u])], 1u)] = k[u];
        }
        __syncthreads();
        for (uint32_t j0 = threadIdx.x; j0 < n; j0 += 2 * kTripleThreads) {   // flat -> chunks, every key at its rank
            uint64_t k[2];
            uint32_t base[2], end[2], r[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const uint32_t j = j0 + u * kTripleThreads;
                k[u] = j < n ? flat[j] : 0ull;
                const uint32_t grp = group_of(k[u]);
                base[u] = (j < n && grp) ? cnt[grp - 1] : 0u;
                end[u] = j < n ? cnt[grp] : 0u;
                r[u] = 0;
            }
#pragma unroll
            for (int u = 0; u < 2; u++)
                for (uint32_t q = base[u]; q < end[u]; q++) r[u] += (uint32_t)(flat[q] < k[u]);
#pragma unroll
            for (int u = 0; u < 2; u++)
                if (j0 + u * kTripleThreads < n) chunks[heavy_slot(sh.chunkBase, base[u] + r[u])] = k[u];
        }
        __syncthreads();
    } else {
        // ---- radix sort: each warp owns a quarter of the keys and its own digit counters
        const uint32_t quarter = ((n + 127u) >> 7) << 5;                     // keys per warp, a multiple of 32
        const uint32_t lo = min(n, warp * quarter), hi = min(n, lo + quarter);
        for (int pass = 0; pass < 6; pass++) {
            const uint32_t shift = 20u + 8u * pass;
            for (uint32_t i = threadIdx.x; i < 1024; i += kTripleThreads) cnt[i] = 0;
            __syncthreads();
            for (uint32_t L0 = lo; L0 < hi; L0 += 128) {                     // digit counts of this warp's quarter
                uint64_t k[4];
#pragma unroll
                for (int u = 0; u < 4; u++) k[u] = L0 + 32 * u + lane < hi ? load(cur, L0 + 32 * u + lane) : 0ull;
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const bool valid = L0 + 32 * u + lane < hi;
                    const uint32_t d = valid ? (uint32_t)(k[u] >> shift) & 0xFFu : 256u;
                    const uint32_t peers = __match_any_sync(0xffffffffu, d);
                    if (valid && lane == (uint32_t)__ffs(peers) - 1u) sh.radix[warp][d] += (uint32_t)__popc(peers);
                    __syncwarp();
                }
            }
            __syncthreads();
            // exclusive scan over (digit, warp): thread t owns digits 2t and 2t + 1
            uint32_t c[2][4], tot = 0;
#pragma unroll
            for (int q = 0; q < 2; q++)
#pragma unroll
                for (int w = 0; w < 4; w++) { c[q][w] = sh.radix[w][2 * threadIdx.x + q]; tot += c[q][w]; }
            const bool single = (c[0][0] + c[0][1] + c[0][2] + c[0][3] == n) || (c[1][0] + c[1][1] + c[1][2] + c[1][3] == n);
            uint32_t incl = tot;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)lane >= o) incl += up;
            }
            if (lane == 31) sh.warpTot[warp] = incl;
            if (__syncthreads_or(single)) continue;                          // all keys share this digit: nothing to move
            uint32_t run = incl - tot;
            for (uint32_t w = 0; w < warp; w++) run += sh.warpTot[w];
#pragma unroll
            for (int q = 0; q < 2; q++)
#pragma unroll
                for (int w = 0; w < 4; w++) { sh.radix[w][2 * threadIdx.x + q] = run; run += c[q][w]; }
            __syncthreads();
            for (uint32_t L0 = lo; L0 < hi; L0 += 128) {                     // stable scatter, 32 keys of the quarter at a time
                uint64_t k[4];
#pragma unroll
                for (int u = 0; u < 4; u++) k[u] = L0 + 32 * u + lane < hi ? load(cur, L0 + 32 * u + lane) : 0ull;
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const bool valid = L0 + 32 * u + lane < hi;
                    const uint32_t d = valid ? (uint32_t)(k[u] >> shift) & 0xFFu : 256u;
                    const uint32_t peers = __match_any_sync(0xffffffffu, d);
                    const uint32_t rank = (uint32_t)__popc(peers & ((1u << lane) - 1u));
                    uint32_t dst = 0;
                    if (valid) dst = sh.radix[warp][d];
                    __syncwarp();
                    if (valid && rank == 0) sh.radix[warp][d] = dst + (uint32_t)__popc(peers);
                    __syncwarp();
                    if (valid) {
                        if (cur) chunks[heavy_slot(sh.chunkBase, dst + rank)] = k[u]; else flat[dst + rank] = k[u];
                    }
                }
            }
            __syncthreads();
            cur ^= 1u;
        }
    }
    // ordered accumulation (ref :394, :460, :466-502)
    double mit = 0.0, cfd = 0.0;
    bool stop = false;
    if (threadIdx.x == 0) { mit = a.sp.totMit[guide]; cfd = a.sp.totCfd[guide]; }
    for (uint32_t w0 = 0; w0 < n; w0 += kTripleThreads) {
        const uint32_t j = w0 + threadIdx.x;
        if (j < n) {
            const uint64_t key = load(cur, j);
            const uint64_t textKey = (key >> 20) & 0xFFFFFFFFFFull;
            uint32_t occ = (uint32_t)key & kHeavyOccMax;
            if (occ == kHeavyOccMax) occ = occ_by_text(a.sp, textKey);
            double cm, cc;
            int dist;
            hit_contrib(a.sp.tb, g, text_key_site(textKey), occ, a.sp.calcMit, a.sp.calcCfd, cm, cc, dist);
            sh.mit[threadIdx.x] = cm; sh.cfd[threadIdx.x] = cc;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t m = min((uint32_t)kTripleThreads, n - w0);
            if (!a.sp.checkExit) {
#pragma unroll 8
                for (uint32_t i = 0; i < m; i++) { mit = __dadd_rn(mit, sh.mit[i]); cfd = __dadd_rn(cfd, sh.cfd[i]); }
            } else {
                for (uint32_t i = 0; i < m && !stop; i++) {
                    mit = __dadd_rn(mit, sh.mit[i]);
                    cfd = __dadd_rn(cfd, sh.cfd[i]);
                    stop = exit_predicate(a.sp.method, mit, cfd, a.sp.maximumSum);
                }
                if (stop) sh.stop = 1;
            }
        }
        __syncthreads();
        if (sh.stop) break;
    }
    if (threadIdx.x == 0) {
        a.totMitOut[guide] = mit; a.totCfdOut[guide] = cfd;
        if (stop) a.doneOut[guide] = 1;
    }
}

// 8 residuals of one 16-byte vector against the guide's; `valid` masks the slots that belong to the bucket.
// NOSPILL (the scan variant that flushes full record lists): a hit that finds the list full is not sent to the general
// pipeline; its slot is returned, and the caller tries again once the list has been flushed.
template <bool NOSPILL = false, class RecX>
__device__ __forceinline__ uint32_t triple_vector(const TripleArgs &a, TripleShared &sh, uint32_t guide, uint2 v, uint32_t gg,
                                                  const uint4 &r, uint32_t valid, RecX recX, uint32_t recFlag,
                                                  uint32_t *nHits = nullptr, uint2 *hits = nullptr, uint32_t cap = 0, uint32_t gated = 31u)
{
    // NOSPILL: the record list is (nHits, hits, cap); otherwise the CTA's, through triple_push
    const int budget = (int)(v.x >> 28);
    const uint32_t w0 = r.x ^ gg, w1 = r.y ^ gg, w2 = r.z ^ gg, w3 = r.w ^ gg;
    const uint32_t f0 = w0 | (w0 >> 1), f1 = w1 | (w1 >> 1), f2 = w2 | (w2 >> 1), f3 = w3 | (w3 >> 1);
    uint32_t pass = 0;
    pass |= (uint32_t)(__popc(f0 & 0x5555u) <= budget) << 0;
    pass |= (uint32_t)(__popc(f0 & 0x55550000u) <= budget) << 1;
    pass |= (uint32_t)(__popc(f1 & 0x5555u) <= budget) << 2;
    pass |= (uint32_t)(__popc(f1 & 0x55550000u) <= budget) << 3;
    pass |= (uint32_t)(__popc(f2 & 0x5555u) <= budget) << 4;
    pass |= (uint32_t)(__popc(f2 & 0x55550000u) <= budget) << 5;
    pass |= (uint32_t)(__popc(f3 & 0x5555u) <= budget) << 6;
    pass |= (uint32_t)(__popc(f3 & 0x55550000u) <= budget) << 7;
    pass &= valid;
    while (pass) {
        const uint32_t i = __ffs(pass) - 1;
        const uint32_t ws = (i & 4u) ? ((i & 2u) ? w3 : w2) : ((i & 2u) ? w1 : w0);
        const uint32_t x16 = (ws >> ((i & 1u) * 16u)) & 0xFFFFu;
        const uint2 h = make_uint2(recX(i), record_y(v, (uint32_t)((x16 & 0xFFu) == 0), (uint32_t)((x16 >> 8) == 0), recFlag));
        if constexpr (NOSPILL) {
            uint32_t slice;
            if (record_keep(h, slice, gated)) {
                const uint32_t slot = atomicAdd(nHits, 1u);
                if (slot >= cap) return pass;   // this slot and the ones after it: again after the flush
                hits[slot] = h;
            }
        } else {
            triple_push(a, sh, guide, h);
        }
        pass &= pass - 1;
    }
    return 0;
}

// bucket [start, end) of the contiguous copy, `lanes` consecutive lanes (this one is number `gl`) share it
__device__ __forceinline__ void triple_bucket(const TripleArgs &a, TripleShared &sh, uint32_t guide, uint2 v, uint32_t start,
                                              uint32_t end, uint32_t gl, uint32_t lanes)
{
    const uint32_t t = (v.x >> 24) & 15u;
    const uint32_t gg = sh.res[t];
    const uint4 *__restrict__ base = reinterpret_cast<const uint4 *>(a.tv.res + (uint64_t)t * a.tv.stride);
    const uint32_t lastVec = (end - 1) >> 3;
    for (uint32_t vi = (start >> 3) + gl; vi <= lastVec; vi += lanes) {
        const uint4 r = __ldg(base + vi);
        const uint32_t first = vi << 3;
        const uint32_t lo = start > first ? start - first : 0u, hi = min(end - first, 8u);
        triple_vector(a, sh, guide, v, gg, r, ((1u << hi) - 1u) & ~((1u << lo) - 1u),
                      [first](uint32_t i) { return first + i; }, 0u);
    }
}

// end of the scan: finish the guide here (fused), or hand its hits on -- as a segment for k_score_segments or as
// keys for the general pipeline
template <bool FUSED, bool FLUSH = false, class Smem>
__device__ __forceinline__ void triple_epilogue(const TripleArgs &a, Smem &sm, uint32_t guide, uint64_t g,
                                                uint32_t entries, uint32_t visited)
{
    TripleShared &sh = sm.scan;
    if (a.streamed) {
        for (int o = 16; o > 0; o >>= 1) {
            entries += __shfl_down_sync(0xffffffffu, entries, o);
            visited += __shfl_down_sync(0xffffffffu, visited, o);
        }
        if ((threadIdx.x & 31u) == 0) { atomicAdd(&sh.count[0], entries); atomicAdd(&sh.count[1], visited); }
    }
    __syncthreads();
    const uint32_t nAll = sh.nHits, nLocal = min(nAll, kTripleHitCap);
    if (a.streamed && threadIdx.x < 2 && sh.count[threadIdx.x]) atomicAdd(a.streamed + threadIdx.x, (unsigned long long)sh.count[threadIdx.x]);
    if (a.maxRecords && threadIdx.x == 0) atomicMax(a.maxRecords, (unsigned long long)nAll);
    if (a.fuse && threadIdx.x == 0) {   // state after this wave unless the fused tail below changes it
        a.totMitOut[guide] = a.sp.totMit[guide]; a.totCfdOut[guide] = a.sp.totCfd[guide]; a.doneOut[guide] = 0;
    }
    if constexpr (FUSED && FLUSH) if (a.fuse && a.heavyKeys && sh.flushed) {
        // a heavy guide: the rest of its records become keys too, then the CTA sorts and finishes it
        __syncthreads();
        if (nAll) heavy_flush(a, sh, g);
        if (!sh.heavyBroken) {   // (else the launch is repeated with a larger buffer)
            if (threadIdx.x == 0) {
                sh.base = atomicAdd(a.heavyGuides, 1ull);
                atomicAdd(a.heavyHits, (unsigned long long)sh.heavyLen);
                atomicAdd(a.fusedHits, (unsigned long long)sh.heavyLen);
            }
            __syncthreads();
            HeavyDesc &hd = a.heavyDesc[sh.base];
            if (threadIdx.x == 0) { hd.guide = guide; hd.len = sh.heavyLen; }
            if (threadIdx.x <= kHeavyChunks) hd.chunkBase[threadIdx.x] = threadIdx.x < sh.heavyChunks ? sh.chunkBase[threadIdx.x] : 0ull;
        }
        return;
    }
    if constexpr (FUSED) if (a.fuse && nAll <= kTripleHitCap && !sh.flushed) {
        if (nAll == 0) return;
        __syncthreads();   // the defaults above are in place before score_guide's writer thread runs
        score_guide(sm.score, sm.group, nAll, guide, g, a.sp, a.totMitOut, a.totCfdOut, a.doneOut,
                    [&](uint32_t j, uint32_t &slice, uint64_t &site, uint32_t &occ, uint64_t &key) {
                        const uint2 h = sh.hits[j];
                        if (!record_keep(h, slice, sh.gated)) { slice = kNoSlice; return; }
                        site = hit_site(a.tv, h);
                        if (a.tv.siteOrdered && site != kSiteUnknown) {
                            // the site's text rank orders it; its id is only needed when it occurs more than once
                            key = site_text_key(site, 20);
                            if (h.y & kRecMulti)
                                occ = occ_of(a.sp, __ldg(a.tv.ids + (uint64_t)(h.y & 15u) * a.tv.stride + hit_position(a.tv, h)));
                        } else {
                            const uint32_t idRaw = __ldg(a.tv.ids + (uint64_t)(h.y & 15u) * a.tv.stride + hit_position(a.tv, h));
                            const uint32_t id = idRaw & (a.tv.occFlag ? 0x7FFFFFFFu : ~0u);
                            if (site == kSiteUnknown) site = __ldg(a.sp.sig + id);
                            occ = occ_of(a.sp, idRaw);
                            key = a.tv.siteOrdered ? site_text_key(site, 20) : (uint64_t)id;
                        }
                        if (a.tv.nibbleOrder) slice = order_slice(a.tv, site ^ g, slice);
                    });
        if (threadIdx.x == 0 && sm.score.kept) atomicAdd(a.fusedHits, (unsigned long long)sm.score.kept);
        return;
    }
    // hand the hits on: de-duplicate, compact, reserve a range of the segment / key buffer, resolve ids
    constexpr uint32_t kPerThread = kTripleHitCap / kTripleThreads;
    const bool segment = a.segCnt && nAll <= kTripleHitCap && !sh.flushed;
    uint32_t myPos[kPerThread], myMinE[kPerThread];
#pragma unroll
    for (uint32_t k = 0; k < kPerThread; k++) {
        const uint32_t j = threadIdx.x + k * kTripleThreads;
        myPos[k] = 0xFFFFFFFFu;
        if (j < nLocal && record_keep(sh.hits[j], myMinE[k], sh.gated)) myPos[k] = atomicAdd(&sh.nKept, 1u);
    }
    __syncthreads();
    const uint32_t nKept = sh.nKept;
    if (threadIdx.x == 0 && nKept) {
        if (segment) {
            sh.base = atomicAdd(a.segCount, (unsigned long long)nKept);
            a.segOff[guide] = sh.base;
            if (sh.base + nKept <= a.segCap) a.segCnt[guide] = nKept;   // else the launch is repeated with a larger buffer
        } else {
            sh.base = atomicAdd(a.hitCount, (unsigned long long)nKept);
        }
    }
    __syncthreads();
#pragma unroll
    for (uint32_t k = 0; k < kPerThread; k++) {
        if (myPos[k] == 0xFFFFFFFFu) continue;
        const uint2 h = sh.hits[threadIdx.x + k * kTripleThreads];
        const uint32_t id = __ldg(a.tv.ids + (uint64_t)(h.y & 15u) * a.tv.stride + hit_position(a.tv, h));
        const unsigned long long slot = sh.base + myPos[k];
        if (!segment) {
            if (slot < a.hitCap) a.hitKeys[slot] = general_key(a.tv, guide, myMinE[k], id);
        } else if (slot < a.segCap) {
            a.segKeys[slot] = ((uint64_t)record_order_slice(a.tv, a.sp.sig, h, g, id, myMinE[k]) << 32) | id;
            a.segSites[slot] = hit_site(a.tv, h);
        }
    }
}

// contiguous copy: offsets, then the bucket (two dependent round trips; the next visit's offsets are requested
// before the current bucket is processed)
template <bool FUSED>
__global__ void __launch_bounds__(kTripleThreads) k_scan_triple(const TripleArgs a)
{
    const uint32_t guide = blockIdx.x;
    if (a.done && a.done[guide]) {
        if (a.fuse && threadIdx.x == 0) {
            a.totMitOut[guide] = a.sp.totMit[guide]; a.totCfdOut[guide] = a.sp.totCfd[guide]; a.doneOut[guide] = 1;
        }
        return;
    }
    __shared__ typename std::conditional<FUSED, TripleSmem, TripleSmemScan>::type sm;
    TripleShared &sh = sm.scan;
    const uint64_t g = a.guides[guide];
    triple_prologue(a, sh, g, guide);

    const uint32_t octet = threadIdx.x >> 3, lane8 = threadIdx.x & 7u;
    constexpr uint32_t kOctets = kTripleThreads / 8;
    const uint32_t v0 = blockIdx.y * a.visitsPerCta, v1 = min(a.nVisits, v0 + a.visitsPerCta);
    uint32_t entries = 0, visited = 0;
    const uint2 *__restrict__ visits = reinterpret_cast<const uint2 *>(a.visits);

    if (a.tv.nonEmpty) {
        // small index: every lane looks one visit's bucket up in the bitmap; the warp's four octets then follow the offsets
        // of the buckets that hold something, four at a time
        const uint32_t lane = threadIdx.x & 31u, oct = lane >> 3;
        for (uint32_t e0 = v0; e0 < v1; e0 += kTripleThreads) {
            const uint32_t e = e0 + threadIdx.x;
            uint2 v = make_uint2(0, 0);
            bool some = false;
            if (e < v1) {
                v = __ldg(visits + e);
                const uint32_t t = (v.x >> 24) & 15u, key = sh.key[t] ^ (v.x & 0xFFFFFFu);
                some = (__ldg(a.tv.nonEmpty + ((size_t)t << 19) + (key >> 5)) >> (key & 31u)) & 1u;
                visited++;
            }
            uint32_t m = __ballot_sync(0xffffffffu, some);
            while (m) {
                uint32_t mm = m;
                for (uint32_t k = 0; k < oct; k++) mm &= mm - 1u;          // this octet's bucket: the oct-th one left
                const bool has = mm != 0;
                const int src = has ? __ffs(mm) - 1 : 0;
                const uint2 vo = make_uint2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
                if (has) {
                    const uint32_t t = (vo.x >> 24) & 15u;
                    const uint32_t *o = a.tv.offs + (uint64_t)t * (kTripleBuckets + 1) + (sh.key[t] ^ (vo.x & 0xFFFFFFu));
                    const uint32_t start = __ldg(o), end = __ldg(o + 1);
                    if (start < end) {
                        triple_bucket(a, sh, guide, vo, start, end, lane8, 8);
                        if (lane8 == 0) entries += end - start;
                    }
                }
                for (int k = 0; k < 4 && m; k++) m &= m - 1u;
            }
        }
        triple_epilogue<FUSED>(a, sm, guide, g, entries, visited);
        return;
    }
    uint32_t e = v0 + octet;
    uint2 v = make_uint2(0, 0);
    uint32_t start = 0, end = 0;
    if (e < v1) {
        v = __ldg(visits + e);
        const uint32_t t = (v.x >> 24) & 15u;
        const uint32_t *o = a.tv.offs + (uint64_t)t * (kTripleBuckets + 1) + (sh.key[t] ^ (v.x & 0xFFFFFFu));
        start = __ldg(o); end = __ldg(o + 1);
    }
    while (e < v1) {
        const uint32_t en = e + kOctets;
        uint2 vn = make_uint2(0, 0);
        uint32_t startn = 0, endn = 0;
        if (en < v1) {
            vn = __ldg(visits + en);
            const uint32_t tn = (vn.x >> 24) & 15u;
            const uint32_t *o = a.tv.offs + (uint64_t)tn * (kTripleBuckets + 1) + (sh.key[tn] ^ (vn.x & 0xFFFFFFu));
            startn = __ldg(o); endn = __ldg(o + 1);
        }
        if (start < end) {
            triple_bucket(a, sh, guide, v, start, end, lane8, 8);
            if (lane8 == 0) entries += end - start;
        }
        if (lane8 == 0) visited++;
        e = en; v = vn; start = startn; end = endn;
    }
    triple_epilogue<FUSED>(a, sm, guide, g, entries, visited);
}

// blocked, bit-sliced copy: ONE aligned read per visit.  SUBS lanes share a visit, each owning a 64-byte
// sub-block: two 32-byte loads (load_sub_block), then all 31 residuals at once --
//   mismatch flag of base b (one bit per slot):  x_b = (plane[2b] ^ G[2b]) | (plane[2b+1] ^ G[2b+1])
//   count of the 8 flags per slot with a carry-save adder (4 full + 3 half adders), compared with the budget,
// 30 LOP3 for 31 (guide, site) pairs instead of ~7 instructions per pair; no POPC, no offset lookup.
#ifndef ISSL_GATHER_IMAD
#define ISSL_GATHER_IMAD 1      // 1: the hit loop's 16 pre-shifts are IMADs (fma pipe) instead of SHFs (alu pipe): 3.705 -> 3.667 ms
#endif
// a * b as an IMAD the compiler cannot turn back into a shift
__device__ __forceinline__ uint32_t mul_lo(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}

#ifndef ISSL_BUDGET_MULHI
#define ISSL_BUDGET_MULHI 1     // 1: budget bit -> 32-bit mask through IMAD.HI instead of ISETP + SEL: 3.667 -> 3.660 ms
#endif
// all ones when x is negative, as an IMAD.HI
__device__ __forceinline__ uint32_t sign_mask(uint32_t x)
{
    int r;
    asm("mul.hi.s32 %0, %1, 1;" : "=r"(r) : "r"((int)x));
    return (uint32_t)r;
}

__device__ __forceinline__ void bs_full_add(uint32_t a, uint32_t b, uint32_t c, uint32_t &sum, uint32_t &carry)
{
    sum = a ^ b ^ c;
    carry = (a & b) | (c & (a ^ b));
}

// One 64-byte sub-block of the blocked copy: four 16-byte words = 16 bit planes of 32 slots; all 31 residuals at once --
//   mismatch flag of base b (one bit per slot):  x_b = (plane[2b] ^ G[2b]) | (plane[2b+1] ^ G[2b+1])
//   count of the 8 flags per slot with a carry-save adder (4 full + 3 half adders), compared with the budget,
// 30 LOP3 for 31 (guide, site) pairs instead of ~7 instructions per pair; no POPC, no offset lookup.
// Hits are recorded in hits[0 .. cap) through one reservation on *nHits.  `allowed`: the slots still to be reported.
// A record list found full: kFullSpill = the hit goes to the general pipeline's buffer on its own; kFullRetry / kFullDrop =
// the slots that did not fit are returned (the flush variant tries them again after emptying the list; the warp-per-guide
// kernel hands the whole guide to the CTA-per-guide kernel).
enum { kFullSpill = 0, kFullRetry = 1, kFullDrop = 2 };

template <int FULL>
__device__ __forceinline__ uint32_t triple_sub_block(const TripleArgs &a, const uint32_t *guide, const uint4 *mask, uint32_t *nHits, uint2 *hits,
                                                     const uint32_t cap, const uint2 v, const uint32_t key, const uint32_t sub, const uint4 q0,
                                                     const uint4 q1, const uint4 q2, const uint4 q3, const uint32_t allowed, uint32_t &cnt,
                                                     const uint32_t gated = 31u)
{
    cnt = (q0.x & 1u) | ((q0.y & 1u) << 1) | ((q0.z & 1u) << 2) | ((q0.w & 1u) << 3) | ((q1.x & 1u) << 4);
#if !ISSL_SUB_NOEARLY
    if (cnt == 0) return 0;
#endif
    const uint4 m0 = mask[0], m1 = mask[1], m2 = mask[2], m3 = mask[3];
    const uint32_t x0 = (q0.x ^ m0.x) | (q0.y ^ m0.y), x1 = (q0.z ^ m0.z) | (q0.w ^ m0.w);
    const uint32_t x2 = (q1.x ^ m1.x) | (q1.y ^ m1.y), x3 = (q1.z ^ m1.z) | (q1.w ^ m1.w);
    const uint32_t x4 = (q2.x ^ m2.x) | (q2.y ^ m2.y), x5 = (q2.z ^ m2.z) | (q2.w ^ m2.w);
    const uint32_t x6 = (q3.x ^ m3.x) | (q3.y ^ m3.y), x7 = (q3.z ^ m3.z) | (q3.w ^ m3.w);
    uint32_t sa, ca, sb, cb, sc, cc, t1, u1;
    bs_full_add(x0, x1, x2, sa, ca);
    bs_full_add(x3, x4, x5, sb, cb);
    bs_full_add(x6, x7, sa, sc, cc);
    const uint32_t s0 = sb ^ sc, cd = sb & sc;          // count bit 0
    bs_full_add(ca, cb, cc, t1, u1);
    const uint32_t s1 = t1 ^ cd, u2 = t1 & cd;          // count bit 1
    const uint32_t s2 = u1 ^ u2, s3 = u1 & u2;          // count bits 2, 3
    // slots whose count exceeds the budget: a bit-sliced comparator, most significant bit last (lanes of a warp hold
    // visits with different budgets -- no branches)
#if ISSL_BUDGET_MULHI
    // budget bit -> mask on the fma pipe: the bit moved to the sign, then the high word of a signed product with 1
    const uint32_t b0 = sign_mask(v.x << 3), b1 = sign_mask(v.x << 2), b2 = sign_mask(v.x << 1);
#else
    const uint32_t bud = v.x >> 28;
    const uint32_t b0 = 0u - (bud & 1u), b1 = 0u - ((bud >> 1) & 1u), b2 = 0u - ((bud >> 2) & 1u);
#endif
    uint32_t over = s0 & ~b0;
    over = (s1 & ~b1) | (~(s1 ^ b1) & over);
    over = (s2 & ~b2) | (~(s2 ^ b2) & over);
    over |= s3;
    // ... and of those within budget, the ones this triple is responsible for (the stateless replacement of the
    // reference's toggle bitset, ref :385-390): a function of the visit and of "the residual matches exactly on slice
    // p / on slice q", tabulated per visit.  Nearly all visits keep only entries that match on neither.
    const uint32_t pEx = ~(x0 | x1 | x2 | x3), qEx = ~(x4 | x5 | x6 | x7);
    const uint32_t kt = v.y >> 16;
    uint32_t keep = ~(pEx | qEx);
    if (kt != 1u)
        keep = ((kt & 1u) ? ~(pEx | qEx) : 0u) | ((kt & 2u) ? (pEx & ~qEx) : 0u) | ((kt & 4u) ? (~pEx & qEx) : 0u) | ((kt & 8u) ? (pEx & qEx) : 0u);
    uint32_t pass = ~over & keep & ((2u << cnt) - 2u) & allowed;   // slots 1..cnt hold residuals
    if (pass) {
        // one reservation for all hits of the sub-block; the sub-block's entries that occur more than once are its first ones
        uint32_t slot = atomicAdd(nHits, (uint32_t)__popc(pass));
        const uint32_t multi = (q1.z & 1u) | ((q1.w & 1u) << 1) | ((q2.x & 1u) << 2) | ((q2.y & 1u) << 3) | ((q2.z & 1u) << 4);
        const uint32_t multiMask = (2u << multi) - 2u;
        const uint32_t y0 = ((v.x >> 24) & 15u) | ((v.y & 31u) << 4) | kRecBlocked;
        do {
            if (FULL != kFullSpill && slot >= cap) break;   // list full: these slots are handed back
            const uint32_t sl = __ffs(pass) - 1;
            pass &= pass - 1;
            // the entry's residual, gathered back from the 16 planes (reading it back later costs more, see the record's
            // description): every plane is shifted so that the slot's bit becomes its top bit, which a funnel shift then
            // feeds into r
            const uint32_t up = 31u - sl;
            uint32_t r = 0;
#if ISSL_GATHER_IMAD
            // the 16 pre-shifts as multiplications by 2^up: IMAD runs on the fma pipe, which idles while LOP3 / SHF keep the
            // alu pipe 67 % busy (ncu)
            const uint32_t mul = 1u << up;
#define ISSL_UP(x) mul_lo(x, mul)
#else
#define ISSL_UP(x) ((x) << up)
#endif
            r = __funnelshift_l(ISSL_UP(q3.w), r, 1); r = __funnelshift_l(ISSL_UP(q3.z), r, 1);
            r = __funnelshift_l(ISSL_UP(q3.y), r, 1); r = __funnelshift_l(ISSL_UP(q3.x), r, 1);
            r = __funnelshift_l(ISSL_UP(q2.w), r, 1); r = __funnelshift_l(ISSL_UP(q2.z), r, 1);
            r = __funnelshift_l(ISSL_UP(q2.y), r, 1); r = __funnelshift_l(ISSL_UP(q2.x), r, 1);
            r = __funnelshift_l(ISSL_UP(q1.w), r, 1); r = __funnelshift_l(ISSL_UP(q1.z), r, 1);
            r = __funnelshift_l(ISSL_UP(q1.y), r, 1); r = __funnelshift_l(ISSL_UP(q1.x), r, 1);
            r = __funnelshift_l(ISSL_UP(q0.w), r, 1); r = __funnelshift_l(ISSL_UP(q0.z), r, 1);
            r = __funnelshift_l(ISSL_UP(q0.y), r, 1); r = __funnelshift_l(ISSL_UP(q0.x), r, 1);
            const uint32_t y = y0 | (((pEx >> sl) & 1u) << 9) | (((qEx >> sl) & 1u) << 10) | (((multiMask >> sl) & 1u) << 12) | (r << 16);
#undef ISSL_UP
            // entry number sub*31 + sl - 1 of the bucket, recorded as entry + 1
            const uint2 h = make_uint2(key | ((sub * kSubEntries + sl) << 24), y);
            if (FULL != kFullSpill || slot < cap) hits[slot] = h; else triple_spill(a, *guide, h, gated);
            slot++;
        } while (pass);
    }
    return pass;
}

// SUBS lanes share a visit, each owning one 64-byte sub-block (two 32-byte loads in flight per lane).  A lane owning the
// whole 128-byte block -- eight loads in flight, one visit-table read and one address per two sub-blocks, 64 registers and
// 8 CTAs per SM -- measured SLOWER (5.2 against 3.8 ms per 100 000 guides, profiles/r02_ab_lane_subs.jsonl), and so did
// asynchronous copies into a shared-memory ring (cp.async.bulk 4.99 ms, cp.async 4.29 ms against 2.74 ms for the bare
// loads, profiles/r02_scan_variants_ring_tma.jsonl): the kernel is bound by instruction issue and by latency across many
// warps, not by bytes in flight per lane.
//
// GATES = sliceWidth 10 (TripleView::perm10): the visit's keep table is narrowed by the guide's open gates, and a visit none
// of whose entries the guide may keep is not read at all.
//
// FLUSH = the variant for guides with more hits than the record list holds (maxDist 5-6, repeat families): one barrier
// per round of visits; a full list is emptied by all threads -- into the guide's own sort keys (heavy_flush; the guide is
// then sorted and finished by this CTA, heavy_finish) or, without that buffer, into the general pipeline's -- and a hit
// that found the list full is tried again afterwards: nothing is ever dropped and nothing leaves the CTA hit by hit.
template <int SUBS, bool FUSED, bool FLUSH, bool GATES = false>
__global__ void __launch_bounds__(kTripleThreads, FLUSH ? ISSL_TRIPLE_MIN_CTAS_FLUSH : ISSL_TRIPLE_MIN_CTAS) k_scan_triple_blocked(const TripleArgs a)
{
    constexpr int LSUBS = 1;   // sub-blocks per lane
    const uint32_t guide = a.guideList ? a.guideList[blockIdx.x] : blockIdx.x;   // (the guides the warp-per-guide kernel left)
    if (a.done && a.done[guide]) {
        if (a.fuse && threadIdx.x == 0) {
            a.totMitOut[guide] = a.sp.totMit[guide]; a.totCfdOut[guide] = a.sp.totCfd[guide]; a.doneOut[guide] = 1;
        }
        return;
    }
    __shared__ typename std::conditional<FUSED, TripleSmem, TripleSmemScan>::type sm;
    TripleShared &sh = sm.scan;
    const uint64_t g = a.guides[guide];
    triple_prologue(a, sh, g, guide);

    constexpr uint32_t LPV = SUBS / LSUBS;                  // lanes per visit
    constexpr uint32_t V = kTripleThreads / LPV;            // visits in flight per CTA
    const uint32_t sub0 = (threadIdx.x % LPV) * LSUBS, vslot = threadIdx.x / LPV;
    const uint32_t v0 = blockIdx.y * a.visitsPerCta, v1 = min(a.nVisits, v0 + a.visitsPerCta);
    uint32_t entries = 0, visited = 0;
    const uint2 *__restrict__ visits = reinterpret_cast<const uint2 *>(a.visits);
    const bool heavy = FUSED && FLUSH && a.heavyKeys != nullptr;

    // a full record list is emptied by all threads (FLUSH only)
    auto flush = [&]() {
        if constexpr (FUSED) { if (heavy) { heavy_flush(a, sh, a.guides[sh.guide]); return; } }
        triple_flush(a, sh, sh.guide);
    };

    // one sub-block: 31 residuals against the guide's, hits recorded.  `allowed`: the slots still to be reported; returns
    // the slots that found the record list full (FLUSH only: 0 otherwise, such hits go to the general pipeline one by one)
    auto sub_block = [&](const uint2 v, const uint32_t t, const uint32_t key, const uint32_t sub, const uint4 q0, const uint4 q1,
                         const uint4 q2, const uint4 q3, const uint32_t allowed, const bool count) -> uint32_t {
        uint32_t cnt = 0;
        const uint32_t left = triple_sub_block<FLUSH ? kFullRetry : kFullSpill>(a, &sh.guide, sh.mask[t], &sh.nHits, sh.hits, kTripleHitCap, v, key, sub,
                                                                                 q0, q1, q2, q3, allowed, cnt, sh.gated);
        if (count) entries += cnt;
        return left;
    };

    // entries [start, end) of the contiguous copy (the rest of a bucket that does not fit its block), shared by `lanes`
    // threads of which this one is number `gl`.  FLUSH: every thread of the CTA calls this the same number of times
    // (it contains barriers); an empty range is fine.
    auto scan_range = [&](const uint2 v, const uint32_t start, const uint32_t end, const uint32_t gl, const uint32_t lanes) {
        const uint32_t t = (v.x >> 24) & 15u;
        if constexpr (!FLUSH) {
            if (start < end) triple_bucket(a, sh, sh.guide, v, start, end, gl, lanes);
        } else {
            const uint32_t gg = sh.res[t];
            const uint4 *__restrict__ base = reinterpret_cast<const uint4 *>(a.tv.res + (uint64_t)t * a.tv.stride);
            uint32_t vi = (start >> 3) + gl, allowed = 0xFFu;
            const uint32_t vecEnd = start < end ? ((end - 1) >> 3) + 1u : 0u;   // one past the last vector
            for (;;) {
                if (vi < vecEnd) {
                    const uint4 r = __ldg(base + vi);
                    const uint32_t first = vi << 3;
                    const uint32_t lo = start > first ? start - first : 0u, hi = min(end - first, 8u);
                    const uint32_t left = triple_vector<true>(a, sh, 0u, v, gg, r, ((1u << hi) - 1u) & ~((1u << lo) - 1u) & allowed,
                                                              [first](uint32_t i) { return first + i; }, 0u, &sh.nHits, sh.hits, kTripleHitCap, sh.gated);
                    if (left) allowed = left; else { vi += lanes; allowed = 0xFFu; }
                }
                if (__syncthreads_or(sh.nHits > kTripleFlushAt)) flush();
                if (!__syncthreads_or(vi < vecEnd)) break;
            }
        }
        if (gl == 0 && start < end) entries += end - start;
    };

    // the first n buckets of the noted list (visits whose bucket has more entries than its block holds): their remainders
    // from the contiguous copy, sixteen buckets at a time, eight lanes each; long remainders (repeat families) are noted once
    // more and shared by all threads.  One remainder is a chain of dependent loads (offsets, then residuals): what counts is
    // how many chains are in flight.  Called by all threads of the CTA.
    auto process_noted = [&](const uint32_t n) {
        for (uint32_t j0 = 0; j0 < n; j0 += kTripleThreads / 8) {
            const uint32_t j = j0 + (threadIdx.x >> 3), gl = threadIdx.x & 7u;
            uint2 vo = make_uint2(0, 0);
            uint32_t start = 0, end = 0;
            if (j < n) {
                vo = __ldg(visits + sh.ovf[j]);
                const uint32_t to = (vo.x >> 24) & 15u;
                const uint32_t *o = a.tv.offs + (uint64_t)to * (kTripleBuckets + 1) + (sh.key[to] ^ (vo.x & 0xFFFFFFu));
                start = __ldg(o) + SUBS * kSubEntries; end = __ldg(o + 1);
                uint32_t isLong = 0;
                if (gl == 0 && end - start > kTripleLongBucket) {
                    const uint32_t slot = atomicAdd(&sh.nLong, 1u);
                    if (slot < kTripleLongCap) { sh.longList[slot] = sh.ovf[j]; isLong = 1; }
                }
                isLong = __shfl_sync(0xffu << (threadIdx.x & 24u), isLong, 0, 8);
                if (isLong) end = start;
            }
            scan_range(vo, start, end, gl, 8);
        }
    };
    auto process_long = [&]() {   // one bucket at a time, all threads
        __syncthreads();
        const uint32_t nLong = min(sh.nLong, kTripleLongCap);
        for (uint32_t j = 0; j < nLong; j++) {
            const uint2 vo = __ldg(visits + sh.longList[j]);
            const uint32_t to = (vo.x >> 24) & 15u;
            const uint32_t *o = a.tv.offs + (uint64_t)to * (kTripleBuckets + 1) + (sh.key[to] ^ (vo.x & 0xFFFFFFu));
            scan_range(vo, __ldg(o) + SUBS * kSubEntries, __ldg(o + 1), threadIdx.x, kTripleThreads);
        }
    };

#if ISSL_TRIPLE_PIPE
    // the visit of the NEXT round is requested while this round's blocks are on their way: the visit-table entry (an L1 hit)
    // heads the chain entry -> bucket key -> address -> block loads, and a warp that waits for it has no block in flight
    uint2 vNext = make_uint2(0, 0);
    if (v0 + vslot < v1) vNext = __ldg(visits + v0 + vslot);
#endif
    for (uint32_t e0 = v0; e0 < v1; e0 += V) {   // the same number of rounds for every thread of the CTA
        const uint32_t e = e0 + vslot;
        uint2 v = make_uint2(0, 0);
        uint32_t t = 0, key = 0;
        uint4 q[LSUBS][4];
        uint32_t left[LSUBS];
        bool ovfPending = false;
#pragma unroll
        for (int s = 0; s < LSUBS; s++) left[s] = 0;
        bool live = e < v1;
        if (live) {
#if ISSL_TRIPLE_PIPE
            v = vNext;
            if (e + V < v1) vNext = __ldg(visits + e + V);
#else
            v = ld_visit(visits + e);
#endif
            if constexpr (GATES) {
                // sliceWidth 10: of the four ways the residual's two slices can match exactly, only those that leave the hit
                // with an exact unit behind an open gate; a visit none of whose entries this guide may keep is not read
                const uint32_t gate = sh.gated, p = (v.y >> 8) & 15u, q = (v.y >> 12) & 15u;
                const uint32_t ok = (v.y & gate & 31u) ? 0xFu : ((((gate >> p) & 1u) ? 0xAu : 0u) | (((gate >> q) & 1u) ? 0xCu : 0u));
                const uint32_t kt = (v.y >> 16) & ok;
                v.y = (v.y & 0xFFFFu) | (kt << 16);
                live = kt != 0;
            }
        }
#if ISSL_TRIPLE_PREFETCH > 0
        {   // the block of this lane's visit ISSL_TRIPLE_PREFETCH rounds ahead: requested into L2 now, so that the round that
            // needs it waits for an L2 hit rather than for DRAM (no registers held, unlike a second visit in flight)
            const uint32_t en = e + ISSL_TRIPLE_PREFETCH * V;
            if (en < v1) {
                const uint2 vn = __ldg(visits + en);
                const uint32_t tn = (vn.x >> 24) & 15u;
                const uint4 *pn = a.tv.blk + ((((uint64_t)tn << 24) | (sh.key[tn] ^ (vn.x & 0xFFFFFFu))) * SUBS + sub0) * 4;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pn));
            }
        }
#endif
        if (live) {
            t = (v.x >> 24) & 15u; key = sh.key[t] ^ (v.x & 0xFFFFFFu);
            const uint4 *__restrict__ p = a.tv.blk + ((((uint64_t)t << 24) | key) * SUBS + sub0) * 4;
            // read once: no line left behind in L1 (where the visit table and the offsets live), evict-first in L2
#pragma unroll
            for (int s = 0; s < LSUBS; s++) {
                load_sub_block(p + 4 * s, q[s][0], q[s][1], q[s][2], q[s][3]);
            }
            if constexpr (GATES) { if (sub0 == 0) visited++; }   // (without gates every visit is read: counted after the loop)
            if ((q[0][1].y & 1u) && sub0 == 0) {   // more entries than the block holds: noted for after the loop
                const uint32_t slot = atomicAdd(&sh.nOvf, 1u);
                if (slot < kTripleOvfCap) sh.ovf[slot] = (uint16_t)e;   // visit tables have fewer than 2^16 entries
                else if constexpr (FLUSH) ovfPending = true;
                else if (a.ovfBits) {
                    // list full: one bit per visit in this CTA's row of a global bitmap (L2), read back after the loop.  Indexes
                    // whose sites are not uniform need this: the reference's extractor takes the reverse strand's site from
                    // the wrong end of its match (extractOfftargets.py:97-106), so half of a real index's sites end in AG / GG,
                    // the buckets keyed on those values are 4.5 x fuller than the mean and most visits of such a guide overflow
                    atomicOr(a.ovfBits + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * a.ovfWords + ((e - v0) >> 5), 1u << ((e - v0) & 31u));
                } else {   // (no bitmap: this lane reads the rest alone)
                    const uint32_t *o = a.tv.offs + (uint64_t)t * (kTripleBuckets + 1) + key;
                    scan_range(v, __ldg(o) + SUBS * kSubEntries, __ldg(o + 1), 0, 1);
                }
            }
#pragma unroll
            for (int s = 0; s < LSUBS; s++) left[s] = sub_block(v, t, key, sub0 + s, q[s][0], q[s][1], q[s][2], q[s][3], ~0u, true);
        }
        if constexpr (FLUSH) {
            // a flush is due when the list is nearly full -- always the case when a hit found it full
            while (__syncthreads_or(sh.nHits > kTripleFlushAt)) {
                flush();
#pragma unroll
                for (int s = 0; s < LSUBS; s++)
                    if (left[s]) left[s] = sub_block(v, t, key, sub0 + s, q[s][0], q[s][1], q[s][2], q[s][3], left[s], false);
            }
            // the list of noted buckets is full: finish them now, so that none is ever skipped (sixteen at a time -- one bucket
            // at a time with all threads, as this was first written, left 116 of 128 threads idle on a 94-entry remainder:
            // 23.4 ms per 100 000 guides on an index made by the reference's extractor)
            while (__syncthreads_or(sh.nOvf >= kTripleOvfCap)) {
                process_noted(kTripleOvfCap);
                process_long();
                __syncthreads();
                if (threadIdx.x == 0) { sh.nOvf = 0; sh.nLong = 0; }
                __syncthreads();
                if (ovfPending) {   // buckets that found the list full
                    const uint32_t slot = atomicAdd(&sh.nOvf, 1u);
                    if (slot < kTripleOvfCap) { sh.ovf[slot] = (uint16_t)e; ovfPending = false; }
                }
            }
        }
    }
    if constexpr (!GATES) { if (sub0 == 0 && v0 + vslot < v1) visited = (v1 - v0 - vslot + V - 1) / V; }
    __syncthreads();
    if (sh.nOvf) {   // CTA-uniform: the buckets noted during the scan
        process_noted(min(sh.nOvf, kTripleOvfCap));
        if constexpr (!FLUSH) {
            if (sh.nOvf > kTripleOvfCap && a.ovfBits) {   // CTA-uniform: the visits noted in the bitmap, a word per group of eight lanes
                // (two lanes per bucket, dealt out evenly through a prefix sum over the words, was measured: 12.6 against 9.8 ms)
                const uint32_t *bits = a.ovfBits + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * a.ovfWords;
                const uint32_t gl = threadIdx.x & 7u, words = (v1 - v0 + 31u) >> 5;
                for (uint32_t w = threadIdx.x >> 3; w < words; w += kTripleThreads / 8) {
                    uint32_t word = __ldcg(bits + w);
                    while (word) {
                        const uint32_t e = v0 + (w << 5) + (uint32_t)__ffs(word) - 1u;
                        word &= word - 1u;
                        const uint2 vo = __ldg(visits + e);
                        const uint32_t to = (vo.x >> 24) & 15u;
                        const uint32_t *o = a.tv.offs + (uint64_t)to * (kTripleBuckets + 1) + (sh.key[to] ^ (vo.x & 0xFFFFFFu));
                        uint32_t start = __ldg(o) + SUBS * kSubEntries;
                        const uint32_t end = __ldg(o + 1);
                        uint32_t isLong = 0;
                        if (gl == 0 && end - start > kTripleLongBucket) {
                            const uint32_t slot = atomicAdd(&sh.nLong, 1u);
                            if (slot < kTripleLongCap) { sh.longList[slot] = (uint16_t)e; isLong = 1; }
                        }
                        isLong = __shfl_sync(0xffu << (threadIdx.x & 24u), isLong, 0, 8);
                        if (isLong) start = end;
                        scan_range(vo, start, end, gl, 8);
                    }
                }
            }
        }
        process_long();
    }
    triple_epilogue<FUSED, FLUSH>(a, sm, sh.guide, a.guides[sh.guide], entries, visited);   // (the guide is read again: two registers the loop needs)
}

// ------------------------------------------------------------------------------------------------
// K1s: warp per guide, for visit tables of a few dozen buckets (maxDist <= 3: 10 / 130 visits per guide, about one hit).
// A CTA per guide leaves most of its 128 lanes idle there and pays a prologue, barriers and a 512-entry tail for one
// round of loads; here a CTA holds four guides, each warp builds its guide's masks, reads the guide's blocks (a lane per
// sub-block), records the hits in a list of its own and finishes the guide: sites from the records, scores (ref :392-461),
// rank by (slice, site text) among the warp's hits, ordered accumulation with the early exit (ref :466-502).  No
// barriers.  A guide that does not fit (a bucket larger than its block, more than 32 hits: repeat families) is flagged and
// taken by the CTA-per-guide kernel afterwards.  Needs the blocked copy and an index in text order.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kSmallHitCap = 32;
constexpr uint32_t kSmallMaxVisits = 256;

struct SmallWarp {
    uint4 mask[kTripleCount][4];
    uint2 hits[kSmallHitCap];
    uint64_t order[kSmallHitCap];
    double mit[kSmallHitCap], cfd[kSmallHitCap];
    uint32_t key[kTripleCount];
    uint32_t nHits, pad;
};

template <int SUBS>
__global__ void __launch_bounds__(kTripleThreads) k_scan_triple_small(const TripleArgs a)
{
    __shared__ SmallWarp sw[kTripleThreads / 32];
    // the launch's counters (entries, visits, hits) are summed per CTA first and added to the global ones by the warp that
    // finishes last: one same-address atomic per guide and counter was a measurable part of a 0.24 ms launch
    __shared__ uint32_t ctaStat[4];   // entries, visits, hits, warps that are through
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t guide = blockIdx.x * (kTripleThreads / 32) + warp;
    if (threadIdx.x < 4) ctaStat[threadIdx.x] = 0;
    __syncthreads();
    uint32_t statEntries = 0, statVisited = 0, statHits = 0;
    auto leave = [&]() {
        if (lane != 0) return;
        if (statEntries) atomicAdd(&ctaStat[0], statEntries);
        if (statVisited) atomicAdd(&ctaStat[1], statVisited);
        if (statHits) atomicAdd(&ctaStat[2], statHits);
        __threadfence_block();
        if (atomicAdd(&ctaStat[3], 1u) != kTripleThreads / 32 - 1) return;
        __threadfence_block();
        const uint32_t e = atomicAdd(&ctaStat[0], 0u), v = atomicAdd(&ctaStat[1], 0u), h = atomicAdd(&ctaStat[2], 0u);
        if (a.streamed && (e | v)) { atomicAdd(a.streamed, (unsigned long long)e); atomicAdd(a.streamed + 1, (unsigned long long)v); }
        if (h) atomicAdd(a.fusedHits, (unsigned long long)h);
    };
    if (guide >= a.nGuides) { leave(); return; }
    if (a.done && a.done[guide]) {
        if (lane == 0) { a.totMitOut[guide] = a.sp.totMit[guide]; a.totCfdOut[guide] = a.sp.totCfd[guide]; a.doneOut[guide] = 1; }
        leave();
        return;
    }
    SmallWarp &w = sw[warp];
    const uint64_t g = a.guides[guide];
    if (lane < kTripleCount) w.key[lane] = triple_key(g, c_tripleSlices[lane][0], c_tripleSlices[lane][1], c_tripleSlices[lane][2]);
    for (uint32_t i = lane; i < kTripleCount * 16; i += 32) {
        const uint32_t t = i >> 4, r = triple_res(g, c_tripleSlices[t][3], c_tripleSlices[t][4]);
        reinterpret_cast<uint32_t *>(w.mask)[i] = 0u - ((r >> (i & 15u)) & 1u);
    }
    if (lane == 0) w.nHits = 0;
    __syncwarp();
    const uint2 *__restrict__ visits = reinterpret_cast<const uint2 *>(a.visits);
    uint32_t entries = 0, visited = 0;
    bool bad = false;
    for (uint32_t i = lane; i < a.nVisits * SUBS; i += 32) {
        const uint32_t e = i / SUBS, sub = i % SUBS;
        const uint2 v = __ldg(visits + e);
        const uint32_t t = (v.x >> 24) & 15u, key = w.key[t] ^ (v.x & 0xFFFFFFu);
        const uint4 *__restrict__ p = a.tv.blk + ((((uint64_t)t << 24) | key) * SUBS + sub) * 4;
        uint4 q0, q1, q2, q3;
        load_sub_block(p, q0, q1, q2, q3);
        if (sub == 0) visited++;
        if ((q1.y & 1u) && sub == 0) {   // more entries than the block holds: this lane reads the rest from the contiguous copy
            const uint32_t *o = a.tv.offs + (uint64_t)t * (kTripleBuckets + 1) + key;
            const uint32_t start = __ldg(o) + SUBS * kSubEntries, end = __ldg(o + 1);
            if (end - start > 64u) bad = true;   // a repeat family's bucket: the CTA-per-guide kernel shares it among its threads
            else {
                const uint32_t gg = triple_res(g, c_tripleSlices[t][3], c_tripleSlices[t][4]) * 0x10001u;
                const uint4 *__restrict__ base = reinterpret_cast<const uint4 *>(a.tv.res + (uint64_t)t * a.tv.stride);
                for (uint32_t vi = start >> 3; vi <= (end - 1) >> 3; vi++) {
                    const uint4 r = __ldg(base + vi);
                    const uint32_t first = vi << 3, lo = start > first ? start - first : 0u, hi = min(end - first, 8u);
                    if (triple_vector<true>(a, *reinterpret_cast<TripleShared *>(&w), guide, v, gg, r, ((1u << hi) - 1u) & ~((1u << lo) - 1u),
                                            [first](uint32_t i) { return first + i; }, 0u, &w.nHits, w.hits, kSmallHitCap, 31u)) bad = true;
                }
                entries += end - start;
            }
        }
        uint32_t cnt;
        if (triple_sub_block<kFullDrop>(a, nullptr, w.mask[t], &w.nHits, w.hits, kSmallHitCap, v, key, sub, q0, q1, q2, q3, ~0u, cnt)) bad = true;
        entries += cnt;
    }
    __syncwarp();
    if (__any_sync(0xffffffffu, bad)) {
        if (lane == 0) a.redo[atomicAdd(a.redoCount, 1ull)] = guide;
        leave();
        return;
    }
    if (a.streamed) {
        for (int o = 16; o > 0; o >>= 1) {
            entries += __shfl_down_sync(0xffffffffu, entries, o);
            visited += __shfl_down_sync(0xffffffffu, visited, o);
        }
        statEntries = entries; statVisited = visited;
    }
    const uint32_t n = w.nHits;
    double mit = 0.0, cfd = 0.0;
    if (lane == 0) { mit = a.sp.totMit[guide]; cfd = a.sp.totCfd[guide]; }
    bool stop = false;
    if (n) {
        uint64_t okey = 0;
        double cm = 0.0, cc = 0.0;
        if (lane < n) {
            uint32_t slice, occ;
            uint64_t site;
            record_resolve(a, w.hits[lane], g, 31u, slice, site, occ);
            int dist;
            hit_contrib(a.sp.tb, g, site, occ, a.sp.calcMit, a.sp.calcCfd, cm, cc, dist);
            okey = ((uint64_t)slice << 60) | site_text_key(site, 20);
            w.order[lane] = okey;
        }
        __syncwarp();
        if (lane < n) {
            uint32_t r = 0;
            for (uint32_t q = 0; q < n; q++) r += (uint32_t)(w.order[q] < okey);
            w.mit[r] = cm; w.cfd[r] = cc;
        }
        __syncwarp();
        if (lane == 0) {
            for (uint32_t i = 0; i < n && !stop; i++) {
                mit = __dadd_rn(mit, w.mit[i]);
                cfd = __dadd_rn(cfd, w.cfd[i]);
                if (a.sp.checkExit) stop = exit_predicate(a.sp.method, mit, cfd, a.sp.maximumSum);
            }
            statHits = n;
        }
    }
    if (lane == 0) { a.totMitOut[guide] = mit; a.totCfdOut[guide] = cfd; a.doneOut[guide] = stop ? 1 : 0; }
    leave();
}

// sliceWidth 4: the ordering slice of every general-pipeline key, from the site itself (ref :330-390: a hit is met
// first in the lowest exactly matching 2-base slice)
__global__ void k_fix_order_slices(uint64_t *keys, uint64_t n, const uint64_t *guides, const uint64_t *sig)
{
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint64_t key = keys[j];
    const uint64_t x = sig[(uint32_t)key] ^ guides[(key & ~kKeyOccursOnce) >> kTripleKeyBits];
    uint64_t s = 0;
    while (s < 9 && ((x >> (4 * s)) & 15ull) != 0) s++;
    keys[j] = (key & ~(15ull << 32)) | (s << 32);
}

// ------------------------------------------------------------------------------------------------
// K2t: finish one guide per CTA from its own segment of the key buffer: sort the keys back into the
// reference's visiting order (slice, then ascending id inside the list, ref :330-344) with a bitonic
// network in shared memory, score every hit (ref :392-461), and accumulate in that order with the
// reference's early exit (ref :394, :460, :466-502) -- replaces radix sort + k_contrib + k_accumulate
// when every guide's hits fit a segment.
// ------------------------------------------------------------------------------------------------
struct SegmentArgs {
    const uint64_t *segKeys, *segSites;
    const uint64_t *segOff;
    const uint32_t *segCnt;
    const uint64_t *guides;
    ScoreParams sp;
};

__global__ void __launch_bounds__(kTripleThreads) k_score_segments(const SegmentArgs a)
{
    const uint32_t guide = blockIdx.x;
    const uint32_t n = a.segCnt[guide];
    if (n == 0) return;
    __shared__ ScoreShared ss;
    __shared__ uint64_t group[kScoreGroupWords];
    const uint64_t off = a.segOff[guide];
    const uint64_t g = a.guides[guide];
    score_guide(ss, group, n, guide, g, a.sp, a.sp.totMit, a.sp.totCfd, a.sp.done,
                [&](uint32_t j, uint32_t &slice, uint64_t &site, uint32_t &occ, uint64_t &key) {
                    const uint64_t k = a.segKeys[off + j];
                    const uint32_t idRaw = (uint32_t)k;
                    slice = (uint32_t)(k >> 32) & 15u;
                    site = a.segSites[off + j];
                    const uint32_t id = idRaw & (a.sp.occFlag ? 0x7FFFFFFFu : ~0u);
                    if (site == kSiteUnknown) site = __ldg(a.sp.sig + id);
                    occ = occ_of(a.sp, idRaw);
                    key = id;
                });
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
