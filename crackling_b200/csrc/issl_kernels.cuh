// issl_kernels.cuh -- hand-written sm_100a kernels of libissl_cuda.
//
// "ref:" = /root/reference/src/ISSL/.  See DESIGN.md for the data layout and the roofline of
// each kernel.  Everything here is integer/bitwise work except the survivor scoring
// (k_contrib / k_accumulate), which is fp64 with explicitly unfused multiply/add so the sums
// come out bit-identical to the reference's x86-64 build (no FMA, Makefile:5 of the reference).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "cfd_tables.h"

namespace issl {

constexpr int kScanThreads = 256;      // threads per scan CTA
constexpr uint32_t kListAlign = 32;    // list starts are padded to 32 entries (128 B of residuals)
constexpr uint32_t kChunkQuantum = 4096;   // scan-item sizes are multiples of this many entries

enum Layout : int { kRes32 = 1, kSig64 = 2, kGather = 3 };

// The index as it lies in HBM.  Position space: p in [0, P) enumerates the entries of all slice
// lists, slice-major then list value then ascending site id -- the order in which the reference
// walks them for one guide (ref isslScoreOfftargets.cpp:330-344) -- with every list start rounded
// up to kListAlign entries.
struct IndexView {
    const uint64_t *sig;        // [N]  packed site signatures (ref offtargets[], :200-204)
    const uint32_t *occ;        // [N]  occurrences per site (ref: high half of every list entry, :348)
    const uint32_t *ids;        // [P]  site id per list position (ref: low half of the entry, :347)
    const uint32_t *res32;      // [P]  kRes32: signature with the slice's known bits removed
    const uint64_t *sig64;      // [P]  kSig64: signature inline
    const uint64_t *listStart;  // [nLists] first position of list (slice * sliceLimit + value)
    const uint64_t *listLen;    // [nLists] ref allSlicelistSizes, :221-226
    uint64_t N, P;
    uint32_t seqLength, sliceWidth, sliceCount, sliceLimit;
    uint32_t sliceMask;         // 2^sliceWidth - 1
    uint32_t knownBits;         // min(sliceWidth, 8): bits every member of a list shares with the list value
    int layout;
};

// One unit of scan work: `count` consecutive list positions starting at p0, tested against a GROUP
// of up to kMaxGroup guides that all selected this list (same slice, same slice value).  The chunk
// is streamed from HBM once and every guide of the group is tested against it from registers.
struct ScanItem {
    uint64_t p0Slice;     // bits 0..39 position p0, bits 40..47 group size (1..8), bits 56..63 slice index
    uint32_t count;
    uint32_t groupStart;  // index of the group's first member in the list-sorted guide index array
};
constexpr int kMaxGroup = 8;       // largest group of the POPC path (guides held in registers)
constexpr int kBigGroup = 32;      // group of the bit-sliced path (one bit per guide in every word)
constexpr int kBigGroupMin = 14;   // a remainder of at least this many guides still gets a bit-sliced block

__constant__ double c_cfdPos[320];
__constant__ double c_cfdPam[16];
__device__ double g_cfdPos[320];   // the same table in global memory: survivors index it divergently, which the
                                   // constant cache serialises; through L1 it is one transaction per distinct line

// ------------------------------------------------------------------------------------------------
// bit helpers
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t remove_bits(uint64_t x, uint32_t off, uint32_t nbits)
{
    const uint64_t low = x & ((1ull << off) - 1ull);
    const uint64_t high = (off + nbits >= 64) ? 0ull : (x >> (off + nbits));
    return low | (high << off);
}

__host__ __device__ __forceinline__ uint64_t insert_bits(uint64_t x, uint32_t off, uint32_t nbits, uint64_t v)
{
    const uint64_t low = x & ((1ull << off) - 1ull);
    const uint64_t high = x >> off;
    return low | (v << off) | ((off + nbits >= 64) ? 0ull : (high << (off + nbits)));
}

// per-base mismatch flags, ref :376-379: ((x & 0xAAAA..) >> 1) | (x & 0x5555..)
__device__ __forceinline__ uint64_t mismatch_mask64(uint64_t x)
{
    return (x | (x >> 1)) & 0x5555555555555555ull;
}
__device__ __forceinline__ int distance32(uint32_t x)
{
    return __popc((x | (x >> 1)) & 0x55555555u);
}
__device__ __forceinline__ int distance64(uint64_t x)
{
    return __popcll(mismatch_mask64(x));
}

// Does list (slice k, value of the guide at slice k) contain this site?  The builder files a site
// under `uint8_t sliceVal` (ref isslCreateIndex.cpp:228), i.e. under the LOW 8 BITS of its slice
// value; the scorer looks a guide up under its full slice value (ref isslScoreOfftargets.cpp:336-341).
__device__ __forceinline__ bool in_list(uint64_t siteSig, uint64_t guideSig, uint32_t k, uint32_t w, uint32_t smask)
{
    const uint32_t sv = (uint32_t)(siteSig >> (w * k)) & smask & 0xFFu;
    const uint32_t gv = (uint32_t)(guideSig >> (w * k)) & smask;
    return sv == gv;
}

// ------------------------------------------------------------------------------------------------
// K1: candidate scan.  ref isslScoreOfftargets.cpp:344-390 (the inner hot loop) for one
// (group of guides, slice, chunk of the list they all selected).
//
// Streams the chunk with coalesced 16-byte loads (4 residuals / 2 signatures / 4 ids per load,
// 2-4 loads in flight per thread), XOR + fold + popcount against every guide of the group held in
// registers, keeps dist <= maxDist.
// De-duplication is stateless: a site reached through several slices is counted only in the
// lowest slice whose looked-up list contains it (in_list), which is where the reference's toggle
// bitset (:385-390, :463) lets it through, because every list holds each site at most once and
// the distance does not depend on the slice.  Survivors (~4e-5 of candidates on a uniform
// genome) are appended to a key buffer: key = guide << pbits | position.  Sorting the keys
// restores the reference's visiting order for every guide.
// ------------------------------------------------------------------------------------------------
struct ScanArgs {
    IndexView iv;
    const ScanItem *items;
    const uint64_t *guides;        // the batch's packed guides
    const uint32_t *sortedGuide;   // guide indices sorted by the list they select
    uint64_t *hitKeys;
    unsigned long long *hitCount;
    uint64_t hitCap;
    int maxDist;
    int pbits;
};

__device__ __forceinline__ void emit_hit(const ScanArgs &a, uint64_t guideSig, uint64_t siteSig, uint32_t slice,
                                         uint32_t guide, uint64_t pos)
{
    for (uint32_t k = 0; k < slice; k++)
        if (in_list(siteSig, guideSig, k, a.iv.sliceWidth, a.iv.sliceMask)) return;
    const unsigned long long slot = atomicAdd(a.hitCount, 1ull);
    if (slot < a.hitCap) a.hitKeys[slot] = ((uint64_t)guide << a.pbits) | pos;
}

// Rare path (about one 16-byte vector in 800 per guide on a uniform genome): some candidate of a
// vector is within maxDist of some guide of the group.  Re-tests every (candidate, guide) pair of
// the vector exactly and emits the survivors.  Kept out of line so that the streaming loop stays
// small; everything it needs is re-read from global memory.
template <int LAYOUT>
__device__ __noinline__ void scan_slow(const ScanArgs &a, uint32_t slice, uint32_t groupStart, uint32_t groupSize,
                                       uint64_t pos0, int nCand, uint64_t c0, uint64_t c1, uint64_t c2, uint64_t c3)
{
    const uint32_t off = a.iv.sliceWidth * slice, kb = a.iv.knownBits;
    const uint64_t cand[4] = {c0, c1, c2, c3};
    for (uint32_t j = 0; j < groupSize; j++) {
        const uint32_t guide = a.sortedGuide[groupStart + j];
        const uint64_t g = a.guides[guide];
        for (int c = 0; c < nCand; c++) {
            uint64_t site = cand[c];
            if (LAYOUT == kRes32) site = insert_bits(site, off, kb, (g >> off) & ((1ull << kb) - 1ull));
            if (distance64(site ^ g) <= a.maxDist) emit_hit(a, g, site, slice, guide, pos0 + c);
        }
    }
}

// Streaming loop for one group class (G = 1, 2, 4, 8 guides held in registers).
//
// kRes32: 4 candidates per 16-byte load.  G <= 2 uses x = c ^ g; popc((x | x >> 1) & 0x5555..);
// G >= 4 pre-splits candidate and guides into their even-bit and odd-bit planes once
// (cE = c & M, cO = (c >> 1) & M), so that one pair costs two LOP3 and one POPC:
// popc((cE ^ gE) | (cO ^ gO)).  The minimum distance of all pairs of a vector is reduced with
// integer min and tested with a single branch.
template <int LAYOUT, int G>
__device__ __forceinline__ void scan_body(const ScanArgs &a, const ScanItem &it)
{
    constexpr uint32_t M32 = 0x55555555u;
    constexpr uint64_t M64 = 0x5555555555555555ull;
    constexpr int U = (G >= 4) ? 2 : 4;          // 16-byte loads in flight per thread
    const uint32_t slice = (uint32_t)(it.p0Slice >> 56);
    const uint32_t groupSize = (uint32_t)(it.p0Slice >> 40) & 0xFFu;
    const uint64_t p0 = it.p0Slice & ((1ull << 40) - 1ull);
    const uint32_t n = it.count, tid = threadIdx.x;
    const int maxDist = a.maxDist;
    const uint32_t off = a.iv.sliceWidth * slice, kb = a.iv.knownBits;

    if (LAYOUT == kRes32) {
        uint32_t gE[G], gO[G];
#pragma unroll
        for (int j = 0; j < G; j++) {
            const uint32_t member = (uint32_t)j < groupSize ? (uint32_t)j : 0u;   // padding slots repeat member 0
            const uint32_t r = (uint32_t)remove_bits(a.guides[a.sortedGuide[it.groupStart + member]], off, kb);
            if (G >= 4) { gE[j] = r & M32; gO[j] = (r >> 1) & M32; } else { gE[j] = r; gO[j] = 0; }
        }
        auto vec_min = [&](const uint4 &r) {
            const uint32_t c[4] = {r.x, r.y, r.z, r.w};
            int m = 64;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (G >= 4) {
                    const uint32_t cE = c[k] & M32, cO = (c[k] >> 1) & M32;
#pragma unroll
                    for (int j = 0; j < G; j++) m = min(m, __popc((cE ^ gE[j]) | (cO ^ gO[j])));
                } else {
#pragma unroll
                    for (int j = 0; j < G; j++) { const uint32_t x = c[k] ^ gE[j]; m = min(m, __popc((x | (x >> 1)) & M32)); }
                }
            }
            return m;
        };
        const uint4 *__restrict__ v4 = reinterpret_cast<const uint4 *>(a.iv.res32 + p0);
        const uint32_t nvec = n >> 2;
        uint32_t i = tid;
        for (; i + (U - 1) * kScanThreads < nvec; i += U * kScanThreads) {
            uint4 r[U];
#pragma unroll
            for (int u = 0; u < U; u++) r[u] = __ldcs(v4 + i + u * kScanThreads);
#pragma unroll
            for (int u = 0; u < U; u++)
                if (vec_min(r[u]) <= maxDist)
                    scan_slow<kRes32>(a, slice, it.groupStart, groupSize, p0 + 4ull * (i + u * kScanThreads), 4,
                                      r[u].x, r[u].y, r[u].z, r[u].w);
        }
        for (; i < nvec; i += kScanThreads) {
            const uint4 r = __ldcs(v4 + i);
            if (vec_min(r) <= maxDist)
                scan_slow<kRes32>(a, slice, it.groupStart, groupSize, p0 + 4ull * i, 4, r.x, r.y, r.z, r.w);
        }
        if (tid == 0 && (n & 3u)) {
            const uint64_t pos = p0 + 4ull * nvec;
            uint64_t c[4] = {0, 0, 0, 0};
            for (uint32_t k = 0; k < (n & 3u); k++) c[k] = a.iv.res32[pos + k];
            scan_slow<kRes32>(a, slice, it.groupStart, groupSize, pos, (int)(n & 3u), c[0], c[1], c[2], c[3]);
        }
    } else {
        // 64-bit signatures: inline (kSig64, 2 per 16-byte load) or gathered through ids (kGather, the
        // reference's own access pattern, isslScoreOfftargets.cpp:346-376).  Two POPC per pair.
        uint64_t gE[G], gO[G];
#pragma unroll
        for (int j = 0; j < G; j++) {
            const uint32_t member = (uint32_t)j < groupSize ? (uint32_t)j : 0u;
            const uint64_t g = a.guides[a.sortedGuide[it.groupStart + member]];
            gE[j] = g & M64; gO[j] = (g >> 1) & M64;
        }
        auto pair_min = [&](uint64_t s, int m) {
            const uint64_t cE = s & M64, cO = (s >> 1) & M64;
#pragma unroll
            for (int j = 0; j < G; j++) m = min(m, __popcll((cE ^ gE[j]) | (cO ^ gO[j])));
            return m;
        };
        if (LAYOUT == kSig64) {
            const ulonglong2 *__restrict__ v2 = reinterpret_cast<const ulonglong2 *>(a.iv.sig64 + p0);
            const uint32_t nvec = n >> 1;
            uint32_t i = tid;
            for (; i + (U - 1) * kScanThreads < nvec; i += U * kScanThreads) {
                ulonglong2 r[U];
#pragma unroll
                for (int u = 0; u < U; u++) r[u] = __ldcs(v2 + i + u * kScanThreads);
#pragma unroll
                for (int u = 0; u < U; u++)
                    if (pair_min(r[u].y, pair_min(r[u].x, 64)) <= maxDist)
                        scan_slow<kSig64>(a, slice, it.groupStart, groupSize, p0 + 2ull * (i + u * kScanThreads), 2,
                                          r[u].x, r[u].y, 0, 0);
            }
            for (; i < nvec; i += kScanThreads) {
                const ulonglong2 r = __ldcs(v2 + i);
                if (pair_min(r.y, pair_min(r.x, 64)) <= maxDist)
                    scan_slow<kSig64>(a, slice, it.groupStart, groupSize, p0 + 2ull * i, 2, r.x, r.y, 0, 0);
            }
            if (tid == 0 && (n & 1u))
                scan_slow<kSig64>(a, slice, it.groupStart, groupSize, p0 + 2ull * nvec, 1, a.iv.sig64[p0 + 2ull * nvec], 0, 0, 0);
        } else {
            const uint4 *__restrict__ v4 = reinterpret_cast<const uint4 *>(a.iv.ids + p0);
            const uint64_t *__restrict__ sig = a.iv.sig;
            const uint32_t nvec = n >> 2;
            uint32_t i = tid;
            for (; i + (U - 1) * kScanThreads < nvec; i += U * kScanThreads) {
                uint4 r[U];
#pragma unroll
                for (int u = 0; u < U; u++) r[u] = __ldcs(v4 + i + u * kScanThreads);
                uint64_t s[U][4];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    s[u][0] = __ldg(sig + r[u].x); s[u][1] = __ldg(sig + r[u].y);
                    s[u][2] = __ldg(sig + r[u].z); s[u][3] = __ldg(sig + r[u].w);
                }
#pragma unroll
                for (int u = 0; u < U; u++)
                    if (pair_min(s[u][3], pair_min(s[u][2], pair_min(s[u][1], pair_min(s[u][0], 64)))) <= maxDist)
                        scan_slow<kGather>(a, slice, it.groupStart, groupSize, p0 + 4ull * (i + u * kScanThreads), 4,
                                           s[u][0], s[u][1], s[u][2], s[u][3]);
            }
            for (; i < nvec; i += kScanThreads) {
                const uint4 r = __ldcs(v4 + i);
                const uint64_t s0 = __ldg(sig + r.x), s1 = __ldg(sig + r.y), s2 = __ldg(sig + r.z), s3 = __ldg(sig + r.w);
                if (pair_min(s3, pair_min(s2, pair_min(s1, pair_min(s0, 64)))) <= maxDist)
                    scan_slow<kGather>(a, slice, it.groupStart, groupSize, p0 + 4ull * i, 4, s0, s1, s2, s3);
            }
            if (tid == 0 && (n & 3u)) {
                const uint64_t pos = p0 + 4ull * nvec;
                uint64_t c[4] = {0, 0, 0, 0};
                for (uint32_t k = 0; k < (n & 3u); k++) c[k] = sig[a.iv.ids[pos + k]];
                scan_slow<kGather>(a, slice, it.groupStart, groupSize, pos, (int)(n & 3u), c[0], c[1], c[2], c[3]);
            }
        }
    }
}

// Bit-sliced path (kRes32, maxDist <= 7, groups of 9..32 guides): no POPC at all.
//
// Every 32-bit word holds one bit per GUIDE of the group.  For each pair of adjacent residual
// bases (a nibble of the candidate, 8 nibbles) a 16-entry shared-memory table gives, for the 16
// possible nibble values, the two words "guide j differs from this base" -- 1 KB per group, built
// once per scan item.  A thread handles one candidate at a time: 8 conflict-free LDS.64 (16
// entries x 8 B = one bank row, equal entries broadcast) yield the 16 per-base mismatch words, a
// carry-save adder tree (11 full adders = 22 LOP3) counts them per bit lane, and 4 more LOP3
// give the lanes whose count is <= 4.  24 + 16 ALU-pipe instructions test 32 (guide, candidate)
// pairs, against 32 POPC (XU pipe, 16 lanes/clk) + 80 ALU for the same pairs on the register path.
// The accept word is exact for maxDist = 4..7 and a superset for maxDist < 4; survivors are re-tested.
struct BitSliceTable { uint2 e[8][16]; };

__device__ __forceinline__ void full_add(uint32_t a, uint32_t b, uint32_t c, uint32_t &sum, uint32_t &carry)
{
    sum = a ^ b ^ c;
    carry = (a & b) | (c & (a ^ b));
}

// MD selects the threshold the accept word encodes: count <= 4 (used for every maxDist <= 4), <= 5, <= 6, <= 7.
template <int MD>
__device__ __forceinline__ uint32_t accept_le(const BitSliceTable &tb, uint32_t c)
{
    uint32_t f[16];
#pragma unroll
    for (int p = 0; p < 8; p++) {
        // (measured: issuing the shift as IMAD.HI on the FMA pipe instead of SHF is slower, 102.5 vs 98.3 ms/step)
        const uint2 e = tb.e[p][(c >> (4 * p)) & 15u];
        f[2 * p] = e.x; f[2 * p + 1] = e.y;
    }
    uint32_t s0, s1, s2, s3, s4, k0, k1, k2, k3, k4, k5, k6, t0, t1, u0, u1, q0, q1, q2, m;
    full_add(f[0], f[1], f[2], s0, k0);
    full_add(f[3], f[4], f[5], s1, k1);
    full_add(f[6], f[7], f[8], s2, k2);
    full_add(f[9], f[10], f[11], s3, k3);
    full_add(f[12], f[13], f[14], s4, k4);
    full_add(s0, s1, s2, t0, k5);          // weight-1 bits left: t0, t1
    full_add(s3, s4, f[15], t1, k6);
    full_add(k0, k1, k2, u0, q0);          // weight 2
    full_add(k3, k4, k5, u1, q1);
    full_add(u0, u1, k6, m, q2);           // weight-2 bit left: m; weight-4 bits: q0, q1, q2
    // count = low + 4 (q0 + q1 + q2) with low = t0 + t1 + 2 m in 0..4.  Two q bits set means count >= 8: always
    // rejected.  With exactly one q bit set, count = 4 + low is rejected when low exceeds MD - 4.
    const uint32_t two = (q0 & q1) | (q2 & (q0 ^ q1)), any = q0 | q1 | q2;
    uint32_t lowTooBig;
    if (MD <= 4) lowTooBig = t0 | t1 | m;               // low >= 1
    else if (MD == 5) lowTooBig = m | (t0 & t1);        // low >= 2
    else if (MD == 6) lowTooBig = m & (t0 | t1);        // low >= 3
    else lowTooBig = m & t0 & t1;                       // low >= 4  (MD == 7)
    return ~(two | (any & lowTooBig));
}

__device__ __noinline__ void scan_emit32(const ScanArgs &a, uint32_t slice, uint32_t groupStart, uint32_t accept,
                                         uint32_t residual, uint64_t pos)
{
    const uint32_t off = a.iv.sliceWidth * slice, kb = a.iv.knownBits;
    while (accept) {
        const uint32_t j = __ffs(accept) - 1;
        accept &= accept - 1;
        const uint32_t guide = a.sortedGuide[groupStart + j];
        const uint64_t g = a.guides[guide];
        const uint64_t site = insert_bits(residual, off, kb, (g >> off) & ((1ull << kb) - 1ull));
        if (distance64(site ^ g) <= a.maxDist) emit_hit(a, g, site, slice, guide, pos);
    }
}

template <int MD>
__device__ __forceinline__ void scan_body_bitsliced(const ScanArgs &a, const ScanItem &it, BitSliceTable &tb, uint32_t *gres)
{
    const uint32_t slice = (uint32_t)(it.p0Slice >> 56);
    const uint32_t groupSize = (uint32_t)(it.p0Slice >> 40) & 0xFFu;
    const uint64_t p0 = it.p0Slice & ((1ull << 40) - 1ull);
    const uint32_t n = it.count, tid = threadIdx.x;
    const uint32_t off = a.iv.sliceWidth * slice, kb = a.iv.knownBits;

    if (tid < 32) gres[tid] = tid < groupSize ? (uint32_t)remove_bits(a.guides[a.sortedGuide[it.groupStart + tid]], off, kb) : 0u;
    __syncthreads();
    if (tid < 128) {
        const uint32_t p = tid >> 4, nib = tid & 15u;
        uint32_t w0 = 0, w1 = 0;
        for (uint32_t j = 0; j < 32; j++) {
            const uint32_t gn = (gres[j] >> (4 * p)) & 15u;
            const bool unused = j >= groupSize;            // unused lanes always mismatch, so they never accept
            w0 |= (uint32_t)(unused || ((gn & 3u) != (nib & 3u))) << j;
            w1 |= (uint32_t)(unused || ((gn >> 2) != (nib >> 2))) << j;
        }
        tb.e[p][nib] = make_uint2(w0, w1);
    }
    __syncthreads();

    const uint4 *__restrict__ v4 = reinterpret_cast<const uint4 *>(a.iv.res32 + p0);
    const uint32_t nvec = n >> 2;
    auto vec = [&](const uint4 &r, uint64_t pos) {
        const uint32_t a0 = accept_le<MD>(tb, r.x), a1 = accept_le<MD>(tb, r.y), a2 = accept_le<MD>(tb, r.z), a3 = accept_le<MD>(tb, r.w);
        if (a0 | a1 | a2 | a3) {
            if (a0) scan_emit32(a, slice, it.groupStart, a0, r.x, pos);
            if (a1) scan_emit32(a, slice, it.groupStart, a1, r.y, pos + 1);
            if (a2) scan_emit32(a, slice, it.groupStart, a2, r.z, pos + 2);
            if (a3) scan_emit32(a, slice, it.groupStart, a3, r.w, pos + 3);
        }
    };
    uint32_t i = tid;
    for (; i + kScanThreads < nvec; i += 2 * kScanThreads) {
        const uint4 r0 = __ldcs(v4 + i), r1 = __ldcs(v4 + i + kScanThreads);
        vec(r0, p0 + 4ull * i);
        vec(r1, p0 + 4ull * (i + kScanThreads));
    }
    for (; i < nvec; i += kScanThreads) {
        const uint4 r = __ldcs(v4 + i);
        vec(r, p0 + 4ull * i);
    }
    if (tid < (n & 3u)) {
        const uint64_t pos = p0 + 4ull * nvec + tid;
        const uint32_t r = a.iv.res32[pos];
        const uint32_t acc = accept_le<MD>(tb, r);
        if (acc) scan_emit32(a, slice, it.groupStart, acc, r, pos);
    }
}

template <int LAYOUT>
__global__ void __launch_bounds__(kScanThreads) k_scan(const ScanArgs a)
{
    const ScanItem it = a.items[blockIdx.x];
    const uint32_t groupSize = (uint32_t)(it.p0Slice >> 40) & 0xFFu;
    if (LAYOUT == kRes32) {
        __shared__ BitSliceTable tb;
        __shared__ uint32_t gres[32];
        if (groupSize > kMaxGroup) {
            if (a.maxDist <= 4) scan_body_bitsliced<4>(a, it, tb, gres);
            else if (a.maxDist == 5) scan_body_bitsliced<5>(a, it, tb, gres);
            else if (a.maxDist == 6) scan_body_bitsliced<6>(a, it, tb, gres);
            else scan_body_bitsliced<7>(a, it, tb, gres);
            return;
        }
    }
    if (groupSize > 4) scan_body<LAYOUT, 8>(a, it);
    else if (groupSize > 2) scan_body<LAYOUT, 4>(a, it);
    else if (groupSize == 2) scan_body<LAYOUT, 2>(a, it);
    else scan_body<LAYOUT, 1>(a, it);
}

// ------------------------------------------------------------------------------------------------
// Scan-item construction (ref :330-341: slice value -> list -> length), with guides grouped by
// the list they select so that a list chunk is streamed once per group instead of once per guide.
// One "pair" = (guide, slice) for slices in [slice0, slice0 + nSlices).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t pair_list(const IndexView &iv, uint64_t g, uint32_t slice)
{
    const uint32_t v = (uint32_t)(g >> (iv.sliceWidth * slice)) & iv.sliceMask;
    return (uint64_t)slice * iv.sliceLimit + v;
}

// pass 1: sort key (list id; nLists = "no work") and value (guide index) of every pair, plus the
// total number of candidates of the wave = the reported unit count (entries the reference visits).
__global__ void k_pair_keys(IndexView iv, const uint64_t *guides, const uint8_t *done, uint32_t nGuides,
                            uint32_t slice0, uint32_t nSlices, uint32_t nLists, uint32_t *keys, uint32_t *vals,
                            unsigned long long *total)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long len = 0;
    if (t < (uint64_t)nGuides * nSlices) {
        const uint32_t gi = (uint32_t)(t / nSlices), s = slice0 + (uint32_t)(t % nSlices);
        uint32_t key = nLists;
        if (!done || !done[gi]) {
            const uint64_t list = pair_list(iv, guides[gi], s);
            len = iv.listLen[list];
            if (len) key = (uint32_t)list;
        }
        keys[t] = key;
        vals[t] = gi;
    }
    for (int o = 16; o > 0; o >>= 1) len += __shfl_down_sync(0xffffffffu, len, o);
    if ((threadIdx.x & 31) == 0 && len) atomicAdd(total, len);
}

// Group shape of sorted pair i: runs of equal list id are cut into groups of 8; a remainder r is
// cut as 7..8 -> one group, 5..6 -> 4 + (r - 4), 1..4 -> one group.  Returns the group size when i
// is the first member of a group, else 0.
__device__ __forceinline__ uint32_t group_head_size(uint32_t rank, uint32_t runLen, uint32_t maxGroup)
{
    if (maxGroup >= (uint32_t)kBigGroup) {   // bit-sliced blocks of up to 32 first
        const uint32_t rem = runLen % kBigGroup;
        const uint32_t bigEnd = runLen - rem + (rem >= (uint32_t)kBigGroupMin ? rem : 0u);
        if (rank < bigEnd) return (rank % kBigGroup) == 0 ? min((uint32_t)kBigGroup, bigEnd - rank) : 0u;
        rank -= bigEnd; runLen -= bigEnd;
    } else if (maxGroup < 8) {
        return (rank % maxGroup) == 0 ? min(maxGroup, runLen - rank) : 0u;
    }
    const uint32_t base = runLen & ~7u, r = runLen - base;
    if (rank < base) return (rank & 7u) == 0 ? 8u : 0u;
    const uint32_t q = rank - base;
    if (q == 0) return (r == 5 || r == 6) ? 4u : r;
    if (q == 4 && (r == 5 || r == 6)) return r - 4;
    return 0u;
}

__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t *keys, uint32_t n, uint32_t key)
{
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// pass 2: items per sorted pair (non-zero only at group heads)
__global__ void k_group_count(IndexView iv, const uint32_t *sortedKeys, uint32_t nPairs, uint32_t nLists, uint32_t chunk,
                              uint32_t maxGroup, uint32_t *counts, unsigned long long *streamed)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nPairs) return;
    uint32_t c = 0;
    const uint32_t key = sortedKeys[i];
    if (key < nLists) {
        const uint32_t lb = lower_bound_u32(sortedKeys, nPairs, key), ub = lower_bound_u32(sortedKeys, nPairs, key + 1);
        if (group_head_size(i - lb, ub - lb, maxGroup)) {
            const uint64_t len = iv.listLen[key];
            c = (uint32_t)((len + chunk - 1) / chunk);
            atomicAdd(streamed, (unsigned long long)len);
        }
    }
    counts[i] = c;
}

// pass 3: write the items of every group at its scanned offset
__global__ void k_group_fill(IndexView iv, const uint32_t *sortedKeys, uint32_t nPairs, uint32_t nLists, uint32_t chunk,
                             uint32_t maxGroup, const uint32_t *offsets, ScanItem *items)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nPairs) return;
    const uint32_t key = sortedKeys[i];
    if (key >= nLists) return;
    const uint32_t lb = lower_bound_u32(sortedKeys, nPairs, key), ub = lower_bound_u32(sortedKeys, nPairs, key + 1);
    const uint32_t size = group_head_size(i - lb, ub - lb, maxGroup);
    if (!size) return;
    const uint64_t len = iv.listLen[key], start = iv.listStart[key];
    const uint64_t slice = key / iv.sliceLimit;
    uint32_t o = offsets[i];
    for (uint64_t c0 = 0; c0 < len; c0 += chunk, o++) {
        ScanItem it;
        it.p0Slice = (start + c0) | ((uint64_t)size << 40) | (slice << 56);
        it.count = (uint32_t)((len - c0 < chunk) ? (len - c0) : chunk);
        it.groupStart = i;
        items[o] = it;
    }
}

// ------------------------------------------------------------------------------------------------
// K2a: score every survivor.  ref :392-461.
// MIT: table[mismatch mask] * occ when dist > 0 (missing mask -> 0.0, the inserting operator[] at :394).
// CFD: 1 when dist == 0, else PAM(GG) * prod over mismatching positions < 20, ascending (:411-458); * occ.
// ------------------------------------------------------------------------------------------------
// The score tables of one index.  mitDense (seqLength <= 20 only) is the file's table spread over all 2^20
// position sets -- one load instead of a binary search; absent masks hold 0.0, which is what the reference's
// inserting operator[] yields for them (:394).
struct ScoreTables {
    const uint64_t *mitMasks;   // sorted ascending
    const double *mitScores;
    uint32_t mitCount;
    const double *mitDense;     // [2^20] or nullptr
};

// bits 0, 2, 4, ... of x packed into the low half
__device__ __forceinline__ uint32_t compress_even_bits(uint64_t x)
{
    x &= 0x5555555555555555ull;
    x = (x | (x >> 1)) & 0x3333333333333333ull;
    x = (x | (x >> 2)) & 0x0F0F0F0F0F0F0F0Full;
    x = (x | (x >> 4)) & 0x00FF00FF00FF00FFull;
    x = (x | (x >> 8)) & 0x0000FFFF0000FFFFull;
    x = (x | (x >> 16)) & 0x00000000FFFFFFFFull;
    return (uint32_t)x;
}

// local MIT and CFD contribution of one scored site, ref :392-461, both already multiplied by the occurrences
__device__ __forceinline__ void hit_contrib(const ScoreTables &tb, uint64_t g, uint64_t site, uint32_t occ, int calcMit,
                                            int calcCfd, double &cm, double &cc, int &dist)
{
    const uint64_t mm = mismatch_mask64(g ^ site);
    dist = __popcll(mm);
    const double docc = (double)occ;
    uint32_t m20 = compress_even_bits(mm);   // mismatch flags of positions 0..31, one bit each
    cm = 0.0; cc = 0.0;
    if (calcMit && dist > 0) {
        double s;
        if (tb.mitDense && (mm >> 40) == 0) {
            s = __ldg(tb.mitDense + m20);
        } else {
            uint32_t lo = 0, hi = tb.mitCount;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (tb.mitMasks[mid] < mm) lo = mid + 1; else hi = mid;
            }
            s = (lo < tb.mitCount && tb.mitMasks[lo] == mm) ? tb.mitScores[lo] : 0.0;
        }
        cm = __dmul_rn(s, docc);
    }
    if (calcCfd) {
        double cfd = 1.0;
        if (dist > 0) {
            cfd = c_cfdPam[10];   // 0b1010 = GG, ref :411
            m20 &= 0xFFFFFu;      // the reference's loop covers positions 0..19 (:413)
            while (m20) {
                const int p = __ffs(m20) - 1;
                m20 &= m20 - 1;
                const uint32_t gb = (uint32_t)(g >> (2 * p)) & 3u, ob = (uint32_t)(site >> (2 * p)) & 3u;
                cfd = __dmul_rn(cfd, __ldg(&g_cfdPos[(p << 4) | (gb << 2) | (ob ^ 3u)]));
            }
        }
        cc = __dmul_rn(cfd, docc);
    }
}

constexpr uint64_t kKeyOccursOnce = 1ull << 63;   // survivor key flag, above every sorted bit: occurrences = 1

struct ContribArgs {
    IndexView iv;
    const uint64_t *keys;     // sorted
    uint64_t nHits;
    const uint64_t *guides;
    ScoreTables tb;
    int pbits;
    int idInKey;              // TRIPLE: the key's low 32 bits are the site id itself (bits 32..34: slice), not a list position
    int calcMit, calcCfd;
    double *contribMit;       // [nHits]
    double *contribCfd;       // [nHits]
    uint32_t *hitId;          // optional dump
    int32_t *hitDist;
    uint32_t *hitOcc;
};

__global__ void __launch_bounds__(256) k_contrib(const ContribArgs a)
{
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.nHits) return;
    const uint64_t key = a.keys[j];
    const uint64_t pos = key & ((1ull << a.pbits) - 1ull);
    const uint64_t g = a.guides[(key & ~kKeyOccursOnce) >> a.pbits];
    const uint32_t id = a.idInKey ? (uint32_t)pos : a.iv.ids[pos];
    const uint64_t site = a.iv.sig[id];
    const uint32_t occ = (key & kKeyOccursOnce) ? 1u : a.iv.occ[id];
    double cm, cc;
    int dist;
    hit_contrib(a.tb, g, site, occ, a.calcMit, a.calcCfd, cm, cc, dist);
    a.contribMit[j] = cm;
    a.contribCfd[j] = cc;
    if (a.hitId) { a.hitId[j] = id; a.hitDist[j] = dist; a.hitOcc[j] = occ; }
}

// ------------------------------------------------------------------------------------------------
// K2b: per-guide ordered accumulation with the reference's early exit.  ref :322-327, :394, :460,
// :466-502.  One thread per guide walks its survivors in (slice, list position) order -- the
// sorted key order -- adding one rounded product at a time, exactly as the CPU does, and stops at
// the first survivor after which the method's predicate exceeds maximum_sum.
// ------------------------------------------------------------------------------------------------
struct AccumArgs {
    const uint64_t *keys;
    uint64_t nHits;
    const double *contribMit, *contribCfd;
    uint32_t nGuides;
    int pbits;
    int method;
    int checkExit;          // 0: maximum_sum is +inf/NaN or the method is unknown
    double maximumSum;
    double *totMit, *totCfd;   // [nGuides] running sums, carried across slice waves
    uint8_t *done;             // [nGuides] set when the guide exited early
    uint64_t *scoredEnd;       // optional [nGuides]: index one past the guide's last scored survivor
    uint64_t *segBegin;        // optional [nGuides]: index of the guide's first survivor of this wave
};

__device__ __forceinline__ uint64_t lower_bound_key(const uint64_t *keys, uint64_t n, uint64_t key)
{
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if ((keys[mid] & ~kKeyOccursOnce) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// the method's early-exit predicate, ref :466-496
__device__ __forceinline__ bool exit_predicate(int method, double mit, double cfd, double mx)
{
    switch (method) {
    case ISSL_METHOD_AND: return mit > mx && cfd > mx;
    case ISSL_METHOD_OR:  return mit > mx || cfd > mx;
    case ISSL_METHOD_AVG: return __ddiv_rn(__dadd_rn(mit, cfd), 2.0) > mx;
    case ISSL_METHOD_MIT: return mit > mx;
    case ISSL_METHOD_CFD: return cfd > mx;
    default: return false;
    }
}

__global__ void __launch_bounds__(128) k_accumulate(const AccumArgs a)
{
    const uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= a.nGuides) return;
    const uint64_t lo = lower_bound_key(a.keys, a.nHits, (uint64_t)gi << a.pbits);
    if (a.segBegin) a.segBegin[gi] = lo;
    if (a.done[gi]) { if (a.scoredEnd) a.scoredEnd[gi] = lo; return; }
    const uint64_t hi = lower_bound_key(a.keys, a.nHits, ((uint64_t)gi + 1) << a.pbits);
    double mit = a.totMit[gi], cfd = a.totCfd[gi];
    const double mx = a.maximumSum;
    uint64_t j = lo;
    bool stop = false;
    for (; j < hi && !stop; j++) {
        mit = __dadd_rn(mit, a.contribMit[j]);
        cfd = __dadd_rn(cfd, a.contribCfd[j]);
        if (a.checkExit) stop = exit_predicate(a.method, mit, cfd, mx);
    }
    a.totMit[gi] = mit;
    a.totCfd[gi] = cfd;
    if (stop) a.done[gi] = 1;
    if (a.scoredEnd) a.scoredEnd[gi] = j;
}

// ref :505-506
__global__ void k_finalize(const double *totMit, const double *totCfd, uint32_t n, double *mitOut, double *cfdOut)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (mitOut) mitOut[i] = __ddiv_rn(10000.0, __dadd_rn(100.0, totMit[i]));
    if (cfdOut) cfdOut[i] = __ddiv_rn(10000.0, __dadd_rn(100.0, totCfd[i]));
}

__global__ void k_count_done(const uint8_t *done, uint32_t n, unsigned long long *count)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned m = __ballot_sync(0xffffffffu, i < n && done[i]);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, (unsigned long long)__popc(m));
}

// ------------------------------------------------------------------------------------------------
// K3: one-time re-layout of the file's list section into position space, with validation of
// the builder's invariants (ref isslCreateIndex.cpp:216-234).
// entries[] holds file entries [q0, q0 + n) of ONE slice, preceded by entry q0-1 when hasPrev.
// errors[0] counts violations (id range, membership, ascending ids, occurrence mismatch).
// ------------------------------------------------------------------------------------------------
struct RelayoutArgs {
    IndexView iv;             // ids/res32/sig64/occ are written through const_cast'ed pointers
    const uint64_t *entries;
    const uint64_t *filePrefix;   // [nLists + 1] exclusive prefix of listLen in file order
    uint64_t q0, n;
    uint32_t slice;
    int hasPrev;
    unsigned long long *errors;
};

__global__ void __launch_bounds__(256) k_relayout(const RelayoutArgs a)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n) return;
    const uint64_t q = a.q0 + t;
    const uint64_t e = a.entries[t + (a.hasPrev ? 1 : 0)];
    const uint32_t id = (uint32_t)(e & 0xFFFFFFFFull), occ = (uint32_t)(e >> 32);
    // list of this file position: last list whose prefix <= q
    uint64_t lo = (uint64_t)a.slice * a.iv.sliceLimit, hi = lo + a.iv.sliceLimit;
    while (hi - lo > 1) {
        const uint64_t mid = (lo + hi) >> 1;
        if (a.filePrefix[mid] <= q) lo = mid; else hi = mid;
    }
    const uint64_t list = lo, j = q - a.filePrefix[list];
    const uint32_t value = (uint32_t)(list - (uint64_t)a.slice * a.iv.sliceLimit);
    bool bad = id >= a.iv.N;
    if (!bad) {
        const uint64_t s = a.iv.sig[id];
        const uint32_t off = a.iv.sliceWidth * a.slice;
        const uint32_t sv = (uint32_t)(s >> off) & a.iv.sliceMask & 0xFFu;
        bad |= (sv != value);
        if (j > 0) {
            const uint64_t prev = a.entries[t + (a.hasPrev ? 1 : 0) - 1];
            bad |= ((uint32_t)(prev & 0xFFFFFFFFull) >= id);
        }
        uint32_t *occOut = const_cast<uint32_t *>(a.iv.occ);
        if (a.slice == 0) occOut[id] = occ; else bad |= (occOut[id] != occ);
        if (a.iv.ids) {   // (nullptr: validation only -- TRIPLE builds its slice lists when something asks for them)
            const uint64_t p = a.iv.listStart[list] + j;
            const_cast<uint32_t *>(a.iv.ids)[p] = id;
            if (a.iv.layout == kRes32)
                const_cast<uint32_t *>(a.iv.res32)[p] = (uint32_t)remove_bits(s, off, a.iv.knownBits);
            else if (a.iv.layout == kSig64)
                const_cast<uint64_t *>(a.iv.sig64)[p] = s;
        }
    }
    if (bad) atomicAdd(a.errors, 1ull);
}

// Inverse of K3: file-order entries (occ << 32 | id) of positions [q0, q0 + n), for
// issl_device_write_issl (ref isslCreateIndex.cpp:285-289).
__global__ void k_export_entries(IndexView iv, const uint64_t *filePrefix, uint64_t nLists, uint64_t q0, uint64_t n,
                                 uint64_t *out)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint64_t q = q0 + t;
    uint64_t lo = 0, hi = nLists;
    while (hi - lo > 1) {
        const uint64_t mid = (lo + hi) >> 1;
        if (filePrefix[mid] <= q) lo = mid; else hi = mid;
    }
    const uint64_t p = iv.listStart[lo] + (q - filePrefix[lo]);
    const uint32_t id = iv.ids[p];
    out[t] = ((uint64_t)iv.occ[id] << 32) | id;
}

// ------------------------------------------------------------------------------------------------
// Index construction on the device (synthetic indexes; ref isslCreateIndex.cpp:184-234).
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x)
{   // splitmix64 finaliser
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ uint64_t rng3(uint64_t seed, uint64_t a, uint64_t b)
{
    return mix64(mix64(mix64(seed) ^ a) ^ (b * 0xD6E8FEB86659FD93ull));
}

// Sort key of a site: base 0 most significant, so that ascending keys = the lexicographic order of
// the text file the reference builder expects (extractOfftargets.py sorts strings).
__host__ __device__ __forceinline__ uint64_t sig_to_sortkey(uint64_t sig, uint32_t L)
{
    uint64_t k = 0;
    for (uint32_t j = 0; j < L; j++) k |= ((sig >> (2 * j)) & 3ull) << (2 * (L - 1 - j));
    return k;
}

// uniform sites: i.i.d. bases, first base in {A,C,G}; family members: a per-family random root, each base substituted
// with the family's rate (family f owns members [familyStart[f], familyStart[f + 1]): sizes are the caller's, fixed or
// log-uniform); low-complexity sites: windows that overlap a poly-A / poly-T / dinucleotide tract on one side (k = 10..20
// bases of the repeat, 2 % of them substituted, the rest random) -- they skew the slice lists towards a few values and
// produce sites that occur thousands of times, as the repeats of a real genome do.
struct SynthArgs {
    uint64_t seed, nUniform, nFamilySites, nLow;
    uint32_t families;
    const uint64_t *familyStart;   // [families + 1]
    double maxSubRate;
    uint32_t L;
    uint64_t *keys;
};

__global__ void k_synth_sites(const SynthArgs a)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t total = a.nUniform + a.nFamilySites + a.nLow;
    if (t >= total) return;
    const uint64_t seed = a.seed;
    const uint32_t L = a.L;
    const uint64_t lmask = (L >= 32) ? ~0ull : ((1ull << (2 * L)) - 1ull);
    uint64_t sig;
    if (t < a.nUniform) {
        const uint64_t r = rng3(seed, 1, t);
        sig = r & lmask & ~3ull;
        sig |= (rng3(seed, 2, t) % 3ull);
    } else if (t < a.nUniform + a.nFamilySites) {
        const uint64_t m = t - a.nUniform;
        uint32_t lo = 0, hi = a.families;   // the family that owns member m
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (a.familyStart[mid] <= m) lo = mid; else hi = mid;
        }
        const uint64_t f = lo;
        const uint64_t root = (rng3(seed, 3, f) & lmask & ~3ull) | (rng3(seed, 4, f) % 3ull);
        const double rate = a.maxSubRate * (double)(rng3(seed, 5, f) >> 11) * (1.0 / 9007199254740992.0);
        sig = root;
        for (uint32_t j = 0; j < L; j++) {
            const uint64_t r = rng3(seed, 6 + j, m);
            const double u = (double)(r >> 11) * (1.0 / 9007199254740992.0);
            if (u < rate) {
                const uint64_t b = ((sig >> (2 * j)) & 3ull), nb = (b + 1 + (r % 3ull)) & 3ull;
                sig = (sig & ~(3ull << (2 * j))) | (nb << (2 * j));
            }
        }
    } else {
        const uint64_t m = t - a.nUniform - a.nFamilySites;
        const uint64_t r = rng3(seed, 40, m);
        const uint32_t kind = (uint32_t)(r % 14ull);            // 0 poly-A, 1 poly-T, 2..13 the twelve dinucleotides XY, X != Y
        const uint32_t side = (uint32_t)(r >> 8) & 1u;          // tract at the start / at the end of the window
        const uint32_t k = 10u + (uint32_t)((r >> 16) % 11ull); // bases of the window inside the tract
        const uint32_t phase = (uint32_t)(r >> 32) & 1u;
        uint32_t x = 0, y = 0;
        if (kind == 1) { x = 3; y = 3; }
        else if (kind >= 2) { x = (kind - 2) / 3; y = (kind - 2) % 3; y += (y >= x); }
        sig = rng3(seed, 41, m) & lmask;
        for (uint32_t j = 0; j < L; j++) {
            const bool inTract = side ? (j + k >= L) : (j < k);
            if (!inTract) continue;
            uint64_t b = ((j + phase) & 1u) ? y : x;
            const uint64_t rs = rng3(seed, 42 + j, m);
            if (rs % 50ull == 0) b = (b + 1 + ((rs >> 8) % 3ull)) & 3ull;
            sig = (sig & ~(3ull << (2 * j))) | (b << (2 * j));
        }
        if ((sig & 3ull) == 3ull) sig = (sig & ~3ull) | (rng3(seed, 39, m) % 3ull);   // no site starts with T
    }
    a.keys[t] = sig_to_sortkey(sig, L);
}

// run heads of the sorted key array
__global__ void k_run_flags(const uint64_t *keys, uint64_t n, uint32_t *flags)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    flags[t] = (t == 0 || keys[t] != keys[t - 1]) ? 1u : 0u;
}

// scatter run heads: runStart[rank] = t, sig[rank] = signature (rank = inclusive scan of flags - 1)
__global__ void k_run_scatter(const uint64_t *keys, const uint32_t *flags, const uint64_t *rankIncl, uint64_t n,
                              uint32_t L, int valuesAreSortKeys, uint64_t *sig, uint64_t *runStart)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n || !flags[t]) return;
    const uint64_t r = rankIncl[t] - 1;
    runStart[r] = t;
    sig[r] = valuesAreSortKeys ? sig_to_sortkey(keys[t], L) : keys[t];   // the key transform is an involution (base order reversal)
}

// isslCreateIndex's input: fixed-width text lines (ref isslCreateIndex.cpp:39-47 packing, :158-161 table: anything
// but C, G, T packs as 0).  text holds n lines preceded by one more line when hasPrev.  flags[i] = 1 when line i
// differs from its predecessor as TEXT (the reference's memcmp at :192), which starts a new distinct site.
__global__ void k_pack_lines(const char *text, uint64_t n, int hasPrev, uint32_t L, uint64_t *sig, uint32_t *flags)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint64_t line = L + 1;
    const char *p = text + (t + (hasPrev ? 1 : 0)) * line;
    uint64_t s = 0;
    bool differs = !(hasPrev || t > 0);
    for (uint32_t j = 0; j < L; j++) {
        const char c = p[j];
        const uint64_t code = c == 'C' ? 1ull : c == 'G' ? 2ull : c == 'T' ? 3ull : 0ull;
        s |= code << (2 * j);
        if (hasPrev || t > 0) differs |= (c != p[(long long)j - (long long)line]);
    }
    sig[t] = s;
    flags[t] = differs ? 1u : 0u;
}

__global__ void k_run_lengths(const uint64_t *runStart, uint64_t nRuns, uint64_t n, uint32_t *occ)
{
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nRuns) return;
    const uint64_t end = (r + 1 < nRuns) ? runStart[r + 1] : n;
    occ[r] = (uint32_t)(end - runStart[r]);
}

// per-slice list value of every site: (sig >> w*slice) & mask, truncated to 8 bits (ref :228)
__global__ void k_slice_values(const uint64_t *sig, uint64_t n, uint32_t w, uint32_t smask, uint32_t slice,
                               uint8_t *values, uint32_t *ids, unsigned long long *hist)
{
    __shared__ unsigned int h[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) h[i] = 0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
        const uint8_t v = (uint8_t)((uint32_t)(sig[t] >> (w * slice)) & smask);
        values[t] = v;
        ids[t] = (uint32_t)t;
        atomicAdd(&h[v], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (h[i]) atomicAdd(&hist[i], (unsigned long long)h[i]);
}

// place the stably sorted ids of one slice into position space
__global__ void k_place_slice(IndexView iv, const uint8_t *sortedValues, const uint32_t *sortedIds, uint64_t n,
                              uint32_t slice, const uint64_t *valueStart /* [256] rank of first id per value */)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint32_t v = sortedValues[t], id = sortedIds[t];
    const uint64_t list = (uint64_t)slice * iv.sliceLimit + v;
    const uint64_t p = iv.listStart[list] + (t - valueStart[v]);
    const uint64_t s = iv.sig[id];
    const_cast<uint32_t *>(iv.ids)[p] = id;
    if (iv.layout == kRes32)
        const_cast<uint32_t *>(iv.res32)[p] = (uint32_t)remove_bits(s, iv.sliceWidth * slice, iv.knownBits);
    else if (iv.layout == kSig64)
        const_cast<uint64_t *>(iv.sig64)[p] = s;
}

// Guide-side pre-filters of the pipeline (ref src/crackling/Crackling.py:312-384) + packing of target23[0:20]
// (ref :747-752 and isslScoreOfftargets.cpp:63-71, :99-102), one thread per 24-byte line.
__global__ void k_guide_filters(const char *text, uint64_t n, uint8_t *flags, double *at, uint64_t *packed)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const char *t = text + i * 24;
    uint32_t f = 0, atCount = 0, run = 0;
    uint64_t sig = 0;
    bool tttt = false;
    for (int j = 0; j < 23; j++) {
        const char c = t[j];
        if (j < 20) {
            atCount += (c == 'A' || c == 'T');
            sig |= (uint64_t)(c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 0) << (2 * j);
        }
        run = (c == 'T') ? run + 1 : 0;
        tttt |= run >= 4;
    }
    if (t[19] != 'G') f |= ISSL_FILTER_G20;
    if ((t[21] == 'G' && t[22] == 'G' && t[0] == 'T') || (t[0] == 'C' && t[1] == 'C' && t[22] == 'A')) f |= ISSL_FILTER_LEADING_T;
    const double pct = __ddiv_rn(__dmul_rn(100.0, (double)atCount), 20.0);
    if (pct < 20.0 || pct > 65.0) f |= ISSL_FILTER_AT;
    if (tttt) f |= ISSL_FILTER_TTTT;
    if (flags) flags[i] = (uint8_t)f;
    if (at) at[i] = pct;
    if (packed) packed[i] = sig;
}

// Duplicate candidate guides, ref /root/reference/src/crackling/Crackling.py:211-240 (the first occurrence of a 23-mer is
// recorded, every later one is dropped and marks the sequence as not unique) and :291-296 (isUnique = rejected).
// k_pack_targets: 23 bases -> 46-bit key (2 bits per base, base 0 most significant; anything but C, G, T packs as A, as
// issl_pack_guides does) + the target's index; after a STABLE sort by key, k_mark_duplicates reads runs of equal keys:
// the head of a run is the occurrence seen first.
__global__ void k_pack_targets(const char *text, uint64_t n, uint64_t *keys, uint32_t *idx)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const char *t = text + i * 24;
    uint64_t k = 0;
    for (int j = 0; j < 23; j++) {
        const char c = t[j];
        k = (k << 2) | (uint64_t)(c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 0);
    }
    keys[i] = k;
    idx[i] = (uint32_t)i;
}

__global__ void k_mark_duplicates(const uint64_t *sortedKeys, const uint32_t *sortedIdx, uint64_t n, uint8_t *flags,
                                  unsigned long long *counters /* [0] later occurrences, [1] sequences seen more than once */)
{
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t later = 0, multiHead = 0;
    if (j < n) {
        const uint64_t k = sortedKeys[j];
        const bool head = j == 0 || sortedKeys[j - 1] != k;
        const bool more = j + 1 < n && sortedKeys[j + 1] == k;
        flags[sortedIdx[j]] = (uint8_t)((head ? 0u : ISSL_FILTER_DUPLICATE) | ((!head || more) ? ISSL_FILTER_NOT_UNIQUE : 0u));
        later = head ? 0u : 1u;
        multiHead = (head && more) ? 1u : 0u;
    }
    const uint32_t nl = __popc(__ballot_sync(0xffffffffu, later)), nm = __popc(__ballot_sync(0xffffffffu, multiHead));
    if ((threadIdx.x & 31u) == 0) {
        if (nl) atomicAdd(counters, (unsigned long long)nl);
        if (nm) atomicAdd(counters + 1, (unsigned long long)nm);
    }
}

__global__ void k_gather_sites(const uint64_t *sig, uint64_t N, const uint64_t *siteIds, uint64_t n, uint64_t *out)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    out[t] = sig[siteIds[t] % N];
}

}  // namespace issl
