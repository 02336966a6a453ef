// scan_variants.cu -- where does the bucket scan's time go?  k_scan_triple_blocked took 4.34 ms per 100 000 guides in
// round 1 while the bare access pattern (tools/visit_order_bw.cu) runs in 2.6 ms.  This tool rebuilds the kernel from
// the library's own device functions (issl_triple.cuh) in steps -- load flavour, visits in flight per lane, shared
// memory footprint, what is done per hit inside the loop, what the tail gathers -- over synthetic blocks of the real
// size (10 x 2^24 x 128 B) and reports each step's time.  It measures instruction issue and the memory system; it does
// not check parity (the library's tests do).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false --expt-relaxed-constexpr -std=c++17 \
//             -o tools/_build/scan_variants tools/scan_variants.cu -Iinclude -Icrackling_b200/csrc \
//             -Lcrackling_b200/lib -lissl_cuda -Xlinker -rpath,'$ORIGIN/../../crackling_b200/lib'
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#include "issl_cuda.h"
#include "issl_triple.cuh"

using namespace issl;

__global__ void k_fill(uint4 *blk, uint64_t nSub)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nSub) return;
    uint32_t w[16];
    for (int p = 0; p < 16; p += 2) { const uint64_t r = mix64(i * 8 + p / 2); w[p] = (uint32_t)r & ~1u; w[p + 1] = (uint32_t)(r >> 32) & ~1u; }
    const uint32_t n = (i & 1) ? 4u : 31u;
    for (int p = 0; p < 5; p++) w[p] |= (n >> p) & 1u;
    if (mix64((i >> 1) ^ 0x5555ull) % 100 == 0) w[5] |= 1u;   // "more entries than the block holds" on 1 % of the blocks
    uint4 *o = blk + i * 4;
    o[0] = make_uint4(w[0], w[1], w[2], w[3]); o[1] = make_uint4(w[4], w[5], w[6], w[7]);
    o[2] = make_uint4(w[8], w[9], w[10], w[11]); o[3] = make_uint4(w[12], w[13], w[14], w[15]);
}
__global__ void k_fill_u32(uint32_t *p, uint64_t n, uint32_t mod)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)(mix64(i) % mod);
}

enum { LD_CS = 0, LD_NC = 1, LD_NC_NOALLOC = 2, LD_CA = 3 };
template <int LD> __device__ __forceinline__ uint4 load16(const uint4 *p)
{
    if constexpr (LD == LD_CS) return __ldcs(p);
    else if constexpr (LD == LD_NC) return __ldg(p);
    else if constexpr (LD == LD_CA) return *p;
    else {
        uint4 r;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
        return r;
    }
}

struct VArgs {
    const uint4 *blk;
    const uint64_t *guides;
    const uint2 *visits;
    uint32_t nVisits;
    const uint32_t *offs;   // [10][2^24 + 1]
    const uint32_t *ids;    // [10][stride]
    const uint16_t *res;    // [10][stride]
    uint64_t stride;
    unsigned long long *sum;
};

// HIT: 0 count only; 1 round 1's loop (keep test, residual gathered from the planes, 8-byte record);
//      2 one 16-byte record per sub-block with any slot within budget, everything else in the tail (sub-block re-read)
// TAIL: 0 none; 1 sites from the records, checksum; 2 + offs -> ids gathers per hit (round 1's fused tail)
// OVF: what a block's "more entries follow" flag costs: 0 ignored; 1 handled where it is met (two dependent loads in the
//      middle of the warp's round, round 1); 2 noted in shared memory and handled for the whole CTA after the loop
template <int HIT, int LD, int UNROLL, int TAIL, int OVF>
__global__ void __launch_bounds__(128, 10) k_variant(const VArgs a)
{
    __shared__ uint32_t nOvf;
    __shared__ uint32_t ovf[64];
    extern __shared__ uint4 pad[];
    __shared__ uint32_t key[10];
    __shared__ uint4 mask[10][4];
    __shared__ uint32_t nHits;
    __shared__ uint4 list[256];   // 4 KB: 512 8-byte records or 256 16-byte ones
    const uint64_t g = a.guides[blockIdx.x];
    if (threadIdx.x < 10) {
        const uint32_t t = threadIdx.x;
        key[t] = triple_key(g, c_tripleSlices[t][0], c_tripleSlices[t][1], c_tripleSlices[t][2]);
        const uint32_t r = triple_res(g, c_tripleSlices[t][3], c_tripleSlices[t][4]);
        uint32_t *m = reinterpret_cast<uint32_t *>(mask[t]);
        for (int p = 0; p < 16; p++) m[p] = 0u - ((r >> p) & 1u);
    }
    if (threadIdx.x == 0) { nHits = 0; nOvf = 0; }
    __syncthreads();
    const uint32_t sub = threadIdx.x & 1u, vslot = threadIdx.x >> 1;
    auto overflow = [&](uint32_t t, uint32_t k) {   // the rest of the bucket through its offsets (emulated: 8 more entries)
        const uint32_t *o = a.offs + (uint64_t)t * (kTripleBuckets + 1) + k;
        const uint32_t start = __ldg(o) + 62, end = __ldg(o + 1);
        const uint4 r = __ldg(reinterpret_cast<const uint4 *>(a.res + (uint64_t)t * a.stride) + (start >> 3) + sub);
        if (((r.x ^ r.y ^ r.z ^ r.w) & 0xFFFFFu) == (end & 0xFFFFFu)) atomicAdd(&nHits, 1u);
    };
    uint2 *list8 = reinterpret_cast<uint2 *>(list);

    auto process = [&](uint32_t e, uint2 v, uint32_t t, uint32_t k, const uint4 &q0, const uint4 &q1, const uint4 &q2, const uint4 &q3) {
        if constexpr (OVF == 1) { if (q1.y & 1u) overflow(t, k); }
        if constexpr (OVF == 2) { if ((q1.y & 1u) && sub == 0) { const uint32_t s = atomicAdd(&nOvf, 1u); if (s < 64) ovf[s] = e; } }
        const uint32_t cnt = (q0.x & 1u) | ((q0.y & 1u) << 1) | ((q0.z & 1u) << 2) | ((q0.w & 1u) << 3) | ((q1.x & 1u) << 4);
        if (cnt == 0) return;
        const uint4 m0 = mask[t][0], m1 = mask[t][1], m2 = mask[t][2], m3 = mask[t][3];
        const uint32_t x0 = (q0.x ^ m0.x) | (q0.y ^ m0.y), x1 = (q0.z ^ m0.z) | (q0.w ^ m0.w);
        const uint32_t x2 = (q1.x ^ m1.x) | (q1.y ^ m1.y), x3 = (q1.z ^ m1.z) | (q1.w ^ m1.w);
        const uint32_t x4 = (q2.x ^ m2.x) | (q2.y ^ m2.y), x5 = (q2.z ^ m2.z) | (q2.w ^ m2.w);
        const uint32_t x6 = (q3.x ^ m3.x) | (q3.y ^ m3.y), x7 = (q3.z ^ m3.z) | (q3.w ^ m3.w);
        uint32_t sa, ca, sb, cb, sc, cc, t1, u1;
        bs_full_add(x0, x1, x2, sa, ca); bs_full_add(x3, x4, x5, sb, cb); bs_full_add(x6, x7, sa, sc, cc);
        const uint32_t s0 = sb ^ sc, cd = sb & sc;
        bs_full_add(ca, cb, cc, t1, u1);
        const uint32_t s1 = t1 ^ cd, u2 = t1 & cd, s2 = u1 ^ u2, s3 = u1 & u2;
        uint32_t over;
        switch (v.x >> 28) {
        case 0: over = s0 | s1 | s2 | s3; break;
        case 1: over = s1 | s2 | s3; break;
        case 2: over = s2 | s3 | (s1 & s0); break;
        case 3: over = s2 | s3; break;
        case 4: over = s3 | (s2 & (s1 | s0)); break;
        case 5: over = s3 | (s2 & s1); break;
        case 6: over = s3 | (s2 & s1 & s0); break;
        default: over = s3; break;
        }
        uint32_t pass = ~over & ((2u << cnt) - 2u);
        if (!pass) return;
        if constexpr (HIT == 0) {
            atomicAdd(&nHits, (uint32_t)__popc(pass));
        } else if constexpr (HIT == 2) {
            const uint32_t pEx = ~(x0 | x1 | x2 | x3), qEx = ~(x4 | x5 | x6 | x7);
            const uint32_t slot = atomicAdd(&nHits, 1u);
            if (slot < 256) list[slot] = make_uint4(e * 2 + sub, pass, pEx & pass, qEx & pass);
        } else {
            const uint32_t pEx = ~(x0 | x1 | x2 | x3), qEx = ~(x4 | x5 | x6 | x7);
            do {
                const uint32_t sl = __ffs(pass) - 1;
                pass &= pass - 1;
                const uint32_t y = record_y(v, (pEx >> sl) & 1u, (qEx >> sl) & 1u, kRecBlocked);
                uint32_t minE;
                if (!record_keep(make_uint2(0u, y), minE)) continue;
                const uint32_t r =
                    ((q0.x >> sl) & 1u) | (((q0.y >> sl) & 1u) << 1) | (((q0.z >> sl) & 1u) << 2) | (((q0.w >> sl) & 1u) << 3) |
                    (((q1.x >> sl) & 1u) << 4) | (((q1.y >> sl) & 1u) << 5) | (((q1.z >> sl) & 1u) << 6) | (((q1.w >> sl) & 1u) << 7) |
                    (((q2.x >> sl) & 1u) << 8) | (((q2.y >> sl) & 1u) << 9) | (((q2.z >> sl) & 1u) << 10) | (((q2.w >> sl) & 1u) << 11) |
                    (((q3.x >> sl) & 1u) << 12) | (((q3.y >> sl) & 1u) << 13) | (((q3.z >> sl) & 1u) << 14) | (((q3.w >> sl) & 1u) << 15);
                const uint32_t slot = atomicAdd(&nHits, 1u);
                if (slot < 512) list8[slot] = make_uint2(k | ((sub * kSubEntries + sl) << 24), y | (r << 16));
            } while (pass);
        }
    };

    for (uint32_t e0 = vslot; e0 < a.nVisits; e0 += 64 * UNROLL) {
        uint2 v[UNROLL]; uint32_t t[UNROLL], k[UNROLL]; uint4 q[UNROLL][4];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const uint32_t e = e0 + u * 64;
            if (e < a.nVisits) {
                v[u] = __ldg(a.visits + e);
                t[u] = (v[u].x >> 24) & 15u; k[u] = key[t[u]] ^ (v[u].x & 0xFFFFFFu);
                const uint4 *p = a.blk + ((((uint64_t)t[u] << 24) | k[u]) * 2 + sub) * 4;
                q[u][0] = load16<LD>(p); q[u][1] = load16<LD>(p + 1); q[u][2] = load16<LD>(p + 2); q[u][3] = load16<LD>(p + 3);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
            if (e0 + u * 64 < a.nVisits) process(e0 + u * 64, v[u], t[u], k[u], q[u][0], q[u][1], q[u][2], q[u][3]);
    }
    __syncthreads();
    if constexpr (OVF == 2) {
        for (uint32_t j = vslot; j < min(nOvf, 64u); j += 64) {
            const uint2 v = __ldg(a.visits + ovf[j]);
            const uint32_t t = (v.x >> 24) & 15u;
            overflow(t, key[t] ^ (v.x & 0xFFFFFFu));
        }
        __syncthreads();
    }
    unsigned long long acc = 0;
    const TripleView tv{nullptr, a.ids, 1u, a.offs, a.stride, a.blk, 64u, 0u};
    if constexpr (HIT == 1 && TAIL >= 1) {
        const uint32_t n = min(nHits, 512u);
        for (uint32_t j = threadIdx.x; j < n; j += 128) {
            const uint2 h = list8[j];
            acc += sig_to_sortkey(hit_site(tv, h), 20);
            if constexpr (TAIL == 2) acc += __ldg(a.ids + (uint64_t)(h.y & 15u) * a.stride + hit_position(tv, h));
        }
    }
    if constexpr (HIT == 2 && TAIL >= 1) {
        const uint32_t n = min(nHits, 256u);
        for (uint32_t j = threadIdx.x; j < n; j += 128) {
            const uint4 rec = list[j];
            const uint2 v = __ldg(a.visits + (rec.x >> 1));
            const uint32_t t = (v.x >> 24) & 15u, k = key[t] ^ (v.x & 0xFFFFFFu), sb = rec.x & 1u;
            const uint4 *p = a.blk + ((((uint64_t)t << 24) | k) * 2 + sb) * 4;
            const uint4 q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + 2), q3 = __ldg(p + 3);
            uint32_t pass = rec.y;
            do {
                const uint32_t sl = __ffs(pass) - 1;
                pass &= pass - 1;
                const uint32_t y = record_y(v, (rec.z >> sl) & 1u, (rec.w >> sl) & 1u, kRecBlocked);
                uint32_t minE;
                if (!record_keep(make_uint2(0u, y), minE)) continue;
                const uint32_t r =
                    ((q0.x >> sl) & 1u) | (((q0.y >> sl) & 1u) << 1) | (((q0.z >> sl) & 1u) << 2) | (((q0.w >> sl) & 1u) << 3) |
                    (((q1.x >> sl) & 1u) << 4) | (((q1.y >> sl) & 1u) << 5) | (((q1.z >> sl) & 1u) << 6) | (((q1.w >> sl) & 1u) << 7) |
                    (((q2.x >> sl) & 1u) << 8) | (((q2.y >> sl) & 1u) << 9) | (((q2.z >> sl) & 1u) << 10) | (((q2.w >> sl) & 1u) << 11) |
                    (((q3.x >> sl) & 1u) << 12) | (((q3.y >> sl) & 1u) << 13) | (((q3.z >> sl) & 1u) << 14) | (((q3.w >> sl) & 1u) << 15);
                acc += sig_to_sortkey(hit_site(tv, make_uint2(k | ((sb * kSubEntries + sl) << 24), y | (r << 16))), 20);
            } while (pass);
        }
    }
    if (threadIdx.x == 0) acc += nHits;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31u) == 0 && acc) atomicAdd(a.sum, acc);
}

// ------------------------------------------------------------------------------------------------
// Asynchronous copies into a shared-memory ring instead of loads into registers (round 2: "a cp.async.bulk 128-B ring
// with an mbarrier").  Both kernels do the bit-sliced compare of k_variant<HIT = 0> (hits are only counted), so their
// times compare with its first line.
//   k_ring_bulk    cp.async.bulk (the TMA unit's 1-D copy, SASS UBLKCP): every lane copies its visit's 128-byte block into
//                  its slot of a per-warp stage (32 x 128 B), completion through one mbarrier per (warp, stage)
//   k_ring_cpasync cp.async.cg 16 B (SASS LDGSTS.BYPASS): eight per lane and visit, chunk-major inside the warp's stage
//                  (conflict-free writes and reads), completion through cp.async.wait_group
// DEPTH stages per warp are in flight.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t compare_count(const uint4 *mask, uint32_t budget, const uint4 q0, const uint4 q1, const uint4 q2, const uint4 q3)
{
    const uint32_t cnt = (q0.x & 1u) | ((q0.y & 1u) << 1) | ((q0.z & 1u) << 2) | ((q0.w & 1u) << 3) | ((q1.x & 1u) << 4);
    if (cnt == 0) return 0;
    const uint4 m0 = mask[0], m1 = mask[1], m2 = mask[2], m3 = mask[3];
    const uint32_t x0 = (q0.x ^ m0.x) | (q0.y ^ m0.y), x1 = (q0.z ^ m0.z) | (q0.w ^ m0.w);
    const uint32_t x2 = (q1.x ^ m1.x) | (q1.y ^ m1.y), x3 = (q1.z ^ m1.z) | (q1.w ^ m1.w);
    const uint32_t x4 = (q2.x ^ m2.x) | (q2.y ^ m2.y), x5 = (q2.z ^ m2.z) | (q2.w ^ m2.w);
    const uint32_t x6 = (q3.x ^ m3.x) | (q3.y ^ m3.y), x7 = (q3.z ^ m3.z) | (q3.w ^ m3.w);
    uint32_t sa, ca, sb, cb, sc, cc, t1, u1;
    bs_full_add(x0, x1, x2, sa, ca); bs_full_add(x3, x4, x5, sb, cb); bs_full_add(x6, x7, sa, sc, cc);
    const uint32_t s0 = sb ^ sc, cd = sb & sc;
    bs_full_add(ca, cb, cc, t1, u1);
    const uint32_t s1 = t1 ^ cd, u2 = t1 & cd, s2 = u1 ^ u2, s3 = u1 & u2;
    const uint32_t b0 = 0u - (budget & 1u), b1 = 0u - ((budget >> 1) & 1u), b2 = 0u - ((budget >> 2) & 1u);
    uint32_t over = s0 & ~b0;
    over = (s1 & ~b1) | (~(s1 ^ b1) & over);
    over = (s2 & ~b2) | (~(s2 ^ b2) & over);
    over |= s3;
    return (uint32_t)__popc(~over & ((2u << cnt) - 2u));
}

template <int DEPTH>
__global__ void __launch_bounds__(128) k_ring_bulk(const VArgs a)
{
    extern __shared__ __align__(128) uint8_t ring[];   // [4 warps][DEPTH][32 lanes][128 B]
    __shared__ __align__(8) uint64_t bar[4][DEPTH];
    __shared__ uint32_t key[10];
    __shared__ uint4 mask[10][4];
    __shared__ uint32_t nHits;
    const uint64_t g = a.guides[blockIdx.x];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if (threadIdx.x < 10) {
        const uint32_t t = threadIdx.x;
        key[t] = triple_key(g, c_tripleSlices[t][0], c_tripleSlices[t][1], c_tripleSlices[t][2]);
        const uint32_t r = triple_res(g, c_tripleSlices[t][3], c_tripleSlices[t][4]);
        uint32_t *m = reinterpret_cast<uint32_t *>(mask[t]);
        for (int p = 0; p < 16; p++) m[p] = 0u - ((r >> p) & 1u);
    }
    if (threadIdx.x == 0) nHits = 0;
    if (lane == 0)
        for (int s = 0; s < DEPTH; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" ::"r"(smem_u32(&bar[warp][s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint8_t *mine = ring + (size_t)warp * DEPTH * 4096;
    auto issue = [&](uint32_t round, uint32_t stage) {   // this lane's visit of the warp's round
        const uint32_t e = round * 128 + threadIdx.x;
        const uint32_t mb = smem_u32(&bar[warp][stage]);
        if (e < a.nVisits) {
            const uint2 v = __ldg(a.visits + e);
            const uint32_t t = (v.x >> 24) & 15u, k = key[t] ^ (v.x & 0xFFFFFFu);
            const uint4 *p = a.blk + (((uint64_t)t << 24) | k) * 8;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 128;" ::"r"(mb) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];"
                         ::"r"(smem_u32(mine + stage * 4096 + lane * 128)), "l"(p), "r"(mb) : "memory");
        } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mb) : "memory");
        }
    };
    const uint32_t rounds = (a.nVisits + 127) / 128;
    for (uint32_t r = 0; r < (uint32_t)DEPTH && r < rounds; r++) issue(r, r);
    uint32_t hits = 0;
    for (uint32_t r = 0; r < rounds; r++) {
        const uint32_t stage = r % DEPTH, parity = (r / DEPTH) & 1u;
        const uint32_t mb = smem_u32(&bar[warp][stage]);
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(mb), "r"(parity) : "memory");
        const uint32_t e = r * 128 + threadIdx.x;
        if (e < a.nVisits) {
            const uint2 v = __ldg(a.visits + e);
            const uint32_t t = (v.x >> 24) & 15u;
            const uint4 *q = reinterpret_cast<const uint4 *>(mine + stage * 4096 + lane * 128);
            hits += compare_count(mask[t], v.x >> 28, q[0], q[1], q[2], q[3]);
            hits += compare_count(mask[t], v.x >> 28, q[4], q[5], q[6], q[7]);
        }
        __syncwarp();
        if (r + DEPTH < rounds) issue(r + DEPTH, stage);
    }
    if (hits) atomicAdd(&nHits, hits);
    __syncthreads();
    if (threadIdx.x == 0 && nHits) atomicAdd(a.sum, (unsigned long long)nHits);
}

template <int DEPTH>
__global__ void __launch_bounds__(128) k_ring_cpasync(const VArgs a)
{
    extern __shared__ __align__(128) uint8_t ring[];   // [4 warps][DEPTH][8 chunks][32 lanes][16 B]
    __shared__ uint32_t key[10];
    __shared__ uint4 mask[10][4];
    __shared__ uint32_t nHits;
    const uint64_t g = a.guides[blockIdx.x];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if (threadIdx.x < 10) {
        const uint32_t t = threadIdx.x;
        key[t] = triple_key(g, c_tripleSlices[t][0], c_tripleSlices[t][1], c_tripleSlices[t][2]);
        const uint32_t r = triple_res(g, c_tripleSlices[t][3], c_tripleSlices[t][4]);
        uint32_t *m = reinterpret_cast<uint32_t *>(mask[t]);
        for (int p = 0; p < 16; p++) m[p] = 0u - ((r >> p) & 1u);
    }
    if (threadIdx.x == 0) nHits = 0;
    __syncthreads();
    uint8_t *mine = ring + (size_t)warp * DEPTH * 4096;
    auto issue = [&](uint32_t round, uint32_t stage) {
        const uint32_t e = round * 128 + threadIdx.x;
        if (e < a.nVisits) {
            const uint2 v = __ldg(a.visits + e);
            const uint32_t t = (v.x >> 24) & 15u, k = key[t] ^ (v.x & 0xFFFFFFu);
            const uint4 *p = a.blk + (((uint64_t)t << 24) | k) * 8;
            const uint32_t dst = smem_u32(mine + stage * 4096 + lane * 16);
#pragma unroll
            for (int c = 0; c < 8; c++)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + c * 512), "l"(p + c) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const uint32_t rounds = (a.nVisits + 127) / 128;
    for (uint32_t r = 0; r < (uint32_t)DEPTH; r++) issue(r, r);   // empty groups beyond the last round keep the count uniform
    uint32_t hits = 0;
    for (uint32_t r = 0; r < rounds; r++) {
        const uint32_t stage = r % DEPTH;
        asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
        const uint32_t e = r * 128 + threadIdx.x;
        if (e < a.nVisits) {
            const uint2 v = __ldg(a.visits + e);
            const uint32_t t = (v.x >> 24) & 15u;
            const uint4 *q = reinterpret_cast<const uint4 *>(mine + stage * 4096 + lane * 16);
            hits += compare_count(mask[t], v.x >> 28, q[0], q[32], q[64], q[96]);
            hits += compare_count(mask[t], v.x >> 28, q[128], q[160], q[192], q[224]);
        }
        issue(r + DEPTH, stage);
    }
    if (hits) atomicAdd(&nHits, hits);
    __syncthreads();
    if (threadIdx.x == 0 && nHits) atomicAdd(a.sum, (unsigned long long)nHits);
}

int main(int argc, char **argv)
{
    const uint32_t n = argc > 1 ? (uint32_t)atol(argv[1]) : 100000u;
    const uint64_t nBlocks = 10ull << 24, N = 581045288ull, stride = (N + 64 + 7) / 8 * 8;
    uint4 *blk; uint32_t *offs, *ids; uint16_t *res;
    if (cudaMalloc(&blk, nBlocks * 128) != cudaSuccess || cudaMalloc(&offs, 10 * ((1ull << 24) + 1) * 4) != cudaSuccess ||
        cudaMalloc(&ids, 10 * stride * 4) != cudaSuccess || cudaMalloc(&res, 10 * stride * 2) != cudaSuccess) { printf("{\"error\": \"cudaMalloc\"}\n"); return 1; }
    k_fill<<<(unsigned)((nBlocks * 2 + 255) / 256), 256>>>(blk, nBlocks * 2);
    k_fill_u32<<<(unsigned)((10 * ((1ull << 24) + 1) + 255) / 256), 256>>>(offs, 10 * ((1ull << 24) + 1), (uint32_t)(N - 64));
    cudaMemset(ids, 0, 10 * stride * 4); cudaMemset(res, 0, 10 * stride * 2);
    uint32_t ws[6];
    std::vector<uint32_t> raw(issl_triple_visits(4, nullptr, 0, nullptr));
    issl_triple_visits(4, raw.data(), raw.size(), ws);
    static const uint8_t slices[10][5] = ISSL_TRIPLE_LAYOUT_INIT;
    std::vector<uint2> vis(raw.size());
    for (size_t i = 0; i < raw.size(); i++) {
        const uint32_t t = (raw[i] >> 24) & 15u;
        uint32_t exact = 0;
        for (int k = 0; k < 3; k++) if (((raw[i] >> (8 * k)) & 0xFFu) == 0) exact |= 1u << slices[t][k];
        vis[i] = make_uint2(raw[i], exact | ((uint32_t)slices[t][3] << 8) | ((uint32_t)slices[t][4] << 12));
    }
    uint2 *dvis; cudaMalloc(&dvis, vis.size() * 8);
    cudaMemcpy(dvis, vis.data(), vis.size() * 8, cudaMemcpyHostToDevice);
    std::vector<uint64_t> g(n);
    for (uint32_t i = 0; i < n; i++) g[i] = mix64(0xABCDEFull + i) & ((1ull << 40) - 1);
    uint64_t *dg; cudaMalloc(&dg, n * 8ull);
    cudaMemcpy(dg, g.data(), n * 8ull, cudaMemcpyHostToDevice);
    unsigned long long *dsum; cudaMalloc(&dsum, 8);
    VArgs a{blk, dg, dvis, (uint32_t)vis.size(), offs, ids, res, stride, dsum};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = (double)n * vis.size() * 128;

#define RUN(HIT, LD, UNROLL, TAIL, PADKB, OVF)                                                                                      \
    do {                                                                                                                       \
        float ms = 0, best = 1e9f;                                                                                             \
        int resident = 0;                                                                                                      \
        cudaFuncSetAttribute(k_variant<HIT, LD, UNROLL, TAIL, OVF>, cudaFuncAttributeMaxDynamicSharedMemorySize, PADKB * 1024);     \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_variant<HIT, LD, UNROLL, TAIL, OVF>, 128, PADKB * 1024);         \
        for (int rep = 0; rep < 4; rep++) {                                                                                    \
            cudaMemset(dsum, 0, 8);                                                                                            \
            cudaEventRecord(e0);                                                                                               \
            k_variant<HIT, LD, UNROLL, TAIL, OVF><<<n, 128, PADKB * 1024>>>(a);                                                     \
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);                                  \
            if (rep && ms < best) best = ms;                                                                                   \
        }                                                                                                                      \
        printf("{\"hit\": %d, \"ld\": \"%s\", \"unroll\": %d, \"tail\": %d, \"pad_kb\": %d, \"ovf\": %d, \"ctas_per_sm\": %d, \"ms_per_100k\": %.3f, " \
               "\"GB/s_blocks\": %.1f, \"error\": \"%s\"}\n", HIT, #LD, UNROLL, TAIL, PADKB, OVF, resident, best * 1e5 / n,         \
               bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));                                                    \
        fflush(stdout);                                                                                                        \
    } while (0)

    RUN(0, LD_CS, 1, 0, 0, 0); RUN(0, LD_CS, 1, 0, 0, 1); RUN(0, LD_CS, 1, 0, 0, 2);
    RUN(1, LD_CS, 1, 0, 0, 0); RUN(1, LD_CS, 1, 1, 0, 0); RUN(1, LD_CS, 1, 2, 0, 0);
    RUN(1, LD_CS, 1, 2, 3, 0); RUN(1, LD_CS, 1, 2, 3, 1); RUN(1, LD_CS, 1, 2, 3, 2); RUN(1, LD_CS, 1, 1, 3, 2);
    RUN(2, LD_CS, 1, 0, 0, 0); RUN(2, LD_CS, 1, 1, 0, 0); RUN(2, LD_CS, 1, 1, 3, 0); RUN(2, LD_CS, 1, 1, 3, 2);
    RUN(2, LD_NC, 1, 1, 3, 2); RUN(2, LD_CA, 1, 1, 3, 2);
    RUN(0, LD_CS, 2, 0, 0, 0); RUN(1, LD_CS, 2, 1, 0, 2);

#define RUN_RING(KERNEL, NAME, DEPTH)                                                                                          \
    do {                                                                                                                       \
        float ms = 0, best = 1e9f;                                                                                             \
        int resident = 0;                                                                                                      \
        const int smem = 4 * DEPTH * 4096;                                                                                     \
        cudaFuncSetAttribute(KERNEL<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                                \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, KERNEL<DEPTH>, 128, smem);                                    \
        for (int rep = 0; rep < 4; rep++) {                                                                                    \
            cudaMemset(dsum, 0, 8);                                                                                            \
            cudaEventRecord(e0);                                                                                               \
            KERNEL<DEPTH><<<n, 128, smem>>>(a);                                                                                \
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);                                  \
            if (rep && ms < best) best = ms;                                                                                   \
        }                                                                                                                      \
        unsigned long long hsum = 0; cudaMemcpy(&hsum, dsum, 8, cudaMemcpyDeviceToHost);                                       \
        printf("{\"ring\": \"%s\", \"depth\": %d, \"smem_kb_per_cta\": %d, \"ctas_per_sm\": %d, \"ms_per_100k\": %.3f, \"GB/s_blocks\": %.1f, " \
               "\"hits\": %llu, \"error\": \"%s\"}\n", NAME, DEPTH, smem / 1024, resident, best * 1e5 / n, bytes / best / 1e6, hsum, \
               cudaGetErrorString(cudaGetLastError()));                                                                        \
        fflush(stdout);                                                                                                        \
    } while (0)
    {
        cudaMemset(dsum, 0, 8);
        k_variant<0, LD_CS, 1, 0, 0><<<n, 128>>>(a);
        unsigned long long hsum = 0; cudaMemcpy(&hsum, dsum, 8, cudaMemcpyDeviceToHost);
        printf("{\"ring\": \"none (registers)\", \"hits\": %llu}\n", hsum);
    }
    RUN_RING(k_ring_bulk, "cp.async.bulk 128 B per lane", 1); RUN_RING(k_ring_bulk, "cp.async.bulk 128 B per lane", 2);
    RUN_RING(k_ring_bulk, "cp.async.bulk 128 B per lane", 3);
    RUN_RING(k_ring_cpasync, "cp.async.cg 16 B x 8 per lane", 1); RUN_RING(k_ring_cpasync, "cp.async.cg 16 B x 8 per lane", 2);
    RUN_RING(k_ring_cpasync, "cp.async.cg 16 B x 8 per lane", 3);
    return 0;
}
