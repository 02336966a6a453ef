// visit_order_bw.cu -- does the ORDER in which (guide, bucket) visits are issued matter to the bucket scan?
//
// The blocked copy of ISSL_LAYOUT_TRIPLE is 10 x 2^24 blocks of 128 B (21.5 GB at human scale).  A guide reads 1 390
// of them (maxDist 4).  Order A ("guide-major", round 1's k_scan_triple_blocked): one CTA per guide walks all ten
// triples, so concurrently running CTAs touch unrelated blocks.  Order B ("triple-sorted"): per triple, guides are
// bucketed by the two high key bytes and a CTA takes a run of G neighbouring guides of ONE triple, so that CTAs
// running at the same time read the same 32 KB / 8 MB windows -- repeated blocks come from L2 and DRAM pages stay open.
// Same loads (4 x 16 B per lane, two lanes per block), same bit-sliced compare, hits appended to a shared-memory list.
// Synthetic block contents (random planes, 31 + 4 entries per block); this measures the memory system, not parity.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/visit_order_bw tools/visit_order_bw.cu \
//             -Iinclude -Lcrackling_b200/lib -lissl_cuda -Xlinker -rpath,'$ORIGIN/../../crackling_b200/lib'
// Run:   tools/_build/visit_order_bw [guides ...]   -> JSON lines
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#include "issl_cuda.h"

static const uint8_t kLay[10][5] = {{2, 1, 0, 3, 4}, {1, 0, 3, 2, 4}, {1, 0, 4, 2, 3}, {3, 0, 2, 1, 4}, {0, 2, 4, 1, 3},
                                    {0, 3, 4, 1, 2}, {3, 2, 1, 0, 4}, {2, 1, 4, 0, 3}, {4, 1, 3, 0, 2}, {4, 2, 3, 0, 1}};
__constant__ uint8_t c_lay[10][5];

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__host__ __device__ __forceinline__ uint32_t key_of(uint64_t g, const uint8_t *l)
{
    return (uint32_t)((g >> (8 * l[0])) & 0xFF) | ((uint32_t)((g >> (8 * l[1])) & 0xFF) << 8) | ((uint32_t)((g >> (8 * l[2])) & 0xFF) << 16);
}
__host__ __device__ __forceinline__ uint32_t res_of(uint64_t g, const uint8_t *l)
{
    return (uint32_t)((g >> (8 * l[3])) & 0xFF) | ((uint32_t)((g >> (8 * l[4])) & 0xFF) << 8);
}

__global__ void k_fill(uint4 *blk, uint64_t nSub)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nSub) return;
    uint32_t w[16];
    for (int p = 0; p < 16; p += 2) { const uint64_t r = mix64(i * 8 + p / 2); w[p] = (uint32_t)r & ~1u; w[p + 1] = (uint32_t)(r >> 32) & ~1u; }
    const uint32_t n = (i & 1) ? 4u : 31u;
    for (int p = 0; p < 5; p++) w[p] |= (n >> p) & 1u;
    uint4 *o = blk + i * 4;
    o[0] = make_uint4(w[0], w[1], w[2], w[3]); o[1] = make_uint4(w[4], w[5], w[6], w[7]);
    o[2] = make_uint4(w[8], w[9], w[10], w[11]); o[3] = make_uint4(w[12], w[13], w[14], w[15]);
}

__device__ __forceinline__ void full_add(uint32_t a, uint32_t b, uint32_t c, uint32_t &s, uint32_t &k) { s = a ^ b ^ c; k = (a & b) | (c & (a ^ b)); }

// the compare of k_scan_triple_blocked: 31 residuals against the guide's, slots within budget
__device__ __forceinline__ uint32_t compare(const uint4 &q0, const uint4 &q1, const uint4 &q2, const uint4 &q3, const uint4 *m, uint32_t budget)
{
    const uint32_t cnt = (q0.x & 1u) | ((q0.y & 1u) << 1) | ((q0.z & 1u) << 2) | ((q0.w & 1u) << 3) | ((q1.x & 1u) << 4);
    const uint4 m0 = m[0], m1 = m[1], m2 = m[2], m3 = m[3];
    const uint32_t x0 = (q0.x ^ m0.x) | (q0.y ^ m0.y), x1 = (q0.z ^ m0.z) | (q0.w ^ m0.w);
    const uint32_t x2 = (q1.x ^ m1.x) | (q1.y ^ m1.y), x3 = (q1.z ^ m1.z) | (q1.w ^ m1.w);
    const uint32_t x4 = (q2.x ^ m2.x) | (q2.y ^ m2.y), x5 = (q2.z ^ m2.z) | (q2.w ^ m2.w);
    const uint32_t x6 = (q3.x ^ m3.x) | (q3.y ^ m3.y), x7 = (q3.z ^ m3.z) | (q3.w ^ m3.w);
    uint32_t sa, ca, sb, cb, sc, cc, t1, u1;
    full_add(x0, x1, x2, sa, ca); full_add(x3, x4, x5, sb, cb); full_add(x6, x7, sa, sc, cc);
    const uint32_t s0 = sb ^ sc, cd = sb & sc;
    full_add(ca, cb, cc, t1, u1);
    const uint32_t s1 = t1 ^ cd, u2 = t1 & cd, s2 = u1 ^ u2, s3 = u1 & u2;
    uint32_t over;
    switch (budget) {
    case 0: over = s0 | s1 | s2 | s3; break;
    case 1: over = s1 | s2 | s3; break;
    case 2: over = s2 | s3 | (s1 & s0); break;
    case 3: over = s2 | s3; break;
    default: over = s3 | (s2 & (s1 | s0)); break;
    }
    return ~over & ((2u << cnt) - 2u);
}

struct Args {
    const uint4 *blk;
    const uint64_t *guides;
    const uint2 *visits;       // x: pattern | t << 24 | budget << 28
    uint32_t nVisits;
    const uint32_t *order;     // B: [10][n] guide indices, bucketed per triple
    uint32_t n;
    uint32_t tripleFirst[11];  // B: first visit of every triple in the per-triple table
    uint32_t G;                // B: guides per CTA (power of two)
    uint32_t ctaFirst[11];     // B: first CTA of every triple
    unsigned long long *hits;
};

// order A: CTA = guide
__global__ void __launch_bounds__(128, 10) k_guide_major(const Args a)
{
    __shared__ uint32_t key[10];
    __shared__ uint4 mask[10][4];
    __shared__ uint32_t nHits;
    __shared__ uint2 list[512];
    const uint64_t g = a.guides[blockIdx.x];
    if (threadIdx.x < 10) {
        const uint32_t t = threadIdx.x;
        key[t] = key_of(g, c_lay[t]);
        const uint32_t r = res_of(g, c_lay[t]);
        uint32_t *m = reinterpret_cast<uint32_t *>(mask[t]);
        for (int p = 0; p < 16; p++) m[p] = 0u - ((r >> p) & 1u);
    }
    if (threadIdx.x == 0) nHits = 0;
    __syncthreads();
    const uint32_t sub = threadIdx.x & 1u, vslot = threadIdx.x >> 1;
    for (uint32_t e = vslot; e < a.nVisits; e += 64) {
        const uint2 v = __ldg(a.visits + e);
        const uint32_t t = (v.x >> 24) & 15u, k = key[t] ^ (v.x & 0xFFFFFFu);
        const uint4 *p = a.blk + ((((uint64_t)t << 24) | k) * 2 + sub) * 4;
        const uint4 q0 = __ldcs(p), q1 = __ldcs(p + 1), q2 = __ldcs(p + 2), q3 = __ldcs(p + 3);
        uint32_t pass = compare(q0, q1, q2, q3, mask[t], v.x >> 28);
        while (pass) {
            const uint32_t sl = __ffs(pass) - 1; pass &= pass - 1;
            const uint32_t s = atomicAdd(&nHits, 1u);
            if (s < 512) list[s] = make_uint2(k | (sl << 24), t);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && nHits) atomicAdd(a.hits, (unsigned long long)nHits);
}

// order B: CTA = (triple, run of G neighbouring guides); lanes walk visit-major so that a warp reads the same pattern
// for 16 neighbouring guides.  LDG: default caching (the blocks are meant to be found in L2 by the neighbours).
template <bool STREAM>
__global__ void __launch_bounds__(128, 10) k_triple_sorted(const Args a)
{
    extern __shared__ uint4 smem[];
    __shared__ uint32_t nHits;
    __shared__ uint2 list[512];
    uint32_t t = 0;
    while (blockIdx.x >= a.ctaFirst[t + 1]) t++;
    const uint32_t run0 = (blockIdx.x - a.ctaFirst[t]) * a.G, G = min(a.G, a.n - run0);
    uint4 *mask = smem;                                       // [G][5]: 4 mask vectors + (key, -, -, -)
    for (uint32_t j = threadIdx.x; j < G; j += blockDim.x) {
        const uint64_t g = a.guides[a.order[(uint64_t)t * a.n + run0 + j]];
        const uint32_t r = res_of(g, c_lay[t]);
        uint32_t *m = reinterpret_cast<uint32_t *>(mask + j * 5);
        for (int p = 0; p < 16; p++) m[p] = 0u - ((r >> p) & 1u);
        m[16] = key_of(g, c_lay[t]);
    }
    if (threadIdx.x == 0) nHits = 0;
    __syncthreads();
    const uint32_t sub = threadIdx.x & 1u, vslot = threadIdx.x >> 1;
    const uint32_t v0 = a.tripleFirst[t], nv = a.tripleFirst[t + 1] - v0, items = nv * a.G;
    const uint32_t gshift = 31 - __clz(a.G);
    for (uint32_t i = vslot; i < items; i += 64) {
        const uint32_t j = i & (a.G - 1), vi = i >> gshift;
        if (j >= G) continue;
        const uint2 v = __ldg(a.visits + v0 + vi);
        const uint32_t k = reinterpret_cast<const uint32_t *>(mask + j * 5)[16] ^ (v.x & 0xFFFFFFu);
        const uint4 *p = a.blk + ((((uint64_t)t << 24) | k) * 2 + sub) * 4;
        uint4 q0, q1, q2, q3;
        if (STREAM) { q0 = __ldcs(p); q1 = __ldcs(p + 1); q2 = __ldcs(p + 2); q3 = __ldcs(p + 3); }
        else { q0 = __ldg(p); q1 = __ldg(p + 1); q2 = __ldg(p + 2); q3 = __ldg(p + 3); }
        uint32_t pass = compare(q0, q1, q2, q3, mask + j * 5, v.x >> 28);
        while (pass) {
            const uint32_t sl = __ffs(pass) - 1; pass &= pass - 1;
            const uint32_t s = atomicAdd(&nHits, 1u);
            if (s < 512) list[s] = make_uint2(k | (sl << 24), t);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && nHits) atomicAdd(a.hits, (unsigned long long)nHits);
}

int main(int argc, char **argv)
{
    std::vector<uint32_t> sizes;
    for (int i = 1; i < argc; i++) sizes.push_back((uint32_t)atol(argv[i]));
    if (sizes.empty()) sizes = {100000u, 1000000u};
    const uint64_t nBlocks = 10ull << 24;
    uint4 *blk;
    if (cudaMalloc(&blk, nBlocks * 128) != cudaSuccess) { printf("{\"error\": \"cudaMalloc\"}\n"); return 1; }
    cudaMemcpyToSymbol(c_lay, kLay, sizeof kLay);
    k_fill<<<(unsigned)((nBlocks * 2 + 255) / 256), 256>>>(blk, nBlocks * 2);
    // visit table, guide-major order (as the library) and grouped per triple
    uint32_t ws[6];
    std::vector<uint32_t> raw(issl_triple_visits(4, nullptr, 0, nullptr));
    issl_triple_visits(4, raw.data(), raw.size(), ws);
    std::vector<uint2> va(raw.size()), vb;
    for (size_t i = 0; i < raw.size(); i++) va[i] = make_uint2(raw[i], 0);
    Args a{};
    for (uint32_t t = 0; t < 10; t++) {
        a.tripleFirst[t] = (uint32_t)vb.size();
        for (uint32_t x : raw) if (((x >> 24) & 15u) == t) vb.push_back(make_uint2(x, 0));
    }
    a.tripleFirst[10] = (uint32_t)vb.size();
    uint2 *dva, *dvb;
    cudaMalloc(&dva, va.size() * 8); cudaMalloc(&dvb, vb.size() * 8);
    cudaMemcpy(dva, va.data(), va.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dvb, vb.data(), vb.size() * 8, cudaMemcpyHostToDevice);
    unsigned long long *dHits;
    cudaMalloc(&dHits, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFuncSetAttribute(k_triple_sorted<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 80);
    cudaFuncSetAttribute(k_triple_sorted<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 80);

    for (uint32_t n : sizes) {
        std::vector<uint64_t> g(n);
        for (uint32_t i = 0; i < n; i++) g[i] = mix64(0xABCDEFull + i) & ((1ull << 40) - 1);
        // per triple: guides ordered by the two high key bytes (counting sort on the host; the product would do it on the device)
        std::vector<uint32_t> order((size_t)10 * n);
        for (uint32_t t = 0; t < 10; t++) {
            std::vector<std::pair<uint32_t, uint32_t>> kv(n);
            for (uint32_t i = 0; i < n; i++) kv[i] = {key_of(g[i], kLay[t]) >> 8, i};
            std::sort(kv.begin(), kv.end());
            for (uint32_t i = 0; i < n; i++) order[(size_t)t * n + i] = kv[i].second;
        }
        uint64_t *dg; uint32_t *dorder;
        cudaMalloc(&dg, n * 8ull); cudaMalloc(&dorder, order.size() * 4);
        cudaMemcpy(dg, g.data(), n * 8ull, cudaMemcpyHostToDevice);
        cudaMemcpy(dorder, order.data(), order.size() * 4, cudaMemcpyHostToDevice);
        a.blk = blk; a.guides = dg; a.order = dorder; a.n = n; a.hits = dHits;
        const double bytes = (double)n * raw.size() * 128;
        auto report = [&](const char *name, uint32_t G, float ms, unsigned long long hits) {
            printf("{\"order\": \"%s\", \"guides\": %u, \"guides_per_cta\": %u, \"ms\": %.3f, \"ms_per_100k\": %.3f, \"GB/s_requested\": %.1f, \"hits_per_guide\": %.1f, \"error\": \"%s\"}\n",
                   name, n, G, ms, ms * 1e5 / n, bytes / ms / 1e6, (double)hits / n, cudaGetErrorString(cudaGetLastError()));
            fflush(stdout);
        };
        float ms = 0; unsigned long long hits = 0;
        for (int rep = 0; rep < 3; rep++) {
            a.visits = dva; a.nVisits = (uint32_t)va.size();
            cudaMemset(dHits, 0, 8);
            cudaEventRecord(e0);
            k_guide_major<<<n, 128>>>(a);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        }
        cudaMemcpy(&hits, dHits, 8, cudaMemcpyDeviceToHost);
        report("guide-major", 1, ms, hits);
        for (int stream = 0; stream < 2; stream++)
            for (uint32_t G : {4u, 8u, 16u, 32u, 64u}) {
                a.visits = dvb; a.G = G;
                uint32_t ctas = 0;
                for (uint32_t t = 0; t < 10; t++) { a.ctaFirst[t] = ctas; ctas += (n + G - 1) / G; }
                a.ctaFirst[10] = ctas;
                for (int rep = 0; rep < 3; rep++) {
                    cudaMemset(dHits, 0, 8);
                    cudaEventRecord(e0);
                    if (stream) k_triple_sorted<true><<<ctas, 128, G * 80>>>(a); else k_triple_sorted<false><<<ctas, 128, G * 80>>>(a);
                    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
                }
                cudaMemcpy(&hits, dHits, 8, cudaMemcpyDeviceToHost);
                report(stream ? "triple-sorted, ld.cs" : "triple-sorted, ld.nc", G, ms, hits);
            }
        cudaFree(dg); cudaFree(dorder);
    }
    return 0;
}
