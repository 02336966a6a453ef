// random_read_bw.cu -- what HBM delivers for the access pattern of the bucket scan: independent, aligned 128-byte
// reads at random addresses of a 21.5 GB array (two lanes x 4 x 16 bytes per read, streaming loads), nothing else.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/random_read_bw tools/random_read_bw.cu
// Run on the GPU box: tools/_build/random_read_bw  -> JSON lines {ctas_per_sm, threads, GB/s}
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

template <int UNROLL>
__global__ void k_random_reads(const uint4 *__restrict__ base, uint64_t nBlocks, uint32_t readsPerPair, uint32_t *sink)
{
    const uint64_t pair = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
    const uint32_t sub = threadIdx.x & 1u;
    uint32_t acc = 0;
    for (uint32_t r = 0; r < readsPerPair; r += UNROLL) {
        uint4 q[UNROLL][4];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const uint64_t b = ((mix64(pair * 0x10001ull + r + u) >> 32) * nBlocks) >> 32;   // nBlocks < 2^32
            const uint4 *p = base + (b * 2 + sub) * 4;
            q[u][0] = __ldcs(p); q[u][1] = __ldcs(p + 1); q[u][2] = __ldcs(p + 2); q[u][3] = __ldcs(p + 3);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
            acc ^= q[u][0].x ^ q[u][1].y ^ q[u][2].z ^ q[u][3].w;
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

int main()
{
    const uint64_t nBlocks = 10ull << 24;   // 167.8 M blocks of 128 B = 21.5 GB, as the blocked copy at human scale
    uint4 *buf; uint32_t *sink;
    if (cudaMalloc(&buf, nBlocks * 128) != cudaSuccess) { printf("{\"error\": \"cudaMalloc\"}\n"); return 1; }
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, nBlocks * 128);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int sms = 148;
    // shared memory per CTA is taken from L1, and L1 holds the lines of the loads in flight: the sweep shows what that costs
    const int smemKb[] = {1, 6, 14, 21, 28};
    for (int unroll = 1; unroll <= 2; unroll++)
        for (int si = 0; si < 5; si++) {
            const int threads = 128, grid = sms * 16 * 8;
            const uint32_t reads = 256;
            const double bytes = (double)grid * threads / 2 * reads * 128;
            const size_t smem = (size_t)smemKb[si] * 1024;
            int resident = 0;
            float ms = 0;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(e0);
                if (unroll == 1) {
                    cudaFuncSetAttribute(k_random_reads<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_random_reads<1>, threads, smem);
                    k_random_reads<1><<<grid, threads, smem>>>(buf, nBlocks, reads, sink);
                } else {
                    cudaFuncSetAttribute(k_random_reads<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_random_reads<2>, threads, smem);
                    k_random_reads<2><<<grid, threads, smem>>>(buf, nBlocks, reads, sink);
                }
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1);
            }
            printf("{\"loads_in_flight_per_lane\": %d, \"smem_kb_per_cta\": %d, \"ctas_per_sm\": %d, \"threads\": %d, \"ms\": %.3f, \"GB/s\": %.1f, \"error\": \"%s\"}\n",
                   4 * unroll, smemKb[si], resident, threads, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
