// Development helper: does the random-sector ceiling (scripts/random_gather.cu) depend on the FOOTPRINT (TLB reach,
// DRAM page locality) and on the read / write mix?  One 32-byte sector per request.
// nvcc -O3 -arch=sm_100a scripts/random_gather3.cu -o /tmp/rg3 && /tmp/rg3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// every `wr_every`-th request is a full-sector store instead of a load (0: loads only)
__global__ void mix(uint4* __restrict__ buf, uint64_t n_sectors, uint32_t per_thread, uint32_t wr_every, uint32_t* out)
{
    uint64_t s = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
    uint32_t acc = 0;
    for (uint32_t i = 0; i < per_thread; i += 4) {
        uint4 a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            const uint64_t sec = (s >> 20) % n_sectors;
            if (wr_every && ((i + u) % wr_every) == 0) {
                const uint4 v = make_uint4((uint32_t)s, i, u, 7);
                asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(buf + 2 * sec), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w),
                             "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                a[u] = v; b[u] = v;
            } else {
                asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a[u].x), "=r"(a[u].y), "=r"(a[u].z), "=r"(a[u].w),
                             "=r"(b[u].x), "=r"(b[u].y), "=r"(b[u].z), "=r"(b[u].w) : "l"(buf + 2 * sec) : "memory");
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) acc += a[u].x ^ b[u].w;
    }
    if (acc == 0xdeadbeef) out[0] = acc;
}

int main()
{
    const uint64_t maxbytes = 16ull << 30;
    uint4* buf; uint32_t* out;
    cudaMalloc(&buf, maxbytes); cudaMalloc(&out, 4); cudaMemset(buf, 1, maxbytes);
    const int threads = 256, blocks = 148 * 32 * 32 / threads;
    const uint32_t per_thread = 1024;
    for (uint64_t mb : {512ull, 1024ull, 2048ull, 4096ull, 8192ull, 16384ull})
        for (uint32_t wr : {0u, 4u, 3u, 2u}) {
            const uint64_t n_sectors = (mb << 20) / 32;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            mix<<<blocks, threads>>>(buf, n_sectors, 64, wr, out);
            cudaEventRecord(e0);
            mix<<<blocks, threads>>>(buf, n_sectors, per_thread, wr, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double req = (double)blocks * threads * per_thread;
            printf("footprint %5llu MB, stores 1 in %u: %.2f G requests/s\n", (unsigned long long)mb, wr, req / ms / 1e6);
        }
    return 0;
}
