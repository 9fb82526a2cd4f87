// Development helper: the DRAM ceiling for the encoder's access pattern -- one 32-byte sector per request at
// pseudo-random addresses of a buffer far larger than L2.  nvcc -O3 -arch=sm_100a scripts/random_gather.cu -o /tmp/rg && /tmp/rg
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void gather(const uint4* __restrict__ buf, uint64_t n_sectors, uint32_t per_thread, uint32_t* out)
{
    uint64_t s = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
    uint32_t acc = 0;
    for (uint32_t i = 0; i < per_thread; i += 4) {
        uint4 a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            const uint64_t sec = (s >> 20) % n_sectors;
            a[u] = buf[2 * sec]; b[u] = buf[2 * sec + 1];
        }
#pragma unroll
        for (int u = 0; u < 4; u++) acc += a[u].x ^ b[u].w;
    }
    if (acc == 0xdeadbeef) out[0] = acc;
}

int main()
{
    const uint64_t bytes = 16ull << 30, n_sectors = bytes / 32;
    uint4* buf; uint32_t* out;
    cudaMalloc(&buf, bytes); cudaMalloc(&out, 4); cudaMemset(buf, 1, bytes);
    for (int warps_per_sm : {8, 16, 32, 64}) {
        const int threads = 256, blocks = 148 * warps_per_sm * 32 / threads;
        const uint32_t per_thread = 2048;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        gather<<<blocks, threads>>>(buf, n_sectors, 64, out);
        cudaEventRecord(e0);
        gather<<<blocks, threads>>>(buf, n_sectors, per_thread, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double req = (double)blocks * threads * per_thread;
        printf("warps/SM %2d: %.2f G sectors/s  = %.1f GB/s of 32-byte sectors (%.1f GB/s if DRAM moves 64 B each)\n", warps_per_sm,
               req / ms / 1e6, req * 32 / ms / 1e6, req * 64 / ms / 1e6);
    }
    return 0;
}
