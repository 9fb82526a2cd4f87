// Development helper: does a load qualifier change what a random sector read costs?  (see random_gather.cu)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__device__ __forceinline__ uint4 ld(const uint4* p)
{
    uint4 v;
    if (MODE == 0) asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 1) asm volatile("ld.global.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 2) asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 3) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <int MODE>
__global__ void gather(const uint4* __restrict__ buf, uint64_t n_sectors, uint32_t per_thread, uint32_t* out)
{
    uint64_t s = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
    uint32_t acc = 0;
    for (uint32_t i = 0; i < per_thread; i += 8) {
        uint4 a[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            a[u] = ld<MODE>(buf + 2 * ((s >> 20) % n_sectors));
        }
#pragma unroll
        for (int u = 0; u < 8; u++) acc += a[u].x;
    }
    if (acc == 0xdeadbeef) out[0] = acc;
}

template <int MODE>
void run(const char* name, const uint4* buf, uint64_t n_sectors, uint32_t* out)
{
    const int threads = 256, blocks = 148 * 32 * 32 / threads;
    const uint32_t per_thread = 2048;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather<MODE><<<blocks, threads>>>(buf, n_sectors, 64, out);
    cudaEventRecord(e0);
    gather<MODE><<<blocks, threads>>>(buf, n_sectors, per_thread, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-40s %.2f G requests/s (%s)\n", name, (double)blocks * threads * per_thread / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const uint64_t bytes = 16ull << 30, n_sectors = bytes / 32;
    uint4* buf; uint32_t* out;
    cudaMalloc(&buf, bytes); cudaMalloc(&out, 4); cudaMemset(buf, 1, bytes);
    run<0>("ld.global (16 B of a sector)", buf, n_sectors, out);
    run<1>("ld.global.L2::64B", buf, n_sectors, out);
    run<2>("ld.global.cs", buf, n_sectors, out);
    run<3>("ld.global.nc.L1::no_allocate.L2::64B", buf, n_sectors, out);
    return 0;
}
