/*
 * ref_harness.cpp -- thin C entry points around the UNMODIFIED reference
 * (compiled from /root/reference by oracle/Makefile into oracle/_ref/).
 *
 * TEST INFRASTRUCTURE ONLY (see tsq_oracle.c header).  Nothing from the
 * reference is copied: this file only #includes its public header and calls
 * tsqInit/tsqEncode/tsqDecode (turbosqueeze.h:643-670) and the MT buffer API
 * (turbosqueeze.h:508,580) under the parity contract of SURVEY.md 8(c).
 */
#include <cstdio>
#include <cstdlib>
#include "turbosqueeze.h"
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

extern "C" {

uint32_t ref_encode_block(const uint8_t* in, uint32_t size, uint8_t* out, uint32_t with_ext)
{
    TSQCompressionContext* ctx = tsqAllocateContext();
    uint32_t n = 0;
    tsqInit(ctx);
    tsqEncode(ctx, const_cast<uint8_t*>(in), out, &n, size, with_ext);
    tsqDeallocateContext(ctx);
    return n;
}

/* out needs >= 256 bytes of slack after `size` (the reference over-writes). */
uint32_t ref_decode_block(const uint8_t* in, uint32_t in_size, uint8_t* out, uint32_t with_ext)
{
    uint32_t n = 0;
    tsqDecode(const_cast<uint8_t*>(in), out, &n, in_size, with_ext);
    return n;
}

/* Blocks are sub-ranges of one buffer, encoded in place (tsq_threads.cpp:109);
 * each output slot is zero-filled first.  threads<=0 -> hardware_concurrency. */
double ref_encode_blocks(const uint8_t* buf, uint64_t total, uint32_t block, uint8_t* out, uint64_t stride,
                         uint32_t* sizes, uint32_t with_ext, int threads, int zero_slots)
{
    const uint64_t nb = (total + block - 1) / block;
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if ((uint64_t)threads > nb) threads = (int)(nb ? nb : 1);
    if (zero_slots) memset(out, 0, nb * stride);
    std::atomic<uint64_t> next{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&] {
            TSQCompressionContext* ctx = tsqAllocateContext();
            for (;;) {
                uint64_t b = next.fetch_add(1);
                if (b >= nb) break;
                uint64_t at = b * (uint64_t)block;
                uint32_t n = (uint32_t)((total - at < block) ? total - at : block);
                tsqInit(ctx);                                   /* tsq_threads.cpp:176 */
                tsqEncode(ctx, const_cast<uint8_t*>(buf + at), out + b * stride, &sizes[b], n, with_ext);
            }
            tsqDeallocateContext(ctx);
        });
    for (auto& th : pool) th.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

/* comp slots at `stride`; decoded block b goes to out + b*out_stride (out_stride >= block + 256). */
double ref_decode_blocks(const uint8_t* comp, uint64_t stride, const uint32_t* comp_sizes, uint64_t nb, uint8_t* out,
                         uint64_t out_stride, uint32_t* sizes, uint32_t with_ext, int threads)
{
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if ((uint64_t)threads > nb) threads = (int)(nb ? nb : 1);
    std::atomic<uint64_t> next{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&] {
            for (;;) {
                uint64_t b = next.fetch_add(1);
                if (b >= nb) break;
                tsqDecode(const_cast<uint8_t*>(comp + b * stride), out + b * out_stride, &sizes[b],
                          comp_sizes ? comp_sizes[b] : 0, with_ext);   /* tsq_threads.cpp:590 */
            }
        });
    for (auto& th : pool) th.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

/* The reference's own MT pipeline, memory -> memory (4 MiB blocks, TSQ1 container).
 * *out is malloc'ed by the reference; release with ref_free. */
int ref_compress_mt(const uint8_t* in, uint64_t n, uint8_t** out, uint64_t* out_n, int with_ext)
{
    TSQCompressionContext_MT* ctx = tsqAllocateContextCompression_MT(false);
    size_t sz = 0;
    bool ok = tsqCompress_MT(ctx, const_cast<uint8_t*>(in), n, false, out, &sz, false, with_ext != 0, 0);
    tsqDeallocateContextCompression_MT(ctx);
    *out_n = sz;
    return ok ? 1 : 0;
}

int ref_decompress_mt(const uint8_t* in, uint64_t n, uint8_t** out, uint64_t* out_n)
{
    TSQDecompressionContext_MT* ctx = tsqAllocateContextDecompression_MT(false);
    size_t sz = 0;
    bool ok = tsqDecompress_MT(ctx, const_cast<uint8_t*>(in), n, false, out, &sz, false);
    tsqDeallocateContextDecompression_MT(ctx);
    *out_n = sz;
    return ok ? 1 : 0;
}

void ref_free(void* p) { free(p); }

int ref_hw_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
