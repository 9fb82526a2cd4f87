// C++ caller of the reference-style API of libturbosqueeze_b200.so, modelled on the reference's own tests
// (test/test.cpp:30-54 block API, :149-199 buffer API, :234-331 async jobs chained from callbacks, drain on destroy).
// Built and run by tests/test_gpu_parity.py::test_cpp_caller_of_the_reference_api.  Exit code 0 = all checks passed.
#include "tsq_b200.h"

#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

static std::vector<uint8_t> make_text(size_t n, uint32_t seed)
{
    static const char* words[] = {"the", "of", "and", "compression", "block", "literal", "match", "offset", "table", "hash",
                                  "<page>", "</page>", "[[link]]", "1999", "turbosqueeze", "warp", "lane", "byte", " ", "\n"};
    std::vector<uint8_t> v;
    v.reserve(n + 256);
    uint32_t s = seed;
    while (v.size() < n) {
        s = s * 1664525u + 1013904223u;
        const char* w = words[(s >> 24) % 20];
        v.insert(v.end(), w, w + strlen(w));
        v.push_back(' ');
    }
    v.resize(n);
    v.resize(n + 256, 0);            // readable slack behind the data, as the reference's callers provide
    return v;
}

#define CHECK(cond)                                                              \
    do {                                                                         \
        if (!(cond)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } \
    } while (0)

int main()
{
    // ---- block API (test/test.cpp:30-54), both formats
    {
        std::vector<uint8_t> in = make_text(699, 1), comp(2048), out(2048);
        TSQCompressionContext* ctx = tsqAllocateContext();
        CHECK(ctx && ctx->refhash);
        for (uint32_t ext = 0; ext < 2; ext++) {
            uint32_t n = 0, m = 0;
            tsqInit(ctx);
            tsqEncode(ctx, in.data(), comp.data(), &n, 699, ext);
            CHECK(n > 0 && n < 900);
            tsqDecode(comp.data(), out.data(), &m, n, ext);
            CHECK(m == 699 && memcmp(out.data(), in.data(), 699) == 0);
        }
        tsqDeallocateContext(ctx);
    }
    const size_t N = (9u << 20) + 1234;
    std::vector<uint8_t> text = make_text(N, 7);
    // ---- synchronous buffer API (test/test.cpp:149-199)
    {
        TSQCompressionContext_MT* c = tsqAllocateContextCompression_MT(false);
        TSQDecompressionContext_MT* d = tsqAllocateContextDecompression_MT(false);
        CHECK(c && d);
        for (int ext = 0; ext < 2; ext++) {
            uint8_t *blob = nullptr, *back = nullptr;
            size_t bn = 0, on = 0;
            CHECK(tsqCompress_MT(c, text.data(), N, false, &blob, &bn, false, ext != 0, 0));
            CHECK(blob && bn > 16 && memcmp(blob, "TSQ1", 4) == 0);
            CHECK(tsqDecompress_MT(d, blob, bn, false, &back, &on, false));
            CHECK(on == N && memcmp(back, text.data(), N) == 0);
            free(blob); free(back);
        }
        CHECK(!tsqCompress_MT(c, nullptr, 10, false, nullptr, nullptr, false, false, 0));      // bad arguments -> false
        tsqDeallocateContextCompression_MT(c);
        tsqDeallocateContextDecompression_MT(d);
    }
    // ---- async jobs: decompression launched from inside the compression callback (test/test.cpp:234-270),
    // many jobs in flight, contexts destroyed while jobs are queued (drain on destroy, :202-231)
    {
        TSQCompressionContext_MT* c = tsqAllocateContextCompression_MT(false);
        TSQDecompressionContext_MT* d = tsqAllocateContextDecompression_MT(false);
        const int J = 12;
        std::vector<uint8_t*> blob(J, nullptr), back(J, nullptr);
        std::vector<size_t> bn(J, 0), on(J, 0);
        std::atomic<int> done{0}, good{0}, progress_calls{0};
        std::mutex m;
        std::condition_variable cv;
        for (int j = 0; j < J; j++) {
            const uint32_t id = tsqCompressAsync_MT(
                c, text.data(), N - 1000 * j, false, &blob[j], &bn[j], false, (j & 1) != 0, 0,
                [&, j](uint32_t jobid, bool ok) {
                    if (!ok || jobid == 0) { done++; cv.notify_all(); return; }
                    tsqDecompressAsync_MT(d, blob[j], bn[j], false, &back[j], &on[j], false,
                                          [&, j](uint32_t, bool ok2) {
                                              if (ok2 && on[j] == N - 1000 * j && memcmp(back[j], text.data(), on[j]) == 0) good++;
                                              done++;
                                              cv.notify_all();
                                          },
                                          [&](uint32_t, double) { progress_calls++; });
                },
                [&](uint32_t, double p) { if (p >= 0.0 && p <= 1.0) progress_calls++; });
            CHECK(id == (uint32_t)(j + 1));
        }
        {
            std::unique_lock<std::mutex> lk(m);
            cv.wait(lk, [&] { return done.load() == J; });
        }
        CHECK(good.load() == J);
        CHECK(progress_calls.load() >= 2 * J);
        // early failure: id 0 and completion(0, false)
        bool early = false;
        CHECK(tsqCompressAsync_MT(c, nullptr, 0, false, nullptr, nullptr, false, false, 0, [&](uint32_t id, bool ok) { early = id == 0 && !ok; }, nullptr) == 0);
        CHECK(early);
        // jobs still queued when the context goes away must complete first
        std::atomic<int> late{0};
        uint8_t* b2[4] = {nullptr, nullptr, nullptr, nullptr};
        size_t n2[4] = {0, 0, 0, 0};
        for (int j = 0; j < 4; j++)
            tsqCompressAsync_MT(c, text.data(), N, false, &b2[j], &n2[j], false, false, 0, [&](uint32_t, bool ok) { if (ok) late++; }, nullptr);
        tsqDeallocateContextCompression_MT(c);
        CHECK(late.load() == 4);
        tsqDeallocateContextDecompression_MT(d);
        for (int j = 0; j < J; j++) { free(blob[j]); free(back[j]); }
        for (int j = 0; j < 4; j++) free(b2[j]);
    }
    // ---- the callback contract of the reference's writer thread (tsq_threads.cpp:248-268, :654-668): one progress call per
    // block with (blocks written) / n_blocks, in order, then the completion call; jobs of one context deliver their callbacks
    // in submission order even when two of them run on the GPU at once
    {
        TSQCompressionContext_MT* c = tsqAllocateContextCompression_MT(false);
        TSQDecompressionContext_MT* d = tsqAllocateContextDecompression_MT(false);
        const int J = 6;
        const size_t n_blocks = (N + (4u << 20) - 1) / (4u << 20);                 // 4 MiB blocks (TSQ_BLOCK_SZ): 3 for this input
        std::vector<uint8_t*> blob(J, nullptr);
        std::vector<size_t> bn(J, 0);
        std::mutex m;
        std::vector<std::pair<uint32_t, double>> events;                          // (job id, progress) or (job id, 2.0 = completed ok)
        for (int j = 0; j < J; j++)
            tsqCompressAsync_MT(c, text.data(), N - 4096 * j, false, &blob[j], &bn[j], false, false, 0,
                                [&](uint32_t id, bool ok) { std::lock_guard<std::mutex> lk(m); events.push_back({id, ok ? 2.0 : -1.0}); },
                                [&](uint32_t id, double p) { std::lock_guard<std::mutex> lk(m); events.push_back({id, p}); });
        tsqDeallocateContextCompression_MT(c);                                    // drains
        CHECK(events.size() == (size_t)J * (n_blocks + 1));
        for (int j = 0; j < J; j++)
            for (size_t e = 0; e <= n_blocks; e++) {
                const auto& ev = events[(size_t)j * (n_blocks + 1) + e];
                CHECK(ev.first == (uint32_t)(j + 1));                             // strictly job by job
                if (e < n_blocks) CHECK(ev.second > (double)e / n_blocks && ev.second <= (double)(e + 1) / n_blocks + 1e-9);
                else CHECK(ev.second == 2.0);                                     // completion last, after progress reached 1.0
            }
        // the same for decompression
        events.clear();
        std::vector<uint8_t*> back(J, nullptr);
        std::vector<size_t> on(J, 0);
        for (int j = 0; j < J; j++)
            tsqDecompressAsync_MT(d, blob[j], bn[j], false, &back[j], &on[j], false,
                                  [&](uint32_t id, bool ok) { std::lock_guard<std::mutex> lk(m); events.push_back({id, ok ? 2.0 : -1.0}); },
                                  [&](uint32_t id, double p) { std::lock_guard<std::mutex> lk(m); events.push_back({id, p}); });
        tsqDeallocateContextDecompression_MT(d);
        CHECK(events.size() == (size_t)J * (n_blocks + 1));
        for (int j = 0; j < J; j++) {
            CHECK(events[(size_t)j * (n_blocks + 1) + n_blocks] == std::make_pair((uint32_t)(j + 1), 2.0));
            CHECK(on[j] == N - 4096 * j && memcmp(back[j], text.data(), on[j]) == 0);
            free(blob[j]); free(back[j]);
        }
    }
    // ---- tsqDecode with inputSize == 0 (the reference ignores inputSize, tsq_decode.cpp:42-126): the stream is the last thing
    // in its allocation, so a library that read a worst-case slot from the caller's buffer would run off it
    {
        std::vector<uint8_t> in = make_text(70000, 3), comp(90000);
        TSQCompressionContext* ctx = tsqAllocateContext();
        uint32_t n = 0, mm = 0;
        tsqInit(ctx);
        tsqEncode(ctx, in.data(), comp.data(), &n, 70000, 0);
        CHECK(n > 0);
        uint8_t* exact = (uint8_t*)malloc(n);                                     // exactly the stream, nothing behind it
        memcpy(exact, comp.data(), n);
        std::vector<uint8_t> out(70000 + 256);
        tsqDecode(exact, out.data(), &mm, 0, 0);
        CHECK(mm == 70000 && memcmp(out.data(), in.data(), 70000) == 0);
        free(exact);
        tsqDeallocateContext(ctx);
    }
    // ---- FILE* entry points (turbosqueeze.cpp:48-147)
    {
        const char* raw = "/tmp/tsqb_async_raw.bin"; const char* tsq = "/tmp/tsqb_async.tsq"; const char* rt = "/tmp/tsqb_async_rt.bin";
        FILE* f = fopen(raw, "wb"); CHECK(f); fwrite(text.data(), 1, N, f); fclose(f);
        FILE* fi = fopen(raw, "rb"); FILE* fo = fopen(tsq, "wb"); CHECK(fi && fo);
        tsqCompress(fi, fo, true, 0); fclose(fi); fclose(fo);
        fi = fopen(tsq, "rb"); fo = fopen(rt, "wb"); CHECK(fi && fo);
        tsqDecompress(fi, fo); fclose(fi); fclose(fo);
        f = fopen(rt, "rb"); CHECK(f);
        std::vector<uint8_t> back(N + 1);
        const size_t got = fread(back.data(), 1, N + 1, f); fclose(f);
        CHECK(got == N && memcmp(back.data(), text.data(), N) == 0);
        remove(raw); remove(tsq); remove(rt);
    }
    printf("async_harness ok\n");
    return 0;
}
