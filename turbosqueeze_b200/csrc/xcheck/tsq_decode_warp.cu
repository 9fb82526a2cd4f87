// tsq_decode_warp.cu (CROSS-CHECK ONLY: test library tests/xcheck/libturbosqueeze_b200_xcheck.so, not the product) -- one WARP per block, 32 symbols per step, stream staged in shared memory.
//
// Semantics: reference tsqDecodeNoext (tsq_decode.cpp:42-126), bit-exact on [0, size).
//
// The token walk is a serial chain -- the address of every size byte depends on the payload
// lengths of the pair before it (tsq_decode.cpp:68-86) -- so one block can never go faster than
// (pairs x chain latency).  The kernel therefore (a) keeps that chain as short as possible: the
// compressed stream is staged into a per-warp shared-memory ring by 1-D bulk async copies
// (cp.async.bulk + mbarrier, i.e. TMA without a tensor map), so each hop is one LDS + a handful of
// ALU ops, and (b) takes everything else off the chain: the walk only records (stream position,
// output position, control bits) of up to 16 pairs, then all 32 lanes -- one per symbol -- fetch
// their own size nibble / offset and move their own <= 16 bytes in parallel.
//
// Symbols whose source lies inside the output of the same step (near matches) wait for the stores
// of the pairs before them: they run in follow-up rounds, each round releasing every symbol whose
// source ends before the first still-pending pair (sources always precede their own pair,
// tsq_encode.cpp:139-141, so every round makes progress).
#include "../tsq_device.cuh"

namespace tsqb {

namespace {

constexpr unsigned FULL      = 0xffffffffu;
constexpr uint32_t kChunk    = 1024;                 // bytes per bulk copy
constexpr uint32_t kChunks   = 4;                    // ring slots
constexpr uint32_t kRing     = kChunk * kChunks;     // 4 KiB of stream per warp
constexpr uint32_t kRingMask = kRing - 1;
constexpr uint32_t kPairs    = 16;                   // pairs per step (32 symbols)
constexpr uint32_t kStepSpan = kPairs * 33 + 4 + 16; // most stream bytes one step can touch
constexpr int      kWarps    = 8;                    // warps per CTA

struct __align__(16) WarpSmem {
    uint8_t  ring[kRing];
    uint2    desc[kPairs];
    uint64_t bar[kChunks];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

}  // namespace

__global__ void __launch_bounds__(kWarps * 32) decode_warp_kernel(DecodeArgs a)
{
    __shared__ WarpSmem sm_all[kWarps];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned wid  = threadIdx.x >> 5;
    WarpSmem& sm = sm_all[wid];
    const uint8_t* ring = sm.ring;

    if (lane == 0) {
        for (uint32_t q = 0; q < kChunks; q++) mbar_init(&sm.bar[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const uint64_t nwarps = (uint64_t)gridDim.x * kWarps;
    uint32_t phase = 0;                    // bit s: parity the next completion of ring slot s will have

    for (uint64_t b = (uint64_t)blockIdx.x * kWarps + wid; b < a.nb; b += nwarps) {
        const uint8_t* src = a.comp + (a.offs ? a.offs[b] : b * a.stride);
        uint8_t* __restrict__ o = a.out + b * a.ostride;
        const uint32_t limit = a.csizes ? a.csizes[b] : (a.stride > 0xffffffffull ? 0xffffffffu : (uint32_t)a.stride);

        // the ring holds the 16-byte aligned stream: stream byte k lives at ring[(k + shift) & mask];
        // chunk n (bytes [n*kChunk, ...) of the aligned stream) uses ring slot and mbarrier n % kChunks
        const uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
        const uint8_t* src_al = src - shift;
        const uint32_t total_al = (shift + limit + 15u) & ~15u;                 // bytes worth staging
        const uint32_t nchunks = (total_al + kChunk - 1) / kChunk;
        uint32_t issued = 0, waited = 0;                                         // chunks issued / observed complete (warp-uniform)

        auto issue_upto = [&](uint32_t want) {                                   // all lanes call; lane 0 copies
            want = min(want, nchunks);
            if (lane == 0)
                for (uint32_t n = issued; n < want; n++) {
                    const uint32_t at = n * kChunk, bytes = min(kChunk, total_al - at);
                    mbar_expect_tx(&sm.bar[n % kChunks], bytes);
                    bulk_load(sm.ring + (at & kRingMask), src_al + at, bytes, &sm.bar[n % kChunks]);
                }
            issued = max(issued, want);
        };
        auto wait_upto = [&](uint32_t want) {                                    // all lanes wait
            want = min(want, issued);
            for (; waited < want; waited++) {
                const uint32_t s = waited % kChunks;
                mbar_wait(&sm.bar[s], (phase >> s) & 1u);
                phase ^= 1u << s;
            }
        };

        __syncwarp();                                                            // previous stream fully consumed
        issue_upto(kChunks);
        wait_upto(1);

        auto rb = [&](uint32_t k) -> uint32_t { return ring[(k + shift) & kRingMask]; };

        const uint32_t size = rb(0) | (rb(1) << 8) | (rb(2) << 16);              // tsq_decode.cpp:49-51
        const bool ok = size <= kBlockMax && size <= a.ostride;
        if (lane == 0) a.osizes[b] = ok ? size : 0u;

        uint32_t p = 3, j = 0;
        while (ok && j < size && p < limit) {
            // ---- stream chunks this step may touch are resident; recycle the slots behind p
            {
                const uint32_t cur = (p + shift) / kChunk;                       // chunk holding p; slots behind it are dead
                issue_upto(cur + kChunks);
                wait_upto((p + shift + kStepSpan) / kChunk + 1u);
            }

            // ---- serial token walk: up to 16 pairs; records where each pair starts (:60-90)
            const uint32_t J0 = j;
            uint32_t np = 0;
#pragma unroll 1
            for (int g = 0; g < 4; g++) {
                if (!(j < size && p < limit)) break;
                uint32_t ctl = rb(p);                                            // :62
                p++;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (j >= size) break;
                    const uint32_t nib = rb(p);                                  // :68
                    if (lane == 0) sm.desc[np] = make_uint2(p | ((ctl & 0xC0u) << 24), j);
                    const uint32_t n0 = nib >> 4, n1 = nib & 15u;
                    const uint32_t pay0 = (ctl & 0x80u) ? n0 + 1u : 2u;
                    const uint32_t pay1 = (ctl & 0x40u) ? n1 + 1u : 2u;
                    ctl <<= 2;
                    const uint32_t j1 = j + n0 + 1u;
                    const bool two = j1 < size;                                  // second symbol exists
                    p += 1u + pay0 + (two ? pay1 : 0u);
                    j = j1 + (two ? n1 + 1u : 0u);
                    np++;
                }
            }
            __syncwarp();

            // ---- one lane per symbol
            const uint32_t pi = lane >> 1, half = lane & 1u;
            bool active = pi < np;
            uint32_t len = 0, dst = 0, sp = 0, srcpos = 0;
            bool lit = true;
            if (active) {
                const uint2 d = sm.desc[pi];
                const uint32_t pp = d.x & 0x3FFFFFFFu, jp = d.y;
                const uint32_t nib = rb(pp);
                const uint32_t n0 = nib >> 4, n1 = nib & 15u;
                const bool l0 = (d.x & 0x80000000u) != 0, l1 = (d.x & 0x40000000u) != 0;
                if (half == 0) { lit = l0; len = n0 + 1u; sp = pp + 1u; dst = jp; }
                else { lit = l1; len = n1 + 1u; sp = pp + 1u + (l0 ? n0 + 1u : 2u); dst = jp + n0 + 1u; }
                active = dst < size;
                len = min(len, size - dst);
                if (active && !lit) {
                    const uint32_t off = rb(sp) | (rb(sp + 1u) << 8);            // :69,73,82
                    active = off <= jp;                                          // corrupt stream guard
                    srcpos = jp - off;
                }
            }

            // ---- round 0: literals and matches whose source precedes this step's output
            uint32_t v[4] = {0, 0, 0, 0};
            bool now = active && (lit || srcpos + len <= J0);
            bool pending = active && !now;
            for (;;) {
                if (now) {
                    if (lit) {
                        const uint32_t a0 = (sp + shift) & ~3u, sh = ((sp + shift) & 3u) * 8u;
                        const uint32_t* r32 = reinterpret_cast<const uint32_t*>(ring);
                        uint32_t w[5];
#pragma unroll
                        for (int m = 0; m < 5; m++) w[m] = r32[((a0 + 4u * m) & kRingMask) >> 2];
#pragma unroll
                        for (int m = 0; m < 4; m++) v[m] = __funnelshift_r(w[m], w[m + 1], sh);
                    } else {
                        const uintptr_t ad = reinterpret_cast<uintptr_t>(o + srcpos);
                        const uint32_t* g32 = reinterpret_cast<const uint32_t*>(ad & ~(uintptr_t)3);
                        const uint32_t sh = (uint32_t)(ad & 3u) * 8u;
                        uint32_t w[5];
#pragma unroll
                        for (int m = 0; m < 5; m++)                              // never touch a word past the source
                            w[m] = ((uint32_t)(ad & 3u) + len > 4u * m) ? g32[m] : 0u;
#pragma unroll
                        for (int m = 0; m < 4; m++) v[m] = __funnelshift_r(w[m], w[m + 1], sh);
                    }
                    uint8_t* d8 = o + dst;
#pragma unroll
                    for (int t = 0; t < 16; t++)
                        if ((uint32_t)t < len) d8[t] = (uint8_t)(v[t >> 2] >> (8 * (t & 3)));
                }
                const uint32_t pm = __ballot_sync(FULL, pending);
                if (pm == 0) break;
                __syncwarp();                                                    // stores above are visible to the warp
                // every pending symbol whose source ends before the first pending pair is now safe
                const uint32_t firstlane = (uint32_t)__ffs((int)pm) - 1u;
                const uint32_t frontier = __shfl_sync(FULL, dst, firstlane & ~1u);   // even lane's dst == start of that pair
                // the first pending pair itself is always released (its sources precede its own start in
                // every valid stream; releasing it unconditionally also bounds the loop on corrupt input)
                now = pending && (srcpos + len <= frontier || (lane >> 1) == (firstlane >> 1));
                pending = pending && !now;
            }
            __syncwarp();
        }
        wait_upto(issued);                                                       // drain before the ring is reused
        __syncwarp();
    }
}

cudaError_t launch_decode_warp(const DecodeArgs& a, int sm_count, cudaStream_t st)
{
    if (a.nb == 0) return cudaSuccess;
    uint64_t ctas = (a.nb + kWarps - 1) / kWarps;
    const uint64_t cap = (uint64_t)sm_count * 8u;                     // 8 CTAs x 8 warps = 64 warps / SM
    if (ctas > cap) ctas = cap;
    decode_warp_kernel<<<(unsigned)ctas, kWarps * 32, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace tsqb
