// tsq_decode_split.cu -- the production decoder (both formats): one WALKER lane + one COPIER warp per block.
//
// Semantics: reference tsqDecodeNoext (tsq_decode.cpp:42-126) and the extension variant (:137-314, template
// EXT: matches of 32 / 48 / 64 bytes), bit-exact on [0, size).
//
// Why two roles.  The token walk is a serial chain: the address of every size byte depends on the
// payload lengths of the pair before it (tsq_decode.cpp:68-86), ~50 cycles per pair even from
// shared memory, and it cannot be split inside a block.  Executed by a whole warp (as in
// tsq_decode_warp.cu) that chain costs 32 lanes' worth of issue slots per instruction.  Here the
// chains of up to 30 blocks are walked side by side by the lanes of the walker warps -- a lane walks
// the stream of one block slot of the CTA -- so the serial part costs a fraction of the issue slots, and it
// runs ahead of the copy work instead of alternating with it.  The walker publishes one 8-byte
// descriptor per pair (stream position of the size byte, output position, the pair's two control
// bits) into a per-slot shared-memory queue.
//
// The copier warp of a slot consumes one STEP of descriptors at a time.  Two copiers, chosen per block (the walker reads
// the cadence from the slot): copier_pairs() -- 32 descriptors = 64 symbols per step, lane L owns both symbols of pair L
// (no-extension format; text-like blocks) -- and copier() -- 16 descriptors = 32 symbols per step, one lane per symbol
// (extension format; (nearly) incompressible and very compressible blocks).  Either way:
//   * the compressed stream is staged into a shared-memory ring by 1-D bulk async copies
//     (cp.async.bulk + mbarrier: TMA without a tensor map; SASS UBLKCP) issued by the copier;
//   * decoded bytes go to a shared-memory OUTPUT ring first; near matches (distance < ring) read
//     their source there (29-cycle LDS instead of an L2 round trip), far matches read HBM/L2 bytes
//     that an earlier step flushed: one aligned 128-bit load per far source (a second one only when
//     the bytes cross its end), issued before and realigned after the step's shared-memory loads --
//     a scattered load costs one L1TEX wavefront per lane and instruction, and that pipe is the busiest
//     unit of this kernel (profiles/r02_experiments.md);
//   * every symbol stores exactly its own bytes: the 16 byte stores of a lane are predicated from a
//     length mask (ptxas turns the mask into predicates with two R2P, so this costs the same as the
//     reference's blind 16-byte copy, tsq_decode.cpp:74-85, and needs no ordering between lanes);
//   * symbols whose source lies inside the output of the same step are copied afterwards, in position
//     order, lane-per-byte with their exact length (sources always precede their own pair,
//     tsq_encode.cpp:139-141, so everything such a symbol reads is in place when it is reached); the
//     lane-per-pair copier takes both symbols of a pair at once (half a warp each: they never depend
//     on each other) with one packed shuffle per symbol;
//   * after the step, all complete 16-byte units of the output ring are written to HBM with
//     coalesced 128-bit stores.  Nothing is ever written past the block's decoded size.
// A descriptor is (stream position of the pair's size byte | the group's control byte << 24, output position);
// the copier lane picks its pair's two control bits from its own index in the step.
#include "tsq_device.cuh"

namespace tsqb {

namespace {

constexpr unsigned FULL      = 0xffffffffu;
#ifndef TSQB_DEC_CHUNK
#define TSQB_DEC_CHUNK 512         // bytes per bulk copy of the stream (the 4 KiB ring holds 4096 / TSQB_DEC_CHUNK of them)
#endif
constexpr uint32_t kChunk    = TSQB_DEC_CHUNK;       // bytes per bulk copy
constexpr uint32_t kInRing   = 4096;                 // 4 KiB of stream per block slot
constexpr uint32_t kChunks   = kInRing / kChunk;     // ring slots
constexpr uint32_t kInMask   = kInRing - 1;
#ifndef TSQB_DEC_BULK4
#define TSQB_DEC_BULK4 1           // development knob: walker takes 4 groups per limit test when far from every limit
#endif
#ifndef TSQB_DEC_QUEUE
#define TSQB_DEC_QUEUE 128
#endif
// round-2 development knobs (scripts/build_variants.sh; profiles/r02_experiments.md records what they measured)
#ifndef TSQB_DEC_FARNP
#define TSQB_DEC_FARNP 0           // lane-per-pair copier: far-match source words loaded without per-word predicates (2.73 -> 4.45 ms:
#endif                             // the far loads are bound by L1TEX wavefronts -- every scattered lane is its own 128-byte line)
#ifndef TSQB_DEC_FARWIDE
#define TSQB_DEC_FARWIDE 1         // far-match source = one aligned 128-bit load (a second one only when the bytes cross its end)
#endif                             // instead of up to five 32-bit loads: ~2.6 -> ~1.3 L1TEX wavefronts per far symbol
#ifndef TSQB_DEC_PEND2
#define TSQB_DEC_PEND2 1           // lane-per-pair copier: in-order copies take both symbols of a pair at once (half-warp each), one packed shuffle per symbol
#endif
#ifndef TSQB_DEC_FLUSH2
#define TSQB_DEC_FLUSH2 1          // lane-per-pair copier: flush of a step = two predicated 128-bit moves (a step leaves <= 65 units)
#endif
#ifndef TSQB_DEC_LIT16
#define TSQB_DEC_LIT16 1           // incompressible data: the walker checks a group of eight 16-byte literals with four independent loads
#endif                             // (its layout is fixed), and the lane-per-symbol copier stores aligned 16-byte symbols with one 128-bit store
#ifndef TSQB_DEC_DENSE_PAIRS
#define TSQB_DEC_DENSE_PAIRS 0     // 1: (nearly) incompressible blocks take the lane-per-pair copier too (random 1 GB: 1.19 vs 0.96 ms: worse)
#endif
#ifndef TSQB_DEC_EAGER
#define TSQB_DEC_EAGER 1           // walker notes a landed stream chunk as soon as it comes within the 4-group path's look-ahead of the
#endif                             // landed frontier (0: only within one pair's look-ahead -- which keeps it off the 4-group path for good;
                                   // 2: two chunks per iteration -- no faster once the one-group path has the literal shortcut, and small blocks lose)
#ifndef TSQB_DEC_L2POL
#define TSQB_DEC_L2POL 0           // L2 policies: bit 0 = far-match loads evict_first (their 64-byte fills are used once and push the freshly
#endif                             // written output -- the next far sources -- out of the L2), bit 1 = output stores evict_last, bit 2 = far loads fill 64 B, bit 3 = stream (TMA) loads evict_first
#ifndef TSQB_DEC_DESC2
#define TSQB_DEC_DESC2 1           // descriptor carries the group's whole control byte (the copier lane picks its pair's two bits from its own
#endif                             // index) and the walker publishes `produced` once per step: fewer walker instructions per pair
#ifndef TSQB_DEC_DIAG
#define TSQB_DEC_DIAG 0            // timing diagnostic, WRONG OUTPUT: 1 = the lane-per-pair copier moves no bytes (walker + hand-over + flush only)
#endif
#ifndef TSQB_DEC_WMASK
#define TSQB_DEC_WMASK 1           // walker: the 4-group path masks its ring addresses instead of requiring that it does not wrap (with TSQB_DEC_EAGER:
                                   // 2.600 -> 2.560 ms; without the mask the lanes of a walker warp split over two paths: 2.909 ms)
#endif
#ifndef TSQB_DEC_ONE_ROUND
#define TSQB_DEC_ONE_ROUND 1       // batches of 27..30 / 53..60 blocks per SM run as one / two rounds of up to 30 slots (28 KB of L1 left)
#endif
constexpr uint32_t kQueue    = TSQB_DEC_QUEUE;       // descriptors per slot
constexpr uint32_t kQMask    = kQueue - 1;
constexpr uint32_t kPairs    = 16;                   // pairs per copier step (32 symbols)
constexpr uint32_t kSteps    = kQueue / kPairs;      // steps the walker can be ahead
constexpr uint32_t kLook     = 40;                   // stream bytes one pair can touch (1 + 1 + 16 + 16) + slack
#ifndef TSQB_DEC_WALKERS
#define TSQB_DEC_WALKERS 2
#endif
constexpr uint32_t kWalkers  = TSQB_DEC_WALKERS;                    // walker warps per CTA, slots dealt round-robin (1 or 2: 2.91 ms, 4: 3.05, 6: 3.12)
#ifndef TSQB_DEC_LB
#define TSQB_DEC_LB 1024                             // development knob: launch bound (threads per CTA); 896 lets ptxas use 72 registers
#endif
constexpr uint32_t kMaxSlots = TSQB_DEC_LB / 32 - kWalkers;   // copier warps per CTA (copiers + walkers <= TSQB_DEC_LB threads)


template <uint32_t OUT_RING>
struct __align__(16) SlotSmem {
    uint8_t  in_ring[kInRing];
    uint8_t  out_ring[OUT_RING];
    uint2    desc[kQueue];
    uint64_t bar[kChunks];   // stream chunk landed (TMA transaction barriers)
    uint64_t full[kSteps];   // walker -> copier: one more step of descriptors is complete
    // block hand-over copier -> walker (written before `ready`)
    uint32_t ready;          // block index + 1 the slot is set up for
    uint32_t phase_bits;     // mbarrier parities at block start
    uint32_t limit_al;       // readable stream bytes, counted from the 16-byte aligned start
    uint32_t nchunks;
    uint32_t shift;          // stream start inside its first 16-byte unit
    uint32_t consumed;       // copier -> walker: descriptors of this block consumed so far
    uint32_t produced;       // walker -> copier: descriptors of this block published so far
    uint32_t ended;          // walker -> copier: the walk of this block is complete (set after the last `produced`)
    uint32_t end_j;          // output position the walk stopped at
    uint32_t step_pairs;     // copier -> walker: descriptors per step for this block (16: lane per symbol, 32: lane per pair)
    uint32_t pad[2];    // slot stride = 16 (mod 128): the walker's lanes do not pile onto 4 banks
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#if TSQB_DEC_L2POL
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
#endif

// 16 bytes of already flushed output (a far match source): around L1, optionally marked evict_first in L2
__device__ __forceinline__ uint4 ld_far16(const uint4* g)
{
#if TSQB_DEC_L2POL & 1
    uint4 v;
    const uint64_t pol = l2_policy_evict_first();
#if TSQB_DEC_L2POL & 4
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(g), "l"(pol) : "memory");
#else
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(g), "l"(pol) : "memory");
#endif
    return v;
#else
    return __ldcg(g);
#endif
}

// 16 bytes of decoded output to HBM (a later far match may read them back): optionally marked evict_last in L2
__device__ __forceinline__ void st_out16(uint8_t* p, const uint4& x)
{
#if TSQB_DEC_L2POL & 2
    const uint64_t pol = l2_policy_evict_last();
    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(x.x), "r"(x.y), "r"(x.z), "r"(x.w), "l"(pol) : "memory");
#else
    *reinterpret_cast<uint4*>(p) = x;
#endif
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait_addr(uint32_t bar_addr, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar_addr), "r"(parity), "r"(1000000u) : "memory");
}

// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
#if TSQB_DEC_L2POL & 8
    // the compressed stream is read exactly once: do not let it displace the output the far matches read back
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
#else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
#endif
}

__device__ __forceinline__ uint32_t ld_vol_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}

__device__ __forceinline__ void st_vol_u32(uint32_t* p, uint32_t v)
{
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

__device__ __forceinline__ uint2 ld_vol_u64(const uint2* p)
{
    uint2 v;
    asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(smem_u32(p)) : "memory");
    return v;
}

// decoded length of a symbol: size nibble + 1, except that in the extension format a match with nibble 0 / 1 / 2
// is 32 / 48 / 64 bytes long (tsq_decode.cpp:174-187)
template <bool EXT>
__device__ __forceinline__ uint32_t sym_len(uint32_t nibble, bool lit)
{
    if (EXT && !lit && nibble < 3u) return 16u * (nibble + 2u);
    return nibble + 1u;
}

// where block b's stream starts / how many bytes of it may be read
__device__ __forceinline__ const uint8_t* stream_of(const DecodeArgs& a, uint64_t b, uint32_t& limit)
{
    limit = a.csizes ? a.csizes[b] : (a.stride > 0xffffffffull ? 0xffffffffu : (uint32_t)a.stride);
    return a.comp + (a.offs ? a.offs[b] : b * a.stride);
}

// ------------------------------------------------------------------------------------------ walker
// Lane L walks the token stream of slot L.  All lanes run the same loop; a lane that has to wait
// (stream chunk not landed, descriptor queue full, copier still setting the block up) simply does
// nothing in that iteration.
// LEAN: the flavour for batches of small blocks (< 64 KiB: the 30-slot launch), whose walks are mostly start-up -- no early
// chunk polling, no masked 4-group path, no literal-group check (8 GiB of text in 16 KiB blocks: 28.0 -> 23.4 ms without them).
template <uint32_t OUT_RING, bool EXT, bool LEAN>
__device__ void walker(const DecodeArgs& a, SlotSmem<OUT_RING>* slots, uint32_t nslots, uint32_t widx, unsigned lane)
{
    constexpr bool kEager = TSQB_DEC_EAGER && !LEAN, kWMask = TSQB_DEC_WMASK && !LEAN;
    constexpr bool kLit16 = TSQB_DEC_LIT16 && TSQB_DEC_DESC2 && !LEAN && !EXT;
    // LEAN: how close to the landed frontier the walk comes before it looks for the next chunk.  Blocks of <= 8 KiB (a stream of a
    // few chunks) gain from looking early (1 GB in 4 KiB blocks: 3.53 -> 3.23 ms), 16 KiB blocks lose (2.70 -> 2.91 ms).
    const uint32_t lean_need = (LEAN && TSQB_DEC_EAGER && a.ostride <= 8192u) ? 4u * 133u + 4u * kLook : 4u * kLook;
    uint32_t pmask = kPairs - 1u, smask = kQueue / kPairs - 1u;   // per block: descriptors per step - 1, step barriers in use - 1
    enum { P_DONE = 0, P_WAIT = 1, P_WALK = 2 };
    const uint64_t stride_slots = (uint64_t)gridDim.x * nslots;
    const uint32_t slot = lane * kWalkers + widx;             // walker warp widx owns slots widx, widx + kWalkers, ...
    uint64_t b = (uint64_t)blockIdx.x * nslots + slot;
    uint32_t phase = (slot < nslots && b < a.nb) ? P_WAIT : P_DONE;
    SlotSmem<OUT_RING>& sm = slots[slot < nslots ? slot : 0];

    uint32_t p = 0, j = 0, size = 0, limit_al = 0, nchunks = 0, avail = 0, bits = 0;
    uint32_t ctl = 0;
    uint32_t k = 0, cons = 0;                      // descriptors of the current block written / known consumed
    uint32_t steps = 0;                            // running count of completed steps (16 descriptors, or up to the end)
    const uint32_t rbase = smem_u32(sm.in_ring), dbase = smem_u32(sm.desc);

    auto ring_u8 = [&](uint32_t pos) -> uint32_t {
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(rbase + (pos & kInMask)));
        return v;
    };
    auto put_desc = [&](uint32_t ad, uint32_t x, uint32_t y) {
        asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(ad), "r"(x), "r"(y) : "memory");
    };

    for (;;) {
#pragma unroll 1
        for (int rep = 0; rep < 8; rep++) {
            if (phase == P_WALK) {
                // take note of freshly landed chunks before deciding how far this lane may go
                // (the 4-group path below needs 4 * 133 + 4 * kLook landed bytes ahead of p: look for the next chunk that early)
#if TSQB_DEC_EAGER >= 2
                // (two chunks per iteration: one iteration of the 4-group path can consume more than a whole chunk of dense data, so
                // noting one chunk per iteration would leave the walk within a chunk of the frontier it knows of for good)
#pragma unroll
                for (int t = 0; t < 2; t++)
#endif
                if (avail != nchunks && p + (kEager ? 4u * 133u + 4u * kLook : lean_need) > avail * kChunk) {
                    const uint32_t s = avail % kChunks;
                    if (mbar_test(&sm.bar[s], (bits >> s) & 1u)) { bits ^= 1u << s; avail++; }
                }
                // the 3-byte header needs the first chunk (tsq_decode.cpp:49-53)
                if (size == 0xffffffffu) {
                    if (avail >= 1u || avail == nchunks) {
                        size = ring_u8(p) | (ring_u8(p + 1u) << 8) | (ring_u8(p + 2u) << 16);
                        if (!(size <= kBlockMax && size <= a.ostride) || nchunks == 0) size = 0;
                        p += 3u;
                    }
                    continue;
                }
                const uint32_t have = (avail == nchunks) ? 0xffffffffu : avail * kChunk;
                if ((k - cons) + (TSQB_DEC_BULK4 ? 16u : 4u) > kQueue) cons = ld_vol_u32(&sm.consumed);
                // ---- fast path: a whole group (control byte + 4 full pairs, tsq_decode.cpp:62-86) with no
                // end-of-block inside it.  k % 4 == 0 here, so the 4 descriptors are contiguous in the queue.
                const uint32_t p_safe = min(have, limit_al);
                int groups = 0;
#if TSQB_DEC_BULK4
                // Four groups at once when even the longest possible ones (133 stream bytes, 128 output bytes each) stay clear of
                // every limit -- and of the end of the stream ring, so that the walk can use a plain shared-memory pointer: the
                // per-group tests and the ring wrap (mask + base) drop out of the serial chain (load, 3 ALU, load).
                if ((k & 3u) == 0 && p + 4u * 133u + 4u * kLook <= p_safe && (k - cons) + 16u <= kQueue && j + 4u * 128u + (EXT ? 512u : 128u) < size && !EXT &&
                    (kWMask || (p & kInMask) + 4u * 133u + 8u <= kInRing)) {
                    // kWMask: ring addresses are masked (one LOP3 on the chain) so that a lane near the end of its ring stays on
                    // this path: lanes of one warp that sit on different paths execute them one after the other.  Else the walk
                    // uses a plain shared-memory pointer (stream position = shared address + dlt).
                    uint32_t pr = kWMask ? p : rbase + (p & kInMask);
                    const uint32_t dlt = kWMask ? 0u : p - pr;
                    auto lds_u8 = [&](uint32_t pos) -> uint32_t {
                        uint32_t v;
                        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(kWMask ? rbase + (pos & kInMask) : pos));
                        return v;
                    };
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        const uint32_t dsl = dbase + ((k & kQMask) << 3);
                        const uint32_t c = lds_u8(pr);
                        uint32_t prp = pr + 1u;
                        const uint32_t c24 = c << 24;
                        // Eight 16-byte literals (control byte 0xFF, four size bytes 0xFF: what incompressible data encodes to,
                        // tsq_encode.cpp:85-97) are a group of fixed layout -- size bytes 33 bytes apart -- so they are verified with
                        // four independent loads instead of the chain load -> lengths -> next load.
                        bool lit16 = false;
                        if (kLit16 && c == 0xFFu) {
                            const uint32_t a0 = lds_u8(prp), a1 = lds_u8(prp + 33u), a2 = lds_u8(prp + 66u), a3 = lds_u8(prp + 99u);
                            lit16 = (a0 & a1 & a2 & a3) == 0xFFu;
                        }
                        if (lit16) {
#pragma unroll
                            for (int q = 0; q < 4; q++) put_desc(dsl + 8u * q, (prp + 33u * q + dlt) | c24, (j + 32u * q) | 0x80000000u);   // bit 31: a verified (L16, L16) pair
                            prp += 132u;
                            j += 128u;
                        } else
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const uint32_t nib = lds_u8(prp);
#if TSQB_DEC_DESC2
                            put_desc(dsl + 8u * q, (prp + dlt) | c24, j);
#else
                            put_desc(dsl + 8u * q, (prp + dlt) | ((c << (18 + 2 * q)) & 0x3000000u), j);
#endif
                            const uint32_t n0 = nib >> 4, n1 = nib & 15u;
                            const uint32_t pay0 = (c & (0x80u >> (2 * q))) ? n0 + 2u : 3u;
                            const uint32_t pay1 = (c & (0x40u >> (2 * q))) ? n1 + 1u : 2u;
                            prp += pay0 + pay1;
                            j += n0 + n1 + 2u;
                        }
                        pr = prp;
                        k += 4u;
#if TSQB_DEC_DESC2
                        if ((k & pmask) == 0) { st_vol_u32(&sm.produced, k); mbar_arrive(&sm.full[steps & smask]); steps++; }
#else
                        st_vol_u32(&sm.produced, k);
                        if ((k & pmask) == 0) { mbar_arrive(&sm.full[steps & smask]); steps++; }
#endif
                    }
                    p = pr + dlt;
                    continue;
                }
#endif
#pragma unroll 1
                while (groups < 4 && (k & 3u) == 0 && p + 4u * kLook <= p_safe && (k - cons) + 4u <= kQueue && j + (EXT ? 512u : 128u) < size) {
                    const uint32_t dsl = dbase + ((k & kQMask) << 3);
                    const uint32_t c = ring_u8(p);                                  // :62
                    uint32_t pp = p + 1u;
                    bool lit16 = false;                                             // eight 16-byte literals: see the 4-group path
                    if (kLit16 && c == 0xFFu) lit16 = (ring_u8(pp) & ring_u8(pp + 33u) & ring_u8(pp + 66u) & ring_u8(pp + 99u)) == 0xFFu;
                    if (lit16) {
#pragma unroll
                        for (int q = 0; q < 4; q++) put_desc(dsl + 8u * q, (pp + 33u * q) | (c << 24), (j + 32u * q) | 0x80000000u);
                        pp += 132u;
                        j += 128u;
                    } else
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const uint32_t nib = ring_u8(pp);                           // :68
                        put_desc(dsl + 8u * q, pp | (TSQB_DEC_DESC2 ? (c << 24) : ((c << (18 + 2 * q)) & 0x3000000u)), j);
                        const uint32_t n0 = nib >> 4, n1 = nib & 15u;
                        const uint32_t pay0 = (c & (0x80u >> (2 * q))) ? n0 + 2u : 3u;   // payload + the size byte itself
                        const uint32_t pay1 = (c & (0x40u >> (2 * q))) ? n1 + 1u : 2u;
                        pp += pay0 + pay1;
                        if (EXT) j += sym_len<true>(n0, (c & (0x80u >> (2 * q))) != 0) + sym_len<true>(n1, (c & (0x40u >> (2 * q))) != 0);
                        else     j += n0 + n1 + 2u;
                    }
                    p = pp;
                    k += 4u;
                    st_vol_u32(&sm.produced, k);
                    if ((k & pmask) == 0) { mbar_arrive(&sm.full[steps & smask]); steps++; }
                    groups++;
                }
                if (groups) continue;
                // ---- slow path: one pair, or the end of the block
                if (!(p + kLook <= have)) continue;
                if (!((k - cons) < kQueue)) { cons = ld_vol_u32(&sm.consumed); continue; }
                if (j < size && p < limit_al) {
                    uint32_t pp = p;
                    const uint32_t q = k & 3u;
                    if (q == 0) { ctl = ring_u8(pp); pp++; }
                    const uint32_t nib = ring_u8(pp);
                    put_desc(dbase + ((k & kQMask) << 3), pp | (TSQB_DEC_DESC2 ? (ctl << 24) : ((ctl << (18u + 2u * q)) & 0x3000000u)), j);
                    const uint32_t n0 = nib >> 4, n1 = nib & 15u;
                    const uint32_t pay0 = (ctl & (0x80u >> (2u * q))) ? n0 + 1u : 2u;
                    const uint32_t pay1 = (ctl & (0x40u >> (2u * q))) ? n1 + 1u : 2u;
                    const uint32_t j1 = j + sym_len<EXT>(n0, (ctl & (0x80u >> (2u * q))) != 0);
                    const bool two = j1 < size;                                     // second symbol exists
                    p = pp + 1u + pay0 + (two ? pay1 : 0u);
                    j = j1 + (two ? sym_len<EXT>(n1, (ctl & (0x40u >> (2u * q))) != 0) : 0u);
                    k++;
                    st_vol_u32(&sm.produced, k);
                    if ((k & pmask) == 0) { mbar_arrive(&sm.full[steps & smask]); steps++; }
                } else {
                    st_vol_u32(&sm.produced, k);                                    // (the 4-group path publishes it once per step only)
                    st_vol_u32(&sm.end_j, j);
                    st_vol_u32(&sm.ended, 1u);
                    mbar_arrive(&sm.full[steps & smask]); steps++;                  // closes the last (possibly empty) step
                    b += stride_slots;
                    phase = b < a.nb ? P_WAIT : P_DONE;
                }
            } else if (phase == P_WAIT) {
                if (ld_vol_u32(&sm.ready) == (uint32_t)b + 1u) {
                    __threadfence_block();
                    bits     = ld_vol_u32(&sm.phase_bits);
                    limit_al = ld_vol_u32(&sm.limit_al);
                    nchunks  = ld_vol_u32(&sm.nchunks);
                    p        = ld_vol_u32(&sm.shift);
                    pmask    = ld_vol_u32(&sm.step_pairs) - 1u;
                    smask    = kQueue / (pmask + 1u) - 1u;
                    avail = 0; ctl = 0; j = 0; k = 0; cons = 0; size = 0xffffffffu;
                    phase = P_WALK;
                }
            }
        }
        if (!__any_sync(FULL, phase != P_DONE)) break;
    }
}

// ------------------------------------------------------------------------------------------ copier
// 16 bytes starting at byte `pos` of a power-of-two ring that begins at shared address `base`.
// `wrap` (warp-uniform) = some lane's 20-byte window crosses the end of its ring.
__device__ __forceinline__ void load16_smem(uint32_t base, uint32_t mask, uint32_t pos, uint32_t v[4], bool wrap)
{
    const uint32_t sh = (pos & 3u) * 8u;
    uint32_t w[5];
    if (!wrap) {
        const uint32_t ad = base + (pos & mask & ~3u);
#pragma unroll
        for (int m = 0; m < 5; m++) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[m]) : "r"(ad + 4u * m));
    } else {
#pragma unroll
        for (int m = 0; m < 5; m++) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[m]) : "r"(base + (((pos & ~3u) + 4u * m) & mask)));
    }
#pragma unroll
    for (int m = 0; m < 4; m++) v[m] = __funnelshift_r(w[m], w[m + 1], sh);
}

// The (up to) 16 bytes of a far match source -- bytes of this block that an earlier step flushed to HBM -- as the five
// 32-bit words around them (the caller funnel-shifts by the address's low two bits, as for the narrow loads).
// One aligned 128-bit load covers the source unless it crosses the next 16-byte boundary (then a second one): a scattered
// warp-wide load costs one L1TEX wavefront per lane and instruction, so fewer, wider instructions are what counts.
// Never touches a 16-byte unit the source does not reach.
// far_issue only issues the loads; far_words turns them into w[] -- called where the bytes are needed, so that the DRAM / L2
// latency of a step's far sources runs under its shared-memory loads.
__device__ __forceinline__ void far_issue(const uint8_t* src, uint32_t len, uint4& A, uint4& B)
{
    const uintptr_t ad = reinterpret_cast<uintptr_t>(src);
    const uint4* g = reinterpret_cast<const uint4*>(ad & ~(uintptr_t)15);
    A = ld_far16(g);
    B = make_uint4(0, 0, 0, 0);
    if ((uint32_t)(ad & 15u) + len > 16u) B = ld_far16(g + 1);
}

__device__ __forceinline__ void far_words(const uint8_t* src, const uint4& A, const uint4& B, uint32_t w[5])
{
    // words X[0..7] = A.x A.y A.z A.w B.x B.y B.z B.w; wanted: X[(o >> 2) + m], m = 0..4 -- two levels of selects
    const uint32_t o = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
    const bool s2 = (o & 8u) != 0, s1 = (o & 4u) != 0;
    const uint32_t y0 = s2 ? A.z : A.x, y1 = s2 ? A.w : A.y, y2 = s2 ? B.x : A.z, y3 = s2 ? B.y : A.w, y4 = s2 ? B.z : B.x, y5 = s2 ? B.w : B.y;
    w[0] = s1 ? y1 : y0; w[1] = s1 ? y2 : y1; w[2] = s1 ? y3 : y2; w[3] = s1 ? y4 : y3; w[4] = s1 ? y5 : y4;
}

__device__ __forceinline__ void far_load_wide(const uint8_t* src, uint32_t len, uint32_t w[5])
{
    uint4 A, B;
    far_issue(src, len, A, B);
    far_words(src, A, B, w);
}

// The bytes of one symbol into a ring: byte t is stored iff bit t of `lenmask` is set (ptxas materialises the mask
// as predicates with two R2P, so exact lengths cost no more than a blind 16-byte run).  `wrap`: the run crosses the
// end of the ring (per-lane; the fast path uses immediate offsets).
template <int T>
__device__ __forceinline__ void store_bytes(uint32_t ad, const uint32_t v[4], uint32_t lenmask)
{
    if (lenmask & (1u << T))
        asm volatile("st.shared.u8 [%0+%1], %2;" ::"r"(ad), "n"(T), "r"(v[T >> 2] >> (8 * (T & 3))) : "memory");
    if constexpr (T > 0) store_bytes<T - 1>(ad, v, lenmask);
}

__device__ __forceinline__ void store16(uint32_t base, uint32_t mask, uint32_t q, const uint32_t v[4], bool wrap, uint32_t lenmask)
{
    if (!wrap) {
        store_bytes<15>(base + (q & mask), v, lenmask);
    } else {
#pragma unroll
        for (int t = 15; t >= 0; t--)
            if (lenmask & (1u << t))
                asm volatile("st.shared.u8 [%0], %1;" ::"r"(base + ((q + (uint32_t)t) & mask)), "r"(v[t >> 2] >> (8 * (t & 3))) : "memory");
    }
}

// Barrier bookkeeping of one block slot, carried from block to block (and between the two copiers)
struct CopierState {
    uint32_t phase = 0;                    // bit s: parity the next completion of stream ring slot s will have
    uint32_t fphase = 0;                   // same for the step barriers
    uint32_t sc = 0;                       // running step counter (mirrors the walker's `steps`)
};

// Decodes blocks b0, b0 + stride_slots, ... (ONE: only b0) of slot `sm`.
template <uint32_t OUT_RING, bool EXT, bool ONE>
__device__ void copier(const DecodeArgs& a, SlotSmem<OUT_RING>& sm, uint64_t b0, uint64_t stride_slots, unsigned lane, CopierState& cs)
{
    constexpr uint32_t kOMask = OUT_RING - 1;
    uint32_t& phase = cs.phase;
    uint32_t& fphase = cs.fphase;
    uint32_t& sc = cs.sc;
    uint8_t* oring = sm.out_ring;
    const uint8_t* iring = sm.in_ring;
    const uint32_t ibase = smem_u32(sm.in_ring), obase = smem_u32(sm.out_ring);
    const uint32_t bar_base = smem_u32(sm.bar), full_base = smem_u32(sm.full);
    const uint32_t pi = lane >> 1, half = lane & 1u;

    for (uint64_t b = b0; b < a.nb; b += stride_slots) {
        uint32_t limit;
        const uint8_t* src = stream_of(a, b, limit);
        const uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
        const uint8_t* src_al = src - shift;
        const uint32_t limit_al = shift + limit;
        const uint32_t total_al = (limit_al + 15u) & ~15u;
        const uint32_t nchunks = (total_al + kChunk - 1) / kChunk;
        uint32_t issued = 0, waited = 0, cur_chunk = 0;

        auto issue_upto = [&](uint32_t want) {                                   // all lanes call; lane 0 copies
            want = min(want, nchunks);
            if (lane == 0)
                for (uint32_t n = issued; n < want; n++) {
                    const uint32_t at = n * kChunk, bytes = min(kChunk, total_al - at);
                    mbar_expect_tx(&sm.bar[n % kChunks], bytes);
                    bulk_load(sm.in_ring + (at & kInMask), src_al + at, bytes, &sm.bar[n % kChunks]);
                }
            issued = max(issued, want);
        };
        auto wait_upto = [&](uint32_t want) {                                    // all lanes wait
            want = min(want, issued);
            for (; waited < want; waited++) {
                const uint32_t s = waited % kChunks;
                mbar_wait_addr(bar_base + 8u * s, (phase >> s) & 1u);
                phase ^= 1u << s;
            }
        };

        // ---- hand the block to the walker
        uint32_t kc = 0;                                                         // descriptors of this block consumed
        if (lane == 0) {
            st_vol_u32(&sm.produced, 0u);
            st_vol_u32(&sm.consumed, 0u);
            st_vol_u32(&sm.ended, 0u);
            st_vol_u32(&sm.phase_bits, phase);
            st_vol_u32(&sm.limit_al, limit_al);
            st_vol_u32(&sm.nchunks, nchunks);
            st_vol_u32(&sm.shift, shift);
            st_vol_u32(&sm.step_pairs, kPairs);
        }
        issue_upto(kChunks);
        __syncwarp();
        if (lane == 0) { __threadfence_block(); st_vol_u32(&sm.ready, (uint32_t)b + 1u); }
        wait_upto(1);

        auto rb = [&](uint32_t k) -> uint32_t { return iring[k & kInMask]; };    // k counts from the aligned start
        uint32_t size = rb(shift) | (rb(shift + 1u) << 8) | (rb(shift + 2u) << 16);   // tsq_decode.cpp:49-51
        const bool ok = size <= kBlockMax && size <= a.ostride && nchunks != 0;
        if (!ok) size = 0;
        if (lane == 0) a.osizes[b] = size;

        // output positions are kept in "q" coordinates: q = position + (address of the block's
        // output & 15), so that 16-byte units of the ring are 16-byte units of HBM
        uint8_t* o = a.out + b * a.ostride;
        const uint32_t oal = (uint32_t)(reinterpret_cast<uintptr_t>(o) & 15u);
        uint8_t* o_al = o - oal;
        uint32_t F = oal;                                                        // flushed up to here (q)

        // complete 16-byte units of the ring -> HBM, one 128-bit load/store per lane
        auto flush_units = [&](uint32_t E) {                                     // E 16-byte aligned, F 16-byte aligned
            // a step leaves at most 33 units behind: two predicated 128-bit moves per lane cover it
#pragma unroll
            for (int rep = 0; rep < 2; rep++) {
                const uint32_t at = F + 16u * lane + 512u * rep;
                if (at < E) {
                    uint4 x;
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "r"(obase + (at & kOMask)));
                    st_out16(o_al + at, x);
                }
            }
            for (uint32_t at = F + 16u * lane + 1024u; at < E; at += 512u) {     // never taken for regular steps
                uint4 x;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "r"(obase + (at & kOMask)));
                st_out16(o_al + at, x);
            }
            F = E;
        };
        auto flush = [&](uint32_t E, bool final) {                               // E in q coordinates
            if (!final) E &= ~15u;
            if (F >= E) return;
            if (F & 15u) {                                                       // unaligned head of the block
                const uint32_t h = min(E, (F + 15u) & ~15u);
                if (lane < h - F) o_al[F + lane] = oring[(F + lane) & kOMask];
                F = h;
            }
            flush_units(E & ~15u);
            if (F < E) {                                                         // final tail
                if (lane < E - F) o_al[F + lane] = oring[(F + lane) & kOMask];
                F = E;
            }
        };

        bool done = false;
        while (!done) {
            // ---- sleep on the step barrier until the walker has completed 16 descriptors or reached END
            {
                const uint32_t s = sc % kSteps;
                mbar_wait_addr(full_base + 8u * s, (fphase >> s) & 1u);
                fphase ^= 1u << s;
                sc++;
            }
            // a step is 16 descriptors, except the last one of a block (fewer, possibly none)
            const uint32_t np = min(ld_vol_u32(&sm.produced) - kc, kPairs);
            done = np < kPairs;
            const uint2 d = ld_vol_u64(&sm.desc[(kc + pi) & kQMask]);
            if (np) {
                // stream bytes of these pairs are resident (the walker saw them); observe the barriers
                const uint32_t plast = __shfl_sync(FULL, d.x & 0xFFFFFFu, (np - 1u) * 2u);
                if ((plast + kLook - 2u) / kChunk >= waited) wait_upto((plast + kLook - 2u) / kChunk + 1u);

#if TSQB_DEC_LIT16 && TSQB_DEC_DESC2
                if constexpr (!EXT) {
                    // ---- a step of sixteen pairs the walker has verified as two 16-byte literals each (incompressible data):
                    // symbol L is 16 bytes at stream position pp + 1 + 16 * (L & 1) and output byte 16 * L of the step: no lengths,
                    // no classification.  The 512 bytes go to the output ring (a later near match may read them) as aligned words:
                    // lane L owns words 4L .. 4L + 3 of the step's 4-byte aligned image, realigned from its own symbol and the last
                    // word of lane L - 1 by the step's (warp-uniform) byte offset -- an accidental match early in a block shifts every
                    // later literal off its 16-byte unit.  When the step starts on a 16-byte unit with everything before it flushed
                    // the bytes also go straight to HBM; otherwise the usual ring -> HBM flush follows.
                    const uint32_t Jd = (__shfl_sync(FULL, d.y, 0) & 0x7FFFFFFFu) + oal;
                    if (np == kPairs && __all_sync(FULL, (d.y >> 31) != 0u)) {
                        const uint32_t spd = (d.x & 0xFFFFFFu) + 1u + 16u * half;
                        uint32_t vd[4];
                        const bool wrapd = __any_sync(FULL, (spd & kInMask) + 20u > kInRing);
                        load16_smem(ibase, kInMask, spd, vd, wrapd);
                        const uint32_t sb = Jd & 3u;                              // warp-uniform
                        if ((Jd & 15u) == 0u) {
                            const uint32_t qd = Jd + 16u * lane;
                            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(obase + (qd & kOMask)), "r"(vd[0]), "r"(vd[1]), "r"(vd[2]), "r"(vd[3]) : "memory");
                            if (F == Jd) { st_out16(o_al + qd, make_uint4(vd[0], vd[1], vd[2], vd[3])); F = Jd + 512u; }
                        } else {
                            const uint32_t prev3 = __shfl_up_sync(FULL, vd[3], 1);
                            const uint32_t s8 = sb * 8u;
                            uint32_t wd[4];
                            wd[0] = __funnelshift_l(prev3, vd[0], s8);            // (sb == 0: the words as they are)
                            wd[1] = __funnelshift_l(vd[0], vd[1], s8);
                            wd[2] = __funnelshift_l(vd[1], vd[2], s8);
                            wd[3] = __funnelshift_l(vd[2], vd[3], s8);
                            const uint32_t w0 = Jd - sb + 16u * lane;             // q of this lane's first aligned word
                            if (lane == 0 && sb) {
                                // the first word also holds sb bytes of the symbol before this step: store only this step's bytes
                                for (uint32_t t = sb; t < 4u; t++)
                                    asm volatile("st.shared.u8 [%0], %1;" ::"r"(obase + ((w0 + t) & kOMask)), "r"(vd[0] >> (8u * (t - sb))) : "memory");
                            } else
                                asm volatile("st.shared.u32 [%0], %1;" ::"r"(obase + (w0 & kOMask)), "r"(wd[0]) : "memory");
                            asm volatile("st.shared.u32 [%0], %1;" ::"r"(obase + ((w0 + 4u) & kOMask)), "r"(wd[1]) : "memory");
                            asm volatile("st.shared.u32 [%0], %1;" ::"r"(obase + ((w0 + 8u) & kOMask)), "r"(wd[2]) : "memory");
                            asm volatile("st.shared.u32 [%0], %1;" ::"r"(obase + ((w0 + 12u) & kOMask)), "r"(wd[3]) : "memory");
                            if (lane == 31 && sb) {
                                // ... and the step's last sb bytes sit in a word of their own
                                for (uint32_t t = 0; t < sb; t++)
                                    asm volatile("st.shared.u8 [%0], %1;" ::"r"(obase + ((w0 + 16u + t) & kOMask)), "r"(vd[3] >> (8u * (4u - sb + t))) : "memory");
                            }
                        }
                        __syncwarp();
                        if (F != Jd + 512u) {                                     // not stored directly: the usual flush of complete units
                            const uint32_t J1d = Jd + 512u;
                            if (F & 15u) flush(J1d, false); else if ((J1d & ~15u) > F) flush_units(J1d & ~15u);
                        }
                        kc += np;
                        const uint32_t c0d = __shfl_sync(FULL, d.x & 0xFFFFFFu, 0) / kChunk;
                        if (c0d != cur_chunk) { cur_chunk = c0d; issue_upto(c0d + kChunks); }
                        if (lane == 0) st_vol_u32(&sm.consumed, kc);
                        continue;                                                 // a full step: the block goes on
                    }
                }
#endif
                // ---- one lane per symbol
                bool active = pi < np;
                uint32_t len = 0, q = 0, sp = 0, srcq = 0;
                bool lit = true;
                const uint32_t pp = d.x & 0xFFFFFFu, jp = d.y & 0x7FFFFFFFu;          // (bit 31 of d.y: the walker's (L16, L16) flag)
                {
                    const uint32_t nib = rb(pp);
                    const uint32_t n0 = nib >> 4, n1 = nib & 15u;
#if TSQB_DEC_DESC2
                    const uint32_t cbits = d.x >> (30u - 2u * (pi & 3u));     // a step starts at a multiple of 16 descriptors: pair pi is pair pi & 3 of its group
                    const bool l0 = (cbits & 2u) != 0, l1 = (cbits & 1u) != 0;
#else
                    const bool l0 = (d.x & (0x80u << 18)) != 0, l1 = (d.x & (0x40u << 18)) != 0;
#endif
                    uint32_t dst;
                    if (half == 0) { lit = l0; len = sym_len<EXT>(n0, l0); sp = pp + 1u; dst = jp; }
                    else { lit = l1; len = sym_len<EXT>(n1, l1); sp = pp + 1u + (l0 ? n0 + 1u : 2u); dst = jp + sym_len<EXT>(n0, l0); }
                    active = active && dst < size;
                    len = active ? min(len, size - dst) : 0u;
                    q = dst + oal;
                    if (active && !lit) {
                        const uint32_t off = rb(sp) | (rb(sp + 1u) << 8);        // :69,73,82
                        active = off <= jp;                                      // corrupt stream guard
                        srcq = jp - off + oal;
                    }
                }
                const uint32_t J0 = __shfl_sync(FULL, q, 0);                     // q of the step's first byte
                const uint32_t J1 = __reduce_max_sync(FULL, active ? q + len : 0u);

                // ---- round 0: literals and matches whose source precedes this step's output.
                // Literals and near matches both come out of shared memory (stream ring / output ring);
                // far matches read bytes an earlier step flushed to HBM.
                uint32_t v[4] = {0, 0, 0, 0};
                // extension format: a match of 32 / 48 / 64 bytes does not fit a lane's 16-byte run; it takes the
                // in-order lane-per-byte path below, from the output ring or -- when older than the ring -- from HBM
                const bool longsym = EXT && len > 16u;
                bool now = active && !longsym && (lit || srcq + len <= J0);
                bool pending = active && !now;
                bool placed = false;                                             // v[] holds this symbol's bytes
                const bool far = now && !lit && srcq + OUT_RING < J1 + 16u;
                {
                    const uint32_t fbase = lit ? ibase : obase, fmask = lit ? kInMask : kOMask, fpos = lit ? sp : srcq;
                    const bool sm_src = now && !far;
                    const bool wrap = __any_sync(FULL, sm_src && ((fpos & fmask) + 20u > fmask + 1u));
                    if (sm_src) load16_smem(fbase, fmask, fpos, v, wrap);
                    if (far) {
                        const uintptr_t ad = reinterpret_cast<uintptr_t>(o_al + srcq);
                        const uint32_t sh = (uint32_t)(ad & 3u) * 8u;
                        uint32_t w[5];
#if TSQB_DEC_FARWIDE
                        far_load_wide(o_al + srcq, len, w);
#else
                        const uint32_t* g32 = reinterpret_cast<const uint32_t*>(ad & ~(uintptr_t)3);
#pragma unroll
                        for (int m = 0; m < 5; m++)                              // never touch a word past the source
                            w[m] = ((uint32_t)(ad & 3u) + len > 4u * m) ? __ldcg(g32 + m) : 0u;
#endif
#pragma unroll
                        for (int m = 0; m < 4; m++) v[m] = __funnelshift_r(w[m], w[m + 1], sh);
                    }
                    placed = now;
                }
#if TSQB_DEC_LIT16
                // every symbol of the step a full, 16-byte aligned run (a block of 16-byte literals): one 128-bit store each
                if (__all_sync(FULL, !placed || (len == 16u && (q & 15u) == 0u))) {
                    if (placed) asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(obase + (q & kOMask)), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
                } else
#endif
                if (placed) store16(obase, kOMask, q, v, (q & kOMask) + 16u > OUT_RING, (1u << min(len, 16u)) - 1u);
                // ---- symbols whose source lies inside this step's output: in position order, one at a time, the
                // warp copying a symbol's bytes lane-per-byte.  Sources always precede their own pair (tsq_encode.cpp:139-141), so by
                // the time a symbol is reached everything it reads is in place.
                uint32_t pm = __ballot_sync(FULL, pending);
                while (pm) {
                    __syncwarp();                                                // stores so far are visible to the warp
                    const uint32_t pl = (uint32_t)__ffs((int)pm) - 1u;
                    const uint32_t s_q = __shfl_sync(FULL, q, pl), s_src = __shfl_sync(FULL, srcq, pl), s_len = __shfl_sync(FULL, len, pl);
                    const bool s_far = EXT && s_src + OUT_RING < J1 + 16u;        // only long matches can be far here
                    if constexpr (EXT) {
                        for (uint32_t t = lane; t < s_len; t += 32u) {
                            uint32_t byte;
                            if (s_far) byte = __ldcg(o_al + s_src + t);
                            else asm volatile("ld.volatile.shared.u8 %0, [%1];" : "=r"(byte) : "r"(obase + ((s_src + t) & kOMask)) : "memory");
                            asm volatile("st.volatile.shared.u8 [%0], %1;" ::"r"(obase + ((s_q + t) & kOMask)), "r"(byte) : "memory");
                        }
                    } else if (lane < s_len) {
                        uint32_t byte;
                        asm volatile("ld.volatile.shared.u8 %0, [%1];" : "=r"(byte) : "r"(obase + ((s_src + lane) & kOMask)) : "memory");
                        asm volatile("st.volatile.shared.u8 [%0], %1;" ::"r"(obase + ((s_q + lane) & kOMask)), "r"(byte) : "memory");
                    }
                    pm &= pm - 1u;
                }
                __syncwarp();
                if (F & 15u) flush(J1, false); else if ((J1 & ~15u) > F) flush_units(J1 & ~15u);
                kc += np;
                // recycle the stream ring behind this step
                const uint32_t c0 = __shfl_sync(FULL, pp, 0) / kChunk;
                if (c0 != cur_chunk) { cur_chunk = c0; issue_upto(c0 + kChunks); }
            }
            if (done) {
                __syncwarp();
                flush(min(ld_vol_u32(&sm.end_j), size) + oal, true);
            }
            if (lane == 0) st_vol_u32(&sm.consumed, kc);
        }
        wait_upto(issued);                                                       // drain before the ring is reused
        __syncwarp();
        if (ONE) break;
    }
}

// ------------------------------------------------------------------------------------------ copier, lane per PAIR
// Same protocol as copier(), but a step is 32 descriptors = 64 symbols and lane L owns BOTH symbols of pair L (no-extension
// format only: a step then produces at most 1 KiB, half of the smallest output ring).  Per symbol the work is the same;
// what is paid once per step instead of twice -- barrier wait, descriptor and size-byte loads, the J0 / J1 reductions, chunk
// bookkeeping, flush -- is about a third of copier()'s instructions, and the far-match loads of 64 symbols are in flight
// together, so a block meets half as many DRAM-latency stalls.
template <uint32_t OUT_RING>
__device__ void copier_pairs(const DecodeArgs& a, SlotSmem<OUT_RING>& sm, uint64_t b, unsigned lane, CopierState& cs)
{
    constexpr uint32_t kOMask = OUT_RING - 1;
    constexpr uint32_t kPairs2 = 32, kSteps2 = kQueue / kPairs2;
    static_assert(OUT_RING >= 2048, "a 64-symbol step writes up to 1 KiB");
    uint32_t& phase = cs.phase;
    uint32_t& fphase = cs.fphase;
    uint32_t& sc = cs.sc;
    uint8_t* oring = sm.out_ring;
    const uint8_t* iring = sm.in_ring;
    const uint32_t ibase = smem_u32(sm.in_ring), obase = smem_u32(sm.out_ring);
    const uint32_t bar_base = smem_u32(sm.bar), full_base = smem_u32(sm.full);

    {
        uint32_t limit;
        const uint8_t* src = stream_of(a, b, limit);
        const uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
        const uint8_t* src_al = src - shift;
        const uint32_t limit_al = shift + limit;
        const uint32_t total_al = (limit_al + 15u) & ~15u;
        const uint32_t nchunks = (total_al + kChunk - 1) / kChunk;
        uint32_t issued = 0, waited = 0, cur_chunk = 0;

        auto issue_upto = [&](uint32_t want) {
            want = min(want, nchunks);
            if (lane == 0)
                for (uint32_t n = issued; n < want; n++) {
                    const uint32_t at = n * kChunk, bytes = min(kChunk, total_al - at);
                    mbar_expect_tx(&sm.bar[n % kChunks], bytes);
                    bulk_load(sm.in_ring + (at & kInMask), src_al + at, bytes, &sm.bar[n % kChunks]);
                }
            issued = max(issued, want);
        };
        auto wait_upto = [&](uint32_t want) {
            want = min(want, issued);
            for (; waited < want; waited++) {
                const uint32_t s = waited % kChunks;
                mbar_wait_addr(bar_base + 8u * s, (phase >> s) & 1u);
                phase ^= 1u << s;
            }
        };

        uint32_t kc = 0;
        if (lane == 0) {
            st_vol_u32(&sm.produced, 0u);
            st_vol_u32(&sm.consumed, 0u);
            st_vol_u32(&sm.ended, 0u);
            st_vol_u32(&sm.phase_bits, phase);
            st_vol_u32(&sm.limit_al, limit_al);
            st_vol_u32(&sm.nchunks, nchunks);
            st_vol_u32(&sm.shift, shift);
            st_vol_u32(&sm.step_pairs, kPairs2);
        }
        issue_upto(kChunks);
        __syncwarp();
        if (lane == 0) { __threadfence_block(); st_vol_u32(&sm.ready, (uint32_t)b + 1u); }
        wait_upto(1);

        auto rb = [&](uint32_t k) -> uint32_t { return iring[k & kInMask]; };
        uint32_t size = rb(shift) | (rb(shift + 1u) << 8) | (rb(shift + 2u) << 16);   // tsq_decode.cpp:49-51
        const bool ok = size <= kBlockMax && size <= a.ostride && nchunks != 0;
        if (!ok) size = 0;
        if (lane == 0) a.osizes[b] = size;

        uint8_t* o = a.out + b * a.ostride;
        const uint32_t oal = (uint32_t)(reinterpret_cast<uintptr_t>(o) & 15u);
        uint8_t* o_al = o - oal;
        uint32_t F = oal;

        auto flush_units = [&](uint32_t E) {
#if TSQB_DEC_FLUSH2
            // a step leaves at most 65 complete units behind; a text step ~20: two predicated moves per lane, a third when full
#pragma unroll
            for (int rep = 0; rep < 2; rep++) {
                const uint32_t at = F + 16u * lane + 512u * rep;
                if (at < E) {
                    uint4 x;
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "r"(obase + (at & kOMask)));
                    st_out16(o_al + at, x);
                }
            }
            if (F + 1024u < E)
#endif
            for (uint32_t at = F + 16u * lane + (TSQB_DEC_FLUSH2 ? 1024u : 0u); at < E; at += 512u) {
                uint4 x;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "r"(obase + (at & kOMask)));
                st_out16(o_al + at, x);
            }
            F = E;
        };
        auto flush = [&](uint32_t E, bool final) {
            if (!final) E &= ~15u;
            if (F >= E) return;
            if (F & 15u) {
                const uint32_t h = min(E, (F + 15u) & ~15u);
                if (lane < h - F) o_al[F + lane] = oring[(F + lane) & kOMask];
                F = h;
            }
            flush_units(E & ~15u);
            if (F < E) {
                if (lane < E - F) o_al[F + lane] = oring[(F + lane) & kOMask];
                F = E;
            }
        };
        // raw words of a far source (bytes an earlier step flushed to HBM); never touches a word past the source
        auto far_load = [&](uint32_t srcq, uint32_t len, uint32_t w[5]) {
            const uintptr_t ad = reinterpret_cast<uintptr_t>(o_al + srcq);
            const uint32_t* g32 = reinterpret_cast<const uint32_t*>(ad & ~(uintptr_t)3);
#if TSQB_DEC_FARNP
            // A far source ends more than OUT_RING - 1040 >= 1008 bytes before this step's first output byte (srcq +
            // OUT_RING < J1 + 16 and J1 <= J0 + 1024): all five words lie inside bytes of THIS block that earlier steps
            // flushed, so they can be read without per-word predicates (only `len` of the 16 bytes are used)
            (void)len;
#pragma unroll
            for (int m = 0; m < 5; m++) w[m] = __ldcg(g32 + m);
#else
#pragma unroll
            for (int m = 0; m < 5; m++) w[m] = ((uint32_t)(ad & 3u) + len > 4u * m) ? __ldcg(g32 + m) : 0u;
#endif
        };
        // one pending symbol (source inside this step's output), lane-per-byte, after everything before it is in place
        auto copy_in_order = [&](uint32_t s_q, uint32_t s_src, uint32_t s_len) {
            __syncwarp();
            if (lane < s_len) {
                uint32_t byte;
                asm volatile("ld.volatile.shared.u8 %0, [%1];" : "=r"(byte) : "r"(obase + ((s_src + lane) & kOMask)) : "memory");
                asm volatile("st.volatile.shared.u8 [%0], %1;" ::"r"(obase + ((s_q + lane) & kOMask)), "r"(byte) : "memory");
            }
        };

        bool done = false;
        while (!done) {
            {
                const uint32_t s = sc % kSteps2;
                mbar_wait_addr(full_base + 8u * s, (fphase >> s) & 1u);
                fphase ^= 1u << s;
                sc++;
            }
            const uint32_t np = min(ld_vol_u32(&sm.produced) - kc, kPairs2);
            done = np < kPairs2;
            const uint2 d = ld_vol_u64(&sm.desc[(kc + lane) & kQMask]);
            if (np) {
                const uint32_t plast = __shfl_sync(FULL, d.x & 0xFFFFFFu, np - 1u);
                if ((plast + kLook - 2u) / kChunk >= waited) wait_upto((plast + kLook - 2u) / kChunk + 1u);

                // ---- lane L: both symbols of pair L (tsq_decode.cpp:68-86)
                const bool active = lane < np;
                const uint32_t pp = d.x & 0xFFFFFFu, jp = d.y & 0x7FFFFFFFu;          // (bit 31 of d.y: the walker's (L16, L16) flag)
                const uint32_t nib = rb(pp);
#if TSQB_DEC_DESC2
                const uint32_t cbits = d.x >> (30u - 2u * (lane & 3u));       // a step starts at a multiple of 32 descriptors: pair L is pair L & 3 of its group
                const bool l0 = (cbits & 2u) != 0, l1 = (cbits & 1u) != 0;
#else
                const bool l0 = (d.x & (0x80u << 18)) != 0, l1 = (d.x & (0x40u << 18)) != 0;
#endif
                uint32_t len0 = (nib >> 4) + 1u, len1 = (nib & 15u) + 1u;
                const uint32_t sp0 = pp + 1u, sp1 = sp0 + (l0 ? len0 : 2u);
                const uint32_t dst0 = jp, dst1 = jp + len0;
                bool act0 = active && dst0 < size, act1 = active && dst1 < size;
                len0 = act0 ? min(len0, size - dst0) : 0u;
                len1 = act1 ? min(len1, size - dst1) : 0u;
                const uint32_t q0 = dst0 + oal, q1 = dst1 + oal;
                uint32_t src0 = 0, src1 = 0;
                if (act0 && !l0) {
                    const uint32_t off = rb(sp0) | (rb(sp0 + 1u) << 8);          // :69,73
                    act0 = off <= jp;                                            // corrupt stream guard
                    src0 = jp - off + oal;
                }
                if (act1 && !l1) {
                    const uint32_t off = rb(sp1) | (rb(sp1 + 1u) << 8);          // :82: the same pair start
                    act1 = off <= jp;
                    src1 = jp - off + oal;
                }
                const uint32_t J0 = __shfl_sync(FULL, q0, 0);
                const uint32_t J1 = __reduce_max_sync(FULL, max(act0 ? q0 + len0 : 0u, act1 ? q1 + len1 : 0u));

                // ---- round 0: literals and matches whose source precedes this step's output
                const bool now0 = act0 && (l0 || src0 + len0 <= J0), now1 = act1 && (l1 || src1 + len1 <= J0);
                const bool pend0 = act0 && !now0, pend1 = act1 && !now1;
                const bool far0 = now0 && !l0 && src0 + OUT_RING < J1 + 16u, far1 = now1 && !l1 && src1 + OUT_RING < J1 + 16u;
#if TSQB_DEC_DIAG != 1
                uint32_t w0[5], w1[5];
#if TSQB_DEC_FARWIDE
                uint4 fa0, fb0, fa1, fb1;
                if (far0) far_issue(o_al + src0, len0, fa0, fb0);                // both symbols' far loads fly together,
                if (far1) far_issue(o_al + src1, len1, fa1, fb1);                // under the shared-memory loads below
#else
                if (far0) far_load(src0, len0, w0);                              // both symbols' far loads fly together
                if (far1) far_load(src1, len1, w1);
#endif
                uint32_t v0[4] = {0, 0, 0, 0}, v1[4] = {0, 0, 0, 0};
                {
                    const uint32_t fbase = l0 ? ibase : obase, fmask = l0 ? kInMask : kOMask, fpos = l0 ? sp0 : src0;
                    const bool sm_src = now0 && !far0;
                    const bool wrap = __any_sync(FULL, sm_src && ((fpos & fmask) + 20u > fmask + 1u));
                    if (sm_src) load16_smem(fbase, fmask, fpos, v0, wrap);
                }
                {
                    const uint32_t fbase = l1 ? ibase : obase, fmask = l1 ? kInMask : kOMask, fpos = l1 ? sp1 : src1;
                    const bool sm_src = now1 && !far1;
                    const bool wrap = __any_sync(FULL, sm_src && ((fpos & fmask) + 20u > fmask + 1u));
                    if (sm_src) load16_smem(fbase, fmask, fpos, v1, wrap);
                }
                if (far0) {
#if TSQB_DEC_FARWIDE
                    far_words(o_al + src0, fa0, fb0, w0);
#endif
                    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(o_al + src0) & 3u) * 8u;
#pragma unroll
                    for (int m = 0; m < 4; m++) v0[m] = __funnelshift_r(w0[m], w0[m + 1], sh);
                }
                if (far1) {
#if TSQB_DEC_FARWIDE
                    far_words(o_al + src1, fa1, fb1, w1);
#endif
                    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(o_al + src1) & 3u) * 8u;
#pragma unroll
                    for (int m = 0; m < 4; m++) v1[m] = __funnelshift_r(w1[m], w1[m + 1], sh);
                }
#if TSQB_DEC_LIT16 && TSQB_DEC_DENSE_PAIRS
                // every symbol of the step a full, 16-byte aligned run (16-byte literals): one 128-bit store each
                if (__all_sync(FULL, (!now0 || (len0 == 16u && (q0 & 15u) == 0u)) && (!now1 || (len1 == 16u && (q1 & 15u) == 0u)))) {
                    if (now0) asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(obase + (q0 & kOMask)), "r"(v0[0]), "r"(v0[1]), "r"(v0[2]), "r"(v0[3]) : "memory");
                    if (now1) asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(obase + (q1 & kOMask)), "r"(v1[0]), "r"(v1[1]), "r"(v1[2]), "r"(v1[3]) : "memory");
                } else
#endif
                {
                if (now0) store16(obase, kOMask, q0, v0, (q0 & kOMask) + 16u > OUT_RING, (1u << len0) - 1u);
                if (now1) store16(obase, kOMask, q1, v1, (q1 & kOMask) + 16u > OUT_RING, (1u << len1) - 1u);
                }

                // ---- symbols whose source lies inside this step's output: in position order, pair by pair; sources always
                // precede their own pair (tsq_encode.cpp:139-141), so the two symbols of a pair never depend on each other
#if TSQB_DEC_PEND2
                {
                    // one packed word per symbol: q - J0 (10 bits) | src + 16 - J0 (11 bits: a pending source starts at most
                    // 15 bytes before the step) | len - 1 (4 bits).  Lanes 0..15 copy the pair's first symbol byte per lane,
                    // lanes 16..31 its second one, at the same time.
                    const uint32_t pk0 = pend0 ? ((q0 - J0) | ((src0 + 16u - J0) << 10) | ((len0 - 1u) << 21)) : 0xFFFFFFFFu;
                    const uint32_t pk1 = pend1 ? ((q1 - J0) | ((src1 + 16u - J0) << 10) | ((len1 - 1u) << 21)) : 0xFFFFFFFFu;
                    uint32_t pm = __ballot_sync(FULL, pend0 || pend1);
                    const uint32_t hl = lane & 15u, tb = J0 + hl;
                    const bool second = lane >= 16u;
                    while (pm) {
                        const uint32_t pl = (uint32_t)__ffs((int)pm) - 1u;
                        pm &= pm - 1u;
                        const uint32_t a = __shfl_sync(FULL, pk0, pl), b = __shfl_sync(FULL, pk1, pl);
                        const uint32_t pk = second ? b : a;
                        __syncwarp();                                            // everything before this pair is in place
                        if (pk != 0xFFFFFFFFu && hl <= (pk >> 21)) {
                            uint32_t byte;
                            asm volatile("ld.volatile.shared.u8 %0, [%1];" : "=r"(byte) : "r"(obase + ((tb + ((pk >> 10) & 0x7FFu) - 16u) & kOMask)) : "memory");
                            asm volatile("st.volatile.shared.u8 [%0], %1;" ::"r"(obase + ((tb + (pk & 0x3FFu)) & kOMask)), "r"(byte) : "memory");
                        }
                    }
                }
#else
                uint32_t pm0 = __ballot_sync(FULL, pend0), pm1 = __ballot_sync(FULL, pend1);
                while (pm0 | pm1) {
                    const uint32_t pl = (uint32_t)__ffs((int)(pm0 | pm1)) - 1u;
                    const bool first = (pm0 >> pl) & 1u;                         // warp-uniform: this pair's first symbol is still due
                    const uint32_t s_q = __shfl_sync(FULL, first ? q0 : q1, pl), s_src = __shfl_sync(FULL, first ? src0 : src1, pl),
                                   s_len = __shfl_sync(FULL, first ? len0 : len1, pl);
                    copy_in_order(s_q, s_src, s_len);
                    if (first) pm0 &= ~(1u << pl); else pm1 &= ~(1u << pl);
                }
#endif
#else
                (void)far0; (void)far1; (void)pend0; (void)pend1;
#endif
                __syncwarp();
                if (F & 15u) flush(J1, false); else if ((J1 & ~15u) > F) flush_units(J1 & ~15u);
                kc += np;
                // recycle the stream ring behind this step: every pair up to the last one is consumed, so the chunks before
                // the one the last pair's size byte sits in are free (a step reads up to 1 KiB of stream; freeing only
                // up to the step's FIRST pair would leave the walker less than one step of look-ahead on literal data)
                const uint32_t c0 = plast / kChunk;
                if (c0 != cur_chunk) { cur_chunk = c0; issue_upto(c0 + kChunks); }
            }
            if (done) {
                __syncwarp();
                flush(min(ld_vol_u32(&sm.end_j), size) + oal, true);
            }
            if (lane == 0) st_vol_u32(&sm.consumed, kc);
        }
        wait_upto(issued);
        __syncwarp();
    }
}

template <uint32_t OUT_RING, bool EXT, uint32_t PAIRS, int LB>
__global__ void __launch_bounds__(LB, 1) decode_split_kernel(DecodeArgs a, uint32_t nslots)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    SlotSmem<OUT_RING>* slots = reinterpret_cast<SlotSmem<OUT_RING>*>(smem_raw);
    const unsigned lane = threadIdx.x & 31u;
    const unsigned wid  = threadIdx.x >> 5;

    // the walkers are the LAST warps: the warp arbiter favours the highest warp ids
    if (wid < nslots) {
        SlotSmem<OUT_RING>& sm = slots[wid];
        if (lane == 0) {
            for (uint32_t q = 0; q < kChunks; q++) mbar_init(&sm.bar[q], 1);
            for (uint32_t q = 0; q < kSteps; q++) mbar_init(&sm.full[q], 1);
            sm.ready = 0; sm.consumed = 0; sm.produced = 0; sm.ended = 0;
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();
    if (wid < nslots) {
        const uint64_t stride_slots = (uint64_t)gridDim.x * nslots, b0 = (uint64_t)blockIdx.x * nslots + wid;
        CopierState cs;
        if constexpr (PAIRS == 32) {
            // Per block: lane per pair (64 symbols per step) unless the block is (nearly) incompressible.  A stream of
            // 16-byte literals is walker-bound, and there the 32-descriptor hand-over costs more than it saves (measured:
            // uniform random data 942 vs 669 GB/s, text 341 vs 371 GB/s).
            for (uint64_t b = b0; b < a.nb; b += stride_slots) {
                uint32_t limit;
                stream_of(a, b, limit);
                const bool dense = (uint64_t)limit * 16u >= a.ostride * 15u;      // C/U >= 0.94, or sizes not given
                // ... and unless it is almost all 16-byte matches (C/U < 0.3): periodic data, where every match copies from
                // inside its own step and the in-order copies dominate (8-byte period: 6.5 vs 8.3 ms per 2 GiB)
                const bool runs = (uint64_t)limit * 10u < a.ostride * 3u;
                if ((dense && !TSQB_DEC_DENSE_PAIRS) || runs) copier<OUT_RING, EXT, true>(a, slots[wid], b, stride_slots, lane, cs);
                else       copier_pairs<OUT_RING>(a, slots[wid], b, lane, cs);
            }
        } else copier<OUT_RING, EXT, false>(a, slots[wid], b0, stride_slots, lane, cs);
    } else walker<OUT_RING, EXT, (LB == 1024 && PAIRS == 32)>(a, slots, nslots, wid - nslots, lane);
}

template <uint32_t OUT_RING, bool EXT, uint32_t PAIRS = kPairs, int LB = TSQB_DEC_LB>
cudaError_t launch_split_t(const DecodeArgs& a, uint32_t nslots, unsigned ctas, cudaStream_t st)
{
    const size_t smem = sizeof(SlotSmem<OUT_RING>) * nslots;
    cudaError_t e = cudaFuncSetAttribute(decode_split_kernel<OUT_RING, EXT, PAIRS, LB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    decode_split_kernel<OUT_RING, EXT, PAIRS, LB><<<ctas, (nslots + kWalkers) * 32, smem, st>>>(a, nslots);
    return cudaGetLastError();
}

}  // namespace

// One CTA per SM; every CTA owns `nslots` block slots (copier warps) and kWalkers walker warps.
cudaError_t launch_decode_split(const DecodeArgs& a, bool ext, int sm_count, cudaStream_t st, bool lane_per_pair, int slot_cap)
{
    if (a.nb == 0) return cudaSuccess;
    const size_t budget = 227u * 1024u;
    const uint64_t per_sm = (a.nb + sm_count - 1) / sm_count;
    lane_per_pair = lane_per_pair && !ext;
    // How many block slots per CTA.  Measured (profiles/r01_experiments.md): with blocks of 64 KiB and more a round of
    // blocks takes ~1.7x longer once the slots need more than the 196 KB shared-memory carve-out (the SM then keeps
    // 28 KB of L1 instead of 60 KB), and within that limit fewer slots make a faster round.  So: stay inside the
    // carve-out, take the fewest rounds that allows, and spread the blocks evenly over those rounds.  Small blocks
    // (< 64 KiB) are set-up bound and want every slot the CTA can hold.
    // (a step of the extension format can produce 32 x 64 bytes: its output ring must be >= 4 KiB)
    const size_t slot_bytes = ext ? sizeof(SlotSmem<4096>) : sizeof(SlotSmem<2048>);
    uint32_t cap = kMaxSlots;
    while (slot_bytes * cap > budget) cap--;
    if (slot_cap > 0) { if (cap > (uint32_t)slot_cap) cap = (uint32_t)slot_cap; }      // option "decode_slots": explicit, up to what 227 KB hold
    else if (a.ostride >= 65536u) {
        uint32_t lo = cap;
        while (lo > 1 && slot_bytes * lo > 195u * 1024u) lo--;
        // A round costs ~(1 + 0.03 x slots) walks of a block, 0.4 more above the carve-out, and never less than one walk: the
        // slots above the carve-out pay when they save a round out of two or three (27..30 blocks per SM: one round instead of
        // two, 4096 text blocks of 256 KiB 4.08 -> 3.22 ms; 53..60: two instead of three, 8192 blocks 6.99 -> 6.44 ms) and lose
        // from there on (32768 blocks: 23.1 vs 25.7 ms).  profiles/r02_experiments.md
        const uint64_t rounds_hi = (per_sm + cap - 1) / cap, rounds_lo = (per_sm + lo - 1) / lo;
        if (ext || TSQB_DEC_ONE_ROUND == 0 || !(rounds_hi < rounds_lo && rounds_hi <= 2)) cap = lo;
    }
    const uint64_t rounds = (per_sm + cap - 1) / cap;
    uint32_t nslots = (uint32_t)((per_sm + rounds - 1) / rounds);
    if (nslots == 0) nslots = 1;
    unsigned ctas = (unsigned)((a.nb + nslots - 1) / nslots);
    if (ctas > (unsigned)sm_count) ctas = (unsigned)sm_count;
    // output ring as large as 227 KB of shared memory allows
    if (ext) {
        if (sizeof(SlotSmem<16384>) * nslots <= budget) return launch_split_t<16384, true>(a, nslots, ctas, st);
        if (sizeof(SlotSmem<8192>) * nslots <= budget)  return launch_split_t<8192, true>(a, nslots, ctas, st);
        return launch_split_t<4096, true>(a, nslots, ctas, st);
    }
    // Measured (profiles/r01_experiments.md): a ring that fills all 227 KB leaves the SM without L1 and is ~20 % slower
    // than a 2 KiB ring (3.73 vs 3.03 ms on the bench workload); 1 KiB is no faster.  Keep >= 43 KB for L1.
    const size_t roomy = 184u * 1024u;
    // lane-per-pair copier (64 symbols per step, no-extension format): 896-thread launch bound = 72 registers
    if (lane_per_pair) {
        if ((nslots + kWalkers) * 32u > 896u) return launch_split_t<2048, false, 32, 1024>(a, nslots, ctas, st);   // 27..30 slots: 64 registers
        if (sizeof(SlotSmem<16384>) * nslots <= roomy) return launch_split_t<16384, false, 32, 896>(a, nslots, ctas, st);
        if (sizeof(SlotSmem<8192>) * nslots <= roomy)  return launch_split_t<8192, false, 32, 896>(a, nslots, ctas, st);
        if (sizeof(SlotSmem<4096>) * nslots <= roomy)  return launch_split_t<4096, false, 32, 896>(a, nslots, ctas, st);
        return launch_split_t<2048, false, 32, 896>(a, nslots, ctas, st);
    }
    if (sizeof(SlotSmem<16384>) * nslots <= roomy) return launch_split_t<16384, false>(a, nslots, ctas, st);
    if (sizeof(SlotSmem<8192>) * nslots <= roomy)  return launch_split_t<8192, false>(a, nslots, ctas, st);
    if (sizeof(SlotSmem<4096>) * nslots <= roomy)  return launch_split_t<4096, false>(a, nslots, ctas, st);
    return launch_split_t<2048, false>(a, nslots, ctas, st);
}

}  // namespace tsqb
