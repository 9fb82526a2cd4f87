// tsq_encode_batch.cu -- the production encoder (both formats): one WARP per block, decisions first,
// bytes later.
//
// Reference semantics: tsqEncodeNoext (tsq_encode.cpp:48-189) and the extension variant (:200-341, template EXT:
// match lengths up to 64, mlen[] of :44-45); bit-exact, see SURVEY.md 8(a).
//
// The greedy parse is a serial chain (every probe reads a table entry written by an earlier probe,
// tsq_encode.cpp:74-79), so the warp spends its time in a warp-uniform decision loop.  This kernel
// keeps that loop free of memory traffic and of byte shuffling:
//
//   1. WINDOW PRECOMPUTE (parallel, lane L <-> position base + L): the 4-byte word and hash, the
//      table entry as committed before the window, the candidate's 16 bytes and hence the match
//      length against it (capped at 16, tsq_encode.cpp:126-137).  All of this is independent of the
//      parse, so every global load of the window is issued at once.  Positions of the window with
//      the same hash are found with __match_any_sync; the decision loop redirects a probe to the
//      nearest earlier lane that was really inserted ("last writer wins", :79).
//   2. DECISION LOOP (warp-uniform): first hit of a literal scan = ballot + ffs; match length =
//      one shuffle of the precomputed length, then the clamps of :139-145; the probe that follows a
//      match (:162-170) is another lane of the same window.  The loop only appends TOKENS
//      (literal run <= 16 bytes / match) to a small shared-memory queue and tracks the input
//      position at the start of the open pair (rep_last_i).
//   3. EMISSION (parallel, lane s <-> token s, 32 tokens at a time): the layout of the stream is a
//      pure function of the tokens -- symbol s of a batch that starts at byte Bj sits at
//      Bj + (s/8 + 1) control bytes + (s/2 + 1) size bytes + the payload bytes before it
//      (:57-61,94-95,157-159) -- so one warp prefix sum places all 32 payloads; control bytes are a
//      ballot, size bytes a shuffle.  Payloads are staged in a shared-memory ring with exact-length
//      byte stores (the reference copies a blind 16 bytes per literal run, :88,108, and lets the next
//      symbol overwrite the tail) and leave for HBM as coalesced 128-bit stores.
//   4. The inserts of the window are committed to the table (global memory, one table per block in
//      flight) when the window is left.
//
// Nothing is stored past the returned size; the two trailing never-initialised bytes of the
// reference (:176-188) are reproduced as in tsq_encode_common.cuh.
#include "tsq_encode_common.cuh"

namespace tsqb {

namespace {

constexpr unsigned FULL   = 0xffffffffu;
constexpr uint32_t kTok   = 64;                      // token queue entries per warp
constexpr uint32_t kORing = 1024;                    // output staging ring per warp (bytes)
constexpr uint32_t kOMask = kORing - 1;
#ifndef TSQB_ENC_WARPS
#define TSQB_ENC_WARPS 4
#endif
constexpr int      kWarps = TSQB_ENC_WARPS;          // warps (= blocks in flight) per CTA

// Development knobs (compile-time; scripts/build_variants.sh builds one library per combination, profiles/r01_experiments.md
// records what they measured).  The defaults are the production kernel.
#ifndef TSQB_ENC_FIXED_GRID
#define TSQB_ENC_FIXED_GRID 0      // 1: windows on a fixed grid of 32 positions (1 + 32k) instead of starting at the next probe
#endif
#ifndef TSQB_ENC_PF_NEXT
#define TSQB_ENC_PF_NEXT 0         // 1 (needs the fixed grid): L2 prefetch of the next window's table sectors before the decision loop
#endif
#ifndef TSQB_ENC_HINTS
#define TSQB_ENC_HINTS 0           // 1: the run-time experiment hints (EncodeArgs::hints) are compiled in (costs 5 %)
#endif
#ifndef TSQB_ENC_DIAG
#define TSQB_ENC_DIAG 0            // timing diagnostics, WRONG OUTPUT: 1 = table reads folded onto 256 sectors per block (L2-resident),
#endif                             // 2 = commits not stored, 3 = both
#ifndef TSQB_ENC_L2POL
#define TSQB_ENC_L2POL 0           // development knob: bit 0 = table loads carry an L2 evict_first policy, bit 1 = commits an evict_last one
#endif
#ifndef TSQB_ENC_L2_64B
#define TSQB_ENC_L2_64B 2          // table loads: 0 = plain, 1 = L2 fills 64 bytes per miss instead of a 128-byte line, 2 = that and no L1 allocation
                                   // (27.86 / 27.60 / 27.36 ms; halves the DRAM bytes per probe)
#endif
#ifndef TSQB_ENC_ALIAS8
#define TSQB_ENC_ALIAS8 0          // 1: alias tags taken from the word's top byte only (one byte load instead of two)
#endif

// token: bit 31 = literal, bits 27..30 = length - 1, low bits = literal source position / match offset
constexpr uint32_t kTokLit = 0x80000000u;

// "Was any entry of this group of 4 hashes written in this block?" -- one bit per 4 table entries, 4 KiB per warp.
// A clear bit proves the entry is empty, so the probe needs no memory access at all (exact, never a guess): most
// probes of small blocks and ~13 % of the probes of a 256 KiB text block.
constexpr uint32_t kBloomWords = kHashSlots / 4u / 32u;

struct __align__(16) WarpWs {
    uint8_t  oring[kORing];
    uint32_t tok[kTok];
    uint32_t written[kBloomWords];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 16 bytes at an arbitrary address of the (read-only) input
__device__ __forceinline__ void ldg16(const uint8_t* p, uint32_t v[4])
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3u) * 8u;
    uint32_t t[5];
#pragma unroll
    for (int m = 0; m < 5; m++) t[m] = __ldg(w + m);
#pragma unroll
    for (int m = 0; m < 4; m++) v[m] = __funnelshift_r(t[m], t[m + 1], sh);
}

// number of equal leading bytes of two 16-byte strings (tsq_encode.cpp:126-137)
__device__ __forceinline__ uint32_t prefix16(const uint32_t a[4], const uint32_t b[4])
{
    const uint64_t lo = ((uint64_t)(a[1] ^ b[1]) << 32) | (a[0] ^ b[0]);
    const uint64_t hi = ((uint64_t)(a[3] ^ b[3]) << 32) | (a[2] ^ b[2]);
    if (lo) return (uint32_t)(__ffsll((long long)lo) - 1) >> 3;
    if (hi) return 8u + ((uint32_t)(__ffsll((long long)hi) - 1) >> 3);
    return 16u;
}

// byte t of a payload is stored iff bit t of `lenmask` is set (a literal run of n bytes: n ones; a match offset: 2 ones);
// ptxas materialises the mask as predicates with two R2P, so exact lengths cost no more than a blind 16-byte run
template <int T>
__device__ __forceinline__ void store_bytes(uint32_t ad, const uint32_t v[4], uint32_t lenmask)
{
    if (lenmask & (1u << T))
        asm volatile("st.shared.u8 [%0+%1], %2;" ::"r"(ad), "n"(T), "r"(v[T >> 2] >> (8 * (T & 3))) : "memory");
    if constexpr (T > 0) store_bytes<T - 1>(ad, v, lenmask);
}

// Table entry of this kernel.  The reference stores the low 16 bits of the position (tsq_encode.cpp:79);
// here an entry is one 32-byte sector:
//   A.x  position (22 bits; its low 16 bits play the reference's role, so the candidate is the same)
//        | low 10 bits of the tag;   A.y  high 5 bits of the tag;   A.z / B.z  epoch (low / high word)
//   A.w, B.x, B.y  the 12 input bytes behind the hashed word
//   A.y bits 5..19, B.w bits 0..14 / 15..29  ALIAS TAGS: the tags of the words 64 / 128 / 192 KiB behind the position.
//        An entry older than 64 KiB names, through its low 16 bits, a position m x 64 KiB further on
//        (tsq_encode.cpp:77-78); the probe hits only if the word THERE equals the probing word, which needs equal
//        tags.  The committing lane reads those (sequential, cache-resident) words once, so a later probe of the
//        aged entry is answered from the entry itself: unequal tags = certain miss, no second DRAM access.
// * tag = word >> 17.  hash = (w ^ (w >> 12)) & 0x1FFFF determines bits 0..16 of w once bits 17..31 are known
//   (b_i = h_i ^ b_{i+12} downwards), so "same slot and same tag" is EXACTLY "same 4-byte word": the hit test of
//   :100 and the match length (:126-137) need no access to the candidate's bytes -- the second, dependent DRAM
//   access of a probe disappears.
// * an entry counts only if it was written under the current epoch (one per encoded block), which replaces
//   tsqInit's memset of the table (tsq_context.cpp:77-80): nothing is zeroed per block.
// * a commit writes the whole sector, so DRAM needs no read-modify-write for it.
// Blackwell has 256-bit global loads / stores (SASS LDG.E.ENL2.256 / STG.E.ENL2.256): one request per probe,
// and a commit that covers its whole sector.
__device__ __forceinline__ void load_entry(const uint4* table, uint32_t h, uint4& A, uint4& B, uint64_t pol)
{
    if (pol)
        asm volatile("ld.global.L1::no_allocate.L2::cache_hint.L2::64B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                     : "=r"(A.x), "=r"(A.y), "=r"(A.z), "=r"(A.w), "=r"(B.x), "=r"(B.y), "=r"(B.z), "=r"(B.w) : "l"(table + 2u * h), "l"(pol) : "memory");
    else
#if TSQB_ENC_L2_64B == 1
        asm volatile("ld.global.L2::64B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(A.x), "=r"(A.y), "=r"(A.z), "=r"(A.w), "=r"(B.x), "=r"(B.y), "=r"(B.z), "=r"(B.w) : "l"(table + 2u * h) : "memory");
#elif TSQB_ENC_L2_64B == 2
        asm volatile("ld.global.L1::no_allocate.L2::64B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(A.x), "=r"(A.y), "=r"(A.z), "=r"(A.w), "=r"(B.x), "=r"(B.y), "=r"(B.z), "=r"(B.w) : "l"(table + 2u * h) : "memory");
#else
        asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(A.x), "=r"(A.y), "=r"(A.z), "=r"(A.w), "=r"(B.x), "=r"(B.y), "=r"(B.z), "=r"(B.w) : "l"(table + 2u * h) : "memory");
#endif
}

// tag (word >> 17) of the 4-byte word at in[pos]: bits 17..31 live in bytes 2 and 3
__device__ __forceinline__ uint32_t tag_at(const uint8_t* __restrict__ in, uint32_t pos)
{
#if TSQB_ENC_ALIAS8
    return (uint32_t)__ldg(in + pos + 3u) << 7;                      // top byte only: bits 7..14 of the tag
#endif
    return ((uint32_t)__ldg(in + pos + 3u) << 7) | ((uint32_t)__ldg(in + pos + 2u) >> 1);
}

// a1..a3: alias tags, i.e. tag_at(pos + 64 KiB * {1, 2, 3}) where that position lies inside the block (a candidate
// always precedes the probing position, so no other alias can ever be asked for).
__device__ __forceinline__ void store_entry(uint4* table, uint32_t h, uint32_t pos, const uint32_t own[4], uint64_t epoch,
                                            uint32_t a1, uint32_t a2, uint32_t a3, uint64_t pol)
{
    const uint32_t tag = own[0] >> 17;
    if (pol)
        asm volatile("st.global.L2::cache_hint.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8}, %9;" ::"l"(table + 2u * h), "r"(pos | (tag << 22)), "r"((tag >> 10) | (a1 << 5)),
                     "r"((uint32_t)epoch), "r"(own[1]), "r"(own[2]), "r"(own[3]), "r"((uint32_t)(epoch >> 32)), "r"(a2 | (a3 << 15)), "l"(pol) : "memory");
    else
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(table + 2u * h), "r"(pos | (tag << 22)), "r"((tag >> 10) | (a1 << 5)),
                     "r"((uint32_t)epoch), "r"(own[1]), "r"(own[2]), "r"(own[3]), "r"((uint32_t)(epoch >> 32)), "r"(a2 | (a3 << 15)) : "memory");
}

__device__ __forceinline__ uint32_t lanes_from_to(uint32_t lo, uint32_t hi)   // bits lo..hi inclusive
{
    return ((2u << hi) - 1u) & ~((1u << lo) - 1u);
}

// Piece-streamed input: wait until `want` bytes of the block's prefix have landed.  The flag is written by a copy that is
// stream-ordered behind the data, and read around L1; the loads that follow are issued after the branch on its value.
// A wait that lasts seconds means the host side died: report it and carry on (the call fails) instead of hanging the GPU.
__device__ __forceinline__ uint32_t wait_arrived(const uint32_t* flag, uint32_t want, uint32_t* err)
{
    uint32_t v;
    const long long t0 = clock64();
    for (;;) {
        asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v >= want) return v;
        __nanosleep(500);
        if (clock64() - t0 > 20000000000ll) { *err = 1u; return 0xFFFFFFFFu; }   // ~10 s
    }
}

struct BlockEncoder {
    // ---- fixed for the block
    const uint8_t* __restrict__ in;
    uint8_t*  o_al;           // output slot, rounded down to 16 bytes
    uint32_t  oal;            // slot address & 15: ring/flush positions are q = byte offset + oal
    uint32_t  size;
    uint32_t  obase, tbase;   // shared addresses of the output ring / token queue
    unsigned  lane;
    // ---- parse state (warp-uniform)
    uint32_t  n;              // symbols decided so far
    uint32_t  rep;            // input position at the start of the open pair (rep_last_i)
    uint32_t  th, nt;         // token queue head / count
    // ---- emission state (warp-uniform)
    uint32_t  Bj;             // byte offset of the control byte of the next batch's first group
    uint32_t  F;              // flushed up to here (q coordinates)
    uint32_t  lit_js, lit_src;// output offset / input position of the last literal run (0x80000000: none)
    bool      stream_out;     // experiment: flushed output leaves with st.global.cs

    __device__ __forceinline__ void push(uint32_t tok)
    {
        if (lane == 0) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(tbase + (((th + nt) & (kTok - 1)) << 2)), "r"(tok) : "memory");
        nt++;
    }

    // tsq_encode.cpp:93-95 / :157-159, the part the parse needs
    __device__ __forceinline__ void symbol_done(uint32_t in_pos)
    {
        n++;
        if ((n & 1u) == 0) rep = in_pos;
    }

    // tsq_encode.cpp:85-97 / :105-117 -- pending literals [from, upto) leave as runs of at most 16
    __device__ __forceinline__ void literals(uint32_t& from, uint32_t upto)
    {
        do {
            uint32_t cnt = upto - from;
            if (cnt > 16u) cnt = 16u;
            push(kTokLit | ((cnt - 1u) << 27) | from);
            from += cnt;
            symbol_done(from);
        } while (upto - from > 0);
    }

    __device__ __forceinline__ void match(uint32_t offset, uint32_t nibble, uint32_t in_pos_after)
    {
        push((nibble << 27) | offset);
        symbol_done(in_pos_after);
    }

    __device__ __forceinline__ void ring_put(uint32_t q, uint32_t v)
    {
        asm volatile("st.volatile.shared.u8 [%0], %1;" ::"r"(obase + (q & kOMask)), "r"(v) : "memory");
    }

    __device__ __forceinline__ uint32_t ring_get(uint32_t q)
    {
        uint32_t v;
        asm volatile("ld.volatile.shared.u8 %0, [%1];" : "=r"(v) : "r"(obase + (q & kOMask)) : "memory");
        return v;
    }

    // complete 16-byte units below E (q coordinates) -> HBM
    __device__ __forceinline__ void flush_to(uint32_t E, bool final)
    {
        if (!final) E &= ~15u;
        if (F >= E) return;
        if (F & 15u) {                                                           // unaligned head of the slot
            const uint32_t h = min(E, (F + 15u) & ~15u);
            if (lane < h - F) o_al[F + lane] = (uint8_t)ring_get(F + lane);
            F = h;
        }
        const uint32_t Ev = E & ~15u;
        for (uint32_t at = F + 16u * lane; at < Ev; at += 512u) {
            uint4 x;
            asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "r"(obase + (at & kOMask)) : "memory");
            if (stream_out) __stcs(reinterpret_cast<uint4*>(o_al + at), x);
            else *reinterpret_cast<uint4*>(o_al + at) = x;
        }
        if (Ev > F) F = Ev;
        if (F < E) {                                                             // final tail
            if (lane < E - F) o_al[F + lane] = (uint8_t)ring_get(F + lane);
            F = E;
        }
    }

    // Emit `cnt` (<= 32) tokens from the head of the queue: lane s places symbol s.
    // Returns the byte offset just behind the last payload.
    __device__ uint32_t emit(uint32_t cnt)
    {
        __syncwarp();                                                            // tokens visible
        uint32_t tok = 0;
        if (lane < cnt) asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(tok) : "r"(tbase + (((th + lane) & (kTok - 1)) << 2)) : "memory");
        const bool valid = lane < cnt;
        const bool lit = valid && (tok & kTokLit);
        const uint32_t nibble = (tok >> 27) & 15u;
        const uint32_t pl = valid ? (lit ? nibble + 1u : 2u) : 0u;               // payload bytes

        // inclusive warp prefix sum of the payload lengths
        uint32_t incl = pl;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, incl, d);
            if (lane >= (unsigned)d) incl += t;
        }
        const uint32_t P = incl - pl;
        const uint32_t total = __shfl_sync(FULL, incl, 31);
        const uint32_t pos = Bj + (lane >> 3) + 1u + (lane >> 1) + 1u + P;        // byte offset of this payload
        const uint32_t q = pos + oal;

        // ---- payloads: literal run = its 1..16 bytes (:88-91), match = 2-byte offset (:152-153)
        uint32_t v[4] = {tok & 0xFFFFu, 0, 0, 0};
        if (lit) ldg16(in + (tok & 0x3FFFFFu), v);
        if (valid) {
            const uint32_t lenmask = (1u << pl) - 1u;
            if ((q & kOMask) + 16u <= kORing) store_bytes<15>(obase + (q & kOMask), v, lenmask);
            else {
#pragma unroll
                for (int t = 15; t >= 0; t--)
                    if (lenmask & (1u << t)) ring_put(q + (uint32_t)t, v[t >> 2] >> (8 * (t & 3)));
            }
        }
        // remember the last literal run for the never-initialised trailing bytes (finish)
        {
            const uint32_t lm = __ballot_sync(FULL, lit);
            if (lm) {
                const uint32_t L = 31u - (uint32_t)__clz((int)lm);
                lit_js = __shfl_sync(FULL, pos, L);
                lit_src = __shfl_sync(FULL, tok & 0x3FFFFFu, L);
            }
        }
        __syncwarp();                                                            // payloads are down
        // ---- size bytes: one per pair, right before the pair's first payload (:95,159).  A trailing
        // odd symbol's byte is (nibble << 4) (:183-186).
        {
            const uint32_t other = __shfl_down_sync(FULL, nibble, 1);
            if (valid && !(lane & 1u)) ring_put(q - 1u, (nibble << 4) | ((lane + 1u < cnt) ? other : 0u));
        }
        // ---- control bytes: one per 8 symbols, right before the size byte of the group's first pair;
        // bit (7 - t) = symbol t is a literal (:94,158); a trailing partial group is padded with ones (:176-182)
        {
            const uint32_t lm = __ballot_sync(FULL, lit);
            if (valid && !(lane & 7u)) {
                const uint32_t have = min(8u, cnt - lane);
                const uint32_t bits = __brev((lm >> lane) & 0xFFu) >> 24;        // symbol t -> bit 7 - t
                ring_put(q - 2u, bits | ((1u << (8u - have)) - 1u));
            }
        }
        __syncwarp();
        th += cnt; nt -= cnt;
        const uint32_t end = Bj + ((cnt + 7u) >> 3) + ((cnt + 1u) >> 1) + total;  // behind the last payload
        if (cnt == 32u) {
            Bj = end;
            flush_to(end + oal, false);
        }
        return end;
    }

    // End of block, after everything queued was emitted: the padding rules of tsq_encode.cpp:176-188.
    // Returns the stream size; `flags` as Emitter::finish.
    __device__ uint32_t finish(uint32_t& flags)
    {
        if (nt >= 32u) emit(32u);
        const uint32_t cnt = nt;                                                 // < 32
        const uint32_t jend = cnt ? emit(cnt) : Bj;                               // cnt == 0: Bj is where the next control byte goes
        const uint32_t r8 = n & 7u;
        uint32_t out_size, keep = 0;                                             // keep: trailing bytes left as pre-filled
        flags = 0;
        auto stale = [&](uint32_t a, uint32_t& val) -> bool {
            if (a - lit_js < 16u) { val = in[lit_src + (a - lit_js)]; return true; }
            return false;
        };
        uint32_t val;
        if (r8 == 0) {
            // a fresh control byte and a fresh size byte were allocated and never written
            out_size = jend + 2u;
            const bool s0 = stale(jend, val);
            if (s0) { if (lane == 0) ring_put(jend + oal, val); } else flags |= kTailCtlPrefill;
            const bool s1 = stale(jend + 1u, val);
            if (s1) { if (lane == 0) ring_put(jend + 1u + oal, val); } else flags |= kTailNibPrefill;
            keep = s1 ? 0u : (s0 ? 1u : 2u);
        } else if ((n & 1u) == 0) {
            // a fresh size byte: the reference shifts whatever it held (:183-186)
            out_size = jend + 1u;
            if (stale(jend, val)) { if (lane == 0) ring_put(jend + oal, val << 4); }
            else { flags |= kTailNibShifted; keep = 1u; }
        } else {
            out_size = jend;
        }
        __syncwarp();
        flush_to(out_size - keep + oal, true);
        if ((flags & kTailNibShifted) && lane == 0) {                             // byte = (pre-fill << 4)
            uint8_t* p = o_al + oal + jend;
            *p = (uint8_t)(*p << 4);
        }
        return out_size;
    }
};

// FAT = sector entries in a 4 MiB table (thousands of blocks in flight: the tables live in HBM and a probe must
// not need a second, dependent access); !FAT = the reference's own 2^17 x u16 table, zeroed per block, for a few
// hundred blocks in flight, whose tables (256 KiB each) and 64 KiB back-windows stay resident in the 126 MB L2.
// STREAM (sector entries only): the block's bytes arrive while it is being encoded (EncodeArgs::arrived); the alias tags,
// which read up to 192 KiB ahead of the scan, are left out -- an aged entry is then answered by reading the candidate's bytes.
template <bool FAT, bool EXT, bool STREAM>
__device__ uint32_t encode_block_batch(void* __restrict__ table_, const uint64_t epoch, const uint8_t* __restrict__ in, const uint32_t size,
                                       uint8_t* __restrict__ out, const unsigned lane, WarpWs& ws, uint32_t& flags, const uint32_t hints_,
                                       const uint32_t* arrived = nullptr, const uint32_t* arrived_next = nullptr, uint32_t* stream_error = nullptr)
{
    const uint32_t hints = TSQB_ENC_HINTS ? hints_ : 0u;
    uint64_t pol = 0;                                                  // experiment: table traffic marked evict-first in L2
    if (hints & 1u) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    uint64_t pol_ld = pol, pol_st = pol;
    if (TSQB_ENC_L2POL & 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_ld));
    if (TSQB_ENC_L2POL & 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_st));
    BlockEncoder e;
    e.stream_out = (hints & 2u) != 0;
    e.in = in; e.size = size; e.lane = lane;
    e.oal = (uint32_t)(reinterpret_cast<uintptr_t>(out) & 15u);
    e.o_al = out - e.oal;
    e.obase = smem_u32(ws.oring); e.tbase = smem_u32(ws.tok);
    e.n = 0; e.rep = 0; e.th = 0; e.nt = 0;
    e.Bj = 3; e.F = e.oal; e.lit_js = 0x80000000u; e.lit_src = 0;
    if (lane < 3) e.ring_put(lane + e.oal, size >> (8u * lane));              // tsq_encode.cpp:53-55
    __syncwarp();

    uint4* const table = static_cast<uint4*>(table_);                  // FAT
    uint16_t* const table16 = static_cast<uint16_t*>(table_);           // !FAT
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t i = 0, lit_from = 0;
    uint32_t base = 1;                // first probe is position 1 (:70-72)
    bool chain_pending = false;       // the next probe is a post-match probe (:162-170)
    uint32_t c0 = 0;                  // lane of the window's first probe (fixed grid: the parse may enter a window anywhere)

    uint32_t landed = 0;              // STREAM: bytes of this block known to have arrived
    bool next_ok = arrived_next == nullptr;
    for (;;) {                                                         // one window per iteration
        if constexpr (STREAM) {
            // everything this window can read: the lanes' 16 bytes (x + 19), a match attempt (i + 64), a literal run's blind
            // 16 bytes -- all below base + 128.  Past the block's end that is the first bytes of the block behind it.
            const uint32_t need = base + 128u;
            if (min(need, size) > landed) landed = wait_arrived(arrived, min(need, size), stream_error);
            if (need > size && !next_ok) { wait_arrived(arrived_next, 128u, stream_error); next_ok = true; }
        }
        // ---------------- window precompute (parallel over 32 positions)
        const uint32_t x = base + lane;
        uint32_t own[4], cb[4];
        ldg16(in + x, own);
        const uint32_t w = own[0];
        const uint32_t h = hash17(w);
        // alias tags for this position's own entry, should it be committed: loaded now, with everything else of the
        // window (sequential streams 64 / 128 / 192 KiB ahead of the scan), so that the commit waits for nothing
        uint32_t a1 = 0, a2 = 0, a3 = 0;
        if constexpr (FAT && !STREAM) {
            if (x + 65536u < size)  a1 = tag_at(in, x + 65536u);
            if (x + 131072u < size) a2 = tag_at(in, x + 131072u);
            if (x + 196608u < size) a3 = tag_at(in, x + 196608u);
        }
        const uint32_t M = __match_any_sync(FULL, h);
        uint32_t tab_cand, m_tab;
        if constexpr (FAT) {
            uint4 A = make_uint4(0, 0, 0, 0), B = A;                   // all-zero = an entry of no epoch
            if ((ws.written[h >> 7] >> ((h >> 2) & 31u)) & 1u) load_entry(table, (TSQB_ENC_DIAG & 1) ? (h & 255u) : h, A, B, pol_ld);
            // An entry of another epoch is the reference's zero entry: candidate = start of the 64 KiB segment
            // (expand_pos(0, x)), one shared, cache-resident location.
            const bool live = A.z == (uint32_t)epoch && B.z == (uint32_t)(epoch >> 32);
            const uint32_t p22 = live ? (A.x & 0x3FFFFFu) : 0u;
            tab_cand = expand_pos(p22 & 0xFFFFu, x);
            if (live && tab_cand == p22) {
                // the entry really describes in[tab_cand]: same hash and same high word bits <=> same word (:100)
                const uint32_t tag = ((A.x >> 22) | (A.y << 10)) & 0x7FFFu;
                cb[0] = own[0]; cb[1] = A.w; cb[2] = B.x; cb[3] = B.y;
                m_tab = tag == (w >> 17) ? prefix16(own, cb) : 0u;
            } else {                                                   // empty entry, or one older than 64 KiB (aliased)
                // aliased by 1..3 segments: the entry carries the tag of the word at the aliased position
                const uint32_t seg = live ? (tab_cand - p22) >> 16 : 0u;
                const uint32_t atag = seg == 1u ? (A.y >> 5) : (seg == 2u ? B.w : (B.w >> 15));
                constexpr uint32_t kAliasMask = TSQB_ENC_ALIAS8 ? 0x7F80u : 0x7FFFu;
                if (!STREAM && seg - 1u < 3u && ((atag ^ (w >> 17)) & kAliasMask) != 0u) m_tab = 0u;   // different words: the probe fails (:100)
                else {
                    ldg16(in + tab_cand, cb);
                    m_tab = prefix16(own, cb);                         // >= 4  <=>  the 4-byte words are equal (:100)
                }
            }
        } else {
            tab_cand = expand_pos(table16[h], x);                      // :76-78
            ldg16(in + tab_cand, cb);
            m_tab = prefix16(own, cb);
        }
        // ---- experiments (hints): warm the L2 for the windows to come while this window's loads are in flight
        if (hints & 4u) {                                              // the input line ~6 windows ahead (first touch = DRAM)
            if (lane == 0 && base + 192u < size) asm volatile("prefetch.global.L2 [%0];" ::"l"(in + base + 192u));
        }
        if constexpr (FAT) {
            if (hints & 8u) {                                          // table sectors of the next window (it starts 32..48 positions on)
                const uint32_t xn = x + 32u;
                if (xn < size) {
                    const uint32_t hn = hash17(ld_le32(in + xn));
                    if ((ws.written[hn >> 7] >> ((hn >> 2) & 31u)) & 1u) asm volatile("prefetch.global.L2 [%0];" ::"l"(table + 2u * hn));
                }
                if (lane < 16u && xn + 32u < size) {
                    const uint32_t hn = hash17(ld_le32(in + xn + 32u));
                    if ((ws.written[hn >> 7] >> ((hn >> 2) & 31u)) & 1u) asm volatile("prefetch.global.L2 [%0];" ::"l"(table + 2u * hn));
                }
            }
        }
#if TSQB_ENC_FIXED_GRID && TSQB_ENC_PF_NEXT
        if constexpr (FAT) {
            // the next window is positions x + 32: its table sectors start their way from DRAM to L2 now and are read,
            // after this window's commits, from L2
            if (x + 32u < size) {
                const uint32_t hn = hash17(ld_le32(in + x + 32u));
                if ((ws.written[hn >> 7] >> ((hn >> 2) & 31u)) & 1u) asm volatile("prefetch.global.L2 [%0];" ::"l"(table + 2u * hn));
            }
        }
#endif
        const bool anydup = __any_sync(FULL, M != (1u << lane));
        uint32_t inP = 0, c = c0;
        bool done = false;

        // A lane is a SIMPLE hit when its probe succeeds and yields the match (pos = tab_cand, k = m_tab)
        // whatever the parse did before it: the open pair started at most 16 bytes earlier (rep in
        // [x - 16, x]), so with D = x - tab_cand in [32, 0xFFFE] the offset test (:100,:144-145,:170) holds
        // and the source cannot reach the pair start (:139-141); x + 16 < size - 5 keeps the end-of-block
        // conditions (:170-172) out of the picture.  A lane whose hash occurs twice in the window is left to the general path.
        const uint32_t D = x - tab_cand;
        // Extension format: symbols are up to 64 bytes, so the pair start is up to 64 bytes back and a 16-byte
        // prefix may be the start of a longer match (:276-290), which the general path measures.
        const bool simple = M == (1u << lane) && m_tab >= 4u && (!EXT || m_tab < 16u) && D >= (EXT ? 128u : 32u) && D <= 0xFFFEu &&
                            size > 21u && x + 16u < size - 5u;
        const uint32_t simple_mask = __ballot_sync(FULL, simple);
        // The same for the hit that ENDS a literal scan: there the test uses the pair start from before the
        // pending literals are flushed (:80-100), up to 16 + 31 bytes back, hence the wider margin.
        const uint32_t simple_scan_mask = __ballot_sync(FULL, simple && D >= (EXT ? 192u : 64u));
        // A lane is a CERTAIN MISS when its word differs from its candidate's and no other lane of the window
        // can change that candidate: the probe fails whatever the parse did (and the block does not end nearby).
        const uint32_t miss_mask = __ballot_sync(FULL, M == (1u << lane) && m_tab < 4u && size > 21u && x + 16u < size - 5u);
        const uint32_t nxt = lane + m_tab;                             // lane of the probe that follows this lane's match

        while (c < 32u) {
            if (e.nt >= 32u) e.emit(32u);

            // ------------ fast path: a chain of simple matches starting at the post-match probe of lane c.
            // One shuffle per match finds the chain; its tokens are then built by the matched lanes in parallel.
            if (chain_pending && ((simple_mask >> c) & 1u)) {
                uint32_t V = 0, cur = c;
                do {
                    V |= 1u << cur;
                    cur = __shfl_sync(FULL, nxt, cur);
                } while (cur < 32u && ((simple_mask >> cur) & 1u));
                const uint32_t below = V & lt;
                const uint32_t rank = (uint32_t)__popc(below), cnt = (uint32_t)__popc(V);
                const uint32_t pred = below ? 31u - (uint32_t)__clz(below) : lane;
                const uint32_t x_pred = __shfl_sync(FULL, x, pred);
                const uint32_t last = 31u - (uint32_t)__clz(V);
                // pair start seen by this lane's symbol (rep_last_i): its own start when it opens a pair, else
                // the start of the symbol before it (:159: rep moves when the symbol count becomes even)
                const uint32_t r = (((e.n + rank) & 1u) == 0) ? x : (below ? x_pred : e.rep);
                if ((V >> lane) & 1u)
                    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(e.tbase + (((e.th + e.nt + rank) & (kTok - 1)) << 2)),
                                 "r"(((m_tab - 1u) << 27) | (r - tab_cand)) : "memory");
                inP |= V;
                e.nt += cnt;
                e.n += cnt;
                i = base + cur;                                        // cur = last + k(last): position after the chain
                e.rep = (e.n & 1u) ? base + last : i;
                c = cur;                                               // the next probe is again a post-match probe
                if (cur < 32u && ((miss_mask >> cur) & 1u)) {
                    // ... and it certainly fails (:170): the byte at lane cur starts a literal run (:66-68)
                    chain_pending = false;
                    inP |= 1u << cur;
                    lit_from = i;
                    const uint32_t rest = ~miss_mask & ~((2u << cur) - 1u);    // lanes behind cur that might hit
                    const uint32_t H2 = rest ? (uint32_t)__ffs((int)rest) - 1u : 32u;
                    if (H2 < 32u && ((simple_scan_mask >> H2) & 1u)) {
                        // certain misses up to H2 - 1, a certain hit at H2: the scan of :70-100 ends there.  The run
                        // is H2 - cur <= 31 bytes, so no forced flush happens on the way (:80-98).
                        inP |= lanes_from_to(cur + 1u, H2);
                        i = base + H2;
                        e.literals(lit_from, i);                       // :103-118
                        c = H2;
                        chain_pending = true;                          // same outcome as a post-match probe at H2
                    } else if (H2 == 32u) {
                        // certain misses up to the end of the window: the run continues in the next one
                        if (cur < 31u) inP |= lanes_from_to(cur + 1u, 31u);
                        i = base + 31u;
                        c = 32u;
                    } else {
                        c = cur + 1u;                                  // an uncertain lane ahead: the general scan takes over
                    }
                }
                continue;
            }
            // ------------ effective candidate of every lane given the lanes in P
            uint32_t cand = tab_cand;
            uint32_t m = m_tab;
            bool ovr = false;                                          // candidate is a lane of this window
            if (anydup) {
                const uint32_t cm = M & lt & (inP | ~((1u << c) - 1u));
                const uint32_t q = cm ? 31u - (uint32_t)__clz(cm) : lane;
                const uint32_t wq = __shfl_sync(FULL, w, q);
                if (cm) { cand = base + q; m = (wq == w) ? 4u : 0u; ovr = true; }   // exact length is taken on demand
            }

            uint32_t H, pos;
            if (chain_pending) {
                // lane c is the probe that follows a match (:162-170); i == base + c
                chain_pending = false;
                inP |= 1u << c;
                pos = __shfl_sync(FULL, cand, c);
                const bool eq = (__ballot_sync(FULL, m >= 4u) >> c) & 1u;
                const uint32_t off = e.rep - pos;
                if (!(i < size - 5u && eq && (off - 4u) < 0xFFFBu)) {
                    if (!(i < size)) { done = true; break; }
                    lit_from = i;                                      // outer loop restarts (:66-68)
                    c++;
                    continue;
                }
                H = c;
            } else {
                // ------------ literal scan from lane c (:70-100)
                const uint32_t Fl = lit_from + 32u - base;             // lane of the forced flush (:80-98)
                const uint32_t E = size - base;                        // lane with x == size
                const uint32_t hi = min(31u, min(Fl, E));
                const bool lanehit = m >= 4u && ((e.rep - cand - 4u) < 0xFFFBu) && lane >= c && lane <= hi && x < size;
                const uint32_t hm = __ballot_sync(FULL, lanehit);
                if (hm == 0) {
                    if (E <= hi) {                                     // ran into the end of the block
                        i = size;
                        if (i - lit_from > 31u) e.literals(lit_from, i);
                        if (i - lit_from > 0u) e.literals(lit_from, i);
                        done = true;
                        break;
                    }
                    inP |= lanes_from_to(c, hi);
                    i = base + hi;
                    if (Fl <= hi) e.literals(lit_from, i);             // 32 pending literals: rep moves, scan goes on
                    c = hi + 1u;
                    continue;
                }
                H = (uint32_t)__ffs((int)hm) - 1u;
                inP |= lanes_from_to(c, H);
                i = base + H;
                if (i - lit_from > 31u) e.literals(lit_from, i);       // flush precedes the loop test (:80-100)
                if (i - lit_from > 0u) e.literals(lit_from, i);        // :103-118
                if ((simple_mask >> H) & 1u) { c = H; chain_pending = true; continue; }   // same outcome as a post-match probe
                pos = __shfl_sync(FULL, cand, H);
            }

            // ---------------- one match attempt at i == base + H against pos (:126-160)
            {
                uint32_t k = __shfl_sync(FULL, m, H);                  // common prefix, capped at 16
                const bool in_window = __shfl_sync(FULL, (uint32_t)ovr, H) != 0;
                if (in_window || (EXT && k == 16u)) {
                    // compare now: an in-window candidate, or (extension format) a prefix that may run on to 64 bytes
                    const uint32_t lim = EXT ? 64u : 16u;
                    k = lim;
                    for (uint32_t t0 = 0; t0 < lim; t0 += 32u) {
                        const uint32_t t = t0 + lane;
                        const bool ne = t < lim && __ldg(in + i + t) != __ldg(in + pos + t);
                        const uint32_t nm = __ballot_sync(FULL, ne);
                        if (nm) { k = t0 + (uint32_t)__ffs((int)nm) - 1u; break; }
                    }
                }
                const uint32_t room = e.rep - pos;
                if (k > room) k = room - 1u;                           // :139-141
                if (k < 4u || !((room - 4u) < 0xFFFBu)) {              // :142-145 -> byte stays a literal
                    lit_from = i;
                    c = H + 1u;
                    continue;
                }
                uint32_t nibble, adv;
                match_code(k, nibble, adv);                            // mlen[] (:44-45): 4..16 -> k-1; 17..31 -> 16; 32 / 48 / 64 -> nibble 0 / 1 / 2
                i += adv;                                              // :154 / :307
                e.match(room, nibble, i);                              // :152-159
                if (!(i < size) && !(i < size - 5u)) { done = true; break; }      // probe would be unobservable
                c = H + adv;
                chain_pending = true;
            }
        }
        if (done) break;

        // ---------------- leave the window: commit inserts (last writer per hash wins, :79)
        {
            const uint32_t mine = M & inP;
            if (((inP >> lane) & 1u) && (mine >> lane) == 1u) {
                if constexpr (FAT) { if (!(TSQB_ENC_DIAG & 2)) store_entry(table, h, x, own, epoch, a1, a2, a3, pol_st); atomicOr(&ws.written[h >> 7], 1u << ((h >> 2) & 31u)); }
                else table16[h] = (uint16_t)x;
            }
            __syncwarp();
        }
#if TSQB_ENC_FIXED_GRID
        // the next probe is position base + c (c >= 32) on every path out of the loop; windows stay on the grid 1 + 32k,
        // a window that a long match (extension format) jumps over is skipped
        base += c & ~31u;
        c0 = c & 31u;
#else
        base = chain_pending ? i : i + 1u;
#endif
    }

    return e.finish(flags);
}

template <bool FAT, bool EXT, bool STREAM = false>
__global__ void __launch_bounds__(kWarps * 32) encode_batch_kernel(EncodeArgs a)
{
    __shared__ WarpWs ws_all[kWarps];
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (slot >= a.n_slots) return;
    WarpWs& ws = ws_all[threadIdx.x >> 5];
    uint8_t* table = reinterpret_cast<uint8_t*>(a.tables) + (size_t)slot * (FAT ? kFatTableBytes : kTableBytes);
    uint64_t epoch = a.epoch;
    for (uint64_t b = slot; b < a.nb; b += a.n_slots) {
        if constexpr (FAT) {
            epoch++;                                                   // a fresh (empty) table: tsqInit (tsq_context.cpp:77-80)
            uint4* w4 = reinterpret_cast<uint4*>(ws.written);
            for (uint32_t q = lane; q < kBloomWords / 4u; q += 32u) w4[q] = make_uint4(0, 0, 0, 0);
            __syncwarp();
        }
        else {
            uint4* t4 = reinterpret_cast<uint4*>(table);
            for (uint32_t q = lane; q < kTableBytes / 16u; q += 32u) t4[q] = make_uint4(0, 0, 0, 0);
            __syncwarp();
        }
        const uint64_t at = b * (uint64_t)a.block;
        const uint32_t n = (uint32_t)((a.total - at < a.block) ? a.total - at : a.block);
        uint32_t flags;
        // (streamed input: what follows the launch's last block belongs to the next launch's blocks)
        const uint32_t c = encode_block_batch<FAT, EXT, STREAM>(table, epoch, a.in + at, n, a.slots + b * a.stride, lane, ws, flags, a.hints, a.arrived,
                                                                b + 1 == a.nb ? a.arrived_next : nullptr, a.stream_error);
        if (lane == 0) { a.sizes[b] = c; if (a.tailflags) a.tailflags[b] = flags; }
        __syncwarp();
    }
}

}  // namespace

// Warps (= blocks in flight) of the batch encoder that are RESIDENT at once on this device: registers (72 per thread)
// allow 7 CTAs of 4 warps per SM, not 8.  A launch with more slots than that runs its last CTAs as a second wave behind
// the first one's whole block list (8 GiB of text in 256 KiB blocks: 304 ms with 148 x 32 slots), so the slot count is
// capped here and the blocks are dealt evenly over slots that all run from the start.
uint32_t encode_batch_resident_warps(int sm_count)
{
    static int per_sm = 0;
    if (per_sm == 0) {
        int n = 0;
        // the most register-hungry instantiation decides (they differ by a CTA at most)
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, encode_batch_kernel<true, true, false>, kWarps * 32, 0) != cudaSuccess || n <= 0) {
            cudaGetLastError();
            n = 6;
        }
        int m = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&m, encode_batch_kernel<true, false, false>, kWarps * 32, 0) == cudaSuccess && m > 0 && m < n) n = m;
        per_sm = n * kWarps;
    }
    return (uint32_t)(per_sm * sm_count);
}

cudaError_t launch_encode_batch(const EncodeArgs& a, bool ext, cudaStream_t st)
{
    if (a.nb == 0) return cudaSuccess;
    const unsigned ctas = (a.n_slots + kWarps - 1) / kWarps;
    if (a.arrived) {                                                   // piece-streamed input: sector entries only
        if (!a.fat) return cudaErrorInvalidValue;
        if (ext) encode_batch_kernel<true, true, true><<<ctas, kWarps * 32, 0, st>>>(a);
        else     encode_batch_kernel<true, false, true><<<ctas, kWarps * 32, 0, st>>>(a);
    } else if (ext) {
        if (a.fat) encode_batch_kernel<true, true><<<ctas, kWarps * 32, 0, st>>>(a);
        else       encode_batch_kernel<false, true><<<ctas, kWarps * 32, 0, st>>>(a);
    } else {
        if (a.fat) encode_batch_kernel<true, false><<<ctas, kWarps * 32, 0, st>>>(a);
        else       encode_batch_kernel<false, false><<<ctas, kWarps * 32, 0, st>>>(a);
    }
    return cudaGetLastError();
}

}  // namespace tsqb
