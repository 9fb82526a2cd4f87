// tsq_encode_common.cuh -- pieces shared by the scalar and the warp-per-block encoders.
#pragma once
#include "tsq_device.cuh"

namespace tsqb {

// Unaligned little-endian 32-bit read built from two aligned words (reference reads unaligned
// u32/u64 directly: tsq_encode.cpp:74,126-128).  Touches at most the aligned word after p+3.
__device__ __forceinline__ uint32_t ld_le32(const uint8_t* p)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const uint32_t lo = w[0], hi = w[1];
    return __funnelshift_r(lo, hi, (uint32_t)(a & 3u) * 8u);
}

// hash of the 4 bytes at a position (tsq_encode.cpp:75)
__device__ __forceinline__ uint32_t hash17(uint32_t w) { return (w ^ (w >> 12)) & kHashMask; }

// 16-bit table entry -> absolute candidate position in the 64 KiB behind i (tsq_encode.cpp:77-78)
__device__ __forceinline__ uint32_t expand_pos(uint32_t entry, uint32_t i)
{
    const uint32_t base = i & 0xFFFF0000u;
    return entry + ((entry >= (i & 0xFFFFu)) ? base - 65536u : base);
}

// match length -> (size nibble, input advance).  tsq_encode.cpp:44-45 (mlen) and :154 / :307.
__device__ __forceinline__ void match_code(uint32_t k, uint32_t& nibble, uint32_t& adv)
{
    if (k < 17u)      { nibble = k - 1u; adv = k; }
    else if (k < 32u) { nibble = 15u;    adv = 16u; }
    else if (k < 48u) { nibble = 0u;     adv = 32u; }
    else if (k < 64u) { nibble = 1u;     adv = 48u; }
    else              { nibble = 2u;     adv = 64u; }
}

// finish() flags: which trailing bytes are a function of the output buffer's previous contents
constexpr uint32_t kTailCtlPrefill = 1u;   // byte j-2 left as pre-filled
constexpr uint32_t kTailNibPrefill = 2u;   // byte j-1 left as pre-filled
constexpr uint32_t kTailNibShifted = 4u;   // byte j-1 = (pre-fill << 4)

// Token-stream writer state.  The reference read-modify-writes the open control / size byte in
// the output buffer for every symbol (tsq_encode.cpp:94-95,158-159); here both live in registers
// and are stored once when full.  `lit_js/lit_src` remember the last literal chunk so that the
// trailing never-initialised bytes can be reproduced (see finish()).
struct Emitter {
    uint8_t* out;
    uint32_t j;         // next free output byte
    uint32_t ctl_at;    // address of the open control byte
    uint32_t nib_at;    // address of the open size byte
    uint32_t n;         // symbols so far
    uint32_t rep;       // input position at the start of the open pair (rep_last_i)
    uint32_t ctl_acc;
    uint32_t nib_acc;
    uint32_t lit_js;    // output address of the last literal chunk (0x80000000: none yet)
    uint32_t lit_src;   // input position that chunk was copied from

    __device__ __forceinline__ void begin(uint8_t* o)
    {
        out = o; j = 5; ctl_at = 3; nib_at = 4; n = 0; rep = 0; ctl_acc = 0; nib_acc = 0;
        lit_js = 0x80000000u; lit_src = 0;
    }

    // What the reference's memory holds at an address it allocated but never initialised: the
    // spill of its last blind 16-byte literal store (tsq_encode.cpp:88,108) if that covered the
    // address, otherwise whatever the output buffer held before the call.
    __device__ __forceinline__ bool stale(const uint8_t* __restrict__ in, uint32_t a, uint32_t& v) const
    {
        if (a - lit_js < 16u) { v = in[lit_src + (a - lit_js)]; return true; }
        return false;
    }

    // Padding, tsq_encode.cpp:176-188; called by ONE thread.  n % 8 == 0: the freshly opened
    // control and size byte are emitted untouched; n even otherwise: the open size byte is
    // (stale << 4); n odd: its one real nibble moves up.  Returns the kTail* flags describing
    // which of the last two bytes still depend on the caller's pre-fill of the output buffer.
    __device__ __forceinline__ uint32_t finish(const uint8_t* __restrict__ in)
    {
        const uint32_t r = n & 7u;
        uint32_t v, flags = 0;
        if (r == 0) {
            if (stale(in, ctl_at, v)) out[ctl_at] = (uint8_t)v; else flags |= kTailCtlPrefill;
            if (stale(in, nib_at, v)) out[nib_at] = (uint8_t)v; else flags |= kTailNibPrefill;
        } else {
            out[ctl_at] = (uint8_t)((ctl_acc << (8u - r)) | ((1u << (8u - r)) - 1u));
            if (n & 1u) out[nib_at] = (uint8_t)(nib_acc << 4);
            else {
                if (!stale(in, nib_at, v)) { v = out[nib_at]; flags |= kTailNibShifted; }
                out[nib_at] = (uint8_t)(v << 4);
            }
        }
        return flags;
    }
};

}  // namespace tsqb
