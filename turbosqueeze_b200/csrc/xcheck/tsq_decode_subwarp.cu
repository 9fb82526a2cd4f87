// tsq_decode_subwarp.cu -- the sub-warp pair-step decoder (round-1 v1).  CROSS-CHECK ONLY: built into the test-only
// library tests/xcheck/libturbosqueeze_b200_xcheck.so (-DTSQB_XCHECK), not into the product; the production decoder is
// tsq_decode_split.cu.
//
// Semantics: reference tsqDecodeNoext (tsq_decode.cpp:42-126) and the extension variant
// (tsq_decode.cpp:137-314), restated around a "pair step": the reference's inner body handles one
// size byte and its two symbols; the two symbols of a pair never depend on each other because every
// match source lies before the start of its pair (tsq_encode.cpp:139-141), so a pair is the natural
// unit of lane parallelism.
//
// Mapping: W lanes (a sub-warp, W = 1..32) own one block.  All W lanes walk the token stream
// redundantly (uniform registers, broadcast loads), then lane t copies byte t (t+W, ...) of both
// symbols of the pair.  W is chosen from the number of independent blocks so that the grid fills
// 148 SMs: few big blocks -> W = 32 (one warp per block), many small blocks -> narrow sub-warps that
// amortise the serial token walk over several blocks per instruction.
//
// Unlike the reference, which copies a blind 16 bytes per symbol and over-writes up to ~100 bytes
// past the block (tsq_decode.cpp:60-123), every store is clipped at the block's decoded size so
// that blocks can be packed back to back in HBM.
#include "../tsq_device.cuh"

namespace tsqb {

template <int W>
__device__ __forceinline__ unsigned group_mask(unsigned lane)
{
    if constexpr (W == 32) return 0xffffffffu;
    else return ((1u << W) - 1u) << (lane & ~(unsigned)(W - 1));
}

template <int W, bool EXT>
__global__ void __launch_bounds__(128) decode_kernel(DecodeArgs a)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned sub  = lane & (unsigned)(W - 1);
    const unsigned mask = group_mask<W>(lane);
    const uint64_t ngroups = (uint64_t)gridDim.x * (blockDim.x / W);
    uint64_t b = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / W;

    for (; b < a.nb; b += ngroups) {
        const uint8_t* __restrict__ in = a.comp + (a.offs ? a.offs[b] : b * a.stride);
        uint8_t* o = a.out + b * a.ostride;
        // tsq_decode.cpp:49-53
        const uint32_t size  = (uint32_t)in[0] | ((uint32_t)in[1] << 8) | ((uint32_t)in[2] << 16);
        const uint32_t limit = a.csizes ? a.csizes[b] : (a.stride > 0xffffffffull ? 0xffffffffu : (uint32_t)a.stride);
        const bool ok = size <= kBlockMax && size <= a.ostride;
        if (sub == 0) a.osizes[b] = ok ? size : 0u;
        if (!ok) continue;

        uint32_t i = 3, j = 0;
        while (j < size && i < limit) {
            uint32_t ctl = in[i++];                                   // tsq_decode.cpp:62
#pragma unroll
            for (int p = 0; p < 4; p++) {
                if (j >= size) break;
                const uint32_t nib = in[i++];                         // :68
                const uint32_t org = j;                               // :69  rep_last_j
                uint32_t len0 = (nib >> 4) + 1u, len1 = (nib & 15u) + 1u;
                const bool lit0 = (ctl & 0x80u) != 0, lit1 = (ctl & 0x40u) != 0;
                ctl <<= 2;

                // ---- first symbol (:70-77)
                const uint8_t* s0;
                bool v0 = true;
                if (lit0) { s0 = in + i; i += len0; }
                else {
                    const uint32_t off = (uint32_t)in[i] | ((uint32_t)in[i + 1] << 8);
                    i += 2;
                    if (EXT && len0 <= 3u) len0 = 16u * (len0 + 1u);   // tsq_decode.cpp:174-187
                    v0 = off <= org;                                  // corrupt stream guard
                    s0 = o + (org - off);
                }
                const uint32_t d0 = j;
                j += len0;

                // ---- second symbol (:79-86); skipped once the block is complete so that a valid
                // stream is never read past its last real symbol
                const bool two = j < size;
                const uint8_t* s1 = s0;
                bool v1 = true;
                uint32_t d1 = j;
                if (two) {
                    if (lit1) { s1 = in + i; i += len1; }
                    else {
                        const uint32_t off = (uint32_t)in[i] | ((uint32_t)in[i + 1] << 8);
                        i += 2;
                        if (EXT && len1 <= 3u) len1 = 16u * (len1 + 1u);
                        v1 = off <= org;
                        s1 = o + (org - off);
                    }
                    j += len1;
                } else {
                    len1 = 0;
                }

                // ---- copy: lane t moves byte t of both symbols; loads first, then stores
                const uint32_t n0 = v0 ? min(len0, size - d0) : 0u;
                const uint32_t n1 = (two && v1) ? min(len1, size - d1) : 0u;
                if (!EXT && W >= 16) {
                    uint8_t x0 = 0, x1 = 0;
                    if (sub < n0) x0 = s0[sub];
                    if (sub < n1) x1 = s1[sub];
                    if (sub < n0) o[d0 + sub] = x0;
                    if (sub < n1) o[d1 + sub] = x1;
                } else {
                    const uint32_t nmax = max(n0, n1);
                    for (uint32_t t = sub; t < nmax; t += W) {
                        uint8_t x0 = 0, x1 = 0;
                        if (t < n0) x0 = s0[t];
                        if (t < n1) x1 = s1[t];
                        if (t < n0) o[d0 + t] = x0;
                        if (t < n1) o[d1 + t] = x1;
                    }
                }
                if (W > 1) __syncwarp(mask);                          // next pair may read these bytes
            }
        }
    }
}

template <int W, bool EXT>
static cudaError_t launch_decode_t(const DecodeArgs& a, int sm_count, cudaStream_t st)
{
    const int threads = 128;
    const uint64_t groups_per_cta = threads / W;
    uint64_t ctas = (a.nb + groups_per_cta - 1) / groups_per_cta;
    const uint64_t cap = (uint64_t)sm_count * 16u;                    // 16 CTAs x 4 warps = 64 warps / SM
    if (ctas > cap) ctas = cap;
    if (ctas == 0) return cudaSuccess;
    decode_kernel<W, EXT><<<(unsigned)ctas, threads, 0, st>>>(a);
    return cudaGetLastError();
}

// lanes = 1, 2, 4, 8, 16, 32 lanes per block
cudaError_t launch_decode_subwarp(const DecodeArgs& a, int lanes, bool ext, int sm_count, cudaStream_t st)
{
#define TSQB_CASE(Wv) case Wv: return ext ? launch_decode_t<Wv, true>(a, sm_count, st) : launch_decode_t<Wv, false>(a, sm_count, st);
    switch (lanes) {
        TSQB_CASE(1) TSQB_CASE(2) TSQB_CASE(4) TSQB_CASE(8) TSQB_CASE(16) TSQB_CASE(32)
        default: return cudaErrorInvalidValue;
    }
#undef TSQB_CASE
}

}  // namespace tsqb
