// tsq_encode_warp.cu (CROSS-CHECK ONLY: test library tests/xcheck/libturbosqueeze_b200_xcheck.so, not the product) -- one WARP per block: the greedy parse with 32 positions probed at a time.
//
// Reference semantics: tsqEncodeNoext (tsq_encode.cpp:48-189); bit-exact, see SURVEY.md 8(a).
//
// The reference parse is a serial chain: every probe reads a table entry written by an earlier
// probe, and which positions are inserted depends on every earlier decision (positions inside a
// match are skipped, tsq_encode.cpp:154).  What can run in parallel is the literal scan: between
// two hits EVERY position is probed and inserted in order (:70-79), so a warp evaluates a WINDOW of
// 32 consecutive positions at once:
//
//   lane L <-> position x = base + L
//     w   = LE32(in + x), h = hash(w)                      (:74-75)
//     s   = table[h]   as committed before this window      (:76)
//     M   = lanes of the window with the same hash          (__match_any_sync)
//   candidate of lane L = nearest earlier lane of M that is "in P" (probed => inserted), else the
//   table entry.  `inP` holds the lanes already known to be probed; during a scan from lane c every
//   lane >= c is in P exactly when the scan reaches it, hence the effective set inP | lanes>=c.
//   The first hit is a ballot + ffs; inserts are committed to the table when the window is left
//   (highest in-P lane per hash wins, which is "last writer wins" of :79).
//
// The match chain (:123-170) stays serial, but its probes reuse the window: after a match of
// length k from lane H the next probe is lane H+k of the same window (lanes in between are not in P).
// Match length is a 16-lane byte compare + ballot.  The token stream writer keeps the open control
// and size bytes in (warp-uniform) registers.
#include "../tsq_encode_common.cuh"

namespace tsqb {

constexpr unsigned FULL = 0xffffffffu;

struct WarpEmitter : Emitter {
    unsigned lane;

    __device__ __forceinline__ void symbol(uint32_t is_lit, uint32_t nibble, uint32_t in_pos)
    {
        n++;
        ctl_acc = (ctl_acc << 1) | is_lit;
        if ((n & 7u) == 0) { if (lane == 0) out[ctl_at] = (uint8_t)ctl_acc; ctl_at = j++; }
        nib_acc = (nib_acc << 4) | nibble;
        if ((n & 1u) == 0) { if (lane == 1) out[nib_at] = (uint8_t)nib_acc; nib_at = j++; rep = in_pos; }
    }

    // tsq_encode.cpp:85-97 / :105-117 -- lanes 0..15 move one byte each per chunk
    __device__ __forceinline__ void literals(const uint8_t* __restrict__ in, uint32_t& from, uint32_t upto)
    {
        do {
            uint32_t cnt = upto - from;
            if (cnt > 16u) cnt = 16u;
            if (lane < cnt) out[j + lane] = __ldg(in + from + lane);
            lit_js = j; lit_src = from;
            from += cnt; j += cnt;
            symbol(1u, cnt - 1u, from);
        } while (upto - from > 0);
    }
};

__device__ __forceinline__ uint32_t lanes_from_to(uint32_t lo, uint32_t hi)   // bits lo..hi inclusive
{
    return ((2u << hi) - 1u) & ~((1u << lo) - 1u);
}

__device__ uint32_t encode_block_warp(uint16_t* __restrict__ table, const uint8_t* __restrict__ in,
                                      const uint32_t size, uint8_t* __restrict__ out, const unsigned lane, uint32_t& flags)
{
    WarpEmitter e;
    e.begin(out);
    e.lane = lane;
    if (lane < 3) out[lane] = (uint8_t)(size >> (8u * lane));        // tsq_encode.cpp:53-55

    const uint32_t lt = (1u << lane) - 1u;
    uint32_t i = 0, lit_from = 0;
    uint32_t base = 1;                // first probe is position 1 (:70-72)
    bool chain_pending = false;       // lane 0 of the next window is a post-match probe (:162-170)

    for (;;) {                                                         // one window per iteration
        // ---------------- window precompute (parallel over 32 positions)
        const uint32_t x = base + lane;
        const uint32_t w = ld_le32(in + x);
        const uint32_t h = hash17(w);
        const uint32_t s = table[h];
        const uint32_t M = __match_any_sync(FULL, h);
        const uint32_t tab_cand = expand_pos(s, x);
        const bool teq = ld_le32(in + tab_cand) == w;
        const bool anydup = __any_sync(FULL, M != (1u << lane));
        uint32_t inP = 0, c = 0;
        bool done = false;

        while (c < 32u) {
            // ------------ effective candidate of every lane given the lanes in P
            uint32_t cand = tab_cand;
            bool weq = teq;
            if (anydup) {
                const uint32_t cm = M & lt & (inP | ~((1u << c) - 1u));
                const uint32_t q = cm ? 31u - (uint32_t)__clz(cm) : lane;
                const uint32_t wq = __shfl_sync(FULL, w, q);
                if (cm) { cand = base + q; weq = wq == w; }
            }

            uint32_t H, pos;
            if (chain_pending) {
                // lane c is the probe that follows a match (:162-170); i == base + c
                chain_pending = false;
                inP |= 1u << c;
                pos = __shfl_sync(FULL, cand, c);
                const bool eq = (__ballot_sync(FULL, weq) >> c) & 1u;
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
                const uint32_t F = lit_from + 32u - base;              // lane of the forced flush (:80-98)
                const uint32_t E = size - base;                        // lane with x == size
                const uint32_t hi = min(31u, min(F, E));
                const bool lanehit = weq && ((e.rep - cand - 4u) < 0xFFFBu) && lane >= c && lane <= hi && x < size;
                const uint32_t hm = __ballot_sync(FULL, lanehit);
                if (hm == 0) {
                    if (E <= hi) {                                     // ran into the end of the block
                        i = size;
                        if (i - lit_from > 31u) e.literals(in, lit_from, i);
                        if (i - lit_from > 0u) e.literals(in, lit_from, i);
                        done = true;
                        break;
                    }
                    inP |= lanes_from_to(c, hi);
                    i = base + hi;
                    if (F <= hi) e.literals(in, lit_from, i);          // 32 pending literals: rep moves, scan goes on
                    c = hi + 1u;
                    continue;
                }
                H = (uint32_t)__ffs((int)hm) - 1u;
                inP |= lanes_from_to(c, H);
                i = base + H;
                if (i - lit_from > 31u) e.literals(in, lit_from, i);   // flush precedes the loop test (:80-100)
                if (i - lit_from > 0u) e.literals(in, lit_from, i);    // :103-118
                pos = __shfl_sync(FULL, cand, H);
            }

            // ---------------- one match attempt at i == base + H against pos (:126-160)
            {
                const uint32_t t = lane & 15u;
                const bool ne = __ldg(in + i + t) != __ldg(in + pos + t);
                const uint32_t nm = __ballot_sync(FULL, ne) | 0xFFFF0000u;
                uint32_t k = (uint32_t)__ffs((int)nm) - 1u;           // common prefix, capped at 16
                const uint32_t room = e.rep - pos;
                if (k > room) k = room - 1u;                           // :139-141
                if (k < 4u || !((room - 4u) < 0xFFFBu)) {              // :142-145 -> byte stays a literal
                    lit_from = i;
                    c = H + 1u;
                    continue;
                }
                if (lane < 2) out[e.j + lane] = (uint8_t)(room >> (8u * lane));   // :152-153
                e.j += 2;
                i += k;                                                // :154
                e.symbol(0u, k - 1u, i);                               // :157-159 (mlen[k] = k-1, 16 -> 15)
                if (!(i < size) && !(i < size - 5u)) { done = true; break; }      // probe would be unobservable
                c = H + k;
                chain_pending = true;
            }
        }
        if (done) break;

        // ---------------- leave the window: commit inserts (last writer per hash wins, :79)
        {
            const uint32_t mine = M & inP;
            if (((inP >> lane) & 1u) && (mine >> lane) == 1u) table[h] = (uint16_t)x;
            __syncwarp();
        }
        base = chain_pending ? i : i + 1u;
    }

    flags = lane == 0 ? e.finish(in) : 0u;
    return e.j;
}

__global__ void __launch_bounds__(128) encode_warp_kernel(EncodeArgs a)
{
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (slot >= a.n_slots) return;
    uint16_t* table = a.tables + (size_t)slot * kHashSlots;
    for (uint64_t b = slot; b < a.nb; b += a.n_slots) {
        uint4* t4 = reinterpret_cast<uint4*>(table);                   // tsqInit (tsq_context.cpp:77-80)
        for (uint32_t q = lane; q < kTableBytes / 16u; q += 32u) t4[q] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        const uint64_t at = b * (uint64_t)a.block;
        const uint32_t n = (uint32_t)((a.total - at < a.block) ? a.total - at : a.block);
        uint32_t flags;
        const uint32_t c = encode_block_warp(table, a.in + at, n, a.slots + b * a.stride, lane, flags);
        if (lane == 0) { a.sizes[b] = c; if (a.tailflags) a.tailflags[b] = flags; }
        __syncwarp();
    }
}

cudaError_t launch_encode_warp(const EncodeArgs& a, cudaStream_t st)
{
    if (a.nb == 0) return cudaSuccess;
    const unsigned warps_per_cta = 4;
    const unsigned ctas = (a.n_slots + warps_per_cta - 1) / warps_per_cta;
    encode_warp_kernel<<<ctas, warps_per_cta * 32, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace tsqb
