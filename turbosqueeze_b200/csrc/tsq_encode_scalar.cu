// tsq_encode_scalar.cu -- one THREAD per block: the plain serial statement of the encoder.
//
// This is the correctness anchor of the device path and the kernel used for the extension format
// and for workloads with very many small blocks (32 independent chains share every instruction).
// Semantics: reference tsqEncodeNoext (tsq_encode.cpp:48-189) / extension variant (:200-341),
// following the behavioural spec in SURVEY.md 8(a): literal scan with one probe+insert per byte
// (:70-100), literal flush in 16-byte chunks (:82-98,:103-118), match chain (:123-170), padding
// (:176-188).  Differences that do not change bytes [0, outputSize):
//   * control / size bytes are assembled in registers and stored once (Emitter);
//   * literals are stored with their exact length -- the reference's blind 16-byte store only
//     matters for the <= 2 trailing never-initialised bytes, reproduced in finish().
#include "tsq_encode_common.cuh"

namespace tsqb {

struct ScalarEmitter : Emitter {
    __device__ __forceinline__ void symbol(uint32_t is_lit, uint32_t nibble, uint32_t in_pos)
    {
        n++;
        ctl_acc = (ctl_acc << 1) | is_lit;
        if ((n & 7u) == 0) { out[ctl_at] = (uint8_t)ctl_acc; ctl_at = j++; }
        nib_acc = (nib_acc << 4) | nibble;
        if ((n & 1u) == 0) { out[nib_at] = (uint8_t)nib_acc; nib_at = j++; rep = in_pos; }
    }

    // tsq_encode.cpp:85-97 / :105-117
    __device__ __forceinline__ void literals(const uint8_t* in, uint32_t& from, uint32_t upto)
    {
        do {
            uint32_t cnt = upto - from;
            if (cnt > 16u) cnt = 16u;
            for (uint32_t t = 0; t < cnt; t++) out[j + t] = in[from + t];
            lit_js = j; lit_src = from;
            from += cnt; j += cnt;
            symbol(1u, cnt - 1u, from);
        } while (upto - from > 0);
    }
};

template <bool EXT>
__device__ uint32_t encode_block_scalar(uint16_t* __restrict__ table, const uint8_t* __restrict__ in,
                                        uint32_t size, uint8_t* __restrict__ out, uint32_t& flags)
{
    constexpr uint32_t cap = EXT ? 64u : 16u;
    ScalarEmitter e;
    e.begin(out);
    out[0] = (uint8_t)size; out[1] = (uint8_t)(size >> 8); out[2] = (uint8_t)(size >> 16);

    uint32_t i = 0, lit_from, word, pos, off;
    do {
        lit_from = i;
        do {                                                       // literal scan :70-100
            i++;
            word = ld_le32(in + i);
            const uint32_t h = hash17(word);
            pos = expand_pos(table[h], i);
            table[h] = (uint16_t)i;
            off = e.rep - pos;                                     // not refreshed by the flush below
            if (i - lit_from > 31u) e.literals(in, lit_from, i);
        } while (i < size && !(word == ld_le32(in + pos) && (off - 4u) < 0xFFFBu));

        if (i - lit_from > 0) e.literals(in, lit_from, i);         // :103-118
        if (!(i < size)) break;

        do {                                                       // match chain :123-170
            uint32_t k = 0;
            while (k < cap && in[i + k] == in[pos + k]) k++;       // :126-137 / :276-290
            const uint32_t room = e.rep - pos;
            if (k > room) k = room - 1u;                           // :139-141
            if (k < 4u) break;
            off = e.rep - pos;
            if (!((off - 4u) < 0xFFFBu)) break;                    // :144-145
            uint32_t nibble, adv;
            match_code(k, nibble, adv);
            out[e.j] = (uint8_t)off; out[e.j + 1] = (uint8_t)(off >> 8);
            e.j += 2;
            i += adv;
            e.symbol(0u, nibble, i);
            word = ld_le32(in + i);                                // :162-167
            const uint32_t h = hash17(word);
            pos = expand_pos(table[h], i);
            table[h] = (uint16_t)i;
            off = e.rep - pos;
        } while (i < size - 5u && word == ld_le32(in + pos) && (off - 4u) < 0xFFFBu);
    } while (i < size);

    flags = e.finish(in);
    return e.j;
}

template <bool EXT>
__global__ void __launch_bounds__(32) encode_scalar_kernel(EncodeArgs a)
{
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= a.n_slots) return;
    uint16_t* table = a.tables + (size_t)slot * kHashSlots;
    for (uint64_t b = slot; b < a.nb; b += a.n_slots) {
        uint4* t4 = reinterpret_cast<uint4*>(table);               // tsqInit (tsq_context.cpp:77-80)
        for (uint32_t q = 0; q < kTableBytes / 16u; q++) t4[q] = make_uint4(0, 0, 0, 0);
        const uint64_t at = b * (uint64_t)a.block;
        const uint32_t n = (uint32_t)((a.total - at < a.block) ? a.total - at : a.block);
        uint32_t flags;
        a.sizes[b] = encode_block_scalar<EXT>(table, a.in + at, n, a.slots + b * a.stride, flags);
        if (a.tailflags) a.tailflags[b] = flags;
    }
}

cudaError_t launch_encode_scalar(const EncodeArgs& a, bool ext, cudaStream_t st)
{
    if (a.nb == 0) return cudaSuccess;
    const unsigned ctas = (a.n_slots + 31u) / 32u;
    if (ext) encode_scalar_kernel<true><<<ctas, 32, 0, st>>>(a);
    else     encode_scalar_kernel<false><<<ctas, 32, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace tsqb
