// tsq_container.cu -- TSQ1 framing on the device and the encode / decode dispatchers.
//
// Container (reference turbosqueeze.cpp:64-67,78-84; tsq_threads.cpp:333-335,218-240):
//   "TSQ1" | n_blocks u32 | total_uncompressed u64 | n_blocks x { u24 (size | 0x800000 if ext), stream }
#include "tsq_device.cuh"

namespace tsqb {

cudaError_t launch_encode_scalar(const EncodeArgs& a, bool ext, cudaStream_t st);
cudaError_t launch_encode_batch(const EncodeArgs& a, bool ext, cudaStream_t st);
uint32_t    encode_batch_resident_warps(int sm_count);
cudaError_t launch_decode_split(const DecodeArgs& a, bool ext, int sm_count, cudaStream_t st, bool lane_per_pair, int slot_cap);
#ifdef TSQB_XCHECK   // round-1 v1 kernels, kept as cross-checks in the test-only library (csrc/xcheck/)
cudaError_t launch_encode_warp(const EncodeArgs& a, cudaStream_t st);
cudaError_t launch_decode_warp(const DecodeArgs& a, int sm_count, cudaStream_t st);
cudaError_t launch_decode_subwarp(const DecodeArgs& a, int lanes, bool ext, int sm_count, cudaStream_t st);
#endif

uint32_t encode_slots_for(int impl, uint64_t nb, int sm_count, int64_t user_override)
{
    uint64_t slots;
    if (user_override > 0) slots = (uint64_t)user_override;
    else if (impl == 1)    slots = (uint64_t)sm_count * 64u;          // threads: 2 CTAs of 32 per SM
    else if (impl == 3)    slots = encode_batch_resident_warps(sm_count);   // every slot resident from the start (28 warps per SM)
    else                   slots = (uint64_t)sm_count * 32u;          // warps: 32 per SM
    if (slots > nb) slots = nb;
    if (slots == 0) slots = 1;
    return (uint32_t)slots;
}

uint32_t encode_table_bytes(int impl, bool fat) { return impl == 3 && fat ? kFatTableBytes : kTableBytes; }

// 384 tables x (256 KiB + a 64 KiB back-window) = 120 MB: what the 126 MB L2 can keep resident
bool encode_wants_fat(int impl, uint32_t n_slots) { return impl == 3 && n_slots > 384u; }

cudaError_t launch_encode(const EncodeArgs& a, int impl, bool ext, int /*sm_count*/, cudaStream_t st)
{
    if (impl == 1) return launch_encode_scalar(a, ext, st);
#ifdef TSQB_XCHECK
    if (impl == 2) return ext ? cudaErrorInvalidValue : launch_encode_warp(a, st);
#else
    if (impl == 2) return cudaErrorInvalidValue;                     // the v1 warp kernel lives in the cross-check library only
#endif
    return launch_encode_batch(a, ext, st);
}

// lanes: 0 = auto (35 for the no-extension format, 34 for the extension format);
// 34 = walker + copier kernel (tsq_decode_split.cu), lane per symbol, both formats;
// 35 = the same kernel choosing per block between the lane-per-pair copier (64 symbols per step) and, for (nearly)
//      incompressible blocks, the lane-per-symbol one; no-extension format (the extension format runs as 34).
// Cross-check library only: 1..32 = sub-warp pair-step kernel with that many lanes per block, 33 = v1 warp-per-block
// step kernel (no-extension format).
cudaError_t launch_decode(const DecodeArgs& a, int lanes, bool ext, int sm_count, cudaStream_t st, int slot_cap)
{
    if (lanes <= 0) lanes = ext ? 34 : 35;
    if (lanes == 34 || lanes == 35) return launch_decode_split(a, ext, sm_count, st, lanes == 35, slot_cap);
#ifdef TSQB_XCHECK
    if (lanes == 33) return ext ? cudaErrorInvalidValue : launch_decode_warp(a, sm_count, st);
    return launch_decode_subwarp(a, lanes, ext, sm_count, st);
#else
    return cudaErrorInvalidValue;
#endif
}

// ---- pack: offsets = exclusive scan of (3 + size), one CTA (n_blocks is at most a few million)
__global__ void __launch_bounds__(1024) pack_scan_kernel(const uint32_t* __restrict__ sizes, uint64_t nb,
                                                         uint64_t total_u, uint8_t* __restrict__ container,
                                                         uint64_t* __restrict__ offsets, uint64_t* __restrict__ total_out)
{
    __shared__ uint64_t part[1024];
    const unsigned t = threadIdx.x;
    const uint64_t per = (nb + 1023u) / 1024u;
    const uint64_t lo = min(nb, t * per), hi = min(nb, lo + per);
    uint64_t sum = 0;
    for (uint64_t b = lo; b < hi; b++) sum += 3u + (uint64_t)sizes[b];
    part[t] = sum;
    __syncthreads();
    if (t == 0) {
        uint64_t acc = 16;                                           // header
        for (int q = 0; q < 1024; q++) { const uint64_t v = part[q]; part[q] = acc; acc += v; }
        *total_out = acc;
        const uint8_t magic[4] = {'T', 'S', 'Q', '1'};
        for (int q = 0; q < 4; q++) container[q] = magic[q];
        for (int q = 0; q < 4; q++) container[4 + q] = (uint8_t)((uint32_t)nb >> (8 * q));
        for (int q = 0; q < 8; q++) container[8 + q] = (uint8_t)(total_u >> (8 * q));
    }
    __syncthreads();
    uint64_t acc = part[t];
    for (uint64_t b = lo; b < hi; b++) { offsets[b] = acc; acc += 3u + (uint64_t)sizes[b]; }
}

__global__ void __launch_bounds__(256) pack_copy_kernel(const uint8_t* __restrict__ slots, uint64_t stride,
                                                        const uint32_t* __restrict__ sizes, uint64_t nb, uint32_t ext,
                                                        uint8_t* __restrict__ container, const uint64_t* __restrict__ offsets)
{
    for (uint64_t b = blockIdx.x; b < nb; b += gridDim.x) {
        const uint8_t* src = slots + b * stride;
        uint8_t* dst = container + offsets[b];
        const uint32_t n = sizes[b];
        if (threadIdx.x < 3) dst[threadIdx.x] = (uint8_t)((n | (ext ? 0x800000u : 0u)) >> (8 * threadIdx.x));
        dst += 3;
        // head bytes until dst is 16-byte aligned, then 16-byte stores fed by byte-realigned loads
        const uint32_t head = min(n, (uint32_t)((16u - ((uintptr_t)dst & 15u)) & 15u));
        if (threadIdx.x < head) dst[threadIdx.x] = src[threadIdx.x];
        const uint32_t body = (n - head) & ~15u;
        const uint8_t* s = src + head;
        uint4* d4 = reinterpret_cast<uint4*>(dst + head);
        const uint32_t sh = (uint32_t)((uintptr_t)s & 3u) * 8u;
        const uint32_t* sw = reinterpret_cast<const uint32_t*>((uintptr_t)s & ~(uintptr_t)3);
        for (uint32_t q = threadIdx.x; q < body / 16u; q += blockDim.x) {
            const uint32_t* p = sw + q * 4u;
            const uint32_t a0 = p[0], a1 = p[1], a2 = p[2], a3 = p[3], a4 = p[4];
            d4[q] = make_uint4(__funnelshift_r(a0, a1, sh), __funnelshift_r(a1, a2, sh),
                               __funnelshift_r(a2, a3, sh), __funnelshift_r(a3, a4, sh));
        }
        const uint32_t tail0 = head + body;
        if (tail0 + threadIdx.x < n) dst[tail0 + threadIdx.x] = src[tail0 + threadIdx.x];
    }
}

cudaError_t launch_pack(const uint8_t* slots, uint64_t stride, const uint32_t* sizes, uint64_t nb, uint64_t total_u,
                        uint32_t ext, uint8_t* container, uint64_t* total_out, uint64_t* scratch_offsets,
                        cudaStream_t st)
{
    pack_scan_kernel<<<1, 1024, 0, st>>>(sizes, nb, total_u, container, scratch_offsets, total_out);
    if (nb) {
        const unsigned ctas = (unsigned)(nb < 148u * 8u ? nb : 148u * 8u);
        pack_copy_kernel<<<ctas, 256, 0, st>>>(slots, stride, sizes, nb, ext, container, scratch_offsets);
    }
    return cudaGetLastError();
}

// ---- index: the u24 chain can only be walked serially (tsq_threads.cpp:480-484,513-524)
__global__ void index_kernel(const uint8_t* __restrict__ c, uint64_t csize, uint64_t max_blocks,
                             uint64_t* __restrict__ offs, uint32_t* __restrict__ sizes, uint32_t* __restrict__ ext,
                             uint64_t* __restrict__ n_blocks)
{
    uint64_t at = 16, n = 0;
    if (csize < 16 || c[0] != 'T' || c[1] != 'S' || c[2] != 'Q' || c[3] != '1') { *n_blocks = 0; return; }
    while (at + 3 <= csize && n < max_blocks) {
        const uint32_t v = (uint32_t)c[at] | ((uint32_t)c[at + 1] << 8) | ((uint32_t)c[at + 2] << 16);
        const uint32_t len = v & 0x7FFFFFu;
        at += 3;
        if (len < 3 || at + len > csize) break;                      // turbosqueeze.cpp:136; a block is at least its u24 size header
        offs[n] = at; sizes[n] = len; ext[n] = v >> 23;
        n++;
        at += len;
    }
    *n_blocks = n;
}

cudaError_t launch_index(const uint8_t* container, uint64_t csize, uint64_t max_blocks, uint64_t* offs,
                         uint32_t* sizes, uint32_t* ext_flags, uint64_t* n_blocks, cudaStream_t st)
{
    index_kernel<<<1, 1, 0, st>>>(container, csize, max_blocks, offs, sizes, ext_flags, n_blocks);
    return cudaGetLastError();
}

}  // namespace tsqb
