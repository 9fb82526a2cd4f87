// tsq_device.cuh -- launch interface between the C-ABI (tsq_capi.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tsqb {

constexpr uint32_t kBlockMax   = 1u << 22;          // TSQ_BLOCK_SZ   (reference turbosqueeze.h:38)
constexpr uint32_t kHashSlots  = 1u << 17;          // 2^TSQ_HASH_BITS (reference turbosqueeze.h:41)
constexpr uint32_t kHashMask   = kHashSlots - 1u;
constexpr uint32_t kTableBytes = kHashSlots * 2u;   // TSQ_HASH_SZ
constexpr uint32_t kFatTableBytes = kHashSlots * 32u;  // one 32-byte sector per entry: batch encoder (tsq_encode_batch.cu)

struct EncodeArgs {
    const uint8_t* in;        // contiguous input, `total` bytes + >= 128 readable
    uint64_t       total;
    uint32_t       block;     // block size
    uint64_t       nb;        // number of blocks
    uint8_t*       slots;     // block b -> slots + b * stride
    uint64_t       stride;
    uint32_t*      sizes;     // compressed size per block
    uint32_t*      tailflags; // optional: kTail* flags per block (see tsq_encode_common.cuh)
    uint16_t*      tables;    // n_slots tables (kTableBytes each; kFatTableBytes for the batch encoder)
    uint64_t       epoch;     // batch encoder: entries written under another epoch are empty (no per-block zeroing)
    uint32_t       fat;       // batch encoder: 1 = sector entries (kFatTableBytes per table), 0 = u16 tables
    uint32_t       n_slots;
    // Piece-streamed input (host path, tsq_capi.cu compress_streamed): the bytes of the blocks arrive over PCIe WHILE the kernel
    // runs, the same prefix of every block at a time.  *arrived = bytes of each block's prefix that have landed (written by a
    // stream-ordered copy behind the data); *arrived_next = the same for the blocks that follow this launch's last block, whose
    // first bytes that block reads (nullptr: nothing follows but the zero pad); *stream_error is set when a wait times out.
    const uint32_t* arrived = nullptr;
    const uint32_t* arrived_next = nullptr;
    uint32_t*      stream_error = nullptr;
    uint32_t       hints = 0; // batch encoder experiments (only in a build with -DTSQB_ENC_HINTS=1; ignored otherwise):
                              // 1 = table traffic evict-first in L2, 2 = streaming output stores, 4 / 8 = L2 prefetches
};

struct DecodeArgs {
    const uint8_t*  comp;
    const uint64_t* offs;     // optional explicit stream offsets
    uint64_t        stride;
    const uint32_t* csizes;   // optional readable bytes per stream
    uint64_t        nb;
    uint8_t*        out;      // block b -> out + b * ostride
    uint64_t        ostride;
    uint32_t*       osizes;
};

// encode_impl: 1 = scalar (one thread per block), 3 = warp per block with token batches (tsq_encode_batch.cu, the default,
// both formats); 2 = round-1 v1 warp kernel, cross-check library only (-DTSQB_XCHECK).
cudaError_t launch_encode(const EncodeArgs& a, int impl, bool ext, int sm_count, cudaStream_t st);
// bytes of one hash table of launch_encode(impl)
uint32_t    encode_table_bytes(int impl, bool fat);
// batch encoder: sector entries pay off once the tables no longer fit the L2
bool        encode_wants_fat(int impl, uint32_t n_slots);
// how many hash tables launch_encode(impl) will use for nb blocks (caller sizes a.tables from it)
uint32_t    encode_slots_for(int impl, uint64_t nb, int sm_count, int64_t user_override);
// batch encoder: warps that are resident at once on the device (registers: 7 CTAs of 4 warps per SM)
uint32_t    encode_batch_resident_warps(int sm_count);

// lanes: see tsq_container.cu (34 / 35 = walker + copier kernel; 1..33 only in the cross-check library)
cudaError_t launch_decode(const DecodeArgs& a, int lanes, bool ext, int sm_count, cudaStream_t st, int slot_cap = 0);

cudaError_t launch_pack(const uint8_t* slots, uint64_t stride, const uint32_t* sizes, uint64_t nb,
                        uint64_t total_u, uint32_t ext, uint8_t* container, uint64_t* total_out,
                        uint64_t* scratch_offsets, cudaStream_t st);
cudaError_t launch_index(const uint8_t* container, uint64_t csize, uint64_t max_blocks, uint64_t* offs,
                         uint32_t* sizes, uint32_t* ext_flags, uint64_t* n_blocks, cudaStream_t st);

}  // namespace tsqb
