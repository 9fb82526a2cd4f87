/*
 * tsq_b200.h -- C-ABI of libturbosqueeze_b200.so: the B200 (sm_100a) implementation of
 * Turbosqueeze's per-block encode/decode hot path.
 *
 * Two layers, both `extern "C"`, plain pointers and sizes only:
 *
 *   1. tsqb_*  -- the device-resident batch path (what the roofline is measured on): encode /
 *      decode n independent blocks that already live in HBM, on a caller-supplied CUDA stream.
 *      Not in the reference (it has no device), kept as thin as possible.
 *
 *   2. tsq*    -- the reference's own entry points (reference turbosqueeze.h:458-670), same names,
 *      argument meaning and error behaviour, taking HOST buffers / FILE*: each one stages H2D,
 *      launches the kernels of layer 1 and copies the result back.  A caller of the reference
 *      relinks against this library for that path and nothing else changes.
 *
 * There is no CPU fallback anywhere: every entry point fails (status != 0, message in
 * tsqb_last_error(); the void reference-style functions report *outputSize = 0 and print to
 * stderr) when no CUDA device is usable.
 *
 * Parity contract (identical to the reference's memory path,
 * tsq_threads.cpp:109,176-177):
 *   - blocks are sub-ranges of ONE contiguous input buffer; the encoder of block b reads up to 19
 *     bytes past the block (tsq_encode.cpp:74,126-128), i.e. into block b+1, and the buffer must be
 *     followed by >= TSQB_INPUT_PAD readable bytes (zeros for bit-exactness with the oracle);
 *   - the hash table is defined to be all-zero at the start of every block (every reference caller
 *     runs tsqInit first: turbosqueeze.cpp:75, tsq_threads.cpp:176, test/test.cpp:42);
 *   - output bytes [0, size) of every block equal the reference's for a zero-filled output slot.
 *     The kernel never stores past `size`; the up-to-two trailing control/size bytes the reference
 *     leaves uninitialised (tsq_encode.cpp:176-188) are reproduced: when the reference would leak the
 *     spill of its last 16-byte literal store the same input bytes are written, otherwise the byte
 *     is left as the caller pre-filled it.
 */
#ifndef TSQ_B200_H
#define TSQ_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference turbosqueeze.h:37-43 */
#define TSQB_BLOCK_MAX   (1u << 22)               /* TSQ_BLOCK_SZ  */
#define TSQB_OUTPUT_MAX  ((1u << 22) + (1u << 20)) /* TSQ_OUTPUT_SZ */
#define TSQB_HASH_BITS   17                        /* TSQ_HASH_BITS */
#define TSQB_HASH_BYTES  ((1u << 17) * 2u)         /* TSQ_HASH_SZ   */
#define TSQB_INPUT_PAD   128                       /* readable bytes required after an input buffer */

/* ------------------------------------------------------------------------------------------------
 * Layer 1: device-resident batch path
 * ---------------------------------------------------------------------------------------------- */

typedef struct tsqb_context tsqb_context;          /* opaque: device ordinal + hash-table scratch */

/* Last error of the calling thread ("" when none). */
const char* tsqb_last_error(void);

/* Kernels launched by this library so far in this process (bench.py's gpu_launches). */
uint64_t tsqb_launch_count(void);

/* Number of usable CUDA devices (0 when there is none; never falls back to the CPU). */
int tsqb_device_count(void);

/* Create / destroy a context bound to CUDA device `device`.  Returns 0 on success. */
int  tsqb_create(tsqb_context** ctx, int device);
void tsqb_destroy(tsqb_context* ctx);

/* Worst-case compressed size of a block of `block_size` bytes plus store slack, rounded up to 128:
 * all-literal stream = 3 + size + ceil(size/16)*(1 + 1/8 + 1/2) (tsq_encode.cpp:53-61,88-95). */
uint64_t tsqb_slot_stride(uint32_t block_size);

/* Knobs for benchmarking / tests.  Returns 0 when the key is known.
 *   "encode_impl"   0 = auto (token-batch kernel), 1 = scalar thread-per-block, 3 = warp-per-block token batches
 *                   (tsq_encode_batch.cu); 2 = round-1 v1 warp kernel: only in the test-only cross-check library
 *   "encode_slots"  0 = auto: hash tables (= blocks) in flight
 *   "encode_fat"    table format of the batch encoder: -1 = auto, 0 = 2^17 x u16, 1 = 32-byte sector entries
 *   "decode_lanes"  0 = auto (35, or 34 for the extension format); 35 = walker + copier kernel (tsq_decode_split.cu) choosing
 *                   per block between the lane-per-pair and the lane-per-symbol copier, 34 = lane per symbol only;
 *                   1..33 = the superseded round-1 kernels: only in the test-only cross-check library
 *   "decode_slots"  0 = auto; 1..30 = at most this many block slots (copier warps) per CTA of the walker + copier kernel
 *   "pipeline"      1 = the host-buffer calls overlap PCIe copies with kernels in chunks, 0 = one-shot staging
 *   "pipeline_min"  bytes below which the host-buffer calls stage in one shot
 *   "pipe_chunks"   chunks a host buffer is cut into (1..16, default 6); "pipe_taper" 1 = chunks shrink towards the end
 *   "stream_in"     1 (default) = compression streams every block's input in pieces while its encoder already runs
 *                   (blocks >= 64 KiB), 0 = whole chunks; "stream_stagger" = transfers between the chunks' kernel starts */
int tsqb_set_option(tsqb_context* ctx, const char* key, int64_t value);

/*
 * Encode ceil(total / block_size) blocks.  Replaces, per block, tsqInit + tsqEncode
 * (tsq_context.cpp:77-80, tsq_encode.cpp:192-198 -> :48-189 / :200-341).
 *   d_in       device pointer, `total` bytes followed by >= TSQB_INPUT_PAD readable bytes
 *   d_slots    device pointer, block b's stream is written at d_slots + b * slot_stride
 *   d_sizes    device pointer, n_blocks x u32: compressed size of each block
 *   with_ext   reference `withExtensions`
 *   stream     cudaStream_t (NULL = default stream); the call is asynchronous on it
 */
int tsqb_encode_blocks(tsqb_context* ctx, const uint8_t* d_in, uint64_t total, uint32_t block_size,
                       uint8_t* d_slots, uint64_t slot_stride, uint32_t* d_sizes, uint32_t with_ext,
                       void* stream);

/*
 * Decode n_blocks streams (d_comp must stay readable for 16 bytes behind the last stream: it is staged in
 * 16-byte units).  Replaces, per block, tsqDecode (tsq_decode.cpp:129-135 -> :42-126 /
 * :137-314).  Stream b starts at d_comp + (d_offsets ? d_offsets[b] : b * slot_stride); its decoded
 * bytes go to d_out + b * out_stride (never more than out_stride bytes, never past the header
 * size: unlike the reference nothing is written beyond the decoded size).
 *   d_comp_sizes  readable bytes of each stream: bounds the staging look-ahead and stops a corrupt stream from
 *                 walking off its buffer (the reference ignores inputSize).  May be NULL only with d_offsets ==
 *                 NULL (then slot_stride bounds every stream)
 *   d_out_sizes   n_blocks x u32: decoded size, 0 when the header exceeds 4 MiB (tsq_decode.cpp:53)
 *                 or out_stride
 */
int tsqb_decode_blocks(tsqb_context* ctx, const uint8_t* d_comp, const uint64_t* d_offsets,
                       uint64_t slot_stride, const uint32_t* d_comp_sizes, uint64_t n_blocks,
                       uint8_t* d_out, uint64_t out_stride, uint32_t* d_out_sizes, uint32_t with_ext,
                       void* stream);

/*
 * TSQ1 container body on the device (turbosqueeze.cpp:64-67,78-84): exclusive scan of the block
 * sizes, then every block is copied behind its u24 length prefix (| 0x800000 when with_ext).
 *   d_container   receives 16-byte header + body; capacity >= 16 + sum(sizes) + 3 * n_blocks
 *   d_total_out   one u64: container length in bytes
 */
int tsqb_pack_container(tsqb_context* ctx, const uint8_t* d_slots, uint64_t slot_stride,
                        const uint32_t* d_sizes, uint64_t n_blocks, uint64_t total_uncompressed,
                        uint32_t with_ext, uint8_t* d_container, uint64_t* d_total_out, void* stream);

/*
 * Inverse: walk the u24 chain of a TSQ1 container resident on the device (serial by construction,
 * tsq_threads.cpp:480-484) and produce per-block stream offsets / sizes / ext flags.
 *   d_offsets, d_sizes  capacity max_blocks;  d_n_blocks: one u64 (blocks found, <= max_blocks)
 */
int tsqb_index_container(tsqb_context* ctx, const uint8_t* d_container, uint64_t container_size,
                         uint64_t max_blocks, uint64_t* d_offsets, uint32_t* d_sizes,
                         uint32_t* d_ext_flags, uint64_t* d_n_blocks, void* stream);

/* Peer memory for the multi-GPU gather (one process per GPU; not in the reference, which has no devices).  The root rank
 * exports the buffer that receives the container (tsqb_ipc_export: 64-byte CUDA IPC handle of the allocation + the
 * pointer's offset inside it), every other rank maps it (tsqb_ipc_open -> base pointer in its own address space) and
 * copies its bytes to base + offset + its prefix-summed position with tsqb_copy_d2d (asynchronous on `stream`, over
 * NVLink / NVSwitch), then unmaps it (tsqb_ipc_close).  turbosqueeze_b200/sharding.py is the caller. */
int tsqb_ipc_export(const void* d_ptr, uint8_t handle[64], uint64_t* offset);
int tsqb_ipc_open(const uint8_t handle[64], void** base);
int tsqb_ipc_close(void* base);
int tsqb_copy_d2d(void* dst, const void* src, uint64_t n, void* stream);

/* Host-buffer convenience used by the reference-style entry points and the end-to-end benchmark:
 * H2D, tsqb_encode_blocks, D2H of sizes and of the slots' used bytes, inside one call.
 *   in        host pointer (pageable or pinned), `total` bytes; the library pads on the device
 *   slots     host pointer, capacity n_blocks * tsqb_slot_stride(block_size)
 *   sizes     host pointer, n_blocks x u32
 */
int tsqb_encode_host(tsqb_context* ctx, const uint8_t* in, uint64_t total, uint32_t block_size,
                     uint8_t* slots, uint32_t* sizes, uint32_t with_ext);
int tsqb_decode_host(tsqb_context* ctx, const uint8_t* slots, uint64_t slot_stride, const uint32_t* comp_sizes,
                     uint64_t n_blocks, uint8_t* out, uint64_t out_stride, uint32_t* out_sizes,
                     uint32_t with_ext);

/* Host TSQ1 container <-> host buffer through the device (block_size <= 4 MiB; the reference's
 * own pipeline is fixed at 4 MiB, turbosqueeze.h:37-38).  *out is malloc'ed; release with free().
 * Replaces the memory->memory mode of tsqCompress_MT / tsqDecompress_MT (tsq_threads.cpp:413-441,
 * :862-890).  Returns 0 on success. */
int tsqb_compress_buffer(tsqb_context* ctx, const uint8_t* in, uint64_t total, uint32_t block_size,
                         uint32_t with_ext, uint8_t** out, uint64_t* out_size);
int tsqb_decompress_buffer(tsqb_context* ctx, const uint8_t* in, uint64_t in_size, uint8_t** out,
                           uint64_t* out_size);
/* Same, into a caller-owned host buffer (pinned memory makes both PCIe legs run at link speed). */
int tsqb_compress_into(tsqb_context* ctx, const uint8_t* in, uint64_t total, uint32_t block_size,
                       uint32_t with_ext, uint8_t* out, uint64_t out_capacity, uint64_t* out_size);
int tsqb_decompress_into(tsqb_context* ctx, const uint8_t* in, uint64_t in_size, uint8_t* out,
                         uint64_t out_capacity, uint64_t* out_size);

/* ------------------------------------------------------------------------------------------------
 * Layer 2: the reference's entry points (reference turbosqueeze.h, line cited on each)
 * ---------------------------------------------------------------------------------------------- */

/* turbosqueeze.h:57-63 -- layout kept: the reference test pokes refhash (test/test.cpp:42). */
struct TSQCompressionContext {
    uint16_t* refhash;
};

struct TSQCompressionContext*  tsqAllocateContext(void);                         /* turbosqueeze.h:625 */
void tsqDeallocateContext(struct TSQCompressionContext* ctx);                    /* turbosqueeze.h:634 */
void tsqInit(struct TSQCompressionContext* ctx);                                 /* turbosqueeze.h:643 */

/* turbosqueeze.h:657 -- one block, host buffers.  inputBlock must be followed by >= 19 readable
 * bytes exactly as for the reference (tsq_encode.cpp:74,126-128).  The table is taken as zero. */
void tsqEncode(struct TSQCompressionContext* ctx, uint8_t* inputBlock, uint8_t* outputBlock,
               uint32_t* outputSize, uint32_t inputSize, uint32_t withExtensions);

/* turbosqueeze.h:670 -- one block, host buffers; *outputSize = 0 when the header exceeds 4 MiB. */
void tsqDecode(uint8_t* inputBlock, uint8_t* outputBlock, uint32_t* outputSize, uint32_t inputSize,
               uint32_t withExtensions);

/* Block size used by tsqCompress / tsqCompress_MT / tsqCompressAsync_MT (process-wide; default 4 MiB = the
 * reference's TSQ_BLOCK_SZ, which makes the containers byte-identical to the reference's).  The container does not
 * record the block size and the reference decodes any block <= 4 MiB, so e.g. 256 KiB keeps the files readable by
 * the reference while giving the GPU 16x more independent blocks.  Returns 0 on success.  The environment variable
 * TSQB_CONTAINER_BLOCK=<bytes>, read when the library is loaded, sets the same default. */
int tsqb_set_container_block_size(uint32_t block_size);

/* turbosqueeze.h:458,470 -- TSQ1 files, 4 MiB blocks by default; `level` is ignored as in the reference
 * (turbosqueeze.cpp:48-95); silent return on failure (turbosqueeze.cpp:103,107-117). */
void tsqCompress(FILE* in, FILE* out, bool useextensions, uint32_t level);
void tsqDecompress(FILE* in, FILE* out);

/* turbosqueeze.h:480-616 -- the synchronous "buffer API".  The contexts are opaque here (the
 * reference's are thread pools; on the GPU the pool is the grid).  `in` / `*out` are file NAMES when
 * infile / outfile are set (tsq_threads.cpp:284-360); a memory *out is malloc'ed by the library
 * and released by the caller with free() (turbosqueeze.h:505,577). */
struct TSQCompressionContext_MT;
struct TSQDecompressionContext_MT;
struct TSQCompressionContext_MT*   tsqAllocateContextCompression_MT(bool verbose);     /* :480 */
void tsqDeallocateContextCompression_MT(struct TSQCompressionContext_MT* ctx);         /* :489 */
bool tsqCompress_MT(struct TSQCompressionContext_MT* ctx, uint8_t* in, size_t szin, bool infile,
                    uint8_t** out, size_t* szout, bool outfile, bool useextensions,
                    uint32_t level);                                                   /* :508 */
struct TSQDecompressionContext_MT* tsqAllocateContextDecompression_MT(bool verbose);   /* :554 */
void tsqDeallocateContextDecompression_MT(struct TSQDecompressionContext_MT* ctx);     /* :563 */
bool tsqDecompress_MT(struct TSQDecompressionContext_MT* ctx, uint8_t* in, size_t szin, bool infile,
                      uint8_t** out, size_t* szout, bool outfile);                     /* :580 */

#ifdef __cplusplus
}

/* turbosqueeze.h:543-544,615-616 -- the asynchronous job API.  As in the reference these two have C linkage but
 * C++ parameter types.  Up to two jobs of one context run at the same time on the GPU (two job threads, each with device
 * scratch of its own); their callbacks are delivered strictly in submission order, as from the reference's single writer
 * thread (tsq_threads.cpp:192-275): progress_cb(id, blocks_done / n_blocks) once per block (:248-254, :654-655), then
 * completion_cb(id, success) (:256-268).  Returns the job id (>= 1); on an early failure completion_cb(0, false) is
 * invoked and 0 returned (tsq_threads.cpp:296-306).  Callbacks may submit jobs to another context
 * (test/test.cpp:247-262).  tsqDeallocateContext*_MT waits for the jobs in flight (tsq_context.cpp:150-155).
 * The synchronous calls above are "submit, then wait for the completion callback", as in the reference (:413-441).
 * Memory input: the last block reads the caller's bytes behind the buffer (<= 19, 72 with extensions) exactly as the
 * reference's in-place encoder does (tsq_threads.cpp:109); pages that are not mapped count as zeros. */
#include <functional>
extern "C" {
uint32_t tsqCompressAsync_MT(struct TSQCompressionContext_MT* ctx, uint8_t* in, size_t szin, bool infile, uint8_t** out, size_t* szout,
                             bool outfile, bool useextensions, uint32_t level,
                             std::function<void(uint32_t jobid, bool)> user_completion_cb,
                             std::function<void(uint32_t jobid, double)> user_progress_cb);                     /* :543 */
uint32_t tsqDecompressAsync_MT(struct TSQDecompressionContext_MT* ctx, uint8_t* in, size_t szin, bool infile, uint8_t** out,
                               size_t* szout, bool outfile,
                               std::function<void(uint32_t jobid, bool)> user_completion_cb,
                               std::function<void(uint32_t jobid, double)> user_progress_cb);                   /* :615 */
}
#endif

#endif /* TSQ_B200_H */
