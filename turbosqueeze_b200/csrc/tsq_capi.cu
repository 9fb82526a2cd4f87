// tsq_capi.cu -- the C-ABI of libturbosqueeze_b200.so (include/tsq_b200.h).
//
// Layer 1 (tsqb_*) launches the kernels on device-resident buffers.  Layer 2 re-implements the
// reference's entry points (reference turbosqueeze.h:458-670) as host wrappers around layer 1:
// stage H2D, launch, copy back.  No CPU codec exists in this library: without a CUDA device every
// call fails.
#include "../../include/tsq_b200.h"
#include "tsq_device.cuh"

#include <atomic>
#include <sys/mman.h>
#include <unistd.h>
#include <condition_variable>
#include <cstdarg>
#include <deque>
#include <functional>
#include <thread>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

using namespace tsqb;

// ------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;

static int fail(const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail("%s: %s", #call, cudaGetErrorString(e_));              \
    } while (0)

extern "C" const char* tsqb_last_error(void) { return g_err.c_str(); }

static std::atomic<uint64_t> g_launches{0};
extern "C" uint64_t tsqb_launch_count(void) { return g_launches.load(); }

extern "C" int tsqb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ----------------------------------------------------------------------------------------- context
struct DevBuf {
    void*  p = nullptr;
    size_t cap = 0;
    bool zero_on_alloc = false;
    int ensure(size_t n)
    {
        if (n <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = (n + (1u << 20)) & ~(size_t)((1u << 20) - 1);
        if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); p = nullptr; return 1; }
        // The library's streams are non-blocking, so nothing orders a legacy-stream memset before their kernels:
        // wait for it here (allocations are rare; the epoch scheme relies on "zero = never written").
        if (zero_on_alloc && (cudaMemset(p, 0, want) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess)) {
            cudaGetLastError(); cudaFree(p); p = nullptr; return 1;
        }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct tsqb_context {
    int device = 0;
    int sm_count = 148;
    int encode_impl = 0;       // 0 auto, 1 scalar, 2 warp
    int decode_lanes = 0;      // 0 auto
    int decode_slots = 0;      // 0 auto; else at most this many block slots (copier warps) per CTA of the walker + copier kernel
    int64_t encode_slots = 0;  // 0 auto
    int encode_hints = 0;      // see EncodeArgs::hints (a development option; a default build ignores it)
    int encode_fat = -1;       // batch encoder table format: -1 auto, 0 u16 tables, 1 sector entries
    DevBuf tables;             // hash tables of the blocks in flight (zeroed when allocated: epoch 0 = empty)
    DevBuf ftables;            // batch encoder: 32-byte entries, only ever written by that kernel, zeroed at allocation
    uint64_t launch_id = 1u << 12;  // encode launches so far: the epoch of the batch encoder's table entries.  Starts high enough
                               // that (epoch >> 32) is never 0: a half-zeroed sector can never pass for a live entry
    cudaEvent_t ev_scratch = nullptr;   // last launch that used the context's shared scratch (tables, pack offsets)
    bool scratch_busy = false;
    // staging for the host-buffer entry points
    DevBuf in, slots, sizes, out, osizes, cont, offs, ext, misc;
    cudaStream_t stream = nullptr;
    // pipelined host path: copy-in / copy-out streams, one compute stream + events per chunk in flight
    static constexpr int kPipe = 16;           // most chunks a buffer is cut into
    int pipe_taper = 0;                        // compress: chunks shrink towards the end (option "pipe_taper")
    int pipe_chunks = 6;                       // chunks actually used (option "pipe_chunks", 1..kPipe; 4 / 6 / 8 / 12: 70.3 / 68.9 / 70.6 / 80.7 ms per GB round trip)
    cudaStream_t s_in = nullptr, s_out = nullptr, s_chunk[kPipe] = {};
    cudaEvent_t ev_in[kPipe] = {}, ev_done[kPipe] = {};
    uint64_t* h_len = nullptr;                 // pinned: per-chunk container length
    int pipeline = 1;                          // 0: one-shot staging (round-1 v1 behaviour)
    int stream_in = 1;                         // compress: 1 = piece-streamed input (compress_streamed), 0 = whole chunks (option "stream_in")
    int stream_stagger = 0;                    // compress_streamed: chunk k's kernel starts behind transfer k * stream_stagger (option "stream_stagger";
                                               // measured 0 / 3 / 5 / 8: 43.3 / 44.9 / 45.6 / 47.0 ms per GB -- starting late only finishes late)
    std::vector<cudaEvent_t> ev_xfer;          // compress_streamed: one event per (piece, chunk) transfer
    uint32_t* h_piece = nullptr;               // pinned: the values the arrival flags take
    DevBuf flags;                              // arrival flags of the chunks + the kernels' error word
    uint64_t pipeline_min = 64ull << 20;       // buffers below this many bytes are staged in one shot
    std::recursive_mutex mtx;          // host entry points hold it across their layer-1 calls
};

extern "C" int tsqb_create(tsqb_context** out, int device)
{
    if (!out) return fail("tsqb_create: null out pointer");
    *out = nullptr;
    const int n = tsqb_device_count();
    if (n <= 0) return fail("tsqb_create: no CUDA device available (this library has no CPU path)");
    if (device < 0 || device >= n) return fail("tsqb_create: device %d out of range (%d devices)", device, n);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    tsqb_context* c = new tsqb_context();
    c->ftables.zero_on_alloc = true;                                  // epoch 0 = never written
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking) == cudaSuccess &&
              cudaMallocHost((void**)&c->h_len, sizeof(uint64_t) * tsqb_context::kPipe) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_scratch, cudaEventDisableTiming) == cudaSuccess;
    for (int k = 0; ok && k < tsqb_context::kPipe; k++)
        ok = cudaStreamCreateWithFlags(&c->s_chunk[k], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->ev_in[k], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->ev_done[k], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        delete c;
        return fail("tsqb_create: cannot create CUDA streams / events");
    }
    *out = c;
    g_err.clear();
    return 0;
}

extern "C" void tsqb_destroy(tsqb_context* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (DevBuf* b : {&c->tables, &c->ftables, &c->in, &c->slots, &c->sizes, &c->out, &c->osizes, &c->cont, &c->offs, &c->ext, &c->misc})
        b->release();
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    for (int k = 0; k < tsqb_context::kPipe; k++) {
        if (c->s_chunk[k]) cudaStreamDestroy(c->s_chunk[k]);
        if (c->ev_in[k]) cudaEventDestroy(c->ev_in[k]);
        if (c->ev_done[k]) cudaEventDestroy(c->ev_done[k]);
    }
    if (c->h_len) cudaFreeHost(c->h_len);
    if (c->h_piece) cudaFreeHost(c->h_piece);
    for (cudaEvent_t e : c->ev_xfer) cudaEventDestroy(e);
    c->flags.release();
    if (c->ev_scratch) cudaEventDestroy(c->ev_scratch);
    delete c;
}

extern "C" uint64_t tsqb_slot_stride(uint32_t block)
{
    const uint64_t n = 5ull + block + (block >> 4) + ((block + 15u) >> 4) + 32u;
    return (n + 127u) / 128u * 128u;
}

extern "C" int tsqb_set_option(tsqb_context* c, const char* key, int64_t v)
{
    if (!c || !key) return 1;
    if (!strcmp(key, "encode_impl"))  { c->encode_impl = (int)v; return 0; }
    if (!strcmp(key, "decode_lanes")) { c->decode_lanes = (int)v; return 0; }
    if (!strcmp(key, "decode_slots")) { if (v < 0 || v > 30) return 1; c->decode_slots = (int)v; return 0; }
    if (!strcmp(key, "encode_slots")) { c->encode_slots = v; return 0; }
    if (!strcmp(key, "encode_fat")) { c->encode_fat = (int)v; return 0; }
    if (!strcmp(key, "encode_hints")) { c->encode_hints = (int)v; return 0; }
    if (!strcmp(key, "pipeline")) { c->pipeline = (int)v; return 0; }
    if (!strcmp(key, "pipeline_min")) { c->pipeline_min = (uint64_t)v; return 0; }
    if (!strcmp(key, "pipe_taper")) { c->pipe_taper = (int)v; return 0; }
    if (!strcmp(key, "stream_in")) { c->stream_in = (int)v; return 0; }
    if (!strcmp(key, "stream_stagger")) { if (v < 0 || v > 64) return 1; c->stream_stagger = (int)v; return 0; }
    if (!strcmp(key, "pipe_chunks")) { if (v < 1 || v > tsqb_context::kPipe) return 1; c->pipe_chunks = (int)v; return 0; }
    if (!strcmp(key, "l2_fetch")) {                                  // 32 / 64 / 128: DRAM fetch granularity hint
        cudaSetDevice(c->device);
        return cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)v) == cudaSuccess ? 0 : 1;
    }
    return 1;
}

// ------------------------------------------------------------------------------ layer 1: device path
// Layer 1 may be called from several threads and on several streams with one context, but the context's scratch (hash
// tables, pack offsets, the epoch counter) is shared: a launch takes c->mtx for its set-up, makes its stream wait for
// the previous user of the scratch (an event), launches, and records the event again.  Launches on different streams
// are therefore serialised on the device where they share tables -- never interleaved.  Launches that bring their own
// table region (`tables`, the pipelined host path) skip the wait.
static int scratch_acquire(tsqb_context* c, cudaStream_t st)
{
    if (c->scratch_busy) CU(cudaStreamWaitEvent(st, c->ev_scratch, 0));
    return 0;
}

static int scratch_release(tsqb_context* c, cudaStream_t st)
{
    CU(cudaEventRecord(c->ev_scratch, st));
    c->scratch_busy = true;
    return 0;
}

// `tables` (optional): caller-provided region of encode_slots_for(...) tables, for launches that overlap in time
static int encode_blocks_impl(tsqb_context* c, const uint8_t* d_in, uint64_t total, uint32_t block, uint8_t* d_slots,
                              uint64_t stride, uint32_t* d_sizes, uint32_t* d_tailflags, uint32_t with_ext, void* stream,
                              uint16_t* tables = nullptr, int64_t slot_cap = 0, const uint32_t* arrived = nullptr,
                              const uint32_t* arrived_next = nullptr, uint32_t* stream_error = nullptr)
{
    if (!c) return fail("tsqb_encode_blocks: null context");
    if (block == 0 || block > kBlockMax) return fail("tsqb_encode_blocks: block size %u not in 1..%u", block, kBlockMax);
    if (stride < tsqb_slot_stride(block)) return fail("tsqb_encode_blocks: slot stride %llu < %llu", (unsigned long long)stride,
                                                      (unsigned long long)tsqb_slot_stride(block));
    if (total == 0) return 0;
    std::lock_guard<std::recursive_mutex> lk(c->mtx);
    CU(cudaSetDevice(c->device));
    EncodeArgs a;
    a.in = d_in; a.total = total; a.block = block; a.nb = (total + block - 1) / block;
    a.slots = d_slots; a.stride = stride; a.sizes = d_sizes; a.tailflags = d_tailflags;
    a.hints = (uint32_t)c->encode_hints;
    a.arrived = arrived; a.arrived_next = arrived_next; a.stream_error = stream_error;
    a.epoch = (++c->launch_id) << 20;                                 // + the slot's block counter, < 2^20
    const int impl = c->encode_impl == 1 ? 1 : ((c->encode_impl == 2 && !with_ext) ? 2 : 3);
    a.n_slots = encode_slots_for(impl, a.nb, c->sm_count, slot_cap > 0 ? slot_cap : c->encode_slots);
    a.fat = (c->encode_fat < 0 ? encode_wants_fat(impl, a.n_slots) : (impl == 3 && c->encode_fat != 0)) ? 1u : 0u;
    if (tables) { a.tables = tables; a.fat = impl == 3 ? 1u : 0u; }            // the pipelined path provisions sector tables
    else {
        // The tables of all blocks in flight: up to sm_count * 28 x 4 MiB = 16.2 GiB.  On a GPU that cannot give that
        // much (shared, or smaller), run with fewer blocks in flight instead of failing: halve until it fits.
        DevBuf& tb = a.fat ? c->ftables : c->tables;
        for (;;) {
            const size_t need = (size_t)a.n_slots * encode_table_bytes(impl, a.fat != 0);
            size_t free_b = 0, total_b = 0;
            const bool fits = need <= tb.cap || cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || need + (256u << 20) <= free_b + tb.cap;
            if (fits && tb.ensure(need) == 0) break;
            if (a.n_slots <= 1) return fail("tsqb_encode_blocks: cannot allocate a hash table (%zu bytes)", need);
            a.n_slots = (a.n_slots + 1) / 2;
        }
        a.tables = (uint16_t*)tb.p;
        if (scratch_acquire(c, (cudaStream_t)stream)) return 1;
    }
    CU(launch_encode(a, impl, with_ext != 0, c->sm_count, (cudaStream_t)stream));
    if (!tables && scratch_release(c, (cudaStream_t)stream)) return 1;
    g_launches += 1;
    return 0;
}

extern "C" int tsqb_encode_blocks(tsqb_context* c, const uint8_t* d_in, uint64_t total, uint32_t block, uint8_t* d_slots,
                                  uint64_t stride, uint32_t* d_sizes, uint32_t with_ext, void* stream)
{
    return encode_blocks_impl(c, d_in, total, block, d_slots, stride, d_sizes, nullptr, with_ext, stream);
}

extern "C" int tsqb_decode_blocks(tsqb_context* c, const uint8_t* d_comp, const uint64_t* d_offsets, uint64_t stride,
                                  const uint32_t* d_comp_sizes, uint64_t nb, uint8_t* d_out, uint64_t out_stride,
                                  uint32_t* d_out_sizes, uint32_t with_ext, void* stream)
{
    if (!c) return fail("tsqb_decode_blocks: null context");
    if (nb == 0) return 0;
    // packed streams have no slot to bound a corrupt stream (or the staging look-ahead): their sizes are required
    if (d_offsets && !d_comp_sizes) return fail("tsqb_decode_blocks: d_comp_sizes is required together with d_offsets");
    CU(cudaSetDevice(c->device));
    DecodeArgs a;
    a.comp = d_comp; a.offs = d_offsets; a.stride = stride; a.csizes = d_comp_sizes; a.nb = nb;
    a.out = d_out; a.ostride = out_stride; a.osizes = d_out_sizes;
    CU(launch_decode(a, c->decode_lanes, with_ext != 0, c->sm_count, (cudaStream_t)stream, c->decode_slots));
    g_launches += 1;
    return 0;
}

extern "C" int tsqb_pack_container(tsqb_context* c, const uint8_t* d_slots, uint64_t stride, const uint32_t* d_sizes,
                                   uint64_t nb, uint64_t total_u, uint32_t with_ext, uint8_t* d_container,
                                   uint64_t* d_total_out, void* stream)
{
    if (!c) return fail("tsqb_pack_container: null context");
    std::lock_guard<std::recursive_mutex> lk(c->mtx);
    CU(cudaSetDevice(c->device));
    if (c->offs.cap < (nb + 1) * sizeof(uint64_t) && c->scratch_busy) CU(cudaEventSynchronize(c->ev_scratch));   // about to be re-allocated
    if (c->offs.ensure((nb + 1) * sizeof(uint64_t))) return fail("tsqb_pack_container: out of device memory");
    if (scratch_acquire(c, (cudaStream_t)stream)) return 1;
    CU(launch_pack(d_slots, stride, d_sizes, nb, total_u, with_ext, d_container, d_total_out, (uint64_t*)c->offs.p,
                   (cudaStream_t)stream));
    if (scratch_release(c, (cudaStream_t)stream)) return 1;
    g_launches += nb ? 2 : 1;
    return 0;
}

extern "C" int tsqb_index_container(tsqb_context* c, const uint8_t* d_container, uint64_t csize, uint64_t max_blocks,
                                    uint64_t* d_offsets, uint32_t* d_sizes, uint32_t* d_ext, uint64_t* d_n, void* stream)
{
    if (!c) return fail("tsqb_index_container: null context");
    CU(cudaSetDevice(c->device));
    CU(launch_index(d_container, csize, max_blocks, d_offsets, d_sizes, d_ext, d_n, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

// ------------------------------------------------------------------ peer memory (multi-GPU gather, one process per GPU)
// The one exchange step of the path is the gather of the per-rank streams into ONE container on the root GPU
// (SURVEY.md 8(e)).  Instead of NCCL send/recv (staged through NCCL's channel buffers, ~320 GB/s into the root with seven
// senders), every rank writes its bytes straight into the root's buffer over NVLink / NVSwitch: the root exports its
// buffer as a CUDA IPC handle, the peers map it and issue one device-to-device copy each at their prefix-summed offset.
// These four calls are that plumbing, nothing else.
extern "C" int tsqb_ipc_export(const void* d_ptr, uint8_t handle[64], uint64_t* offset)
{
    if (!d_ptr || !handle || !offset) return fail("tsqb_ipc_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    // An IPC handle names a whole allocation: find its base (the pointer may sit inside a caching allocator's segment).
    typedef int (*GetRange)(unsigned long long*, size_t*, unsigned long long);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CU(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr));
    if (!fn || qr != cudaDriverEntryPointSuccess) return fail("tsqb_ipc_export: cuMemGetAddressRange is not available");
    unsigned long long base = 0;
    size_t size = 0;
    if (((GetRange)fn)(&base, &size, (unsigned long long)(uintptr_t)d_ptr) != 0) return fail("tsqb_ipc_export: not a device allocation");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, (void*)(uintptr_t)base));
    memcpy(handle, &h, 64);
    *offset = (uint64_t)((uintptr_t)d_ptr - (uintptr_t)base);
    return 0;
}

extern "C" int tsqb_ipc_open(const uint8_t handle[64], void** base)
{
    if (!handle || !base) return fail("tsqb_ipc_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CU(cudaIpcOpenMemHandle(base, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int tsqb_ipc_close(void* base)
{
    if (!base) return 0;
    CU(cudaIpcCloseMemHandle(base));
    return 0;
}

extern "C" int tsqb_copy_d2d(void* dst, const void* src, uint64_t n, void* stream)
{
    if (n == 0) return 0;
    CU(cudaMemcpyAsync(dst, src, n, cudaMemcpyDefault, (cudaStream_t)stream));    // unified addressing: works for peer memory
    return 0;
}

// --------------------------------------------------------------------------- host-buffer convenience
// `tail` (optional, <= TSQB_INPUT_PAD bytes) are the bytes that follow the input in the caller's memory.
static int stage_input(tsqb_context* c, const uint8_t* in, uint64_t total, const uint8_t* tail, uint32_t tail_n)
{
    if (c->in.ensure(total + 2 * TSQB_INPUT_PAD)) return fail("out of device memory for %llu input bytes", (unsigned long long)total);
    uint8_t* d = (uint8_t*)c->in.p;
    CU(cudaMemsetAsync(d + total, 0, 2 * TSQB_INPUT_PAD, c->stream));
    CU(cudaMemcpyAsync(d, in, total, cudaMemcpyHostToDevice, c->stream));
    if (tail && tail_n) CU(cudaMemcpyAsync(d + total, tail, tail_n, cudaMemcpyHostToDevice, c->stream));
    return 0;
}

extern "C" int tsqb_encode_host(tsqb_context* c, const uint8_t* in, uint64_t total, uint32_t block, uint8_t* slots,
                                uint32_t* sizes, uint32_t with_ext)
{
    if (!c) return fail("tsqb_encode_host: null context");
    if (block == 0 || block > kBlockMax) return fail("tsqb_encode_host: bad block size %u", block);
    if (total == 0) return 0;
    std::lock_guard<std::recursive_mutex> lk(c->mtx);
    CU(cudaSetDevice(c->device));
    const uint64_t nb = (total + block - 1) / block, stride = tsqb_slot_stride(block);
    if (stage_input(c, in, total, nullptr, 0)) return 1;
    if (c->slots.ensure(nb * stride) || c->sizes.ensure(nb * 4)) return fail("tsqb_encode_host: out of device memory");
    CU(cudaMemsetAsync(c->slots.p, 0, nb * stride, c->stream));          // zero-filled slots (parity contract)
    if (tsqb_encode_blocks(c, (uint8_t*)c->in.p, total, block, (uint8_t*)c->slots.p, stride, (uint32_t*)c->sizes.p, with_ext, c->stream)) return 1;
    CU(cudaMemcpyAsync(sizes, c->sizes.p, nb * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    // only the used part of the slots travels back: ONE strided copy, as wide as the largest stream
    uint32_t widest = 0;
    for (uint64_t b = 0; b < nb; b++) if (sizes[b] > widest) widest = sizes[b];
    if (widest) CU(cudaMemcpy2DAsync(slots, stride, c->slots.p, stride, widest, nb, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int tsqb_decode_host(tsqb_context* c, const uint8_t* slots, uint64_t stride, const uint32_t* comp_sizes, uint64_t nb,
                                uint8_t* out, uint64_t out_stride, uint32_t* out_sizes, uint32_t with_ext)
{
    if (!c) return fail("tsqb_decode_host: null context");
    if (nb == 0) return 0;
    std::lock_guard<std::recursive_mutex> lk(c->mtx);
    CU(cudaSetDevice(c->device));
    if (c->slots.ensure(nb * stride + 256) || c->sizes.ensure(nb * 4) || c->out.ensure(nb * out_stride) || c->osizes.ensure(nb * 4))
        return fail("tsqb_decode_host: out of device memory");
    if (comp_sizes) {
        uint64_t widest = 0;
        for (uint64_t b = 0; b < nb; b++) if (comp_sizes[b] > widest) widest = comp_sizes[b];
        if (widest > stride) widest = stride;
        if (widest) CU(cudaMemcpy2DAsync(c->slots.p, stride, slots, stride, widest, nb, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->sizes.p, comp_sizes, nb * 4, cudaMemcpyHostToDevice, c->stream));
    } else {
        CU(cudaMemcpyAsync(c->slots.p, slots, nb * stride, cudaMemcpyHostToDevice, c->stream));
    }
    if (tsqb_decode_blocks(c, (uint8_t*)c->slots.p, nullptr, stride, comp_sizes ? (uint32_t*)c->sizes.p : nullptr, nb,
                           (uint8_t*)c->out.p, out_stride, (uint32_t*)c->osizes.p, with_ext, c->stream)) return 1;
    CU(cudaMemcpyAsync(out_sizes, c->osizes.p, nb * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    // contiguous output: one copy when blocks are packed back to back
    CU(cudaMemcpyAsync(out, c->out.p, (nb - 1) * out_stride + out_sizes[nb - 1], cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

// Per-block progress of a host-path job (tsq_threads.cpp:248-254, :654-655: the reference's writer thread reports
// (blocks written) / n_blocks after every block).  Called on the thread that runs the job, in block order.
typedef std::function<void(uint64_t blocks_done, uint64_t n_blocks)> ProgressFn;
static void report_blocks(const ProgressFn* prog, uint64_t from, uint64_t to, uint64_t nb)
{
    if (prog && *prog) for (uint64_t b = from; b < to; b++) (*prog)(b + 1, nb);
}

// `tail`: see stage_input.  Container is assembled on the device and comes back in one copy, either
// into a fresh malloc (host_out == nullptr) or into the caller's buffer of `host_cap` bytes.
static int compress_locked(tsqb_context* c, const uint8_t* in, uint64_t total, const uint8_t* tail, uint32_t tail_n, uint32_t block,
                           uint32_t with_ext, uint8_t* host_out, uint64_t host_cap, uint8_t** out, uint64_t* out_size,
                           const ProgressFn* prog = nullptr)
{
    const uint64_t nb = (total + block - 1) / block, stride = tsqb_slot_stride(block);
    if (stage_input(c, in, total, tail, tail_n)) return 1;
    const uint64_t cap = 16 + nb * (stride + 3);
    if (c->slots.ensure(nb * stride + 256) || c->sizes.ensure(nb * 4 + 4) || c->cont.ensure(cap + 256) || c->misc.ensure(64))
        return fail("compress: out of device memory");
    // parity contract: every block's output slot is zero-filled before it is encoded
    if (nb) CU(cudaMemsetAsync(c->slots.p, 0, nb * stride, c->stream));
    if (nb && tsqb_encode_blocks(c, (uint8_t*)c->in.p, total, block, (uint8_t*)c->slots.p, stride, (uint32_t*)c->sizes.p, with_ext, c->stream)) return 1;
    if (tsqb_pack_container(c, (uint8_t*)c->slots.p, stride, (uint32_t*)c->sizes.p, nb, total, with_ext, (uint8_t*)c->cont.p,
                            (uint64_t*)c->misc.p, c->stream)) return 1;
    uint64_t clen = 0;
    CU(cudaMemcpyAsync(&clen, c->misc.p, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    uint8_t* host = host_out;
    if (!host) {
        host = (uint8_t*)malloc(clen ? clen : 1);
        if (!host) return fail("compress: malloc(%llu) failed", (unsigned long long)clen);
    } else if (clen > host_cap) {
        return fail("compress: output needs %llu bytes, caller gave %llu", (unsigned long long)clen, (unsigned long long)host_cap);
    }
    cudaError_t e = cudaMemcpyAsync(host, c->cont.p, clen, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
        if (!host_out) free(host);
        return fail("compress: %s", cudaGetErrorString(e));
    }
    report_blocks(prog, 0, nb, nb);
    if (out) *out = host;
    *out_size = clen;
    return 0;
}

// ------------------------------------------------------------------------------ pipelined host path
// The host entry points move 2 x (U + C) bytes over PCIe around ~40 ms of kernels.  Above kPipeMin
// bytes the input is cut into up to kPipe chunks of whole blocks: chunk k's H2D copy, its kernels and
// the D2H copy of its result run on different streams, so the copies of one chunk hide behind the
// kernels of another (this is what the reference's reader / worker / writer threads do on the CPU,
// tsq_threads.cpp:52-275).  The chunks' kernels overlap in time, which keeps thousands of blocks in
// flight; every chunk therefore gets its own hash-table region.

static int compress_pipelined(tsqb_context* c, const uint8_t* in, uint64_t total, const uint8_t* tail, uint32_t tail_n, uint32_t block,
                              uint32_t with_ext, uint8_t* host_out, uint64_t host_cap, uint8_t** out, uint64_t* out_size,
                              const ProgressFn* prog = nullptr)
{
    constexpr int KMAX = tsqb_context::kPipe;
    const int K = c->pipe_chunks;
    const uint64_t nb = (total + block - 1) / block, stride = tsqb_slot_stride(block);
    // chunk boundaries (in blocks).  Equal chunks, or -- option "pipe_taper" -- chunks that shrink towards the end:
    // a block takes ~20 ms from the arrival of its bytes however few blocks are in flight (its parse is a serial
    // chain), so the job ends one block latency after the LAST chunk has landed; a small last chunk lands and
    // leaves quickly.
    uint64_t cb[KMAX + 1];
    int nchunks = 0;
    cb[0] = 0;
    if (c->pipe_taper && K >= 3) {
        double w[KMAX], sum = 0;
        for (int k = 0; k < K; k++) { w[k] = 1.0 / (1.0 + 0.6 * k * k / (double)K); sum += w[k]; }
        double acc = 0;
        for (int k = 0; k < K; k++) {
            acc += w[k];
            uint64_t e = k == K - 1 ? nb : (uint64_t)(nb * (acc / sum) + 0.5);
            if (e > nb) e = nb;
            if (e > cb[nchunks]) cb[++nchunks] = e;
        }
    } else {
        const uint64_t per = (nb + K - 1) / K;
        for (uint64_t b = 0; b < nb; b += per) cb[++nchunks] = (b + per < nb) ? b + per : nb;
    }
    uint64_t per = 0;                                                         // blocks in the largest chunk
    for (int k = 0; k < nchunks; k++) if (cb[k + 1] - cb[k] > per) per = cb[k + 1] - cb[k];
    const int impl = c->encode_impl == 1 ? 1 : ((c->encode_impl == 2 && !with_ext) ? 2 : 3);
    // the chunks' kernels share the GPU: together they get the tables of one full grid (32 warps per SM)
    // (each chunk its share, in proportion to its blocks)
    int64_t slot_cap[KMAX];
    uint64_t slots_tab[KMAX], tab_at[KMAX], tab_total = 0;
    for (int k = 0; k < nchunks; k++) {
        const uint64_t grid_slots = impl == 3 ? encode_batch_resident_warps(c->sm_count) : (uint64_t)c->sm_count * 32;
        slot_cap[k] = c->encode_slots > 0 ? c->encode_slots : (int64_t)((grid_slots * (cb[k + 1] - cb[k]) + nb - 1) / nb);
        slots_tab[k] = encode_slots_for(impl, cb[k + 1] - cb[k], c->sm_count, slot_cap[k]);
        tab_at[k] = tab_total; tab_total += slots_tab[k];
    }
    const uint64_t ccap = 16 + per * (stride + 3) + 256;                     // container capacity of one chunk
    if (c->in.ensure(total + 2 * TSQB_INPUT_PAD) || c->slots.ensure(nb * stride + 256) || c->sizes.ensure(nb * 4 + 4) ||
        c->cont.ensure(ccap * nchunks) || c->offs.ensure((nb + KMAX) * 8) || c->misc.ensure(64 * KMAX) || (impl == 3 ? c->ftables : c->tables).ensure(tab_total * encode_table_bytes(impl, true)))
        return fail("compress: out of device memory");
    // an error after the first chunk is in flight: let the device finish before the context's buffers are touched again
#define CUP(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) { cudaDeviceSynchronize(); return fail("%s: %s", #call, cudaGetErrorString(e_)); } \
    } while (0)
    uint8_t* d_in = (uint8_t*)c->in.p;
    CUP(cudaMemsetAsync(d_in + total, 0, 2 * TSQB_INPUT_PAD, c->s_in));
    if (tail && tail_n) CUP(cudaMemcpyAsync(d_in + total, tail, tail_n, cudaMemcpyHostToDevice, c->s_in));
    for (int k = 0; k < nchunks; k++) {
        const uint64_t b0 = cb[k], b1 = cb[k + 1];
        const uint64_t lo = b0 * block, hi = (b1 * block < total) ? b1 * block : total;
        // the last block of the chunk reads a few bytes past it (tsq_encode.cpp:74,126-128): ship them with this chunk
        const uint64_t hi_tail = (hi + TSQB_INPUT_PAD < total) ? hi + TSQB_INPUT_PAD : total;
        CUP(cudaMemcpyAsync(d_in + lo, in + lo, hi_tail - lo, cudaMemcpyHostToDevice, c->s_in));
        CUP(cudaEventRecord(c->ev_in[k], c->s_in));
        cudaStream_t st = c->s_chunk[k];
        CUP(cudaStreamWaitEvent(st, c->ev_in[k], 0));
        CUP(cudaMemsetAsync((uint8_t*)c->slots.p + b0 * stride, 0, (b1 - b0) * stride, st));   // zero-filled slots (parity contract)
        // a chunk is encoded as a buffer of its own: (hi - lo) bytes that happen to be followed by the next chunk
        if (encode_blocks_impl(c, d_in + lo, hi - lo, block, (uint8_t*)c->slots.p + b0 * stride, stride, (uint32_t*)c->sizes.p + b0, nullptr,
                               with_ext, st, (uint16_t*)((uint8_t*)(impl == 3 ? c->ftables : c->tables).p + tab_at[k] * encode_table_bytes(impl, true)), slot_cap[k])) { cudaDeviceSynchronize(); return 1; }
        uint8_t* d_cont = (uint8_t*)c->cont.p + (uint64_t)k * ccap;
        uint64_t* d_len = (uint64_t*)c->misc.p + 8 * k;
        CUP(launch_pack((uint8_t*)c->slots.p + b0 * stride, stride, (uint32_t*)c->sizes.p + b0, b1 - b0, hi - lo, with_ext, d_cont, d_len,
                       (uint64_t*)c->offs.p + b0 + k, st));
        g_launches += 2;
        CUP(cudaMemcpyAsync(&c->h_len[k], d_len, 8, cudaMemcpyDeviceToHost, st));
        CUP(cudaEventRecord(c->ev_done[k], st));
    }
#undef CUP
    // results leave in order; the body of chunk k goes behind the bodies before it
    uint8_t* host = host_out;
    if (!host) {                                                              // malloc mode: the size is known last -> worst case
        host = (uint8_t*)malloc(16 + nb * (stride + 3));
        if (!host) return fail("compress: malloc failed");
    }
    uint64_t at = 16;
    cudaError_t err = cudaSuccess;
    bool too_small = false;
    for (int k = 0; k < nchunks && err == cudaSuccess && !too_small; k++) {
        err = cudaEventSynchronize(c->ev_done[k]);
        if (err != cudaSuccess) break;
        const uint64_t body = c->h_len[k] - 16;
        if (host_out && at + body > host_cap) { too_small = true; break; }
        err = cudaMemcpyAsync(host + at, (uint8_t*)c->cont.p + (uint64_t)k * ccap + 16, body, cudaMemcpyDeviceToHost, c->s_out);
        at += body;
        report_blocks(prog, cb[k], k + 1 == nchunks ? nb - 1 : cb[k + 1], nb);   // the last block is reported when everything is home
    }
    if (err == cudaSuccess) err = cudaStreamSynchronize(c->s_out);
    if (err != cudaSuccess || too_small) {
        cudaDeviceSynchronize();                                              // let the other chunks finish before the buffers are reused
        if (!host_out) free(host);
        if (too_small) return fail("compress: output needs more than the %llu bytes the caller gave", (unsigned long long)host_cap);
        return fail("compress: %s", cudaGetErrorString(err));
    }
    memcpy(host, "TSQ1", 4);                                                  // turbosqueeze.cpp:64-67
    const uint32_t nb32 = (uint32_t)nb;
    memcpy(host + 4, &nb32, 4);
    memcpy(host + 8, &total, 8);
    if (nb) report_blocks(prog, nb - 1, nb, nb);
    if (out) *out = host;
    *out_size = at;
    return 0;
}

// Piece-streamed compression.  A block's parse is a serial chain that takes ~20 ms however few blocks run, so with whole-chunk
// staging the job ends one block latency after the LAST byte has crossed PCIe.  Here every block starts as soon as its FIRST
// piece has landed: the input crosses PCIe piece-major -- piece p of every block of chunk 0, of chunk 1, ... then piece p + 1
// (strided copies) -- each transfer followed, in stream order, by a 4-byte copy that publishes "bytes of every block's prefix
// landed" for that chunk; the encoder polls that word before it reads on (EncodeArgs::arrived).  Same bytes as the one-shot
// path (tests: test_pipelined_host_path_equals_one_shot).  Measured (1 GB, 256 KiB blocks, pinned buffers): 46.4 -> 43.3 ms;
// what remains is the parse of all blocks (they share the GPU fairly, so they all finish together, ~31 ms) and then the
// container's trip home (0.62 GB, ~11 ms), which nothing can overlap with: the bytes are not final before the blocks end.
// Starting the chunks' kernels apart (stream_stagger) does not make them finish apart; it only delays the last one.
// Returns -1 when the shape does not suit it (small or odd block sizes, few blocks): the caller takes the chunked path.
static int compress_streamed(tsqb_context* c, const uint8_t* in, uint64_t total, const uint8_t* tail, uint32_t tail_n, uint32_t block,
                             uint32_t with_ext, uint8_t* host_out, uint64_t host_cap, uint8_t** out, uint64_t* out_size,
                             const ProgressFn* prog)
{
    constexpr int KMAX = tsqb_context::kPipe, P = 8;
    const int impl = c->encode_impl == 1 ? 1 : ((c->encode_impl == 2 && !with_ext) ? 2 : 3);
    const uint64_t nb = (total + block - 1) / block, stride = tsqb_slot_stride(block);
    if (impl != 3 || block < 65536u || (block % (P * 256u)) != 0 || c->encode_slots > 0 || c->encode_fat == 0) return -1;
    // the kernels wait for their input with a ~10 s bail-out: keep the whole transfer far below that even from pageable memory
    if (total > (4ull << 30)) return -1;
    int K = c->pipe_chunks;
    if (nb < (uint64_t)K * 8u) return -1;
    const uint32_t S = block / P;
    uint64_t cb[KMAX + 1];
    for (int k = 0; k <= K; k++) cb[k] = nb * k / K;
    uint64_t per = 0;
    for (int k = 0; k < K; k++) if (cb[k + 1] - cb[k] > per) per = cb[k + 1] - cb[k];
    // every block in flight: one table per block (as the chunked path provisions them, a full grid in all)
    uint64_t tab_at[KMAX], tab_total = 0;
    for (int k = 0; k < K; k++) { tab_at[k] = tab_total; tab_total += cb[k + 1] - cb[k]; }
    if (tab_total > encode_batch_resident_warps(c->sm_count)) return -1;      // more blocks than resident warps: chunked path
    const uint64_t ccap = 16 + per * (stride + 3) + 256;
    if (c->in.ensure(total + 2 * TSQB_INPUT_PAD) || c->slots.ensure(nb * stride + 256) || c->sizes.ensure(nb * 4 + 4) ||
        c->cont.ensure(ccap * K) || c->offs.ensure((nb + KMAX) * 8) || c->misc.ensure(64 * KMAX) || c->flags.ensure(256) ||
        c->ftables.ensure(tab_total * kFatTableBytes))
        return fail("compress: out of device memory");
    if (!c->h_piece) {
        if (cudaMallocHost((void**)&c->h_piece, sizeof(uint32_t) * (P + 1)) != cudaSuccess) { cudaGetLastError(); return -1; }
    }
    for (int p = 0; p < P; p++) c->h_piece[p] = (uint32_t)(p + 1) * S;
    while (c->ev_xfer.size() < (size_t)P * KMAX) {
        cudaEvent_t e;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return -1; }
        c->ev_xfer.push_back(e);
    }
#define CUP(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) { cudaDeviceSynchronize(); return fail("%s: %s", #call, cudaGetErrorString(e_)); } \
    } while (0)
    uint8_t* d_in = (uint8_t*)c->in.p;
    uint32_t* d_flags = (uint32_t*)c->flags.p;                                // [0..K) arrival flags, [32] error word
    CUP(cudaMemsetAsync(d_flags, 0, 256, c->s_in));
    CUP(cudaMemsetAsync(d_in + total, 0, 2 * TSQB_INPUT_PAD, c->s_in));
    if (tail && tail_n) CUP(cudaMemcpyAsync(d_in + total, tail, tail_n, cudaMemcpyHostToDevice, c->s_in));
    // ---- the input, piece-major
    const uint64_t nfull = total / block;                                     // blocks that are complete
    for (int p = 0; p < P; p++)
        for (int k = 0; k < K; k++) {
            const uint64_t b0 = cb[k], b1 = cb[k + 1];
            const uint64_t rows = (b1 <= nfull ? b1 : nfull) > b0 ? (b1 <= nfull ? b1 : nfull) - b0 : 0;
            const uint64_t at = b0 * (uint64_t)block + (uint64_t)p * S;
            if (rows) CUP(cudaMemcpy2DAsync(d_in + at, block, in + at, block, S, rows, cudaMemcpyHostToDevice, c->s_in));
            if (b1 > nfull && nfull >= b0) {                                  // the ragged last block: what it has of this piece
                const uint64_t lo = nfull * (uint64_t)block + (uint64_t)p * S;
                const uint64_t hi = lo + S < total ? lo + S : total;
                if (lo < hi) CUP(cudaMemcpyAsync(d_in + lo, in + lo, hi - lo, cudaMemcpyHostToDevice, c->s_in));
            }
            CUP(cudaMemcpyAsync(d_flags + k, &c->h_piece[p], 4, cudaMemcpyHostToDevice, c->s_in));
            CUP(cudaEventRecord(c->ev_xfer[p * K + k], c->s_in));
        }
    // ---- the kernels: chunk k starts behind transfer k * stagger (never before its own first piece)
    for (int k = 0; k < K; k++) {
        const uint64_t b0 = cb[k], b1 = cb[k + 1];
        const uint64_t lo = b0 * block, hi = (b1 * block < total) ? b1 * block : total;
        int after = k * c->stream_stagger;
        if (after < k) after = k;
        if (after > P * K - 1) after = P * K - 1;
        cudaStream_t st = c->s_chunk[k];
        CUP(cudaStreamWaitEvent(st, c->ev_xfer[after], 0));
        CUP(cudaMemsetAsync((uint8_t*)c->slots.p + b0 * stride, 0, (b1 - b0) * stride, st));   // zero-filled slots (parity contract)
        if (encode_blocks_impl(c, d_in + lo, hi - lo, block, (uint8_t*)c->slots.p + b0 * stride, stride, (uint32_t*)c->sizes.p + b0, nullptr, with_ext,
                               st, (uint16_t*)((uint8_t*)c->ftables.p + tab_at[k] * kFatTableBytes), (int64_t)(b1 - b0), d_flags + k,
                               k + 1 < K ? d_flags + k + 1 : nullptr, d_flags + 32)) { cudaDeviceSynchronize(); return 1; }
        uint8_t* d_cont = (uint8_t*)c->cont.p + (uint64_t)k * ccap;
        uint64_t* d_len = (uint64_t*)c->misc.p + 8 * k;
        CUP(launch_pack((uint8_t*)c->slots.p + b0 * stride, stride, (uint32_t*)c->sizes.p + b0, b1 - b0, hi - lo, with_ext, d_cont, d_len,
                        (uint64_t*)c->offs.p + b0 + k, st));
        g_launches += 2;
        CUP(cudaMemcpyAsync(&c->h_len[k], d_len, 8, cudaMemcpyDeviceToHost, st));
        CUP(cudaEventRecord(c->ev_done[k], st));
    }
#undef CUP
    // ---- results leave in order; the body of chunk k goes behind the bodies before it
    uint8_t* host = host_out;
    if (!host) {
        host = (uint8_t*)malloc(16 + nb * (stride + 3));
        if (!host) { cudaDeviceSynchronize(); return fail("compress: malloc failed"); }
    }
    uint64_t at = 16;
    cudaError_t err = cudaSuccess;
    bool too_small = false;
    for (int k = 0; k < K && err == cudaSuccess && !too_small; k++) {
        err = cudaEventSynchronize(c->ev_done[k]);
        if (err != cudaSuccess) break;
        const uint64_t body = c->h_len[k] - 16;
        if (host_out && at + body > host_cap) { too_small = true; break; }
        err = cudaMemcpyAsync(host + at, (uint8_t*)c->cont.p + (uint64_t)k * ccap + 16, body, cudaMemcpyDeviceToHost, c->s_out);
        at += body;
        report_blocks(prog, cb[k], k + 1 == K ? nb - 1 : cb[k + 1], nb);
    }
    uint32_t stream_err = 0;
    if (err == cudaSuccess) err = cudaStreamSynchronize(c->s_out);
    if (err == cudaSuccess) err = cudaMemcpy(&stream_err, d_flags + 32, 4, cudaMemcpyDeviceToHost);
    if (err != cudaSuccess || too_small || stream_err) {
        cudaDeviceSynchronize();
        if (!host_out) free(host);
        if (too_small) return fail("compress: output needs more than the %llu bytes the caller gave", (unsigned long long)host_cap);
        if (stream_err) return fail("compress: the encoder timed out waiting for its input to arrive");
        return fail("compress: %s", cudaGetErrorString(err));
    }
    memcpy(host, "TSQ1", 4);                                                  // turbosqueeze.cpp:64-67
    const uint32_t nb32 = (uint32_t)nb;
    memcpy(host + 4, &nb32, 4);
    memcpy(host + 8, &total, 8);
    if (nb) report_blocks(prog, nb - 1, nb, nb);
    if (out) *out = host;
    *out_size = at;
    return 0;
}

// Container in host memory: the u24 chain is walked on the host (it is serial by construction,
// tsq_threads.cpp:480-484, and the bytes are right here), then chunks of blocks are copied, decoded and
// copied back on separate streams.  Returns 1 with g_err set on failure, -1 when the container is not
// regular (mixed block sizes / flags): the caller then falls back to the one-shot path.
static int decompress_pipelined(tsqb_context* c, const uint8_t* in, uint64_t in_size, uint8_t* host_out, uint64_t host_cap,
                                uint8_t** out, uint64_t* out_size, const ProgressFn* prog = nullptr)
{
    const int K = c->pipe_chunks;
    std::vector<uint64_t> offs;
    std::vector<uint32_t> sizes;
    uint32_t with_ext = 0, block = 0;
    uint64_t total = 0;
    uint32_t nb_hdr;
    bool short_seen = false;                                                  // a block shorter than the first one: must be the last
    memcpy(&nb_hdr, in + 4, 4);                                               // tsq_threads.cpp:728-757 reads this many blocks
    for (uint64_t at = 16; at + 3 <= in_size && offs.size() < nb_hdr;) {
        uint32_t len = (uint32_t)in[at] | ((uint32_t)in[at + 1] << 8) | ((uint32_t)in[at + 2] << 16);
        const uint32_t ext = len >> 23;
        len &= 0x7FFFFFu;
        if (len < 3 || at + 3 + len > in_size) break;                         // turbosqueeze.cpp:119-136: stop at a short read
        const uint8_t* h = in + at + 3;
        const uint32_t u = (uint32_t)h[0] | ((uint32_t)h[1] << 8) | ((uint32_t)h[2] << 16);
        if (offs.empty()) { with_ext = ext; block = u; }
        else if (ext != with_ext) return -1;
        // Block b is decoded to b * block and the output copied back as one range: every block but the last must
        // decode to exactly `block` bytes (a 0-byte or short block in the middle goes to the one-shot path)
        if (short_seen || u > kBlockMax || u > block) return -1;
        if (u != block) short_seen = true;
        offs.push_back(at + 3); sizes.push_back(len);
        total += u;
        at += 3 + (uint64_t)len;
    }
    const uint64_t nb = offs.size();
    if (nb == 0 || block == 0) return -1;
    uint8_t* host = host_out;
    if (!host) {
        host = (uint8_t*)malloc(total + 128);                                 // tsq_threads.cpp:795 allocates outsize+128
        if (!host) return fail("decompress: malloc failed");
    } else if (total > host_cap) {
        return fail("decompress: output needs %llu bytes, caller gave %llu", (unsigned long long)total, (unsigned long long)host_cap);
    }
    // any failure from here on: wait for the streams that still use the context's buffers, release the output
    auto bail = [&](int rc) { cudaDeviceSynchronize(); if (!host_out) free(host); return rc; };
#define CUP(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) { fail("%s: %s", #call, cudaGetErrorString(e_)); return bail(1); }  \
    } while (0)
    if (c->cont.ensure(in_size + 512) || c->offs.ensure(nb * 8) || c->sizes.ensure(nb * 4) || c->osizes.ensure(nb * 4) ||
        c->out.ensure(nb * (uint64_t)block)) { fail("decompress: out of device memory"); return bail(1); }
    CUP(cudaMemcpyAsync(c->offs.p, offs.data(), nb * 8, cudaMemcpyHostToDevice, c->s_in));
    CUP(cudaMemcpyAsync(c->sizes.p, sizes.data(), nb * 4, cudaMemcpyHostToDevice, c->s_in));
    // Chunk boundaries: the output's trip home (U bytes of D2H) is the long leg and can only start once the first chunk
    // is decoded, so the first chunks are small (1 : 2 : 4 : 8 : 8 : ...) and the D2H stream starts ~1.5 ms earlier.
    constexpr int KMAX = tsqb_context::kPipe;
    uint64_t cbd[KMAX + 1];
    int nchunks = 0;
    {
        int want = K + 2 > KMAX ? KMAX : K + 2;
        double w[KMAX], sum = 0;
        for (int k = 0; k < want; k++) { w[k] = k < 3 ? (double)(1 << k) : 8.0; sum += w[k]; }
        double acc = 0;
        cbd[0] = 0;
        for (int k = 0; k < want; k++) {
            acc += w[k];
            uint64_t e = k == want - 1 ? nb : (uint64_t)(nb * (acc / sum) + 0.5);
            if (e > nb) e = nb;
            if (e > cbd[nchunks]) cbd[++nchunks] = e;
        }
        if (cbd[nchunks] < nb) cbd[nchunks] = nb;
    }
    for (int k = 0; k < nchunks; k++) {
        const uint64_t b0 = cbd[k], b1 = cbd[k + 1];
        const uint64_t lo = offs[b0], hi = offs[b1 - 1] + sizes[b1 - 1];
        CUP(cudaMemcpyAsync((uint8_t*)c->cont.p + lo, in + lo, hi - lo, cudaMemcpyHostToDevice, c->s_in));
        CUP(cudaEventRecord(c->ev_in[k], c->s_in));
        cudaStream_t st = c->s_chunk[k];
        CUP(cudaStreamWaitEvent(st, c->ev_in[k], 0));
        if (tsqb_decode_blocks(c, (uint8_t*)c->cont.p, (uint64_t*)c->offs.p + b0, 0, (uint32_t*)c->sizes.p + b0, b1 - b0,
                               (uint8_t*)c->out.p + b0 * (uint64_t)block, block, (uint32_t*)c->osizes.p + b0, with_ext, st)) return bail(1);
        CUP(cudaEventRecord(c->ev_done[k], st));
        CUP(cudaStreamWaitEvent(c->s_out, c->ev_done[k], 0));
        const uint64_t olo = b0 * (uint64_t)block, ohi = (b1 * (uint64_t)block < total) ? b1 * (uint64_t)block : total;
        CUP(cudaMemcpyAsync(host + olo, (uint8_t*)c->out.p + olo, ohi - olo, cudaMemcpyDeviceToHost, c->s_out));
        CUP(cudaEventRecord(c->ev_in[k], c->s_out));                          // reused: chunk k's bytes are home
    }
    for (int k = 0; k < nchunks; k++) {                                       // progress, chunk by chunk as the output lands
        CUP(cudaEventSynchronize(c->ev_in[k]));
        report_blocks(prog, cbd[k], cbd[k + 1], nb);
    }
    CUP(cudaStreamSynchronize(c->s_out));
#undef CUP
    if (out) *out = host;
    *out_size = total;
    return 0;
}

// One host buffer -> one TSQ1 container (pipelined above pipeline_min bytes).  `tail`: the bytes that follow the input in
// the caller's memory (the reference's memory path encodes in place and its last block reads them, tsq_threads.cpp:109).
static int compress_any(tsqb_context* c, const uint8_t* in, uint64_t total, const uint8_t* tail, uint32_t tail_n, uint32_t block,
                        uint32_t with_ext, uint8_t* host_out, uint64_t host_cap, uint8_t** out, uint64_t* out_size, const ProgressFn* prog)
{
    std::lock_guard<std::recursive_mutex> lk(c->mtx);
    CU(cudaSetDevice(c->device));
    if (c->pipeline && total >= c->pipeline_min) {
        if (c->stream_in) {
            const int r = compress_streamed(c, in, total, tail, tail_n, block, with_ext, host_out, host_cap, out, out_size, prog);
            if (r >= 0) return r;
        }
        return compress_pipelined(c, in, total, tail, tail_n, block, with_ext, host_out, host_cap, out, out_size, prog);
    }
    return compress_locked(c, in, total, tail, tail_n, block, with_ext, host_out, host_cap, out, out_size, prog);
}

extern "C" int tsqb_compress_buffer(tsqb_context* c, const uint8_t* in, uint64_t total, uint32_t block, uint32_t with_ext,
                                    uint8_t** out, uint64_t* out_size)
{
    if (!c || !out || !out_size) return fail("tsqb_compress_buffer: null argument");
    if (block == 0 || block > kBlockMax) return fail("tsqb_compress_buffer: bad block size %u", block);
    return compress_any(c, in, total, nullptr, 0, block, with_ext, nullptr, 0, out, out_size, nullptr);
}

extern "C" int tsqb_compress_into(tsqb_context* c, const uint8_t* in, uint64_t total, uint32_t block, uint32_t with_ext,
                                  uint8_t* out, uint64_t out_capacity, uint64_t* out_size)
{
    if (!c || !out || !out_size) return fail("tsqb_compress_into: null argument");
    if (block == 0 || block > kBlockMax) return fail("tsqb_compress_into: bad block size %u", block);
    return compress_any(c, in, total, nullptr, 0, block, with_ext, out, out_capacity, nullptr, out_size, nullptr);
}

static int decompress_locked(tsqb_context* c, const uint8_t* in, uint64_t in_size, uint8_t* host_out, uint64_t host_cap,
                             uint8_t** out, uint64_t* out_size, const ProgressFn* prog = nullptr)
{
    uint32_t nb_hdr; uint64_t total_hdr;
    memcpy(&nb_hdr, in + 4, 4); memcpy(&total_hdr, in + 8, 8);
    // the header's counts are not trusted beyond sizing (turbosqueeze.cpp:110-117 ignores them)
    const uint64_t max_blocks = (in_size - 16) / 4 + 1 < (uint64_t)nb_hdr + 1 ? (in_size - 16) / 4 + 1 : (uint64_t)nb_hdr + 1;
    if (c->cont.ensure(in_size + 512) || c->offs.ensure(max_blocks * 8) || c->sizes.ensure(max_blocks * 4) || c->ext.ensure(max_blocks * 4) ||
        c->osizes.ensure(max_blocks * 4) || c->misc.ensure(64))
        return fail("decompress: out of device memory");
    CU(cudaMemcpyAsync(c->cont.p, in, in_size, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync((uint8_t*)c->cont.p + in_size, 0, 512, c->stream));
    if (tsqb_index_container(c, (uint8_t*)c->cont.p, in_size, max_blocks, (uint64_t*)c->offs.p, (uint32_t*)c->sizes.p,
                             (uint32_t*)c->ext.p, (uint64_t*)c->misc.p, c->stream)) return 1;
    uint64_t nb = 0;
    CU(cudaMemcpyAsync(&nb, c->misc.p, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    // every block but the last must be a full block of the same size: read the first header to get it
    uint32_t with_ext = 0, block = 0;
    std::vector<uint64_t> offs(nb);
    std::vector<uint32_t> exts(nb);
    if (nb) {
        CU(cudaMemcpy(offs.data(), c->offs.p, nb * 8, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(exts.data(), c->ext.p, nb * 4, cudaMemcpyDeviceToHost));
        with_ext = exts[0];
        for (uint64_t b = 0; b < nb; b++) {
            if (exts[b] != with_ext) return fail("decompress: mixed extension flags in one container are not supported on the device path");
            const uint8_t* h = in + offs[b];
            const uint32_t u = (uint32_t)h[0] | ((uint32_t)h[1] << 8) | ((uint32_t)h[2] << 16);
            if (u > kBlockMax) return fail("decompress: block %llu declares %u bytes (> 4 MiB)", (unsigned long long)b, u);
            if (b == 0) block = u;
            else if (b + 1 < nb ? u != block : u > block) { block = 0; break; }
        }
        if (block == 0) block = kBlockMax;      // irregular container: one 4 MiB stride per block, compacted on the host
    }
    const uint64_t ostride = block;
    if (nb && c->out.ensure(nb * ostride)) return fail("decompress: out of device memory");
    if (nb && tsqb_decode_blocks(c, (uint8_t*)c->cont.p, (uint64_t*)c->offs.p, 0, (uint32_t*)c->sizes.p, nb, (uint8_t*)c->out.p, ostride,
                                 (uint32_t*)c->osizes.p, with_ext, c->stream)) return 1;
    std::vector<uint32_t> osz(nb);
    if (nb) CU(cudaMemcpyAsync(osz.data(), c->osizes.p, nb * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    uint64_t total = 0;
    bool packed = true;
    for (uint64_t b = 0; b < nb; b++) { if (b + 1 < nb && osz[b] != ostride) packed = false; total += osz[b]; }
    uint8_t* host = host_out;
    if (!host) {
        host = (uint8_t*)malloc(total + 128);                      // tsq_threads.cpp:795 allocates outsize+128
        if (!host) return fail("decompress: malloc failed");
    } else if (total > host_cap) {
        return fail("decompress: output needs %llu bytes, caller gave %llu", (unsigned long long)total, (unsigned long long)host_cap);
    }
    cudaError_t e = cudaSuccess;
    if (packed) {
        if (total) e = cudaMemcpyAsync(host, c->out.p, total, cudaMemcpyDeviceToHost, c->stream);
    } else {
        uint64_t at = 0;
        for (uint64_t b = 0; b < nb && e == cudaSuccess; b++) {
            if (osz[b]) e = cudaMemcpyAsync(host + at, (uint8_t*)c->out.p + b * ostride, osz[b], cudaMemcpyDeviceToHost, c->stream);
            at += osz[b];
        }
    }
    const cudaError_t e2 = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess || e2 != cudaSuccess) {
        if (!host_out) free(host);
        return fail("decompress: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    }
    report_blocks(prog, 0, nb, nb);
    if (out) *out = host;
    *out_size = total;
    (void)total_hdr;
    return 0;
}

static int decompress_any(tsqb_context* c, const uint8_t* in, uint64_t in_size, uint8_t* host_out, uint64_t host_cap, uint8_t** out,
                          uint64_t* out_size, const ProgressFn* prog)
{
    std::lock_guard<std::recursive_mutex> lk(c->mtx);
    CU(cudaSetDevice(c->device));
    if (c->pipeline && in_size >= c->pipeline_min) {
        const int r = decompress_pipelined(c, in, in_size, host_out, host_cap, out, out_size, prog);
        if (r >= 0) return r;
    }
    return decompress_locked(c, in, in_size, host_out, host_cap, out, out_size, prog);
}

extern "C" int tsqb_decompress_buffer(tsqb_context* c, const uint8_t* in, uint64_t in_size, uint8_t** out, uint64_t* out_size)
{
    if (!c || !in || !out || !out_size) return fail("tsqb_decompress_buffer: null argument");
    if (in_size < 16 || memcmp(in, "TSQ1", 4) != 0) return fail("tsqb_decompress_buffer: not a TSQ1 container");
    return decompress_any(c, in, in_size, nullptr, 0, out, out_size, nullptr);
}

extern "C" int tsqb_decompress_into(tsqb_context* c, const uint8_t* in, uint64_t in_size, uint8_t* out, uint64_t out_capacity,
                                    uint64_t* out_size)
{
    if (!c || !in || !out || !out_size) return fail("tsqb_decompress_into: null argument");
    if (in_size < 16 || memcmp(in, "TSQ1", 4) != 0) return fail("tsqb_decompress_into: not a TSQ1 container");
    return decompress_any(c, in, in_size, out, out_capacity, nullptr, out_size, nullptr);
}

// ------------------------------------------------------------- layer 2: the reference's entry points
// Block size of the container-producing entry points (tsqCompress, tsqCompress_MT, tsqCompressAsync_MT).  The
// reference cuts 4 MiB blocks (turbosqueeze.h:37-38) and so does the default here, which makes the containers
// identical; the container does not record the block size and the reference's decoder accepts any block <= 4 MiB
// (tsq_decode.cpp:53), so a smaller block stays readable by `tsq d` while giving the GPU 16x more blocks in flight.
static uint32_t container_block_default()
{
    // TSQB_CONTAINER_BLOCK=<bytes>: the same switch as tsqb_set_container_block_size for a caller that cannot be recompiled
    const char* e = getenv("TSQB_CONTAINER_BLOCK");
    if (e && *e) {
        const unsigned long v = strtoul(e, nullptr, 0);
        if (v >= 1 && v <= kBlockMax) return (uint32_t)v;
        fprintf(stderr, "turbosqueeze_b200: TSQB_CONTAINER_BLOCK=%s ignored (not in 1..%u)\n", e, kBlockMax);
    }
    return kBlockMax;
}
static std::atomic<uint32_t> g_container_block{container_block_default()};

extern "C" int tsqb_set_container_block_size(uint32_t block_size)
{
    if (block_size == 0 || block_size > kBlockMax) return fail("tsqb_set_container_block_size: %u not in 1..%u", block_size, kBlockMax);
    g_container_block = block_size;
    return 0;
}

static tsqb_context* g_default = nullptr;
static std::mutex g_default_mtx;

static tsqb_context* default_context()
{
    std::lock_guard<std::mutex> lk(g_default_mtx);
    if (!g_default && tsqb_create(&g_default, 0) != 0) {
        fprintf(stderr, "turbosqueeze_b200: %s\n", tsqb_last_error());
        g_default = nullptr;
    }
    return g_default;
}

// tsq_context.cpp:56-74 -- the table is kept (128-byte aligned, 256 KiB) because callers poke it
// (test/test.cpp:42); the device path defines the table as zero at entry and never reads this copy.
extern "C" struct TSQCompressionContext* tsqAllocateContext(void)
{
    TSQCompressionContext* ctx = (TSQCompressionContext*)malloc(sizeof(TSQCompressionContext));
    if (!ctx) return nullptr;
    ctx->refhash = (uint16_t*)aligned_alloc(128, TSQB_HASH_BYTES);
    if (!ctx->refhash) { free(ctx); return nullptr; }
    return ctx;
}

extern "C" void tsqDeallocateContext(struct TSQCompressionContext* ctx)
{
    if (!ctx) return;
    free(ctx->refhash);
    free(ctx);
}

extern "C" void tsqInit(struct TSQCompressionContext* ctx)          // tsq_context.cpp:77-80
{
    if (ctx && ctx->refhash) memset(ctx->refhash, 0, TSQB_HASH_BYTES);
}

extern "C" void tsqEncode(struct TSQCompressionContext* ctx, uint8_t* in, uint8_t* out, uint32_t* outputSize, uint32_t inputSize,
                          uint32_t withExtensions)
{
    if (outputSize) *outputSize = 0;
    tsqb_context* c = default_context();
    if (!c || !in || !out || !outputSize || inputSize == 0 || inputSize > kBlockMax) return;
    // Documented divergence: the device path DEFINES the table as zero at entry (every reference caller runs tsqInit or
    // memsets refhash first).  A caller that skips that would get different bytes from the reference: say so, once.
    if (ctx && ctx->refhash) {
        static std::atomic<bool> warned{false};
        const uint64_t* w = reinterpret_cast<const uint64_t*>(ctx->refhash);
        uint64_t any = 0;
        for (size_t q = 0; q < TSQB_HASH_BYTES / 8; q++) any |= w[q];
        if (any && !warned.exchange(true))
            fprintf(stderr, "turbosqueeze_b200: tsqEncode called with a non-zero refhash table; the device path encodes as if tsqInit had been called\n");
    }
    std::lock_guard<std::recursive_mutex> lk(c->mtx);
    const uint64_t stride = tsqb_slot_stride(inputSize);
    // the reference reads <= 19 bytes past the block (72 with extensions): ship exactly those
    const uint32_t tail_n = withExtensions ? 72u : 19u;
    if (cudaSetDevice(c->device) != cudaSuccess || stage_input(c, in, inputSize, in + inputSize, tail_n)) goto bad;
    if (c->slots.ensure(stride) || c->sizes.ensure(8)) goto bad;
    if (encode_blocks_impl(c, (uint8_t*)c->in.p, inputSize, inputSize, (uint8_t*)c->slots.p, stride, (uint32_t*)c->sizes.p,
                           (uint32_t*)c->sizes.p + 1, withExtensions, c->stream)) goto bad;
    {
        uint32_t nf[2] = {0, 0};                                     // size, kTail* flags
        if (cudaMemcpyAsync(nf, c->sizes.p, 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) goto bad;
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) goto bad;
        const uint32_t n = nf[0];
        // The last two bytes may be ones the reference never initialises (tsq_encode.cpp:176-188):
        // there the caller's own pre-fill of `outputBlock` must survive, exactly as in the reference.
        uint8_t tail[2] = {0, 0};
        const uint32_t body = n >= 2 ? n - 2 : 0;
        if (body && cudaMemcpy(out, c->slots.p, body, cudaMemcpyDeviceToHost) != cudaSuccess) goto bad;
        if (cudaMemcpy(tail, (uint8_t*)c->slots.p + body, n - body, cudaMemcpyDeviceToHost) != cudaSuccess) goto bad;
        if (n >= 2) {
            if (!(nf[1] & 1u)) out[n - 2] = tail[0];
            if (nf[1] & 4u) out[n - 1] = (uint8_t)(out[n - 1] << 4);
            else if (!(nf[1] & 2u)) out[n - 1] = tail[1];
        } else if (n == 1) out[0] = tail[0];
        *outputSize = n;
    }
    return;
bad:
    cudaGetLastError();
    fprintf(stderr, "turbosqueeze_b200: tsqEncode failed: %s\n", tsqb_last_error());
}

static uint64_t stream_extent(const uint8_t* in, uint32_t size, bool ext);

extern "C" void tsqDecode(uint8_t* in, uint8_t* out, uint32_t* outputSize, uint32_t inputSize, uint32_t withExtensions)
{
    if (outputSize) *outputSize = 0;
    tsqb_context* c = default_context();
    if (!c || !in || !out || !outputSize) return;
    const uint32_t size = (uint32_t)in[0] | ((uint32_t)in[1] << 8) | ((uint32_t)in[2] << 16);
    if (size > kBlockMax) return;                                    // tsq_decode.cpp:53
    std::lock_guard<std::recursive_mutex> lk(c->mtx);
    // the reference ignores inputSize (tsq_decode.cpp:42-126); without it the stream is measured by its own size bytes,
    // so that nothing behind the stream is read from the caller's buffer
    uint64_t n_in = inputSize ? inputSize : stream_extent(in, size, withExtensions != 0);
    if (cudaSetDevice(c->device) != cudaSuccess) goto bad;
    if (c->slots.ensure(n_in + 512) || c->sizes.ensure(4) || c->out.ensure((uint64_t)size + 16) || c->osizes.ensure(4)) goto bad;
    if (cudaMemsetAsync((uint8_t*)c->slots.p + n_in, 0, 512, c->stream) != cudaSuccess) goto bad;
    if (cudaMemcpyAsync(c->slots.p, in, n_in, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) goto bad;
    {
        const uint32_t lim = (uint32_t)(n_in + 256);
        if (cudaMemcpyAsync(c->sizes.p, &lim, 4, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) goto bad;
    }
    if (tsqb_decode_blocks(c, (uint8_t*)c->slots.p, nullptr, 0, (uint32_t*)c->sizes.p, 1, (uint8_t*)c->out.p, size ? size : 1,
                           (uint32_t*)c->osizes.p, withExtensions, c->stream)) goto bad;
    if (size && cudaMemcpyAsync(out, c->out.p, size, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) goto bad;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) goto bad;
    *outputSize = size;                                              // tsq_decode.cpp:125
    return;
bad:
    cudaGetLastError();
    fprintf(stderr, "turbosqueeze_b200: tsqDecode failed: %s\n", tsqb_last_error());
}

// Length of a block's token stream, found the way the reference's decoder finds it: by walking the control and size
// bytes (tsq_decode.cpp:60-123; :137-314 with extensions).  Touches exactly the bytes the reference would read, never
// one past them, and copies nothing.  Used only when the caller passes inputSize == 0 (the reference ignores inputSize).
static uint64_t stream_extent(const uint8_t* in, uint32_t size, bool ext)
{
    uint64_t i = 3;
    uint32_t j = 0;
    while (j < size) {
        const uint32_t ctl = in[i++];
        for (int p = 0; p < 4 && j < size; p++) {
            const uint32_t nib = in[i++];
            const uint32_t n0 = nib >> 4, n1 = nib & 15u;
            const bool l0 = (ctl & (0x80u >> (2 * p))) != 0, l1 = (ctl & (0x40u >> (2 * p))) != 0;
            i += l0 ? n0 + 1u : 2u;
            j += (ext && !l0 && n0 < 3u) ? 16u * (n0 + 2u) : n0 + 1u;
            if (j >= size) break;                                    // the second symbol is padding: nothing of it is read
            i += l1 ? n1 + 1u : 2u;
            j += (ext && !l1 && n1 < 3u) ? 16u * (n1 + 2u) : n1 + 1u;
        }
    }
    return i;
}

static bool read_all(FILE* f, std::vector<uint8_t>& v)
{
    if (fseek(f, 0, SEEK_END) == 0) {                                // turbosqueeze.cpp:57-59
        const long n = ftell(f);
        if (n < 0 || fseek(f, 0, SEEK_SET) != 0) return false;
        v.resize((size_t)n);
        return n == 0 || fread(v.data(), 1, (size_t)n, f) == (size_t)n;
    }
    uint8_t buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) v.insert(v.end(), buf, buf + n);
    return true;
}

extern "C" void tsqCompress(FILE* in, FILE* out, bool useextensions, uint32_t level)
{
    (void)level;                                                     // never read by the reference either
    tsqb_context* c = default_context();
    if (!c || !in || !out) return;
    std::vector<uint8_t> data;
    if (!read_all(in, data)) return;
    uint8_t* blob = nullptr; uint64_t n = 0;
    if (tsqb_compress_buffer(c, data.data(), data.size(), g_container_block.load(), useextensions ? 1u : 0u, &blob, &n) != 0) {
        fprintf(stderr, "turbosqueeze_b200: tsqCompress failed: %s\n", tsqb_last_error());
        return;
    }
    fwrite(blob, 1, n, out);
    free(blob);
}

extern "C" void tsqDecompress(FILE* in, FILE* out)
{
    tsqb_context* c = default_context();
    if (!c || !in || !out) return;
    std::vector<uint8_t> data;
    if (!read_all(in, data)) return;
    if (data.size() < 16 || memcmp(data.data(), "TSQ1", 4) != 0) return;   // turbosqueeze.cpp:107-109
    uint8_t* blob = nullptr; uint64_t n = 0;
    if (tsqb_decompress_buffer(c, data.data(), data.size(), &blob, &n) != 0) {
        fprintf(stderr, "turbosqueeze_b200: tsqDecompress failed: %s\n", tsqb_last_error());
        return;
    }
    fwrite(blob, 1, n, out);
    free(blob);
}

// ---- buffer API and asynchronous job API (tsq_threads.cpp:278-441, :679-890)
// The reference pushes the blocks of consecutive jobs through one pipeline of reader / worker / writer threads
// (tsq_threads.cpp:52-275): jobs overlap in flight, and results and callbacks leave strictly in submission order on the
// writer thread (:199,:611) -- one progress callback per written block (:248-254,:654-655), then the completion
// callback (:256-268).  Here a context owns kLanes job threads, each with a device context of its own (streams +
// scratch), which take jobs from one queue and run them concurrently on the GPU.  The CALLBACKS are delivered in
// submission order: what a job reports while an older job is still unfinished is held back until it is the oldest.
namespace {
struct JobEngine {
    static constexpr int kLanes = 2;
    struct Job {
        uint32_t id = 0;
        uint64_t seq = 0;
        std::function<bool(tsqb_context*, const ProgressFn*)> run;        // the device work; true = success
        std::function<void(uint32_t, bool)> completion_cb;
        std::function<void(uint32_t, double)> progress_cb;
    };

    std::mutex m;                       // queue
    std::condition_variable cv;
    std::deque<Job> jobs;
    bool stop = false, lane0_busy = false;
    uint32_t next_id = 1;
    uint64_t next_seq = 0;
    std::mutex dm;                      // ordered delivery of callbacks
    std::condition_variable dcv;
    uint64_t head_seq = 0;              // oldest job whose completion callback has not run yet
    tsqb_context* dev[kLanes] = {};
    std::thread th[kLanes];

    explicit JobEngine(tsqb_context* dev0)
    {
        dev[0] = dev0;
        for (int l = 0; l < kLanes; l++) th[l] = std::thread([this, l] { lane_main(l); });
    }
    ~JobEngine()                        // drains: tsq_context.cpp:150-155 waits for the jobs in flight
    {
        { std::lock_guard<std::mutex> lk(m); stop = true; }
        cv.notify_all();
        for (int l = 0; l < kLanes; l++) th[l].join();
        for (int l = 1; l < kLanes; l++) if (dev[l]) tsqb_destroy(dev[l]);
    }
    uint32_t submit(Job j)
    {
        std::lock_guard<std::mutex> lk(m);
        j.id = next_id++;
        if (next_id == 0) next_id = 1;
        j.seq = next_seq++;
        const uint32_t id = j.id;
        jobs.push_back(std::move(j));
        cv.notify_all();
        return id;
    }
    void lane_main(int lane)
    {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(m);
                // lane 0 takes whatever comes; the other lanes only help while lane 0 is busy (a context that sees one
                // job at a time never creates a second device context)
                cv.wait(lk, [&] { return (!jobs.empty() && (lane == 0 || lane0_busy)) || (stop && jobs.empty()); });
                if (jobs.empty()) return;                            // stop requested and nothing left
                j = std::move(jobs.front());
                jobs.pop_front();
                if (lane == 0) lane0_busy = true;
            }
            if (!dev[lane] && tsqb_create(&dev[lane], dev[0]->device) != 0) dev[lane] = nullptr;
            std::vector<double> held;                                // progress reported while an older job was unfinished
            ProgressFn progress = [&](uint64_t done, uint64_t nb) {
                if (!j.progress_cb) return;
                const double v = nb ? (double)done / (double)nb : 1.0;
                std::lock_guard<std::mutex> lk(dm);
                if (head_seq != j.seq) { held.push_back(v); return; }
                for (double h : held) j.progress_cb(j.id, h);
                held.clear();
                j.progress_cb(j.id, v);
            };
            const bool ok = dev[lane] != nullptr && j.run(dev[lane], &progress);
            {
                std::unique_lock<std::mutex> lk(dm);
                dcv.wait(lk, [&] { return head_seq == j.seq; });
                if (j.progress_cb) for (double h : held) j.progress_cb(j.id, h);
                if (j.completion_cb) j.completion_cb(j.id, ok);     // may submit jobs to other contexts (test/test.cpp:247-262)
                head_seq++;
            }
            dcv.notify_all();
            if (lane == 0) { std::lock_guard<std::mutex> lk(m); lane0_busy = false; }
            cv.notify_all();
        }
    }
};

// [p, p + n) are the bytes BEHIND a caller's buffer.  The reference's memory path encodes in place, so its last block
// simply reads them (tsq_threads.cpp:109; <= 19 bytes, 72 with extensions) and they shape the last block's stream.  Same
// here -- unless a page of that range is not mapped at all, where the reference would have faulted: then zeros.
static bool tail_is_mapped(const uint8_t* p, size_t n)
{
    const uintptr_t ps = (uintptr_t)sysconf(_SC_PAGESIZE);
    const uintptr_t own = ((uintptr_t)p - 1) & ~(ps - 1);           // page of the buffer's last byte: mapped by definition
    for (uintptr_t page = (uintptr_t)p & ~(ps - 1); page <= (((uintptr_t)p + n - 1) & ~(ps - 1)); page += ps) {
        unsigned char vec;
        if (page != own && mincore((void*)page, ps, &vec) != 0) return false;
    }
    return true;
}
}  // namespace

struct TSQCompressionContext_MT   { tsqb_context* dev; bool verbose; JobEngine* jobs; };
struct TSQDecompressionContext_MT { tsqb_context* dev; bool verbose; JobEngine* jobs; };

extern "C" struct TSQCompressionContext_MT* tsqAllocateContextCompression_MT(bool verbose)
{
    tsqb_context* d = nullptr;
    if (tsqb_create(&d, 0) != 0) { if (verbose) printf("Error: %s\n", tsqb_last_error()); return nullptr; }
    return new TSQCompressionContext_MT{d, verbose, new JobEngine(d)};
}

extern "C" void tsqDeallocateContextCompression_MT(struct TSQCompressionContext_MT* ctx)
{
    if (!ctx) return;
    delete ctx->jobs;                                                // waits for the queued jobs
    tsqb_destroy(ctx->dev);
    delete ctx;
}

extern "C" struct TSQDecompressionContext_MT* tsqAllocateContextDecompression_MT(bool verbose)
{
    tsqb_context* d = nullptr;
    if (tsqb_create(&d, 0) != 0) { if (verbose) printf("Error: %s\n", tsqb_last_error()); return nullptr; }
    return new TSQDecompressionContext_MT{d, verbose, new JobEngine(d)};
}

extern "C" void tsqDeallocateContextDecompression_MT(struct TSQDecompressionContext_MT* ctx)
{
    if (!ctx) return;
    delete ctx->jobs;
    tsqb_destroy(ctx->dev);
    delete ctx;
}

static bool load_input(uint8_t* in, size_t szin, bool infile, std::vector<uint8_t>& store, const uint8_t** p, size_t* n, bool verbose)
{
    if (!infile) { *p = in; *n = szin; return true; }
    FILE* f = fopen((const char*)in, "rb");                          // tsq_threads.cpp:294
    if (!f) { if (verbose) printf("Error: could not open input file.\n"); return false; }
    const bool ok = read_all(f, store);
    fclose(f);
    *p = store.data(); *n = store.size();
    return ok;
}

static bool store_output(uint8_t* blob, uint64_t n, uint8_t** out, size_t* szout, bool outfile, bool verbose)
{
    if (!outfile) { *out = blob; *szout = (size_t)n; return true; }
    FILE* f = fopen((const char*)*out, "wb");                        // tsq_threads.cpp:319
    bool ok = f != nullptr;
    if (!f && verbose) printf("Error: could not open output file.\n");
    if (f) { ok = fwrite(blob, 1, n, f) == n; fclose(f); }
    free(blob);
    if (ok && szout) *szout = (size_t)n;
    return ok;
}

// one compression job on device context `dev` (tsq_threads.cpp:278-410 + the pipeline behind it)
static bool compress_job(tsqb_context* dev, bool verbose, uint8_t* in, size_t szin, bool infile, uint8_t** out, size_t* szout, bool outfile,
                         bool useextensions, const ProgressFn* prog)
{
    std::vector<uint8_t> store;
    const uint8_t* p; size_t n;
    if (!load_input(in, szin, infile, store, &p, &n, verbose)) return false;
    // memory input: the last block sees the caller's bytes behind the buffer, as in the reference (tsq_threads.cpp:109)
    const uint32_t tail_n = useextensions ? 72u : 19u;
    const uint8_t* tail = (!infile && n && tail_is_mapped(p + n, tail_n)) ? p + n : nullptr;
    uint8_t* blob = nullptr; uint64_t bn = 0;
    if (compress_any(dev, p, n, tail, tail ? tail_n : 0u, g_container_block.load(), useextensions ? 1u : 0u, nullptr, 0, &blob, &bn, prog) != 0) {
        if (verbose) printf("Error: %s\n", tsqb_last_error());
        return false;
    }
    return store_output(blob, bn, out, szout, outfile, verbose);
}

static bool decompress_job(tsqb_context* dev, bool verbose, uint8_t* in, size_t szin, bool infile, uint8_t** out, size_t* szout, bool outfile,
                           const ProgressFn* prog)
{
    std::vector<uint8_t> store;
    const uint8_t* p; size_t n;
    if (!load_input(in, szin, infile, store, &p, &n, verbose)) return false;
    if (n < 16 || memcmp(p, "TSQ1", 4) != 0) { if (verbose) printf("Error: not a TSQ1 container.\n"); return false; }   // tsq_threads.cpp:728-730
    uint8_t* blob = nullptr; uint64_t bn = 0;
    if (decompress_any(dev, p, n, nullptr, 0, &blob, &bn, prog) != 0) {
        if (verbose) printf("Error: %s\n", tsqb_last_error());
        return false;
    }
    return store_output(blob, bn, out, szout, outfile, verbose);
}

// ---- asynchronous job API (tsq_threads.cpp:278-410, :679-859)
extern "C" uint32_t tsqCompressAsync_MT(struct TSQCompressionContext_MT* ctx, uint8_t* in, size_t szin, bool infile, uint8_t** out,
                                        size_t* szout, bool outfile, bool useextensions, uint32_t level,
                                        std::function<void(uint32_t, bool)> completion_cb, std::function<void(uint32_t, double)> progress_cb)
{
    (void)level;                                                     // stored and never read by the reference (tsq_threads.cpp:98,112)
    if (!ctx || !in || (!infile && szin == 0) || !out || !szout) {    // early failure: completion(0, false), id 0 (:296-306)
        if (completion_cb) completion_cb(0, false);
        return 0;
    }
    JobEngine::Job j;
    const bool verbose = ctx->verbose;
    j.run = [=](tsqb_context* dev, const ProgressFn* prog) { return compress_job(dev, verbose, in, szin, infile, out, szout, outfile, useextensions, prog); };
    j.completion_cb = [=](uint32_t id, bool ok) {
        if (verbose) printf("Compression job %u %s\n", id, ok ? "done" : "failed");
        if (completion_cb) completion_cb(id, ok);
    };
    j.progress_cb = progress_cb;
    return ctx->jobs->submit(std::move(j));
}

extern "C" uint32_t tsqDecompressAsync_MT(struct TSQDecompressionContext_MT* ctx, uint8_t* in, size_t szin, bool infile, uint8_t** out,
                                          size_t* szout, bool outfile, std::function<void(uint32_t, bool)> completion_cb,
                                          std::function<void(uint32_t, double)> progress_cb)
{
    if (!ctx || !in || (!infile && szin == 0) || !out || !szout) {
        if (completion_cb) completion_cb(0, false);
        return 0;
    }
    JobEngine::Job j;
    const bool verbose = ctx->verbose;
    j.run = [=](tsqb_context* dev, const ProgressFn* prog) { return decompress_job(dev, verbose, in, szin, infile, out, szout, outfile, prog); };
    j.completion_cb = [=](uint32_t id, bool ok) {
        if (verbose) printf("Decompression job %u %s\n", id, ok ? "done" : "failed");
        if (completion_cb) completion_cb(id, ok);
    };
    j.progress_cb = progress_cb;
    return ctx->jobs->submit(std::move(j));
}

// ---- synchronous buffer API: submit, then wait for the completion callback -- exactly what the reference does
// (tsq_threads.cpp:413-441, :862-890)
namespace {
struct Waiter {
    std::mutex m; std::condition_variable cv; bool done = false, ok = false;
    void set(bool v) { { std::lock_guard<std::mutex> lk(m); done = true; ok = v; } cv.notify_all(); }
    bool wait() { std::unique_lock<std::mutex> lk(m); cv.wait(lk, [&] { return done; }); return ok; }
};
}  // namespace

extern "C" bool tsqCompress_MT(struct TSQCompressionContext_MT* ctx, uint8_t* in, size_t szin, bool infile, uint8_t** out, size_t* szout,
                               bool outfile, bool useextensions, uint32_t level)
{
    if (!ctx || !in || szin == 0 || !out || szout == 0) return false;          // tsq_threads.cpp:415-418
    Waiter w;
    tsqCompressAsync_MT(ctx, in, szin, infile, out, szout, outfile, useextensions, level, [&w](uint32_t, bool ok) { w.set(ok); }, nullptr);
    return w.wait();
}

extern "C" bool tsqDecompress_MT(struct TSQDecompressionContext_MT* ctx, uint8_t* in, size_t szin, bool infile, uint8_t** out, size_t* szout,
                                 bool outfile)
{
    if (!ctx || !in || szin == 0 || !out || szout == 0) return false;          // tsq_threads.cpp:864-867
    Waiter w;
    tsqDecompressAsync_MT(ctx, in, szin, infile, out, szout, outfile, [&w](uint32_t, bool ok) { w.set(ok); }, nullptr);
    return w.wait();
}
