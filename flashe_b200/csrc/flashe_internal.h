// flashe_internal.h — host-side glue shared by the translation units of libflashe_b200.so
// (not part of the C ABI; include/flashe_b200.h is).
#ifndef FLASHE_INTERNAL_H
#define FLASHE_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/flashe_b200.h"

// Records `msg` as the calling thread's last error and returns `code`.
int flashe_fail(int code, const std::string& msg);
// Bench bookkeeping behind flashe_launch_count().
void flashe_count_launches(int n);
// What the other translation units need to know about a context.
struct flashe_ctx_info { int device; int int_bits; int words; int num_sms; };
int flashe_ctx_get_info(const flashe_ctx* ctx, flashe_ctx_info* out);

struct alignas(16) KeySched { uint32_t rk[60]; };   // AES-256 round keys, big-endian words

struct flashe_ctx {
    int device;
    int int_bits;
    int words;       // 1, 2, 4
    int num_sms;
    uint32_t m;
    uint8_t key[32];
    KeySched ks;
    uint32_t* d_te0;   // device copy of the Te0 table (256 words): source of the kernels' shared-memory tables
    uint32_t* d_tickets;              // (ticket, done) counter pairs of k_stream's dynamic deal, all zero between launches
    struct flashe_ticket_state* tickets;   // host bookkeeping of the slots (flashe_kernels.cu)
};

// Counter pair for one k_stream launch on `stream`, or NULL (static deal: FLASHE_DYNAMIC=0, slots exhausted, per-thread
// default stream).  Launches that may run concurrently never share a slot: eager launches get the slot of their stream
// (one stream serialises its launches; a dependent launch only draws tickets after griddepcontrol.wait), every launch
// recorded during stream capture gets a slot of its own for the life of the context.
uint32_t* flashe_ticket_slot(const flashe_ctx* ctx, cudaStream_t stream);


int flashe_check_span(const flashe_span* s);
// flashe_elementwise.cu: dense[index_c[p]] += (subtract: -=) compact_c[p] mod 2^int_bits for every client c, in place, built
// tile by tile in shared memory (index lists sorted unique, ks[c] < 2^32; not capturable: uploads a table)
int flashe_sparse_accumulate_tiled(flashe_ctx* ctx, const void* const* compacts, const int64_t* const* indexes, const uint64_t* ks, int n,
                                   int subtract, uint64_t total, void* dense, cudaStream_t cs);
static inline uint64_t ceil_div(uint64_t a, uint64_t b) { return (a + b - 1) / b; }
static inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }
static inline int grid_1d(const flashe_ctx* ctx, uint64_t work_items, int threads, int per_sm) {
    uint64_t blocks = ceil_div(work_items ? work_items : 1, (uint64_t)threads);
    uint64_t cap = (uint64_t)ctx->num_sms * per_sm;
    return (int)(blocks < cap ? blocks : cap);
}

// Grid of a grid-stride kernel: exactly the CTAs that are resident at once (occupancy x SMs), or fewer when the
// work is small.  A fixed "SMs x 8" grid with a kernel that fits only 5 CTAs per SM runs 1.6 waves, the second
// at 60 % occupancy (decode lost 20 % to that).  The occupancy is queried once per kernel and device.
int flashe_resident_ctas(const void* kernel, int threads, size_t dyn_smem, int device, int num_sms);
static inline int grid_occ(int num_sms, int device, const void* kernel, uint64_t work_items, int threads, size_t dyn_smem = 0) {
    uint64_t blocks = (work_items + (uint64_t)threads - 1) / (uint64_t)threads;
    if (blocks < 1) blocks = 1;
    const uint64_t cap = (uint64_t)flashe_resident_ctas(kernel, threads, dyn_smem, device, num_sms);
    return (int)(blocks < cap ? blocks : cap);
}
#define GRID_OCC(ctx, kernel, work, threads) grid_occ((ctx)->num_sms, (ctx)->device, (const void*)(kernel), (work), (threads))

#define FLASHE_CUDA_TRY(expr)                                                                            \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return flashe_fail(FLASHE_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));       \
    } while (0)

struct FlasheDeviceGuard {
    int prev;
    bool ok;
    explicit FlasheDeviceGuard(int dev) : prev(-1), ok(true) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~FlasheDeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

#define CUDA_TRY(expr) FLASHE_CUDA_TRY(expr)
#define ENTER(ctx)                                                                 \
    if (!(ctx)) return flashe_fail(FLASHE_EINVAL, "ctx is NULL");                  \
    FlasheDeviceGuard guard__((ctx)->device);                                      \
    if (!guard__.ok) return flashe_fail(FLASHE_ECUDA, "cudaSetDevice failed");     \
    cudaStream_t cs = (cudaStream_t)stream

#endif
