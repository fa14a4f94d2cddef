// flashe_kernels.cu — B200 (sm_100a) kernels and C ABI of the FLASHE hot path.
//
// Path (SamuelGong/FLASHE, federatedml/secureprotol/jzf_flashe.py + jzf_quantize.py, and the server
// sum of framework/homo/procedure/jzf_aggregator.py:404-430); see include/flashe_b200.h for the
// per-entry-point citations and DESIGN.md for the layout/roofline discussion.
//
// Kernel families
//   k_stream<WORDS, MMAX, MODE, SHARE, ALIGNED>   persistent, one 512-thread CTA per SM.  A warp owns
//       work units of consecutive "items" of one reference chunk; an item = 64 AES blocks whose
//       counters share their upper 56 bits (lane <-> blocks lane and lane + 32).  Lanes run AES-256
//       from a bank-conflict-free shared-memory T-table (4 tables x 32 replicas = 128 KB, one LDS + one
//       PRMT per lookup, 197 lookups per block: round 1 hoisted on the host, rounds 1-2 factored per
//       counter window), reduce the signed streams into a combined mask in registers and apply it,
//       fused with encode / decode.  Full items of the common layouts (4-byte words with m = 4, 5, 6;
//       16-byte words) run in a lane-local loop with 16-byte global accesses; everything else (edge
//       items, other widths, unaligned buffers) goes through a per-warp slab and element pairs.
//   elementwise kernels           aggregate (element-wise and packed-carry), premasked add, encode,
//       decode, lane batching, sparse expand / sum, noise.
//   flashe_wire.cu                wire bit-packing, layer-wise top-k sparsify, per-layer statistics.
//
// No tensor cores: nothing here is a dense contraction (integer PRF + modular adds).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <tuple>
#include <string>
#include <vector>

#include "../../include/flashe_b200.h"
#include "flashe_internal.h"
#include "flashe_device.cuh"
#include "flashe_codec_host.cuh"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const std::string& msg) { g_err = msg; return code; }
int flashe_fail(int code, const std::string& msg) { return fail(code, msg); }
void flashe_count_launches(int n) { g_launches.fetch_add((uint64_t)n); }
// ------------------------------------------------------------------------------------------------
// host AES-256 (key schedule, T-table, round-1 hoisting).  FIPS-197; big-endian word convention:
// word = b0<<24 | b1<<16 | b2<<8 | b3, so the reference's block iter(4B BE)||prf(4B BE)||ctr(8B BE)
// (jzf_flashe.py:34,304,308) is simply the words {iter, prf, ctr>>32, ctr&0xffffffff}.
// ------------------------------------------------------------------------------------------------
namespace haes {
static uint8_t sbox[256];
static uint32_t te0[256];
static bool ready = false;

static uint8_t xtime(uint8_t a) { return (uint8_t)((a << 1) ^ ((a & 0x80) ? 0x1b : 0)); }
static uint8_t mul(uint8_t a, uint8_t b) {
    uint8_t p = 0;
    while (b) { if (b & 1) p ^= a; a = xtime(a); b >>= 1; }
    return p;
}
static void init() {
    if (ready) return;
    // multiplicative inverse via exponentiation tables on generator 3
    uint8_t exp[256], log[256];
    uint8_t x = 1;
    for (int i = 0; i < 255; ++i) { exp[i] = x; log[x] = (uint8_t)i; x = (uint8_t)(x ^ xtime(x)); }
    exp[255] = exp[0];
    for (int v = 0; v < 256; ++v) {
        uint8_t inv = v ? exp[(255 - log[v]) % 255] : 0;
        uint8_t s = inv;
        for (int i = 1; i < 5; ++i) s ^= (uint8_t)((inv << i) | (inv >> (8 - i)));
        sbox[v] = s ^ 0x63;
    }
    for (int v = 0; v < 256; ++v) {
        uint8_t s = sbox[v], s2 = xtime(s), s3 = (uint8_t)(s2 ^ s);
        te0[v] = ((uint32_t)s2 << 24) | ((uint32_t)s << 16) | ((uint32_t)s << 8) | s3;
    }
    ready = true;
}
static inline uint32_t ror(uint32_t v, int n) { return n ? ((v >> n) | (v << (32 - n))) : v; }
static inline uint32_t te(int t, uint32_t idx) { return ror(te0[idx & 0xff], 8 * t); }

static void expand(const uint8_t key[32], uint32_t rk[60]) {
    init();
    for (int i = 0; i < 8; ++i)
        rk[i] = ((uint32_t)key[4 * i] << 24) | ((uint32_t)key[4 * i + 1] << 16) | ((uint32_t)key[4 * i + 2] << 8) | key[4 * i + 3];
    uint32_t rcon = 1;
    for (int i = 8; i < 60; ++i) {
        uint32_t t = rk[i - 1];
        if (i % 8 == 0) {
            t = (t << 8) | (t >> 24);
            t = ((uint32_t)sbox[t >> 24] << 24) | ((uint32_t)sbox[(t >> 16) & 0xff] << 16) | ((uint32_t)sbox[(t >> 8) & 0xff] << 8) | sbox[t & 0xff];
            t ^= rcon << 24;
            rcon = mul((uint8_t)rcon, 2);
        } else if (i % 8 == 4) {
            t = ((uint32_t)sbox[t >> 24] << 24) | ((uint32_t)sbox[(t >> 16) & 0xff] << 16) | ((uint32_t)sbox[(t >> 8) & 0xff] << 8) | sbox[t & 0xff];
        }
        rk[i] = rk[i - 8] ^ t;
    }
}
// Round-1 terms that do not depend on the low counter word (input words w0,w1,w2 fixed).
static void hoist_round1(const uint32_t rk[60], uint32_t w0, uint32_t w1, uint32_t w2, uint32_t pre[4]) {
    uint32_t s0 = w0 ^ rk[0], s1 = w1 ^ rk[1], s2 = w2 ^ rk[2];
    pre[0] = te(0, s0 >> 24) ^ te(1, s1 >> 16) ^ te(2, s2 >> 8) ^ rk[4];
    pre[1] = te(0, s1 >> 24) ^ te(1, s2 >> 16) ^ te(3, s0) ^ rk[5];
    pre[2] = te(0, s2 >> 24) ^ te(2, s0 >> 8) ^ te(3, s1) ^ rk[6];
    pre[3] = te(1, s0 >> 16) ^ te(2, s1 >> 8) ^ te(3, s2) ^ rk[7];
}
}  // namespace haes

#include "flashe_stream_decl.h"

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int words_of(int b) { return b <= 32 ? 1 : (b <= 64 ? 2 : 4); }

int flashe_ctx_get_info(const flashe_ctx* ctx, flashe_ctx_info* out) {
    if (!ctx) return fail(FLASHE_EINVAL, "ctx is NULL");
    out->device = ctx->device; out->int_bits = ctx->int_bits; out->words = ctx->words; out->num_sms = ctx->num_sms;
    return FLASHE_OK;
}

int flashe_resident_ctas(const void* kernel, int threads, size_t dyn_smem, int device, int num_sms) {
    static std::mutex mu;
    static std::map<std::tuple<const void*, int, size_t, int>, int> cache;
    std::lock_guard<std::mutex> lock(mu);
    const auto key = std::make_tuple(kernel, threads, dyn_smem, device);
    auto it = cache.find(key);
    if (it == cache.end()) {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, dyn_smem) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 1; }
        it = cache.emplace(key, occ * num_sms).first;
    }
    return it->second;
}

int flashe_check_span(const flashe_span* s) {
    if (!s) return fail(FLASHE_EINVAL, "span is NULL");
    if (s->n_jobs == 0) return fail(FLASHE_EINVAL, "span.n_jobs must be >= 1");
    if (s->reserved != 0) return fail(FLASHE_EINVAL, "span.reserved must be 0");
    if (s->begin > s->total_len || s->count > s->total_len - s->begin)
        return fail(FLASHE_EINVAL, "span [begin, begin+count) exceeds total_len");
    return FLASHE_OK;
}

static void make_geom(const flashe_ctx* ctx, const flashe_span* s, uint32_t sup, Geom* g) {
    memset(g, 0, sizeof(*g));
    g->L = s->total_len; g->begin = s->begin; g->end = s->begin + s->count;
    g->m = ctx->m; g->b = (uint32_t)ctx->int_bits;
    g->d = s->total_len / s->n_jobs; g->r = s->total_len % s->n_jobs;
    const uint64_t nbA = ceil_div(g->d + 1, g->m), nbB = g->d ? ceil_div(g->d, g->m) : 0;
    // items per chunk: ceil((blocks + (cb & 63)) / 64) depends on the chunk; the bound for cb & 63 = 63 is
    // used for every chunk (at most one empty trailing item, skipped by the kernel)
    g->nwA = ceil_div(nbA + ITEM_BLOCKS - 1, ITEM_BLOCKS); g->nwB = nbB ? ceil_div(nbB + ITEM_BLOCKS - 1, ITEM_BLOCKS) : 0;
    if (sup < 1) sup = 1;
    g->sup = sup;
    g->nsA = ceil_div(g->nwA, sup); g->nsB = ceil_div(g->nwB, sup);
    g->rSA = g->r * g->nsA;
    g->aligned4 = ((s->begin & 3u) == 0 && (g->r <= 1 || ((g->d + 1) & 3u) == 0) && ((g->r * (g->d + 1)) & 3u) == 0 &&
                   ((g->d & 3u) == 0 || s->n_jobs - g->r <= 1)) ? 1u : 0u;
    if (s->count == 0) { g->S_lo = 0; g->S_cnt = 0; return; }
    auto unit_of = [&](uint64_t j) {
        uint64_t k, cb;
        if (j < g->r * (g->d + 1)) { k = j / (g->d + 1); cb = k * (g->d + 1); }
        else { k = g->r + (j - g->r * (g->d + 1)) / g->d; cb = g->r * (g->d + 1) + (k - g->r) * g->d; }
        const uint64_t u = (((j - cb) / g->m + (cb & (ITEM_BLOCKS - 1))) / ITEM_BLOCKS) / sup;
        return (k < g->r ? k * g->nsA : g->rSA + (k - g->r) * g->nsB) + u;
    };
    g->S_lo = unit_of(g->begin);
    g->S_cnt = unit_of(g->end - 1) - g->S_lo + 1;
}

// Items per work unit: as large as possible (amortises the per-unit chunk decode: two 64-bit divisions) while every
// warp of the persistent grid still gets at least FLASHE_UNITS_PER_WARP units (the last wave is dealt in pieces of
// sup / 8 items, so the tail does not grow with the unit size).
#ifndef FLASHE_UNITS_PER_WARP
#define FLASHE_UNITS_PER_WARP 8       // measured (profiles/r3n_ab_units_per_warp.txt): 32 -> 8 is 4-9 % on the 2.5M-element configs, 4 loses to its tail
#endif
static uint32_t pick_sup(const flashe_ctx* ctx, const flashe_span* s, uint64_t rows) {
    const uint64_t items = ceil_div(ceil_div(s->count ? s->count : 1, ctx->m), ITEM_BLOCKS) * (rows ? rows : 1);
    const uint64_t warps = (uint64_t)ctx->num_sms * (STREAM_THREADS / 32);
    uint64_t sup = items / (warps * FLASHE_UNITS_PER_WARP);
    return (uint32_t)(sup < 1 ? 1 : (sup > 32 ? 32 : sup));
}

// ---- ticket slots of the dynamic deal (flashe_internal.h) ---------------------------------------------------------
#define TICKET_STREAM_SLOTS 64
#define TICKET_CAPTURE_SLOTS 4032
struct flashe_ticket_state {
    std::mutex mu;
    cudaStream_t streams[TICKET_STREAM_SLOTS];
    int n_streams = 0, next_capture = 0;
};
uint32_t* flashe_ticket_slot(const flashe_ctx* ctx, cudaStream_t stream) {
    static const bool on = [] { const char* e = getenv("FLASHE_DYNAMIC"); return !(e && e[0] == '0'); }();
    if (!on || !ctx->d_tickets || !ctx->tickets || stream == cudaStreamPerThread) return nullptr;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    flashe_ticket_state* ts = ctx->tickets;
    std::lock_guard<std::mutex> lk(ts->mu);
    if (cap != cudaStreamCaptureStatusNone) {
        if (ts->next_capture >= TICKET_CAPTURE_SLOTS) return nullptr;
        return ctx->d_tickets + 2 * (TICKET_STREAM_SLOTS + ts->next_capture++);
    }
    for (int i = 0; i < ts->n_streams; ++i)
        if (ts->streams[i] == stream) return ctx->d_tickets + 2 * i;
    if (ts->n_streams >= TICKET_STREAM_SLOTS) return nullptr;
    ts->streams[ts->n_streams] = stream;
    return ctx->d_tickets + 2 * ts->n_streams++;
}

static int make_streams(const flashe_ctx* ctx, uint32_t iter, const int32_t* prf, const int32_t* sign, int n, StreamTab* st) {
    if (n < 1 || n > MAXS) return fail(FLASHE_EINVAL, "nstreams must be in [1, FLASHE_MAX_STREAMS]");
    if (!prf) return fail(FLASHE_EINVAL, "prf_idx is NULL");
    memset(st, 0, sizeof(*st));
    st->n = (uint32_t)n; st->iter = iter;
    for (int k = 0; k < n; ++k) {
        st->prf[k] = (uint32_t)prf[k];
        st->sign[k] = sign ? (sign[k] >= 0 ? 1 : -1) : 1;
        haes::hoist_round1(ctx->ks.rk, iter, st->prf[k], 0u, st->pre[k]);
    }
    return FLASHE_OK;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int flashe_abi_version(void) { return FLASHE_ABI_VERSION; }
const char* flashe_last_error(void) { return g_err.c_str(); }
uint64_t flashe_launch_count(void) { return g_launches.load(); }

int flashe_word_bytes(int int_bits) {
    if (int_bits < 1 || int_bits > 128) return fail(FLASHE_EINVAL, "int_bits must be in [1, 128]");
    return 4 * words_of(int_bits);
}

int flashe_ctx_create(const uint8_t* seed, size_t seed_len, int int_bits, int device, flashe_ctx** out) {
    if (!out) return fail(FLASHE_EINVAL, "out is NULL");
    *out = nullptr;
    if (!seed || seed_len == 0) return fail(FLASHE_EINVAL, "seed must hold at least one byte");
    if (int_bits < 8 || int_bits > 128)
        return fail(FLASHE_EUNSUPPORTED, "int_bits must be in [8, 128] (the reference itself fails above 128: merge_size = 128 // int_bits = 0)");
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(FLASHE_EINVAL, "no such CUDA device");
    FlasheDeviceGuard guard(device);
    if (!guard.ok) return fail(FLASHE_ECUDA, "cudaSetDevice failed");
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(FLASHE_EUNSUPPORTED, std::string("built for sm_100a (B200) only; device is ") + prop.name);
    if ((size_t)prop.sharedMemPerBlockOptin < SMEM_BYTES) return fail(FLASHE_EUNSUPPORTED, "device lacks 192 KB of shared memory per block");
    flashe_ctx* ctx = new (std::nothrow) flashe_ctx();
    if (!ctx) return fail(FLASHE_ENOMEM, "out of host memory");
    ctx->device = device; ctx->int_bits = int_bits; ctx->words = words_of(int_bits);
    ctx->num_sms = prop.multiProcessorCount; ctx->m = (uint32_t)(128 / int_bits);
    // low 32 bytes of the seed, left-zero-padded (jzf_aes.py:21-28)
    memset(ctx->key, 0, 32);
    if (seed_len >= 32) memcpy(ctx->key, seed + (seed_len - 32), 32);
    else memcpy(ctx->key + (32 - seed_len), seed, seed_len);
    haes::expand(ctx->key, ctx->ks.rk);
    // Workspaces come from the device's default stream-ordered pool; keep freed blocks cached across
    // synchronisation points (the default threshold of 0 returns them to the driver at every sync, and
    // the next cudaMallocAsync then costs milliseconds).
    {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    cudaError_t e = cudaMalloc((void**)&ctx->d_te0, sizeof(haes::te0));
    if (e == cudaSuccess) e = cudaMemcpy(ctx->d_te0, haes::te0, sizeof(haes::te0), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(ctx->d_te0); delete ctx; return fail(FLASHE_ECUDA, std::string("Te0 table upload: ") + cudaGetErrorString(e)); }
    const size_t ticket_bytes = 2 * sizeof(uint32_t) * (TICKET_STREAM_SLOTS + TICKET_CAPTURE_SLOTS);
    e = cudaMalloc((void**)&ctx->d_tickets, ticket_bytes);
    if (e == cudaSuccess) e = cudaMemset(ctx->d_tickets, 0, ticket_bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    ctx->tickets = new (std::nothrow) flashe_ticket_state();
    if (e != cudaSuccess || !ctx->tickets) {
        cudaFree(ctx->d_tickets); cudaFree(ctx->d_te0); delete ctx->tickets; delete ctx;
        return fail(FLASHE_ECUDA, std::string("ticket counters: ") + cudaGetErrorString(e));
    }
    *out = ctx;
    return FLASHE_OK;
}

int flashe_ctx_destroy(flashe_ctx* ctx) {
    if (!ctx) return FLASHE_OK;
    memset(ctx->key, 0, sizeof(ctx->key));
    memset(&ctx->ks, 0, sizeof(ctx->ks));
    {
        FlasheDeviceGuard guard(ctx->device);
        if (ctx->d_te0) cudaFree(ctx->d_te0);
        if (ctx->d_tickets) cudaFree(ctx->d_tickets);
    }
    delete ctx->tickets;
    delete ctx;
    return FLASHE_OK;
}
// ---- peer buffers (cudaIpc): see the header ----------------------------------------------------------
static_assert(sizeof(cudaIpcMemHandle_t) == FLASHE_PEER_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
int flashe_peer_alloc(flashe_ctx* ctx, uint64_t bytes, void** ptr_out, uint8_t* handle_out) {
    if (!ctx) return fail(FLASHE_EINVAL, "ctx is NULL");
    if (!ptr_out || !handle_out || bytes == 0) return fail(FLASHE_EINVAL, "need bytes > 0, ptr_out and handle_out");
    *ptr_out = nullptr;
    FlasheDeviceGuard guard(ctx->device);
    if (!guard.ok) return fail(FLASHE_ECUDA, "cudaSetDevice failed");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(e == cudaErrorMemoryAllocation ? FLASHE_ENOMEM : FLASHE_ECUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); cudaGetLastError(); return fail(FLASHE_ECUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
    memcpy(handle_out, &h, sizeof(h));
    *ptr_out = p;
    return FLASHE_OK;
}
int flashe_peer_open(flashe_ctx* ctx, const uint8_t* handle, void** ptr_out) {
    if (!ctx) return fail(FLASHE_EINVAL, "ctx is NULL");
    if (!handle || !ptr_out) return fail(FLASHE_EINVAL, "need handle and ptr_out");
    *ptr_out = nullptr;
    FlasheDeviceGuard guard(ctx->device);
    if (!guard.ok) return fail(FLASHE_ECUDA, "cudaSetDevice failed");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(FLASHE_ECUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e)); }
    *ptr_out = p;
    return FLASHE_OK;
}
int flashe_peer_close(flashe_ctx* ctx, void* ptr) {
    if (!ctx) return fail(FLASHE_EINVAL, "ctx is NULL");
    if (!ptr) return FLASHE_OK;
    FlasheDeviceGuard guard(ctx->device);
    if (!guard.ok) return fail(FLASHE_ECUDA, "cudaSetDevice failed");
    CUDA_TRY(cudaIpcCloseMemHandle(ptr));
    return FLASHE_OK;
}
int flashe_peer_free(flashe_ctx* ctx, void* ptr) {
    if (!ctx) return fail(FLASHE_EINVAL, "ctx is NULL");
    if (!ptr) return FLASHE_OK;
    FlasheDeviceGuard guard(ctx->device);
    if (!guard.ok) return fail(FLASHE_ECUDA, "cudaSetDevice failed");
    CUDA_TRY(cudaFree(ptr));
    return FLASHE_OK;
}

int flashe_ctx_int_bits(const flashe_ctx* ctx) { return ctx ? ctx->int_bits : fail(FLASHE_EINVAL, "ctx is NULL"); }
int flashe_ctx_device(const flashe_ctx* ctx) { return ctx ? ctx->device : fail(FLASHE_EINVAL, "ctx is NULL"); }

int flashe_prp_block(flashe_ctx* ctx, const uint8_t in16[16], uint8_t out16[16], void* stream) {
    ENTER(ctx);
    if (!in16 || !out16) return fail(FLASHE_EINVAL, "NULL buffer");
    uint32_t w[4], o[8];
    for (int i = 0; i < 4; ++i) w[i] = ((uint32_t)in16[4 * i] << 24) | ((uint32_t)in16[4 * i + 1] << 16) | ((uint32_t)in16[4 * i + 2] << 8) | in16[4 * i + 3];
    uint32_t* d = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d, 12 * sizeof(uint32_t)));
    cudaError_t e = cudaMemcpyAsync(d, w, 16, cudaMemcpyHostToDevice, cs);
    if (e == cudaSuccess && flashe_launch_prp_block(ctx, d, d + 4, cs) != FLASHE_OK) { cudaFree(d); return FLASHE_ECUDA; }
    if (e == cudaSuccess) e = cudaMemcpyAsync(o, d + 4, 32, cudaMemcpyDeviceToHost, cs);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cs);
    cudaFree(d);
    if (e != cudaSuccess) return fail(FLASHE_ECUDA, std::string("flashe_prp_block: ") + cudaGetErrorString(e));
    // o = hoisted round-1 path, o+4 = generic path: both must give the same block
    const uint32_t* r = o;
    if (memcmp(o, o + 4, 16) != 0) return fail(FLASHE_ECUDA, "internal: hoisted and generic AES paths disagree");
    for (int i = 0; i < 4; ++i) { out16[4 * i] = (uint8_t)(r[i] >> 24); out16[4 * i + 1] = (uint8_t)(r[i] >> 16); out16[4 * i + 2] = (uint8_t)(r[i] >> 8); out16[4 * i + 3] = (uint8_t)r[i]; }
    return FLASHE_OK;
}

int flashe_masks(flashe_ctx* ctx, uint32_t iter, const int32_t* prf_idx, const int32_t* sign, int nstreams,
                 const flashe_span* span, void* out, void* stream) {
    ENTER(ctx);
    int rc = flashe_check_span(span); if (rc) return rc;
    if (span->count && !out) return fail(FLASHE_EINVAL, "out is NULL");
    StreamTab st; rc = make_streams(ctx, iter, prf_idx, sign, nstreams, &st); if (rc) return rc;
    Geom g; make_geom(ctx, span, pick_sup(ctx, span, 1), &g);
    IoDev io; memset(&io, 0, sizeof(io)); io.out = out; io.n_clients = 1;
    CodecDev cd; memset(&cd, 0, sizeof(cd)); NoiseDev nz; memset(&nz, 0, sizeof(nz));
    return flashe_launch_stream_masks(ctx, st, g, io, cd, nz, cs);
}

int flashe_precompute(flashe_ctx* ctx, uint32_t iter_from, int n_rounds, const int32_t* prf_idx, const int32_t* sign, int nstreams,
                      const flashe_span* span, void* out, uint64_t round_stride, void* stream) {
    if (!ctx) return fail(FLASHE_EINVAL, "ctx is NULL");
    if (n_rounds < 0) return fail(FLASHE_EINVAL, "n_rounds must be >= 0");
    int rc = flashe_check_span(span); if (rc) return rc;
    if (n_rounds > 1 && round_stride < span->count) return fail(FLASHE_EINVAL, "round_stride must be >= span.count");
    for (int r = 0; r < n_rounds; ++r) {
        rc = flashe_masks(ctx, iter_from + (uint32_t)r, prf_idx, sign, nstreams, span,
                          (uint8_t*)out + (size_t)r * round_stride * 4u * (size_t)ctx->words, stream);
        if (rc) return rc;
    }
    return FLASHE_OK;
}

int flashe_apply_masks(flashe_ctx* ctx, uint32_t iter, const int32_t* prf_idx, const int32_t* sign, int nstreams,
                       const flashe_span* span, const void* in, void* out, void* stream) {
    ENTER(ctx);
    int rc = flashe_check_span(span); if (rc) return rc;
    if (span->count && (!in || !out)) return fail(FLASHE_EINVAL, "in/out is NULL");
    StreamTab st; rc = make_streams(ctx, iter, prf_idx, sign, nstreams, &st); if (rc) return rc;
    Geom g; make_geom(ctx, span, pick_sup(ctx, span, 1), &g);
    IoDev io; memset(&io, 0, sizeof(io)); io.in = in; io.out = out; io.n_clients = 1;
    CodecDev cd; memset(&cd, 0, sizeof(cd)); NoiseDev nz; memset(&nz, 0, sizeof(nz));
    return flashe_launch_stream_apply(ctx, st, g, io, cd, nz, cs);
}

int flashe_encrypt(flashe_ctx* ctx, uint32_t iter, int32_t idx, int scheme, const flashe_span* span, const void* q_in,
                   void* ct_out, void* stream) {
    if (scheme != FLASHE_SCHEME_SINGLE && scheme != FLASHE_SCHEME_DOUBLE) return fail(FLASHE_EINVAL, "unknown masking scheme");
    if (idx < 0) return fail(FLASHE_EINVAL, "client idx must be >= 0");
    const int32_t prf[2] = {idx, idx + 1}, sg[2] = {1, -1};
    return flashe_apply_masks(ctx, iter, prf, sg, scheme == FLASHE_SCHEME_DOUBLE ? 2 : 1, span, q_in, ct_out, stream);
}

static int build_decrypt_streams(const int32_t* add_idx, int na, const int32_t* minus_idx, int ns, int32_t* prf, int32_t* sg) {
    if (na < 0 || ns < 0 || na + ns < 1 || na + ns > MAXS) return fail(FLASHE_EINVAL, "na + ns must be in [1, FLASHE_MAX_STREAMS]");
    if ((na && !add_idx) || (ns && !minus_idx)) return fail(FLASHE_EINVAL, "index list is NULL");
    for (int k = 0; k < na; ++k) { prf[k] = add_idx[k]; sg[k] = 1; }
    for (int k = 0; k < ns; ++k) { prf[na + k] = minus_idx[k]; sg[na + k] = -1; }
    return FLASHE_OK;
}

int flashe_decrypt(flashe_ctx* ctx, uint32_t iter, const int32_t* add_idx, int na, const int32_t* minus_idx, int ns,
                   const flashe_span* span, const void* agg_in, void* p_out, void* stream) {
    int32_t prf[MAXS], sg[MAXS];
    int rc = build_decrypt_streams(add_idx, na, minus_idx, ns, prf, sg); if (rc) return rc;
    return flashe_apply_masks(ctx, iter, prf, sg, na + ns, span, agg_in, p_out, stream);
}



static int encode_encrypt_impl(flashe_ctx* ctx, uint32_t iter, int32_t idx0, int n_clients, int scheme, const flashe_span* span,
                               const float* x, uint64_t x_stride, const flashe_codec* codec, const flashe_noise* noise,
                               uint64_t u_stride, void* ct_out, uint64_t ct_stride, uint32_t* q_out, int share, void* stream) {
    ENTER(ctx);
    int rc = flashe_check_span(span); if (rc) return rc;
    if (scheme != FLASHE_SCHEME_SINGLE && scheme != FLASHE_SCHEME_DOUBLE) return fail(FLASHE_EINVAL, "unknown masking scheme");
    if (idx0 < 0) return fail(FLASHE_EINVAL, "client idx must be >= 0");
    if (n_clients < 1 || n_clients > MAXS - 1) return fail(FLASHE_EINVAL, "n_clients must be in [1, FLASHE_MAX_STREAMS-1] per call");
    if (span->count == 0) return FLASHE_OK;
    if (!x || !ct_out) return fail(FLASHE_EINVAL, "NULL buffer");
    if (n_clients > 1 && (x_stride < span->count || ct_stride < span->count)) return fail(FLASHE_EINVAL, "client strides must be >= span.count");
    const bool dbl = scheme == FLASHE_SCHEME_DOUBLE;
    int32_t prf[MAXS];
    const int nent = n_clients + (dbl ? 1 : 0);
    for (int k = 0; k < nent; ++k) prf[k] = idx0 + k;
    StreamTab st; rc = make_streams(ctx, iter, prf, nullptr, nent, &st); if (rc) return rc;
    st.batch = 1; st.dbl = dbl ? 1u : 0u;
    Geom g; make_geom(ctx, span, pick_sup(ctx, span, (share && dbl) ? 1 : (uint64_t)n_clients), &g);
    CodecHost ch; rc = make_codec(ctx, span, codec, false, cs, &ch); if (rc) return rc;
    rc = check_noise(noise); if (rc) { free_codec(&ch, cs); return rc; }
    NoiseDev nz; make_noise(noise, u_stride, &nz);
    IoDev io; memset(&io, 0, sizeof(io));
    io.in = x; io.in_stride = x_stride; io.out = ct_out; io.out_stride = ct_stride; io.aux = q_out;
    io.n_clients = (uint32_t)n_clients; io.share = (share && dbl) ? 1u : 0u; io.elem0 = ch.elem0;
    const bool n32 = nz.res32 != 0u && nz.u == nullptr;
    if (n32) rc = io.share ? flashe_launch_stream_encode_shared_n32(ctx, st, g, io, ch.dev, nz, cs)
                           : flashe_launch_stream_encode_n32(ctx, st, g, io, ch.dev, nz, cs);
    else rc = io.share ? flashe_launch_stream_encode_shared(ctx, st, g, io, ch.dev, nz, cs)
                       : flashe_launch_stream_encode(ctx, st, g, io, ch.dev, nz, cs);
    free_codec(&ch, cs);
    return rc;
}

int flashe_encode_encrypt(flashe_ctx* ctx, uint32_t iter, int32_t idx, int scheme, const flashe_span* span, const float* x,
                          const flashe_codec* codec, const flashe_noise* noise, void* ct_out, uint32_t* q_out, void* stream) {
    return encode_encrypt_impl(ctx, iter, idx, 1, scheme, span, x, 0, codec, noise, 0, ct_out, 0, q_out, 0, stream);
}

int flashe_encode_encrypt_batch(flashe_ctx* ctx, uint32_t iter, int32_t idx0, int n_clients, int scheme, const flashe_span* span,
                                const float* x, uint64_t x_stride, const flashe_codec* codec, const flashe_noise* noise,
                                uint64_t u_stride, void* ct_out, uint64_t ct_stride, int share_streams, void* stream) {
    return encode_encrypt_impl(ctx, iter, idx0, n_clients, scheme, span, x, x_stride, codec, noise, u_stride, ct_out, ct_stride,
                               nullptr, share_streams, stream);
}





int flashe_decrypt_decode(flashe_ctx* ctx, uint32_t iter, const int32_t* add_idx, int na, const int32_t* minus_idx, int ns,
                          const flashe_span* span, const void* agg_in, const flashe_codec* codec, double* out, void* p_out,
                          void* stream) {
    ENTER(ctx);
    int rc = flashe_check_span(span); if (rc) return rc;
    if (ctx->words == 4 && !(codec && codec->batch_lane_bits))
        return fail(FLASHE_EUNSUPPORTED, "decrypt_decode of 16-byte words needs a lane-batching codec (batch_lane_bits): one element per 128-bit word has no decode");
    int32_t prf[MAXS], sg[MAXS];
    rc = build_decrypt_streams(add_idx, na, minus_idx, ns, prf, sg); if (rc) return rc;
    if (span->count == 0) return FLASHE_OK;
    if (!agg_in || !out) return fail(FLASHE_EINVAL, "NULL buffer");
    StreamTab st; rc = make_streams(ctx, iter, prf, sg, na + ns, &st); if (rc) return rc;
    Geom g; make_geom(ctx, span, pick_sup(ctx, span, 1), &g);
    CodecHost ch; rc = make_codec(ctx, span, codec, true, cs, &ch); if (rc) return rc;
    IoDev io; memset(&io, 0, sizeof(io)); io.in = agg_in; io.outf = out; io.aux = p_out; io.n_clients = 1; io.elem0 = ch.elem0;
    NoiseDev nz; memset(&nz, 0, sizeof(nz));
    rc = flashe_launch_stream_decode(ctx, st, g, io, ch.dev, nz, cs);
    free_codec(&ch, cs);
    return rc;
}







int flashe_sparse_apply_masks(flashe_ctx* ctx, uint32_t iter, const int32_t* prf_idx, const int32_t* sign, int nstreams,
                              const flashe_span* span, const int64_t* index, void* dense, uint64_t dense_len, void* stream) {
    ENTER(ctx);
    int rc = flashe_check_span(span); if (rc) return rc;
    if (span->count == 0) return FLASHE_OK;
    if (!index || !dense) return fail(FLASHE_EINVAL, "NULL buffer");
    if (span->count > dense_len) return fail(FLASHE_EINVAL, "more compact positions than dense words");
    StreamTab st; rc = make_streams(ctx, iter, prf_idx, sign, nstreams, &st); if (rc) return rc;
    Geom g; make_geom(ctx, span, pick_sup(ctx, span, 1), &g);
    IoDev io; memset(&io, 0, sizeof(io)); io.out = dense; io.aux = (void*)index; io.n_clients = 1; io.dense_len = dense_len;
    CodecDev cd; memset(&cd, 0, sizeof(cd)); NoiseDev nz; memset(&nz, 0, sizeof(nz));
    return flashe_launch_stream_scatter(ctx, st, g, io, cd, nz, cs);
}

int flashe_sparse_apply_masks_batch(flashe_ctx* ctx, uint32_t iter, const int32_t* prf_idx, int sign, int n_clients, const uint64_t* ks,
                                    uint32_t n_jobs, const int64_t* const* indexes, void* dense, uint64_t dense_len, void* stream) {
    ENTER(ctx);
    if (n_clients < 1 || !prf_idx || !ks || !indexes) return fail(FLASHE_EINVAL, "need n_clients >= 1, prf_idx, ks and indexes");
    if (n_jobs == 0) return fail(FLASHE_EINVAL, "n_jobs must be >= 1");
    if (!dense) return fail(FLASHE_EINVAL, "dense is NULL");
    uint64_t words_total = 0;
    std::vector<uint64_t> off((size_t)n_clients);
    for (int c = 0; c < n_clients; ++c) {
        if (ks[c] > dense_len) return fail(FLASHE_EINVAL, "more compact positions than dense words");
        if (ks[c] >= (1ull << 32)) return fail(FLASHE_EUNSUPPORTED, "index lists of 2^32 or more entries");
        if (ks[c] && !indexes[c]) return fail(FLASHE_EINVAL, "NULL index list");
        off[c] = words_total;
        words_total += (ks[c] + 3) & ~3ull;                              // rows stay 16-byte aligned
    }
    if (words_total == 0) return FLASHE_OK;
    const size_t wb = 4u * (size_t)ctx->words;
    uint8_t* masks = nullptr;
    CUDA_TRY(cudaMallocAsync((void**)&masks, words_total * wb, cs));
    // 1. F(iter, prf_c) over the compact positions of every client: one launch per run of clients whose lists have the
    //    same length (the reference's sparsity is per layer of a common model: every client's list has the same length)
    int rc = FLASHE_OK;
    for (int c0 = 0; rc == FLASHE_OK && c0 < n_clients;) {
        int c1 = c0 + 1;
        while (c1 < n_clients && ks[c1] == ks[c0] && c1 - c0 < MAXS) ++c1;
        if (ks[c0]) {
            flashe_span span; memset(&span, 0, sizeof(span));
            span.total_len = ks[c0]; span.begin = 0; span.count = ks[c0]; span.n_jobs = n_jobs;
            StreamTab st; rc = make_streams(ctx, iter, prf_idx + c0, nullptr, c1 - c0, &st);
            if (rc == FLASHE_OK) {
                st.batch = 1; st.dbl = 0;
                Geom g; make_geom(ctx, &span, pick_sup(ctx, &span, (uint64_t)(c1 - c0)), &g);
                IoDev io; memset(&io, 0, sizeof(io));
                io.out = masks + off[c0] * wb; io.out_stride = (ks[c0] + 3) & ~3ull; io.n_clients = (uint32_t)(c1 - c0);
                CodecDev cd; memset(&cd, 0, sizeof(cd)); NoiseDev nz; memset(&nz, 0, sizeof(nz));
                rc = flashe_launch_stream_masks(ctx, st, g, io, cd, nz, cs);
            }
        }
        c0 = c1;
    }
    // 2. dense[index_c[p]] += sign * mask_c[p], every client at once, tile by tile
    if (rc == FLASHE_OK) {
        std::vector<const void*> rows((size_t)n_clients);
        for (int c = 0; c < n_clients; ++c) rows[c] = masks + off[c] * wb;
        rc = flashe_sparse_accumulate_tiled(ctx, rows.data(), indexes, ks, n_clients, sign < 0 ? 1 : 0, dense_len, dense, cs);
    }
    cudaFreeAsync(masks, cs);
    return rc;
}

}  // extern "C"
