// flashe_wire.cu — the callers and data formats either side of the FLASHE hot path:
//
//   f1  wire bit-packing          framework/jzf_weights.py:45-137 (_to_bytes / _from_bytes), 155-231
//   f2  top-k sparsify + residual framework/homo/procedure/jzf_aggregator.py:578-623 (Client.sparsify)
//   (f3, the per-layer statistics of secureprotol/jzf_quantize.py:542-564, lives in flashe_stats.cu)
//
// All three are HBM-bound byte / integer / reduction work: grid-stride kernels with grids capped at a
// multiple of the SM count, wide accesses where the layout allows.  No tensor cores.
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "flashe_internal.h"

typedef unsigned __int128 u128_t;

static uint64_t ceil_div_u64(uint64_t a, uint64_t b) { return (a + b - 1) / b; }
static int grid_cap(int num_sms, uint64_t blocks, int per_sm) {
    uint64_t cap = (uint64_t)num_sms * per_sm;
    if (blocks < 1) blocks = 1;
    return (int)(blocks < cap ? blocks : cap);
}

#define WIRE_ENTER(ctx)                                                                   \
    flashe_ctx_info info;                                                                 \
    { int rc__ = flashe_ctx_get_info((ctx), &info); if (rc__) return rc__; }              \
    FlasheDeviceGuard guard__(info.device);                                               \
    if (!guard__.ok) return flashe_fail(FLASHE_ECUDA, "cudaSetDevice failed");            \
    cudaStream_t cs = (cudaStream_t)stream

// =================================================================================================
// f1  wire bit-packing
//
// _to_bytes(flatten_array, num_bits) (jzf_weights.py:45-84) builds the Python integer
//     s = sum_j  a[j] << ((L-1-j) * num_bits)            (first element most significant)
// batch by batch; the batching (lcm(num_bits, 8) bits at a time) only bounds the size of the
// intermediate integers and does not change s.  On the wire s travels as a big integer; here it is the
// big-endian byte string of length ceil(L*num_bits/8) (int.to_bytes(nbytes, 'big')), i.e. output byte k
// holds bits [8*(nbytes-1-k), 8*(nbytes-k)) of s, with zero bits above L*num_bits.
// _from_bytes (jzf_weights.py:98-137) + the reverse() in decompress (:224) is the inverse.
// =================================================================================================
template <int WB> struct WireWord;
template <> struct WireWord<4> {
    static __device__ __forceinline__ u128_t load(const void* p, uint64_t j) { return reinterpret_cast<const uint32_t*>(p)[j]; }
    static __device__ __forceinline__ void store(void* p, uint64_t j, u128_t v) { reinterpret_cast<uint32_t*>(p)[j] = (uint32_t)v; }
};
template <> struct WireWord<8> {
    static __device__ __forceinline__ u128_t load(const void* p, uint64_t j) { return reinterpret_cast<const uint64_t*>(p)[j]; }
    static __device__ __forceinline__ void store(void* p, uint64_t j, u128_t v) { reinterpret_cast<uint64_t*>(p)[j] = (uint64_t)v; }
};
template <> struct WireWord<16> {
    static __device__ __forceinline__ u128_t load(const void* p, uint64_t j) {
        const uint4 r = reinterpret_cast<const uint4*>(p)[j];
        return ((u128_t)(((uint64_t)r.w << 32) | r.z) << 64) | (((uint64_t)r.y << 32) | r.x);
    }
    static __device__ __forceinline__ void store(void* p, uint64_t j, u128_t v) {
        const uint64_t lo = (uint64_t)v, hi = (uint64_t)(v >> 64);
        reinterpret_cast<uint4*>(p)[j] = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
    }
};

__device__ __forceinline__ u128_t field_mask(uint32_t bits) { return bits >= 128 ? ~(u128_t)0 : (((u128_t)1 << bits) - 1); }

// One thread = one 16-byte chunk of the output stream (chunk k = bytes [16k, 16k+16)).  The chunk is the
// 128-bit window [lo_bit, lo_bit+128) of s with lo_bit = 8*(nbytes - 16k) - 128 (negative only for a
// short final chunk); the window is assembled from the <= 128/bits + 2 fields that intersect it.
// Consecutive threads read consecutive elements and write consecutive 16-byte chunks.
template <int WB>
__global__ void __launch_bounds__(256)
k_wire_pack(const void* __restrict__ words, uint64_t count, uint32_t bits, uint64_t nbytes, uint8_t* __restrict__ out, uint64_t k0) {
    const uint64_t nchunks = (nbytes + 15) >> 4;
    const u128_t fm = field_mask(bits);
    for (uint64_t k = k0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nchunks; k += (uint64_t)gridDim.x * blockDim.x) {
        const int64_t hi_bit = (int64_t)(8 * (nbytes - 16 * k));     // exclusive, > 0
        const int64_t lo_bit = hi_bit - 128;
        // fields are numbered from the END: field e holds element count-1-e at bits [e*bits, (e+1)*bits)
        const uint64_t e_lo = lo_bit > 0 ? (uint64_t)lo_bit / bits : 0;
        uint64_t e_hi = (uint64_t)(hi_bit - 1) / bits;
        if (e_hi >= count) e_hi = count - 1;                          // zero padding above the top field
        u128_t v = 0;
        for (uint64_t e = e_hi + 1; e-- > e_lo;) {                    // ascending element index
            const u128_t a = WireWord<WB>::load(words, count - 1 - e) & fm;
            const int64_t sh = (int64_t)(e * bits) - lo_bit;
            if (sh >= 128 || sh <= -128) continue;
            v |= sh >= 0 ? (a << sh) : (a >> (-sh));
        }
        // big-endian bytes of the window
        const uint64_t vh = (uint64_t)(v >> 64), vl = (uint64_t)v;
        const uint32_t w0 = __byte_perm((uint32_t)(vh >> 32), 0, 0x0123), w1 = __byte_perm((uint32_t)vh, 0, 0x0123);
        const uint32_t w2 = __byte_perm((uint32_t)(vl >> 32), 0, 0x0123), w3 = __byte_perm((uint32_t)vl, 0, 0x0123);
        if (16 * k + 16 <= nbytes) {
            reinterpret_cast<uint4*>(out)[k] = make_uint4(w0, w1, w2, w3);
        } else {
            const uint32_t w[4] = {w0, w1, w2, w3};
            for (uint64_t i = 0; 16 * k + i < nbytes; ++i) out[16 * k + i] = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
        }
    }
}

// One thread = one element: field e = count-1-j lives at bits [p, p+bits), p = e*bits, of s; its bytes
// are gathered most significant first.  Neighbouring threads read neighbouring (overlapping) bytes.
template <int WB>
__global__ void __launch_bounds__(256)
k_wire_unpack(const uint8_t* __restrict__ in, uint64_t count, uint32_t bits, uint64_t nbytes, void* __restrict__ words) {
    const u128_t fm = field_mask(bits);
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t p = (count - 1 - j) * (uint64_t)bits;
        const uint64_t b_lo = p >> 3, b_hi = (p + bits - 1) >> 3;    // little-endian byte numbers of s
        const uint32_t sh = (uint32_t)(p & 7);
        u128_t v = 0;
        uint32_t top = 0;                                            // 17th byte when bits + sh > 128
        for (uint64_t bn = b_hi + 1; bn-- > b_lo;) {
            const uint32_t byte = __ldg(in + (nbytes - 1 - bn));
            const uint64_t rel = bn - b_lo;
            if (rel >= 16) top = byte; else v |= (u128_t)byte << (8 * rel);
        }
        v = sh ? ((v >> sh) | ((u128_t)top << (128 - sh))) : v;
        WireWord<WB>::store(words, j, v & fm);
    }
}

// -------------------------------------------------------------------------------------------------
// Fast forms for 4-byte words (8 <= bits <= 32 for pack: the ciphertext widths FLASHE ships).  Seen from
// its first byte the wire string is a big-endian BIT stream: `pad` = 8*nbytes - count*bits zero bits,
// then the elements in index order, most significant bit first; element j occupies stream bits
// [pad + j*bits, pad + (j+1)*bits), i.e. at most two 32-bit stream words.
//   pack    a CTA owns a tile of 1024 stream words; every thread reads its elements straight from global
//           memory (coalesced 4-byte loads, ~6 in flight per thread), ORs the one or two word pieces into
//           the tile image in shared memory (shared atomics), and the image leaves as 16-byte stores.
//   unpack  the stream words of a tile of 1024 elements are staged in shared memory with 16-byte
//           cp.async (next tile in flight while this one is converted); an element is one funnel shift
//           of two staged words.
// Both are a few tens of instructions per 16 bytes moved, so HBM, not the issue rate, bounds them.
// -------------------------------------------------------------------------------------------------
#define WP_THREADS 256
#define WP_WORDS (WP_THREADS * 4)                 // stream words (32 bit) per pack tile

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct PackTile { uint64_t f_lo; uint32_t nf; int32_t rel0; };

// pack: tile t = stream words [t*WP_WORDS, (t+1)*WP_WORDS); only whole 16-byte chunks (nvec of them)
__global__ void __launch_bounds__(WP_THREADS)
k_wire_pack32(const uint32_t* __restrict__ words, uint64_t count, uint32_t bits, uint32_t pad, uint64_t nvec,
              uint4* __restrict__ out) {
    __shared__ __align__(16) uint32_t so[2][WP_WORDS];
    __shared__ PackTile sg[2];
    const uint32_t fm = bits >= 32u ? 0xffffffffu : ((1u << bits) - 1u);
    const uint64_t ntiles = (nvec + WP_THREADS - 1) / WP_THREADS;
    uint64_t t = blockIdx.x;
    if (t >= ntiles) return;

    // thread 0 prepares the geometry of a tile one step ahead (the only 64-bit divisions of the kernel)
    auto geom = [&](uint64_t tile, int slot) {
        const uint64_t bit0 = tile * (uint64_t)(WP_WORDS * 32);             // stream bit of the tile's first word
        PackTile g;
        g.f_lo = bit0 > pad ? (bit0 - pad) / bits : 0;                       // first element that reaches into the tile
        uint64_t f_hi = (bit0 + (uint64_t)(WP_WORDS * 32) - 1 - pad) / bits; // last element that starts inside it
        if (f_hi >= count) f_hi = count - 1;
        g.nf = (uint32_t)(f_hi - g.f_lo + 1);
        g.rel0 = (int32_t)((int64_t)(pad + g.f_lo * bits) - (int64_t)bit0);  // in (-bits, 32)
        sg[slot] = g;
    };
    for (uint32_t i = threadIdx.x; i < 2 * WP_WORDS; i += WP_THREADS) (&so[0][0])[i] = 0u;
    if (threadIdx.x == 0) geom(t, 0);
    __syncthreads();
    for (int buf = 0; t < ntiles; t += gridDim.x, buf ^= 1) {
        const PackTile g = sg[buf];
        if (threadIdx.x == 0 && t + gridDim.x < ntiles) geom(t + gridDim.x, buf ^ 1);
        const uint32_t* src = words + g.f_lo;
        uint32_t* img = so[buf];
        for (uint32_t i0 = threadIdx.x; i0 < g.nf; i0 += 4u * WP_THREADS) {  // four independent loads in flight per thread
            uint32_t v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { const uint32_t i = i0 + k * WP_THREADS; v[k] = i < g.nf ? __ldcs(src + i) & fm : 0u; }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t i = i0 + k * WP_THREADS;
                const int32_t rel = g.rel0 + (int32_t)(i * bits);           // element start, in tile bits
                const int32_t wi = rel >> 5;                                // -1 for an element that begins before the tile
                const uint64_t V = (uint64_t)v[k] << (64u - bits - ((uint32_t)rel & 31u));   // placed in the word pair (wi, wi+1)
                const uint32_t hi = (uint32_t)(V >> 32), lo = (uint32_t)V;
                if (hi && (uint32_t)wi < (uint32_t)WP_WORDS) atomicOr(img + wi, hi);
                if (lo && (uint32_t)(wi + 1) < (uint32_t)WP_WORDS) atomicOr(img + wi + 1, lo);
            }
        }
        __syncthreads();                                                    // the tile image is complete
        const uint64_t vec = t * WP_THREADS + threadIdx.x;
        uint4* cell = reinterpret_cast<uint4*>(img) + threadIdx.x;
        const uint4 w = *cell;
        *cell = make_uint4(0u, 0u, 0u, 0u);                                 // ready for the tile after next
        if (vec < nvec)
            __stcs(out + vec, make_uint4(__byte_perm(w.x, 0, 0x0123), __byte_perm(w.y, 0, 0x0123),
                                         __byte_perm(w.z, 0, 0x0123), __byte_perm(w.w, 0, 0x0123)));   // first stream byte first
        // no second barrier: the next tile ORs into the other image (cleared one round ago, before the
        // barrier above) and reads the other geometry slot
    }
}

// -------------------------------------------------------------------------------------------------
// pack, second form (the one flashe_wire_pack launches when `words` is 16-byte aligned): no shared-memory
// atomics.  A CTA walks tiles of WB_WORDS stream words.  The elements that reach into a tile are ONE contiguous
// run of the input, so thread 0 fetches them with a single bulk asynchronous copy (cp.async.bulk global ->
// shared, completion counted in bytes on an mbarrier: the TMA path, no per-thread load instructions), two tiles
// in flight.  Every thread then builds runs of four consecutive stream words from the staged elements (the
// first one is found with one multiply-high, the rest are walked once through a 64-bit bit accumulator) and
// writes each as one 16-byte store.
// -------------------------------------------------------------------------------------------------
#define WB_THREADS 256
#define WB_RUN 4                                     // consecutive stream words per run (one 16-byte store)
#define WB_WORDS (WB_THREADS * WB_RUN * 4)            // stream words per tile (16 KB of output): four runs per thread

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// FB: the element width as a compile-time constant (0 = any width, taken from `bits`): the elements that feed a run of
// four stream words then land through constant shifts (see the run loop).
template <int FB>
__global__ void __launch_bounds__(WB_THREADS)
k_wire_pack32_bulk(const uint32_t* __restrict__ words, uint64_t count, uint32_t bits, uint32_t pad, uint64_t nwords,
                   uint32_t magic, uint32_t stage_words, uint32_t* __restrict__ out) {
    extern __shared__ __align__(128) uint32_t stage[];            // 2 x stage_words
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ int64_t s_rel[2];                                   // stream bit of staged element 0 minus the tile's first bit
    const uint32_t la = 32u - bits;                                // left-alignment shift of an element
    const uint64_t ntiles = (nwords + WB_WORDS - 1) / WB_WORDS;
    uint64_t t = blockIdx.x;
    if (t >= ntiles) return;
    const uint32_t bar0 = smem_u32(&bars[0]), st0 = smem_u32(stage);
    if (threadIdx.x == 0) {
        mbar_init(bar0, 1); mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // thread 0: stage the elements that intersect tile `tile` (run [f_al, f_al + n), f_al a multiple of 4 elements)
    auto issue = [&](uint64_t tile, int buf) {
        const uint64_t bit0 = tile * (uint64_t)(WB_WORDS * 32);
        const uint64_t f_lo = bit0 > pad ? (bit0 - pad) / bits : 0;
        uint64_t f_hi = (bit0 + (uint64_t)(WB_WORDS * 32) - 1 - pad) / bits;
        if (f_hi >= count) f_hi = count - 1;
        const uint64_t f_al = f_lo & ~3ull;
        uint64_t n = f_hi - f_al + 1;
        const uint64_t n_bulk_max = (count & ~3ull) > f_al ? (count & ~3ull) - f_al : 0;   // whole 16-byte pieces inside the vector
        const uint64_t n_bulk = ((n + 3) & ~3ull) <= n_bulk_max ? ((n + 3) & ~3ull) : n_bulk_max;
        s_rel[buf] = (int64_t)(pad + f_al * bits) - (int64_t)bit0;
        const uint32_t dst = st0 + (uint32_t)buf * stage_words * 4u;
        for (uint64_t i = n_bulk; i < n; ++i) stage[(uint32_t)buf * stage_words + i] = words[f_al + i];   // <= 3 elements at the vector's end
        if (f_hi + 1 >= count)                                                                            // last tile: slots past the last element read as zero
            for (uint64_t i = n; i < n + 8 && i < stage_words; ++i) stage[(uint32_t)buf * stage_words + i] = 0u;
        mbar_expect_tx(bar0 + 8u * buf, (uint32_t)(n_bulk * 4));
        if (n_bulk) bulk_g2s(dst, words + f_al, (uint32_t)(n_bulk * 4), bar0 + 8u * buf);
    };
    if (threadIdx.x == 0) issue(t, 0);
    uint32_t phase[2] = {0u, 0u};
    for (int buf = 0; t < ntiles; t += gridDim.x, buf ^= 1) {
        if (threadIdx.x == 0 && t + gridDim.x < ntiles) issue(t + gridDim.x, buf ^ 1);
        mbar_wait(bar0 + 8u * buf, phase[buf]);
        phase[buf] ^= 1u;
        const int32_t rel_tile = (int32_t)s_rel[buf];                  // <= 0 except for the first tile (pad), > -(4 * bits)
        const uint32_t* el = stage + (uint32_t)buf * stage_words;
        const uint64_t w_base = t * (uint64_t)WB_WORDS;
        // a thread owns runs of WB_RUN = 4 consecutive stream words (one 16-byte store, consecutive lanes -> consecutive
        // 16-byte pieces: a warp writes 512 contiguous bytes per store) and walks the elements that feed a run ONCE, left
        // to right, through a 64-bit bit accumulator (`fill` valid bits at the top): every element costs one
        // shared-memory load, two shifts and an OR, every finished word one shift
#pragma unroll
        for (uint32_t rr = 0; rr < WB_WORDS / WB_RUN / WB_THREADS; ++rr) {
            const uint32_t w0 = (threadIdx.x + rr * WB_THREADS) * WB_RUN;
            if (w_base + w0 >= nwords) break;
            const int32_t x = (int32_t)(32u * w0) - rel_tile;          // bit offset of the run from staged element 0
            uint32_t i = x > 0 ? __umulhi((uint32_t)x, magic) : 0u;     // element that holds the run's first bit (x < 2^27)
            const int32_t rel = (int32_t)(i * bits) - x;                // its start relative to the run: in (-bits, 0], or pad > 0 for the stream's first word
            // (hi, lo): 64 stream bits, `fill` of them valid from the top.  An element, left-aligned in 32 bits (which
            // also drops anything above its `bits`), lands `fill` bits from the top: two shifts, two ORs.  Staged slots
            // past the vector's end hold zeros, so no bounds test is needed here.
            if (FB > 0 && rel <= 0) {
                // Fixed width: the NE elements i .. i+NE-1 cover the run's 128 bits from any start inside element i.  They
                // are laid end to end in S (constant shifts, resolved at compile time); the run is S shifted left by the
                // start's offset inside element i: four funnel shifts.  (Bits of S past the run may come from a slot the
                // tile does not own: they are never selected.)
                constexpr int NE = (128 + 2 * FB - 2) / (FB > 0 ? FB : 1);
                constexpr int NW = (NE * FB + 31) / 32;
                uint32_t S[NW + 1];
#pragma unroll
                for (int k = 0; k <= NW; ++k) S[k] = 0u;
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    const uint32_t v = FB == 32 ? el[i + e] : (el[i + e] & ((1u << (FB & 31)) - 1u));
                    const int p = e * FB, w = p >> 5, off = p & 31;
                    if (off + FB <= 32) S[w] |= v << ((32 - FB - off) & 31);
                    else { S[w] |= v >> ((off + FB - 32) & 31); S[w + 1] |= v << ((64 - FB - off) & 31); }
                }
                const uint32_t sh = (uint32_t)(-rel);                   // 0 <= sh < FB
                uint32_t r[WB_RUN];
#pragma unroll
                for (int k = 0; k < WB_RUN; ++k) r[k] = __byte_perm(__funnelshift_l(S[k + 1], S[k], sh), 0, 0x0123);
                __stcs(reinterpret_cast<uint4*>(out + w_base + w0), make_uint4(r[0], r[1], r[2], r[3]));
                continue;
            }
            uint32_t hi = 0u, lo = 0u, fill = 0u;
            if (rel > 0) fill = (uint32_t)rel;                          // leading zero bits (the stream's padding)
            else if (rel < 0) {                                         // the first element starts before the run: keep its low bits
                hi = el[i] << (la + (uint32_t)(-rel));                  // drops the bits that belong to the previous word
                fill = bits + (uint32_t)rel; ++i;
            }
            uint32_t r[WB_RUN];
#pragma unroll
            for (int k = 0; k < WB_RUN; ++k) {
                while (fill < 32u) {                                    // one or two elements complete a word at the shipped widths
                    const uint32_t a = el[i] << la;
                    hi |= a >> fill;
                    lo |= __funnelshift_r(0u, a, fill);                 // (a:0) >> fill, low word; 0 when fill == 0
                    fill += bits; ++i;
                }
                r[k] = __byte_perm(hi, 0, 0x0123);                      // first stream byte first
                hi = lo; lo = 0u; fill -= 32u;
            }
            __stcs(reinterpret_cast<uint4*>(out + w_base + w0), make_uint4(r[0], r[1], r[2], r[3]));    // nwords is a multiple of 4
        }
        __syncthreads();                                               // the stage is free for the tile after next
    }
}

// unpack: tile = WU_ELEMS consecutive elements, thread = 4 of them (one 16-byte store)
#define WU_ELEMS (WP_THREADS * 4)
#define WU_BUF (WU_ELEMS + 16)
__global__ void __launch_bounds__(WP_THREADS)
k_wire_unpack32(const uint8_t* __restrict__ in, uint64_t count, uint32_t bits, uint32_t pad, uint64_t nbytes,
                uint32_t* __restrict__ words) {
    __shared__ __align__(16) uint32_t sw[2][WU_BUF];
    const uint32_t fm = bits >= 32u ? 0xffffffffu : ((1u << bits) - 1u);
    const uint64_t ntiles = (count + WU_ELEMS - 1) / WU_ELEMS;

    // stream words [w_al, w_hi) of tile t -> sw[buf], as stored (big-endian); w_al = 16-byte aligned start
    auto issue = [&](uint64_t tile, int buf) {
        const uint64_t j0 = tile * WU_ELEMS;
        const uint32_t ne = (uint32_t)(count - j0 < WU_ELEMS ? count - j0 : WU_ELEMS);
        const uint64_t b_lo = pad + j0 * bits, b_hi = b_lo + (uint64_t)ne * bits;
        const uint64_t w_al = (b_lo >> 5) & ~3ull;
        const uint32_t nq = (uint32_t)((((b_hi + 31) >> 5) - w_al + 3) >> 2);    // 16-byte pieces, <= WU_ELEMS / 4 + 2
        const uint32_t base = smem_u32(&sw[buf][0]);
        for (uint32_t q = threadIdx.x; q < nq; q += WP_THREADS) {
            const uint64_t byte0 = 4 * (w_al + 4ull * q);
            if (byte0 + 16 <= nbytes) {
                cp_async16(base + 16u * q, in + byte0);
            } else {                                                         // the stream's last, partial piece
                for (uint32_t k = 0; k < 4; ++k) {
                    uint32_t v = 0;
                    for (uint32_t b = 0; b < 4; ++b) { const uint64_t a = byte0 + 4 * k + b; v |= (a < nbytes ? (uint32_t)in[a] : 0u) << (8 * b); }
                    sw[buf][4 * q + k] = v;
                }
            }
        }
        cp_async_commit();
    };

    uint64_t t = blockIdx.x;
    if (t >= ntiles) return;
    issue(t, 0);
    for (int buf = 0; t < ntiles; t += gridDim.x, buf ^= 1) {
        if (t + gridDim.x < ntiles) { issue(t + gridDim.x, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        const uint64_t j0 = t * WU_ELEMS;
        const uint32_t ne = (uint32_t)(count - j0 < WU_ELEMS ? count - j0 : WU_ELEMS);
        const uint64_t b_lo = pad + j0 * bits;
        const uint32_t o_base = (uint32_t)(b_lo - (((b_lo >> 5) & ~3ull) << 5));   // bit offset of element j0 in sw[buf]
        const uint32_t e0 = 4u * threadIdx.x;
        if (e0 < ne) {
            const uint32_t* w = sw[buf];
            uint32_t r[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t o = o_base + (e0 + k) * bits;
                const uint32_t wi = o >> 5;
                const uint32_t hi = __byte_perm(w[wi], 0, 0x0123), lo = __byte_perm(w[wi + 1], 0, 0x0123);
                // 32 stream bits from bit (o & 31) of hi:lo; a word past the staged ones only feeds bits that are shifted out
                r[k] = (__funnelshift_l(lo, hi, o & 31u) >> (32u - bits)) & fm;
            }
            uint32_t* dst = words + j0 + e0;
            if (e0 + 4u <= ne && ((uintptr_t)dst & 15u) == 0) __stcs(reinterpret_cast<uint4*>(dst), make_uint4(r[0], r[1], r[2], r[3]));
            else for (uint32_t k = 0; k < 4u && e0 + k < ne; ++k) dst[k] = r[k];
        }
        __syncthreads();
    }
}

// unpack, second form (launched when the stream is 16-byte aligned): tiles of WUB_ELEMS elements, the tile's bytes of
// the stream fetched by ONE bulk asynchronous copy per tile (thread 0, mbarrier completion), two tiles in flight;
// a thread converts four groups of four consecutive elements (one 16-byte store each).
#define WUB_ELEMS 4096
__global__ void __launch_bounds__(WB_THREADS)
k_wire_unpack32_bulk(const uint8_t* __restrict__ in, uint64_t count, uint32_t bits, uint32_t pad, uint64_t nbytes,
                     uint32_t stage_words, uint32_t* __restrict__ words) {
    extern __shared__ __align__(128) uint32_t stage[];            // 2 x stage_words, stream words as stored (big-endian)
    __shared__ __align__(8) uint64_t bars[2];
    const uint32_t fm = bits >= 32u ? 0xffffffffu : ((1u << bits) - 1u);
    const uint64_t ntiles = (count + WUB_ELEMS - 1) / WUB_ELEMS;
    uint64_t t = blockIdx.x;
    if (t >= ntiles) return;
    const uint32_t bar0 = smem_u32(&bars[0]), st0 = smem_u32(stage);
    if (threadIdx.x == 0) {
        mbar_init(bar0, 1); mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](uint64_t tile, int buf) {
        const uint64_t j0 = tile * WUB_ELEMS;
        const uint32_t ne = (uint32_t)(count - j0 < WUB_ELEMS ? count - j0 : WUB_ELEMS);
        const uint64_t b_lo = pad + j0 * bits, b_hi = b_lo + (uint64_t)ne * bits;      // stream bits of the tile
        const uint64_t byte_al = (b_lo >> 3) & ~15ull;
        const uint64_t byte_end = ((b_hi + 7) >> 3) + 4;                                // one word of slack for the funnel shift
        const uint64_t want = (byte_end - byte_al + 15) & ~15ull;
        const uint64_t avail = (nbytes & ~15ull) > byte_al ? (nbytes & ~15ull) - byte_al : 0;
        const uint64_t n_bulk = want <= avail ? want : avail;
        uint8_t* sb = reinterpret_cast<uint8_t*>(stage + (uint32_t)buf * stage_words);
        for (uint64_t i = n_bulk; i < want; ++i) sb[i] = byte_al + i < nbytes ? in[byte_al + i] : (uint8_t)0;   // the stream's last < 16 bytes
        mbar_expect_tx(bar0 + 8u * buf, (uint32_t)n_bulk);
        if (n_bulk) bulk_g2s(st0 + (uint32_t)buf * stage_words * 4u, in + byte_al, (uint32_t)n_bulk, bar0 + 8u * buf);
    };
    if (threadIdx.x == 0) issue(t, 0);
    uint32_t phase[2] = {0u, 0u};
    for (int buf = 0; t < ntiles; t += gridDim.x, buf ^= 1) {
        if (threadIdx.x == 0 && t + gridDim.x < ntiles) issue(t + gridDim.x, buf ^ 1);
        mbar_wait(bar0 + 8u * buf, phase[buf]);
        phase[buf] ^= 1u;
        const uint64_t j0 = t * WUB_ELEMS;
        const uint32_t ne = (uint32_t)(count - j0 < WUB_ELEMS ? count - j0 : WUB_ELEMS);
        const uint64_t b_lo = pad + j0 * bits;
        const uint32_t o_base = (uint32_t)(b_lo - (((b_lo >> 3) & ~15ull) << 3));        // bit offset of element j0 in the stage
        const uint32_t* w = stage + (uint32_t)buf * stage_words;
#pragma unroll
        for (uint32_t q = 0; q < WUB_ELEMS / 4 / WB_THREADS; ++q) {
            const uint32_t e0 = 4u * (threadIdx.x + q * WB_THREADS);
            if (e0 >= ne) break;
            uint32_t r[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t o = o_base + (e0 + k) * bits;
                const uint32_t wi = o >> 5;
                const uint32_t hi = __byte_perm(w[wi], 0, 0x0123), lo = __byte_perm(w[wi + 1], 0, 0x0123);
                r[k] = (__funnelshift_l(lo, hi, o & 31u) >> (32u - bits)) & fm;
            }
            uint32_t* dst = words + j0 + e0;
            if (e0 + 4u <= ne && ((uintptr_t)dst & 15u) == 0) __stcs(reinterpret_cast<uint4*>(dst), make_uint4(r[0], r[1], r[2], r[3]));
            else for (uint32_t k = 0; k < 4u && e0 + k < ne; ++k) dst[k] = r[k];
        }
        __syncthreads();
    }
}

// =================================================================================================
// f2  layer-wise top-s% sparsification with residual accumulation (Client.sparsify,
// proc/jzf_aggregator.py:578-623).  Per layer, in the reference's order:
//     abs_flatten = |x|                                  (:594, BEFORE the residual is added)
//     flatten     = x + remain                           (:595-596, float32)
//     location    = sorted(abs_flatten.argsort()[-k:])   (:598-599)   k = max(1, floor(s * size))
//     compact     = flatten[location]; flatten[location] = 0; remain = flatten   (:600-604)
//     locations  += location + base                      (:606)
// The set `location` is the k largest |x|; elements tied with the k-th largest are taken from the
// HIGHEST indices first (what a stable ascending argsort followed by [-k:] yields; numpy's default
// introsort leaves the choice among exact ties unspecified).
//
// Selection = 3-pass MSD radix select on key = bits(|x|) (monotone for non-negative floats, NaN above
// inf as numpy sorts it): 11 + 11 + 9 bits, block-private shared-memory histograms merged with global
// atomics, one small block picks the bucket.  Compaction keeps index order (= sorted locations): tile
// counts, one-block scan, ordered write.
//
// Two routes to the threshold, chosen per layer ON THE DEVICE, with identical results:
//   candidate route (sparse selections, k <= n/40): a strided sample of the layer (<= 16 K elements) gives a
//     conservative lower bound `lo` (the lower edge of the 11-bit bucket holding the sample's r-th largest key,
//     r ~ 1.5x the expected rank + 6 sigma); ONE pass over x appends every element with key >= lo (2-4 % of
//     the layer) to a candidate list; the three radix passes and the tile counts then run on that list.  x is
//     read twice in total (filter, ordered write) instead of five times and no histogram sees every element
//     (a shared-memory atomic per element costs ~2 clk per lane, 10x the HBM time of the pass).
//   exact route: the radix passes and the tile counts read x itself.  Taken by layers that are too small or
//     too dense for sampling, and by any layer whose candidate list turns out to hold fewer than k elements
//     (bound above the true threshold: skewed sample) or to overflow its region (n/8 slots: heavy ties).
//   Both routes end in the same TopkState (threshold, ties to skip) because the bound is bucket-aligned:
//   every bucket at or above the threshold's is complete in the candidate list.
// =================================================================================================
struct TopkState { uint32_t prefix; uint32_t k_rem; uint32_t c_eq; uint32_t pad; };
// one layer: elements [begin, begin+n) of x, k to keep, first tile number, first output slot; candidate route:
// region [cbegin, cbegin+cap) of the candidate arrays (tiles from ctile0), sample lines (32 elements each)
// numbered from samp0: line i sits at element 32 * (i * samp_stride + jitter(i)); samp_rank = r above.
struct TopkSeg {
    uint64_t begin, tile0, out_off;
    uint32_t n, k;
    uint64_t cbegin, ctile0, samp0;
    uint32_t cap, samp_lines, samp_stride, samp_rank;
    uint32_t small, pad;               // small: the three radix passes of this layer run inside ONE block (k_topk_select_small)
};

#define TK_THREADS 256
#define TK_PER 16
#define TK_TILE (TK_THREADS * TK_PER)
#define TK_BINS 2048
#define TK_SAMPLE_ALL 16384u         // layers up to this size are "sampled" completely
#define TK_SMALL_N (1u << 22)         // candidate-route layers up to this size (<= 2^19 candidates) and ...
#define TK_SMALL_EXACT 65536u        // ... exact-route layers up to this size select inside one block

__device__ __forceinline__ uint32_t key_of(float x) { return __float_as_uint(x) & 0x7fffffffu; }
__device__ __forceinline__ uint32_t key_of_bits(uint32_t w) { return w & 0x7fffffffu; }
// Every kernel of the chain is launched with programmatic stream serialization (topk_launch): it may start while
// its predecessor drains, lets its own successor do the same, and waits here before touching any data.
__device__ __forceinline__ void topk_pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// All layers are processed by the same launches: the tiles (TK_TILE elements, never across a layer
// boundary) of every layer are numbered consecutively and a tile finds its layer by binary search: the LAST
// layer whose first tile is <= the tile (layers without tiles of a kind share their successor's number).
// WHICH: 0 = x tiles (tile0), 1 = candidate tiles (ctile0), 2 = sample lines (samp0).
template <int WHICH>
__device__ __forceinline__ uint64_t topk_first(const TopkSeg& s) { return WHICH == 0 ? s.tile0 : (WHICH == 1 ? s.ctile0 : s.samp0); }
template <int WHICH>
__device__ __forceinline__ int topk_seg_of(const TopkSeg* __restrict__ segs, int nseg, uint64_t t) {
    int lo = 0, hi = nseg - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (topk_first<WHICH>(segs[mid]) <= t) lo = mid; else hi = mid - 1; }
    return lo;
}
// The route of a layer, decided by its candidate count once the filter pass has run.
__device__ __forceinline__ bool topk_exact_route(const TopkSeg& s, uint32_t count) { return s.cap == 0u || count > s.cap || count < s.k; }
// What a histogram / count pass reads for a layer: CAND = the candidate keys of a candidate-route layer,
// !CAND = x itself for an exact-route layer; n = 0 when the layer belongs to the other route.
struct TopkView { uint64_t begin; uint64_t tile0; uint32_t n; };
// RADIX: the view of the multi-block histogram passes, which leave the small layers to k_topk_select_small.
template <bool CAND, bool RADIX = false>
__device__ __forceinline__ TopkView topk_view(const TopkSeg& s, const uint32_t* __restrict__ cand_count, int idx) {
    const uint32_t count = cand_count ? cand_count[idx] : 0u;
    const bool exact = topk_exact_route(s, count);
    TopkView v;
    if (CAND) { v.begin = s.cbegin; v.tile0 = s.ctile0; v.n = exact ? 0u : count; }
    else { v.begin = s.begin; v.tile0 = s.tile0; v.n = exact ? s.n : 0u; }
    if (RADIX && s.small) v.n = 0u;
    return v;
}

__global__ void k_topk_init(const TopkSeg* __restrict__ segs, int nseg, TopkState* __restrict__ st) {
    topk_pdl_enter();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < nseg) { TopkState v; v.prefix = 0u; v.k_rem = segs[s].k; v.c_eq = 0u; v.pad = 0u; st[s] = v; }
}

__device__ __forceinline__ uint32_t topk_hash(uint32_t i) { i *= 0x9E3779B1u; i ^= i >> 15; i *= 0x85EBCA77u; i ^= i >> 13; return i; }

// Sample histogram: one warp per sample line (32 consecutive elements, one or two 128-byte lines), top 11 key bits.
__global__ void __launch_bounds__(256)
k_topk_sample(const uint32_t* __restrict__ x, const TopkSeg* __restrict__ segs, int nseg, uint64_t nlines, uint32_t* __restrict__ hist) {
    topk_pdl_enter();
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t line = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; line < nlines; line += nwarps) {
        const int s = topk_seg_of<2>(segs, nseg, line);
        const TopkSeg sg = segs[s];
        const uint32_t i = (uint32_t)(line - sg.samp0);
        const uint32_t jit = sg.samp_stride > 1u ? topk_hash(i) % sg.samp_stride : 0u;
        const uint64_t e = ((uint64_t)i * sg.samp_stride + jit) * 32u + lane;
        if (e < sg.n) atomicAdd(&hist[(size_t)s * TK_BINS + (key_of_bits(x[sg.begin + e]) >> 20)], 1u);
    }
}

// One block of 1024 threads per layer: suffix sums of the histogram from the top bin.  Clears the histogram.
//   lo_out == NULL  the bucket holding the k_rem-th largest key extends the layer's prefix (radix passes);
//   lo_out != NULL  the lower edge of the bucket holding the sample's samp_rank-th largest key -> lo_out[layer].
__global__ void __launch_bounds__(1024)
k_topk_pick(uint32_t* __restrict__ hist_all, TopkState* __restrict__ st_all, uint32_t shift, const TopkSeg* __restrict__ segs,
            uint32_t* __restrict__ lo_out) {
    __shared__ uint32_t warp_tot[32];
    topk_pdl_enter();
    uint32_t* hist = hist_all + (size_t)blockIdx.x * TK_BINS;
    TopkState* st = st_all + blockIdx.x;
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    // thread t owns bins hi = TK_BINS-1-2t and hi-1 (descending order)
    const uint32_t b0 = TK_BINS - 1 - 2 * t, b1 = b0 - 1;
    const uint32_t c0 = hist[b0], c1 = hist[b1];
    hist[b0] = 0; hist[b1] = 0;
    uint32_t incl = c0 + c1;
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += o; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_tot[lane];
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, w, d); if (lane >= (uint32_t)d) w += o; }
        warp_tot[lane] = w;
    }
    __syncthreads();
    const uint32_t before = (warp ? warp_tot[warp - 1] : 0u) + incl - (c0 + c1);   // keys in higher bins
    const uint32_t k_rem = lo_out ? segs[blockIdx.x].samp_rank : st->k_rem;
    __syncthreads();
    // the k_rem-th largest lies in the first bin (from the top) whose inclusive count reaches k_rem
    const bool in0 = before < k_rem && k_rem <= before + c0;
    const bool in1 = before + c0 < k_rem && k_rem <= before + c0 + c1;
    if (lo_out) {
        if (in0) lo_out[blockIdx.x] = b0 << 20; else if (in1) lo_out[blockIdx.x] = b1 << 20;   // (stays 0 when the sample is shorter than the rank)
    } else if (in0) {
        st->prefix |= b0 << shift; st->k_rem = k_rem - before; st->c_eq = c0;
    } else if (in1) {
        st->prefix |= b1 << shift; st->k_rem = k_rem - before - c0; st->c_eq = c1;
    }
}

// The sample bound of a layer in ONE block: k_topk_sample's histogram of the top 11 key bits over the layer's sample
// lines (at most 512 lines = 16 K elements whatever the layer's size) in shared memory, then k_topk_pick's suffix scan
// for the bucket of the samp_rank-th largest sample -> lo_out[layer] (stays 0 when the sample is shorter than the rank
// or the layer is not sampled).  One launch instead of two, no global histogram traffic.
__global__ void __launch_bounds__(1024)
k_topk_bound(const uint32_t* __restrict__ x, const TopkSeg* __restrict__ segs, uint32_t* __restrict__ lo_out) {
    __shared__ uint32_t sh[TK_BINS];
    __shared__ uint32_t warp_tot[32];
    topk_pdl_enter();
    const TopkSeg sg = segs[blockIdx.x];
    if (sg.samp_lines == 0u) return;
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    sh[t] = 0u; sh[t + 1024u] = 0u;
    __syncthreads();
    // at most 512 lines (host: TK_SAMPLE_ALL / 32 and the cap on `want`): 16 per warp, all loads in flight before the counts
    for (uint32_t i0 = warp; i0 < sg.samp_lines; i0 += 32u * 16u) {
        uint32_t v[16]; uint32_t ok = 0u;
#pragma unroll
        for (uint32_t q = 0; q < 16u; ++q) {
            const uint32_t i = i0 + 32u * q;
            if (i < sg.samp_lines) {
                const uint32_t jit = sg.samp_stride > 1u ? topk_hash(i) % sg.samp_stride : 0u;
                const uint64_t e = ((uint64_t)i * sg.samp_stride + jit) * 32u + lane;
                if (e < sg.n) { v[q] = x[sg.begin + e]; ok |= 1u << q; }
            }
        }
#pragma unroll
        for (uint32_t q = 0; q < 16u; ++q) if ((ok >> q) & 1u) atomicAdd(&sh[key_of_bits(v[q]) >> 20], 1u);
    }
    __syncthreads();
    const uint32_t b0 = TK_BINS - 1 - 2 * t, b1 = b0 - 1;
    const uint32_t c0 = sh[b0], c1 = sh[b1];
    uint32_t incl = c0 + c1;
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += o; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_tot[lane];
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, w, d); if (lane >= (uint32_t)d) w += o; }
        warp_tot[lane] = w;
    }
    __syncthreads();
    const uint32_t before = (warp ? warp_tot[warp - 1] : 0u) + incl - (c0 + c1);
    const uint32_t k_rem = sg.samp_rank;
    if (before < k_rem && k_rem <= before + c0) lo_out[blockIdx.x] = b0 << 20;
    else if (before + c0 < k_rem && k_rem <= before + c0 + c1) lo_out[blockIdx.x] = b1 << 20;
}

// Small layers: all three radix passes and their picks inside ONE block of 1024 threads per layer - the keys of the
// layer's view (its candidate list, or x itself on the exact route) are read three times by the same block, the
// histogram lives in shared memory, nothing is exchanged between blocks.  A model of many layers spent six launches
// (~80 us for 200 layers of 250 k, most of it launch and pick latency) on a few thousand candidates per layer.
// Leaves the same TopkState as k_topk_hist + k_topk_pick x 3.
__global__ void __launch_bounds__(1024)
k_topk_select_small(const uint32_t* __restrict__ x, const uint32_t* __restrict__ cand_key, const uint32_t* __restrict__ cand_idx,
                    const TopkSeg* __restrict__ segs, const uint32_t* __restrict__ cand_count, TopkState* __restrict__ st_all,
                    uint2* __restrict__ tile_counts) {
    __shared__ uint32_t sh[TK_BINS];
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t s_prefix, s_krem, s_ceq;
    topk_pdl_enter();
    const TopkSeg sg = segs[blockIdx.x];
    if (sg.small != 1u) return;
    const uint32_t count = cand_count ? cand_count[blockIdx.x] : 0u;
    const bool exact = topk_exact_route(sg, count);
    const uint32_t* __restrict__ keys = exact ? x + sg.begin : cand_key + sg.cbegin;
    const uint32_t n = exact ? sg.n : count;
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    if (t == 0) { s_prefix = 0u; s_krem = sg.k; s_ceq = 0u; }
    const uint32_t shifts[3] = {20u, 9u, 0u}, nbits[3] = {11u, 11u, 9u};
#pragma unroll 1
    for (int p = 0; p < 3; ++p) {
        sh[t] = 0u; sh[t + 1024u] = 0u;
        __syncthreads();
        const uint32_t prefix = s_prefix, k_rem = s_krem;
        const uint32_t shift = shifts[p], hi_shift = shift + nbits[p], dmask = (1u << nbits[p]) - 1u;
        for (uint32_t i0 = t; i0 < n; i0 += 8u * 1024u) {                     // eight loads in flight per thread
            uint32_t kq[8];
#pragma unroll
            for (uint32_t q = 0; q < 8u; ++q) { const uint32_t i = i0 + q * 1024u; kq[q] = i < n ? key_of_bits(keys[i]) : 0xffffffffu; }
#pragma unroll
            for (uint32_t q = 0; q < 8u; ++q)
                if (kq[q] != 0xffffffffu && (hi_shift >= 31u || (kq[q] >> hi_shift) == (prefix >> hi_shift))) atomicAdd(&sh[(kq[q] >> shift) & dmask], 1u);
        }
        __syncthreads();
        // the pick of k_topk_pick: thread t owns bins TK_BINS-1-2t and the one below, suffix sums from the top bin
        const uint32_t b0 = TK_BINS - 1 - 2 * t, b1 = b0 - 1;
        const uint32_t c0 = sh[b0], c1 = sh[b1];
        uint32_t incl = c0 + c1;
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += o; }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_tot[lane];
            for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, w, d); if (lane >= (uint32_t)d) w += o; }
            warp_tot[lane] = w;
        }
        __syncthreads();
        const uint32_t before = (warp ? warp_tot[warp - 1] : 0u) + incl - (c0 + c1);
        const bool in0 = before < k_rem && k_rem <= before + c0;
        const bool in1 = before + c0 < k_rem && k_rem <= before + c0 + c1;
        if (in0) { s_prefix = prefix | (b0 << shift); s_krem = k_rem - before; s_ceq = c0; }
        else if (in1) { s_prefix = prefix | (b1 << shift); s_krem = k_rem - before - c0; s_ceq = c1; }
        __syncthreads();
    }
    if (t == 0) { TopkState v; v.prefix = s_prefix; v.k_rem = s_krem; v.c_eq = s_ceq; v.pad = 0u; st_all[blockIdx.x] = v; }
    // ... and the per-tile counts of (key > T, key == T) that k_topk_count would make for this layer (tile_counts starts zeroed)
    const uint32_t T = s_prefix;
    for (uint32_t i = t; i < n; i += 1024u) {
        const uint32_t key = key_of_bits(keys[i]);
        if (key >= T) {
            const uint32_t e = exact ? i : cand_idx[sg.cbegin + i];       // element index inside the layer
            atomicAdd(reinterpret_cast<uint32_t*>(&tile_counts[sg.tile0 + e / TK_TILE]) + (key == T ? 1 : 0), 1u);
        }
    }
}

// Sixteen keys of a tile per thread, four rows of one 16-byte quad each: row r, thread t holds elements
// 4 * (256 r + t) .. + 3 of the tile, so that a warp covers 512 contiguous bytes per access.  `valid` = bit per key.
__device__ __forceinline__ uint32_t topk_load_tile(const uint32_t* __restrict__ xs, uint64_t base, uint32_t n, bool vec, uint32_t (&w)[TK_PER]) {
    uint32_t valid = 0u;
#pragma unroll
    for (int r = 0; r < TK_PER / 4; ++r) {
        const uint64_t i = base + 4ull * ((uint64_t)r * TK_THREADS + threadIdx.x);
        if (vec && i + 4 <= n) {
            const uint4 v = __ldcs(reinterpret_cast<const uint4*>(xs + i));
            w[4 * r] = v.x; w[4 * r + 1] = v.y; w[4 * r + 2] = v.z; w[4 * r + 3] = v.w;
            valid |= 0xfu << (4 * r);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                w[4 * r + k] = 0u;
                if (i + k < n) { w[4 * r + k] = xs[i + k]; valid |= 1u << (4 * r + k); }
            }
        }
    }
    return valid;
}

// element k (dynamic) of a thread's sixteen registers: a select tree (used on the rare hit paths only)
template <typename T>
__device__ __forceinline__ T topk_pick16(const T (&v)[TK_PER], uint32_t k) {
    const bool b0 = k & 1u, b1 = k & 2u, b2 = k & 4u, b3 = k & 8u;
    const T a0 = b0 ? v[1] : v[0], a1 = b0 ? v[3] : v[2], a2 = b0 ? v[5] : v[4], a3 = b0 ? v[7] : v[6];
    const T a4 = b0 ? v[9] : v[8], a5 = b0 ? v[11] : v[10], a6 = b0 ? v[13] : v[12], a7 = b0 ? v[15] : v[14];
    const T c0 = b1 ? a1 : a0, c1 = b1 ? a3 : a2, c2 = b1 ? a5 : a4, c3 = b1 ? a7 : a6;
    const T d0 = b2 ? c1 : c0, d1 = b2 ? c3 : c2;
    return b3 ? d1 : d0;
}
// index inside the tile of a thread's element k (topk_load_tile's layout)
__device__ __forceinline__ uint32_t topk_elem_of(uint32_t k) { return 4u * ((k >> 2) * TK_THREADS + threadIdx.x) + (k & 3u); }

// Candidate filter: every element with key >= lo[layer] is appended (key, index within the layer) to the layer's
// region.  A block owns a contiguous run of tiles (the layer table is walked, not searched) and collects its
// candidates in shared memory; the region's cursor is bumped once per flush (buffer full, layer change, end of the
// run) instead of once per tile, so no tile waits for the round trip of an atomic, and the flush writes are
// coalesced.  The cursor keeps counting past the region's end, which is how the later passes see an overflow.
// The loads of the next tile are in flight while the current one is scanned; the few candidates of a thread (2-4 %
// of the elements) are visited through the set bits of its hit mask.
#define TK_FBUF 2048u
#define TKF_STAGES 3
__global__ void __launch_bounds__(TK_THREADS)
k_topk_filter(const uint32_t* __restrict__ x, const TopkSeg* __restrict__ segs, int nseg, uint64_t ntiles, uint64_t tiles_per_block,
              const uint32_t* __restrict__ lo_all, uint32_t* __restrict__ cand_count, uint32_t* __restrict__ cand_key, uint32_t* __restrict__ cand_idx) {
    extern __shared__ __align__(128) uint32_t tkf_ring[];             // TKF_STAGES tiles of x (whole tiles of aligned layers, bulk copy engine)
    __shared__ __align__(8) uint64_t full[TKF_STAGES];
    __shared__ uint2 buf[TK_FBUF];
    __shared__ uint32_t wtot[2][TK_THREADS / 32];
    __shared__ uint32_t slot0;
    topk_pdl_enter();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint64_t t_begin = (uint64_t)blockIdx.x * tiles_per_block;
    uint64_t t_end = t_begin + tiles_per_block;
    if (t_end > ntiles) t_end = ntiles;
    if (t_begin >= t_end) return;
    const uint32_t bar0 = smem_u32(&full[0]), ring0 = smem_u32(tkf_ring);
    if (threadIdx.x == 0) {
        for (int q = 0; q < TKF_STAGES; ++q) mbar_init(bar0 + 8u * q, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    int s = topk_seg_of<0>(segs, nseg, t_begin);
    TopkSeg sg = segs[s];
    uint32_t lo = lo_all[s];
    // Per-layer state in 32 bits: tile index inside the layer, its tile count, how many of them are whole, whether the
    // layer's tiles can come through the ring.  Tiles are consecutive and every layer has at least one, so the layer
    // of the next tile is s or s + 1: the table is walked without a search and without 64-bit arithmetic per tile.
    auto layer_bulk = [&](const TopkSeg& g) { return g.cap != 0u && (((uintptr_t)(x + g.begin)) & 15u) == 0; };
    uint32_t lt = (uint32_t)(t_begin - sg.tile0), ntl = (sg.n + TK_TILE - 1u) / TK_TILE, nfull = sg.n / TK_TILE;
    bool lbulk = layer_bulk(sg);
    const uint32_t nrun = (uint32_t)(t_end - t_begin);
    // thread 0: the same cursor for the tile being issued, TKF_STAGES - 1 tiles ahead
    int s_is = s;
    uint32_t lt_is = lt, ntl_is = ntl, nfull_is = nfull, q_is = 0u;
    bool lbulk_is = lbulk;
    uint64_t e_is = sg.begin + (uint64_t)lt * TK_TILE;                  // first element of the tile being issued
    auto issue_next = [&]() {                                           // issues the cursor's tile and advances the cursor
        if (lbulk_is && lt_is < nfull_is) {
            mbar_expect_tx(bar0 + 8u * q_is, TK_TILE * 4u);
            bulk_g2s(ring0 + q_is * (TK_TILE * 4u), x + e_is, TK_TILE * 4u, bar0 + 8u * q_is);
        }
        q_is = q_is + 1u == TKF_STAGES ? 0u : q_is + 1u;
        if (++lt_is == ntl_is) {
            if (++s_is < nseg) {
                const TopkSeg g = segs[s_is];
                lt_is = 0u; ntl_is = (g.n + TK_TILE - 1u) / TK_TILE; nfull_is = g.n / TK_TILE; lbulk_is = layer_bulk(g); e_is = g.begin;
            }
        } else e_is += TK_TILE;
    };
    __syncthreads();                                                    // barriers initialised
    if (threadIdx.x == 0)
        for (uint32_t j = 0; j < nrun && j < TKF_STAGES - 1u; ++j) issue_next();
    uint32_t phase = 0u;                                                // one parity bit per stage
    uint32_t fill = 0u;                                                 // (block-uniform) entries of layer s waiting in buf
    // claims `count` slots of layer s; returns the first one, or 0xffffffff when the region cannot hold them
    auto claim = [&](uint32_t count) -> uint32_t {
        __syncthreads();                                                // buf complete / slot0 free
        if (threadIdx.x == 0) slot0 = atomicAdd(&cand_count[s], count);
        __syncthreads();
        const uint32_t first = slot0;
        return first + count <= sg.cap ? first : 0xffffffffu;
    };
    auto flush = [&]() {
        if (fill == 0u) return;
        const uint32_t first = claim(fill);
        if (first != 0xffffffffu)
            for (uint32_t j = threadIdx.x; j < fill; j += TK_THREADS) { const uint2 c = buf[j]; cand_key[sg.cbegin + first + j] = c.x; cand_idx[sg.cbegin + first + j] = c.y; }
        __syncthreads();                                                // buf may be refilled
        fill = 0u;
    };
    uint32_t q = 0u;                                                    // ring stage of the current tile
    for (uint32_t j = 0; j < nrun; ++j) {
        const bool more = j + (TKF_STAGES - 1u) < nrun;                 // a tile to issue: into the stage of tile j - 1
        if (sg.cap) {                                                   // (block-uniform)
            const bool bulk = lbulk && lt < nfull;
            const uint32_t* ring = tkf_ring + q * TK_TILE;
            uint32_t w[TK_PER], valid;
            if (bulk) {
                mbar_wait(bar0 + 8u * q, (phase >> q) & 1u);
                phase ^= 1u << q;
                const uint4* xs4 = reinterpret_cast<const uint4*>(ring);
#pragma unroll
                for (int r = 0; r < TK_PER / 4; ++r) {
                    const uint4 v = xs4[r * TK_THREADS + threadIdx.x];
                    w[4 * r] = v.x; w[4 * r + 1] = v.y; w[4 * r + 2] = v.z; w[4 * r + 3] = v.w;
                }
                valid = 0xffffu;
            } else {
                valid = topk_load_tile(x + sg.begin, (uint64_t)lt * TK_TILE, sg.n, (((uintptr_t)(x + sg.begin)) & 15u) == 0, w);
            }
            const uint32_t base = lt * TK_TILE;
            const float lo_f = __uint_as_float(lo);
            uint32_t sel = 0u;
#pragma unroll
            for (int k = 0; k < TK_PER; ++k) sel |= (uint32_t)(!(fabsf(__uint_as_float(w[k])) < lo_f)) << k;   // key >= lo, NaN included (its key is above inf)
            sel &= valid;
            const uint32_t cnt = __popc(sel);
            uint32_t incl = cnt;
            for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += o; }
            const uint32_t pb = j & 1u;                                 // alternating shared slots: one barrier per tile
            if (lane == 31) wtot[pb][warp] = incl;
            __syncthreads();                                            // ... which also ends every read of tile j - 1's stage
            if (threadIdx.x == 0 && more) issue_next();
            uint32_t before = incl - cnt, total = 0u;
#pragma unroll
            for (uint32_t v = 0; v < TK_THREADS / 32; ++v) { const uint32_t c = wtot[pb][v]; if (v < warp) before += c; total += c; }
            // a candidate's key: still in the ring stage for a bulk tile (one shared load), else picked from the registers
            auto key_at = [&](uint32_t k) { return key_of_bits(bulk ? ring[topk_elem_of(k)] : topk_pick16(w, k)); };
            if (fill + total > TK_FBUF) flush();
            if (total > TK_FBUF) {                                      // a tile denser than the buffer (bound too low): straight to the region
                const uint32_t first = claim(total);
                if (first != 0xffffffffu) {
                    uint32_t m = sel, jj = first + before;
                    while (m) {
                        const uint32_t k = (uint32_t)__ffs((int)m) - 1u;
                        m &= m - 1u;
                        cand_key[sg.cbegin + jj] = key_at(k);
                        cand_idx[sg.cbegin + jj] = base + topk_elem_of(k);
                        ++jj;
                    }
                }
            } else {
                uint32_t m = sel, jj = fill + before;
                while (m) {
                    const uint32_t k = (uint32_t)__ffs((int)m) - 1u;
                    m &= m - 1u;
                    buf[jj] = make_uint2(key_at(k), base + topk_elem_of(k));
                    ++jj;
                }
                fill += total;
            }
        } else {
            __syncthreads();                                            // every read of tile j - 1's stage has ended
            if (threadIdx.x == 0 && more) issue_next();
        }
        q = q + 1u == TKF_STAGES ? 0u : q + 1u;
        if (++lt == ntl && j + 1u < nrun) {                             // the run enters the next layer
            flush();                                                    // the buffer holds one layer's candidates
            ++s;
            sg = segs[s]; lo = lo_all[s];
            lt = 0u; ntl = (sg.n + TK_TILE - 1u) / TK_TILE; nfull = sg.n / TK_TILE; lbulk = layer_bulk(sg);
        }
    }
    flush();
}

// pass 0: shift 20 / 11 bits; pass 1: shift 9 / 11 bits; pass 2: shift 0 / 9 bits.
// A block owns a contiguous run of tiles; its shared histogram is merged into the layer's global one
// whenever the run crosses into the next layer (and at the end).  CAND: the tiles are those of the candidate
// regions and `x` is the candidate key array; otherwise the tiles of x itself (exact-route layers only).
template <bool CAND>
__device__ __forceinline__ void topk_hist_run(uint32_t* sh, uint32_t bid, const uint32_t* __restrict__ x, const TopkSeg* __restrict__ segs, int nseg,
                                              uint64_t ntiles, uint64_t tiles_per_block, const TopkState* __restrict__ st,
                                              const uint32_t* __restrict__ cand_count, uint32_t shift, uint32_t nbits, uint32_t* __restrict__ hist) {
    const uint64_t t_begin = (uint64_t)bid * tiles_per_block;
    uint64_t t_end = t_begin + tiles_per_block;
    if (t_end > ntiles) t_end = ntiles;
    if (t_begin >= t_end) return;
    for (uint32_t i = threadIdx.x; i < TK_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const uint32_t hi_shift = shift + nbits;               // bits above the digit must equal the prefix
    const uint32_t dmask = (1u << nbits) - 1u;
    int s = topk_seg_of<CAND ? 1 : 0>(segs, nseg, t_begin);
    bool dirty = false;                                    // (block-uniform) the shared histogram holds counts of layer s
    for (uint64_t tile = t_begin; tile < t_end; ++tile) {
        while (s + 1 < nseg && topk_first<CAND ? 1 : 0>(segs[s + 1]) <= tile) {   // the run enters the next layer
            if (dirty) {
                __syncthreads();
                for (uint32_t i = threadIdx.x; i < TK_BINS; i += blockDim.x) {
                    const uint32_t c = sh[i];
                    if (c) { atomicAdd(&hist[(size_t)s * TK_BINS + i], c); sh[i] = 0; }
                }
                __syncthreads();
                dirty = false;
            }
            ++s;
        }
        const TopkView vw = topk_view<CAND, true>(segs[s], cand_count, s);
        const uint64_t base = (tile - vw.tile0) * TK_TILE;
        if (base >= vw.n) {                                // other route (or a small layer), or past the candidates that exist:
            if (s + 1 >= nseg) break;                      // on to the next layer's first tile
            const uint64_t nxt = topk_first<CAND ? 1 : 0>(segs[s + 1]);
            tile = (nxt > tile ? nxt : tile + 1) - 1;
            continue;
        }
        dirty = true;
        const uint32_t prefix = st[s].prefix;
        const uint32_t* xs = x + vw.begin;
        uint32_t w[TK_PER];
        const uint32_t valid = topk_load_tile(xs, base, vw.n, (((uintptr_t)xs) & 15u) == 0, w);   // element order does not matter here
#pragma unroll
        for (int k = 0; k < TK_PER; ++k) {
            const uint32_t key = key_of_bits(w[k]);
            if (((valid >> k) & 1u) && (hi_shift >= 31u || (key >> hi_shift) == (prefix >> hi_shift))) atomicAdd(&sh[(key >> shift) & dmask], 1u);
        }
    }
    if (!dirty) return;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < TK_BINS; i += blockDim.x) {
        const uint32_t c = sh[i];
        if (c) atomicAdd(&hist[(size_t)s * TK_BINS + i], c);
    }
}
// ONE launch per radix pass for both routes: the first `gc` blocks walk the candidate tiles, the others the tiles of x
// (where they find nothing to do unless a layer takes the exact route).
__global__ void __launch_bounds__(TK_THREADS)
k_topk_hist(const uint32_t* __restrict__ x, const uint32_t* __restrict__ cand_key, const TopkSeg* __restrict__ segs, int nseg,
            uint32_t gc, uint64_t nctiles, uint64_t ctpb, uint64_t ntiles, uint64_t tpb,
            const TopkState* __restrict__ st, const uint32_t* __restrict__ cand_count, uint32_t shift, uint32_t nbits, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[TK_BINS];
    topk_pdl_enter();
    if (blockIdx.x < gc) topk_hist_run<true>(sh, blockIdx.x, cand_key, segs, nseg, nctiles, ctpb, st, cand_count, shift, nbits, hist);
    else topk_hist_run<false>(sh, blockIdx.x - gc, x, segs, nseg, ntiles, tpb, st, cand_count, shift, nbits, hist);
}

// Per-tile counts of (key > T, key == T), both routes in one launch.  Exact route (blocks >= gc): read from x; a block
// owns a contiguous run of tiles and jumps over the layers of the candidate route.  Candidate route (blocks < gc): see
// topk_cand_count_run.
__device__ __forceinline__ void topk_count_run(uint32_t bid, const uint32_t* __restrict__ x, const TopkSeg* __restrict__ segs, int nseg, uint64_t ntiles,
                                               uint64_t tiles_per_block, const TopkState* __restrict__ st, const uint32_t* __restrict__ cand_count,
                                               uint2* __restrict__ tile_counts) {
    __shared__ uint32_t sg_[TK_THREADS / 32], se_[TK_THREADS / 32];
    const uint64_t t_begin = (uint64_t)bid * tiles_per_block;
    uint64_t t_end = t_begin + tiles_per_block;
    if (t_end > ntiles) t_end = ntiles;
    if (t_begin >= t_end) return;
    int s = topk_seg_of<0>(segs, nseg, t_begin);
    for (uint64_t tile = t_begin; tile < t_end; ++tile) {
        while (s + 1 < nseg && segs[s + 1].tile0 <= tile) ++s;
        const TopkView vw = topk_view<false, true>(segs[s], cand_count, s);
        if (vw.n == 0u) {                                   // candidate route (k_topk_cand_count fills these tiles) or a small layer
            if (s + 1 >= nseg) break;
            const uint64_t nxt = segs[s + 1].tile0;
            tile = (nxt > tile ? nxt : tile + 1) - 1;
            continue;
        }
        const uint32_t T = st[s].prefix;
        const uint64_t base = (tile - vw.tile0) * TK_TILE;
        const uint32_t* xs = x + vw.begin;
        uint32_t w[TK_PER];
        const uint32_t valid = topk_load_tile(xs, base, vw.n, (((uintptr_t)xs) & 15u) == 0, w);   // counts do not need index order
        uint32_t g = 0, e = 0;
#pragma unroll
        for (int k = 0; k < TK_PER; ++k) {
            const uint32_t key = key_of_bits(w[k]), ok = (valid >> k) & 1u;
            g += ok & (uint32_t)(key > T); e += ok & (uint32_t)(key == T);
        }
        for (int d = 16; d > 0; d >>= 1) { g += __shfl_down_sync(0xffffffffu, g, d); e += __shfl_down_sync(0xffffffffu, e, d); }
        if ((threadIdx.x & 31u) == 0) { sg_[threadIdx.x >> 5] = g; se_[threadIdx.x >> 5] = e; }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t G = 0, E = 0;
            for (int v = 0; v < TK_THREADS / 32; ++v) { G += sg_[v]; E += se_[v]; }
            tile_counts[tile] = make_uint2(G, E);
        }
        __syncthreads();
    }
}

// candidate route: the same per-tile counts from the candidate list (every element at or above the threshold is
// a candidate); tile_counts starts zeroed.  The blocks are dealt to the layers directly (P blocks per layer, layers
// round-robin over the nblk / P slots): the list lengths are only known on the device, and a search per work unit
// through a table of mostly empty regions cost more than the counting.
__device__ __forceinline__ void topk_cand_count_run(uint32_t bid, uint32_t nblk, const uint32_t* __restrict__ cand_key, const uint32_t* __restrict__ cand_idx,
                                                    const TopkSeg* __restrict__ segs, int nseg, const TopkState* __restrict__ st,
                                                    const uint32_t* __restrict__ cand_count, uint2* __restrict__ tile_counts) {
    const uint32_t P = nblk / (uint32_t)nseg > 0u ? nblk / (uint32_t)nseg : 1u, slots = nblk / P;
    const uint32_t slot = bid / P, j = bid % P;
    if (slot >= slots) return;
    for (uint32_t s = slot; s < (uint32_t)nseg; s += slots) {
        const TopkSeg sg = segs[s];
        const TopkView vw = topk_view<true, true>(sg, cand_count, (int)s);   // (small layers: counted by k_topk_select_small)
        if (vw.n == 0u) continue;
        const uint32_t T = st[s].prefix;
        for (uint32_t i = j * TK_THREADS + threadIdx.x; i < vw.n; i += P * TK_THREADS) {
            const uint32_t key = cand_key[vw.begin + i];
            if (key >= T) {
                uint32_t* tc = reinterpret_cast<uint32_t*>(&tile_counts[sg.tile0 + cand_idx[vw.begin + i] / TK_TILE]);
                atomicAdd(tc + (key == T ? 1 : 0), 1u);
            }
        }
    }
}
__global__ void __launch_bounds__(TK_THREADS)
k_topk_count(const uint32_t* __restrict__ x, const uint32_t* __restrict__ cand_key, const uint32_t* __restrict__ cand_idx, const TopkSeg* __restrict__ segs, int nseg,
             uint32_t gc, uint64_t nctiles, uint64_t ntiles, uint64_t tpb, const TopkState* __restrict__ st, const uint32_t* __restrict__ cand_count,
             uint2* __restrict__ tile_counts) {
    topk_pdl_enter();
    if (blockIdx.x < gc) topk_cand_count_run(blockIdx.x, gc, cand_key, cand_idx, segs, nseg, st, cand_count, tile_counts);
    else topk_count_run(blockIdx.x - gc, x, segs, nseg, ntiles, tpb, st, cand_count, tile_counts);
}

// Ordered write.  A tile is four rows of 256 16-byte quads (topk_load_tile's layout: every access of a warp
// covers 512 contiguous bytes of x, the residual and the new residual); the index order inside the tile is
// row-major, so a thread's output slots follow from a block scan of its per-row counts, packed 16 bits per row
// into one 64-bit word each for (key > T) and (key == T).  An element is selected when key > T, or key == T and
// at least c_eq - k_rem tied elements precede it.  Output slot = (#selected before it in its layer) + out_off.
// Every row leaves as x + residual in one 16-byte store; the ~1 % of elements at or above the threshold are then
// visited through the set bits of the thread's hit mask (compact output, zero into the new residual: a later store
// of the same thread to the same address).  A block owns a contiguous run of tiles; the loads of the next tile are
// in flight while the current one is scanned and written.
struct TopkTile { uint32_t w[TK_PER]; float rv[TK_PER]; uint32_t valid; bool vec; };
__device__ __forceinline__ void topk_write_load(TopkTile& t, const uint32_t* __restrict__ x, const float* __restrict__ res_in, const float* res_out,
                                                const TopkSeg& sg, uint64_t tile) {
    const uint32_t* xs = x + sg.begin;
    const uint64_t base = (tile - sg.tile0) * TK_TILE;
    t.vec = ((((uintptr_t)xs) | (res_in ? (uintptr_t)(res_in + sg.begin) : 0) | (res_out ? (uintptr_t)(res_out + sg.begin) : 0)) & 15u) == 0;
    t.valid = topk_load_tile(xs, base, sg.n, t.vec, t.w);
    if (res_in) {
#pragma unroll
        for (int r = 0; r < TK_PER / 4; ++r) {
            const uint64_t i = base + 4ull * ((uint64_t)r * TK_THREADS + threadIdx.x);
            const float* rp = res_in + sg.begin + i;
            if (t.vec && i + 4 <= sg.n) {
                const float4 v = __ldcs(reinterpret_cast<const float4*>(rp));
                t.rv[4 * r] = v.x; t.rv[4 * r + 1] = v.y; t.rv[4 * r + 2] = v.z; t.rv[4 * r + 3] = v.w;
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) t.rv[4 * r + k] = i + k < sg.n ? rp[k] : 0.0f;
            }
        }
    }
}
// Whole tiles of 16-byte aligned layers reach the block through a ring of TKW_STAGES shared-memory stages filled by
// the bulk copy engine (cp.async.bulk + mbarrier, issued TKW_STAGES - 1 tiles ahead by thread 0): the prefetch holds
// no registers, so two CTAs per SM keep six tiles (192 KB) in flight.  Partial tiles and unaligned layers are read
// straight from global memory.
#define TKW_STAGES 3
__global__ void __launch_bounds__(TK_THREADS)
k_topk_write(const uint32_t* __restrict__ x, const float* __restrict__ res_in, const TopkSeg* __restrict__ segs, int nseg,
             uint64_t ntiles, uint64_t tiles_per_block, const TopkState* __restrict__ st_all, const uint2* __restrict__ tile_counts,
             float* __restrict__ values_all, int64_t* __restrict__ index_all, float* res_out) {
    extern __shared__ __align__(128) uint32_t tkw_ring[];          // TKW_STAGES x (x tile | residual tile)
    __shared__ __align__(8) uint64_t full[TKW_STAGES];
    __shared__ uint64_t wg[2][TK_THREADS / 32], we[2][TK_THREADS / 32];
    topk_pdl_enter();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint64_t t_begin = (uint64_t)blockIdx.x * tiles_per_block;
    uint64_t t_end = t_begin + tiles_per_block;
    if (t_end > ntiles) t_end = ntiles;
    if (t_begin >= t_end) return;
    const uint32_t bar0 = smem_u32(&full[0]), ring0 = smem_u32(tkw_ring);
    if (threadIdx.x == 0) {
        for (int q = 0; q < TKW_STAGES; ++q) mbar_init(bar0 + 8u * q, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    int s = topk_seg_of<0>(segs, nseg, t_begin);
    TopkSeg sgm = segs[s];
    // 32-bit per-layer cursor (see k_topk_filter): tile inside the layer, tiles of the layer, whole tiles, ring-eligible
    auto layer_bulk = [&](const TopkSeg& g) {
        return ((((uintptr_t)(x + g.begin)) | (res_in ? (uintptr_t)(res_in + g.begin) : 0) | (res_out ? (uintptr_t)(res_out + g.begin) : 0)) & 15u) == 0;
    };
    uint32_t lt = (uint32_t)(t_begin - sgm.tile0), ntl = (sgm.n + TK_TILE - 1u) / TK_TILE, nfull = sgm.n / TK_TILE;
    bool lbulk = layer_bulk(sgm);
    const uint32_t nrun = (uint32_t)(t_end - t_begin);
    int s_is = s;                                                   // thread 0: cursor of the tile being issued
    uint32_t lt_is = lt, ntl_is = ntl, nfull_is = nfull, q_is = 0u;
    bool lbulk_is = lbulk;
    uint64_t e_is = sgm.begin + (uint64_t)lt * TK_TILE;
    const uint32_t stage_bytes = 2u * TK_TILE * 4u;
    auto issue_next = [&]() {
        if (lbulk_is && lt_is < nfull_is) {
            mbar_expect_tx(bar0 + 8u * q_is, (res_in ? 2u : 1u) * TK_TILE * 4u);
            bulk_g2s(ring0 + q_is * stage_bytes, x + e_is, TK_TILE * 4u, bar0 + 8u * q_is);
            if (res_in) bulk_g2s(ring0 + q_is * stage_bytes + TK_TILE * 4u, res_in + e_is, TK_TILE * 4u, bar0 + 8u * q_is);
        }
        q_is = q_is + 1u == TKW_STAGES ? 0u : q_is + 1u;
        if (++lt_is == ntl_is) {
            if (++s_is < nseg) {
                const TopkSeg g = segs[s_is];
                lt_is = 0u; ntl_is = (g.n + TK_TILE - 1u) / TK_TILE; nfull_is = g.n / TK_TILE; lbulk_is = layer_bulk(g); e_is = g.begin;
            }
        } else e_is += TK_TILE;
    };
    __syncthreads();                                               // barriers initialised
    if (threadIdx.x == 0)
        for (uint32_t j = 0; j < nrun && j < TKW_STAGES - 1u; ++j) issue_next();
    uint32_t phase = 0u;                                           // one parity bit per stage
    // (key > T, key == T) elements of the layer before the current tile: the per-tile counts of the tiles between the
    // layer's first one and the run's first one, summed once per run by the whole block (no scan pass over the table);
    // from there the block keeps the count itself, and a run enters every further layer at its first tile (0, 0)
    uint32_t run_g = 0u, run_e = 0u;
    {
        uint32_t g = 0u, e = 0u;
        for (uint64_t t = sgm.tile0 + threadIdx.x; t < t_begin; t += TK_THREADS) { const uint2 c = tile_counts[t]; g += c.x; e += c.y; }
        for (int d = 16; d > 0; d >>= 1) { g += __shfl_down_sync(0xffffffffu, g, d); e += __shfl_down_sync(0xffffffffu, e, d); }
        if (lane == 0) { wg[0][warp] = g; we[0][warp] = e; }
        __syncthreads();
#pragma unroll
        for (uint32_t v = 0; v < TK_THREADS / 32; ++v) { run_g += (uint32_t)wg[0][v]; run_e += (uint32_t)we[0][v]; }
        __syncthreads();
    }
    TopkState st = st_all[s];
    uint32_t q = 0u;
    for (uint32_t j = 0; j < nrun; ++j) {
        const bool more = j + (TKW_STAGES - 1u) < nrun;
        const bool bulk = lbulk && lt < nfull;
        const uint32_t* ring_x = tkw_ring + q * (2u * TK_TILE);
        const float* ring_r = reinterpret_cast<const float*>(ring_x + TK_TILE);
        TopkTile cur;
        if (bulk) {
            mbar_wait(bar0 + 8u * q, (phase >> q) & 1u);
            phase ^= 1u << q;
            const uint4* xs4 = reinterpret_cast<const uint4*>(ring_x);
            const float4* rs4 = reinterpret_cast<const float4*>(ring_r);
#pragma unroll
            for (int r = 0; r < TK_PER / 4; ++r) {
                const uint4 v = xs4[r * TK_THREADS + threadIdx.x];
                cur.w[4 * r] = v.x; cur.w[4 * r + 1] = v.y; cur.w[4 * r + 2] = v.z; cur.w[4 * r + 3] = v.w;
                if (res_in) { const float4 u = rs4[r * TK_THREADS + threadIdx.x]; cur.rv[4 * r] = u.x; cur.rv[4 * r + 1] = u.y; cur.rv[4 * r + 2] = u.z; cur.rv[4 * r + 3] = u.w; }
            }
            cur.valid = 0xffffu; cur.vec = true;
        } else {
            topk_write_load(cur, x, res_in, res_out, sgm, sgm.tile0 + lt);
        }
        const uint32_t T = st.prefix, skip_eq = st.c_eq - st.k_rem;   // ties with rank < skip_eq are not taken
        const uint32_t n = sgm.n;
        const uint32_t base = lt * TK_TILE;
        float f[TK_PER];
        uint32_t gm = 0u, em = 0u;                              // hit masks: key > T, key == T
#pragma unroll
        for (int k = 0; k < TK_PER; ++k) {
            const uint32_t key = key_of_bits(cur.w[k]);
            gm |= (uint32_t)(key > T) << k; em |= (uint32_t)(key == T) << k;
            f[k] = res_in ? __fadd_rn(__uint_as_float(cur.w[k]), cur.rv[k]) : __uint_as_float(cur.w[k]);
        }
        gm &= cur.valid; em &= cur.valid;
        uint64_t G = 0ull, E = 0ull;
#pragma unroll
        for (int r = 0; r < TK_PER / 4; ++r) {
            G |= (uint64_t)__popc(gm & (0xfu << (4 * r))) << (16 * r);
            E |= (uint64_t)__popc(em & (0xfu << (4 * r))) << (16 * r);
        }
        // block scan of the packed counts in thread order (a row holds at most 1024 elements: 16 bits suffice)
        uint64_t sg = G, se = E;
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t og = __shfl_up_sync(0xffffffffu, sg, d), oe = __shfl_up_sync(0xffffffffu, se, d);
            if (lane >= (uint32_t)d) { sg += og; se += oe; }
        }
        const uint32_t pb = j & 1u;                             // alternating slots: one barrier per tile
        if (lane == 31) { wg[pb][warp] = sg; we[pb][warp] = se; }
        // the rows of the new residual (selected positions are zeroed below)
        if (res_out) {
#pragma unroll
            for (int r = 0; r < TK_PER / 4; ++r) {
                const uint32_t i0 = base + 4u * ((uint32_t)r * TK_THREADS + threadIdx.x);
                float* op = res_out + sgm.begin + i0;
                if (cur.vec && i0 + 4u <= n) __stcs(reinterpret_cast<float4*>(op), make_float4(f[4 * r], f[4 * r + 1], f[4 * r + 2], f[4 * r + 3]));
                else for (int k = 0; k < 4; ++k) if (i0 + k < n) op[k] = f[4 * r + k];
            }
        }
        __syncthreads();                                        // ... which also ends every read of tile j - 1's stage
        if (threadIdx.x == 0 && more) issue_next();
        uint32_t hits = gm | em;
        uint64_t bg = 0ull, be = 0ull, tg = 0ull, te = 0ull;    // threads before this warp; the whole tile
#pragma unroll
        for (uint32_t v = 0; v < TK_THREADS / 32; ++v) { const uint64_t a = wg[pb][v], b = we[pb][v]; if (v < warp) { bg += a; be += b; } tg += a; te += b; }
        if (hits) {
            // per row: hits before this thread's quad = rows before (totals) + threads before in the row; rows packed 16 bits each
            const uint64_t rows_g = (tg << 16) + (tg << 32) + (tg << 48), rows_e = (te << 16) + (te << 32) + (te << 48);   // exclusive prefix over rows
            const uint64_t xg = bg + sg - G + rows_g, xe = be + se - E + rows_e;
            float* values = values_all + sgm.out_off;
            int64_t* index = index_all + sgm.out_off;
            while (hits) {
                const uint32_t k = (uint32_t)__ffs((int)hits) - 1u;
                hits &= hits - 1u;
                const uint32_t r = k >> 2, below = (1u << k) - 1u, rowm = 0xfu << (4u * r);
                const uint32_t gt_before = run_g + (uint32_t)((xg >> (16u * r)) & 0xffffu) + __popc(gm & rowm & below);
                const uint32_t eq_before = run_e + (uint32_t)((xe >> (16u * r)) & 0xffffu) + __popc(em & rowm & below);
                const bool is_gt = (gm >> k) & 1u;
                if (is_gt || eq_before >= skip_eq) {
                    const uint32_t taken_eq = eq_before > skip_eq ? eq_before - skip_eq : 0u;   // selected ties before it
                    const uint64_t pos = (uint64_t)gt_before + taken_eq;
                    const uint32_t el = topk_elem_of(k);
                    const uint64_t gi = sgm.begin + base + el;
                    // the value: still in the ring stage for a bulk tile (same operands, same rounding), else picked from the registers
                    values[pos] = bulk ? (res_in ? __fadd_rn(__uint_as_float(ring_x[el]), ring_r[el]) : __uint_as_float(ring_x[el])) : topk_pick16(f, k);
                    index[pos] = (int64_t)gi;
                    if (res_out) res_out[gi] = 0.0f;
                }
            }
        }
        q = q + 1u == TKW_STAGES ? 0u : q + 1u;
        if (++lt == ntl && j + 1u < nrun) {                     // the run enters the next layer, at its first tile
            ++s;
            sgm = segs[s]; st = st_all[s];
            lt = 0u; ntl = (sgm.n + TK_TILE - 1u) / TK_TILE; nfull = sgm.n / TK_TILE; lbulk = layer_bulk(sgm);
            run_g = 0u; run_e = 0u;
        } else {
            run_g += (uint32_t)((tg & 0xffffu) + ((tg >> 16) & 0xffffu) + ((tg >> 32) & 0xffffu) + (tg >> 48));
            run_e += (uint32_t)((te & 0xffffu) + ((te >> 16) & 0xffffu) + ((te >> 32) & 0xffffu) + (te >> 48));
        }
    }
}

// Launch with programmatic stream serialization (the kernels call topk_pdl_enter first): the launch latency of the
// ~20 small kernels of the chain overlaps the tail of their predecessors.  FLASHE_PDL=0 turns it off.
template <typename... KArgs, typename... Args>
static cudaError_t topk_launch_s(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t cs, Args... args) {
    static const bool pdl = [] { const char* e = getenv("FLASHE_PDL"); return !(e && e[0] == '0'); }();
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = cs;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1u : 0u;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
static cudaError_t topk_launch(void (*kern)(KArgs...), int grid, int block, cudaStream_t cs, Args... args) {
    return topk_launch_s(kern, grid, block, 0, cs, args...);
}

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int flashe_wire_nbytes(int bits, uint64_t count, uint64_t* nbytes_out) {
    if (!nbytes_out) return flashe_fail(FLASHE_EINVAL, "nbytes_out is NULL");
    if (bits < 1 || bits > 128) return flashe_fail(FLASHE_EINVAL, "bits must be in [1, 128]");
    if (count > (UINT64_MAX - 7) / (uint64_t)bits) return flashe_fail(FLASHE_EINVAL, "count * bits overflows");
    *nbytes_out = (count * (uint64_t)bits + 7) / 8;
    return FLASHE_OK;
}

static int wire_check(int bits, int word_bytes) {
    if (bits < 1 || bits > 128) return flashe_fail(FLASHE_EINVAL, "bits must be in [1, 128]");
    if (word_bytes != 4 && word_bytes != 8 && word_bytes != 16) return flashe_fail(FLASHE_EINVAL, "word_bytes must be 4, 8 or 16");
    if (bits > 8 * word_bytes) return flashe_fail(FLASHE_EINVAL, "bits exceeds the word width");
    return FLASHE_OK;
}

int flashe_wire_pack(flashe_ctx* ctx, const void* words, int word_bytes, uint64_t count, int bits, uint8_t* out, void* stream) {
    WIRE_ENTER(ctx);
    int rc = wire_check(bits, word_bytes); if (rc) return rc;
    if (count == 0) return FLASHE_OK;
    if (!words || !out) return flashe_fail(FLASHE_EINVAL, "NULL buffer");
    if (((uintptr_t)out & 15u) != 0) return flashe_fail(FLASHE_EINVAL, "out must be 16-byte aligned");
    uint64_t nbytes; rc = flashe_wire_nbytes(bits, count, &nbytes); if (rc) return rc;
    if (word_bytes == 4 && bits >= 8) {
        // whole 16-byte chunks through the tiled kernel, a short final chunk through the generic one
        const uint64_t nvec = nbytes >> 4;
        const uint32_t pad = (uint32_t)(8 * nbytes - count * (uint64_t)bits);
        int launches = 0;
        if (nvec && ((uintptr_t)words & 15u) == 0) {
            // bulk-staged form: stage = the elements of one tile (+ alignment slack), two stages
            const uint32_t stage_words = (uint32_t)(((WB_WORDS * 32 + bits - 1) / bits + 2 + 8 + 31) & ~31);
            const size_t smem = 2u * (size_t)stage_words * 4u;
            const uint32_t magic = (uint32_t)((1ull << 32) / (uint64_t)bits) + 1u;      // floor(x / bits) = umulhi(x, magic) for x < 2^17
            const uint64_t nwords = nvec * 4;
            // the shipped widths (20: un-batched ciphertexts, 32: the wide configs, 24) have their own instantiation
            auto kern = bits == 20 ? k_wire_pack32_bulk<20> : (bits == 32 ? k_wire_pack32_bulk<32> : (bits == 24 ? k_wire_pack32_bulk<24> : k_wire_pack32_bulk<0>));
            FLASHE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4 * ((WB_WORDS * 32 / 8) + 64)));
            const int grid = grid_occ(info.num_sms, info.device, (const void*)kern, ceil_div_u64(nwords, WB_WORDS) * WB_THREADS, WB_THREADS, smem);
            kern<<<grid, WB_THREADS, smem, cs>>>(reinterpret_cast<const uint32_t*>(words), count, (uint32_t)bits, pad, nwords, magic,
                                                 stage_words, reinterpret_cast<uint32_t*>(out));
            ++launches;
        } else if (nvec) {
            const int grid = grid_cap(info.num_sms, ceil_div_u64(nvec, WP_THREADS), 8);
            k_wire_pack32<<<grid, WP_THREADS, 0, cs>>>(reinterpret_cast<const uint32_t*>(words), count, (uint32_t)bits, pad, nvec,
                                                        reinterpret_cast<uint4*>(out));
            ++launches;
        }
        if (nbytes & 15u) { k_wire_pack<4><<<1, 32, 0, cs>>>(words, count, (uint32_t)bits, nbytes, out, nvec); ++launches; }
        flashe_count_launches(launches);
        FLASHE_CUDA_TRY(cudaGetLastError());
        return FLASHE_OK;
    }
    const int grid = grid_cap(info.num_sms, ceil_div_u64((nbytes + 15) / 16, 256), 8);
    if (word_bytes == 4) k_wire_pack<4><<<grid, 256, 0, cs>>>(words, count, (uint32_t)bits, nbytes, out, 0);
    else if (word_bytes == 8) k_wire_pack<8><<<grid, 256, 0, cs>>>(words, count, (uint32_t)bits, nbytes, out, 0);
    else k_wire_pack<16><<<grid, 256, 0, cs>>>(words, count, (uint32_t)bits, nbytes, out, 0);
    flashe_count_launches(1);
    FLASHE_CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}

int flashe_wire_unpack(flashe_ctx* ctx, const uint8_t* in, uint64_t count, int bits, int word_bytes, void* words_out, void* stream) {
    WIRE_ENTER(ctx);
    int rc = wire_check(bits, word_bytes); if (rc) return rc;
    if (count == 0) return FLASHE_OK;
    if (!in || !words_out) return flashe_fail(FLASHE_EINVAL, "NULL buffer");
    uint64_t nbytes; rc = flashe_wire_nbytes(bits, count, &nbytes); if (rc) return rc;
    if (word_bytes == 4 && ((uintptr_t)words_out & 3u) == 0 && ((uintptr_t)in & 15u) == 0) {
        const uint32_t pad = (uint32_t)(8 * nbytes - count * (uint64_t)bits);
        if (bits >= 8) {
            const uint32_t stage_words = (uint32_t)(((WUB_ELEMS * (uint32_t)bits / 8 + 16 + 4 + 16) / 4 + 31) & ~31);
            const size_t smem = 2u * (size_t)stage_words * 4u;
            static bool attr_done = false;
            if (!attr_done) { FLASHE_CUDA_TRY(cudaFuncSetAttribute(k_wire_unpack32_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (WUB_ELEMS * 4 + 256))); attr_done = true; }
            const int grid = grid_occ(info.num_sms, info.device, (const void*)k_wire_unpack32_bulk, ceil_div_u64(count, WUB_ELEMS) * WB_THREADS, WB_THREADS, smem);
            k_wire_unpack32_bulk<<<grid, WB_THREADS, smem, cs>>>(in, count, (uint32_t)bits, pad, nbytes, stage_words, reinterpret_cast<uint32_t*>(words_out));
        } else {
        const int grid = grid_cap(info.num_sms, ceil_div_u64(count, WU_ELEMS), 8);
        k_wire_unpack32<<<grid, WP_THREADS, 0, cs>>>(in, count, (uint32_t)bits, pad, nbytes, reinterpret_cast<uint32_t*>(words_out));
        }
        flashe_count_launches(1);
        FLASHE_CUDA_TRY(cudaGetLastError());
        return FLASHE_OK;
    }
    const int grid = grid_cap(info.num_sms, ceil_div_u64(count, 256), 8);
    if (word_bytes == 4) k_wire_unpack<4><<<grid, 256, 0, cs>>>(in, count, (uint32_t)bits, nbytes, words_out);
    else if (word_bytes == 8) k_wire_unpack<8><<<grid, 256, 0, cs>>>(in, count, (uint32_t)bits, nbytes, words_out);
    else k_wire_unpack<16><<<grid, 256, 0, cs>>>(in, count, (uint32_t)bits, nbytes, words_out);
    flashe_count_launches(1);
    FLASHE_CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}

int flashe_topk_sparsify(flashe_ctx* ctx, const float* x, const float* residual_in, uint64_t total, const uint64_t* seg_end,
                         const uint64_t* k, int nseg, float* values_out, int64_t* index_out, float* residual_out, void* stream) {
    WIRE_ENTER(ctx);
    if (nseg < 1 || !seg_end || !k) return flashe_fail(FLASHE_EINVAL, "need nseg >= 1, seg_end and k");
    if (seg_end[nseg - 1] != total) return flashe_fail(FLASHE_EINVAL, "seg_end[nseg-1] must equal total");
    uint64_t prev = 0, max_n = 0;
    for (int s = 0; s < nseg; ++s) {
        if (seg_end[s] < prev) return flashe_fail(FLASHE_EINVAL, "seg_end must be ascending");
        const uint64_t n = seg_end[s] - prev;
        if (k[s] > n) return flashe_fail(FLASHE_EINVAL, "k[s] exceeds the layer size");
        if (n && k[s] == 0) return flashe_fail(FLASHE_EINVAL, "k[s] must be >= 1 for a non-empty layer (max(1, floor(s*size)))");
        if (n >= (1ull << 32)) return flashe_fail(FLASHE_EUNSUPPORTED, "layers of 2^32 or more elements");
        if (n > max_n) max_n = n;
        prev = seg_end[s];
    }
    if (total == 0) return FLASHE_OK;
    if (!x || !values_out || !index_out) return flashe_fail(FLASHE_EINVAL, "NULL buffer");
    // FLASHE_TOPK_ROUTE (tests): "exact" = no layer samples; "badbound" = the sample rank is forced to 1, so the bound
    // lands above the threshold and the device has to fall back to the exact route.
    const char* route_env = getenv("FLASHE_TOPK_ROUTE");
    const bool force_exact = route_env && strcmp(route_env, "exact") == 0;
    const bool bad_bound = route_env && strcmp(route_env, "badbound") == 0;
    // Layers are processed in groups of at most TK_GROUP by the same launches (the per-layer
    // histograms of a group are 8 KB each).  Empty layers are dropped from the table.
    const int TK_GROUP = 4096;
    static const bool small_on = [] { const char* ev = getenv("FLASHE_TOPK_SMALL"); return !(ev && ev[0] == '0'); }();   // (A/B switch)
    std::vector<TopkSeg> segs;
    segs.reserve((size_t)(nseg < TK_GROUP ? nseg : TK_GROUP));
    uint8_t* ws = nullptr;
    cudaError_t e = cudaSuccess;
    int launches = 0;
    uint64_t begin = 0, out_off = 0;
    int s = 0;
    const uint32_t* xw = reinterpret_cast<const uint32_t*>(x);
    while (e == cudaSuccess && s < nseg) {
        segs.clear();
        uint64_t ntiles = 0, nctiles = 0, nlines = 0;
        int n_small = 0;
        for (; s < nseg && (int)segs.size() < TK_GROUP; ++s) {
            const uint64_t n = seg_end[s] - begin;
            if (n) {
                TopkSeg g;
                memset(&g, 0, sizeof(g));
                g.begin = begin; g.tile0 = ntiles; g.out_off = out_off; g.n = (uint32_t)n; g.k = (uint32_t)k[s];
                g.cbegin = nctiles * TK_TILE; g.ctile0 = nctiles; g.samp0 = nlines;
                // candidate route: sparse selections of layers large enough to sample
                if (!force_exact && n >= 4096 && k[s] * 40 <= n) {
                    uint32_t lines, stride, rank;
                    if (n <= TK_SAMPLE_ALL) { lines = (uint32_t)((n + 31) / 32); stride = 1; rank = (uint32_t)k[s]; }
                    else {
                        const uint64_t lines_total = n / 32;
                        uint64_t want = lines_total / 64;
                        want = want < 128 ? 128 : (want > 512 ? 512 : want);
                        lines = (uint32_t)want; stride = (uint32_t)(lines_total / want);
                        const double sp = 32.0 * (double)lines * (double)k[s] / (double)n;     // expected sample elements above the threshold
                        rank = (uint32_t)(1.5 * sp + 6.0 * sqrt(sp) + 8.0) + 1u;
                        if (bad_bound) rank = 1u;
                        if (rank >= 32u * lines) lines = 0;
                    }
                    if (lines) {
                        g.cap = (uint32_t)(((n / 8) + 3) & ~3ull);
                        g.samp_lines = lines; g.samp_stride = stride; g.samp_rank = rank;
                        nctiles += ceil_div_u64(g.cap, TK_TILE);
                        nlines += lines;
                    }
                }
                // small: the radix passes of the layer fit one block (candidate lists of up to TK_SMALL_N / 8 keys, or a short x)
                // (larger layers stay on the multi-block passes: a thread-block cluster of 8 per layer, histograms added up through
                //  distributed shared memory, was measured at 0.33 ms against 0.197 for 1 % of 50 M - the candidates of a layer
                //  share a handful of top-digit bins, and the shared-memory atomics on them need all the SMs, not eight)
                g.small = (small_on && ((g.cap != 0u && n <= TK_SMALL_N) || n <= TK_SMALL_EXACT)) ? 1u : 0u;
                if (g.small) ++n_small;
                segs.push_back(g);
                ntiles += ceil_div_u64(n, TK_TILE);
            }
            begin = seg_end[s];
            out_off += k[s];
        }
        const int ng = (int)segs.size();
        if (ng == 0) continue;
        // workspace: layer table | states | [histograms | candidate counts | bounds | tile counts] (zeroed) | candidate keys | candidate indices
        const size_t seg_bytes = (sizeof(TopkSeg) * (size_t)ng + 255) & ~(size_t)255;
        const size_t st_bytes = (sizeof(TopkState) * (size_t)ng + 255) & ~(size_t)255;
        const size_t hist_bytes = sizeof(uint32_t) * TK_BINS * (size_t)ng;
        const size_t cnt_bytes = (sizeof(uint32_t) * (size_t)ng + 255) & ~(size_t)255;
        const size_t tile_bytes = (sizeof(uint2) * (size_t)ntiles + 255) & ~(size_t)255;
        const size_t zero_bytes = hist_bytes + 2 * cnt_bytes + tile_bytes;
        const size_t cand_bytes = sizeof(uint32_t) * (size_t)nctiles * TK_TILE;
        e = cudaMallocAsync((void**)&ws, seg_bytes + st_bytes + zero_bytes + 2 * cand_bytes, cs);
        if (e != cudaSuccess) break;
        TopkSeg* dseg = reinterpret_cast<TopkSeg*>(ws);
        TopkState* st = reinterpret_cast<TopkState*>(ws + seg_bytes);
        uint8_t* zero0 = ws + seg_bytes + st_bytes;
        uint32_t* hist = reinterpret_cast<uint32_t*>(zero0);
        uint32_t* cand_count = reinterpret_cast<uint32_t*>(zero0 + hist_bytes);
        uint32_t* lo = reinterpret_cast<uint32_t*>(zero0 + hist_bytes + cnt_bytes);
        uint2* tiles = reinterpret_cast<uint2*>(zero0 + hist_bytes + 2 * cnt_bytes);
        uint32_t* cand_key = reinterpret_cast<uint32_t*>(zero0 + zero_bytes);
        uint32_t* cand_idx = reinterpret_cast<uint32_t*>(zero0 + zero_bytes + cand_bytes);
        e = cudaMemcpyAsync(dseg, segs.data(), sizeof(TopkSeg) * (size_t)ng, cudaMemcpyHostToDevice, cs);
        // (a pageable source has been staged by the time cudaMemcpyAsync returns: `segs` may be refilled)
        if (e == cudaSuccess) e = cudaMemsetAsync(zero0, 0, zero_bytes, cs);
        const int gh = grid_cap(info.num_sms, ntiles, 8);
        const uint64_t tpb = ceil_div_u64(ntiles, (uint64_t)gh);
        const int gc = grid_cap(info.num_sms, nctiles, 8);
        const uint64_t ctpb = nctiles ? ceil_div_u64(nctiles, (uint64_t)gc) : 0;
        uint32_t* null_lo = nullptr;
#define TK_GO(call) do { if (e == cudaSuccess) { e = (call); ++launches; } } while (0)
        if (n_small != ng) TK_GO(topk_launch(k_topk_init, (ng + 255) / 256, 256, cs, dseg, ng, st));   // (k_topk_select_small writes the whole state of a small layer)
        if (nctiles) {
            if (small_on && ng >= 8) TK_GO(topk_launch(k_topk_bound, ng, 1024, cs, xw, dseg, lo));   // (a lone block is slower than 64 sampling blocks + the pick: 50 M single layer 0.197 -> 0.202 ms)
            else {
                TK_GO(topk_launch(k_topk_sample, grid_cap(info.num_sms, ceil_div_u64(nlines, 8), 8), 256, cs, xw, dseg, ng, nlines, hist));
                TK_GO(topk_launch(k_topk_pick, ng, 1024, cs, hist, st, 0u, dseg, lo));
            }
            const size_t f_smem = (size_t)TKF_STAGES * TK_TILE * 4u;
            { static bool attr_done = false; if (!attr_done && e == cudaSuccess) { e = cudaFuncSetAttribute(k_topk_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f_smem); attr_done = true; } }
            const int gf = grid_occ(info.num_sms, info.device, (const void*)k_topk_filter, ntiles * TK_THREADS, TK_THREADS, f_smem);   // exactly the resident CTAs: equal runs, one wave
            TK_GO(topk_launch_s(k_topk_filter, gf, TK_THREADS, f_smem, cs, xw, dseg, ng, ntiles, ceil_div_u64(ntiles, (uint64_t)gf), lo, cand_count, cand_key, cand_idx));
        }
        const uint32_t shifts[3] = {20u, 9u, 0u}, nbits[3] = {11u, 11u, 9u};
        const uint32_t gcu = nctiles ? (uint32_t)gc : 0u;
        if (n_small > 0) TK_GO(topk_launch(k_topk_select_small, ng, 1024, cs, xw, cand_key, cand_idx, dseg, cand_count, st, tiles));
        for (int p = 0; p < 3 && n_small != ng; ++p) {         // the multi-block passes serve the layers that are not small
            TK_GO(topk_launch(k_topk_hist, (int)gcu + gh, TK_THREADS, cs, xw, cand_key, dseg, ng, gcu, nctiles, ctpb, ntiles, tpb, st, cand_count, shifts[p], nbits[p], hist));
            TK_GO(topk_launch(k_topk_pick, ng, 1024, cs, hist, st, shifts[p], dseg, null_lo));
        }
        const uint32_t gcc = nctiles ? (uint32_t)grid_cap(info.num_sms, 4 * nctiles, 8) : 0u;
        if (n_small != ng) TK_GO(topk_launch(k_topk_count, (int)gcc + gh, TK_THREADS, cs, xw, cand_key, cand_idx, dseg, ng, gcc, nctiles, ntiles, tpb, st, cand_count, tiles));
        const size_t w_smem = (size_t)TKW_STAGES * 2u * TK_TILE * 4u;
        { static bool attr_done = false; if (!attr_done && e == cudaSuccess) { e = cudaFuncSetAttribute(k_topk_write, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w_smem); attr_done = true; } }
        const int gw = grid_occ(info.num_sms, info.device, (const void*)k_topk_write, ntiles * TK_THREADS, TK_THREADS, w_smem);
        TK_GO(topk_launch_s(k_topk_write, gw, TK_THREADS, w_smem, cs, xw, residual_in, dseg, ng, ntiles, ceil_div_u64(ntiles, (uint64_t)gw), st, tiles, values_out, index_out, residual_out));
#undef TK_GO
        if (e == cudaSuccess) e = cudaGetLastError();
        cudaFreeAsync(ws, cs);
        ws = nullptr;
    }
    flashe_count_launches(launches);
    if (e != cudaSuccess) return flashe_fail(FLASHE_ECUDA, std::string("flashe_topk_sparsify: ") + cudaGetErrorString(e));
    return FLASHE_OK;
}

}  // extern "C"
