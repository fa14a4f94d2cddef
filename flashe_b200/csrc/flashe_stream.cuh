// flashe_stream.cuh — the PRF stream kernel (k_stream) of the FLASHE hot path and its launcher.  Included by
// one small translation unit per mode (flashe_stream_*.cu) so that the ~35 instantiations compile in parallel;
// flashe_kernels.cu holds the host side and the C ABI.  See the header of flashe_kernels.cu.
#ifndef FLASHE_STREAM_CUH
#define FLASHE_STREAM_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <string>

#include "../../include/flashe_b200.h"
#include "flashe_internal.h"
#include "flashe_device.cuh"
#include "flashe_stream_decl.h"

// ------------------------------------------------------------------------------------------------
// device: shared-memory T-tables
// Layout (absolute addresses in the CTA's shared window):
//   [0x10000, 0x20000)  T0/T1 interleaved: entry e, table t, replica l at 0x10000 + e*256 + t*128 + l*4
//   [0x20000, 0x30000)  T2/T3 likewise
// Replica l is only ever read by lane l, so every lookup instruction hits 32 distinct banks.  The
// address of a lookup is PRMT(state, y, sel) with y = 0x00010000 | lane*4: one ALU op builds
// 0x0001_<byte>_<lane*4>, the table select rides in the LDS immediate.
// Below 0x10000 (from wherever the driver starts dynamic shared memory) live the per-warp slabs.
// ------------------------------------------------------------------------------------------------
#define TAB_BASE 0x10000u


template <int OFF>
__device__ __forceinline__ uint32_t lds_tab(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t smem_window_base() {
    extern __shared__ __align__(16) uint8_t dyn_smem[];
    return (uint32_t)__cvta_generic_to_shared(dyn_smem);
}

__device__ __forceinline__ void fill_tables(const uint32_t* __restrict__ g_te0) {   // g_te0: Te0, 256 words (flashe_ctx::d_te0)
    // One warp store writes the 32 replicas (128 bytes) of one (entry, table) pair.  A warp owns 256 / nwarps consecutive
    // entries: ONE coalesced load brings them in (lane i holds entry e0 + i), each is broadcast by shuffle and stored
    // four times, rotated (Te_t = ror(Te0, 8 t)).  The per-word loop this replaces paid one dependent global load per
    // store: 3.3 us per launch on the per-warp timeline, never hidden - a k_stream CTA owns its SM's whole register
    // file, so a dependent launch cannot become resident next to its predecessor.
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t per = (256u + nwarps - 1u) / nwarps;          // entries per warp: 16 with 16 warps (<= 32 for >= 8 warps)
    for (uint32_t base = warp * per; base < 256u && base < (warp + 1u) * per; base += 32u) {
        const uint32_t n_here = min(min(32u, 256u - base), (warp + 1u) * per - base);
        const uint32_t mine = lane < n_here ? __ldg(g_te0 + base + lane) : 0u;
#pragma unroll 4
        for (uint32_t i = 0; i < n_here; ++i) {
            const uint32_t v = __shfl_sync(0xffffffffu, mine, i);
            const uint32_t a = TAB_BASE + (base + i) * 256u + lane * 4u;
            sts32(a, v);                                              // T0: region 0, first half of the entry's row
            sts32(a + 128u, __funnelshift_r(v, v, 8));                // T1
            sts32(a + 0x10000u, __funnelshift_r(v, v, 16));           // T2: region 1
            sts32(a + 0x10080u, __funnelshift_r(v, v, 24));           // T3
        }
    }
}

#define SEL_B3 0x7634
#define SEL_B2 0x7624
#define SEL_B1 0x7614
#define SEL_B0 0x7604
#ifndef FLASHE_IMAD_B3
#define FLASHE_IMAD_B3 0
#endif
// Address of the byte-3 lookup on the FMA pipe instead of the ALU pipe: (s >> 24) via mad.hi, then
// * 256 + y via mad.lo (the ALU pipe is as loaded as the LSU; the FMA pipe idles).
__device__ __forceinline__ uint32_t addr_b3(uint32_t s, uint32_t y) {
#if FLASHE_IMAD_B3
    uint32_t hi, a;
    asm("mul.hi.u32 %0, %1, 256;" : "=r"(hi) : "r"(s));
    asm("mad.lo.u32 %0, %1, 256, %2;" : "=r"(a) : "r"(hi), "r"(y));
    return a;
#else
    return __byte_perm(s, y, SEL_B3);
#endif
}
#ifndef FLASHE_IMAD_B0
#define FLASHE_IMAD_B0 0
#endif
// Same for the byte-0 lookup: (s << 24) via mul.lo, then ((s << 24) >> 16) + y via mad.hi.
__device__ __forceinline__ uint32_t addr_b0(uint32_t s, uint32_t y) {
#if FLASHE_IMAD_B0
    uint32_t t, a;
    asm("mul.lo.u32 %0, %1, 0x1000000;" : "=r"(t) : "r"(s));
    asm("mad.hi.u32 %0, %1, 0x10000, %2;" : "=r"(a) : "r"(t), "r"(y));
    return a;
#else
    return __byte_perm(s, y, SEL_B0);
#endif
}
#define T0(s) lds_tab<0>(addr_b3((s), y))
#define T1(s) lds_tab<128>(__byte_perm((s), y, SEL_B2))
#define T2(s) lds_tab<0x10000>(__byte_perm((s), y, SEL_B1))
#define T3(s) lds_tab<0x10080>(addr_b0((s), y))

// AES-256 of the block {w0,w1,w2,w3}; `pre` = round-1 terms hoisted by the host for (w0,w1,w2=0).
// Output o[0..3] big-endian words (o[0] most significant).
struct Pre { uint32_t p0, p1, p2, p3; };
__device__ __forceinline__ void aes256_block(const KeySched& ks, uint32_t y, uint32_t w0, uint32_t w1,
                                             uint32_t w2, uint32_t w3, Pre pre, uint32_t o[4]) {
    uint32_t s0, s1, s2, s3, t0, t1, t2, t3;
    s3 = w3 ^ ks.rk[3];
    if (w2 == 0) {
        t0 = pre.p0 ^ T3(s3);
        t1 = pre.p1 ^ T2(s3);
        t2 = pre.p2 ^ T1(s3);
        t3 = pre.p3 ^ T0(s3);
    } else {
        s0 = w0 ^ ks.rk[0]; s1 = w1 ^ ks.rk[1]; s2 = w2 ^ ks.rk[2];
        t0 = T0(s0) ^ T1(s1) ^ T2(s2) ^ T3(s3) ^ ks.rk[4];
        t1 = T0(s1) ^ T1(s2) ^ T2(s3) ^ T3(s0) ^ ks.rk[5];
        t2 = T0(s2) ^ T1(s3) ^ T2(s0) ^ T3(s1) ^ ks.rk[6];
        t3 = T0(s3) ^ T1(s0) ^ T2(s1) ^ T3(s2) ^ ks.rk[7];
    }
#pragma unroll
    for (int r = 2; r < 14; r += 2) {
        s0 = T0(t0) ^ T1(t1) ^ T2(t2) ^ T3(t3) ^ ks.rk[4 * r + 0];
        s1 = T0(t1) ^ T1(t2) ^ T2(t3) ^ T3(t0) ^ ks.rk[4 * r + 1];
        s2 = T0(t2) ^ T1(t3) ^ T2(t0) ^ T3(t1) ^ ks.rk[4 * r + 2];
        s3 = T0(t3) ^ T1(t0) ^ T2(t1) ^ T3(t2) ^ ks.rk[4 * r + 3];
        t0 = T0(s0) ^ T1(s1) ^ T2(s2) ^ T3(s3) ^ ks.rk[4 * r + 4];
        t1 = T0(s1) ^ T1(s2) ^ T2(s3) ^ T3(s0) ^ ks.rk[4 * r + 5];
        t2 = T0(s2) ^ T1(s3) ^ T2(s0) ^ T3(s1) ^ ks.rk[4 * r + 6];
        t3 = T0(s3) ^ T1(s0) ^ T2(s1) ^ T3(s2) ^ ks.rk[4 * r + 7];
    }
    // t = state after round 13.  Final round: SubBytes + ShiftRows + AddRoundKey; the S-box byte is
    // taken from the table whose entry carries S[x] in the wanted byte lane:
    //   byte3 <- T2 (S<<24), byte2 <- T3 (S<<16), byte1 <- T0 (S<<8), byte0 <- T1 (S).
#define LAST(a, b, c, d, k)                                                                         \
    (__byte_perm(__byte_perm(lds_tab<128>(__byte_perm((d), y, SEL_B0)),                             \
                             lds_tab<0>(__byte_perm((c), y, SEL_B1)), 0x3250),                      \
                 __byte_perm(lds_tab<0x10080>(__byte_perm((b), y, SEL_B2)),                         \
                             lds_tab<0x10000>(__byte_perm((a), y, SEL_B3)), 0x7210), 0x7610) ^ ks.rk[k])
    o[0] = LAST(t0, t1, t2, t3, 56);
    o[1] = LAST(t1, t2, t3, t0, 57);
    o[2] = LAST(t2, t3, t0, t1, 58);
    o[3] = LAST(t3, t0, t1, t2, 59);
#undef LAST
}

// ------------------------------------------------------------------------------------------------
// device: geometry of the reference's chunked counter rule (jzf_flashe.py:12-16, 24-34)
// ------------------------------------------------------------------------------------------------
// Items are cut on multiples of 64 of the AES COUNTER (counter = chunk begin + block, jzf_flashe.py:34),
// not of the block number: item 0 of a chunk holds its first 64 - (cb & 63) blocks, item w >= 1 the
// blocks [64w - (cb & 63), +64).  All counters of an item then share their upper 56 bits, which is
// what lets the first two AES rounds be factored per item (window_consts below).
struct Item { uint64_t cb; uint64_t clen; uint64_t w; uint32_t off; };  // chunk begin, chunk length, first item, cb & 63

// Work unit S -> first warp item of the unit and the number of items in it.  The two 64-bit divisions
// happen once per unit; the items inside are walked incrementally.
__device__ __forceinline__ Item decode_unit(const Geom& g, uint64_t S, uint32_t& nsub) {
    Item it;
    uint64_t s, nw;
    if (S < g.rSA) {
        const uint64_t k = S / g.nsA; s = S - k * g.nsA;
        it.cb = k * (g.d + 1); it.clen = g.d + 1; nw = g.nwA;
    } else {
        const uint64_t Sp = S - g.rSA;
        const uint64_t k = Sp / g.nsB; s = Sp - k * g.nsB;
        it.cb = g.r * (g.d + 1) + k * g.d; it.clen = g.d; nw = g.nwB;
    }
    const uint64_t w0 = s * g.sup, left = nw - w0;
    it.w = w0;
    it.off = (uint32_t)(it.cb & (ITEM_BLOCKS - 1u));
    nsub = (uint32_t)(left < g.sup ? left : g.sup);
    return it;
}

// Slots of one AES output (jzf_flashe.py:37-43): s = big-endian 128-bit integer; slot k is
// (s >> k*b) & mask.  acc[k] += sign * slot.
template <int WORDS, int MMAX>
__device__ __forceinline__ void accumulate_slots(const uint32_t o[4], uint32_t b, uint32_t m, int sign,
                                                 typename Word<WORDS>::T (&acc)[MMAX]) {
    if constexpr (WORDS == 1) {
        if (b == 32u) {          // four whole words: no shifting (slot k = word 3-k)
#pragma unroll
            for (int k = 0; k < 4 && k < MMAX; ++k) acc[k] += (uint32_t)sign * o[3 - k];
            return;
        }
        // The accumulators are only meaningful mod 2^b (every consumer masks the final sum), so the bits a
        // slot word carries above bit b need not be cleared here.
        uint32_t v0 = o[3], v1 = o[2], v2 = o[1], v3 = o[0];
        const uint32_t sg = (uint32_t)sign;
        if (MMAX >= 6 && b == 20u) {             // the shipped un-batched width: slots at bits 0, 20, .. 100
            acc[0] += sg * v0;
            acc[1] += sg * __funnelshift_r(v0, v1, 20);
            acc[2] += sg * (v1 >> 8);
            acc[3] += sg * __funnelshift_r(v1, v2, 28);
            acc[MMAX >= 6 ? 4 : 0] += sg * __funnelshift_r(v2, v3, 16);
            acc[MMAX >= 6 ? 5 : 0] += sg * (v3 >> 4);           // (index guarded for the narrower instantiations)
            return;
        }
        if (MMAX >= 5 && b == 24u) {             // slots at bits 0, 24, 48, 72, 96
            acc[0] += sg * v0;
            acc[1] += sg * __funnelshift_r(v0, v1, 24);
            acc[2] += sg * __funnelshift_r(v1, v2, 16);
            acc[3] += sg * __funnelshift_r(v2, v3, 8);
            acc[MMAX >= 5 ? 4 : 0] += sg * v3;
            return;
        }
#pragma unroll
        for (int k = 0; k < MMAX; ++k) {
            if ((uint32_t)k < m) {
                acc[k] += sg * v0;
                v0 = __funnelshift_rc(v0, v1, b);
                v1 = __funnelshift_rc(v1, v2, b);
                v2 = __funnelshift_rc(v2, v3, b);
                v3 = __funnelshift_rc(v3, 0u, b);
            }
        }
    } else if constexpr (WORDS == 2) {
        uint64_t V0 = ((uint64_t)o[2] << 32) | o[3], V1 = ((uint64_t)o[0] << 32) | o[1];
        const uint64_t mk = Word<2>::mask(b);
#pragma unroll
        for (int k = 0; k < MMAX; ++k) {
            if ((uint32_t)k < m) {
                const uint64_t slot = V0 & mk;
                acc[k] = sign >= 0 ? acc[k] + slot : acc[k] - slot;
                if (b >= 64) { V0 = V1; V1 = 0; }
                else { V0 = (V0 >> b) | (V1 << (64 - b)); V1 >>= b; }
            }
        }
    } else {
        u128 s; s.lo = ((uint64_t)o[2] << 32) | o[3]; s.hi = ((uint64_t)o[0] << 32) | o[1];
        s = Word<4>::band(s, Word<4>::mask(b));
        acc[0] = sign >= 0 ? Word<4>::add(acc[0], s) : Word<4>::sub(acc[0], s);
    }
}

// ------------------------------------------------------------------------------------------------
// the stream kernel
// ------------------------------------------------------------------------------------------------
template <int WORDS>
__device__ __forceinline__ void slab_store(uint32_t addr, typename Word<WORDS>::T v);
template <> __device__ __forceinline__ void slab_store<1>(uint32_t addr, uint32_t v) { sts32(addr, v); }
template <> __device__ __forceinline__ void slab_store<2>(uint32_t addr, uint64_t v) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"((uint32_t)v), "r"((uint32_t)(v >> 32)) : "memory");
}
template <> __device__ __forceinline__ void slab_store<4>(uint32_t addr, u128 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"((uint32_t)v.lo), "r"((uint32_t)(v.lo >> 32)),
                 "r"((uint32_t)v.hi), "r"((uint32_t)(v.hi >> 32)) : "memory");
}
template <int WORDS>
__device__ __forceinline__ typename Word<WORDS>::T slab_load(uint32_t addr);
template <> __device__ __forceinline__ uint32_t slab_load<1>(uint32_t addr) { return lds32(addr); }
template <> __device__ __forceinline__ uint64_t slab_load<2>(uint32_t addr) {
    uint32_t a, b;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr) : "memory");
    return ((uint64_t)b << 32) | a;
}
template <> __device__ __forceinline__ u128 slab_load<4>(uint32_t addr) {
    uint32_t a, b, c, d;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
    u128 r; r.lo = ((uint64_t)b << 32) | a; r.hi = ((uint64_t)d << 32) | c; return r;
}

// Four consecutive 4-byte elements whose first element sits `r` elements past a 16-byte boundary
// (r is warp-uniform: it is a property of the reference chunk the item belongs to).  r = 0: one
// 128-bit access; r = 2: two 64-bit accesses; r odd: 32 + 64 + 32 bits.  The narrower loads allocate
// in L1 (the lane's accesses share sectors), the stores merge in L2.
__device__ __forceinline__ void ldg_quad(const void* p, uint32_t r, uint32_t (&v)[4]) {
    const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
    if (r == 0u) {
        ldg_v4(q, v[0], v[1], v[2], v[3]);
    } else if (r == 2u) {
        asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "l"(q));
        asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(v[2]), "=r"(v[3]) : "l"(q + 2));
    } else {
        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v[0]) : "l"(q));
        asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(v[1]), "=r"(v[2]) : "l"(q + 1));
        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v[3]) : "l"(q + 3));
    }
}
__device__ __forceinline__ void stg_quad(void* p, uint32_t r, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    uint32_t* q = reinterpret_cast<uint32_t*>(p);
    if (r == 0u) {
        stg_v4(q, a, b, c, d);
    } else if (r == 2u) {
        asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(q), "r"(a), "r"(b) : "memory");
        asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(q + 2), "r"(c), "r"(d) : "memory");
    } else {
        asm volatile("st.global.u32 [%0], %1;" ::"l"(q), "r"(a) : "memory");
        asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(q + 1), "r"(b), "r"(c) : "memory");
        asm volatile("st.global.u32 [%0], %1;" ::"l"(q + 3), "r"(d) : "memory");
    }
}
// four consecutive float64 outputs, first one `r` elements past a 32-byte boundary of the element grid
__device__ __forceinline__ void stg_quad_f64(double* p, uint32_t r, const double (&v)[4]) {
    if ((reinterpret_cast<uintptr_t>(p) & 31u) == 0u) {        // (warp-uniform) whole 32-byte sectors in one store
        stg_d4(p, v[0], v[1], v[2], v[3]);
    } else if ((r & 1u) == 0u) {
        stg_d2(p, v[0], v[1]);
        stg_d2(p + 2, v[2], v[3]);
    } else {
        asm volatile("st.global.f64 [%0], %1;" ::"l"(p), "d"(v[0]) : "memory");
        stg_d2(p + 1, v[1], v[2]);
        asm volatile("st.global.f64 [%0], %1;" ::"l"(p + 3), "d"(v[3]) : "memory");
    }
}

// Out-of-line single block for the rare paths (chunk tails, counters >= 2^32).
static __device__ __noinline__ void aes256_block_slow(const KeySched& ks, uint32_t y, uint32_t w0, uint32_t w1, uint32_t w2,
                                               uint32_t w3, Pre pre, uint32_t* o) {
    uint32_t t[4];
    aes256_block(ks, y, w0, w1, w2, w3, pre, t);
    o[0] = t[0]; o[1] = t[1]; o[2] = t[2]; o[3] = t[3];
}

// Unroll factor of the double-round loop of aes256_x2w (5 iterations).  Fully unrolled (5) is the measured
// optimum now that the hot loop holds ONE inlined copy (~13 KB of SASS; round keys become constant-bank
// operands): 77.2 ms vs 78.7 ms rolled for the 64-client encode.  With several inlined copies (the kernel
// before the lane-local item loop) the unrolled form lost ~5 % to instruction-fetch stalls.
#define FLASHE_PRAGMA_(x) _Pragma(#x)
#define FLASHE_PRAGMA(x) FLASHE_PRAGMA_(x)
#ifndef FLASHE_AES_UNROLL
#define FLASHE_AES_UNROLL 5
#endif
#define AES_ROUNDS_UNROLL FLASHE_PRAGMA(unroll FLASHE_AES_UNROLL)
#ifndef FLASHE_AES_UNROLL_M6
// The m = 5, 6 instantiation has three unrolled quad bodies in its hot loop; with the rounds unrolled as
// well it lost 12 % of its issue slots to instruction fetch (ncu stall_no_inst).  Rolled rounds there:
// 25M x 10 clients at int_bits 20, encode 2.60 -> 2.45 ms.
#define FLASHE_AES_UNROLL_M6 1
#endif
#ifndef FLASHE_AES_UNROLL_A6
#define FLASHE_AES_UNROLL_A6 1      // the m == 6, aligned instantiation (b = 20)
#endif

// ------------------------------------------------------------------------------------------------
// k_stream: persistent, one 512-thread CTA per SM (128 KB of tables + per-warp slabs).
//
// Work unit ("warp item") = NB*32 = 64 consecutive AES blocks of one reference chunk: lane l owns
// blocks i0+l and i0+32+l, i.e. 2*m elements.  Per item and client:
//   1. prefetch the item's input elements (pairs (2p, 2p+1) of the global index) into registers so the
//      DRAM latency hides under the AES work;
//   2. per stream: two interleaved AES-256 blocks per lane, slots accumulated with sign in registers;
//   3. transpose lane-major -> element-major through the warp's shared slab;
//   4. walk the item's element pairs: noise (one Philox per pair), encode / decode, modular add,
//      coalesced stores.
// ------------------------------------------------------------------------------------------------
#define NB 2

// Counter-window factoring.  The AES input is iter || prf || ctr_hi || ctr_lo and only ctr_lo's low
// byte differs between the counters of one 256-aligned window.  After round 1 that byte has reached
// column 0 only (p0); columns 1-3 are window constants.  In round 2 every output column takes exactly
// one byte of column 0, so three of its four lookups are window constants too: c0..c3 below (round key
// folded in).  Per block, rounds 1-2 then cost 1 + 4 lookups instead of 4 + 16 (197 per block instead
// of 212); the 15 lookups of window_consts are paid once per window and stream, or once per lane pair.
struct WinC { uint32_t c0, c1, c2, c3; };
__device__ __forceinline__ WinC window_consts(const KeySched& ks, uint32_t y, Pre pre, uint32_t w3) {
    const uint32_t s3 = w3 ^ ks.rk[3];
    const uint32_t p1 = pre.p1 ^ T2(s3), p2 = pre.p2 ^ T1(s3), p3 = pre.p3 ^ T0(s3);
    WinC c;
    c.c0 = T1(p1) ^ T2(p2) ^ T3(p3) ^ ks.rk[8];
    c.c1 = T0(p1) ^ T1(p2) ^ T2(p3) ^ ks.rk[9];
    c.c2 = T0(p2) ^ T1(p3) ^ T3(p1) ^ ks.rk[10];
    c.c3 = T0(p3) ^ T2(p1) ^ T3(p2) ^ ks.rk[11];
    return c;
}

// Two AES-256 blocks of the SAME stream and the same counter window (counters w3a, w3b; words 0-2 shared,
// word 2 == 0) computed in one instruction stream: twice the independent lookups per round, so the
// round-boundary latency (LDS ~30 clk + LOP3) of one block hides under the other's.
template <int UNROLL = FLASHE_AES_UNROLL>
__device__ __forceinline__ void aes256_x2w(const KeySched& ks, uint32_t y, uint32_t pre_p0, WinC c, uint32_t w3a, uint32_t w3b,
                                           uint32_t oa[4], uint32_t ob[4]) {
    uint32_t a0, a1, a2, a3, b0, b1, b2, b3, p0, p1, p2, p3, q0, q1, q2, q3;
    p0 = pre_p0 ^ T3(w3a ^ ks.rk[3]); q0 = pre_p0 ^ T3(w3b ^ ks.rk[3]);      // round 1, column 0
    a0 = c.c0 ^ T0(p0); b0 = c.c0 ^ T0(q0);                                   // round 2
    a1 = c.c1 ^ T3(p0); b1 = c.c1 ^ T3(q0);
    a2 = c.c2 ^ T2(p0); b2 = c.c2 ^ T2(q0);
    a3 = c.c3 ^ T1(p0); b3 = c.c3 ^ T1(q0);
    p0 = T0(a0) ^ T1(a1) ^ T2(a2) ^ T3(a3) ^ ks.rk[12];                       // round 3
    q0 = T0(b0) ^ T1(b1) ^ T2(b2) ^ T3(b3) ^ ks.rk[12];
    p1 = T0(a1) ^ T1(a2) ^ T2(a3) ^ T3(a0) ^ ks.rk[13];
    q1 = T0(b1) ^ T1(b2) ^ T2(b3) ^ T3(b0) ^ ks.rk[13];
    p2 = T0(a2) ^ T1(a3) ^ T2(a0) ^ T3(a1) ^ ks.rk[14];
    q2 = T0(b2) ^ T1(b3) ^ T2(b0) ^ T3(b1) ^ ks.rk[14];
    p3 = T0(a3) ^ T1(a0) ^ T2(a1) ^ T3(a2) ^ ks.rk[15];
    q3 = T0(b3) ^ T1(b0) ^ T2(b1) ^ T3(b2) ^ ks.rk[15];
#pragma unroll UNROLL
    for (int r = 4; r < 14; r += 2) {                                         // rounds 4..13
        // both round keys of the iteration as two 128-bit constant-bank loads (the rolled loop indexes them)
        const uint4 k0 = *reinterpret_cast<const uint4*>(&ks.rk[4 * r]), k1 = *reinterpret_cast<const uint4*>(&ks.rk[4 * r + 4]);
        a0 = T0(p0) ^ T1(p1) ^ T2(p2) ^ T3(p3) ^ k0.x;
        b0 = T0(q0) ^ T1(q1) ^ T2(q2) ^ T3(q3) ^ k0.x;
        a1 = T0(p1) ^ T1(p2) ^ T2(p3) ^ T3(p0) ^ k0.y;
        b1 = T0(q1) ^ T1(q2) ^ T2(q3) ^ T3(q0) ^ k0.y;
        a2 = T0(p2) ^ T1(p3) ^ T2(p0) ^ T3(p1) ^ k0.z;
        b2 = T0(q2) ^ T1(q3) ^ T2(q0) ^ T3(q1) ^ k0.z;
        a3 = T0(p3) ^ T1(p0) ^ T2(p1) ^ T3(p2) ^ k0.w;
        b3 = T0(q3) ^ T1(q0) ^ T2(q1) ^ T3(q2) ^ k0.w;
        p0 = T0(a0) ^ T1(a1) ^ T2(a2) ^ T3(a3) ^ k1.x;
        q0 = T0(b0) ^ T1(b1) ^ T2(b2) ^ T3(b3) ^ k1.x;
        p1 = T0(a1) ^ T1(a2) ^ T2(a3) ^ T3(a0) ^ k1.y;
        q1 = T0(b1) ^ T1(b2) ^ T2(b3) ^ T3(b0) ^ k1.y;
        p2 = T0(a2) ^ T1(a3) ^ T2(a0) ^ T3(a1) ^ k1.z;
        q2 = T0(b2) ^ T1(b3) ^ T2(b0) ^ T3(b1) ^ k1.z;
        p3 = T0(a3) ^ T1(a0) ^ T2(a1) ^ T3(a2) ^ k1.w;
        q3 = T0(b3) ^ T1(b0) ^ T2(b1) ^ T3(b2) ^ k1.w;
    }
#define LAST(a, b, c, d, k)                                                                         \
    (__byte_perm(__byte_perm(lds_tab<128>(__byte_perm((d), y, SEL_B0)),                             \
                             lds_tab<0>(__byte_perm((c), y, SEL_B1)), 0x3250),                      \
                 __byte_perm(lds_tab<0x10080>(__byte_perm((b), y, SEL_B2)),                         \
                             lds_tab<0x10000>(__byte_perm((a), y, SEL_B3)), 0x7210), 0x7610) ^ ks.rk[k])
    oa[0] = LAST(p0, p1, p2, p3, 56); ob[0] = LAST(q0, q1, q2, q3, 56);
    oa[1] = LAST(p1, p2, p3, p0, 57); ob[1] = LAST(q1, q2, q3, q0, 57);
    oa[2] = LAST(p2, p3, p0, p1, 58); ob[2] = LAST(q2, q3, q0, q1, 58);
    oa[3] = LAST(p3, p0, p1, p2, 59); ob[3] = LAST(q3, q0, q1, q2, 59);
#undef LAST
}

template <int MODE, int WORDS> struct InType { typedef typename Word<WORDS>::T T; };
template <int WORDS> struct InType<M_ENCODE, WORDS> { typedef float T; };

// ALIGNED (4-byte words, m = 4 only): the host has checked that every chunk of the span starts on a
// multiple of 4 elements, so the lane-local path is 128-bit accesses without an alignment switch.
// M6S (4-byte words, MMAX = 6): m is exactly 6 like in the ALIGNED instantiation, but the chunks may start anywhere
// (SHIFT6 in the kernel); the purely aligned layout keeps its own instantiation, which is 3 % faster without that code.
template <int WORDS, int MMAX, int MODE, bool SHARE, bool ALIGNED, bool N32 = false, bool M6S = false>
__global__ void __launch_bounds__(STREAM_THREADS, 1)
k_stream(const __grid_constant__ KeySched ks, const __grid_constant__ StreamTab st, const __grid_constant__ Geom g,
         const __grid_constant__ IoDev io, const __grid_constant__ CodecDev cd, const __grid_constant__ NoiseDev nz) {
    typedef Word<WORDS> WT;
    typedef typename WT::T word_t;
    typedef typename InType<MODE, WORDS>::T in_t;
    constexpr bool HAS_IN = (MODE == M_APPLY || MODE == M_ENCODE || MODE == M_DECODE);
    constexpr int PF = MMAX + 1;                 // pair iterations per item: ceil((64m + 2) / 64)
    constexpr uint32_t WB = WORDS * 4u;
    constexpr bool QUAD_OK = (WORDS == 1 && MODE != M_SCATTER && MMAX <= 6);   // 4-byte words, m = 4 (b 25..32) or 5, 6 (b 20..25)
    constexpr int NQ = (MMAX + 1) / 2;            // 16-byte element quads per lane and item: 64 m / 4 / 32, rounded up
    // m = 4, chunk starts off a 16-byte boundary: instead of 64/32-bit pieces the lanes own the memory-ALIGNED
    // quads and receive the 1-3 mask words that belong to the neighbouring block by shuffle (see fast_item)
    constexpr bool SHIFT_OK = (WORDS == 1 && MMAX == 4 && !ALIGNED && !SHARE && MODE != M_SCATTER);
    // ALIGNED or M6S with MMAX == 6: m is exactly 6 (b = 19..21: the reference's shipped width is 20), a compile-time
    // constant: an item is 96 whole 16-byte quads (three per lane).  ALIGNED: every chunk starts on a multiple of 4
    // elements; M6S: chunks that start off a 16-byte boundary are handled by SHIFT6 below
    constexpr bool A6 = (WORDS == 1 && MMAX == 6 && (ALIGNED || M6S));
    // m = 6, chunk starts off a 16-byte boundary of the buffers (what an arbitrary model length gives: L / n_jobs is
    // rarely a multiple of 4): the masks already pass through the warp's slab, so they are written there SHIFTED by the
    // misalignment and read back as the MEMORY-aligned quads; the quad that straddles two items is completed by the
    // previous item's last words, which stay in the slab (see fast_item).  Same idea as SHIFT_OK, through the slab.
    constexpr bool SHIFT6 = (WORDS == 1 && MMAX == 6 && M6S && !ALIGNED && !SHARE && MODE != M_SCATTER);
    // 16-byte words (the shipped 120-bit batch mode), m = 1: masks / apply always; encode / decode when the codec
    // batches lanes into the word (cd.bs != 0: encode -> pack -> mask and unmask -> unpack -> decode fused)
    constexpr bool W4_OK = (WORDS == 4 && !SHARE && MODE != M_SCATTER);
    constexpr bool W4_CODEC = (WORDS == 4 && (MODE == M_ENCODE || MODE == M_DECODE));
    const bool w4_batched = W4_CODEC && cd.bs != 0u;
    constexpr bool W2_OK = (WORDS == 2 && MODE != M_SCATTER);                      // 8-byte words with m = 2 (b = 43..64)
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t y = 0x00010000u | (lane << 2);
    const uint32_t sbase = smem_window_base();
    const uint32_t slab_bytes = (NB * 32u * MMAX + 2u + (WORDS == 1 ? 2u * MMAX + 1u : 0u)) * WB;
    if (sbase + nwarps * slab_bytes > TAB_BASE) { __trap(); }
    const uint32_t slab = sbase + warp * slab_bytes;
    // per-warp cache of window terms: one 16-byte slot per stream-table entry, above the slabs
    const uint32_t wcache_all = (sbase + nwarps * slab_bytes + 15u) & ~15u;
    const bool cache_ok = wcache_all + nwarps * (MAXS * 16u) <= TAB_BASE;
    const uint32_t wcache = wcache_all + warp * (MAXS * 16u);
#define PRE_OF(k) Pre{st.pre[k][0], st.pre[k][1], st.pre[k][2], st.pre[k][3]}

    // Programmatic dependent launch: this grid may have been started while the previous kernel of the stream was
    // still draining (launch_stream_t sets the attribute).  Let OUR successor start as early as it can, build the
    // tables (they depend on nothing a predecessor writes), and only then wait for the predecessor's memory.
#ifdef FLASHE_TRACE
    unsigned long long tr_begin, tr_wait, tr_edge = 0ull, tr_items = 0ull, tr_nedge = 0ull;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_begin));
#endif
    asm volatile("griddepcontrol.launch_dependents;");
    fill_tables(io.te0);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __syncthreads();
#ifdef FLASHE_TRACE
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_wait));
#endif

    // slab word index -> byte offset; 4-byte words are skewed by one word per 32 so that the
    // lane-major stores (stride m) and the element-major loads never pile onto one bank
    auto sl = [&](uint32_t i) -> uint32_t { return slab + (WORDS == 1 ? (i + (i >> 5)) : i) * WB; };

    const word_t mk = WT::mask(g.b);
    const uint32_t m = g.m;
    const uint64_t n_units = (st.batch && !io.share) ? g.S_cnt * io.n_clients : g.S_cnt;
    // CTA-minor numbering of the warps: consecutive units go to DIFFERENT SMs first, so the units of a partial
    // last wave spread evenly over the SMs instead of filling the first CTAs' warps (at 1 M elements x 3 clients
    // that was 64 items on 45 SMs against 48 on the rest)
    const uint64_t gw = (uint64_t)warp * gridDim.x + blockIdx.x, gstride = (uint64_t)gridDim.x * nwarps;

    // Units are dealt round-robin, one per warp and wave.  The LAST full wave and the leftover units behind it are
    // cut into `fine` pieces of consecutive items each and dealt the same way, so that the warps finish within
    // one piece (sup / 8 items) of each other instead of one unit (at 12.5 M elements x 64 clients per GPU a
    // unit is ~240 us of a ~10 ms launch).
    const uint64_t full_waves = n_units / gstride;
    const uint64_t n_main = full_waves >= 1 ? (full_waves - 1) * gstride : 0;
    const uint32_t fine = g.sup < 8u ? g.sup : 8u;
    const uint32_t piece_items = (g.sup + fine - 1u) / fine;
    const uint64_t n_virtual = n_main + (n_units - n_main) * fine;
    // Dealing.  Static (io.tickets == NULL): warp gw takes v = gw, gw + gstride, ...  Dynamic: the warps claim v from a
    // ticket counter in global memory, so a warp that drew slow units (edge items of a chunk run 1.6x a whole item; SMs
    // and warps of one SM do not progress at exactly the same rate) simply claims fewer: measured on the per-warp
    // timeline, the static deal left 8 % (2.5 M elements x 10 clients) to 20 % (1 M x 3) between the median and the last
    // warp.  The fine pieces are claimed in REVERSE order: the very last piece of the list is the last chunk's trailing
    // edge item, which would otherwise be the tail of every launch.  Every warp draws exactly one ticket past the end;
    // it then bumps the `done` counter, and the last warp to do so zeroes both for the next launch on this slot
    // (flashe_ticket_slot: one slot per stream, or per captured launch).
    uint32_t* const tk = io.tickets;
    uint64_t v_static = gw;
    bool first_draw = true;
    const uint64_t rev_lo = n_main > gstride ? n_main : gstride;      // (the first wave is dealt statically, see below)
    auto next_v = [&]() -> uint64_t {
        if (tk && !first_draw) {
            uint32_t k32 = 0u;
            if (lane == 0u) k32 = atomicAdd(tk, 1u);
            const uint64_t k = (uint64_t)__shfl_sync(0xffffffffu, k32, 0) + gstride;
            return k < rev_lo || k >= n_virtual ? k : rev_lo + (n_virtual - 1u - k);
        }
        // static deal, and the first unit of every warp under the dynamic deal (2368 warps drawing from one address at
        // the same instant is 2 us of serialised atomics before any work starts)
        first_draw = false;
        const uint64_t v = v_static;
        v_static += gstride;
        return v;
    };
    for (uint64_t v = next_v(); v < n_virtual; v = next_v()) {
        const bool whole = v < n_main;
        const uint64_t t = whole ? v : n_main + (v - n_main) / fine;
        const uint32_t piece = whole ? 0u : (uint32_t)((v - n_main) % fine);
        uint32_t c_first = 0, c_count = 1;
        uint64_t S = t;
        if (st.batch) {
            if (io.share) { c_first = 0; c_count = io.n_clients; }
            else { c_first = (uint32_t)(t / g.S_cnt); S = t - (uint64_t)c_first * g.S_cnt; }
        }
        // a work unit = up to g.sup consecutive warp items of one chunk: the 64-bit divisions of the
        // chunk rule are paid once per unit, the items inside advance by ITEM_BLOCKS
        uint32_t nsub;
        Item it = decode_unit(g, g.S_lo + S, nsub);
        if (!whole) {                                                  // this warp's piece of the unit's items
            const uint32_t lo = piece * piece_items;
            if (lo >= nsub) continue;
            it.w += lo;
            nsub = nsub - lo < piece_items ? nsub - lo : piece_items;
        }
      uint32_t cached_win = 0xffffffffu;                             // counter window the cached round-2 terms belong to
      const uint32_t n_iter = SHARE ? c_count + 1 : c_count;
      // ---- lane-local items ------------------------------------------------------------------------
      // Items [wf_lo, wf_hi) of the unit's chunk are FULL (64 blocks of m = 4 elements), lie inside the
      // shard and have 32-bit counters: a lane's AES block IS four consecutive elements, so the lane
      // loads / stores them itself (a warp covers 512 contiguous bytes per access), nothing goes
      // through the slab, and all per-item geometry is a handful of additions.  The bounds are
      // computed once per unit.
      uint64_t wf_lo = 1, wf_hi = 0;
      const uint32_t mm = (WORDS == 1 && MMAX == 4) ? 4u : (A6 ? 6u : (WORDS == 4 ? 1u : m));   // compile-time where the instantiation fixes it
      const uint32_t item_elems = ITEM_BLOCKS * mm;
      // (8-byte words: only m = 2, and only chunks that start on an even element of an even shard, so that a
      //  block is one aligned 16-byte pair and one noise pair)
      const bool w2_here = W2_OK && m == 2u && ((it.cb | g.begin) & 1ull) == 0ull;
      const bool w4_here = W4_OK && (!W4_CODEC || w4_batched);
      if ((QUAD_OK || w4_here || w2_here) && io.quad && !((SHIFT_OK || SHIFT6) && (g.begin & 1ull))) {
          const uint64_t shift = (uint64_t)mm * it.off;             // e0(w) = cb - shift + 64 m w
          const uint64_t full_end = it.cb + (MMAX == 4 ? (it.clen & ~3ull) : (it.clen / mm) * mm);   // end of the chunk's last whole block
          const uint64_t hi_e = (full_end < g.end ? full_end : g.end) + shift;
          const uint64_t lo_e = (g.begin > it.cb ? g.begin : it.cb) + shift;   // w = 0 is lane-local only when off == 0
          wf_lo = lo_e > it.cb ? (lo_e - it.cb + item_elems - 1u) / item_elems : 0;
          wf_hi = hi_e > it.cb ? (hi_e - it.cb) / item_elems : 0;    // items w with 64 m (w+1) <= hi_e - cb
          const uint64_t c0 = it.cb - it.off;                       // counter of item 0's (virtual) first block
          const uint64_t w32 = c0 < (1ull << 32) ? ((1ull << 32) - c0) >> 6 : 0;   // 64 (w+1) <= 2^32 - c0
          if (wf_hi > w32) wf_hi = w32;
      }
      // m = 5, 6: the lane-major masks are turned element-major through the warp's slab (16-byte aligned part)
      const uint32_t fslab = (slab + 15u) & ~15u;
      // shifted mode (SHIFT_OK, misaligned chunk): the unit's lane-local items [wA, wBx) form one run; the last
      // qr mask words of an item travel to the next item in lane 31's `carry` registers
      const uint64_t wA = it.w > wf_lo ? it.w : wf_lo, wBx = it.w + nsub < wf_hi ? it.w + nsub : wf_hi;
      uint32_t carry0 = 0u, carry1 = 0u, carry2 = 0u;
      auto fast_item = [&](uint64_t w) {
        if constexpr (QUAD_OK) {
          const uint64_t e0 = it.cb - (uint64_t)mm * it.off + w * item_elems;   // first global element of the item
          const uint32_t ctr0 = (uint32_t)(it.cb - it.off) + ((uint32_t)w << 6);   // jzf_flashe.py:34 "(i + begin)"
          const uint64_t o0 = e0 - g.begin;
          const uint32_t qr0 = ALIGNED ? 0u : ((uint32_t)o0 & 3u);   // misalignment of the chunk in the buffers
          // Shifted mode: quads start qr0 elements BEFORE the item (aligned in memory, aligned noise pairs); quad
          // q's first qr0 mask words come from the block before it.  Quad 0 of the run's first item is partial
          // (lane 0 handles its own elements one by one), and so are the qr0 elements after the run's last quad.
          const bool shifted6 = SHIFT6 && qr0 != 0u;
          const bool shifted = (SHIFT_OK && qr0 != 0u) || shifted6;
          const bool run_first = shifted && w == wA, run_last = shifted && w + 1 == wBx;
          // alignment switch of the 16-byte accesses: with SHIFT_OK every quad is aligned (qr0 != 0 => shifted),
          // which removes the 64/32-bit piece code from this instantiation's hot loop
          const uint32_t qr = (SHIFT_OK || SHIFT6) ? 0u : qr0;
          const uint64_t o0q = shifted ? o0 - qr0 : o0, e0q = shifted ? e0 - qr0 : e0;
          const uint32_t nquads = item_elems >> 2;                  // 16 m; lane owns quads lane + 32 k
          const uint32_t ctrA = ctr0 + lane, ctrB = ctrA + 32u;
          const uint32_t win = ctr0 >> 8;                           // same for every counter of the item
          const bool stale = !cache_ok || win != cached_win;
          const uint32_t mk32 = Word<1>::mask(g.b);
          const bool one_seg = cd.nseg == 1;
          const bool one_rcp = one_seg && MODE == M_ENCODE && cd.seg[0].rcp_two_a != 0.0f;
          uint32_t prev[NB][MMAX];
          const uint32_t n_iter_here = SHARE ? n_iter : 1u;          // without SHARE a unit serves exactly one client
          for (uint32_t cc = 0; cc < n_iter_here; ++cc) {
              const uint32_t c = SHARE ? (cc ? c_first + cc - 1 : 0) : c_first + cc;
              const bool emit = !SHARE || cc > 0;
              uint32_t r[NQ][4];
              if (HAS_IN && emit) {                                  // inputs first: their latency hides under the AES rounds
                  const uint32_t* in = reinterpret_cast<const uint32_t*>(io.in) + (uint64_t)c * io.in_stride + o0q;
#pragma unroll
                  for (int k = 0; k < NQ; ++k) {
                      const uint32_t q = lane + 32u * k;
                      if ((MMAX == 4 || A6 || q < nquads) && !(run_first && q == 0u)) ldg_quad(in + 4u * q, qr, r[k]);
                  }
              }
              uint32_t acc[NB][MMAX];
#pragma unroll
              for (int h = 0; h < NB; ++h)
#pragma unroll
                  for (int k = 0; k < MMAX; ++k) acc[h][k] = 0u;
              uint32_t s_begin, s_count;
              if (!st.batch) { s_begin = 0; s_count = st.n; }
              else if (SHARE) { s_begin = cc; s_count = 1; }
              else { s_begin = c; s_count = st.dbl ? 2u : 1u; }
              for (uint32_t si = 0; si < s_count; ++si) {
                  const uint32_t sidx = s_begin + si;
                  const int sign = st.batch ? (si == 0 ? +1 : -1) : st.sign[sidx];
                  // window terms of stream sidx: the warp's cache slot (broadcast read), or recomputed by
                  // every lane (same inputs, same result) when the item opens a new counter window
                  WinC wc;
                  const uint32_t slot = wcache + sidx * 16u;
                  if (stale) {
                      wc = window_consts(ks, y, PRE_OF(sidx), ctr0);
                      if (cache_ok) {
                          if (lane == 0) asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(slot), "r"(wc.c0), "r"(wc.c1), "r"(wc.c2), "r"(wc.c3) : "memory");
                          __syncwarp();
                      }
                  } else {
                      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(wc.c0), "=r"(wc.c1), "=r"(wc.c2), "=r"(wc.c3) : "r"(slot) : "memory");
                  }
                  uint32_t oa[4], ob[4];
                  aes256_x2w<(MMAX == 4 ? FLASHE_AES_UNROLL : (A6 ? FLASHE_AES_UNROLL_A6 : FLASHE_AES_UNROLL_M6))>(ks, y, st.pre[sidx][0], wc, ctrA, ctrB, oa, ob);
                  accumulate_slots<1, MMAX>(oa, g.b, mm, sign, acc[0]);
                  accumulate_slots<1, MMAX>(ob, g.b, mm, sign, acc[1]);
              }
              if (SHARE) {                                           // acc = F(cc); mask of client cc-1 = prev - acc
#pragma unroll
                  for (int h = 0; h < NB; ++h)
#pragma unroll
                      for (int k = 0; k < MMAX; ++k) {
                          const uint32_t cur = acc[h][k];
                          if (cc > 0) acc[h][k] = prev[h][k] - cur;
                          prev[h][k] = cur;
                      }
                  if (!emit) continue;
              }
              if (MMAX != 4) {
                  // lane-major -> element-major: block (lane + 32 h) holds elements [(lane + 32 h) m, +m).  Stride m
                  // words (m = 5: odd; m = 6: written as 64-bit pairs, 16 lanes x 24 bytes hit 32 distinct banks):
                  // conflict-free.  Read back as 16-byte quads.
                  __syncwarp();                                      // the previous round's quads have been read
                  // shifted: word i of the item goes to slab word i + qr0, so that memory-aligned quad q sits at 16 q; the
                  // first qr0 slab words (the head of quad 0) are the previous item's last qr0 mask words, which lie at
                  // slab words 384 .. 384 + qr0 - 1 until this item overwrites them
                  uint32_t carry_w = 0u;
                  if (SHIFT6 && shifted6 && !run_first && lane < qr0) carry_w = lds32(fslab + 4u * (384u + lane));
                  if (SHIFT6 && shifted6) __syncwarp();
                  if (SHIFT6 && shifted6 && !run_first && lane < qr0) sts32(fslab + 4u * lane, carry_w);
#pragma unroll
                  for (int h = 0; h < NB; ++h) {
                      const uint32_t a0 = fslab + (lane + 32u * h) * mm * 4u + (SHIFT6 && shifted6 ? 4u * qr0 : 0u);
                      if (SHIFT6 && shifted6 && (qr0 & 1u)) {        // odd shift: the pairs are not 8-byte aligned
#pragma unroll
                          for (int k = 0; k < MMAX; ++k) sts32(a0 + 4u * k, acc[h][k]);
                      } else if (A6 || mm == 6u) {
#pragma unroll
                          for (int k = 0; k + 1 < MMAX; k += 2)
                              asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a0 + 4u * k), "r"(acc[h][k]), "r"(acc[h][k + 1]) : "memory");
                      } else {
#pragma unroll
                          for (int k = 0; k < MMAX; ++k)
                              if ((uint32_t)k < mm) sts32(a0 + 4u * k, acc[h][k]);
                      }
                  }
                  __syncwarp();
              }
              uint32_t e4[4] = {0u, 0u, 0u, 0u};                     // shifted mode: the words of the run's edge elements
              if constexpr (SHIFT_OK) {
                  if (shifted) {
                      const bool head = run_first && lane == 0u;     // lane 0: its block A; lane 31: its block B
#pragma unroll
                      for (int k = 0; k < 4; ++k) e4[k] = head ? acc[0][k] : acc[1][k];
                      const uint32_t srcl = (lane + 31u) & 31u;      // every lane reads its left neighbour, lane 0 reads lane 31
                      const bool l31 = lane == 31u;                  // ... which forwards the tail of the block BEFORE lane 0's
#define ROT(own, before) __shfl_sync(0xffffffffu, l31 ? (before) : (own), srcl)
                      uint32_t n0[4], n1[4];
                      if (qr0 == 1u) {
                          n0[0] = ROT(acc[0][3], carry0); n1[0] = ROT(acc[1][3], acc[0][3]);
                          n0[1] = acc[0][0]; n0[2] = acc[0][1]; n0[3] = acc[0][2];
                          n1[1] = acc[1][0]; n1[2] = acc[1][1]; n1[3] = acc[1][2];
                          carry0 = acc[1][3];
                      } else if (qr0 == 2u) {
                          n0[0] = ROT(acc[0][2], carry0); n0[1] = ROT(acc[0][3], carry1);
                          n1[0] = ROT(acc[1][2], acc[0][2]); n1[1] = ROT(acc[1][3], acc[0][3]);
                          n0[2] = acc[0][0]; n0[3] = acc[0][1]; n1[2] = acc[1][0]; n1[3] = acc[1][1];
                          carry0 = acc[1][2]; carry1 = acc[1][3];
                      } else {
                          n0[0] = ROT(acc[0][1], carry0); n0[1] = ROT(acc[0][2], carry1); n0[2] = ROT(acc[0][3], carry2);
                          n1[0] = ROT(acc[1][1], acc[0][1]); n1[1] = ROT(acc[1][2], acc[0][2]); n1[2] = ROT(acc[1][3], acc[0][3]);
                          n0[3] = acc[0][0]; n1[3] = acc[1][0];
                          carry0 = acc[1][1]; carry1 = acc[1][2]; carry2 = acc[1][3];
                      }
#undef ROT
#pragma unroll
                      for (int k = 0; k < 4; ++k) { acc[0][k] = n0[k]; acc[1][k] = n1[k]; }
                  }
              }
#pragma unroll
              for (int h = 0; h < NQ; ++h) {
                  const uint32_t q = lane + 32u * h;                 // this lane's h-th quad of the item
                  if (MMAX != 4 && !A6 && q >= nquads) break;
                  if ((SHIFT_OK || SHIFT6) && run_first && q == 0u) continue;    // partial quad: handled element-wise below
                  const uint64_t o = o0q + 4u * q;
                  const uint64_t j = e0q + 4u * q;
                  uint32_t mw[4];
                  if (MMAX == 4) {
#pragma unroll
                      for (int k = 0; k < 4; ++k) mw[k] = acc[h < NB ? h : 0][k];
                  } else {
                      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(mw[0]), "=r"(mw[1]), "=r"(mw[2]), "=r"(mw[3]) : "r"(fslab + 16u * q) : "memory");
                  }
                  // (mw is reduced mod 2^b together with the sum below; only the mask output needs it by itself)
                  if (MODE == M_MASKS) {                       // (batch: one row per stream entry; otherwise c = 0)
                      stg_quad(reinterpret_cast<uint32_t*>(io.out) + (uint64_t)c * io.out_stride + o, qr, mw[0] & mk32, mw[1] & mk32, mw[2] & mk32, mw[3] & mk32);
                  } else if (MODE == M_APPLY) {
                      uint32_t* out = reinterpret_cast<uint32_t*>(io.out) + (uint64_t)c * io.out_stride + o;
                      stg_quad(out, qr, (r[h][0] + mw[0]) & mk32, (r[h][1] + mw[1]) & mk32, (r[h][2] + mw[2]) & mk32, (r[h][3] + mw[3]) & mk32);
                  } else if (MODE == M_ENCODE) {
                      double u[4];
                      if (nz.u) {
                          const double* up = nz.u + (uint64_t)c * nz.u_stride + o;
#pragma unroll
                          for (int k = 0; k < 4; ++k) u[k] = up[k];
                      } else if (N32 && (j & 3ull) == 0ull) {                // 32-bit resolution: the quad is one generator call
                          noise_quad<true>(nz, nz.stream + c, j >> 2, u);
                      } else if (ALIGNED || SHIFT_OK || SHIFT6 || (j & 1ull) == 0ull) {   // (SHIFT_OK / SHIFT6: quads start at begin + 4k, begin even)
                          noise_pair<N32>(nz, nz.stream + c, j >> 1, u[0], u[1]);
                          noise_pair<N32>(nz, nz.stream + c, (j >> 1) + 1, u[2], u[3]);
                      } else {                      // odd chunk start: the four elements touch three pairs
                          double lo, hi;
                          noise_pair<N32>(nz, nz.stream + c, j >> 1, lo, u[0]);
                          noise_pair<N32>(nz, nz.stream + c, (j >> 1) + 1, u[1], u[2]);
                          noise_pair<N32>(nz, nz.stream + c, (j >> 1) + 2, u[3], hi);
                      }
                      uint32_t q4[4];
                      if (one_rcp) {                                 // single layer with a usable reciprocal (warp-uniform)
#pragma unroll
                          for (int k = 0; k < 4; ++k) q4[k] = encode_one<true>(__uint_as_float(r[h][k]), u[k], cd.seg[0], cd.scale);
                      } else {
                          Seg sg = find_seg(cd, j);
#pragma unroll
                          for (int k = 0; k < 4; ++k) {
                              if (k && !one_seg && j + k >= sg.end) sg = find_seg(cd, j + k);
                              q4[k] = encode_one(__uint_as_float(r[h][k]), u[k], sg, cd.scale);
                          }
                      }
                      if (io.aux) stg_quad(reinterpret_cast<uint32_t*>(io.aux) + (uint64_t)c * io.out_stride + o, qr, q4[0], q4[1], q4[2], q4[3]);
                      uint32_t* out = reinterpret_cast<uint32_t*>(io.out) + (uint64_t)c * io.out_stride + o;
                      stg_quad(out, qr, (q4[0] + mw[0]) & mk32, (q4[1] + mw[1]) & mk32, (q4[2] + mw[2]) & mk32, (q4[3] + mw[3]) & mk32);
                  } else if (MODE == M_DECODE) {
                      uint32_t pw[4];
                      double dv[4];
                      Seg sg = find_seg(cd, j);
#pragma unroll
                      for (int k = 0; k < 4; ++k) {
                          pw[k] = (r[h][k] + mw[k]) & mk32;
                          if (k && !one_seg && j + k >= sg.end) sg = find_seg(cd, j + k);
                          dv[k] = decode_one((double)pw[k], sg.two_an, cd.den, cd.den_rcp, sg.an);
                      }
                      if (io.aux) stg_quad(reinterpret_cast<uint32_t*>(io.aux) + o, qr, pw[0], pw[1], pw[2], pw[3]);
                      stg_quad_f64(io.outf + o, qr, dv);
                  }
              }
              if constexpr (SHIFT6) {
                  // edges of a shifted m = 6 run, one element per LANE: the 4 - qr0 elements before the first item's first
                  // aligned quad (lanes 0 ..) and the last qr0 elements of the last item (lanes .. 31); their masks are in the
                  // slab, so any lane can take one (a single lane looping over them was 2 x 3 dependent passes of noise +
                  // encode per run: 7 % of a launch whose runs are three items long)
                  uint32_t i_edge = 0xffffffffu;
                  if (shifted6 && run_first && lane < 4u - qr0) i_edge = lane;
                  else if (shifted6 && run_last && lane >= 32u - qr0) i_edge = 352u + lane;      // 384 - qr0 + (lane - (32 - qr0))
                  if (i_edge != 0xffffffffu) {
                      {
                          const uint32_t i = i_edge;
                          const uint32_t mword = lds32(fslab + 4u * (i + qr0));
                          const uint64_t o = o0 + i, j = e0 + i;
                          if (MODE == M_MASKS) {
                              reinterpret_cast<uint32_t*>(io.out)[(uint64_t)c * io.out_stride + o] = mword & mk32;
                          } else if (MODE == M_APPLY) {
                              const uint64_t oc = (uint64_t)c * io.out_stride + o;
                              reinterpret_cast<uint32_t*>(io.out)[oc] = (reinterpret_cast<const uint32_t*>(io.in)[(uint64_t)c * io.in_stride + o] + mword) & mk32;
                          } else if (MODE == M_ENCODE) {
                              const float x = reinterpret_cast<const float*>(io.in)[(uint64_t)c * io.in_stride + o];
                              const double u = nz.u ? nz.u[(uint64_t)c * nz.u_stride + o] : noise_one<N32>(nz, nz.stream + c, j);
                              const Seg sg = find_seg(cd, j);
                              const uint32_t qv = encode_one(x, u, sg, cd.scale);
                              const uint64_t oc = (uint64_t)c * io.out_stride + o;
                              if (io.aux) reinterpret_cast<uint32_t*>(io.aux)[oc] = qv;
                              reinterpret_cast<uint32_t*>(io.out)[oc] = (qv + mword) & mk32;
                          } else if (MODE == M_DECODE) {
                              const uint32_t pw = (reinterpret_cast<const uint32_t*>(io.in)[o] + mword) & mk32;
                              if (io.aux) reinterpret_cast<uint32_t*>(io.aux)[o] = pw;
                              const Seg sg = find_seg(cd, j);
                              io.outf[o] = decode_one((double)pw, sg.two_an, cd.den, cd.den_rcp, sg.an);
                          }
                      }
                  }
              }
              if constexpr (SHIFT_OK) {
                  // edges of a shifted run, one element at a time: the 4 - qr0 elements of the first item's block 0
                  // (lane 0) and the last qr0 elements of the last item's block 63 (lane 31)
                  const bool head = run_first && lane == 0u, tail = run_last && lane == 31u;
                  if (head || tail) {
                      const uint32_t i_lo = head ? 0u : 4u - qr0, i_hi = head ? 4u - qr0 : 4u;
                      const uint64_t ob = o0 + (head ? 0u : 252u), jb = e0 + (head ? 0u : 252u);
#pragma unroll 1
                      for (uint32_t i = i_lo; i < i_hi; ++i) {
                          const uint32_t mword = i == 0u ? e4[0] : (i == 1u ? e4[1] : (i == 2u ? e4[2] : e4[3]));
                          const uint64_t o = ob + i, j = jb + i;
                          if (MODE == M_MASKS) {
                              reinterpret_cast<uint32_t*>(io.out)[(uint64_t)c * io.out_stride + o] = mword & mk32;
                          } else if (MODE == M_APPLY) {
                              const uint64_t oc = (uint64_t)c * io.out_stride + o;
                              reinterpret_cast<uint32_t*>(io.out)[oc] = (reinterpret_cast<const uint32_t*>(io.in)[(uint64_t)c * io.in_stride + o] + mword) & mk32;
                          } else if (MODE == M_ENCODE) {
                              const float x = reinterpret_cast<const float*>(io.in)[(uint64_t)c * io.in_stride + o];
                              const double u = nz.u ? nz.u[(uint64_t)c * nz.u_stride + o] : noise_one<N32>(nz, nz.stream + c, j);
                              const Seg sg = find_seg(cd, j);
                              const uint32_t qv = encode_one(x, u, sg, cd.scale);
                              const uint64_t oc = (uint64_t)c * io.out_stride + o;
                              if (io.aux) reinterpret_cast<uint32_t*>(io.aux)[oc] = qv;
                              reinterpret_cast<uint32_t*>(io.out)[oc] = (qv + mword) & mk32;
                          } else if (MODE == M_DECODE) {
                              const uint32_t pw = (reinterpret_cast<const uint32_t*>(io.in)[o] + mword) & mk32;
                              if (io.aux) reinterpret_cast<uint32_t*>(io.aux)[o] = pw;
                              const Seg sg = find_seg(cd, j);
                              io.outf[o] = decode_one((double)pw, sg.two_an, cd.den, cd.den_rcp, sg.an);
                          }
                      }
                  }
              }
          }
          cached_win = win;                                          // every stream of the unit now has this window cached
        }
      };
      // Lane-batched word jw (offset o in the shard) of client c, given its combined mask:
      //   encode: the word's <= 8 lanes are quantised (jzf_quantize.py:55-67), packed first element most significant
      //           (:162-185; lanes past the layer's end are the zero padding), masked and stored;
      //   decode: the aggregate word is unmasked, split into its lanes (:234-251) and every real lane decoded (:102-107).
      auto codec_word = [&](uint32_t c, uint64_t jw, uint64_t o, u128 mask) {
        if constexpr (W4_CODEC) {
          const WordPos wp = locate_word(cd, jw);
          const uint32_t lb = cd.lane_bits;
          const u128 mk128 = Word<4>::mask(g.b);
          if (MODE == M_ENCODE) {
              const float* xin = reinterpret_cast<const float*>(io.in) + (uint64_t)c * io.in_stride;
              // The shipped geometry (jzf_quantize.py:162-185 with int_bits 120, 6 lanes of <= 21 bits), a word inside its
              // layer, device noise, element pairs aligned for the 8-byte loads and the noise generator: straight-line.
              const float* xp = xin + (wp.e0 - io.elem0);
              if (cd.bs == 6u && lb <= 21u && wp.e0 + 6u <= wp.eend && !nz.u && !io.aux && (wp.e0 & 1ull) == 0ull && ((uintptr_t)xp & 7u) == 0u) {
                  uint32_t xr[6];
                  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(xr[0]), "=r"(xr[1]) : "l"(xp));
                  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(xr[2]), "=r"(xr[3]) : "l"(xp + 2));
                  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(xr[4]), "=r"(xr[5]) : "l"(xp + 4));
                  double u[6];
                  const uint64_t p0 = wp.e0 >> 1;
                  if (N32) {                                        // two generator calls: the aligned quad and the pair beside it
                      double uq[4], ua, ub;
                      if ((p0 & 1ull) == 0ull) {
                          noise_quad<true>(nz, nz.stream + c, p0 >> 1, uq);
                          noise_pair<true>(nz, nz.stream + c, p0 + 2, ua, ub);
                          u[0] = uq[0]; u[1] = uq[1]; u[2] = uq[2]; u[3] = uq[3]; u[4] = ua; u[5] = ub;
                      } else {
                          noise_pair<true>(nz, nz.stream + c, p0, ua, ub);
                          noise_quad<true>(nz, nz.stream + c, (p0 + 1) >> 1, uq);
                          u[0] = ua; u[1] = ub; u[2] = uq[0]; u[3] = uq[1]; u[4] = uq[2]; u[5] = uq[3];
                      }
                  } else {
                      noise_pair<false>(nz, nz.stream + c, p0, u[0], u[1]);
                      noise_pair<false>(nz, nz.stream + c, p0 + 1, u[2], u[3]);
                      noise_pair<false>(nz, nz.stream + c, p0 + 2, u[4], u[5]);
                  }
                  uint32_t q[6];
#pragma unroll
                  for (int i = 0; i < 6; ++i) q[i] = encode_one(__uint_as_float(xr[i]), u[i], wp.sg, cd.scale);
                  // first element most significant: two halves of three lanes (3 lb <= 63 bits each)
                  const uint64_t H = ((uint64_t)q[0] << (2u * lb)) | ((uint64_t)q[1] << lb) | q[2];
                  const uint64_t Lw = ((uint64_t)q[3] << (2u * lb)) | ((uint64_t)q[4] << lb) | q[5];
                  u128 v; v.lo = Lw | (H << (3u * lb)); v.hi = H >> (64u - 3u * lb);
                  v = Word<4>::band(Word<4>::add(v, mask), mk128);
                  stg_v4(reinterpret_cast<u128*>(io.out) + (uint64_t)c * io.out_stride + o, (uint32_t)v.lo, (uint32_t)(v.lo >> 32), (uint32_t)v.hi, (uint32_t)(v.hi >> 32));
                  return;
              }
              uint64_t plo = 0ull, phi = 0ull, pc = ~0ull;
              double u0 = 0.0, u1 = 0.0;
              double uq[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
              for (uint32_t i = 0; i < cd.bs; ++i) {
                  const uint64_t e = wp.e0 + i;
                  uint32_t q = 0u;
                  if (e < wp.eend) {
                      double u;
                      if (nz.u) u = nz.u[(uint64_t)c * nz.u_stride + (e - io.elem0)];
                      else if (N32) {                               // 32-bit resolution: one generator call per four elements
                          if ((e >> 2) != pc) { pc = e >> 2; noise_quad<true>(nz, nz.stream + c, pc, uq); }
                          const uint32_t k = (uint32_t)e & 3u;
                          u = k == 0u ? uq[0] : (k == 1u ? uq[1] : (k == 2u ? uq[2] : uq[3]));
                      } else {
                          if ((e >> 1) != pc) { pc = e >> 1; noise_pair<false>(nz, nz.stream + c, pc, u0, u1); }
                          u = (e & 1ull) ? u1 : u0;
                      }
                      q = encode_one(xin[e - io.elem0], u, wp.sg, cd.scale);
                      if (io.aux) reinterpret_cast<uint32_t*>(io.aux)[(uint64_t)c * io.in_stride + (e - io.elem0)] = q;
                  }
                  phi = (phi << lb) | (plo >> (64u - lb));
                  plo = (plo << lb) | q;
              }
              u128 v; v.lo = plo; v.hi = phi;
              v = Word<4>::band(Word<4>::add(v, mask), mk128);
              stg_v4(reinterpret_cast<u128*>(io.out) + (uint64_t)c * io.out_stride + o, (uint32_t)v.lo, (uint32_t)(v.lo >> 32), (uint32_t)v.hi, (uint32_t)(v.hi >> 32));
          } else {
              uint32_t r0, r1, r2, r3;
              ldg_v4(reinterpret_cast<const u128*>(io.in) + o, r0, r1, r2, r3);
              u128 v; v.lo = ((uint64_t)r1 << 32) | r0; v.hi = ((uint64_t)r3 << 32) | r2;
              v = Word<4>::band(Word<4>::add(v, mask), mk128);
              if (io.aux) stg_v4(reinterpret_cast<u128*>(io.aux) + o, (uint32_t)v.lo, (uint32_t)(v.lo >> 32), (uint32_t)v.hi, (uint32_t)(v.hi >> 32));
              const uint64_t lm = (1ull << lb) - 1ull;
              double* op = io.outf + (wp.e0 - io.elem0);
              if (cd.bs == 6u && lb <= 21u && wp.e0 + 6u <= wp.eend && ((uintptr_t)op & 15u) == 0u) {   // the shipped geometry, straight-line
                  const uint64_t H = (v.lo >> (3u * lb)) | (v.hi << (64u - 3u * lb)), Lw = v.lo;      // lanes 0-2 (most significant) and 3-5
                  double d[6];
#pragma unroll
                  for (int i = 0; i < 3; ++i) {
                      d[i] = decode_one((double)((H >> ((2 - i) * lb)) & lm), wp.sg.two_an, cd.den, cd.den_rcp, wp.sg.an);
                      d[3 + i] = decode_one((double)((Lw >> ((2 - i) * lb)) & lm), wp.sg.two_an, cd.den, cd.den_rcp, wp.sg.an);
                  }
                  stg_d2(op, d[0], d[1]); stg_d2(op + 2, d[2], d[3]); stg_d2(op + 4, d[4], d[5]);
                  return;
              }
#pragma unroll 1
              for (int i = (int)cd.bs - 1; i >= 0; --i) {
                  const uint64_t e = wp.e0 + (uint32_t)i;
                  if (e < wp.eend) io.outf[e - io.elem0] = decode_one((double)(v.lo & lm), wp.sg.two_an, cd.den, cd.den_rcp, wp.sg.an);
                  v.lo = (v.lo >> lb) | (v.hi << (64u - lb));
                  v.hi >>= lb;
              }
          }
        }
      };
      // 16-byte words, m = 1: a lane's AES block masks exactly one word (words lane and lane + 32 of the item)
      auto fast_item_w4 = [&](uint64_t w) {
        if constexpr (W4_OK) {
          const uint64_t e0 = it.cb - it.off + (w << 6);            // first word (= first block) of the item
          const uint32_t ctr0 = (uint32_t)(it.cb - it.off) + ((uint32_t)w << 6);
          const uint64_t o0 = e0 - g.begin;
          const uint32_t ctrA = ctr0 + lane, ctrB = ctrA + 32u;
          const uint32_t win = ctr0 >> 8;
          const bool stale = !cache_ok || win != cached_win;
          const u128 mk128 = Word<4>::mask(g.b);
          const uint32_t c = c_first;                               // (SHARE exists for the encode mode only)
          uint32_t r[NB][4];
          if (MODE == M_APPLY) {
              const u128* in = reinterpret_cast<const u128*>(io.in) + (uint64_t)c * io.in_stride + o0 + lane;
              ldg_v4(in, r[0][0], r[0][1], r[0][2], r[0][3]);
              ldg_v4(in + 32, r[1][0], r[1][1], r[1][2], r[1][3]);
          }
          u128 acc[NB][1];
          acc[0][0] = Word<4>::zero(); acc[1][0] = Word<4>::zero();
          uint32_t s_begin, s_count;
          if (!st.batch) { s_begin = 0; s_count = st.n; }
          else { s_begin = c; s_count = st.dbl ? 2u : 1u; }
          for (uint32_t si = 0; si < s_count; ++si) {
              const uint32_t sidx = s_begin + si;
              const int sign = st.batch ? (si == 0 ? +1 : -1) : st.sign[sidx];
              WinC wc;
              const uint32_t slot = wcache + sidx * 16u;
              if (stale) {
                  wc = window_consts(ks, y, PRE_OF(sidx), ctr0);
                  if (cache_ok) {
                      if (lane == 0) asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(slot), "r"(wc.c0), "r"(wc.c1), "r"(wc.c2), "r"(wc.c3) : "memory");
                      __syncwarp();
                  }
              } else {
                  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(wc.c0), "=r"(wc.c1), "=r"(wc.c2), "=r"(wc.c3) : "r"(slot) : "memory");
              }
              uint32_t oa[4], ob[4];
              aes256_x2w(ks, y, st.pre[sidx][0], wc, ctrA, ctrB, oa, ob);
              accumulate_slots<4, 1>(oa, g.b, 1u, sign, acc[0]);
              accumulate_slots<4, 1>(ob, g.b, 1u, sign, acc[1]);
          }
#pragma unroll
          for (int h = 0; h < NB; ++h) {
              u128 v = acc[h][0];
              if (W4_CODEC) { codec_word(c, e0 + lane + 32u * h, o0 + lane + 32u * h, v); continue; }
              if (MODE == M_APPLY) {
                  u128 x;
                  x.lo = ((uint64_t)r[h][1] << 32) | r[h][0]; x.hi = ((uint64_t)r[h][3] << 32) | r[h][2];
                  v = Word<4>::add(x, v);
              }
              v = Word<4>::band(v, mk128);
              u128* out = reinterpret_cast<u128*>(io.out) + (uint64_t)c * io.out_stride + o0 + lane + 32u * h;
              stg_v4(out, (uint32_t)v.lo, (uint32_t)(v.lo >> 32), (uint32_t)v.hi, (uint32_t)(v.hi >> 32));
          }
          cached_win = win;
        }
      };
      // 8-byte words, m = 2: a lane's AES block masks one aligned pair of elements
      auto fast_item_w2 = [&](uint64_t w) {
        if constexpr (W2_OK) {
          const uint64_t e0 = it.cb - 2ull * it.off + (w << 7);     // first element of the item (even)
          const uint32_t ctr0 = (uint32_t)(it.cb - it.off) + ((uint32_t)w << 6);
          const uint64_t o0 = e0 - g.begin;
          const uint32_t ctrA = ctr0 + lane, ctrB = ctrA + 32u;
          const uint32_t win = ctr0 >> 8;
          const bool stale = !cache_ok || win != cached_win;
          const uint64_t mk64 = Word<2>::mask(g.b);
          const bool one_seg = cd.nseg == 1;
          uint64_t prev[NB][2];
          const uint32_t n_iter_here = SHARE ? n_iter : 1u;
          for (uint32_t cc = 0; cc < n_iter_here; ++cc) {
              const uint32_t c = SHARE ? (cc ? c_first + cc - 1 : 0) : c_first + cc;
              const bool emit = !SHARE || cc > 0;
              uint32_t r[NB][4];                                     // two 8-byte words, or two floats in r[h][0..1]
              if (HAS_IN && emit) {
                  if (MODE == M_ENCODE) {
                      const float* in = reinterpret_cast<const float*>(io.in) + (uint64_t)c * io.in_stride + o0 + 2u * lane;
                      asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r[0][0]), "=r"(r[0][1]) : "l"(in));
                      asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r[1][0]), "=r"(r[1][1]) : "l"(in + 64));
                  } else {
                      const uint64_t* in = reinterpret_cast<const uint64_t*>(io.in) + (uint64_t)c * io.in_stride + o0 + 2u * lane;
                      ldg_v4(in, r[0][0], r[0][1], r[0][2], r[0][3]);
                      ldg_v4(in + 64, r[1][0], r[1][1], r[1][2], r[1][3]);
                  }
              }
              uint64_t acc[NB][MMAX];
#pragma unroll
              for (int h = 0; h < NB; ++h)
#pragma unroll
                  for (int k = 0; k < MMAX; ++k) acc[h][k] = 0ull;
              uint32_t s_begin, s_count;
              if (!st.batch) { s_begin = 0; s_count = st.n; }
              else if (SHARE) { s_begin = cc; s_count = 1; }
              else { s_begin = c; s_count = st.dbl ? 2u : 1u; }
              for (uint32_t si = 0; si < s_count; ++si) {
                  const uint32_t sidx = s_begin + si;
                  const int sign = st.batch ? (si == 0 ? +1 : -1) : st.sign[sidx];
                  WinC wc;
                  const uint32_t slot = wcache + sidx * 16u;
                  if (stale) {
                      wc = window_consts(ks, y, PRE_OF(sidx), ctr0);
                      if (cache_ok) {
                          if (lane == 0) asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(slot), "r"(wc.c0), "r"(wc.c1), "r"(wc.c2), "r"(wc.c3) : "memory");
                          __syncwarp();
                      }
                  } else {
                      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(wc.c0), "=r"(wc.c1), "=r"(wc.c2), "=r"(wc.c3) : "r"(slot) : "memory");
                  }
                  uint32_t oa[4], ob[4];
                  aes256_x2w(ks, y, st.pre[sidx][0], wc, ctrA, ctrB, oa, ob);
                  accumulate_slots<2, MMAX>(oa, g.b, 2u, sign, acc[0]);
                  accumulate_slots<2, MMAX>(ob, g.b, 2u, sign, acc[1]);
              }
              if (SHARE) {
#pragma unroll
                  for (int h = 0; h < NB; ++h)
#pragma unroll
                      for (int k = 0; k < 2; ++k) {
                          const uint64_t cur = acc[h][k];
                          if (cc > 0) acc[h][k] = prev[h][k] - cur;
                          prev[h][k] = cur;
                      }
                  if (!emit) continue;
              }
#pragma unroll
              for (int h = 0; h < NB; ++h) {
                  const uint64_t o = o0 + 2u * lane + 64u * h;
                  const uint64_t j = e0 + 2u * lane + 64u * h;
                  const uint64_t m0 = acc[h][0], m1 = acc[h][1];
                  uint64_t w0, w1;                                   // the two output words
                  if (MODE == M_MASKS) {
                      w0 = m0 & mk64; w1 = m1 & mk64;
                  } else if (MODE == M_ENCODE) {
                      double u0, u1;
                      if (nz.u) { const double* up = nz.u + (uint64_t)c * nz.u_stride + o; u0 = up[0]; u1 = up[1]; }
                      else noise_pair<N32>(nz, nz.stream + c, j >> 1, u0, u1);
                      Seg sg = find_seg(cd, j);
                      const uint32_t q0 = encode_one(__uint_as_float(r[h][0]), u0, sg, cd.scale);
                      if (!one_seg && j + 1 >= sg.end) sg = find_seg(cd, j + 1);
                      const uint32_t q1 = encode_one(__uint_as_float(r[h][1]), u1, sg, cd.scale);
                      if (io.aux) {
                          uint32_t* qo = reinterpret_cast<uint32_t*>(io.aux) + (uint64_t)c * io.out_stride + o;
                          asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(qo), "r"(q0), "r"(q1) : "memory");
                      }
                      w0 = ((uint64_t)q0 + m0) & mk64; w1 = ((uint64_t)q1 + m1) & mk64;
                  } else {                                           // M_APPLY, M_DECODE
                      w0 = ((((uint64_t)r[h][1] << 32) | r[h][0]) + m0) & mk64;
                      w1 = ((((uint64_t)r[h][3] << 32) | r[h][2]) + m1) & mk64;
                  }
                  if (MODE == M_DECODE) {
                      Seg sg = find_seg(cd, j);
                      const double d0 = decode_one((double)w0, sg.two_an, cd.den, cd.den_rcp, sg.an);
                      if (!one_seg && j + 1 >= sg.end) sg = find_seg(cd, j + 1);
                      const double d1 = decode_one((double)w1, sg.two_an, cd.den, cd.den_rcp, sg.an);
                      if (io.aux) stg_v4(reinterpret_cast<uint64_t*>(io.aux) + o, (uint32_t)w0, (uint32_t)(w0 >> 32), (uint32_t)w1, (uint32_t)(w1 >> 32));
                      stg_d2(io.outf + o, d0, d1);
                  } else {
                      uint64_t* out = reinterpret_cast<uint64_t*>(io.out) + (uint64_t)c * io.out_stride + o;
                      stg_v4(out, (uint32_t)w0, (uint32_t)(w0 >> 32), (uint32_t)w1, (uint32_t)(w1 >> 32));
                  }
              }
          }
          cached_win = win;
        }
      };
#ifdef FLASHE_TRACE
      unsigned long long tr_e0 = 0ull;
      bool tr_in_edge = false;
#define TR_EDGE_CLOSE() do { if (tr_in_edge) { unsigned long long t__; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__)); tr_edge += t__ - tr_e0; tr_in_edge = false; } } while (0)
#else
#define TR_EDGE_CLOSE() do { } while (0)
#endif
      for (uint32_t sub = 0; sub < nsub; ++sub, ++it.w) {
        TR_EDGE_CLOSE();
#ifdef FLASHE_TRACE
        ++tr_items;
#endif
        if (QUAD_OK && it.w >= wf_lo && it.w < wf_hi) { fast_item(it.w); continue; }
        if (W2_OK && it.w >= wf_lo && it.w < wf_hi) { fast_item_w2(it.w); continue; }
        if (w4_here && it.w >= wf_lo && it.w < wf_hi) { fast_item_w4(it.w); continue; }
#ifdef FLASHE_TRACE
        ++tr_nedge; tr_in_edge = true;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_e0));
#endif
        const uint64_t blk0 = it.w ? it.w * ITEM_BLOCKS - it.off : 0;  // first block of the item
        const uint32_t nblk = it.w ? ITEM_BLOCKS : ITEM_BLOCKS - it.off;
        if (blk0 * m >= it.clen) break;                                // past the chunk's last item
        const uint64_t item_e0 = it.cb + blk0 * m;                     // first global element of the item
        const uint64_t rem = it.clen - blk0 * m;
        const uint32_t item_n = (uint32_t)(rem < (uint64_t)nblk * m ? rem : (uint64_t)nblk * m);
        if (item_e0 + item_n <= g.begin || item_e0 >= g.end) continue;   // item outside this shard
        const uint64_t blkA = blk0 + lane, blkB = blkA + 32;
        const bool onA = lane < nblk && blkA * m < it.clen, onB = lane + 32u < nblk && blkB * m < it.clen;
        const uint64_t ctr0 = it.cb + blk0;                            // jzf_flashe.py:34 "(i + begin)"
        const uint64_t ctrA = ctr0 + lane, ctrB = ctrA + 32;
        const bool fast = (((ctr0 + ITEM_BLOCKS - 1u) >> 32) == 0);  // hoisted round 1 needs word 2 == 0
        const uint32_t par = (uint32_t)(item_e0 & 1ull);
        const uint64_t base_e = item_e0 - par;                       // even; slab index = j - base_e
        const uint32_t npairs = (par + item_n + 1u) >> 1;
        // shard clipping in slab-index space
        const uint32_t lo_i = g.begin > item_e0 ? (uint32_t)(g.begin - base_e) : par;
        const uint32_t hi_i = (item_e0 + item_n) > g.end ? (uint32_t)(g.end > base_e ? g.end - base_e : 0) : par + item_n;
        const int64_t off0 = (int64_t)(base_e - g.begin);            // offset of slab index 0 in the shard buffers
        // one AES pass: F(iter, prf) for this lane's two blocks, accumulated with sign
        // (edge items and layouts without a lane-local path: the window terms are recomputed per call)
        auto stream_into = [&](uint32_t sidx, int sign, word_t (&acc)[NB][MMAX]) {
            uint32_t oa[4], ob[4];
            if (fast) {
                // (warp-uniform) EVERY lane runs both of its blocks, whether the chunk holds them or not - a lane without
                // block B would otherwise send the warp through the one-block routine as well (both sides of a divergent
                // branch: twice the lookups) - and, where this path only serves the edge items of a lane-local layout, the
                // ROLLED form of the rounds: edge items are rare, what they cost is instruction fetch (ncu on a launch of
                // edge items only: 67 % of the stall samples were no_inst)
                const WinC wc = window_consts(ks, y, PRE_OF(sidx), (uint32_t)ctr0);
                aes256_x2w<(QUAD_OK || W2_OK || W4_OK) ? 1 : FLASHE_AES_UNROLL>(ks, y, st.pre[sidx][0], wc, (uint32_t)ctrA, (uint32_t)ctrB, oa, ob);
                if (onA) accumulate_slots<WORDS, MMAX>(oa, g.b, m, sign, acc[0]);
                if (onB) accumulate_slots<WORDS, MMAX>(ob, g.b, m, sign, acc[1]);
            } else {
                const uint32_t prf = st.prf[sidx];
                if (onA) {
                    aes256_block_slow(ks, y, st.iter, prf, (uint32_t)(ctrA >> 32), (uint32_t)ctrA, PRE_OF(sidx), oa);
                    accumulate_slots<WORDS, MMAX>(oa, g.b, m, sign, acc[0]);
                }
                if (onB) {
                    aes256_block_slow(ks, y, st.iter, prf, (uint32_t)(ctrB >> 32), (uint32_t)ctrB, PRE_OF(sidx), ob);
                    accumulate_slots<WORDS, MMAX>(ob, g.b, m, sign, acc[1]);
                }
            }
        };

        // SHARE: iteration 0 only produces F(iter, first client); iteration cc >= 1 serves client cc-1
        // with F(c) - F(c+1), reusing F(c+1) as the next client's add term.
        word_t prev[SHARE ? NB : 1][SHARE ? MMAX : 1];
        for (uint32_t cc = 0; cc < n_iter; ++cc) {
            const uint32_t c = SHARE ? (cc ? c_first + cc - 1 : 0) : c_first + cc;
            const bool emit = !SHARE || cc > 0;
            // ---- 1. prefetch inputs (pairs) ----
            in_t pf[PF][2];
            if (HAS_IN && emit && !w4_batched) {
                const in_t* in = reinterpret_cast<const in_t*>(io.in) + (uint64_t)c * io.in_stride;
#pragma unroll
                for (int k = 0; k < PF; ++k) {
                    const uint32_t i0 = 2u * (lane + 32u * k);
                    if (i0 >= lo_i && i0 < hi_i) pf[k][0] = in[off0 + i0];
                    if (i0 + 1 >= lo_i && i0 + 1 < hi_i) pf[k][1] = in[off0 + i0 + 1];
                }
            }
            // ---- 2. keystreams ----
            word_t acc[NB][MMAX];
#pragma unroll
            for (int h = 0; h < NB; ++h)
#pragma unroll
                for (int k = 0; k < MMAX; ++k) acc[h][k] = WT::zero();
            {
                uint32_t s_begin, s_count;
                if (!st.batch) { s_begin = 0; s_count = st.n; }
                else if (SHARE) { s_begin = cc; s_count = 1; }
                else { s_begin = c; s_count = st.dbl ? 2u : 1u; }
                for (uint32_t s = 0; s < s_count; ++s) {
                    const int sign = st.batch ? (s == 0 ? +1 : -1) : st.sign[s_begin + s];
                    stream_into(s_begin + s, sign, acc);
                }
            }
            if (SHARE) {
                // acc = F(cc); mask of client cc-1 = prev - acc
#pragma unroll
                for (int h = 0; h < NB; ++h)
#pragma unroll
                    for (int k = 0; k < MMAX; ++k) {
                        const word_t cur = acc[h][k];
                        if (cc > 0) acc[h][k] = WT::sub(prev[SHARE ? h : 0][SHARE ? k : 0], cur);
                        prev[SHARE ? h : 0][SHARE ? k : 0] = cur;
                    }
                if (!emit) continue;
            }
            // ---- 3. lane-major -> element-major through the warp's slab ----
            __syncwarp();
#pragma unroll
            for (int h = 0; h < NB; ++h)
#pragma unroll
                for (int k = 0; k < MMAX; ++k)
                    if ((uint32_t)k < m) slab_store<WORDS>(sl(par + (lane + 32u * h) * m + k), acc[h][k]);
            __syncwarp();

            // ---- 4. element pairs ----
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                const uint32_t p = lane + 32u * k;
                if (p >= npairs) break;
                const uint32_t i0 = 2u * p;
                const bool v0 = i0 >= lo_i && i0 < hi_i, v1 = i0 + 1 >= lo_i && i0 + 1 < hi_i;
                if (!v0 && !v1) continue;
                const uint64_t j0 = base_e + i0;
                word_t mw0 = WT::band(slab_load<WORDS>(sl(i0)), mk);
                word_t mw1 = WT::band(slab_load<WORDS>(sl(i0 + 1)), mk);
                const int64_t o0 = off0 + i0;
                if constexpr (W4_CODEC) {
                    if (w4_batched) {                                  // lane-batched words of an edge item
                        if (v0) codec_word(c, j0, (uint64_t)o0, mw0);
                        if (v1) codec_word(c, j0 + 1, (uint64_t)(o0 + 1), mw1);
                        continue;
                    }
                }
                if (MODE == M_MASKS) {
                    word_t* out = reinterpret_cast<word_t*>(io.out) + (uint64_t)c * io.out_stride;
                    if (v0) out[o0] = mw0;
                    if (v1) out[o0 + 1] = mw1;
                } else if (MODE == M_APPLY) {
                    word_t* out = reinterpret_cast<word_t*>(io.out) + (uint64_t)c * io.out_stride;
                    if (v0) out[o0] = WT::band(WT::add(*reinterpret_cast<word_t*>(&pf[k][0]), mw0), mk);
                    if (v1) out[o0 + 1] = WT::band(WT::add(*reinterpret_cast<word_t*>(&pf[k][1]), mw1), mk);
                } else if (MODE == M_ENCODE) {
                    word_t* out = reinterpret_cast<word_t*>(io.out) + (uint64_t)c * io.out_stride;
                    double u0, u1;
                    if (nz.u) {
                        const double* up = nz.u + (uint64_t)c * nz.u_stride;
                        u0 = v0 ? up[o0] : 0.0; u1 = v1 ? up[o0 + 1] : 0.0;
                    } else {
                        noise_pair<N32>(nz, nz.stream + c, j0 >> 1, u0, u1);
                    }
                    if (v0) {
                        const Seg sg = find_seg(cd, j0);
                        uint32_t q = encode_one(*reinterpret_cast<float*>(&pf[k][0]), u0, sg, cd.scale);
                        if (io.aux) reinterpret_cast<uint32_t*>(io.aux)[(uint64_t)c * io.out_stride + o0] = q;
                        out[o0] = WT::band(WT::add(WT::from_u32(q), mw0), mk);
                    }
                    if (v1) {
                        const Seg sg = find_seg(cd, j0 + 1);
                        uint32_t q = encode_one(*reinterpret_cast<float*>(&pf[k][1]), u1, sg, cd.scale);
                        if (io.aux) reinterpret_cast<uint32_t*>(io.aux)[(uint64_t)c * io.out_stride + o0 + 1] = q;
                        out[o0 + 1] = WT::band(WT::add(WT::from_u32(q), mw1), mk);
                    }
                } else if (MODE == M_DECODE) {
                    if (v0) {
                        word_t pw = WT::band(WT::add(*reinterpret_cast<word_t*>(&pf[k][0]), mw0), mk);
                        if (io.aux) reinterpret_cast<word_t*>(io.aux)[o0] = pw;
                        const Seg sg = find_seg(cd, j0);
                        io.outf[o0] = decode_one(WT::to_double(pw), sg.two_an, cd.den, cd.den_rcp, sg.an);
                    }
                    if (v1) {
                        word_t pw = WT::band(WT::add(*reinterpret_cast<word_t*>(&pf[k][1]), mw1), mk);
                        if (io.aux) reinterpret_cast<word_t*>(io.aux)[o0 + 1] = pw;
                        const Seg sg = find_seg(cd, j0 + 1);
                        io.outf[o0 + 1] = decode_one(WT::to_double(pw), sg.two_an, cd.den, cd.den_rcp, sg.an);
                    }
                } else if (MODE == M_SCATTER) {
                    const int64_t* index = reinterpret_cast<const int64_t*>(io.aux);
                    word_t* dense = reinterpret_cast<word_t*>(io.out);
                    // indices come from other parties: anything outside [0, dense_len) is skipped (k_scatter does the same)
                    if (v0) { const uint64_t d = (uint64_t)index[o0]; if (d < io.dense_len) dense[d] = WT::band(WT::add(dense[d], mw0), mk); }
                    if (v1) { const uint64_t d = (uint64_t)index[o0 + 1]; if (d < io.dense_len) dense[d] = WT::band(WT::add(dense[d], mw1), mk); }
                }
            }
        }
      }
      TR_EDGE_CLOSE();
    }
    if (tk && lane == 0u) {
        // this warp has drawn its one ticket past the end; when every warp of the grid has, nobody touches the slot again
        if (atomicAdd(tk + 1, 1u) == gridDim.x * nwarps - 1u) { tk[0] = 0u; tk[1] = 0u; }
    }
#ifdef FLASHE_TRACE
    if (io.trace && lane == 0u) {
        unsigned long long tr_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_end));
        unsigned long long* rec = io.trace + 6ull * ((unsigned long long)blockIdx.x * nwarps + warp);
        rec[0] = tr_begin; rec[1] = tr_wait; rec[2] = tr_end; rec[3] = tr_edge; rec[4] = tr_items; rec[5] = tr_nedge;
    }
#endif
#undef TR_EDGE_CLOSE
#undef PRE_OF
}

#ifdef FLASHE_TRACE
#include <stdio.h>
#include <algorithm>
#include <vector>
static void flashe_trace_report(const unsigned long long* r, size_t n, int mode, int mmax, int aligned, int m6s, unsigned sup, unsigned long long units) {
    unsigned long long t0 = ~0ull, tw0 = ~0ull, tend = 0;
    std::vector<long long> ends, pro;
    for (size_t i = 0; i < n; ++i) { const unsigned long long* q = r + 6 * i; if (!q[0]) continue; t0 = std::min(t0, q[0]); tw0 = std::min(tw0, q[1]); tend = std::max(tend, q[2]); }
    unsigned long long items = 0, edges = 0, max_items = 0, max_edge = 0, max_edge_ns = 0, sum_edge_ns = 0;
    for (size_t i = 0; i < n; ++i) {
        const unsigned long long* q = r + 6 * i; if (!q[0]) continue;
        ends.push_back((long long)(q[2] - t0)); pro.push_back((long long)(q[1] - q[0]));
        items += q[4]; edges += q[5]; max_items = std::max(max_items, q[4]); max_edge = std::max(max_edge, q[5]); max_edge_ns = std::max(max_edge_ns, q[3]); sum_edge_ns += q[3];
    }
    std::sort(ends.begin(), ends.end()); std::sort(pro.begin(), pro.end());
    const size_t k = ends.size();
    if (!k) return;
    fprintf(stderr, "[trace] mode %d mmax %d aligned %d m6s %d sup %u units %llu | warps %zu items %llu (max/warp %llu) edge items %llu (max/warp %llu) | span %.1f us, prologue p50 %.1f max %.1f us, first wait-done at %.1f | warp end p0 %.1f p10 %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f us | edge time: mean per edge item %.2f us, max per warp %.1f us\n",
            mode, mmax, aligned, m6s, sup, units, k, items, max_items, edges, max_edge, (tend - t0) * 1e-3, pro[k / 2] * 1e-3, pro[k - 1] * 1e-3, (tw0 - t0) * 1e-3,
            ends[0] * 1e-3, ends[k / 10] * 1e-3, ends[k / 2] * 1e-3, ends[k * 9 / 10] * 1e-3, ends[k * 99 / 100] * 1e-3, ends[k - 1] * 1e-3,
            edges ? sum_edge_ns * 1e-3 / edges : 0.0, max_edge_ns * 1e-3);
    // the five last warps: items, edge items, edge time
    std::vector<size_t> idx; for (size_t i = 0; i < n; ++i) if (r[6 * i]) idx.push_back(i);
    std::sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return r[6 * a + 2] > r[6 * b + 2]; });
    for (size_t j = 0; j < 5 && j < idx.size(); ++j) { const unsigned long long* q = r + 6 * idx[j];
        fprintf(stderr, "[trace]   late warp cta %zu w %zu: end %.1f us items %llu edge %llu edge time %.1f us\n", idx[j] / 16, idx[j] % 16, (q[2] - t0) * 1e-3, q[4], q[5], q[3] * 1e-3); }
}
#endif

template <int WORDS, int MMAX, int MODE, bool SHARE, bool ALIGNED = false, bool N32 = false, bool M6S = false>
static int launch_stream_t(const flashe_ctx* ctx, const StreamTab& st, const Geom& g, const IoDev& io_in, const CodecDev& cd,
                           const NoiseDev& nz, cudaStream_t stream) {
    auto kern = k_stream<WORDS, MMAX, MODE, SHARE, ALIGNED, N32, M6S>;
    // the opt-in shared-memory size is a per-device property of the function: set it once per device
    static std::atomic<uint64_t> attr_set{0};
    const uint64_t dev_bit = 1ull << (ctx->device & 63);
    if (!(attr_set.load(std::memory_order_acquire) & dev_bit)) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        attr_set.fetch_or(dev_bit, std::memory_order_release);
    }
    const uint64_t items = (st.batch && !io_in.share) ? g.S_cnt * io_in.n_clients : g.S_cnt;
    if (items == 0) return FLASHE_OK;
    const int slab_bytes = (2 * 32 * MMAX + 2 + (WORDS == 1 ? 2 * MMAX + 1 : 0)) * WORDS * 4;
    int threads = STREAM_THREADS;
    while (threads > 32 && (threads / 32) * slab_bytes > 60 * 1024) threads >>= 1;
    const int wpb = threads / 32;
    uint64_t blocks = ceil_div(items, (uint64_t)wpb);
    if (blocks > (uint64_t)ctx->num_sms) blocks = (uint64_t)ctx->num_sms;
    // programmatic stream serialization: the grid may start (table fill) before the previous kernel has drained;
    // the kernel waits (griddepcontrol.wait) before it touches global data.  FLASHE_PDL=0 turns it off.
    static const bool pdl = [] { const char* e = getenv("FLASHE_PDL"); return !(e && e[0] == '0'); }();
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)blocks); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1u : 0u;
    IoDev io_t = io_in;
    // Dynamic dealing: the ticket (units x <= 8 pieces + one per warp) must fit 32 bits, and the launch must be long
    // enough for the re-balancing to pay for the draws (FLASHE_DYNAMIC_MIN_ITEMS warp items per warp; below that the
    // static deal is as good: with a handful of items per warp the tail is one item either way)
    static const uint64_t dyn_min_items = [] { const char* e = getenv("FLASHE_DYNAMIC_MIN_ITEMS"); return (uint64_t)(e ? atoi(e) : 6); }();
    const bool dyn_ok = items * 8u + blocks * (uint64_t)wpb < 0xfffffff0ull && items * g.sup >= dyn_min_items * blocks * (uint64_t)wpb;
    io_t.tickets = dyn_ok ? flashe_ticket_slot(ctx, stream) : nullptr;
    const IoDev& io = io_t;
#ifdef FLASHE_TRACE
    // tuning builds only: per-warp timeline of this launch, summarised on stderr (FLASHE_TRACE_PRINT=1)
    static const bool tr_print = [] { const char* e = getenv("FLASHE_TRACE_PRINT"); return e && e[0] == '1'; }();
    cudaStreamCaptureStatus tr_cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(stream, &tr_cap);
    if (tr_print && tr_cap == cudaStreamCaptureStatusNone) {
        static unsigned long long* tr_host = nullptr;
        const size_t n_rec = (size_t)blocks * wpb;
        if (!tr_host) CUDA_TRY(cudaHostAlloc((void**)&tr_host, 6 * sizeof(unsigned long long) * 148 * 16 * 2, cudaHostAllocMapped));
        memset(tr_host, 0, 6 * sizeof(unsigned long long) * n_rec);
        IoDev io2 = io;
        CUDA_TRY(cudaHostGetDevicePointer((void**)&io2.trace, tr_host, 0));
        CUDA_TRY(cudaStreamSynchronize(stream));
        CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, ctx->ks, st, g, io2, cd, nz));
        CUDA_TRY(cudaStreamSynchronize(stream));
        flashe_trace_report(tr_host, n_rec, MODE, MMAX, (int)ALIGNED, (int)M6S, (unsigned)g.sup, (unsigned long long)items);
        flashe_count_launches(1);
        return FLASHE_OK;
    }
#endif
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, ctx->ks, st, g, io, cd, nz));
    flashe_count_launches(1);
    CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}


// SHARED = true instantiates only the shared-stream encode kernels (their own translation unit)
template <int MODE, bool SHARED = false, bool N32 = false>
static int launch_stream(const flashe_ctx* ctx, const StreamTab& st, const Geom& g, const IoDev& io_in, const CodecDev& cd,
                         const NoiseDev& nz, cudaStream_t stream) {
    const int b = ctx->int_bits;
    IoDev io = io_in;
    io.te0 = ctx->d_te0;
    // 128-bit fast path preconditions (4-byte words): every row of every buffer starts 16-byte aligned
    // (16-byte words: every word is aligned as soon as the base pointers are)
    const bool batched = cd.bs != 0u && (MODE == M_ENCODE || MODE == M_DECODE);   // element-indexed pointers need no alignment
    if (batched) io.quad = (MODE == M_ENCODE ? aligned16(io.out) : (aligned16(io.in) && aligned16(io.aux))) ? 1u : 0u;
    else
    io.quad = (MODE != M_SCATTER && aligned16(io.in) && aligned16(io.out) && aligned16(io.aux) && aligned16(io.outf) &&
               ((ctx->words == 1 && (io.n_clients <= 1 || ((io.in_stride | io.out_stride) & 3ull) == 0)) ||
                (ctx->words == 2 && (io.n_clients <= 1 || ((io.in_stride | io.out_stride) & 1ull) == 0)) || ctx->words == 4)) ? 1u : 0u;
    if constexpr (MODE == M_ENCODE && SHARED) {
        {
            if (b <= 32) {
                if (ctx->m == 4 && g.aligned4 && io.quad) return launch_stream_t<1, 4, MODE, true, true, N32>(ctx, st, g, io, cd, nz, stream);
                if (ctx->m <= 4) return launch_stream_t<1, 4, MODE, true, false, N32>(ctx, st, g, io, cd, nz, stream);
                if (ctx->m == 6 && g.aligned4 && io.quad) return launch_stream_t<1, 6, MODE, true, true, N32>(ctx, st, g, io, cd, nz, stream);
                if (ctx->m <= 6) return launch_stream_t<1, 6, MODE, true, false, N32>(ctx, st, g, io, cd, nz, stream);
                return launch_stream_t<1, 16, MODE, true, false, N32>(ctx, st, g, io, cd, nz, stream);
            }
            if (b <= 64) return launch_stream_t<2, 3, MODE, true, false, N32>(ctx, st, g, io, cd, nz, stream);
            return launch_stream_t<4, 1, MODE, true, false, N32>(ctx, st, g, io, cd, nz, stream);
        }
    } else {
    if (b <= 32) {
        if (ctx->m == 4 && g.aligned4 && io.quad && MODE != M_SCATTER) return launch_stream_t<1, 4, MODE, false, true, N32>(ctx, st, g, io, cd, nz, stream);
        if (ctx->m <= 4) return launch_stream_t<1, 4, MODE, false, false, N32>(ctx, st, g, io, cd, nz, stream);
        if (ctx->m == 6 && g.aligned4 && io.quad && MODE != M_SCATTER) return launch_stream_t<1, 6, MODE, false, true, N32>(ctx, st, g, io, cd, nz, stream);
        if (ctx->m == 6 && io.quad && MODE != M_SCATTER) return launch_stream_t<1, 6, MODE, false, false, N32, true>(ctx, st, g, io, cd, nz, stream);   // any chunk alignment
        if (ctx->m <= 6) return launch_stream_t<1, 6, MODE, false, false, N32>(ctx, st, g, io, cd, nz, stream);
        return launch_stream_t<1, 16, MODE, false, false, N32>(ctx, st, g, io, cd, nz, stream);
    }
    if (b <= 64) return launch_stream_t<2, 3, MODE, false, false, N32>(ctx, st, g, io, cd, nz, stream);
    return launch_stream_t<4, 1, MODE, false, false, N32>(ctx, st, g, io, cd, nz, stream);
    }
}


#endif
