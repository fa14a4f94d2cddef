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
#include <string>
#include <vector>

#include "../../include/flashe_b200.h"
#include "flashe_internal.h"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const std::string& msg) { g_err = msg; return code; }
int flashe_fail(int code, const std::string& msg) { return fail(code, msg); }
void flashe_count_launches(int n) { g_launches.fetch_add((uint64_t)n); }
#define CUDA_TRY(expr)                                                                             \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(FLASHE_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));        \
    } while (0)

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

// ------------------------------------------------------------------------------------------------
// kernel parameter blocks (all in the constant bank)
// ------------------------------------------------------------------------------------------------
#define MAXS FLASHE_MAX_STREAMS
#define MAX_INLINE_SEG 48
#ifndef STREAM_THREADS
#define STREAM_THREADS 512
#endif

struct alignas(16) KeySched { uint32_t rk[60]; };

struct StreamTab {
    uint32_t n;           // entries
    uint32_t iter;
    uint32_t batch;       // 0: entries are the stream list of the single vector
                          // 1: client c uses entry c (+) and, when dbl, entry c+1 (-)
    uint32_t dbl;
    uint32_t prf[MAXS];
    int32_t sign[MAXS];
    uint32_t pre[MAXS][4];
};

struct Geom {
    uint64_t L, begin, end;  // whole length, shard [begin,end)
    uint64_t d, r;           // divmod(L, n_jobs): first r chunks have d+1 elements
    uint64_t nwA, nwB;       // warp items per chunk (types: d+1 / d elements)
    uint64_t nsA, nsB;       // work units per chunk: ceil(nw / sup)
    uint64_t rSA;            // r * nsA
    uint64_t S_lo, S_cnt;    // work units that intersect the shard
    uint32_t m, b;           // slots per AES block, int_bits
    uint32_t sup;            // warp items per work unit (consecutive items of one chunk)
    uint32_t aligned4;       // every chunk begin and the shard begin are multiples of 4 elements
};

struct Seg { uint64_t end; float a, two_a; float rcp_two_a, pad; double an, two_an; };  // rcp_two_a = RN(1/two_a), 0 = not usable

struct CodecDev {
    int32_t nseg;
    int32_t ebits;
    float scale;             // 2^e - 1 as float32
    double den;              // (2^e - 1) * n as float64
    double den_rcp;          // RN(1 / den), or 0: use the library division
    const Seg* table;        // device table when nseg > MAX_INLINE_SEG, else NULL
    Seg seg[MAX_INLINE_SEG];
};

struct NoiseDev { const double* u; uint64_t u_stride; uint64_t stream; uint32_t rk[10][2]; };  // rk: Philox round keys

struct IoDev {
    const void* in;  uint64_t in_stride;    // words (or floats) between consecutive clients
    void* out;       uint64_t out_stride;
    void* aux;                              // q_out (encode) / p_out (decode) / index (scatter)
    double* outf;
    uint32_t n_clients;
    uint32_t share;                         // batch double masking: compute each stream once
    uint32_t quad;                          // every buffer / stride allows 16-byte accesses per 4 elements
    uint64_t dense_len;                     // scatter: words in the dense target (indices outside are skipped)
};

enum { M_MASKS = 0, M_APPLY = 1, M_ENCODE = 2, M_DECODE = 3, M_SCATTER = 4 };

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
#define SMEM_BYTES 0x30000u  // requested dynamic shared memory: covers [base, 0x30000) for base <= 0x10000

__device__ uint32_t g_te0[256];  // filled once per process by flashe_ctx_create

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

__device__ __forceinline__ void fill_tables() {
    // word w of the 128 KB region: region = w>>14, entry = (w>>6)&255, table-in-region = (w>>5)&1
    for (uint32_t w = threadIdx.x; w < 32768u; w += blockDim.x) {
        uint32_t t = ((w >> 14) << 1) | ((w >> 5) & 1u);
        uint32_t v = g_te0[(w >> 6) & 255u];
        v = __funnelshift_r(v, v, 8 * t);  // Te_t = ror(Te0, 8t)
        sts32(TAB_BASE + 4u * w, v);
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
#define T0(s) lds_tab<0>(addr_b3((s), y))
#define T1(s) lds_tab<128>(__byte_perm((s), y, SEL_B2))
#define T2(s) lds_tab<0x10000>(__byte_perm((s), y, SEL_B1))
#define T3(s) lds_tab<0x10080>(__byte_perm((s), y, SEL_B0))

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
#define ITEM_BLOCKS 64u   // AES blocks per warp item (two per lane)
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

// ------------------------------------------------------------------------------------------------
// device: encode / decode / noise
// ------------------------------------------------------------------------------------------------
// Layer parameters of element j, BY VALUE: the single-layer case reads the constant bank directly,
// an inline table is searched in the constant bank, a large one in global memory (a reference return
// would force generic loads for all three).
__device__ __forceinline__ Seg find_seg(const CodecDev& c, uint64_t j) {
    if (c.nseg == 1) return c.seg[0];
    int lo = 0, hi = c.nseg - 1;
    if (c.table) {
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (j < __ldg(&c.table[mid].end)) hi = mid; else lo = mid + 1;
        }
        Seg r; const Seg* t = c.table + lo;
        r.end = __ldg(&t->end); r.a = __ldg(&t->a); r.two_a = __ldg(&t->two_a); r.rcp_two_a = __ldg(&t->rcp_two_a); r.pad = 0.f;
        r.an = __ldg(&t->an); r.two_an = __ldg(&t->two_an);
        return r;
    }
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (j < c.seg[mid].end) hi = mid; else lo = mid + 1;
    }
    return c.seg[lo];
}

// _static_quantize_padding_asymmetric, jzf_quantize.py:55-67: four float32 ops in the reference's
// order (no FMA contraction), then float64 add of the noise, floor, int.
// IEEE-754 round-to-nearest v / d for a divisor whose correctly rounded reciprocal y = RN(1/d) is known
// (the host computes it exactly).  q0 = RN(v*y) is within 1.5 ulp of v/d; one FMA residual correction
// makes it faithful, and Markstein's theorem (y correctly rounded, q faithful) makes the second
// correction the correctly rounded quotient (tests/native/div_rcp_check.c sweeps it on the CPU).
// The residuals are exact only without underflow.  The host offers y only for alpha in [2^-41, 2^59];
// the numerator v = fl(fl(clip(x) + alpha) * (2^e - 1)) is then 0 or >= alpha * 2^-24 * (2^e - 1) > 2^-60
// (x + alpha is 0 or at least half an ulp of alpha), so no element needs a range check; other
// alphas are flagged by y == 0 and take the library division.
__device__ __forceinline__ float div_rn_known_rcp(float v, float d, float y) {
    if (y == 0.0f) return __fdiv_rn(v, d);
    float q = __fmul_rn(v, y);
    float r = __fmaf_rn(-d, q, v);
    q = __fmaf_rn(r, y, q);
    r = __fmaf_rn(-d, q, v);
    return __fmaf_rn(r, y, q);
}

// RCP = true: the caller has checked sg.rcp_two_a != 0 (one test per layer run instead of one per element).
template <bool RCP = false>
__device__ __forceinline__ uint32_t encode_one(float x, double u, const Seg& sg, float scale) {
    float v = fminf(fmaxf(x, -sg.a), sg.a);
    v = __fadd_rn(v, sg.a);
    v = __fmul_rn(v, scale);
    if (RCP) {
        const float d = sg.two_a, y = sg.rcp_two_a;
        float q = __fmul_rn(v, y);
        float r = __fmaf_rn(-d, q, v);
        q = __fmaf_rn(r, y, q);
        r = __fmaf_rn(-d, q, v);
        v = __fmaf_rn(r, y, q);
    } else {
        v = div_rn_known_rcp(v, sg.two_a, sg.rcp_two_a);
    }
    // floor(t) for 0 <= t < 2^32: t + 2^52 rounded towards -inf lands on the integer grid at
    // 2^52 + floor(t); the integer is the low word of that double.  (v >= 0 by construction.)
    const double t = __dadd_rn((double)v, u);
    return (uint32_t)__double2loint(__dadd_rd(t, 4503599627370496.0));
}

// _static_unquantize_padding_asymmetric, jzf_quantize.py:102-107 (float64, left to right).
// The division uses the host-computed y = RN(1/den) when the host offers it (den_rcp != 0): q0 = RN(n*y)
// is within 2 ulp of n/den, one FMA residual correction makes it faithful and, y being correctly rounded,
// Markstein's theorem makes the second one the IEEE quotient (tests/native/ddiv_rcp_check.c sweeps it on
// the CPU, 2^28 cases).  The host withholds y when a layer's 2*alpha*n lies outside [2^-400, 2^400]
// (the residuals must not underflow).  Six float64 operations instead of the ~30 of the library division.
__device__ __forceinline__ double ddiv_rn_known_rcp(double n, double d, double y) {
    if (y == 0.0) return __ddiv_rn(n, d);
    double q = __dmul_rn(n, y);
    double r = __fma_rn(-d, q, n);
    q = __fma_rn(r, y, q);
    r = __fma_rn(-d, q, n);
    return __fma_rn(r, y, q);
}
__device__ __forceinline__ double decode_one(double v, double two_an, double den, double den_rcp, double an) {
    return __dsub_rn(ddiv_rn_known_rcp(__dmul_rn(v, two_an), den, den_rcp), an);
}

// Philox4x32-10 (Salmon et al. 2011), counter (c0,c1,c2,c3); the ten round keys
// (k0 + i*0x9E3779B9, k1 + i*0xBB67AE85) are expanded on the host (NoiseDev.rk).
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const NoiseDev& nz,
                                              uint32_t out[4]) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ nz.rk[i][0], n2 = (uint32_t)(p0 >> 32) ^ c3 ^ nz.rk[i][1];
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// numpy's res53 construction ((a>>5)*2^26 + (b>>6)) / 2^53 without integer->double conversions:
// A = 2^25 + (a>>5)*2^-27 and B = 2^-1 + (b>>6)*2^-53 are assembled as bit patterns (exponent word |
// mantissa low word); A - (2^25 + 2^-1) and the sum with B are exact, so the value is bit-identical.
__device__ __forceinline__ double res53(uint32_t a, uint32_t b) {
    const double A = __hiloint2double(0x41800000, (int)(a >> 5));
    const double B = __hiloint2double(0x3FE00000, (int)(b >> 6));
    return __dadd_rn(__dadd_rn(A, -33554432.5), B);
}
// u_j in [0,1): counter (j>>1 lo, j>>1 hi, stream lo, stream hi); words (2(j&1), 2(j&1)+1) feed res53.
__device__ __forceinline__ double noise_one(const NoiseDev& nz, uint64_t stream, uint64_t j) {
    uint32_t o[4];
    uint64_t c = j >> 1;
    philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)stream, (uint32_t)(stream >> 32), nz, o);
    return (j & 1) ? res53(o[2], o[3]) : res53(o[0], o[1]);
}

// Both numbers of one Philox call: u for elements 2c and 2c+1 (same values as noise_one).
__device__ __forceinline__ void noise_pair(const NoiseDev& nz, uint64_t stream, uint64_t c, double& u0, double& u1) {
    uint32_t o[4];
    philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)stream, (uint32_t)(stream >> 32), nz, o);
    u0 = res53(o[0], o[1]);
    u1 = res53(o[2], o[3]);
}

// ------------------------------------------------------------------------------------------------
// device: word arithmetic for the three storage widths
// ------------------------------------------------------------------------------------------------
template <int WORDS> struct Word;
template <> struct Word<1> {
    typedef uint32_t T;
    static __host__ __device__ __forceinline__ T mask(uint32_t b) { return b >= 32 ? 0xffffffffu : ((1u << b) - 1u); }
    static __host__ __device__ __forceinline__ T add(T a, T b) { return a + b; }
    static __host__ __device__ __forceinline__ T sub(T a, T b) { return a - b; }
    static __host__ __device__ __forceinline__ T band(T a, T m) { return a & m; }
    static __host__ __device__ __forceinline__ T from_u32(uint32_t q) { return q; }
    static __host__ __device__ __forceinline__ double to_double(T a) { return (double)a; }
    static __host__ __device__ __forceinline__ T zero() { return 0u; }
};
template <> struct Word<2> {
    typedef uint64_t T;
    static __host__ __device__ __forceinline__ T mask(uint32_t b) { return b >= 64 ? ~0ull : ((1ull << b) - 1ull); }
    static __host__ __device__ __forceinline__ T add(T a, T b) { return a + b; }
    static __host__ __device__ __forceinline__ T sub(T a, T b) { return a - b; }
    static __host__ __device__ __forceinline__ T band(T a, T m) { return a & m; }
    static __host__ __device__ __forceinline__ T from_u32(uint32_t q) { return q; }
    static __host__ __device__ __forceinline__ double to_double(T a) { return (double)a; }
    static __host__ __device__ __forceinline__ T zero() { return 0ull; }
};
struct alignas(16) u128 { uint64_t lo, hi; };
template <> struct Word<4> {
    typedef u128 T;
    static __host__ __device__ __forceinline__ T mask(uint32_t b) {
        T m; m.lo = ~0ull; m.hi = b >= 128 ? ~0ull : ((1ull << (b - 64)) - 1ull); return m;
    }
    static __host__ __device__ __forceinline__ T add(T a, T b) { T r; r.lo = a.lo + b.lo; r.hi = a.hi + b.hi + (r.lo < a.lo); return r; }
    static __host__ __device__ __forceinline__ T sub(T a, T b) { T r; r.lo = a.lo - b.lo; r.hi = a.hi - b.hi - (a.lo < b.lo); return r; }
    static __host__ __device__ __forceinline__ T band(T a, T m) { T r; r.lo = a.lo & m.lo; r.hi = a.hi & m.hi; return r; }
    static __host__ __device__ __forceinline__ T from_u32(uint32_t q) { T r; r.lo = q; r.hi = 0; return r; }
    static __host__ __device__ __forceinline__ double to_double(T a) { return (double)a.lo; }
    static __host__ __device__ __forceinline__ T zero() { T r; r.lo = 0; r.hi = 0; return r; }
};

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

__device__ __forceinline__ void ldg_v4(const void* p, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
}
__device__ __forceinline__ void stg_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void stg_d2(double* p, double a, double b) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
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
    if ((r & 1u) == 0u) {
        stg_d2(p, v[0], v[1]);
        stg_d2(p + 2, v[2], v[3]);
    } else {
        asm volatile("st.global.f64 [%0], %1;" ::"l"(p), "d"(v[0]) : "memory");
        stg_d2(p + 1, v[1], v[2]);
        asm volatile("st.global.f64 [%0], %1;" ::"l"(p + 3), "d"(v[3]) : "memory");
    }
}

// Out-of-line single block for the rare paths (chunk tails, counters >= 2^32).
__device__ __noinline__ void aes256_block_slow(const KeySched& ks, uint32_t y, uint32_t w0, uint32_t w1, uint32_t w2,
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
template <int WORDS, int MMAX, int MODE, bool SHARE, bool ALIGNED>
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
    constexpr bool W4_OK = (WORDS == 4 && (MODE == M_MASKS || MODE == M_APPLY));   // 16-byte words (the shipped 120-bit batch mode): m = 1
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

    fill_tables();
    __syncthreads();

    // slab word index -> byte offset; 4-byte words are skewed by one word per 32 so that the
    // lane-major stores (stride m) and the element-major loads never pile onto one bank
    auto sl = [&](uint32_t i) -> uint32_t { return slab + (WORDS == 1 ? (i + (i >> 5)) : i) * WB; };

    const word_t mk = WT::mask(g.b);
    const uint32_t m = g.m;
    const uint64_t n_units = (st.batch && !io.share) ? g.S_cnt * io.n_clients : g.S_cnt;
    const uint64_t gw = (uint64_t)blockIdx.x * nwarps + warp, gstride = (uint64_t)gridDim.x * nwarps;

    for (uint64_t t = gw; t < n_units; t += gstride) {
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
      uint32_t cached_win = 0xffffffffu;                             // counter window the cached round-2 terms belong to
      const uint32_t n_iter = SHARE ? c_count + 1 : c_count;
      // ---- lane-local items ------------------------------------------------------------------------
      // Items [wf_lo, wf_hi) of the unit's chunk are FULL (64 blocks of m = 4 elements), lie inside the
      // shard and have 32-bit counters: a lane's AES block IS four consecutive elements, so the lane
      // loads / stores them itself (a warp covers 512 contiguous bytes per access), nothing goes
      // through the slab, and all per-item geometry is a handful of additions.  The bounds are
      // computed once per unit.
      uint64_t wf_lo = 1, wf_hi = 0;
      const uint32_t mm = (WORDS == 1 && MMAX == 4) ? 4u : (WORDS == 4 ? 1u : m);   // compile-time where the instantiation fixes it
      const uint32_t item_elems = ITEM_BLOCKS * mm;
      // (8-byte words: only m = 2, and only chunks that start on an even element of an even shard, so that a
      //  block is one aligned 16-byte pair and one noise pair)
      const bool w2_here = W2_OK && m == 2u && ((it.cb | g.begin) & 1ull) == 0ull;
      if ((QUAD_OK || W4_OK || w2_here) && io.quad && !(SHIFT_OK && (g.begin & 1ull))) {
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
          const bool shifted = SHIFT_OK && qr0 != 0u;
          const bool run_first = shifted && w == wA, run_last = shifted && w + 1 == wBx;
          // alignment switch of the 16-byte accesses: with SHIFT_OK every quad is aligned (qr0 != 0 => shifted),
          // which removes the 64/32-bit piece code from this instantiation's hot loop
          const uint32_t qr = SHIFT_OK ? 0u : qr0;
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
                      if ((MMAX == 4 || q < nquads) && !(run_first && q == 0u)) ldg_quad(in + 4u * q, qr, r[k]);
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
                  aes256_x2w<(MMAX == 4 ? FLASHE_AES_UNROLL : FLASHE_AES_UNROLL_M6)>(ks, y, st.pre[sidx][0], wc, ctrA, ctrB, oa, ob);
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
#pragma unroll
                  for (int h = 0; h < NB; ++h) {
                      const uint32_t a0 = fslab + (lane + 32u * h) * mm * 4u;
                      if (mm == 6u) {
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
                  if (MMAX != 4 && q >= nquads) break;
                  if (SHIFT_OK && run_first && q == 0u) continue;    // partial quad: handled element-wise below
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
                  if (MODE == M_MASKS) {
                      stg_quad(reinterpret_cast<uint32_t*>(io.out) + o, qr, mw[0] & mk32, mw[1] & mk32, mw[2] & mk32, mw[3] & mk32);
                  } else if (MODE == M_APPLY) {
                      uint32_t* out = reinterpret_cast<uint32_t*>(io.out) + (uint64_t)c * io.out_stride + o;
                      stg_quad(out, qr, (r[h][0] + mw[0]) & mk32, (r[h][1] + mw[1]) & mk32, (r[h][2] + mw[2]) & mk32, (r[h][3] + mw[3]) & mk32);
                  } else if (MODE == M_ENCODE) {
                      double u[4];
                      if (nz.u) {
                          const double* up = nz.u + (uint64_t)c * nz.u_stride + o;
#pragma unroll
                          for (int k = 0; k < 4; ++k) u[k] = up[k];
                      } else if (ALIGNED || SHIFT_OK || (j & 1ull) == 0ull) {   // (SHIFT_OK: quads start at begin + 4k, begin even)
                          noise_pair(nz, nz.stream + c, j >> 1, u[0], u[1]);
                          noise_pair(nz, nz.stream + c, (j >> 1) + 1, u[2], u[3]);
                      } else {                      // odd chunk start: the four elements touch three pairs
                          double lo, hi;
                          noise_pair(nz, nz.stream + c, j >> 1, lo, u[0]);
                          noise_pair(nz, nz.stream + c, (j >> 1) + 1, u[1], u[2]);
                          noise_pair(nz, nz.stream + c, (j >> 1) + 2, u[3], hi);
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
                              reinterpret_cast<uint32_t*>(io.out)[o] = mword & mk32;
                          } else if (MODE == M_APPLY) {
                              const uint64_t oc = (uint64_t)c * io.out_stride + o;
                              reinterpret_cast<uint32_t*>(io.out)[oc] = (reinterpret_cast<const uint32_t*>(io.in)[(uint64_t)c * io.in_stride + o] + mword) & mk32;
                          } else if (MODE == M_ENCODE) {
                              const float x = reinterpret_cast<const float*>(io.in)[(uint64_t)c * io.in_stride + o];
                              const double u = nz.u ? nz.u[(uint64_t)c * nz.u_stride + o] : noise_one(nz, nz.stream + c, j);
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
          if (HAS_IN) {
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
              if (MODE == M_APPLY) {
                  u128 x;
                  x.lo = ((uint64_t)r[h][1] << 32) | r[h][0]; x.hi = ((uint64_t)r[h][3] << 32) | r[h][2];
                  v = Word<4>::add(x, v);
              }
              v = Word<4>::band(v, mk128);
              u128* out = reinterpret_cast<u128*>(io.out) + (MODE == M_APPLY ? (uint64_t)c * io.out_stride : 0ull) + o0 + lane + 32u * h;
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
                      else noise_pair(nz, nz.stream + c, j >> 1, u0, u1);
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
                      uint64_t* out = reinterpret_cast<uint64_t*>(io.out) + (MODE == M_MASKS ? 0ull : (uint64_t)c * io.out_stride) + o;
                      stg_v4(out, (uint32_t)w0, (uint32_t)(w0 >> 32), (uint32_t)w1, (uint32_t)(w1 >> 32));
                  }
              }
          }
          cached_win = win;
        }
      };
      for (uint32_t sub = 0; sub < nsub; ++sub, ++it.w) {
        if (QUAD_OK && it.w >= wf_lo && it.w < wf_hi) { fast_item(it.w); continue; }
        if (W2_OK && it.w >= wf_lo && it.w < wf_hi) { fast_item_w2(it.w); continue; }
        if (W4_OK && it.w >= wf_lo && it.w < wf_hi) { fast_item_w4(it.w); continue; }
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
            WinC wc = {0u, 0u, 0u, 0u};
            if (fast) wc = window_consts(ks, y, PRE_OF(sidx), (uint32_t)ctr0);
            if (!onA) return;
            if (onB && fast) {
                aes256_x2w(ks, y, st.pre[sidx][0], wc, (uint32_t)ctrA, (uint32_t)ctrB, oa, ob);
                accumulate_slots<WORDS, MMAX>(oa, g.b, m, sign, acc[0]);
                accumulate_slots<WORDS, MMAX>(ob, g.b, m, sign, acc[1]);
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
            if (HAS_IN && emit) {
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
                if (MODE == M_MASKS) {
                    word_t* out = reinterpret_cast<word_t*>(io.out);
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
                        noise_pair(nz, nz.stream + c, j0 >> 1, u0, u1);
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
    }
#undef PRE_OF
}

// one AES block, known-answer tests (flashe_prp_block)
__global__ void k_prp_block(const __grid_constant__ KeySched ks, const uint32_t* in, uint32_t* out) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t y = 0x00010000u | (lane << 2);
    if (smem_window_base() > TAB_BASE) { __trap(); }
    fill_tables();
    __syncthreads();
    if (threadIdx.x < 32) {
        Pre pre; uint32_t o[4];
        // w2 forced non-zero path is not wanted here: hoist on the device for this block
        uint32_t w0 = in[0], w1 = in[1], w2 = in[2], w3 = in[3];
        uint32_t s0 = w0 ^ ks.rk[0], s1 = w1 ^ ks.rk[1], s2 = w2 ^ ks.rk[2];
        pre.p0 = T0(s0) ^ T1(s1) ^ T2(s2) ^ ks.rk[4];
        pre.p1 = T0(s1) ^ T1(s2) ^ T3(s0) ^ ks.rk[5];
        pre.p2 = T0(s2) ^ T2(s0) ^ T3(s1) ^ ks.rk[6];
        pre.p3 = T1(s0) ^ T2(s1) ^ T3(s2) ^ ks.rk[7];
        aes256_block(ks, y, w0, w1, 0u, w3, pre, o);   // fast path with the hoisted terms
        uint32_t o2[4];
        aes256_block(ks, y, w0, w1, w2 | 0u, w3, pre, o2);  // generic path when w2 != 0
        if (lane == 0) {
            for (int i = 0; i < 4; ++i) { out[i] = o[i]; out[4 + i] = o2[i]; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// elementwise kernels
// ------------------------------------------------------------------------------------------------
template <int WORDS>
__global__ void k_add_premasked(const typename Word<WORDS>::T* __restrict__ in, const typename Word<WORDS>::T* __restrict__ mask,
                                int sign, uint64_t count, uint32_t b, typename Word<WORDS>::T* __restrict__ out) {
    typedef Word<WORDS> WT;
    const typename WT::T mk = WT::mask(b);
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += (uint64_t)gridDim.x * blockDim.x) {
        typename WT::T a = in[j], mkv = mask[j];
        out[j] = WT::band(sign >= 0 ? WT::add(a, mkv) : WT::sub(a, mkv), mk);
    }
}

// vectorised u32 specialisation: 4 elements per thread, 128-bit accesses
__global__ void k_add_premasked_v4(const uint4* __restrict__ in, const uint4* __restrict__ mask, int sign, uint64_t nvec,
                                   uint32_t mk, uint4* __restrict__ out) {
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (uint64_t)gridDim.x * blockDim.x) {
        uint4 a = __ldg(in + v), k = __ldg(mask + v), r;
        if (sign >= 0) { r.x = a.x + k.x; r.y = a.y + k.y; r.z = a.z + k.z; r.w = a.w + k.w; }
        else { r.x = a.x - k.x; r.y = a.y - k.y; r.z = a.z - k.z; r.w = a.w - k.w; }
        r.x &= mk; r.y &= mk; r.z &= mk; r.w &= mk;
        out[v] = r;
    }
}

template <int WORDS, bool WITH_MASK>
__global__ void k_encode(const float* __restrict__ x, const typename Word<WORDS>::T* __restrict__ mask, uint64_t begin,
                         uint64_t count, uint32_t b, const __grid_constant__ CodecDev cd, const __grid_constant__ NoiseDev nz,
                         uint32_t* __restrict__ q_out, typename Word<WORDS>::T* __restrict__ ct_out) {
    typedef Word<WORDS> WT;
    const typename WT::T mk = WT::mask(b);
    for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < count; o += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t j = begin + o;
        const Seg sg = find_seg(cd, j);
        double u = nz.u ? nz.u[o] : noise_one(nz, nz.stream, j);
        uint32_t q = encode_one(x[o], u, sg, cd.scale);
        if (q_out) q_out[o] = q;
        if (WITH_MASK) ct_out[o] = WT::band(WT::add(WT::from_u32(q), mask[o]), mk);
    }
}

// Online step after mask precomputation, 4-byte words: one thread = 4 consecutive elements (begin and
// every pointer 16-byte aligned), 128-bit loads of x and of the precomputed mask, two Philox calls for
// the four noise values, one 128-bit store.  12 algorithmic bytes per element: HBM-bound.
__global__ void __launch_bounds__(256)
k_encode_premasked_v4(const uint4* __restrict__ x, const uint4* __restrict__ mask, uint64_t begin, uint64_t nvec, uint32_t mk,
                      const __grid_constant__ CodecDev cd, const __grid_constant__ NoiseDev nz, uint4* __restrict__ ct_out) {
    const bool one_seg = cd.nseg == 1, one_rcp = one_seg && cd.seg[0].rcp_two_a != 0.0f;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t j = begin + 4ull * v;
        const uint4 xv = __ldcs(x + v), mv = __ldcs(mask + v);
        double u[4];
        if (nz.u) {
            const double2 a = __ldcs(reinterpret_cast<const double2*>(nz.u) + 2 * v), b = __ldcs(reinterpret_cast<const double2*>(nz.u) + 2 * v + 1);
            u[0] = a.x; u[1] = a.y; u[2] = b.x; u[3] = b.y;
        } else {
            noise_pair(nz, nz.stream, j >> 1, u[0], u[1]);
            noise_pair(nz, nz.stream, (j >> 1) + 1, u[2], u[3]);
        }
        const uint32_t xr[4] = {xv.x, xv.y, xv.z, xv.w};
        uint32_t q[4];
        if (one_rcp) {                                                 // single layer with a usable reciprocal (uniform)
#pragma unroll
            for (int k = 0; k < 4; ++k) q[k] = encode_one<true>(__uint_as_float(xr[k]), u[k], cd.seg[0], cd.scale);
        } else {
            Seg sg = find_seg(cd, j);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k && !one_seg && j + k >= sg.end) sg = find_seg(cd, j + k);
                q[k] = encode_one(__uint_as_float(xr[k]), u[k], sg, cd.scale);
            }
        }
        __stcs(ct_out + v, make_uint4((q[0] + mv.x) & mk, (q[1] + mv.y) & mk, (q[2] + mv.z) & mk, (q[3] + mv.w) & mk));
    }
}

template <int WORDS>
__global__ void k_decode(const typename Word<WORDS>::T* __restrict__ v, uint64_t begin, uint64_t count,
                         const __grid_constant__ CodecDev cd, double* __restrict__ out) {
    typedef Word<WORDS> WT;
    for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < count; o += (uint64_t)gridDim.x * blockDim.x) {
        const Seg sg = find_seg(cd, begin + o);
        out[o] = decode_one(WT::to_double(v[o]), sg.two_an, cd.den, cd.den_rcp, sg.an);
    }
}

// 4-byte words, 16-byte aligned buffers and begin: one thread = 4 elements (128-bit load, two 128-bit stores)
__global__ void __launch_bounds__(256)
k_decode_v4(const uint4* __restrict__ v, uint64_t begin, uint64_t nvec, const __grid_constant__ CodecDev cd, double* __restrict__ out) {
    const bool one_seg = cd.nseg == 1;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 w = __ldcs(v + i);
        const uint32_t p[4] = {w.x, w.y, w.z, w.w};
        const uint64_t j = begin + 4ull * i;
        double d[4];
        Seg sg = find_seg(cd, j);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k && !one_seg && j + k >= sg.end) sg = find_seg(cd, j + k);
            d[k] = decode_one((double)p[k], sg.two_an, cd.den, cd.den_rcp, sg.an);
        }
        stg_d2(out + 4ull * i, d[0], d[1]);
        stg_d2(out + 4ull * i + 2, d[2], d[3]);
    }
}

__global__ void k_rng_uniform(const __grid_constant__ NoiseDev nz, uint64_t begin, uint64_t count, double* __restrict__ out) {
    for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < count; o += (uint64_t)gridDim.x * blockDim.x)
        out[o] = noise_one(nz, nz.stream, begin + o);
}

// Element-wise server sum (jzf_aggregator.py:421-430).  One thread owns one 16-byte column of the
// [n][count] matrix and walks the n client rows with UNROLL independent 128-bit loads in flight.
template <int WORDS>
__global__ void __launch_bounds__(256)
k_aggregate_vec(const uint4* __restrict__ cts, uint64_t stride_vec, int n, uint64_t nvec, uint32_t b, uint4* __restrict__ out) {
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (uint64_t)gridDim.x * blockDim.x) {
        const uint4* p = cts + v;
        if (WORDS == 1) {
            uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
            int c = 0;
            for (; c + 8 <= n; c += 8) {
                uint4 r[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) r[k] = __ldcs(p + (uint64_t)(c + k) * stride_vec);
#pragma unroll
                for (int k = 0; k < 8; ++k) { a0 += r[k].x; a1 += r[k].y; a2 += r[k].z; a3 += r[k].w; }
            }
            for (; c < n; ++c) { uint4 r = __ldcs(p + (uint64_t)c * stride_vec); a0 += r.x; a1 += r.y; a2 += r.z; a3 += r.w; }
            const uint32_t mk = Word<1>::mask(b);
            out[v] = make_uint4(a0 & mk, a1 & mk, a2 & mk, a3 & mk);
        } else if (WORDS == 2) {
            uint64_t a0 = 0, a1 = 0;
            int c = 0;
            for (; c + 8 <= n; c += 8) {
                uint4 r[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) r[k] = __ldcs(p + (uint64_t)(c + k) * stride_vec);
#pragma unroll
                for (int k = 0; k < 8; ++k) { a0 += ((uint64_t)r[k].y << 32) | r[k].x; a1 += ((uint64_t)r[k].w << 32) | r[k].z; }
            }
            for (; c < n; ++c) { uint4 r = __ldcs(p + (uint64_t)c * stride_vec); a0 += ((uint64_t)r.y << 32) | r.x; a1 += ((uint64_t)r.w << 32) | r.z; }
            const uint64_t mk = Word<2>::mask(b);
            a0 &= mk; a1 &= mk;
            out[v] = make_uint4((uint32_t)a0, (uint32_t)(a0 >> 32), (uint32_t)a1, (uint32_t)(a1 >> 32));
        } else {
            u128 a = Word<4>::zero();
            for (int c = 0; c < n; ++c) {
                uint4 r = __ldcs(p + (uint64_t)c * stride_vec);
                u128 w; w.lo = ((uint64_t)r.y << 32) | r.x; w.hi = ((uint64_t)r.w << 32) | r.z;
                a = Word<4>::add(a, w);
            }
            a = Word<4>::band(a, Word<4>::mask(b));
            out[v] = make_uint4((uint32_t)a.lo, (uint32_t)(a.lo >> 32), (uint32_t)a.hi, (uint32_t)(a.hi >> 32));
        }
    }
}

// scalar fallback for unaligned rows / tails (u32 and u64 words)
template <int WORDS>
__global__ void k_aggregate_scalar(const typename Word<WORDS>::T* __restrict__ cts, uint64_t stride, int n, uint64_t j0,
                                   uint64_t count, uint32_t b, typename Word<WORDS>::T* __restrict__ out) {
    typedef Word<WORDS> WT;
    const typename WT::T mk = WT::mask(b);
    for (uint64_t j = j0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += (uint64_t)gridDim.x * blockDim.x) {
        typename WT::T a = WT::zero();
        for (int c = 0; c < n; ++c) a = WT::add(a, cts[(uint64_t)c * stride + j]);
        out[j] = WT::band(a, mk);
    }
}

// Packed-carry server sum (jzf_aggregator.py:404-419): radix-2^b addition of the n packed vectors,
// least significant digit = LAST element.  Digit sum S_j = H_j*2^b + lo_j; the carry into element
// j-1 is H_j + [lo_j + cin_j >= 2^b], i.e. a transfer function cin -> A + [cin >= T] with
// (A,T) = (H_j, 2^b - lo_j).  Such functions compose into the same form, so carries are resolved
// with a reverse scan: thread-serial over its ELEMS elements, shuffle scan across the warp, shared
// memory across warps, and a look-ahead across tiles: a tile obtains its carry-in by composing the
// transfer functions of the elements after it until the composition no longer depends on its own
// carry-in (T = never) — for ciphertext-like data that happens after one element with probability
// 1 - (n-1)/2^b — or the end of the range (carry_in) is reached.
struct Xfer { uint32_t A; uint32_t T; };  // cin -> A + (cin >= T); T == 0xffffffff: never
#define T_NEVER 0xffffffffu
__device__ __forceinline__ uint32_t xfer_apply(Xfer f, uint32_t cin) { return f.A + (cin >= f.T ? 1u : 0u); }
// h = outer ∘ inner  (inner is applied first: it belongs to the element closer to the end)
__device__ __forceinline__ Xfer xfer_compose(Xfer outer, Xfer inner) {
    Xfer h;
    const uint32_t lo = inner.A, hi = inner.A + 1;  // possible outputs of inner
    const bool lo_hit = lo >= outer.T, hi_hit = (inner.T != T_NEVER) && (hi >= outer.T);
    if (inner.T == T_NEVER || lo_hit == hi_hit) { h.A = outer.A + (lo_hit ? 1u : 0u); h.T = T_NEVER; }
    else { h.A = outer.A; h.T = inner.T; }  // lo misses, hi hits: depends on inner's threshold
    return h;
}
// The low part of a digit sum (S_j mod 2^b) in the width of the word; H_j = S_j >> b <= n - 1 fits 32 bits.
template <int WORDS> struct Dig { typedef uint64_t lo_t; };
template <> struct Dig<4> { typedef u128 lo_t; };

template <int WORDS>
__device__ __forceinline__ void digit_sum(const typename Word<WORDS>::T* __restrict__ cts, uint64_t stride, int n, uint64_t j,
                                          uint32_t b, typename Dig<WORDS>::lo_t& lo, uint32_t& H) {
    // returns S_j = H*2^b + lo with lo < 2^b
    if constexpr (WORDS == 1) {
        uint64_t s = 0;
        for (int c = 0; c < n; ++c) s += reinterpret_cast<const uint32_t*>(cts)[(uint64_t)c * stride + j];
        lo = s & ((1ull << b) - 1ull); H = (uint32_t)(s >> b);
    } else if constexpr (WORDS == 2) {
        const uint64_t mk = Word<2>::mask(b);
        uint64_t l = 0; uint32_t h = 0;
        for (int c = 0; c < n; ++c) {
            uint64_t w = reinterpret_cast<const uint64_t*>(cts)[(uint64_t)c * stride + j];
            uint64_t s = l + w;
            if (b >= 64) { h += (s < l); l = s; }
            else { h += (uint32_t)(s >> b); l = s & mk; }
        }
        lo = l; H = h;
    } else {
        // 16-byte words (b = 65..128): three-limb accumulator (lo64, hi64, top), rows read as 128-bit
        // streaming loads with eight of them in flight
        const uint4* col = reinterpret_cast<const uint4*>(cts) + j;
        uint64_t l0 = 0, l1 = 0; uint32_t top = 0;
#pragma unroll 8
        for (int c = 0; c < n; ++c) {
            const uint4 v = __ldcs(col + (uint64_t)c * stride);
            const uint64_t w0 = ((uint64_t)v.y << 32) | v.x, w1 = ((uint64_t)v.w << 32) | v.z;
            const uint64_t s0 = l0 + w0;
            const uint64_t c0 = s0 < l0 ? 1ull : 0ull;
            const uint64_t t1 = l1 + w1;
            const uint32_t ca = t1 < l1 ? 1u : 0u;
            const uint64_t s1 = t1 + c0;
            const uint32_t cb = s1 < t1 ? 1u : 0u;
            l0 = s0; l1 = s1; top += ca + cb;
        }
        u128 r; r.lo = l0;
        if (b >= 128) { r.hi = l1; H = top; }
        else {
            const uint32_t sh = b - 64u;                               // 1..63
            r.hi = l1 & ((1ull << sh) - 1ull);
            H = (uint32_t)(((uint64_t)top << (64u - sh)) | (l1 >> sh));
        }
        lo = r;
    }
}
__device__ __forceinline__ Xfer xfer_of(uint64_t lo, uint32_t H, uint32_t b) {
    Xfer f; f.A = H;
    // threshold 2^b - lo, only relevant when it is small (cin <= n-1 < 2^31)
    uint64_t thr = (b >= 64) ? (0ull - lo) : ((1ull << b) - lo);
    f.T = (lo != 0 && thr < 0x7fffffffull) ? (uint32_t)thr : T_NEVER;
    return f;
}
__device__ __forceinline__ Xfer xfer_of(u128 lo, uint32_t H, uint32_t b) {
    Xfer f; f.A = H;
    // 2^b - lo = (-lo) mod 2^b for 0 < lo < 2^b
    u128 neg; neg.lo = 0ull - lo.lo; neg.hi = ~lo.hi + (lo.lo == 0ull ? 1ull : 0ull);
    neg = Word<4>::band(neg, Word<4>::mask(b));
    const bool nz = (lo.lo | lo.hi) != 0ull;
    f.T = (nz && neg.hi == 0ull && neg.lo < 0x7fffffffull) ? (uint32_t)neg.lo : T_NEVER;
    return f;
}
// lo + c (c small) -> value mod 2^b, carry out of b bits
__device__ __forceinline__ uint64_t add_small(uint64_t lo, uint32_t c, uint32_t b, uint32_t& extra) {
    uint64_t s = lo + c;
    if (b >= 64) { extra = (s < lo) ? 1u : 0u; return s; }
    extra = (uint32_t)(s >> b);
    return s & ((1ull << b) - 1ull);
}
__device__ __forceinline__ u128 add_small(u128 lo, uint32_t c, uint32_t b, uint32_t& extra) {
    u128 s; s.lo = lo.lo + c; s.hi = lo.hi + (s.lo < lo.lo ? 1ull : 0ull);
    if (b >= 128) { extra = (s.hi < lo.hi) ? 1u : 0u; return s; }
    const uint32_t sh = b - 64u;
    extra = (uint32_t)(s.hi >> sh);
    s.hi &= (1ull << sh) - 1ull;
    return s;
}
template <int WORDS> __device__ __forceinline__ typename Word<WORDS>::T word_of(typename Dig<WORDS>::lo_t v);
template <> __device__ __forceinline__ uint32_t word_of<1>(uint64_t v) { return (uint32_t)v; }
template <> __device__ __forceinline__ uint64_t word_of<2>(uint64_t v) { return v; }
template <> __device__ __forceinline__ u128 word_of<4>(u128 v) { return v; }
template <int WORDS> __device__ __forceinline__ typename Dig<WORDS>::lo_t lo_zero() { return 0ull; }
template <> __device__ __forceinline__ u128 lo_zero<4>() { return Word<4>::zero(); }

#define PK_THREADS 256
// elements per thread: four 4- or 8-byte words (one 16- or 32-byte column), one 16-byte word
template <int WORDS> struct PkElems { static constexpr int V = WORDS == 4 ? 1 : 4; };
template <int WORDS>
__global__ void __launch_bounds__(PK_THREADS)
k_aggregate_packed(const typename Word<WORDS>::T* __restrict__ cts, uint64_t stride, int n, uint64_t count, uint32_t b,
                   uint32_t carry_in, typename Word<WORDS>::T* __restrict__ out, uint32_t* __restrict__ desc_out, int vec_ok) {
    typedef Word<WORDS> WT;
    typedef typename Dig<WORDS>::lo_t lo_t;
    constexpr int PK_ELEMS = PkElems<WORDS>::V;
    __shared__ Xfer warp_x[PK_THREADS / 32];
    __shared__ uint32_t tile_cin;
    const uint64_t tile_elems = (uint64_t)PK_THREADS * PK_ELEMS;
    const uint64_t ntiles = (count + tile_elems - 1) / tile_elems;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint64_t mk64 = (b >= 64) ? ~0ull : ((1ull << b) - 1ull);

    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // tiles and threads are numbered from the END of the vector (carry flows towards element 0):
        // thread q of tile t owns elements hi-1 .. hi-ELEMS with hi = count - (t*tile_elems + q*ELEMS)
        const uint64_t base = tile * tile_elems + (uint64_t)threadIdx.x * PK_ELEMS;
        lo_t lo_[PK_ELEMS]; uint32_t H_[PK_ELEMS];
        Xfer mine; mine.A = 0; mine.T = 0;  // identity: cin -> cin is not representable; track validity
        bool have = false;
        // 4-byte words, count and every row 16-byte aligned: the thread's four elements are one 128-bit
        // column of the [n, count] matrix; walk the rows with independent streaming loads in flight
        const bool quad = WORDS == 1 && vec_ok && base + PK_ELEMS <= count;
        if constexpr (WORDS == 1) {
          if (quad) {
            const uint4* col = reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(cts) + (count - PK_ELEMS - base));
            const uint64_t sv = stride >> 2;
            uint64_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll 8
            for (int c = 0; c < n; ++c) {
                const uint4 v = __ldcs(col + (uint64_t)c * sv);
                s0 += v.w; s1 += v.z; s2 += v.y; s3 += v.x;          // element e sits `e` places before the end: .w first
            }
            const uint64_t mkb = (1ull << b) - 1ull;
            lo_[0] = s0 & mkb; H_[0] = (uint32_t)(s0 >> b); lo_[1] = s1 & mkb; H_[1] = (uint32_t)(s1 >> b);
            lo_[2] = s2 & mkb; H_[2] = (uint32_t)(s2 >> b); lo_[3] = s3 & mkb; H_[3] = (uint32_t)(s3 >> b);
#pragma unroll
            for (int e = 0; e < PK_ELEMS; ++e) {
                Xfer f = xfer_of(lo_[e], H_[e], b);
                mine = have ? xfer_compose(f, mine) : f;
                have = true;
            }
          }
        }
        if (!quad) {
#pragma unroll
            for (int e = 0; e < PK_ELEMS; ++e) {
                const uint64_t back = base + e;  // distance from the end
                if (back < count) {
                    digit_sum<WORDS>(cts, stride, n, count - 1 - back, b, lo_[e], H_[e]);
                    Xfer f = xfer_of(lo_[e], H_[e], b);
                    mine = have ? xfer_compose(f, mine) : f;
                    have = true;
                } else { lo_[e] = lo_zero<WORDS>(); H_[e] = 0; }
            }
        }
        // Identity handling: a thread with no elements must pass the carry through unchanged.  That
        // only happens in the last (partial) tile, where such threads sit AFTER all real ones in scan
        // order (larger `back`), so their value is never consumed; give them a harmless constant.
        if (!have) { mine.A = 0; mine.T = T_NEVER; }

        // warp-level inclusive scan in `back` order (lane 0 is closest to the end):
        // incl[l] = f_l ∘ f_{l-1} ∘ ... ∘ f_0
        Xfer incl = mine;
#pragma unroll
        for (int dlt = 1; dlt < 32; dlt <<= 1) {
            Xfer o; o.A = __shfl_up_sync(0xffffffffu, incl.A, dlt); o.T = __shfl_up_sync(0xffffffffu, incl.T, dlt);
            if (lane >= (uint32_t)dlt) incl = xfer_compose(incl, o);
        }
        if (lane == 31) warp_x[warp] = incl;

        // look-ahead for the tile's carry-in (elements closer to the end than this tile)
        if (threadIdx.x == 0) {
            uint32_t cin;
            if (tile == 0) cin = carry_in;
            else {
                // compose f_{j} for j just after the tile, walking towards the end, until constant
                const uint64_t first_back = tile * tile_elems;  // `back` of this tile's first element
                Xfer acc; bool started = false; uint64_t bk = first_back;  // walk bk-1, bk-2, ... 0
                cin = 0; bool resolved = false;
                while (bk > 0) {
                    --bk;
                    lo_t l; uint32_t h;
                    digit_sum<WORDS>(cts, stride, n, count - 1 - bk, b, l, h);
                    Xfer f = xfer_of(l, h, b);
                    // acc currently maps (carry into element bk+1.. chain) ; new element is applied FIRST
                    acc = started ? xfer_compose(acc, f) : f;
                    started = true;
                    if (acc.T == T_NEVER) { cin = acc.A; resolved = true; break; }
                }
                if (!resolved) cin = started ? xfer_apply(acc, carry_in) : carry_in;
            }
            tile_cin = cin;
        }
        __syncthreads();
        // carry into this warp = composition of the previous warps applied to tile_cin
        uint32_t cin = tile_cin;
        for (uint32_t w = 0; w < warp; ++w) cin = xfer_apply(warp_x[w], cin);
        // carry into this lane's first element: exclusive prefix within the warp
        Xfer ex; ex.A = __shfl_up_sync(0xffffffffu, incl.A, 1); ex.T = __shfl_up_sync(0xffffffffu, incl.T, 1);
        uint32_t c = lane == 0 ? cin : xfer_apply(ex, cin);
        bool stored = false;
        if constexpr (WORDS == 1) {
          if (quad && vec_ok > 1) {                                          // out is 16-byte aligned too: one 128-bit store
            uint32_t r[PK_ELEMS];
#pragma unroll
            for (int e = 0; e < PK_ELEMS; ++e) {
                const uint64_t sum = lo_[e] + c;                           // lo < 2^b <= 2^32, c small
                r[e] = (uint32_t)(sum & mk64);
                c = H_[e] + (uint32_t)(sum >> b);
            }
            __stcs(reinterpret_cast<uint4*>(reinterpret_cast<uint32_t*>(out) + (count - PK_ELEMS - base)), make_uint4(r[3], r[2], r[1], r[0]));
            stored = true;
          }
        }
        if (!stored) {
#pragma unroll
            for (int e = 0; e < PK_ELEMS; ++e) {
                const uint64_t back = base + e;
                if (back < count) {
                    uint32_t extra;
                    const lo_t s = add_small(lo_[e], c, b, extra);   // lo < 2^b, c small
                    if constexpr (WORDS == 4) {
                        __stcs(reinterpret_cast<uint4*>(out) + (count - 1 - back),
                               make_uint4((uint32_t)s.lo, (uint32_t)(s.lo >> 32), (uint32_t)s.hi, (uint32_t)(s.hi >> 32)));
                    } else {
                        out[count - 1 - back] = word_of<WORDS>(s);
                    }
                    c = H_[e] + extra;
                }
            }
        }
        // Range descriptor for element-range shards (desc_out = {carry out for the given carry_in,
        // depends, A, T}).  Word 0 comes from the thread that owns element 0.  Words 1-3 come from a
        // walk from the END of the range: if the composed transfer function becomes constant, the
        // carry out of the range cannot depend on carry_in (depends = 0); otherwise the walk has
        // covered the whole range and (A, T) is its exact transfer function (depends = 1).
        if (desc_out && tile == ntiles - 1) {
            const uint64_t last_back = count - 1;
            if (last_back >= base && last_back < base + PK_ELEMS) desc_out[0] = c;
        }
        if (desc_out && tile == 0 && threadIdx.x == 0) {
            Xfer acc; acc.A = 0; acc.T = T_NEVER; bool started = false, resolved = false;
            for (uint64_t bk = 0; bk < count; ++bk) {
                lo_t l; uint32_t h;
                digit_sum<WORDS>(cts, stride, n, count - 1 - bk, b, l, h);
                Xfer f = xfer_of(l, h, b);
                acc = started ? xfer_compose(f, acc) : f;   // later elements are applied after (outer)
                started = true;
                if (acc.T == T_NEVER) { resolved = true; break; }
            }
            desc_out[1] = resolved ? 0u : 1u; desc_out[2] = acc.A; desc_out[3] = acc.T;
        }
        __syncthreads();
    }
}

// Ripple a late carry-in into an already aggregated shard (multi-GPU packed sum): out is the
// radix-2^b number whose least significant digit is the LAST element.
template <int WORDS>
__global__ void k_carry_fixup(typename Word<WORDS>::T* __restrict__ out, uint64_t count, uint32_t b, uint32_t carry_in) {
    if (blockIdx.x || threadIdx.x) return;
    uint32_t c = carry_in;
    for (uint64_t j = count; c && j-- > 0;) {
        uint32_t extra;
        if constexpr (WORDS == 1) {
            const uint64_t v = add_small((uint64_t)out[j], c, b, extra);
            out[j] = (uint32_t)v;
        } else {
            out[j] = add_small(out[j], c, b, extra);
        }
        c = extra;
    }
}

// lane batching (jzf_quantize.py:162-185, 234-251)
__global__ void k_batch_pack(const uint32_t* __restrict__ q, uint64_t count, uint32_t lane_bits, uint32_t bs, uint64_t nwords,
                             u128* __restrict__ out) {
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t lo = 0, hi = 0;
        for (uint32_t i = 0; i < bs; ++i) {
            const uint64_t j = w * bs + i;
            const uint64_t v = j < count ? q[j] : 0u;
            hi = (hi << lane_bits) | (lo >> (64 - lane_bits));   // lane_bits in [1,32]
            lo = (lo << lane_bits) + v;                          // v < 2^lane_bits: no carry
        }
        u128 r; r.lo = lo; r.hi = hi;
        out[w] = r;
    }
}
__global__ void k_batch_unpack(const u128* __restrict__ in, uint64_t nwords, uint32_t lane_bits, uint32_t bs, uint32_t* __restrict__ out) {
    const uint64_t lm = (1ull << lane_bits) - 1ull;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (uint64_t)gridDim.x * blockDim.x) {
        u128 t = in[w];
        for (int i = (int)bs - 1; i >= 0; --i) {
            out[w * bs + i] = (uint32_t)(t.lo & lm);
            t.lo = (t.lo >> lane_bits) | (t.hi << (64 - lane_bits));
            t.hi >>= lane_bits;
        }
    }
}

// expand_to_dense (jzf_aggregator.py:150-165)
template <int WORDS>
__global__ void k_fill(typename Word<WORDS>::T* __restrict__ out, uint64_t count, typename Word<WORDS>::T v) {
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += (uint64_t)gridDim.x * blockDim.x) out[j] = v;
}
template <int WORDS>
__global__ void k_scatter(const typename Word<WORDS>::T* __restrict__ compact, const int64_t* __restrict__ index, uint64_t k,
                          uint64_t total, typename Word<WORDS>::T* __restrict__ dense) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < k; i += (uint64_t)gridDim.x * blockDim.x) {
        const int64_t d = index[i];
        if (d >= 0 && (uint64_t)d < total) dense[d] = compact[i];
    }
}
// |A ∩ B| for sorted unique index lists: each element of A binary-searches B
__global__ void k_overlap(const int64_t* __restrict__ a, uint64_t ka, const int64_t* __restrict__ bq, uint64_t kb,
                          unsigned long long* __restrict__ out) {
    unsigned long long local = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ka; i += (uint64_t)gridDim.x * blockDim.x) {
        const int64_t v = a[i];
        uint64_t lo = 0, hi = kb;
        while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (bq[mid] < v) lo = mid + 1; else hi = mid; }
        local += (lo < kb && bq[lo] == v) ? 1ull : 0ull;
    }
    for (int dlt = 16; dlt > 0; dlt >>= 1) local += __shfl_down_sync(0xffffffffu, local, dlt);
    if ((threadIdx.x & 31u) == 0 && local) atomicAdd(out, local);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct flashe_ctx {
    int device;
    int int_bits;
    int words;       // 1, 2, 4
    int num_sms;
    uint32_t m;
    uint8_t key[32];
    KeySched ks;
};

static int words_of(int b) { return b <= 32 ? 1 : (b <= 64 ? 2 : 4); }

int flashe_ctx_get_info(const flashe_ctx* ctx, flashe_ctx_info* out) {
    if (!ctx) return fail(FLASHE_EINVAL, "ctx is NULL");
    out->device = ctx->device; out->int_bits = ctx->int_bits; out->words = ctx->words; out->num_sms = ctx->num_sms;
    return FLASHE_OK;
}

struct DeviceGuard {
    int prev;
    bool ok;
    explicit DeviceGuard(int dev) : prev(-1), ok(true) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static int check_span(const flashe_span* s) {
    if (!s) return fail(FLASHE_EINVAL, "span is NULL");
    if (s->n_jobs == 0) return fail(FLASHE_EINVAL, "span.n_jobs must be >= 1");
    if (s->reserved != 0) return fail(FLASHE_EINVAL, "span.reserved must be 0");
    if (s->begin > s->total_len || s->count > s->total_len - s->begin)
        return fail(FLASHE_EINVAL, "span [begin, begin+count) exceeds total_len");
    return FLASHE_OK;
}

static uint64_t ceil_div(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

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

// Items per work unit: as large as possible (amortises the per-unit chunk decode) while every warp of
// the persistent grid still gets at least ~32 units.
static uint32_t pick_sup(const flashe_ctx* ctx, const flashe_span* s, uint64_t rows) {
    const uint64_t items = ceil_div(ceil_div(s->count ? s->count : 1, ctx->m), ITEM_BLOCKS) * (rows ? rows : 1);
    const uint64_t warps = (uint64_t)ctx->num_sms * (STREAM_THREADS / 32);
    uint64_t sup = items / (warps * 32);
    return (uint32_t)(sup < 1 ? 1 : (sup > 32 ? 32 : sup));
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

struct CodecHost {
    CodecDev dev;
    Seg* table;  // device allocation to free (stream ordered) or NULL
};

static int make_codec(const flashe_ctx* ctx, const flashe_span* span, const flashe_codec* c, bool decode, cudaStream_t stream,
                      CodecHost* out) {
    (void)ctx;
    memset(out, 0, sizeof(*out));
    if (!c) return fail(FLASHE_EINVAL, "codec is NULL");
    if (c->element_bits < 1 || c->element_bits > 24) return fail(FLASHE_EINVAL, "element_bits must be in [1, 24]");
    if (c->nseg < 1 || !c->seg_end || !c->alpha) return fail(FLASHE_EINVAL, "codec needs nseg >= 1, seg_end and alpha");
    if (decode && c->n_clients < 1) return fail(FLASHE_EINVAL, "codec.n_clients must be >= 1 for decode");
    if (c->seg_end[c->nseg - 1] != span->total_len) return fail(FLASHE_EINVAL, "seg_end[nseg-1] must equal span.total_len");
    std::vector<Seg> segs((size_t)c->nseg);
    const int n = decode ? c->n_clients : 1;
    for (int s = 0; s < c->nseg; ++s) {
        if (s && c->seg_end[s] < c->seg_end[s - 1]) return fail(FLASHE_EINVAL, "seg_end must be ascending");
        segs[s].end = c->seg_end[s];
        segs[s].a = (float)c->alpha[s];
        segs[s].two_a = (float)(2.0 * c->alpha[s]);
        // RN(1/two_a): the double quotient is at least 2^-49 (relative) away from any float rounding
        // boundary, so rounding it to float cannot double-round
        const float ta = segs[s].two_a;
        segs[s].rcp_two_a = (ta >= 9.094947017729282e-13f /* 2^-40 */ && ta <= 1.152921504606847e18f /* 2^60 */) ? (float)(1.0 / (double)ta) : 0.0f;
        segs[s].pad = 0.0f;
        volatile double an = c->alpha[s] * (double)n;     // alpha *= num_clients (jzf_quantize.py:103)
        segs[s].an = an;
        segs[s].two_an = 2.0 * an;
    }
    CodecDev& d = out->dev;
    d.nseg = c->nseg; d.ebits = c->element_bits;
    d.scale = (float)(((int64_t)1 << c->element_bits) - 1);
    d.den = (double)((((int64_t)1 << c->element_bits) - 1) * (int64_t)n);
    {
        bool ok = decode;
        for (int s = 0; ok && s < c->nseg; ++s) {
            const double t = fabs(segs[s].two_an);
            ok = (t == 0.0) || (t >= 0x1p-400 && t <= 0x1p400);
        }
        volatile double y = 1.0 / d.den;                  // IEEE division: correctly rounded
        d.den_rcp = ok ? y : 0.0;
    }
    if (c->nseg <= MAX_INLINE_SEG) {
        memcpy(d.seg, segs.data(), sizeof(Seg) * (size_t)c->nseg);
        d.table = nullptr;
    } else {
        // large layer tables travel through a stream-ordered allocation (pageable copy: the runtime
        // stages it before returning)
        CUDA_TRY(cudaMallocAsync((void**)&out->table, sizeof(Seg) * (size_t)c->nseg, stream));
        CUDA_TRY(cudaMemcpyAsync(out->table, segs.data(), sizeof(Seg) * (size_t)c->nseg, cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));  // segs goes out of scope
        d.table = out->table;
    }
    return FLASHE_OK;
}
static void free_codec(CodecHost* c, cudaStream_t stream) { if (c->table) cudaFreeAsync(c->table, stream); }

static void make_noise(const flashe_noise* nz, uint64_t u_stride, NoiseDev* d) {
    memset(d, 0, sizeof(*d));
    if (!nz) return;
    d->u = nz->u; d->u_stride = u_stride;
    uint32_t k0 = (uint32_t)nz->rng_seed, k1 = (uint32_t)(nz->rng_seed >> 32);
    for (int i = 0; i < 10; ++i) { d->rk[i][0] = k0; d->rk[i][1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    d->stream = nz->rng_stream;
}

static int grid_1d(const flashe_ctx* ctx, uint64_t work_items, int threads, int per_sm) {
    uint64_t blocks = ceil_div(work_items ? work_items : 1, (uint64_t)threads);
    uint64_t cap = (uint64_t)ctx->num_sms * per_sm;
    return (int)(blocks < cap ? blocks : cap);
}

template <int WORDS, int MMAX, int MODE, bool SHARE, bool ALIGNED = false>
static int launch_stream_t(const flashe_ctx* ctx, const StreamTab& st, const Geom& g, const IoDev& io, const CodecDev& cd,
                           const NoiseDev& nz, cudaStream_t stream) {
    auto kern = k_stream<WORDS, MMAX, MODE, SHARE, ALIGNED>;
    // the opt-in shared-memory size is a per-device property of the function: set it once per device
    static std::atomic<uint64_t> attr_set{0};
    const uint64_t dev_bit = 1ull << (ctx->device & 63);
    if (!(attr_set.load(std::memory_order_acquire) & dev_bit)) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        attr_set.fetch_or(dev_bit, std::memory_order_release);
    }
    const uint64_t items = (st.batch && !io.share) ? g.S_cnt * io.n_clients : g.S_cnt;
    if (items == 0) return FLASHE_OK;
    const int slab_bytes = (2 * 32 * MMAX + 2 + (WORDS == 1 ? 2 * MMAX + 1 : 0)) * WORDS * 4;
    int threads = STREAM_THREADS;
    while (threads > 32 && (threads / 32) * slab_bytes > 60 * 1024) threads >>= 1;
    const int wpb = threads / 32;
    uint64_t blocks = ceil_div(items, (uint64_t)wpb);
    if (blocks > (uint64_t)ctx->num_sms) blocks = (uint64_t)ctx->num_sms;
    kern<<<(unsigned)blocks, threads, SMEM_BYTES, stream>>>(ctx->ks, st, g, io, cd, nz);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}

static bool aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }

template <int MODE>
static int launch_stream(const flashe_ctx* ctx, const StreamTab& st, const Geom& g, const IoDev& io_in, const CodecDev& cd,
                         const NoiseDev& nz, cudaStream_t stream) {
    const int b = ctx->int_bits;
    IoDev io = io_in;
    // 128-bit fast path preconditions (4-byte words): every row of every buffer starts 16-byte aligned
    // (16-byte words: every word is aligned as soon as the base pointers are)
    io.quad = (MODE != M_SCATTER && aligned16(io.in) && aligned16(io.out) && aligned16(io.aux) && aligned16(io.outf) &&
               ((ctx->words == 1 && (io.n_clients <= 1 || ((io.in_stride | io.out_stride) & 3ull) == 0)) ||
                (ctx->words == 2 && (io.n_clients <= 1 || ((io.in_stride | io.out_stride) & 1ull) == 0)) || ctx->words == 4)) ? 1u : 0u;
    if constexpr (MODE == M_ENCODE) {
        if (io.share) {
            if (b <= 32) {
                if (ctx->m == 4 && g.aligned4 && io.quad) return launch_stream_t<1, 4, MODE, true, true>(ctx, st, g, io, cd, nz, stream);
                if (ctx->m <= 4) return launch_stream_t<1, 4, MODE, true>(ctx, st, g, io, cd, nz, stream);
                if (ctx->m <= 6) return launch_stream_t<1, 6, MODE, true>(ctx, st, g, io, cd, nz, stream);
                return launch_stream_t<1, 16, MODE, true>(ctx, st, g, io, cd, nz, stream);
            }
            if (b <= 64) return launch_stream_t<2, 3, MODE, true>(ctx, st, g, io, cd, nz, stream);
            return launch_stream_t<4, 1, MODE, true>(ctx, st, g, io, cd, nz, stream);
        }
    }
    if (b <= 32) {
        if (ctx->m == 4 && g.aligned4 && io.quad && MODE != M_SCATTER) return launch_stream_t<1, 4, MODE, false, true>(ctx, st, g, io, cd, nz, stream);
        if (ctx->m <= 4) return launch_stream_t<1, 4, MODE, false>(ctx, st, g, io, cd, nz, stream);
        if (ctx->m <= 6) return launch_stream_t<1, 6, MODE, false>(ctx, st, g, io, cd, nz, stream);
        return launch_stream_t<1, 16, MODE, false>(ctx, st, g, io, cd, nz, stream);
    }
    if (b <= 64) return launch_stream_t<2, 3, MODE, false>(ctx, st, g, io, cd, nz, stream);
    return launch_stream_t<4, 1, MODE, false>(ctx, st, g, io, cd, nz, stream);
}

// dense[index[i]] += compact[i] - zero  (mod 2^b): one client's contribution to the sum of the expanded
// vectors once `dense` holds the sum of every client's zero word (index sorted unique: no conflicts)
template <int WORDS>
__global__ void k_scatter_add(const typename Word<WORDS>::T* __restrict__ compact, const int64_t* __restrict__ index, uint64_t k,
                              uint64_t total, typename Word<WORDS>::T zero, uint32_t b, typename Word<WORDS>::T* __restrict__ dense) {
    typedef Word<WORDS> WT;
    const typename WT::T mk = WT::mask(b);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < k; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t d = (uint64_t)index[i];
        if (d < total) dense[d] = WT::band(WT::add(dense[d], WT::sub(compact[i], zero)), mk);
    }
}

template <int WORDS>
static int sparse_sum_t(flashe_ctx* ctx, const void* const* compacts, const int64_t* const* indexes, const uint64_t* ks,
                        const void* zero_words, int n, uint64_t total, void* dense_out, cudaStream_t cs) {
    typedef Word<WORDS> WT;
    typedef typename WT::T word_t;
    std::vector<word_t> zeros((size_t)n);                              // the caller's buffer need not be aligned
    memcpy(zeros.data(), zero_words, sizeof(word_t) * (size_t)n);
    const word_t mk = WT::mask((uint32_t)ctx->int_bits);
    word_t zsum = WT::zero();
    for (int c = 0; c < n; ++c) zsum = WT::band(WT::add(zsum, WT::band(zeros[c], mk)), mk);
    k_fill<WORDS><<<grid_1d(ctx, total, 256, 16), 256, 0, cs>>>((word_t*)dense_out, total, zsum);
    int launches = 1;
    for (int c = 0; c < n; ++c) {
        if (!ks[c]) continue;
        k_scatter_add<WORDS><<<grid_1d(ctx, ks[c], 256, 16), 256, 0, cs>>>((const word_t*)compacts[c], indexes[c], ks[c], total,
                                                                           WT::band(zeros[c], mk), (uint32_t)ctx->int_bits, (word_t*)dense_out);
        ++launches;
    }
    g_launches.fetch_add(launches);
    CUDA_TRY(cudaGetLastError());
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
    DeviceGuard guard(device);
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
    cudaError_t e = cudaMemcpyToSymbol(g_te0, haes::te0, sizeof(haes::te0));
    if (e != cudaSuccess) { delete ctx; return fail(FLASHE_ECUDA, std::string("cudaMemcpyToSymbol: ") + cudaGetErrorString(e)); }
    *out = ctx;
    return FLASHE_OK;
}

int flashe_ctx_destroy(flashe_ctx* ctx) {
    if (!ctx) return FLASHE_OK;
    memset(ctx->key, 0, sizeof(ctx->key));
    memset(&ctx->ks, 0, sizeof(ctx->ks));
    delete ctx;
    return FLASHE_OK;
}
int flashe_ctx_int_bits(const flashe_ctx* ctx) { return ctx ? ctx->int_bits : fail(FLASHE_EINVAL, "ctx is NULL"); }
int flashe_ctx_device(const flashe_ctx* ctx) { return ctx ? ctx->device : fail(FLASHE_EINVAL, "ctx is NULL"); }

#define ENTER(ctx)                                                        \
    if (!(ctx)) return fail(FLASHE_EINVAL, "ctx is NULL");                \
    DeviceGuard guard__((ctx)->device);                                   \
    if (!guard__.ok) return fail(FLASHE_ECUDA, "cudaSetDevice failed");   \
    cudaStream_t cs = (cudaStream_t)stream

int flashe_prp_block(flashe_ctx* ctx, const uint8_t in16[16], uint8_t out16[16], void* stream) {
    ENTER(ctx);
    if (!in16 || !out16) return fail(FLASHE_EINVAL, "NULL buffer");
    uint32_t w[4], o[8];
    for (int i = 0; i < 4; ++i) w[i] = ((uint32_t)in16[4 * i] << 24) | ((uint32_t)in16[4 * i + 1] << 16) | ((uint32_t)in16[4 * i + 2] << 8) | in16[4 * i + 3];
    uint32_t* d = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d, 12 * sizeof(uint32_t)));
    cudaError_t e = cudaMemcpyAsync(d, w, 16, cudaMemcpyHostToDevice, cs);
    if (e == cudaSuccess) {
        auto kern = k_prp_block;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (e == cudaSuccess) {
            kern<<<1, 256, SMEM_BYTES, cs>>>(ctx->ks, d, d + 4);
            g_launches.fetch_add(1);
            e = cudaGetLastError();
        }
    }
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
    int rc = check_span(span); if (rc) return rc;
    if (span->count && !out) return fail(FLASHE_EINVAL, "out is NULL");
    StreamTab st; rc = make_streams(ctx, iter, prf_idx, sign, nstreams, &st); if (rc) return rc;
    Geom g; make_geom(ctx, span, pick_sup(ctx, span, 1), &g);
    IoDev io; memset(&io, 0, sizeof(io)); io.out = out; io.n_clients = 1;
    CodecDev cd; memset(&cd, 0, sizeof(cd)); NoiseDev nz; memset(&nz, 0, sizeof(nz));
    return launch_stream<M_MASKS>(ctx, st, g, io, cd, nz, cs);
}

int flashe_precompute(flashe_ctx* ctx, uint32_t iter_from, int n_rounds, const int32_t* prf_idx, const int32_t* sign, int nstreams,
                      const flashe_span* span, void* out, uint64_t round_stride, void* stream) {
    if (!ctx) return fail(FLASHE_EINVAL, "ctx is NULL");
    if (n_rounds < 0) return fail(FLASHE_EINVAL, "n_rounds must be >= 0");
    int rc = check_span(span); if (rc) return rc;
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
    int rc = check_span(span); if (rc) return rc;
    if (span->count && (!in || !out)) return fail(FLASHE_EINVAL, "in/out is NULL");
    StreamTab st; rc = make_streams(ctx, iter, prf_idx, sign, nstreams, &st); if (rc) return rc;
    Geom g; make_geom(ctx, span, pick_sup(ctx, span, 1), &g);
    IoDev io; memset(&io, 0, sizeof(io)); io.in = in; io.out = out; io.n_clients = 1;
    CodecDev cd; memset(&cd, 0, sizeof(cd)); NoiseDev nz; memset(&nz, 0, sizeof(nz));
    return launch_stream<M_APPLY>(ctx, st, g, io, cd, nz, cs);
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

int flashe_add_premasked(flashe_ctx* ctx, const void* in, const void* mask, int sign, uint64_t count, void* out, void* stream) {
    ENTER(ctx);
    if (count == 0) return FLASHE_OK;
    if (!in || !mask || !out) return fail(FLASHE_EINVAL, "NULL buffer");
    const uint32_t b = (uint32_t)ctx->int_bits;
    if (ctx->words == 1) {
        const bool aligned = (((uintptr_t)in | (uintptr_t)mask | (uintptr_t)out) & 15u) == 0;
        uint64_t nvec = aligned ? count / 4 : 0;
        if (nvec) {
            k_add_premasked_v4<<<grid_1d(ctx, nvec, 256, 16), 256, 0, cs>>>((const uint4*)in, (const uint4*)mask, sign, nvec, Word<1>::mask(b) , (uint4*)out);
            g_launches.fetch_add(1);
        }
        const uint64_t done = nvec * 4;
        if (done < count) {
            k_add_premasked<1><<<grid_1d(ctx, count - done, 256, 16), 256, 0, cs>>>((const uint32_t*)in + done, (const uint32_t*)mask + done, sign, count - done, b, (uint32_t*)out + done);
            g_launches.fetch_add(1);
        }
    } else if (ctx->words == 2) {
        k_add_premasked<2><<<grid_1d(ctx, count, 256, 16), 256, 0, cs>>>((const uint64_t*)in, (const uint64_t*)mask, sign, count, b, (uint64_t*)out);
        g_launches.fetch_add(1);
    } else {
        k_add_premasked<4><<<grid_1d(ctx, count, 256, 16), 256, 0, cs>>>((const u128*)in, (const u128*)mask, sign, count, b, (u128*)out);
        g_launches.fetch_add(1);
    }
    CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}

int flashe_encode(flashe_ctx* ctx, const flashe_span* span, const float* x, const flashe_codec* codec, const flashe_noise* noise,
                  uint32_t* q_out, void* stream) {
    ENTER(ctx);
    int rc = check_span(span); if (rc) return rc;
    if (span->count == 0) return FLASHE_OK;
    if (!x || !q_out) return fail(FLASHE_EINVAL, "NULL buffer");
    CodecHost ch; rc = make_codec(ctx, span, codec, false, cs, &ch); if (rc) return rc;
    NoiseDev nz; make_noise(noise, 0, &nz);
    k_encode<1, false><<<grid_1d(ctx, span->count, 256, 16), 256, 0, cs>>>(x, nullptr, span->begin, span->count, 32, ch.dev, nz, q_out, nullptr);
    g_launches.fetch_add(1);
    free_codec(&ch, cs);
    CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}

static int encode_encrypt_impl(flashe_ctx* ctx, uint32_t iter, int32_t idx0, int n_clients, int scheme, const flashe_span* span,
                               const float* x, uint64_t x_stride, const flashe_codec* codec, const flashe_noise* noise,
                               uint64_t u_stride, void* ct_out, uint64_t ct_stride, uint32_t* q_out, int share, void* stream) {
    ENTER(ctx);
    int rc = check_span(span); if (rc) return rc;
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
    NoiseDev nz; make_noise(noise, u_stride, &nz);
    IoDev io; memset(&io, 0, sizeof(io));
    io.in = x; io.in_stride = x_stride; io.out = ct_out; io.out_stride = ct_stride; io.aux = q_out;
    io.n_clients = (uint32_t)n_clients; io.share = (share && dbl) ? 1u : 0u;
    rc = launch_stream<M_ENCODE>(ctx, st, g, io, ch.dev, nz, cs);
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

int flashe_encode_add_premasked(flashe_ctx* ctx, const flashe_span* span, const float* x, const flashe_codec* codec,
                                const flashe_noise* noise, const void* mask, void* ct_out, void* stream) {
    ENTER(ctx);
    int rc = check_span(span); if (rc) return rc;
    if (span->count == 0) return FLASHE_OK;
    if (!x || !mask || !ct_out) return fail(FLASHE_EINVAL, "NULL buffer");
    CodecHost ch; rc = make_codec(ctx, span, codec, false, cs, &ch); if (rc) return rc;
    NoiseDev nz; make_noise(noise, 0, &nz);
    const uint32_t b = (uint32_t)ctx->int_bits;
    const int grid = grid_1d(ctx, span->count, 256, 16);
    const bool v4 = ctx->words == 1 && (span->begin & 3ull) == 0 &&
                    ((((uintptr_t)x | (uintptr_t)mask | (uintptr_t)ct_out | (uintptr_t)nz.u) & 15u) == 0);
    if (v4) {
        const uint64_t nvec = span->count / 4, done = nvec * 4;
        if (nvec) k_encode_premasked_v4<<<grid_1d(ctx, nvec, 256, 8), 256, 0, cs>>>((const uint4*)x, (const uint4*)mask, span->begin, nvec, Word<1>::mask(b), ch.dev, nz, (uint4*)ct_out);
        if (done < span->count) {
            NoiseDev nt = nz; if (nt.u) nt.u += done;
            k_encode<1, true><<<1, 32, 0, cs>>>(x + done, (const uint32_t*)mask + done, span->begin + done, span->count - done, b, ch.dev, nt, nullptr, (uint32_t*)ct_out + done);
            g_launches.fetch_add(nvec ? 1 : 0);
        }
    }
    else if (ctx->words == 1) k_encode<1, true><<<grid, 256, 0, cs>>>(x, (const uint32_t*)mask, span->begin, span->count, b, ch.dev, nz, nullptr, (uint32_t*)ct_out);
    else if (ctx->words == 2) k_encode<2, true><<<grid, 256, 0, cs>>>(x, (const uint64_t*)mask, span->begin, span->count, b, ch.dev, nz, nullptr, (uint64_t*)ct_out);
    else k_encode<4, true><<<grid, 256, 0, cs>>>(x, (const u128*)mask, span->begin, span->count, b, ch.dev, nz, nullptr, (u128*)ct_out);
    g_launches.fetch_add(1);
    free_codec(&ch, cs);
    CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}

int flashe_aggregate(flashe_ctx* ctx, const void* cts, uint64_t stride, int n, uint64_t count, int mode, uint32_t carry_in,
                     void* out, uint32_t* carry_out, void* stream) {
    ENTER(ctx);
    if (n < 1) return fail(FLASHE_EINVAL, "n must be >= 1");
    if (mode != FLASHE_AGG_ELEMENTWISE && mode != FLASHE_AGG_PACKED) return fail(FLASHE_EINVAL, "unknown aggregate mode");
    if (count == 0) {
        if (mode == FLASHE_AGG_PACKED && carry_out) {
            // empty range: the carry passes through unchanged: cin -> 0 + (cin >= 1) only holds for
            // cin <= 1, so report it as "given carry, no dependence" and let callers skip empty shards
            uint32_t d[4] = {carry_in, 0u, carry_in, T_NEVER};
            CUDA_TRY(cudaMemcpyAsync(carry_out, d, sizeof(d), cudaMemcpyHostToDevice, cs));
        }
        return FLASHE_OK;
    }
    if (!cts || !out) return fail(FLASHE_EINVAL, "NULL buffer");
    if (n > 1 && stride < count) return fail(FLASHE_EINVAL, "stride must be >= count");
    const uint32_t b = (uint32_t)ctx->int_bits;
    const int wb = 4 * ctx->words;
    if (mode == FLASHE_AGG_ELEMENTWISE) {
        const int per_vec = 16 / wb;
        const bool aligned = (((uintptr_t)cts | (uintptr_t)out) & 15u) == 0 && (stride % (uint64_t)per_vec) == 0;
        const uint64_t nvec = aligned ? count / per_vec : 0;
        if (nvec) {
            const int grid = grid_1d(ctx, nvec, 256, 8);
            const uint64_t sv = stride / per_vec;
            if (ctx->words == 1) k_aggregate_vec<1><<<grid, 256, 0, cs>>>((const uint4*)cts, sv, n, nvec, b, (uint4*)out);
            else if (ctx->words == 2) k_aggregate_vec<2><<<grid, 256, 0, cs>>>((const uint4*)cts, sv, n, nvec, b, (uint4*)out);
            else k_aggregate_vec<4><<<grid, 256, 0, cs>>>((const uint4*)cts, sv, n, nvec, b, (uint4*)out);
            g_launches.fetch_add(1);
        }
        const uint64_t done = nvec * per_vec;
        if (done < count) {
            const int grid = grid_1d(ctx, count - done, 256, 8);
            if (ctx->words == 1) k_aggregate_scalar<1><<<grid, 256, 0, cs>>>((const uint32_t*)cts, stride, n, done, count, b, (uint32_t*)out);
            else if (ctx->words == 2) k_aggregate_scalar<2><<<grid, 256, 0, cs>>>((const uint64_t*)cts, stride, n, done, count, b, (uint64_t*)out);
            else k_aggregate_scalar<4><<<grid, 256, 0, cs>>>((const u128*)cts, stride, n, done, count, b, (u128*)out);
            g_launches.fetch_add(1);
        }
    } else {
        // The carry transfer of one digit is modelled as cin -> A + (cin >= T): at most +1, which needs
        // cin <= n - 1 < 2^b (with more clients than digit values a digit could hand on +2).
        if (b < 31u && (uint64_t)n > (1ull << b))
            return fail(FLASHE_EINVAL, "packed-carry aggregate needs n <= 2^int_bits");
        const uint64_t per_tile = (uint64_t)PK_THREADS * (ctx->words == 4 ? PkElems<4>::V : PkElems<1>::V);
        const uint64_t ntiles = ceil_div(count, per_tile);
        uint64_t cap = (uint64_t)ctx->num_sms * 8;
        const int grid = (int)(ntiles < cap ? ntiles : cap);
        // 1: rows are 128-bit columns; 2: the output too
        const int vec_ok = (ctx->words == 1 && (count & 3u) == 0 && (stride & 3u) == 0 && aligned16(cts)) ? (aligned16(out) ? 2 : 1) : 0;
        if (ctx->words == 4 && !(aligned16(cts) && aligned16(out))) return fail(FLASHE_EINVAL, "16-byte words must be 16-byte aligned");
        if (ctx->words == 1) k_aggregate_packed<1><<<grid, PK_THREADS, 0, cs>>>((const uint32_t*)cts, stride, n, count, b, carry_in, (uint32_t*)out, carry_out, vec_ok);
        else if (ctx->words == 2) k_aggregate_packed<2><<<grid, PK_THREADS, 0, cs>>>((const uint64_t*)cts, stride, n, count, b, carry_in, (uint64_t*)out, carry_out, 0);
        else k_aggregate_packed<4><<<grid, PK_THREADS, 0, cs>>>((const u128*)cts, stride, n, count, b, carry_in, (u128*)out, carry_out, 0);
        g_launches.fetch_add(1);
    }
    CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}

int flashe_aggregate_carry_fixup(flashe_ctx* ctx, void* out, uint64_t count, uint32_t carry_in, void* stream) {
    ENTER(ctx);
    if (count == 0 || carry_in == 0) return FLASHE_OK;
    if (!out) return fail(FLASHE_EINVAL, "out is NULL");
    if (ctx->words == 1) k_carry_fixup<1><<<1, 32, 0, cs>>>((uint32_t*)out, count, (uint32_t)ctx->int_bits, carry_in);
    else if (ctx->words == 2) k_carry_fixup<2><<<1, 32, 0, cs>>>((uint64_t*)out, count, (uint32_t)ctx->int_bits, carry_in);
    else k_carry_fixup<4><<<1, 32, 0, cs>>>((u128*)out, count, (uint32_t)ctx->int_bits, carry_in);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}

int flashe_decode(flashe_ctx* ctx, const flashe_span* span, const void* v, const flashe_codec* codec, double* out, void* stream) {
    ENTER(ctx);
    int rc = check_span(span); if (rc) return rc;
    if (ctx->words == 4) return fail(FLASHE_EUNSUPPORTED, "decode takes int_bits <= 64 (unbatch 128-bit words first)");
    if (span->count == 0) return FLASHE_OK;
    if (!v || !out) return fail(FLASHE_EINVAL, "NULL buffer");
    CodecHost ch; rc = make_codec(ctx, span, codec, true, cs, &ch); if (rc) return rc;
    const int grid = grid_1d(ctx, span->count, 256, 16);
    if (ctx->words == 1 && (((uintptr_t)v | (uintptr_t)out) & 15u) == 0) {
        const uint64_t nvec = span->count / 4, done = nvec * 4;
        if (nvec) k_decode_v4<<<grid_1d(ctx, nvec, 256, 8), 256, 0, cs>>>((const uint4*)v, span->begin, nvec, ch.dev, out);
        if (done < span->count) {
            k_decode<1><<<1, 32, 0, cs>>>((const uint32_t*)v + done, span->begin + done, span->count - done, ch.dev, out + done);
            g_launches.fetch_add(1);
        }
        if (!nvec) g_launches.fetch_sub(1);
    }
    else if (ctx->words == 1) k_decode<1><<<grid, 256, 0, cs>>>((const uint32_t*)v, span->begin, span->count, ch.dev, out);
    else k_decode<2><<<grid, 256, 0, cs>>>((const uint64_t*)v, span->begin, span->count, ch.dev, out);
    g_launches.fetch_add(1);
    free_codec(&ch, cs);
    CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}

int flashe_decrypt_decode(flashe_ctx* ctx, uint32_t iter, const int32_t* add_idx, int na, const int32_t* minus_idx, int ns,
                          const flashe_span* span, const void* agg_in, const flashe_codec* codec, double* out, void* p_out,
                          void* stream) {
    ENTER(ctx);
    int rc = check_span(span); if (rc) return rc;
    if (ctx->words == 4) return fail(FLASHE_EUNSUPPORTED, "decrypt_decode takes int_bits <= 64 (decrypt, unbatch, decode for 128-bit words)");
    int32_t prf[MAXS], sg[MAXS];
    rc = build_decrypt_streams(add_idx, na, minus_idx, ns, prf, sg); if (rc) return rc;
    if (span->count == 0) return FLASHE_OK;
    if (!agg_in || !out) return fail(FLASHE_EINVAL, "NULL buffer");
    StreamTab st; rc = make_streams(ctx, iter, prf, sg, na + ns, &st); if (rc) return rc;
    Geom g; make_geom(ctx, span, pick_sup(ctx, span, 1), &g);
    CodecHost ch; rc = make_codec(ctx, span, codec, true, cs, &ch); if (rc) return rc;
    IoDev io; memset(&io, 0, sizeof(io)); io.in = agg_in; io.outf = out; io.aux = p_out; io.n_clients = 1;
    NoiseDev nz; memset(&nz, 0, sizeof(nz));
    rc = launch_stream<M_DECODE>(ctx, st, g, io, ch.dev, nz, cs);
    free_codec(&ch, cs);
    return rc;
}

int flashe_rng_uniform(flashe_ctx* ctx, uint64_t rng_seed, uint64_t rng_stream, uint64_t begin, uint64_t count, double* out, void* stream) {
    ENTER(ctx);
    if (count == 0) return FLASHE_OK;
    if (!out) return fail(FLASHE_EINVAL, "out is NULL");
    flashe_noise n; n.u = nullptr; n.rng_seed = rng_seed; n.rng_stream = rng_stream;
    NoiseDev nz; make_noise(&n, 0, &nz);
    k_rng_uniform<<<grid_1d(ctx, count, 256, 16), 256, 0, cs>>>(nz, begin, count, out);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}

static int batch_geometry(const flashe_ctx* ctx, int element_bits, int factor, uint32_t* lane, uint32_t* bs) {
    if (ctx->words != 4) return fail(FLASHE_EUNSUPPORTED, "lane batching is built for 64 < int_bits <= 128 (shipped: 120)");
    const int l = element_bits + factor;
    if (element_bits < 1 || factor < 0 || l > 32) return fail(FLASHE_EINVAL, "element_bits + factor must be in [1, 32]");
    *lane = (uint32_t)l; *bs = (uint32_t)(ctx->int_bits / l);
    if (*bs == 0) return fail(FLASHE_EINVAL, "int_bits smaller than one lane");
    return FLASHE_OK;
}

int flashe_batch_pack(flashe_ctx* ctx, const uint32_t* q, uint64_t count, int element_bits, int factor, void* words_out, void* stream) {
    ENTER(ctx);
    uint32_t lane, bs; int rc = batch_geometry(ctx, element_bits, factor, &lane, &bs); if (rc) return rc;
    if (count == 0) return FLASHE_OK;
    if (!q || !words_out) return fail(FLASHE_EINVAL, "NULL buffer");
    const uint64_t nw = ceil_div(count, bs);
    k_batch_pack<<<grid_1d(ctx, nw, 256, 16), 256, 0, cs>>>(q, count, lane, bs, nw, (u128*)words_out);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}

int flashe_batch_unpack(flashe_ctx* ctx, const void* words, uint64_t nwords, int element_bits, int factor, uint32_t* q_out, void* stream) {
    ENTER(ctx);
    uint32_t lane, bs; int rc = batch_geometry(ctx, element_bits, factor, &lane, &bs); if (rc) return rc;
    if (nwords == 0) return FLASHE_OK;
    if (!words || !q_out) return fail(FLASHE_EINVAL, "NULL buffer");
    k_batch_unpack<<<grid_1d(ctx, nwords, 256, 16), 256, 0, cs>>>((const u128*)words, nwords, lane, bs, q_out);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}

int flashe_sparse_sum(flashe_ctx* ctx, const void* const* compacts, const int64_t* const* indexes, const uint64_t* ks,
                      const void* zero_words, int n_clients, uint64_t total, void* dense_out, void* stream) {
    ENTER(ctx);
    if (n_clients < 1 || !ks || !zero_words) return fail(FLASHE_EINVAL, "need n_clients >= 1, ks and zero_words");
    if (total == 0) return FLASHE_OK;
    if (!dense_out) return fail(FLASHE_EINVAL, "dense_out is NULL");
    for (int c = 0; c < n_clients; ++c) {
        if (ks[c] > total) return fail(FLASHE_EINVAL, "k exceeds total");
        if (ks[c] && (!compacts || !indexes || !compacts[c] || !indexes[c])) return fail(FLASHE_EINVAL, "NULL compact / index buffer");
    }
    if (ctx->words == 1) return sparse_sum_t<1>(ctx, compacts, indexes, ks, zero_words, n_clients, total, dense_out, cs);
    if (ctx->words == 2) return sparse_sum_t<2>(ctx, compacts, indexes, ks, zero_words, n_clients, total, dense_out, cs);
    return sparse_sum_t<4>(ctx, compacts, indexes, ks, zero_words, n_clients, total, dense_out, cs);
}

int flashe_sparse_expand(flashe_ctx* ctx, const void* compact, const int64_t* index, uint64_t k, uint64_t total, const void* zero_word,
                         void* dense_out, void* stream) {
    ENTER(ctx);
    if (total == 0) return FLASHE_OK;
    if (!dense_out || !zero_word || (k && (!compact || !index))) return fail(FLASHE_EINVAL, "NULL buffer");
    if (k > total) return fail(FLASHE_EINVAL, "k exceeds total");
    const int gf = grid_1d(ctx, total, 256, 16), gs = grid_1d(ctx, k, 256, 16);
    if (ctx->words == 1) {
        uint32_t z; memcpy(&z, zero_word, 4);
        k_fill<1><<<gf, 256, 0, cs>>>((uint32_t*)dense_out, total, z);
        if (k) k_scatter<1><<<gs, 256, 0, cs>>>((const uint32_t*)compact, index, k, total, (uint32_t*)dense_out);
    } else if (ctx->words == 2) {
        uint64_t z; memcpy(&z, zero_word, 8);
        k_fill<2><<<gf, 256, 0, cs>>>((uint64_t*)dense_out, total, z);
        if (k) k_scatter<2><<<gs, 256, 0, cs>>>((const uint64_t*)compact, index, k, total, (uint64_t*)dense_out);
    } else {
        u128 z; memcpy(&z, zero_word, 16);
        k_fill<4><<<gf, 256, 0, cs>>>((u128*)dense_out, total, z);
        if (k) k_scatter<4><<<gs, 256, 0, cs>>>((const u128*)compact, index, k, total, (u128*)dense_out);
    }
    g_launches.fetch_add(k ? 2 : 1);
    CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}

int flashe_sparse_apply_masks(flashe_ctx* ctx, uint32_t iter, const int32_t* prf_idx, const int32_t* sign, int nstreams,
                              const flashe_span* span, const int64_t* index, void* dense, uint64_t dense_len, void* stream) {
    ENTER(ctx);
    int rc = check_span(span); if (rc) return rc;
    if (span->count == 0) return FLASHE_OK;
    if (!index || !dense) return fail(FLASHE_EINVAL, "NULL buffer");
    if (span->count > dense_len) return fail(FLASHE_EINVAL, "more compact positions than dense words");
    StreamTab st; rc = make_streams(ctx, iter, prf_idx, sign, nstreams, &st); if (rc) return rc;
    Geom g; make_geom(ctx, span, pick_sup(ctx, span, 1), &g);
    IoDev io; memset(&io, 0, sizeof(io)); io.out = dense; io.aux = (void*)index; io.n_clients = 1; io.dense_len = dense_len;
    CodecDev cd; memset(&cd, 0, sizeof(cd)); NoiseDev nz; memset(&nz, 0, sizeof(nz));
    return launch_stream<M_SCATTER>(ctx, st, g, io, cd, nz, cs);
}

int flashe_sparse_overlap(flashe_ctx* ctx, const int64_t* const* index, const uint64_t* k, int n, uint64_t total, uint64_t* overlap_out,
                          void* stream) {
    ENTER(ctx);
    (void)total;
    if (n < 1 || !index || !k) return fail(FLASHE_EINVAL, "bad arguments");
    if (n == 1) return FLASHE_OK;
    if (!overlap_out) return fail(FLASHE_EINVAL, "overlap_out is NULL");
    unsigned long long* d = nullptr;
    CUDA_TRY(cudaMallocAsync((void**)&d, sizeof(unsigned long long) * (size_t)(n - 1), cs));
    cudaError_t e = cudaMemsetAsync(d, 0, sizeof(unsigned long long) * (size_t)(n - 1), cs);
    for (int i = 0; e == cudaSuccess && i + 1 < n; ++i) {
        if (k[i] == 0 || k[i + 1] == 0) continue;
        k_overlap<<<grid_1d(ctx, k[i], 256, 16), 256, 0, cs>>>(index[i], k[i], index[i + 1], k[i + 1], d + i);
        g_launches.fetch_add(1);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(overlap_out, d, sizeof(unsigned long long) * (size_t)(n - 1), cudaMemcpyDeviceToHost, cs);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cs);
    cudaFreeAsync(d, cs);
    if (e != cudaSuccess) return fail(FLASHE_ECUDA, std::string("flashe_sparse_overlap: ") + cudaGetErrorString(e));
    return FLASHE_OK;
}

}  // extern "C"
