// flashe_device.cuh — device-side types and functions shared by the translation units of
// libflashe_b200.so: storage words, layer (segment) tables, encode / decode arithmetic, the counter-based
// noise generator, 128-bit global accesses.  Reference lines are cited at each function.
#ifndef FLASHE_DEVICE_CUH
#define FLASHE_DEVICE_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#define MAX_INLINE_SEG 48

struct Seg { uint64_t end; float a, two_a; float rcp_two_a, pad; double an, two_an; uint64_t wend; };  // rcp_two_a = RN(1/two_a), 0 = not usable;
                                                   // wend: cumulative words up to and including this layer (lane batching)

struct CodecDev {
    int32_t nseg;
    int32_t ebits;
    float scale;             // 2^e - 1 as float32
    double den;              // (2^e - 1) * n as float64
    double den_rcp;          // RN(1 / den), or 0: use the library division
    uint32_t lane_bits;      // lane batching (jzf_quantize.py:162-185): lane width, 0 = one element per word
    uint32_t bs;             // lanes per word = int_bits / lane_bits
    const Seg* table;        // device table when nseg > MAX_INLINE_SEG, else NULL
    Seg seg[MAX_INLINE_SEG];
};

struct NoiseDev { const double* u; uint64_t u_stride; uint64_t stream; uint32_t rk[10][2]; uint32_t res32; uint32_t pad; };  // rk: Philox round keys; res32: one 32-bit word per element


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
        r.an = __ldg(&t->an); r.two_an = __ldg(&t->two_an); r.wend = __ldg(&t->wend);
        return r;
    }
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (j < c.seg[mid].end) hi = mid; else lo = mid + 1;
    }
    return c.seg[lo];
}

// Lane batching: the layer that owns WORD jw of the batched vector, the first element of that word and the
// end of the layer (elements of the word at or past `eend` are the zero padding of jzf_quantize.py:166-171).
struct WordPos { Seg sg; uint64_t e0, eend; };
__device__ __forceinline__ WordPos locate_word(const CodecDev& c, uint64_t jw) {
    WordPos p;
    if (c.nseg == 1) { p.sg = c.seg[0]; p.e0 = jw * c.bs; p.eend = p.sg.end; return p; }
    int lo = 0, hi = c.nseg - 1;
    uint64_t ebeg = 0, wbeg = 0;
    if (c.table) {
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (jw < __ldg(&c.table[mid].wend)) hi = mid; else lo = mid + 1; }
        const Seg* t = c.table + lo;
        p.sg.end = __ldg(&t->end); p.sg.a = __ldg(&t->a); p.sg.two_a = __ldg(&t->two_a); p.sg.rcp_two_a = __ldg(&t->rcp_two_a); p.sg.pad = 0.f;
        p.sg.an = __ldg(&t->an); p.sg.two_an = __ldg(&t->two_an); p.sg.wend = __ldg(&t->wend);
        if (lo) { ebeg = __ldg(&t[-1].end); wbeg = __ldg(&t[-1].wend); }
    } else {
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (jw < c.seg[mid].wend) hi = mid; else lo = mid + 1; }
        p.sg = c.seg[lo];
        if (lo) { ebeg = c.seg[lo - 1].end; wbeg = c.seg[lo - 1].wend; }
    }
    p.e0 = ebeg + (jw - wbeg) * c.bs; p.eend = p.sg.end;
    return p;
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
// Throughput-mode resolution (flashe_noise.resolution = FLASHE_NOISE_32): u = w * 2^-32 from ONE 32-bit word, built
// as the bit pattern of 1 + w * 2^-32 minus 1 (exact).  One Philox call then serves four elements instead of two.
__device__ __forceinline__ double unit32(uint32_t w) {
    return __dadd_rn(__hiloint2double((int)(0x3FF00000u | (w >> 12)), (int)(w << 20)), -1.0);
}
// u_j in [0,1).  53-bit resolution: counter (j>>1 lo, j>>1 hi, stream lo, stream hi), words (2(j&1), 2(j&1)+1) feed
// res53.  32-bit resolution (N32): counter (j>>2, stream), word j&3.  The resolution is a TEMPLATE parameter: the
// encode loop of k_stream is instruction-cache sensitive, a run-time branch around a second generator cost it 5 %.
template <bool N32>
__device__ __forceinline__ double noise_one(const NoiseDev& nz, uint64_t stream, uint64_t j) {
    uint32_t o[4];
    if (N32) {
        const uint64_t c = j >> 2;
        philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)stream, (uint32_t)(stream >> 32), nz, o);
        const uint32_t k = (uint32_t)j & 3u;
        return unit32(k == 0u ? o[0] : (k == 1u ? o[1] : (k == 2u ? o[2] : o[3])));
    }
    uint64_t c = j >> 1;
    philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)stream, (uint32_t)(stream >> 32), nz, o);
    return (j & 1) ? res53(o[2], o[3]) : res53(o[0], o[1]);
}

// u for elements 2c and 2c+1 (same values as noise_one).
template <bool N32>
__device__ __forceinline__ void noise_pair(const NoiseDev& nz, uint64_t stream, uint64_t c, double& u0, double& u1) {
    uint32_t o[4];
    if (N32) {
        const uint64_t q = c >> 1;
        philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)stream, (uint32_t)(stream >> 32), nz, o);
        const bool hi = (c & 1ull) != 0ull;
        u0 = unit32(hi ? o[2] : o[0]);
        u1 = unit32(hi ? o[3] : o[1]);
        return;
    }
    philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)stream, (uint32_t)(stream >> 32), nz, o);
    u0 = res53(o[0], o[1]);
    u1 = res53(o[2], o[3]);
}
// u for elements 4q .. 4q+3: one Philox call at 32-bit resolution, two at 53-bit.
template <bool N32>
__device__ __forceinline__ void noise_quad(const NoiseDev& nz, uint64_t stream, uint64_t q, double (&u)[4]) {
    if (N32) {
        uint32_t o[4];
        philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)stream, (uint32_t)(stream >> 32), nz, o);
        u[0] = unit32(o[0]); u[1] = unit32(o[1]); u[2] = unit32(o[2]); u[3] = unit32(o[3]);
        return;
    }
    noise_pair<false>(nz, stream, 2ull * q, u[0], u[1]);
    noise_pair<false>(nz, stream, 2ull * q + 1ull, u[2], u[3]);
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


__device__ __forceinline__ void ldg_v4(const void* p, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
}
__device__ __forceinline__ void stg_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// 32-byte global accesses (sm_100: LDG.256 / STG.256), streaming (no L1 allocation); the address must be 32-byte aligned
__device__ __forceinline__ void ldg_v8(const void* p, uint32_t (&v)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p));
}
__device__ __forceinline__ void stg_v8(void* p, const uint32_t (&v)[8]) {
    asm volatile("st.global.L1::no_allocate.v8.u32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
// four float64 as ONE 32-byte store (sm_100: STG.256).  A warp then writes 1 KB of whole sectors per instruction — what a
// peer GPU's memory behind NVLink wants (two 16-byte stores per lane arrive there as half-filled sectors).
__device__ __forceinline__ void stg_d4(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void stg_d2(double* p, double a, double b) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}


#endif
