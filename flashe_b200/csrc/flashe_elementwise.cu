// flashe_elementwise.cu — the HBM-bound kernels of the FLASHE hot path and their C ABI entries: server sums
// (element-wise and packed-carry, proc/jzf_aggregator.py:404-430), the online step after mask precomputation,
// stand-alone encode / decode, lane batching, sparse expand / sum / overlap.  (The PRF stream kernel lives in
// flashe_kernels.cu, the wire format / top-k / statistics kernels in flashe_wire.cu.)
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/flashe_b200.h"
#include "flashe_internal.h"
#include "flashe_device.cuh"
#include "flashe_codec_host.cuh"

#define fail flashe_fail
static inline void count_launch(int n = 1) { flashe_count_launches(n); }

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

// the same with 32-byte accesses: 8 elements per thread
__global__ void k_add_premasked_v8(const uint32_t* __restrict__ in, const uint32_t* __restrict__ mask, int sign, uint64_t nvec,
                                   uint32_t mk, uint32_t* __restrict__ out) {
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t a[8], k[8], r[8];
        ldg_v8(in + 8ull * v, a); ldg_v8(mask + 8ull * v, k);
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = (sign >= 0 ? a[i] + k[i] : a[i] - k[i]) & mk;
        stg_v8(out + 8ull * v, r);
    }
}

template <int WORDS, bool WITH_MASK, bool N32 = false>
__global__ void k_encode(const float* __restrict__ x, const typename Word<WORDS>::T* __restrict__ mask, uint64_t begin,
                         uint64_t count, uint32_t b, const __grid_constant__ CodecDev cd, const __grid_constant__ NoiseDev nz,
                         uint32_t* __restrict__ q_out, typename Word<WORDS>::T* __restrict__ ct_out) {
    typedef Word<WORDS> WT;
    const typename WT::T mk = WT::mask(b);
    for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < count; o += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t j = begin + o;
        const Seg sg = find_seg(cd, j);
        double u = nz.u ? nz.u[o] : noise_one<N32>(nz, nz.stream, j);
        uint32_t q = encode_one(x[o], u, sg, cd.scale);
        if (q_out) q_out[o] = q;
        if (WITH_MASK) ct_out[o] = WT::band(WT::add(WT::from_u32(q), mask[o]), mk);
    }
}

// Online step after mask precomputation, 4-byte words: one thread = 4 consecutive elements (begin and
// every pointer 16-byte aligned), 128-bit loads of x and of the precomputed mask, two Philox calls for
// the four noise values, one 128-bit store.  12 algorithmic bytes per element: HBM-bound.
// n_clients rows in one launch: blockIdx.y = client (row c: x + c*xs, mask + c*ms, ct_out + c*cs vectors, noise stream
// id + c).
template <bool N32>
__global__ void __launch_bounds__(256)
k_encode_premasked_v4(const uint4* __restrict__ x, const uint4* __restrict__ mask, uint64_t begin, uint64_t nvec, uint32_t mk,
                      const __grid_constant__ CodecDev cd, const __grid_constant__ NoiseDev nz, uint4* __restrict__ ct_out,
                      uint64_t xs, uint64_t ms, uint64_t cs, uint64_t us) {
    const bool one_seg = cd.nseg == 1, one_rcp = one_seg && cd.seg[0].rcp_two_a != 0.0f;
    const uint32_t c = blockIdx.y;
    x += c * xs; mask += c * ms; ct_out += c * cs;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t j = begin + 4ull * v;
        const uint4 xv = __ldcs(x + v), mv = __ldcs(mask + v);
        double u[4];
        if (nz.u) {
            const double2* up = reinterpret_cast<const double2*>(nz.u + c * us) + 2 * v;
            const double2 a = __ldcs(up), b = __ldcs(up + 1);
            u[0] = a.x; u[1] = a.y; u[2] = b.x; u[3] = b.y;
        } else {
            noise_quad<N32>(nz, nz.stream + c, j >> 2, u);                  // begin is a multiple of 4
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

// 32-byte accesses, 8 elements per thread (device noise only): twice the bytes in flight per thread and half the
// address arithmetic of the 16-byte form
template <bool N32>
__global__ void __launch_bounds__(256)
k_encode_premasked_v8(const uint32_t* __restrict__ x, const uint32_t* __restrict__ mask, uint64_t begin, uint64_t nvec, uint32_t mk,
                      const __grid_constant__ CodecDev cd, const __grid_constant__ NoiseDev nz, uint32_t* __restrict__ ct_out,
                      uint64_t xs, uint64_t ms, uint64_t cs) {
    const bool one_seg = cd.nseg == 1, one_rcp = one_seg && cd.seg[0].rcp_two_a != 0.0f;
    const uint32_t c = blockIdx.y;
    x += c * xs; mask += c * ms; ct_out += c * cs;                   // (strides in elements)
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t j = begin + 8ull * v;                           // begin is a multiple of 8
        uint32_t xr[8], mv[8], q[8];
        ldg_v8(x + 8ull * v, xr); ldg_v8(mask + 8ull * v, mv);
        double u[8];
        {
            double ua[4], ub[4];
            noise_quad<N32>(nz, nz.stream + c, j >> 2, ua);
            noise_quad<N32>(nz, nz.stream + c, (j >> 2) + 1, ub);
#pragma unroll
            for (int k = 0; k < 4; ++k) { u[k] = ua[k]; u[4 + k] = ub[k]; }
        }
        if (one_rcp) {                                                 // single layer with a usable reciprocal (uniform)
#pragma unroll
            for (int k = 0; k < 8; ++k) q[k] = encode_one<true>(__uint_as_float(xr[k]), u[k], cd.seg[0], cd.scale);
        } else {
            Seg sg = find_seg(cd, j);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (k && !one_seg && j + k >= sg.end) sg = find_seg(cd, j + k);
                q[k] = encode_one(__uint_as_float(xr[k]), u[k], sg, cd.scale);
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) q[k] = (q[k] + mv[k]) & mk;
        stg_v8(ct_out + 8ull * v, q);
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
    const bool out32 = (reinterpret_cast<uintptr_t>(out) & 31u) == 0u;     // 32-byte stores (one per four elements)
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    // two independent 16-byte loads in flight per thread (the kernel is pure streaming: 4 B in, 8 B out per element)
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < nvec; i0 += 2 * stride) {
        const uint64_t i1 = i0 + stride;
        const bool two = i1 < nvec;
        const uint4 wa = __ldcs(v + i0);
        uint4 wb = make_uint4(0u, 0u, 0u, 0u);
        if (two) wb = __ldcs(v + i1);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (h && !two) break;
            const uint64_t i = h ? i1 : i0;
            const uint4 w = h ? wb : wa;
            const uint32_t p[4] = {w.x, w.y, w.z, w.w};
            const uint64_t j = begin + 4ull * i;
            double d[4];
            Seg sg = find_seg(cd, j);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k && !one_seg && j + k >= sg.end) sg = find_seg(cd, j + k);
                d[k] = decode_one((double)p[k], sg.two_an, cd.den, cd.den_rcp, sg.an);
            }
            if (out32) stg_d4(out + 4ull * i, d[0], d[1], d[2], d[3]);
            else { stg_d2(out + 4ull * i, d[0], d[1]); stg_d2(out + 4ull * i + 2, d[2], d[3]); }
        }
    }
}

template <bool N32>
__global__ void k_rng_uniform(const __grid_constant__ NoiseDev nz, uint64_t begin, uint64_t count, double* __restrict__ out) {
    for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < count; o += (uint64_t)gridDim.x * blockDim.x)
        out[o] = noise_one<N32>(nz, nz.stream, begin + o);
}

// Element-wise server sum (jzf_aggregator.py:421-430).  One thread owns one 16-byte column of the
// [n][count] matrix and walks the n client rows with UNROLL independent 128-bit loads in flight.
template <int WORDS>
__global__ void __launch_bounds__(256)
k_aggregate_vec(const uint4* __restrict__ cts, uint64_t stride_vec, int n, uint64_t nvec, uint32_t b, uint4* __restrict__ out) {
    asm volatile("griddepcontrol.launch_dependents;");   // a following k_stream may build its tables meanwhile (it waits before reading)
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
        // first look-ahead element of the tile (thread 0 only): issued here so that its latency overlaps the
        // column loads above instead of following the warp scan
        lo_t la_lo = lo_zero<WORDS>(); uint32_t la_H = 0;
        if (threadIdx.x == 0 && tile > 0) digit_sum<WORDS>(cts, stride, n, count - tile * tile_elems, b, la_lo, la_H);
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

        // look-ahead for the tile's carry-in (elements closer to the end than this tile); its first element
        // was loaded up front (la_lo / la_H), together with the thread's own column
        if (threadIdx.x == 0) {
            uint32_t cin;
            if (tile == 0) cin = carry_in;
            else {
                // compose f_{j} for j just after the tile, walking towards the end, until constant
                const uint64_t first_back = tile * tile_elems;  // `back` of this tile's first element
                Xfer acc = xfer_of(la_lo, la_H, b);             // element first_back - 1
                uint64_t bk = first_back - 1;                   // walk bk-1, bk-2, ... 0
                while (acc.T != T_NEVER && bk > 0) {
                    --bk;
                    lo_t l; uint32_t h;
                    digit_sum<WORDS>(cts, stride, n, count - 1 - bk, b, l, h);
                    // acc currently maps (carry into element bk+1.. chain) ; new element is applied FIRST
                    acc = xfer_compose(acc, xfer_of(l, h, b));
                }
                cin = acc.T == T_NEVER ? acc.A : xfer_apply(acc, carry_in);
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

// Per-layer lane batching through a segment table: QuantizingClient.quantize batches every
// layer by itself (jzf_quantize.py:447-452), so each layer is zero-padded to a multiple of batch_size and the
// flat word vector is the concatenation of the layers' words (flatten_weights, jzf_aggregator.py:625-650).
// ebeg / wbeg: first element / first word of layer s (nseg + 1 entries).
__device__ __forceinline__ int layer_of_word(const uint64_t* __restrict__ wbeg, int nseg, uint64_t w) {
    int lo = 0, hi = nseg - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (__ldg(wbeg + mid) <= w) lo = mid; else hi = mid - 1; }
    return lo;
}
__global__ void k_batch_pack_layers(const uint32_t* __restrict__ q, const uint64_t* __restrict__ ebeg, const uint64_t* __restrict__ wbeg,
                                    int nseg, uint32_t lane_bits, uint32_t bs, uint64_t nwords, u128* __restrict__ out) {
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (uint64_t)gridDim.x * blockDim.x) {
        const int s = layer_of_word(wbeg, nseg, w);
        const uint64_t e0 = __ldg(ebeg + s) + (w - __ldg(wbeg + s)) * bs, eend = __ldg(ebeg + s + 1);
        uint64_t lo = 0, hi = 0;
        for (uint32_t i = 0; i < bs; ++i) {
            const uint64_t v = e0 + i < eend ? q[e0 + i] : 0u;
            hi = (hi << lane_bits) | (lo >> (64 - lane_bits));
            lo = (lo << lane_bits) + v;
        }
        u128 r; r.lo = lo; r.hi = hi;
        out[w] = r;
    }
}
__global__ void k_batch_unpack_layers(const u128* __restrict__ in, const uint64_t* __restrict__ ebeg, const uint64_t* __restrict__ wbeg,
                                      int nseg, uint32_t lane_bits, uint32_t bs, uint64_t nwords, uint32_t* __restrict__ out) {
    const uint64_t lm = (1ull << lane_bits) - 1ull;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (uint64_t)gridDim.x * blockDim.x) {
        const int s = layer_of_word(wbeg, nseg, w);
        const uint64_t e0 = __ldg(ebeg + s) + (w - __ldg(wbeg + s)) * bs, eend = __ldg(ebeg + s + 1);
        u128 t = in[w];
        for (int i = (int)bs - 1; i >= 0; --i) {
            if (e0 + i < eend) out[e0 + i] = (uint32_t)(t.lo & lm);         // padding lanes are dropped
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

// ---- the arbiter's sparse sum and the overlap counts by TILES of the dense vector ----------------------------------
// The index lists are sorted, so the entries of one client that fall into a tile of the dense vector are one
// contiguous run of its list.  k_sparse_splits finds the runs (one binary search per client and tile boundary);
// k_sparse_sum_tiled then builds every tile of the output in shared memory - start from the sum of the zero words, add
// (compact - zero) for the entries of all clients with shared-memory atomics - and writes it ONCE, coalesced: total
// words out + sum k_c (index + word) in, instead of one read-modify-write of a 32-byte sector per entry and one launch
// per client (32 clients x 1 % of 50 M: 0.62 ms -> see profiles/).  The overlap counts of dynamic_masking use the same
// runs: an entry of client i searches only client i+1's run of the same tile.
struct SparseClient { const void* compact; const int64_t* index; uint64_t k; uint64_t zero_lo, zero_hi; };
// Up to SPARSE_PARAM_CLIENTS clients travel in the kernel parameters (no table upload: a cudaMemcpyAsync from pageable
// memory would hold the host until the stream reaches it); more clients go through a table in global memory.
#define SPARSE_PARAM_CLIENTS 64
struct SparseClientsParam { SparseClient c[SPARSE_PARAM_CLIENTS]; };
#define SPARSE_TILE_BYTES 32768u
#ifndef SP_DEPTH
#define SP_DEPTH 6u        // measured with SP_MINB (profiles/r4m_ab_sparse_variants.txt): 6 loads in flight x 4 CTAs per SM
#endif

__global__ void k_sparse_splits(const SparseClient* __restrict__ cl_g, const __grid_constant__ SparseClientsParam P, int n, uint64_t total,
                                uint32_t tile_log2, uint64_t n_tiles, uint32_t* __restrict__ splits) {
    const SparseClient* __restrict__ cl = cl_g ? cl_g : P.c;
    const uint64_t per = n_tiles + 1, all = per * (uint64_t)n;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < all; g += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t c = g / per, t = g - c * per;
        const int64_t bound = t == n_tiles ? (int64_t)total : (int64_t)(t << tile_log2);    // first index of tile t
        const int64_t* __restrict__ idx = cl[c].index;
        uint64_t lo = 0, hi = cl[c].k;
        while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (__ldg(idx + mid) < bound) lo = mid + 1; else hi = mid; }
        splits[g] = (uint32_t)lo;
    }
}

template <int WORDS> __device__ __forceinline__ void smem_add(typename Word<WORDS>::T* p, typename Word<WORDS>::T v);
template <> __device__ __forceinline__ void smem_add<1>(uint32_t* p, uint32_t v) { atomicAdd(p, v); }
template <> __device__ __forceinline__ void smem_add<2>(uint64_t* p, uint64_t v) { atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v); }
template <> __device__ __forceinline__ void smem_add<4>(u128* p, u128 v) { *p = Word<4>::add(*p, v); }   // (one client at a time, see below)

#ifndef SP_MINB
#define SP_MINB 4
#endif
template <int WORDS, bool ONE_GROUP>
__global__ void __launch_bounds__(256, SP_MINB)
k_sparse_sum_tiled(const SparseClient* __restrict__ cl_g, const __grid_constant__ SparseClientsParam P, int n, const uint32_t* __restrict__ splits,
                   uint64_t total, uint64_t n_tiles, typename Word<WORDS>::T zsum, uint32_t b, typename Word<WORDS>::T* dense, int inplace,
                   int subtract) {
    const SparseClient* __restrict__ cl = cl_g ? cl_g : P.c;
    // inplace: the tile starts from what `dense` holds (flashe_sparse_apply_masks_batch) instead of the sum of the zero
    // words; subtract: the entries are taken away (compact = masks, zero words unused)
    typedef Word<WORDS> WT;
    typedef typename WT::T word_t;
    constexpr uint32_t TILE = SPARSE_TILE_BYTES / (4u * WORDS);
    constexpr uint32_t PER16 = 4u / WORDS;                                // words per 16 bytes
    __shared__ __align__(16) word_t acc[TILE];
    const word_t mk = WT::mask(b);
    const uint64_t per = n_tiles + 1;
    // n <= blockDim.x (the usual case): a thread keeps ONE client for the whole launch - tpc threads per client - so the
    // client's pointers are fetched once and the run bounds of the NEXT tile are fetched while this tile's run is walked
    constexpr bool one_group = ONE_GROUP && WORDS <= 2;                // (the host launches it when 1 <= n <= blockDim.x)
    uint32_t g_tpc = 1u, g_r = 0u, g_p0 = 0u, g_p1 = 0u;
    const word_t* g_cp = nullptr; const int64_t* g_ix = nullptr; const uint32_t* g_sp = nullptr;
    word_t g_z = WT::zero();
    if (one_group) {
        g_tpc = blockDim.x / (uint32_t)n;
        const uint32_t cg = threadIdx.x / g_tpc;
        g_r = threadIdx.x - cg * g_tpc;
        if (cg < (uint32_t)n) {
            g_cp = reinterpret_cast<const word_t*>(cl[cg].compact); g_ix = cl[cg].index; g_sp = splits + (uint64_t)cg * per;
            if constexpr (WORDS <= 2) g_z = (word_t)cl[cg].zero_lo & mk;
            if (blockIdx.x < n_tiles) { g_p0 = __ldg(g_sp + blockIdx.x); g_p1 = __ldg(g_sp + blockIdx.x + 1); }
        }
    }
    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const uint64_t a = t * TILE;
        const uint32_t cnt = (uint32_t)(total - a < TILE ? total - a : TILE);
        word_t* out = dense + a;
        const bool vec = (reinterpret_cast<uintptr_t>(out) & 15u) == 0u;  // (block-uniform) whole 16-byte accesses
        const uint32_t nvec = vec ? cnt / PER16 : 0u;
        if (inplace) {
            uint4* av = reinterpret_cast<uint4*>(acc);
            for (uint32_t i = threadIdx.x; i < nvec; i += 4u * blockDim.x) {   // four 16-byte loads in flight per thread
                uint4 w[4];
#pragma unroll
                for (uint32_t q = 0; q < 4u; ++q) if (i + q * blockDim.x < nvec) w[q] = reinterpret_cast<const uint4*>(out)[i + q * blockDim.x];
#pragma unroll
                for (uint32_t q = 0; q < 4u; ++q) if (i + q * blockDim.x < nvec) av[i + q * blockDim.x] = w[q];
            }
            for (uint32_t i = nvec * PER16 + threadIdx.x; i < cnt; i += blockDim.x) acc[i] = out[i];
        } else {
            for (uint32_t i = threadIdx.x; i < TILE; i += blockDim.x) acc[i] = zsum;
        }
        if constexpr (one_group) {
            {
                __syncthreads();                                           // the tile is initialised
                uint32_t np0 = 0u, np1 = 0u;
                const uint64_t tn = t + gridDim.x;
                if (g_sp && tn < n_tiles) { np0 = __ldg(g_sp + tn); np1 = __ldg(g_sp + tn + 1); }   // used after this tile
                const uint32_t a32 = (uint32_t)a;                          // offsets in the tile: low words suffice (TILE <= 2^32)
                for (uint32_t p = g_p0 + g_r; p < g_p1; p += SP_DEPTH * g_tpc) {   // SP_DEPTH entries' loads in flight, then their adds
                    uint32_t off[SP_DEPTH]; word_t v[SP_DEPTH];
#pragma unroll
                    for (uint32_t q = 0; q < SP_DEPTH; ++q) {
                        const uint32_t pq = p + q * g_tpc;
                        if (pq < g_p1) { off[q] = (uint32_t)__ldg(reinterpret_cast<const uint2*>(g_ix + pq)).x - a32; v[q] = __ldg(g_cp + pq); }
                    }
#pragma unroll
                    for (uint32_t q = 0; q < SP_DEPTH; ++q)
                        if (p + q * g_tpc < g_p1) smem_add<WORDS>(&acc[off[q]], subtract ? WT::sub(WT::zero(), v[q]) : WT::sub(v[q], g_z));
                }
                g_p0 = np0; g_p1 = np1;
                __syncthreads();
            }
        } else if constexpr (WORDS <= 2 && !ONE_GROUP) {
            // tpc threads per client walk the client's run of this tile, every client at once and no barrier between
            // fetching the run bounds and walking the run (a warp per client, run after run, or a block-wide scan of the
            // run lengths first, both left the loads of a tile in dependent phases: 84 - 92 us for 32 clients x 1 % of
            // 50 M).  Entries of different clients may meet in a word: shared-memory atomics.
            const uint32_t group = n < (int)blockDim.x ? (uint32_t)n : blockDim.x;
            const uint32_t tpc = blockDim.x / group;
            const uint32_t cg = threadIdx.x / tpc, r = threadIdx.x - cg * tpc;
            bool synced = false;
            for (uint32_t c0 = 0; c0 < (uint32_t)n; c0 += group) {
                const uint32_t c = c0 + cg;
                uint32_t p0 = 0u, p1 = 0u;
                const word_t* cp = nullptr; const int64_t* ix = nullptr; word_t z = WT::zero();
                if (cg < group && c < (uint32_t)n) {
                    p0 = __ldg(splits + (uint64_t)c * per + t); p1 = __ldg(splits + (uint64_t)c * per + t + 1);
                    cp = reinterpret_cast<const word_t*>(cl[c].compact); ix = cl[c].index; z = (word_t)cl[c].zero_lo & mk;
                }
                if (!synced) { __syncthreads(); synced = true; }              // the tile is initialised (the loads above are in flight)
                for (uint32_t p = p0 + r; p < p1; p += 4u * tpc) {            // four entries' loads in flight, then their adds
                    uint32_t off[4]; word_t v[4];
#pragma unroll
                    for (uint32_t q = 0; q < 4u; ++q) {
                        const uint32_t pq = p + q * tpc;
                        if (pq < p1) { off[q] = (uint32_t)((uint64_t)__ldg(ix + pq) - a); v[q] = __ldg(cp + pq); }
                    }
#pragma unroll
                    for (uint32_t q = 0; q < 4u; ++q)
                        if (p + q * tpc < p1) smem_add<WORDS>(&acc[off[q]], subtract ? WT::sub(WT::zero(), v[q]) : WT::sub(v[q], z));
                }
            }
            if (!synced) __syncthreads();
            __syncthreads();
        } else {
            __syncthreads();
            // 16-byte words have no atomic: the clients take turns (a client's indices are unique)
            for (int c = 0; c < n; ++c) {
                const uint32_t p0 = __ldg(splits + (uint64_t)c * per + t), p1 = __ldg(splits + (uint64_t)c * per + t + 1);
                if (p0 < p1) {
                    const word_t* __restrict__ cp = reinterpret_cast<const word_t*>(cl[c].compact);
                    const int64_t* __restrict__ ix = cl[c].index;
                    word_t z;
                    if constexpr (WORDS == 4) { z.lo = cl[c].zero_lo; z.hi = cl[c].zero_hi; z = WT::band(z, mk); }
                    for (uint32_t p = p0 + threadIdx.x; p < p1; p += blockDim.x)
                        smem_add<WORDS>(&acc[(uint32_t)((uint64_t)__ldg(ix + p) - a)], subtract ? WT::sub(WT::zero(), cp[p]) : WT::sub(cp[p], z));
                }
                __syncthreads();
            }
        }
        {
            const uint4* av = reinterpret_cast<const uint4*>(acc);
#pragma unroll 4
            for (uint32_t i = threadIdx.x; i < nvec; i += blockDim.x) {
                uint4 v = av[i];
                if constexpr (WORDS == 1) { const uint32_t m = mk; v.x &= m; v.y &= m; v.z &= m; v.w &= m; }
                else if constexpr (WORDS == 2) { const uint64_t m = mk; v.x &= (uint32_t)m; v.y &= (uint32_t)(m >> 32); v.z &= (uint32_t)m; v.w &= (uint32_t)(m >> 32); }
                else { v.x &= (uint32_t)mk.lo; v.y &= (uint32_t)(mk.lo >> 32); v.z &= (uint32_t)mk.hi; v.w &= (uint32_t)(mk.hi >> 32); }
                reinterpret_cast<uint4*>(out)[i] = v;
            }
            for (uint32_t i = nvec * PER16 + threadIdx.x; i < cnt; i += blockDim.x) out[i] = WT::band(acc[i], mk);
        }
        __syncthreads();
    }
}

// overlap[i] += |run_i ∩ run_{i+1}| tile by tile (n <= 256 clients): the runs of ALL clients in the tile are brought into
// shared memory as 32-bit offsets (one coalesced pass over the index lists in total), every entry of clients 0 .. n-2
// then looks for itself in the next client's run there (~7 probes of shared memory), hits are counted per pair in shared
// memory and flushed once per tile.  A tile whose runs do not fit (OV_CAP entries) searches in global memory instead.
// (The per-pair launches this replaces - every entry a binary search over the neighbour's whole list in global memory -
// took 0.57 ms for 32 clients x 1 % of 50 M; confined to the neighbour's run of the tile but still in global memory: 0.16.)
#define OV_CAP 6144u
__global__ void __launch_bounds__(256)
k_sparse_overlap_tiled(const SparseClient* __restrict__ cl_g, const __grid_constant__ SparseClientsParam P, int n,
                       const uint32_t* __restrict__ splits, uint32_t tile_log2, uint64_t n_tiles, unsigned long long* __restrict__ out) {
    const SparseClient* __restrict__ cl = cl_g ? cl_g : P.c;
    __shared__ uint32_t s_off[OV_CAP];
    const uint64_t per = n_tiles + 1;
    // tpc threads share a client for the whole launch: thread (cg, r) brings entries r, r + tpc, .. of client cg's run
    // into the client's SLOT of s_off (OV_CAP / n offsets: no scan of the run lengths, no layout to agree on) and looks
    // them up in the next client's slot; its hits stay in a register until the end.  The run bounds of the next tile are
    // fetched while this tile is worked on.  A tile in which some run exceeds its slot searches in global memory.
    const uint32_t tpc = blockDim.x / (uint32_t)n, cg = threadIdx.x / tpc, r = threadIdx.x - cg * tpc;
    const uint32_t slot = OV_CAP / (uint32_t)n;
    const bool mine = cg < (uint32_t)n, has_next = mine && cg + 1u < (uint32_t)n;
    const int64_t* __restrict__ ix = mine ? cl[cg].index : nullptr;
    const int64_t* __restrict__ ix_next = has_next ? cl[cg + 1u].index : nullptr;
    const uint32_t* __restrict__ sp = splits + (uint64_t)(mine ? cg : 0u) * per;
    unsigned long long hits = 0ull;
    uint32_t p0 = 0u, p1 = 0u, q1 = 0u;                                    // own run [p0, p1), next client's run [p1', q1): list positions
    uint32_t q0 = 0u;
    if (mine && blockIdx.x < n_tiles) {
        p0 = __ldg(sp + blockIdx.x); p1 = __ldg(sp + blockIdx.x + 1);
        if (has_next) { q0 = __ldg(sp + per + blockIdx.x); q1 = __ldg(sp + per + blockIdx.x + 1); }
    }
    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const uint32_t a32 = (uint32_t)(t << tile_log2);                   // offsets inside the tile: low words suffice
        const uint64_t tn = t + gridDim.x;
        uint32_t np0 = 0u, np1 = 0u, nq0 = 0u, nq1 = 0u;
        if (mine && tn < n_tiles) {
            np0 = __ldg(sp + tn); np1 = __ldg(sp + tn + 1);
            if (has_next) { nq0 = __ldg(sp + per + tn); nq1 = __ldg(sp + per + tn + 1); }
        }
        const uint32_t len = p1 - p0, len_next = q1 - q0;
        if (!__syncthreads_or(len > slot)) {                               // (also: the previous tile's lookups are done)
            if (mine) {
                const uint32_t base = cg * slot;
                for (uint32_t k = r; k < len; k += 6u * tpc) {             // six loads in flight
                    uint32_t o[6];
#pragma unroll
                    for (uint32_t q = 0; q < 6u; ++q) if (k + q * tpc < len) o[q] = (uint32_t)__ldg(reinterpret_cast<const uint2*>(ix + p0 + k + q * tpc)).x - a32;
#pragma unroll
                    for (uint32_t q = 0; q < 6u; ++q) if (k + q * tpc < len) s_off[base + k + q * tpc] = o[q];
                }
            }
            __syncthreads();
            if (has_next) {
                const uint32_t base = cg * slot, nb = base + slot;
                for (uint32_t k = r; k < len; k += tpc) {
                    const uint32_t v = s_off[base + k];
                    uint32_t lo = 0u, hi = len_next;
                    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (s_off[nb + mid] < v) lo = mid + 1; else hi = mid; }
                    hits += (lo < len_next && s_off[nb + lo] == v) ? 1ull : 0ull;
                }
            }
        } else if (has_next) {                                             // some run does not fit its slot: search in global memory
            for (uint32_t p = p0 + r; p < p1; p += tpc) {
                const int64_t v = __ldg(ix + p);
                uint32_t lo = q0, hi = q1;
                while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(ix_next + mid) < v) lo = mid + 1; else hi = mid; }
                hits += (lo < q1 && __ldg(ix_next + lo) == v) ? 1ull : 0ull;
            }
        }
        p0 = np0; p1 = np1; q0 = nq0; q1 = nq1;
    }
    if (has_next && hits) atomicAdd(out + cg, hits);
}

// The same counts through per-client BITMAPS of the tile (n <= 64 clients: n x tile / 8 <= 32 KB of shared memory): the
// entries of every client set their bits, then every entry of client i tests its bit in client i+1's bitmap - a dozen
// instructions per entry where the binary search in shared memory spent ~90 (ncu: that kernel was issue-bound, 66 M warp
// instructions, 98 us for 32 clients x 1 % of 50 M).  The entries are read a second time for the test (L1 / L2 hits), and
// a third pass clears exactly the words that were set.
#define OVB_BYTES 32768u
__global__ void __launch_bounds__(256)
k_sparse_overlap_bitmap(const SparseClient* __restrict__ cl_g, const __grid_constant__ SparseClientsParam P, int n,
                        const uint32_t* __restrict__ splits, uint32_t tile_log2, uint64_t n_tiles, unsigned long long* __restrict__ out) {
    const SparseClient* __restrict__ cl = cl_g ? cl_g : P.c;
    __shared__ uint32_t bm[OVB_BYTES / 4u];
    const uint64_t per = n_tiles + 1;
    const uint32_t wpc = 1u << (tile_log2 - 5u);                           // bitmap words per client
    const uint32_t tpc = blockDim.x / (uint32_t)n, cg = threadIdx.x / tpc, r = threadIdx.x - cg * tpc;
    const bool mine = cg < (uint32_t)n, has_next = mine && cg + 1u < (uint32_t)n;
    const int64_t* __restrict__ ix = mine ? cl[cg].index : nullptr;
    const uint32_t* __restrict__ sp = splits + (uint64_t)(mine ? cg : 0u) * per;
    uint32_t* own = bm + cg * wpc;
    const uint32_t* nxt = own + wpc;
    for (uint32_t i = threadIdx.x; i < (uint32_t)n * wpc; i += blockDim.x) bm[i] = 0u;
    unsigned long long hits = 0ull;
    uint32_t p0 = 0u, p1 = 0u;
    if (mine && blockIdx.x < n_tiles) { p0 = __ldg(sp + blockIdx.x); p1 = __ldg(sp + blockIdx.x + 1); }
    __syncthreads();
    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const uint32_t a32 = (uint32_t)(t << tile_log2);
        const uint64_t tn = t + gridDim.x;
        uint32_t np0 = 0u, np1 = 0u;
        if (mine && tn < n_tiles) { np0 = __ldg(sp + tn); np1 = __ldg(sp + tn + 1); }
        for (uint32_t p = p0 + r; p < p1; p += 6u * tpc) {                 // set: six loads in flight
            uint32_t o[6];
#pragma unroll
            for (uint32_t q = 0; q < 6u; ++q) if (p + q * tpc < p1) o[q] = (uint32_t)__ldg(reinterpret_cast<const uint2*>(ix + p + q * tpc)).x - a32;
#pragma unroll
            for (uint32_t q = 0; q < 6u; ++q) if (p + q * tpc < p1) atomicOr(own + (o[q] >> 5), 1u << (o[q] & 31u));
        }
        __syncthreads();
        if (has_next) {
            for (uint32_t p = p0 + r; p < p1; p += 6u * tpc) {             // test in the next client's bitmap
                uint32_t o[6];
#pragma unroll
                for (uint32_t q = 0; q < 6u; ++q) if (p + q * tpc < p1) o[q] = (uint32_t)__ldg(reinterpret_cast<const uint2*>(ix + p + q * tpc)).x - a32;
#pragma unroll
                for (uint32_t q = 0; q < 6u; ++q) if (p + q * tpc < p1) hits += (nxt[o[q] >> 5] >> (o[q] & 31u)) & 1u;
            }
        }
        __syncthreads();
        for (uint32_t p = p0 + r; p < p1; p += 6u * tpc) {                 // clear the words this thread set
            uint32_t o[6];
#pragma unroll
            for (uint32_t q = 0; q < 6u; ++q) if (p + q * tpc < p1) o[q] = (uint32_t)__ldg(reinterpret_cast<const uint2*>(ix + p + q * tpc)).x - a32;
#pragma unroll
            for (uint32_t q = 0; q < 6u; ++q) if (p + q * tpc < p1) own[o[q] >> 5] = 0u;
        }
        __syncthreads();
        p0 = np0; p1 = np1;
    }
    if (has_next && hits) atomicAdd(out + cg, hits);
}

// Uploads the client table and computes the runs; *ws_out (one stream-ordered allocation) holds both.
static int sparse_prepare(flashe_ctx* ctx, const void* const* compacts, const int64_t* const* indexes, const uint64_t* ks, const void* zero_words,
                          int word_bytes, int n, uint64_t total, uint32_t tile_log2, cudaStream_t cs, uint8_t** ws_out, SparseClient** cl_out,
                          SparseClientsParam* P, uint32_t** splits_out, uint64_t* n_tiles_out) {
    const uint64_t n_tiles = (total + (1ull << tile_log2) - 1) >> tile_log2;
    const bool in_params = n <= SPARSE_PARAM_CLIENTS;
    std::vector<SparseClient> h((size_t)(in_params ? 0 : n));
    memset(P, 0, sizeof(*P));
    for (int c = 0; c < n; ++c) {
        SparseClient e; memset(&e, 0, sizeof(e));
        e.compact = compacts ? compacts[c] : nullptr; e.index = indexes[c]; e.k = ks[c];
        if (zero_words) memcpy(&e.zero_lo, (const uint8_t*)zero_words + (size_t)c * word_bytes, (size_t)word_bytes);
        if (in_params) P->c[c] = e; else h[c] = e;
    }
    const size_t cl_bytes = in_params ? 0 : ((sizeof(SparseClient) * (size_t)n + 255) & ~(size_t)255);
    const size_t sp_bytes = sizeof(uint32_t) * (size_t)n * (size_t)(n_tiles + 1);
    uint8_t* ws = nullptr;
    CUDA_TRY(cudaMallocAsync((void**)&ws, cl_bytes + sp_bytes, cs));
    SparseClient* cl = nullptr;
    if (!in_params) {
        cudaError_t e = cudaMemcpyAsync(ws, h.data(), sizeof(SparseClient) * (size_t)n, cudaMemcpyHostToDevice, cs);   // (pageable source: staged before the call returns)
        if (e != cudaSuccess) { cudaFreeAsync(ws, cs); return fail(FLASHE_ECUDA, std::string("sparse client table: ") + cudaGetErrorString(e)); }
        cl = reinterpret_cast<SparseClient*>(ws);
    }
    uint32_t* splits = reinterpret_cast<uint32_t*>(ws + cl_bytes);
    k_sparse_splits<<<GRID_OCC(ctx, k_sparse_splits, (uint64_t)n * (n_tiles + 1), 256), 256, 0, cs>>>(cl, *P, n, total, tile_log2, n_tiles, splits);
    count_launch();
    *ws_out = ws; *cl_out = cl; *splits_out = splits; *n_tiles_out = n_tiles;
    return FLASHE_OK;
}

template <int WORDS>
static int sparse_sum_tiled_t(flashe_ctx* ctx, const void* const* compacts, const int64_t* const* indexes, const uint64_t* ks,
                              const void* zero_words, int n, uint64_t total, void* dense_out, cudaStream_t cs) {
    typedef Word<WORDS> WT;
    typedef typename WT::T word_t;
    constexpr uint32_t TILE = SPARSE_TILE_BYTES / (4u * WORDS);
    uint32_t tile_log2 = 0; while ((1u << tile_log2) < TILE) ++tile_log2;
    std::vector<word_t> zeros((size_t)n);
    memcpy(zeros.data(), zero_words, sizeof(word_t) * (size_t)n);
    const word_t mk = WT::mask((uint32_t)ctx->int_bits);
    word_t zsum = WT::zero();
    for (int c = 0; c < n; ++c) zsum = WT::band(WT::add(zsum, WT::band(zeros[c], mk)), mk);
    uint8_t* ws; SparseClient* cl; uint32_t* splits; uint64_t n_tiles; SparseClientsParam P;
    int rc = sparse_prepare(ctx, compacts, indexes, ks, zero_words, 4 * WORDS, n, total, tile_log2, cs, &ws, &cl, &P, &splits, &n_tiles);
    if (rc) return rc;
    const int grid = (int)(n_tiles < (uint64_t)ctx->num_sms * 32 ? n_tiles : (uint64_t)ctx->num_sms * 32);
    if (WORDS <= 2 && n <= 256) k_sparse_sum_tiled<WORDS, (WORDS <= 2)><<<grid, 256, 0, cs>>>(cl, P, n, splits, total, n_tiles, zsum, (uint32_t)ctx->int_bits, (word_t*)dense_out, 0, 0);
    else k_sparse_sum_tiled<WORDS, false><<<grid, 256, 0, cs>>>(cl, P, n, splits, total, n_tiles, zsum, (uint32_t)ctx->int_bits, (word_t*)dense_out, 0, 0);
    count_launch();
    cudaFreeAsync(ws, cs);
    CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}

// dense[index_c[p]] += / -= compact_c[p] (mod 2^b) for every client, in place, by tiles (flashe_internal.h)
template <int WORDS>
static int sparse_accumulate_t(flashe_ctx* ctx, const void* const* compacts, const int64_t* const* indexes, const uint64_t* ks, int n,
                               int subtract, uint64_t total, void* dense, cudaStream_t cs) {
    typedef typename Word<WORDS>::T word_t;
    constexpr uint32_t TILE = SPARSE_TILE_BYTES / (4u * WORDS);
    uint32_t tile_log2 = 0; while ((1u << tile_log2) < TILE) ++tile_log2;
    uint8_t* ws; SparseClient* cl; uint32_t* splits; uint64_t n_tiles; SparseClientsParam P;
    int rc = sparse_prepare(ctx, compacts, indexes, ks, nullptr, 4 * WORDS, n, total, tile_log2, cs, &ws, &cl, &P, &splits, &n_tiles);
    if (rc) return rc;
    const int grid = (int)(n_tiles < (uint64_t)ctx->num_sms * 32 ? n_tiles : (uint64_t)ctx->num_sms * 32);
    if (WORDS <= 2 && n <= 256) k_sparse_sum_tiled<WORDS, (WORDS <= 2)><<<grid, 256, 0, cs>>>(cl, P, n, splits, total, n_tiles, Word<WORDS>::zero(), (uint32_t)ctx->int_bits, (word_t*)dense, 1, subtract);
    else k_sparse_sum_tiled<WORDS, false><<<grid, 256, 0, cs>>>(cl, P, n, splits, total, n_tiles, Word<WORDS>::zero(), (uint32_t)ctx->int_bits, (word_t*)dense, 1, subtract);
    count_launch();
    cudaFreeAsync(ws, cs);
    CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}
int flashe_sparse_accumulate_tiled(flashe_ctx* ctx, const void* const* compacts, const int64_t* const* indexes, const uint64_t* ks, int n,
                                   int subtract, uint64_t total, void* dense, cudaStream_t cs) {
    if (ctx->words == 1) return sparse_accumulate_t<1>(ctx, compacts, indexes, ks, n, subtract, total, dense, cs);
    if (ctx->words == 2) return sparse_accumulate_t<2>(ctx, compacts, indexes, ks, n, subtract, total, dense, cs);
    return sparse_accumulate_t<4>(ctx, compacts, indexes, ks, n, subtract, total, dense, cs);
}

template <int WORDS>
static int sparse_sum_t(flashe_ctx* ctx, const void* const* compacts, const int64_t* const* indexes, const uint64_t* ks,
                        const void* zero_words, int n, uint64_t total, void* dense_out, cudaStream_t cs) {
    typedef Word<WORDS> WT;
    typedef typename WT::T word_t;
    // the tiled form needs 32-bit list positions and cannot be recorded into a CUDA graph (it uploads a table);
    // FLASHE_SPARSE_TILED=0 keeps the per-client scatter-adds (A/B measurements)
    {
        static const bool tiled = [] { const char* e = getenv("FLASHE_SPARSE_TILED"); return !(e && e[0] == '0'); }();
        bool ok = tiled;
        for (int c = 0; c < n; ++c) ok = ok && ks[c] < (1ull << 32);
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(cs, &cap) != cudaSuccess) { cudaGetLastError(); ok = false; }
        if (ok && cap == cudaStreamCaptureStatusNone) return sparse_sum_tiled_t<WORDS>(ctx, compacts, indexes, ks, zero_words, n, total, dense_out, cs);
    }
    std::vector<word_t> zeros((size_t)n);                              // the caller's buffer need not be aligned
    memcpy(zeros.data(), zero_words, sizeof(word_t) * (size_t)n);
    const word_t mk = WT::mask((uint32_t)ctx->int_bits);
    word_t zsum = WT::zero();
    for (int c = 0; c < n; ++c) zsum = WT::band(WT::add(zsum, WT::band(zeros[c], mk)), mk);
    k_fill<WORDS><<<GRID_OCC(ctx, k_fill<WORDS>, total, 256), 256, 0, cs>>>((word_t*)dense_out, total, zsum);
    int launches = 1;
    for (int c = 0; c < n; ++c) {
        if (!ks[c]) continue;
        k_scatter_add<WORDS><<<GRID_OCC(ctx, k_scatter_add<WORDS>, ks[c], 256), 256, 0, cs>>>((const word_t*)compacts[c], indexes[c], ks[c], total,
                                                                           WT::band(zeros[c], mk), (uint32_t)ctx->int_bits, (word_t*)dense_out);
        ++launches;
    }
    count_launch(launches);
    CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}


extern "C" {

int flashe_add_premasked(flashe_ctx* ctx, const void* in, const void* mask, int sign, uint64_t count, void* out, void* stream) {
    ENTER(ctx);
    if (count == 0) return FLASHE_OK;
    if (!in || !mask || !out) return fail(FLASHE_EINVAL, "NULL buffer");
    const uint32_t b = (uint32_t)ctx->int_bits;
    if (ctx->words == 1) {
        const bool aligned = (((uintptr_t)in | (uintptr_t)mask | (uintptr_t)out) & 15u) == 0;
        uint64_t nvec = aligned ? count / 4 : 0;
        const bool aligned32 = (((uintptr_t)in | (uintptr_t)mask | (uintptr_t)out) & 31u) == 0;
        uint64_t done = 0;
        if (nvec && aligned32 && count >= 8) {                          // 32-byte accesses, the odd quad goes to the generic kernel
            const uint64_t n8 = count / 8;
            k_add_premasked_v8<<<GRID_OCC(ctx, k_add_premasked_v8, n8, 256), 256, 0, cs>>>((const uint32_t*)in, (const uint32_t*)mask, sign, n8, Word<1>::mask(b), (uint32_t*)out);
            count_launch();
            done = n8 * 8;
        } else if (nvec) {
            k_add_premasked_v4<<<GRID_OCC(ctx, k_add_premasked_v4, nvec, 256), 256, 0, cs>>>((const uint4*)in, (const uint4*)mask, sign, nvec, Word<1>::mask(b) , (uint4*)out);
            count_launch();
            done = nvec * 4;
        }
        if (done < count) {
            k_add_premasked<1><<<GRID_OCC(ctx, k_add_premasked<1>, count - done, 256), 256, 0, cs>>>((const uint32_t*)in + done, (const uint32_t*)mask + done, sign, count - done, b, (uint32_t*)out + done);
            count_launch();
        }
    } else if (ctx->words == 2) {
        k_add_premasked<2><<<GRID_OCC(ctx, k_add_premasked<2>, count, 256), 256, 0, cs>>>((const uint64_t*)in, (const uint64_t*)mask, sign, count, b, (uint64_t*)out);
        count_launch();
    } else {
        k_add_premasked<4><<<GRID_OCC(ctx, k_add_premasked<4>, count, 256), 256, 0, cs>>>((const u128*)in, (const u128*)mask, sign, count, b, (u128*)out);
        count_launch();
    }
    CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}

int flashe_encode(flashe_ctx* ctx, const flashe_span* span, const float* x, const flashe_codec* codec, const flashe_noise* noise,
                  uint32_t* q_out, void* stream) {
    ENTER(ctx);
    int rc = flashe_check_span(span); if (rc) return rc;
    if (span->count == 0) return FLASHE_OK;
    if (!x || !q_out) return fail(FLASHE_EINVAL, "NULL buffer");
    if (codec && codec->batch_lane_bits) return fail(FLASHE_EINVAL, "lane batching is fused into flashe_encode_encrypt / flashe_decrypt_decode (or use flashe_batch_pack_layers)");
    CodecHost ch; rc = make_codec(ctx, span, codec, false, cs, &ch); if (rc) return rc;
    rc = check_noise(noise); if (rc) { free_codec(&ch, cs); return rc; }
    NoiseDev nz; make_noise(noise, 0, &nz);
    if (nz.res32) k_encode<1, false, true><<<GRID_OCC(ctx, (k_encode<1, false, true>), span->count, 256), 256, 0, cs>>>(x, nullptr, span->begin, span->count, 32, ch.dev, nz, q_out, nullptr);
    else k_encode<1, false><<<GRID_OCC(ctx, (k_encode<1, false>), span->count, 256), 256, 0, cs>>>(x, nullptr, span->begin, span->count, 32, ch.dev, nz, q_out, nullptr);
    count_launch();
    free_codec(&ch, cs);
    CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}

int flashe_encode_add_premasked(flashe_ctx* ctx, const flashe_span* span, const float* x, const flashe_codec* codec,
                                const flashe_noise* noise, const void* mask, void* ct_out, void* stream) {
    ENTER(ctx);
    int rc = flashe_check_span(span); if (rc) return rc;
    if (span->count == 0) return FLASHE_OK;
    if (!x || !mask || !ct_out) return fail(FLASHE_EINVAL, "NULL buffer");
    if (codec && codec->batch_lane_bits) return fail(FLASHE_EINVAL, "lane batching is fused into flashe_encode_encrypt / flashe_decrypt_decode (or use flashe_batch_pack_layers)");
    CodecHost ch; rc = make_codec(ctx, span, codec, false, cs, &ch); if (rc) return rc;
    rc = check_noise(noise); if (rc) { free_codec(&ch, cs); return rc; }
    NoiseDev nz; make_noise(noise, 0, &nz);
    const uint32_t b = (uint32_t)ctx->int_bits;
    const int grid = ctx->words == 1 ? GRID_OCC(ctx, (k_encode<1, true>), span->count, 256)
                   : (ctx->words == 2 ? GRID_OCC(ctx, (k_encode<2, true>), span->count, 256) : GRID_OCC(ctx, (k_encode<4, true>), span->count, 256));
    const bool v4 = ctx->words == 1 && (span->begin & 3ull) == 0 &&
                    ((((uintptr_t)x | (uintptr_t)mask | (uintptr_t)ct_out | (uintptr_t)nz.u) & 15u) == 0);
    if (v4 && !nz.u && !nz.res32 && (span->begin & 7ull) == 0 && (span->count & 7ull) == 0 && ((((uintptr_t)x | (uintptr_t)mask | (uintptr_t)ct_out) & 31u) == 0)) {
        const uint64_t n8 = span->count / 8;                            // 32-byte accesses
        auto k8 = nz.res32 ? k_encode_premasked_v8<true> : k_encode_premasked_v8<false>;
        k8<<<GRID_OCC(ctx, k8, n8, 256), 256, 0, cs>>>((const uint32_t*)x, (const uint32_t*)mask, span->begin, n8, Word<1>::mask(b), ch.dev, nz, (uint32_t*)ct_out, 0, 0, 0);
    } else
    if (v4) {
        const uint64_t nvec = span->count / 4, done = nvec * 4;
        auto kpm = nz.res32 ? k_encode_premasked_v4<true> : k_encode_premasked_v4<false>;
        if (nvec) kpm<<<GRID_OCC(ctx, kpm, nvec, 256), 256, 0, cs>>>((const uint4*)x, (const uint4*)mask, span->begin, nvec, Word<1>::mask(b), ch.dev, nz, (uint4*)ct_out, 0, 0, 0, 0);
        if (done < span->count) {
            NoiseDev nt = nz; if (nt.u) nt.u += done;
            if (nz.res32) k_encode<1, true, true><<<1, 32, 0, cs>>>(x + done, (const uint32_t*)mask + done, span->begin + done, span->count - done, b, ch.dev, nt, nullptr, (uint32_t*)ct_out + done);
            else k_encode<1, true><<<1, 32, 0, cs>>>(x + done, (const uint32_t*)mask + done, span->begin + done, span->count - done, b, ch.dev, nt, nullptr, (uint32_t*)ct_out + done);
            count_launch(nvec ? 1 : 0);
        }
    }
    else if (nz.res32) {
        if (ctx->words == 1) k_encode<1, true, true><<<grid, 256, 0, cs>>>(x, (const uint32_t*)mask, span->begin, span->count, b, ch.dev, nz, nullptr, (uint32_t*)ct_out);
        else if (ctx->words == 2) k_encode<2, true, true><<<grid, 256, 0, cs>>>(x, (const uint64_t*)mask, span->begin, span->count, b, ch.dev, nz, nullptr, (uint64_t*)ct_out);
        else k_encode<4, true, true><<<grid, 256, 0, cs>>>(x, (const u128*)mask, span->begin, span->count, b, ch.dev, nz, nullptr, (u128*)ct_out);
    }
    else if (ctx->words == 1) k_encode<1, true><<<grid, 256, 0, cs>>>(x, (const uint32_t*)mask, span->begin, span->count, b, ch.dev, nz, nullptr, (uint32_t*)ct_out);
    else if (ctx->words == 2) k_encode<2, true><<<grid, 256, 0, cs>>>(x, (const uint64_t*)mask, span->begin, span->count, b, ch.dev, nz, nullptr, (uint64_t*)ct_out);
    else k_encode<4, true><<<grid, 256, 0, cs>>>(x, (const u128*)mask, span->begin, span->count, b, ch.dev, nz, nullptr, (u128*)ct_out);
    count_launch();
    free_codec(&ch, cs);
    CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}

int flashe_encode_add_premasked_batch(flashe_ctx* ctx, const flashe_span* span, int n_clients, const float* x, uint64_t x_stride,
                                      const flashe_codec* codec, const flashe_noise* noise, uint64_t u_stride, const void* mask,
                                      uint64_t mask_stride, void* ct_out, uint64_t ct_stride, void* stream) {
    ENTER(ctx);
    int rc = flashe_check_span(span); if (rc) return rc;
    if (n_clients < 1) return fail(FLASHE_EINVAL, "n_clients must be >= 1");
    if (span->count == 0) return FLASHE_OK;
    if (!x || !mask || !ct_out) return fail(FLASHE_EINVAL, "NULL buffer");
    if (n_clients > 1 && (x_stride < span->count || mask_stride < span->count || ct_stride < span->count))
        return fail(FLASHE_EINVAL, "client strides must be >= span.count");
    const bool v4 = ctx->words == 1 && (span->begin & 3ull) == 0 && (span->count & 3ull) == 0 &&
                    ((((uintptr_t)x | (uintptr_t)mask | (uintptr_t)ct_out | (uintptr_t)(noise ? noise->u : nullptr)) & 15u) == 0) &&
                    (n_clients == 1 || ((x_stride | mask_stride | ct_stride | u_stride) & 3ull) == 0);
    if (!v4) {                                                         // other layouts: one launch per client
        for (int c = 0; c < n_clients; ++c) {
            flashe_noise nc; if (noise) { nc = *noise; if (nc.u) nc.u += (uint64_t)c * u_stride; nc.rng_stream += (uint64_t)c; }
            const size_t wb = 4u * (size_t)ctx->words;
            rc = flashe_encode_add_premasked(ctx, span, x + (uint64_t)c * x_stride, codec, noise ? &nc : nullptr,
                                             (const uint8_t*)mask + (uint64_t)c * mask_stride * wb, (uint8_t*)ct_out + (uint64_t)c * ct_stride * wb, stream);
            if (rc) return rc;
        }
        return FLASHE_OK;
    }
    if (codec && codec->batch_lane_bits) return fail(FLASHE_EINVAL, "lane batching is fused into flashe_encode_encrypt / flashe_decrypt_decode");
    CodecHost ch; rc = make_codec(ctx, span, codec, false, cs, &ch); if (rc) return rc;
    rc = check_noise(noise); if (rc) { free_codec(&ch, cs); return rc; }
    NoiseDev nz; make_noise(noise, u_stride, &nz);
    const uint64_t nvec = span->count / 4;
    if (n_clients > 65535) { free_codec(&ch, cs); return fail(FLASHE_EINVAL, "at most 65535 clients per launch"); }
    // 32-byte accesses when every row allows them (device noise)
    // (measured: 53-bit noise 70 -> 83 % of HBM; the 32-bit-resolution form is faster with 16-byte accesses, 89 vs 86 %)
    const bool v8 = !nz.u && !nz.res32 && (span->begin & 7ull) == 0 && (span->count & 7ull) == 0 &&
                    ((((uintptr_t)x | (uintptr_t)mask | (uintptr_t)ct_out) & 31u) == 0) &&
                    (n_clients == 1 || ((x_stride | mask_stride | ct_stride) & 7ull) == 0);
    if (v8) {
        const uint64_t n8 = span->count / 8;
        auto k8 = nz.res32 ? k_encode_premasked_v8<true> : k_encode_premasked_v8<false>;
        int g8 = GRID_OCC(ctx, k8, n8 * (uint64_t)n_clients, 256) / n_clients;
        if (g8 < 1) g8 = 1;
        k8<<<dim3((unsigned)g8, (unsigned)n_clients), 256, 0, cs>>>(
            (const uint32_t*)x, (const uint32_t*)mask, span->begin, n8, Word<1>::mask((uint32_t)ctx->int_bits), ch.dev, nz, (uint32_t*)ct_out,
            x_stride, mask_stride, ct_stride);
        count_launch();
        free_codec(&ch, cs);
        CUDA_TRY(cudaGetLastError());
        return FLASHE_OK;
    }
    auto kpm = nz.res32 ? k_encode_premasked_v4<true> : k_encode_premasked_v4<false>;
    int gx = GRID_OCC(ctx, kpm, nvec * (uint64_t)n_clients, 256) / n_clients;     // resident CTAs, split over the rows
    if (gx < 1) gx = 1;
    kpm<<<dim3((unsigned)gx, (unsigned)n_clients), 256, 0, cs>>>(
        (const uint4*)x, (const uint4*)mask, span->begin, nvec, Word<1>::mask((uint32_t)ctx->int_bits), ch.dev, nz, (uint4*)ct_out,
        x_stride / 4, mask_stride / 4, ct_stride / 4, u_stride);
    count_launch();
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
            const int grid = ctx->words == 1 ? GRID_OCC(ctx, k_aggregate_vec<1>, nvec, 256)
                           : (ctx->words == 2 ? GRID_OCC(ctx, k_aggregate_vec<2>, nvec, 256) : GRID_OCC(ctx, k_aggregate_vec<4>, nvec, 256));
            const uint64_t sv = stride / per_vec;
            if (ctx->words == 1) k_aggregate_vec<1><<<grid, 256, 0, cs>>>((const uint4*)cts, sv, n, nvec, b, (uint4*)out);
            else if (ctx->words == 2) k_aggregate_vec<2><<<grid, 256, 0, cs>>>((const uint4*)cts, sv, n, nvec, b, (uint4*)out);
            else k_aggregate_vec<4><<<grid, 256, 0, cs>>>((const uint4*)cts, sv, n, nvec, b, (uint4*)out);
            count_launch();
        }
        const uint64_t done = nvec * per_vec;
        if (done < count) {
            const int grid = grid_1d(ctx, count - done, 256, 8);
            if (ctx->words == 1) k_aggregate_scalar<1><<<grid, 256, 0, cs>>>((const uint32_t*)cts, stride, n, done, count, b, (uint32_t*)out);
            else if (ctx->words == 2) k_aggregate_scalar<2><<<grid, 256, 0, cs>>>((const uint64_t*)cts, stride, n, done, count, b, (uint64_t*)out);
            else k_aggregate_scalar<4><<<grid, 256, 0, cs>>>((const u128*)cts, stride, n, done, count, b, (u128*)out);
            count_launch();
        }
    } else {
        // The carry transfer of one digit is modelled as cin -> A + (cin >= T): at most +1, which needs
        // cin <= n - 1 < 2^b (with more clients than digit values a digit could hand on +2).
        if (b < 31u && (uint64_t)n > (1ull << b))
            return fail(FLASHE_EINVAL, "packed-carry aggregate needs n <= 2^int_bits");
        const uint64_t per_tile = (uint64_t)PK_THREADS * (ctx->words == 4 ? PkElems<4>::V : PkElems<1>::V);
        const uint64_t ntiles = ceil_div(count, per_tile);
        // persistent grid = exactly the CTAs that are resident at once (a partial second wave would run at a
        // fraction of the occupancy)
        int occ = 0;
        if (ctx->words == 1) CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_aggregate_packed<1>, PK_THREADS, 0));
        else if (ctx->words == 2) CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_aggregate_packed<2>, PK_THREADS, 0));
        else CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_aggregate_packed<4>, PK_THREADS, 0));
        uint64_t cap = (uint64_t)ctx->num_sms * (uint64_t)(occ > 0 ? occ : 1);
        const int grid = (int)(ntiles < cap ? ntiles : cap);
        // 1: rows are 128-bit columns; 2: the output too
        const int vec_ok = (ctx->words == 1 && (count & 3u) == 0 && (stride & 3u) == 0 && aligned16(cts)) ? (aligned16(out) ? 2 : 1) : 0;
        if (ctx->words == 4 && !(aligned16(cts) && aligned16(out))) return fail(FLASHE_EINVAL, "16-byte words must be 16-byte aligned");
        if (ctx->words == 1) k_aggregate_packed<1><<<grid, PK_THREADS, 0, cs>>>((const uint32_t*)cts, stride, n, count, b, carry_in, (uint32_t*)out, carry_out, vec_ok);
        else if (ctx->words == 2) k_aggregate_packed<2><<<grid, PK_THREADS, 0, cs>>>((const uint64_t*)cts, stride, n, count, b, carry_in, (uint64_t*)out, carry_out, 0);
        else k_aggregate_packed<4><<<grid, PK_THREADS, 0, cs>>>((const u128*)cts, stride, n, count, b, carry_in, (u128*)out, carry_out, 0);
        count_launch();
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
    count_launch();
    CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}

int flashe_decode(flashe_ctx* ctx, const flashe_span* span, const void* v, const flashe_codec* codec, double* out, void* stream) {
    ENTER(ctx);
    int rc = flashe_check_span(span); if (rc) return rc;
    if (ctx->words == 4) return fail(FLASHE_EUNSUPPORTED, "decode takes int_bits <= 64 (unbatch 128-bit words first)");
    if (span->count == 0) return FLASHE_OK;
    if (!v || !out) return fail(FLASHE_EINVAL, "NULL buffer");
    if (codec && codec->batch_lane_bits) return fail(FLASHE_EINVAL, "lane batching is fused into flashe_encode_encrypt / flashe_decrypt_decode (or use flashe_batch_pack_layers)");
    CodecHost ch; rc = make_codec(ctx, span, codec, true, cs, &ch); if (rc) return rc;
    const int grid = ctx->words == 1 ? GRID_OCC(ctx, k_decode<1>, span->count, 256) : GRID_OCC(ctx, k_decode<2>, span->count, 256);
    if (ctx->words == 1 && (((uintptr_t)v | (uintptr_t)out) & 15u) == 0) {
        const uint64_t nvec = span->count / 4, done = nvec * 4;
        if (nvec) k_decode_v4<<<GRID_OCC(ctx, k_decode_v4, nvec, 256), 256, 0, cs>>>((const uint4*)v, span->begin, nvec, ch.dev, out);
        if (done < span->count) {
            k_decode<1><<<1, 32, 0, cs>>>((const uint32_t*)v + done, span->begin + done, span->count - done, ch.dev, out + done);
            count_launch();
        }
        if (!nvec) count_launch(-1);
    }
    else if (ctx->words == 1) k_decode<1><<<grid, 256, 0, cs>>>((const uint32_t*)v, span->begin, span->count, ch.dev, out);
    else k_decode<2><<<grid, 256, 0, cs>>>((const uint64_t*)v, span->begin, span->count, ch.dev, out);
    count_launch();
    free_codec(&ch, cs);
    CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}

int flashe_rng_uniform(flashe_ctx* ctx, uint64_t rng_seed, uint64_t rng_stream, int resolution, uint64_t begin, uint64_t count, double* out, void* stream) {
    ENTER(ctx);
    if (count == 0) return FLASHE_OK;
    if (!out) return fail(FLASHE_EINVAL, "out is NULL");
    if (resolution != FLASHE_NOISE_53 && resolution != FLASHE_NOISE_32) return fail(FLASHE_EINVAL, "unknown noise resolution");
    flashe_noise n; n.u = nullptr; n.rng_seed = rng_seed; n.rng_stream = rng_stream; n.resolution = resolution; n.reserved = 0;
    NoiseDev nz; make_noise(&n, 0, &nz);
    if (nz.res32) k_rng_uniform<true><<<GRID_OCC(ctx, k_rng_uniform<true>, count, 256), 256, 0, cs>>>(nz, begin, count, out);
    else k_rng_uniform<false><<<GRID_OCC(ctx, k_rng_uniform<false>, count, 256), 256, 0, cs>>>(nz, begin, count, out);
    count_launch();
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
    k_batch_pack<<<GRID_OCC(ctx, k_batch_pack, nw, 256), 256, 0, cs>>>(q, count, lane, bs, nw, (u128*)words_out);
    count_launch();
    CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}

int flashe_batch_unpack(flashe_ctx* ctx, const void* words, uint64_t nwords, int element_bits, int factor, uint32_t* q_out, void* stream) {
    ENTER(ctx);
    uint32_t lane, bs; int rc = batch_geometry(ctx, element_bits, factor, &lane, &bs); if (rc) return rc;
    if (nwords == 0) return FLASHE_OK;
    if (!words || !q_out) return fail(FLASHE_EINVAL, "NULL buffer");
    k_batch_unpack<<<GRID_OCC(ctx, k_batch_unpack, nwords, 256), 256, 0, cs>>>((const u128*)words, nwords, lane, bs, q_out);
    count_launch();
    CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}

int flashe_batch_layout(int int_bits, int element_bits, int factor, const uint64_t* seg_end, int nseg, uint64_t* word_end_out) {
    const int l = element_bits + factor;
    if (int_bits < 1 || int_bits > 128 || element_bits < 1 || factor < 0 || l > 32) return fail(FLASHE_EINVAL, "bad lane geometry");
    const uint64_t bs = (uint64_t)(int_bits / l);
    if (bs == 0) return fail(FLASHE_EINVAL, "int_bits smaller than one lane");
    if (nseg < 1 || !seg_end || !word_end_out) return fail(FLASHE_EINVAL, "need nseg >= 1, seg_end and word_end_out");
    uint64_t prev = 0, words = 0;
    for (int s = 0; s < nseg; ++s) {
        if (seg_end[s] < prev) return fail(FLASHE_EINVAL, "seg_end must be ascending");
        words += ceil_div(seg_end[s] - prev, bs);
        word_end_out[s] = words;
        prev = seg_end[s];
    }
    return FLASHE_OK;
}

// device copy of the layer table {first element, first word} (nseg + 1 entries each)
static int upload_layer_table(flashe_ctx* ctx, const uint64_t* seg_end, int nseg, int element_bits, int factor, cudaStream_t cs,
                              uint64_t** d_out, uint64_t* nwords) {
    std::vector<uint64_t> tab(2 * (size_t)(nseg + 1));
    std::vector<uint64_t> wend((size_t)nseg);
    int rc = flashe_batch_layout(ctx->int_bits, element_bits, factor, seg_end, nseg, wend.data()); if (rc) return rc;
    tab[0] = 0; tab[nseg + 1] = 0;
    for (int s = 0; s < nseg; ++s) { tab[s + 1] = seg_end[s]; tab[nseg + 1 + s + 1] = wend[s]; }
    *nwords = wend[nseg - 1];
    CUDA_TRY(cudaMallocAsync((void**)d_out, sizeof(uint64_t) * tab.size(), cs));
    CUDA_TRY(cudaMemcpyAsync(*d_out, tab.data(), sizeof(uint64_t) * tab.size(), cudaMemcpyHostToDevice, cs));
    CUDA_TRY(cudaStreamSynchronize(cs));   // tab (pageable) goes out of scope
    return FLASHE_OK;
}

int flashe_batch_pack_layers(flashe_ctx* ctx, const uint32_t* q, const uint64_t* seg_end, int nseg, int element_bits, int factor,
                             void* words_out, void* stream) {
    ENTER(ctx);
    uint32_t lane, bs; int rc = batch_geometry(ctx, element_bits, factor, &lane, &bs); if (rc) return rc;
    uint64_t* tab = nullptr; uint64_t nw = 0;
    rc = upload_layer_table(ctx, seg_end, nseg, element_bits, factor, cs, &tab, &nw); if (rc) return rc;
    if (nw) {
        if (!q || !words_out) { cudaFreeAsync(tab, cs); return fail(FLASHE_EINVAL, "NULL buffer"); }
        k_batch_pack_layers<<<GRID_OCC(ctx, k_batch_pack_layers, nw, 256), 256, 0, cs>>>(q, tab, tab + nseg + 1, nseg, lane, bs, nw, (u128*)words_out);
        count_launch();
    }
    cudaFreeAsync(tab, cs);
    CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}

int flashe_batch_unpack_layers(flashe_ctx* ctx, const void* words, const uint64_t* seg_end, int nseg, int element_bits, int factor,
                               uint32_t* q_out, void* stream) {
    ENTER(ctx);
    uint32_t lane, bs; int rc = batch_geometry(ctx, element_bits, factor, &lane, &bs); if (rc) return rc;
    uint64_t* tab = nullptr; uint64_t nw = 0;
    rc = upload_layer_table(ctx, seg_end, nseg, element_bits, factor, cs, &tab, &nw); if (rc) return rc;
    if (nw) {
        if (!words || !q_out) { cudaFreeAsync(tab, cs); return fail(FLASHE_EINVAL, "NULL buffer"); }
        k_batch_unpack_layers<<<GRID_OCC(ctx, k_batch_unpack_layers, nw, 256), 256, 0, cs>>>((const u128*)words, tab, tab + nseg + 1, nseg, lane, bs, nw, q_out);
        count_launch();
    }
    cudaFreeAsync(tab, cs);
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
    count_launch(k ? 2 : 1);
    CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
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
    // one launch for all pairs, every search confined to the neighbour's run of the entry's tile (k_sparse_overlap_tiled)
    static const bool tiled = [] { const char* ev = getenv("FLASHE_SPARSE_TILED"); return !(ev && ev[0] == '0'); }();
    bool tiled_ok = tiled && total > 0 && e == cudaSuccess;
    uint64_t kmax = 0;
    for (int i = 0; i < n; ++i) { tiled_ok = tiled_ok && k[i] < (1ull << 32) && (k[i] == 0 || index[i]); if (k[i] > kmax) kmax = k[i]; }
    if (tiled_ok && kmax && n <= 64) {
        // per-client bitmaps of the tile: the largest power-of-two tile with n bitmaps in OVB_BYTES of shared memory
        uint32_t tile_log2 = 5;
        while (tile_log2 < 16 && ((uint64_t)n << (tile_log2 + 1 - 3)) <= OVB_BYTES) ++tile_log2;
        uint8_t* ws; SparseClient* cl; uint32_t* splits; uint64_t n_tiles; SparseClientsParam P;
        int rc = sparse_prepare(ctx, nullptr, index, k, nullptr, 0, n, total, tile_log2, cs, &ws, &cl, &P, &splits, &n_tiles);
        if (rc) { cudaFreeAsync(d, cs); return rc; }
        const int grid = GRID_OCC(ctx, k_sparse_overlap_bitmap, n_tiles * 256, 256);
        k_sparse_overlap_bitmap<<<grid, 256, 0, cs>>>(cl, P, n, splits, tile_log2, n_tiles, d);
        count_launch();
        e = cudaGetLastError();
        cudaFreeAsync(ws, cs);
    } else if (tiled_ok && kmax && n <= 256) {
        // tile size: the runs of all clients in a tile should fit the kernel's shared-memory buffer twice over on average
        uint64_t ksum = 0;
        for (int i = 0; i < n; ++i) ksum += k[i];
        uint32_t tile_log2 = 8;
        while (tile_log2 < 20 && ((ksum << (tile_log2 + 1)) / total) * 2 <= OV_CAP) ++tile_log2;
        uint8_t* ws; SparseClient* cl; uint32_t* splits; uint64_t n_tiles; SparseClientsParam P;
        int rc = sparse_prepare(ctx, nullptr, index, k, nullptr, 0, n, total, tile_log2, cs, &ws, &cl, &P, &splits, &n_tiles);
        if (rc) { cudaFreeAsync(d, cs); return rc; }
        const int grid = GRID_OCC(ctx, k_sparse_overlap_tiled, n_tiles * 256, 256);
        k_sparse_overlap_tiled<<<grid, 256, 0, cs>>>(cl, P, n, splits, tile_log2, n_tiles, d);
        count_launch();
        e = cudaGetLastError();
        cudaFreeAsync(ws, cs);
    } else
    for (int i = 0; e == cudaSuccess && i + 1 < n; ++i) {
        if (k[i] == 0 || k[i + 1] == 0) continue;
        k_overlap<<<GRID_OCC(ctx, k_overlap, k[i], 256), 256, 0, cs>>>(index[i], k[i], index[i + 1], k[i + 1], d + i);
        count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(overlap_out, d, sizeof(unsigned long long) * (size_t)(n - 1), cudaMemcpyDeviceToHost, cs);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cs);
    cudaFreeAsync(d, cs);
    if (e != cudaSuccess) return fail(FLASHE_ECUDA, std::string("flashe_sparse_overlap: ") + cudaGetErrorString(e));
    return FLASHE_OK;
}

}  // extern "C"
