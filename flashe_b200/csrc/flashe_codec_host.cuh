// flashe_codec_host.cuh — host-side construction of the kernel parameter blocks for encode / decode
// (layer table, exact reciprocals) and for the noise generator; included by every translation unit that
// launches a kernel taking CodecDev / NoiseDev.
#ifndef FLASHE_CODEC_HOST_CUH
#define FLASHE_CODEC_HOST_CUH

#include <math.h>
#include <string.h>
#include <vector>

#include "flashe_internal.h"
#include "flashe_device.cuh"

struct CodecHost {
    CodecDev dev;
    uint64_t elem0;   // lane batching: first element of the span's first word
    Seg* table;  // device allocation to free (stream ordered) or NULL
};

static inline int make_codec(const flashe_ctx* ctx, const flashe_span* span, const flashe_codec* c, bool decode, cudaStream_t stream,
                      CodecHost* out) {
    (void)ctx;
    memset(out, 0, sizeof(*out));
    if (!c) return flashe_fail(FLASHE_EINVAL, "codec is NULL");
    if (c->element_bits < 1 || c->element_bits > 24) return flashe_fail(FLASHE_EINVAL, "element_bits must be in [1, 24]");
    if (c->nseg < 1 || !c->seg_end || !c->alpha) return flashe_fail(FLASHE_EINVAL, "codec needs nseg >= 1, seg_end and alpha");
    if (decode && c->n_clients < 1) return flashe_fail(FLASHE_EINVAL, "codec.n_clients must be >= 1 for decode");
    // lane batching: the span counts WORDS, seg_end counts ELEMENTS; every layer is padded by itself
    const int lane_bits = c->batch_lane_bits;
    uint32_t bs = 0;
    if (lane_bits) {
        if (ctx->words != 4) return flashe_fail(FLASHE_EUNSUPPORTED, "lane batching is built for 64 < int_bits <= 128 (shipped: 120)");
        if (lane_bits < c->element_bits || lane_bits > 32) return flashe_fail(FLASHE_EINVAL, "batch_lane_bits must be in [element_bits, 32]");
        bs = (uint32_t)(ctx->int_bits / lane_bits);
        if (bs == 0) return flashe_fail(FLASHE_EINVAL, "int_bits smaller than one lane");
    } else if (c->seg_end[c->nseg - 1] != span->total_len) {
        return flashe_fail(FLASHE_EINVAL, "seg_end[nseg-1] must equal span.total_len");
    }
    std::vector<Seg> segs((size_t)c->nseg);
    uint64_t words = 0;
    const int n = decode ? c->n_clients : 1;
    for (int s = 0; s < c->nseg; ++s) {
        if (s && c->seg_end[s] < c->seg_end[s - 1]) return flashe_fail(FLASHE_EINVAL, "seg_end must be ascending");
        segs[s].end = c->seg_end[s];
        if (bs) words += (c->seg_end[s] - (s ? c->seg_end[s - 1] : 0) + bs - 1) / bs;
        segs[s].wend = words;
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
    if (bs && words != span->total_len) return flashe_fail(FLASHE_EINVAL, "span.total_len must equal the number of batched words of the layers");
    CodecDev& d = out->dev;
    d.nseg = c->nseg; d.ebits = c->element_bits;
    d.lane_bits = (uint32_t)lane_bits; d.bs = bs;
    out->elem0 = 0;
    if (bs) {                     // first element of word span->begin: what every element-indexed pointer of the call addresses
        int lo = 0;
        while (lo < c->nseg - 1 && span->begin >= segs[lo].wend) ++lo;
        const uint64_t ebeg = lo ? segs[lo - 1].end : 0, wbeg = lo ? segs[lo - 1].wend : 0;
        out->elem0 = span->begin >= words ? segs[c->nseg - 1].end : ebeg + (span->begin - wbeg) * bs;
    }
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
static inline void free_codec(CodecHost* c, cudaStream_t stream) { if (c->table) cudaFreeAsync(c->table, stream); }

static inline int check_noise(const flashe_noise* nz) {
    if (nz && nz->resolution != FLASHE_NOISE_53 && nz->resolution != FLASHE_NOISE_32) return flashe_fail(FLASHE_EINVAL, "unknown noise resolution");
    if (nz && nz->reserved != 0) return flashe_fail(FLASHE_EINVAL, "noise.reserved must be 0");
    return FLASHE_OK;
}
static inline void make_noise(const flashe_noise* nz, uint64_t u_stride, NoiseDev* d) {
    memset(d, 0, sizeof(*d));
    if (!nz) return;
    d->u = nz->u; d->u_stride = u_stride;
    uint32_t k0 = (uint32_t)nz->rng_seed, k1 = (uint32_t)(nz->rng_seed >> 32);
    for (int i = 0; i < 10; ++i) { d->rk[i][0] = k0; d->rk[i][1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    d->stream = nz->rng_stream;
    d->res32 = nz->resolution == FLASHE_NOISE_32 ? 1u : 0u;
}


#endif
