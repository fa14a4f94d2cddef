// flashe_stream_decl.h — parameter blocks of the stream kernel and the per-mode launchers (one translation
// unit each, flashe_stream_*.cu), shared with the host side in flashe_kernels.cu.
#ifndef FLASHE_STREAM_DECL_H
#define FLASHE_STREAM_DECL_H

#include "flashe_internal.h"
#include "flashe_device.cuh"

#define SMEM_BYTES 0x30000u  // dynamic shared memory of k_stream: covers [base, 0x30000) for base <= 0x10000
#define ITEM_BLOCKS 64u   // AES blocks per warp item (two per lane)

// ------------------------------------------------------------------------------------------------
// kernel parameter blocks (all in the constant bank)
// ------------------------------------------------------------------------------------------------
#define MAXS FLASHE_MAX_STREAMS
#ifndef STREAM_THREADS
#define STREAM_THREADS 512
#endif


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

struct IoDev {
    const void* in;  uint64_t in_stride;    // words (or floats) between consecutive clients
    void* out;       uint64_t out_stride;
    void* aux;                              // q_out (encode) / p_out (decode) / index (scatter)
    double* outf;
    uint32_t n_clients;
    uint32_t share;                         // batch double masking: compute each stream once
    uint32_t quad;                          // every buffer / stride allows 16-byte accesses per 4 elements
    uint64_t dense_len;                     // scatter: words in the dense target (indices outside are skipped)
    uint64_t elem0;                         // lane batching: first element of the shard's first word (x / u / q / outf offsets)
    const uint32_t* te0;                    // Te0 table in global memory (flashe_ctx::d_te0), source of the shared-memory tables
    uint32_t* tickets;                      // (ticket, done) counters of the dynamic deal (flashe_ticket_slot), NULL: static deal
#ifdef FLASHE_TRACE
    unsigned long long* trace;              // tuning builds only (scripts/build_variant.sh trace -DFLASHE_TRACE=1): per-warp timeline
#endif
};

enum { M_MASKS = 0, M_APPLY = 1, M_ENCODE = 2, M_DECODE = 3, M_SCATTER = 4 };


int flashe_launch_stream_masks(const flashe_ctx* ctx, const StreamTab& st, const Geom& g, const IoDev& io, const CodecDev& cd,
                               const NoiseDev& nz, cudaStream_t stream);
int flashe_launch_stream_apply(const flashe_ctx* ctx, const StreamTab& st, const Geom& g, const IoDev& io, const CodecDev& cd,
                               const NoiseDev& nz, cudaStream_t stream);
int flashe_launch_stream_encode(const flashe_ctx* ctx, const StreamTab& st, const Geom& g, const IoDev& io, const CodecDev& cd,
                                const NoiseDev& nz, cudaStream_t stream);
int flashe_launch_stream_encode_shared(const flashe_ctx* ctx, const StreamTab& st, const Geom& g, const IoDev& io,
                                       const CodecDev& cd, const NoiseDev& nz, cudaStream_t stream);
int flashe_launch_stream_encode_n32(const flashe_ctx* ctx, const StreamTab& st, const Geom& g, const IoDev& io, const CodecDev& cd,
                                    const NoiseDev& nz, cudaStream_t stream);
int flashe_launch_stream_encode_shared_n32(const flashe_ctx* ctx, const StreamTab& st, const Geom& g, const IoDev& io,
                                           const CodecDev& cd, const NoiseDev& nz, cudaStream_t stream);
int flashe_launch_stream_decode(const flashe_ctx* ctx, const StreamTab& st, const Geom& g, const IoDev& io, const CodecDev& cd,
                                const NoiseDev& nz, cudaStream_t stream);
int flashe_launch_stream_scatter(const flashe_ctx* ctx, const StreamTab& st, const Geom& g, const IoDev& io, const CodecDev& cd,
                                 const NoiseDev& nz, cudaStream_t stream);
// one AES block on the device (known-answer tests): d_in 4 words, d_out 8 words (hoisted path, generic path)
int flashe_launch_prp_block(const flashe_ctx* ctx, const uint32_t* d_in, uint32_t* d_out, cudaStream_t stream);

#endif
