/*
 * flashe_b200.h — C ABI of the B200-native FLASHE hot path (libflashe_b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every entry point names
 * the reference interface it replaces (SamuelGong/FLASHE, paths relative to the reference root;
 * `sp/` = federatedml/secureprotol/, `proc/` = federatedml/framework/homo/procedure/).
 *
 * Conventions
 *   - Device pointers are CUDA device addresses on the context's device; `stream` is a cudaStream_t
 *     passed as void* (NULL = legacy default stream).  Calls enqueue work and return; no entry
 *     point synchronises the device or the stream unless its comment says so.
 *   - All functions return FLASHE_OK (0) or a negative error code and never throw; the message of the
 *     last failure on the calling thread is available from flashe_last_error().
 *   - A context is bound to one (key, int_bits, device) and is not thread-safe.
 *   - Element storage ("words"), little-endian: int_bits <= 32 -> uint32_t, <= 64 -> uint64_t,
 *     <= 128 -> 16 bytes (lo uint64_t, hi uint64_t).  Values are always reduced mod 2^int_bits.
 *   - flashe_span describes which slice of which chunked vector a call covers.  The reference derives
 *     the AES counter from the chunking of the WHOLE vector over N_JOBS = cpu_count() workers
 *     (sp/jzf_flashe.py:7,12-16,33-34), so total_len and n_jobs are part of the ciphertext format;
 *     [begin, begin+count) selects an element-range shard (multi-GPU) of that vector, and every
 *     data pointer of the call addresses element `begin` at offset 0.
 */
#ifndef FLASHE_B200_H
#define FLASHE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLASHE_ABI_VERSION 2   /* (entry points added since 2 without changing an existing one: flashe_sparse_apply_masks_batch) */

enum {
    FLASHE_OK = 0,
    FLASHE_EINVAL = -1,       /* bad argument */
    FLASHE_ECUDA = -2,        /* CUDA runtime error (message has the cudaError string) */
    FLASHE_ENOMEM = -3,
    FLASHE_EUNSUPPORTED = -4  /* valid in the reference, not built here (message says what) */
};

enum { FLASHE_SCHEME_SINGLE = 0, FLASHE_SCHEME_DOUBLE = 1 };  /* FlasheCipher(mask=...) sp/jzf_flashe.py:230 */
enum { FLASHE_AGG_ELEMENTWISE = 0, FLASHE_AGG_PACKED = 1 };   /* proc/jzf_aggregator.py:421-430 / 404-419 */

#define FLASHE_MAX_STREAMS 128

typedef struct flashe_ctx flashe_ctx;

typedef struct flashe_span {
    uint64_t total_len; /* L of the whole vector the reference's encrypt()/decrypt() would receive */
    uint64_t begin;     /* first element of this shard */
    uint64_t count;     /* elements in this shard (pointers address `count` elements) */
    uint32_t n_jobs;    /* the reference's N_JOBS (sp/jzf_flashe.py:7) */
    uint32_t reserved;  /* must be 0 */
} flashe_span;

/* Encode parameters: QuantizingClient.quantize -> _static_quantize_padding_asymmetric
 * (sp/jzf_quantize.py:394-491, 55-67).  The flat vector is a concatenation of `nseg` layers
 * (proc/jzf_aggregator.py:625-650); layer s covers elements [seg_end[s-1], seg_end[s]) of the WHOLE
 * vector and is clipped at alpha[s].  seg_end / alpha are HOST arrays (copied by the call). */
typedef struct flashe_codec {
    int32_t element_bits;    /* 16 in every shipped config */
    int32_t n_clients;       /* decode only: num_clients of _static_unquantize_padding_asymmetric */
    int32_t nseg;            /* >= 1 */
    int32_t batch_lane_bits; /* 0: one element per word.  Otherwise lane batching (_static_batching_padding_asymmetric,
                              * sp/jzf_quantize.py:162-185; the shipped int_bits = 120 config): lanes of this many bits
                              * = element_bits + ceil(log2(num_clients)), int_bits / lanes per word, FIRST element most
                              * significant, every layer zero-padded by itself.  The call's span then counts WORDS
                              * (total_len = sum over layers of ceil(size / lanes)) while seg_end keeps counting ELEMENTS,
                              * and element-indexed pointers (x, noise.u, q_out, the float64 output) address the first
                              * element of word span.begin.  Needs 64 < int_bits <= 128. */
    const uint64_t* seg_end; /* [nseg], ascending, seg_end[nseg-1] == span.total_len (== the element count when batching) */
    const double* alpha;     /* [nseg] Python-float alphas (rounded to float32 inside, as numpy does) */
} flashe_codec;

/* Stochastic-rounding noise source: the reference draws np.random.random(size) (float64 in [0,1),
 * sp/jzf_quantize.py:64).  Either hand the same numbers in (`u` device pointer, parity mode) or let
 * the device generate them with the counter-based generator (u == NULL): element j of client
 * `rng_stream` uses Philox4x32-10(key=rng_seed, counter=(j>>1, rng_stream)) and the res53
 * construction numpy uses; flashe_rng_uniform() materialises exactly those numbers. */
enum { FLASHE_NOISE_53 = 0, FLASHE_NOISE_32 = 1 };
typedef struct flashe_noise {
    const double* u;     /* device, [count] for this shard, or NULL */
    uint64_t rng_seed;
    uint64_t rng_stream;
    int32_t resolution;  /* device generator only.  FLASHE_NOISE_53: numpy's 53-bit construction, two 32-bit words per
                          * element (above).  FLASHE_NOISE_32 (throughput mode): u_j = w * 2^-32 with w = word (j & 3)
                          * of Philox4x32-10(key = rng_seed, counter = (j >> 2, rng_stream)) — one generator call per
                          * four elements.  Rounding a 16-bit quantisation stochastically needs far fewer than 32
                          * random bits; ciphertexts still equal the oracle's when it is fed the same u
                          * (flashe_rng_uniform materialises either stream). */
    int32_t reserved;
} flashe_noise;

int flashe_abi_version(void);
const char* flashe_last_error(void);

/* Bytes per stored element for a given int_bits (4, 8 or 16), or FLASHE_EINVAL. */
int flashe_word_bytes(int int_bits);

/* FlasheCipher(int_bits) + generate_prp_seed(assigned_seed) -> PsuedoRandomPermutation.generate_key
 * -> AESCipher.generate_key(mode="ECB") (sp/jzf_flashe.py:230-295, sp/jzf_aes_prp.py:11-22,
 * sp/jzf_aes.py:14-34): `seed` of any length >= 1 is reduced to the AES-256 key by keeping its low
 * 32 bytes (left-zero-padded when shorter).  int_bits in [8, 128]. */
int flashe_ctx_create(const uint8_t* seed, size_t seed_len, int int_bits, int device, flashe_ctx** out);
int flashe_ctx_destroy(flashe_ctx* ctx);
int flashe_ctx_int_bits(const flashe_ctx* ctx);
int flashe_ctx_device(const flashe_ctx* ctx);

/* One AES-256-ECB block on the device with the context key (PsuedoRandomPermutation.get_permutation,
 * sp/jzf_aes_prp.py:24-30).  in16 / out16 are HOST buffers; synchronises `stream`.  Known-answer
 * tests only. */
int flashe_prp_block(flashe_ctx* ctx, const uint8_t in16[16], uint8_t out16[16], void* stream);

/* Combined keystream  out[j] = sum_k sign[k] * F(iter, prf_idx[k])[j] mod 2^b  — the bodies of
 * _static_prepare_encrypt(_single) / _static_prepare_decrypt(_single) (sp/jzf_flashe.py:19-152) and
 * therefore of prepare_encrypt / prepare_decrypt (sp/jzf_flashe.py:599-666; precompute for round t+k
 * is the same call with iter = t+k).  prf_idx / sign are HOST arrays, sign[k] is +1 or -1,
 * 1 <= nstreams <= FLASHE_MAX_STREAMS. */
int flashe_masks(flashe_ctx* ctx, uint32_t iter, const int32_t* prf_idx, const int32_t* sign, int nstreams,
                 const flashe_span* span, void* out, void* stream);

/* Mask precomputation for several future rounds in one call (prepare_encrypt looks one round ahead,
 * sp/jzf_flashe.py:599-631; BASELINE config 3 looks 16 ahead): for r in [0, n_rounds) the combined
 * keystream of round iter_from + r is written to out + r*round_stride words.  Same stream list for
 * every round (e.g. {idx, idx+1} with signs {+1, -1} for a client's double-masking term). */
int flashe_precompute(flashe_ctx* ctx, uint32_t iter_from, int n_rounds, const int32_t* prf_idx, const int32_t* sign,
                      int nstreams, const flashe_span* span, void* out, uint64_t round_stride, void* stream);

/* out[j] = (in[j] + sum_k sign[k] * F(iter, prf_idx[k])[j]) mod 2^b.  in may equal out.
 * The arithmetic of _multiprocessing_encrypt(_single) / _multiprocessing_decrypt(_single)
 * (sp/jzf_flashe.py:431-488, 506-582) for an arbitrary stream list. */
int flashe_apply_masks(flashe_ctx* ctx, uint32_t iter, const int32_t* prf_idx, const int32_t* sign,
                       int nstreams, const flashe_span* span, const void* in, void* out, void* stream);

/* FlasheCipher.encrypt (sp/jzf_flashe.py:490-504) of an already quantised vector: double masking adds
 * F(iter, idx) - F(iter, idx+1) (:349-353, :480), single masking adds F(iter, idx) (:308-309, :450). */
int flashe_encrypt(flashe_ctx* ctx, uint32_t iter, int32_t idx, int scheme, const flashe_span* span,
                   const void* q_in, void* ct_out, void* stream);

/* FlasheCipher.decrypt (sp/jzf_flashe.py:584-594) after set_idx_list(mode="decrypt") chose the PRF
 * index sets (sp/jzf_flashe.py:354-386): adds F(iter, a) for a in add_idx and subtracts F(iter, s)
 * for s in minus_idx.  Single masking: na = 0 and minus_idx = every surviving client
 * (:306-314, :531).  na + ns <= FLASHE_MAX_STREAMS per call (chain calls for more). */
int flashe_decrypt(flashe_ctx* ctx, uint32_t iter, const int32_t* add_idx, int na, const int32_t* minus_idx,
                   int ns, const flashe_span* span, const void* agg_in, void* p_out, void* stream);

/* out[j] = (in[j] + sign * mask[j]) mod 2^b — the online step when the keystream was produced ahead
 * of time by flashe_masks (consumption of next_iter_encrypt_prepared / next_iter_decrypt_prepared,
 * sp/jzf_flashe.py:480-486, 570-580). */
int flashe_add_premasked(flashe_ctx* ctx, const void* in, const void* mask, int sign, uint64_t count,
                         void* out, void* stream);

/* Encode only: q[j] = floor(float64(clip32(x[j])) + u[j])  (_static_quantize_padding_asymmetric,
 * sp/jzf_quantize.py:55-67).  q_out is uint32_t[count] whatever int_bits is. */
int flashe_encode(flashe_ctx* ctx, const flashe_span* span, const float* x, const flashe_codec* codec,
                  const flashe_noise* noise, uint32_t* q_out, void* stream);

/* Fused encode + encrypt of one client's float32 gradient shard: quantize (above) then
 * FlasheCipher.encrypt.  Replaces the pair QuantizingClient.quantize / weights.encrypted(cipher)
 * (proc/jzf_aggregator.py:722, 739).  q_out may be NULL; when given it receives the quantised
 * plaintext as uint32_t. */
int flashe_encode_encrypt(flashe_ctx* ctx, uint32_t iter, int32_t idx, int scheme, const flashe_span* span,
                          const float* x, const flashe_codec* codec, const flashe_noise* noise,
                          void* ct_out, uint32_t* q_out, void* stream);

/* Same, for n_clients logical clients idx0 .. idx0+n_clients-1 hosted on this device in ONE launch.
 * Client c reads x + c*x_stride floats (and noise->u + c*u_stride doubles when u != NULL; device
 * noise uses stream id noise->rng_stream + c) and writes ct_out + c*ct_stride words.  Every client's
 * ciphertext is produced exactly as by flashe_encode_encrypt.  share_streams != 0 computes
 * F(iter, c+1) once and uses it as client c's minus term and client c+1's add term (same outputs,
 * about half the AES work; only meaningful for FLASHE_SCHEME_DOUBLE). */
int flashe_encode_encrypt_batch(flashe_ctx* ctx, uint32_t iter, int32_t idx0, int n_clients, int scheme,
                                const flashe_span* span, const float* x, uint64_t x_stride,
                                const flashe_codec* codec, const flashe_noise* noise, uint64_t u_stride,
                                void* ct_out, uint64_t ct_stride, int share_streams, void* stream);

/* Fused encode + add of a precomputed combined mask (the online path after prepare_encrypt):
 * ct[j] = (quantize(x[j]) + mask[j]) mod 2^b. */
int flashe_encode_add_premasked(flashe_ctx* ctx, const flashe_span* span, const float* x,
                                const flashe_codec* codec, const flashe_noise* noise, const void* mask,
                                void* ct_out, void* stream);

/* Same for n_clients logical clients hosted on this device in ONE launch (the online round of the mask
 * precomputation schedule): client c reads x + c*x_stride floats and mask + c*mask_stride words, writes
 * ct_out + c*ct_stride words, and draws its noise from noise->u + c*u_stride or stream id rng_stream + c. */
int flashe_encode_add_premasked_batch(flashe_ctx* ctx, const flashe_span* span, int n_clients, const float* x,
                                      uint64_t x_stride, const flashe_codec* codec, const flashe_noise* noise,
                                      uint64_t u_stride, const void* mask, uint64_t mask_stride, void* ct_out,
                                      uint64_t ct_stride, void* stream);

/* Server sum of n ciphertext vectors (vector c at cts + c*stride words):
 *   FLASHE_AGG_ELEMENTWISE  out[j] = sum_c ct_c[j] mod 2^b           (proc/jzf_aggregator.py:421-430)
 *   FLASHE_AGG_PACKED       the (x + y) % (1 << (b*L)) sum of the packed wire integers
 *                           (proc/jzf_aggregator.py:404-419 on framework/jzf_weights.py:45-84): carries
 *                           out of element j leak into element j-1.  carry_in enters at the LAST
 *                           element of this range (0 for a whole vector; for an element-range shard
 *                           it is the carry out of the next shard).  If carry_out (device, uint32_t[4])
 *                           is not NULL it receives {c, depends, A, T}: c = carry out of element 0 for
 *                           the given carry_in; depends = 0 when that carry cannot depend on carry_in
 *                           (always, for ciphertext-like data), else the exact transfer function of the
 *                           range is carry_out = A + (carry_in >= T).  Shards must be non-empty.
 *                           Every width up to 128 bits (the shipped int_bits = 120 batch mode included);
 *                           n <= 2^int_bits (a digit hands on at most +1 beyond its own quotient). */
int flashe_aggregate(flashe_ctx* ctx, const void* cts, uint64_t stride, int n, uint64_t count, int mode,
                     uint32_t carry_in, void* out, uint32_t* carry_out, void* stream);

/* Multi-GPU packed sum, step 2: after the shards' descriptors were exchanged, ripple the true
 * carry-in (carry out of the NEXT shard) into a shard that was aggregated with carry_in = 0. */
int flashe_aggregate_carry_fixup(flashe_ctx* ctx, void* out, uint64_t count, uint32_t carry_in, void* stream);

/* Decode only: out[j] = v[j] * (2*alpha*n) / ((2^e - 1) * n) - alpha*n in float64
 * (_static_unquantize_padding_asymmetric, sp/jzf_quantize.py:102-107).  v holds words of this
 * context's width (int_bits <= 64). */
int flashe_decode(flashe_ctx* ctx, const flashe_span* span, const void* v, const flashe_codec* codec,
                  double* out, void* stream);

/* Fused decrypt + decode of the aggregate: FlasheCipher.decrypt then QuantizingClient.unquantize
 * (proc/jzf_aggregator.py:883-899).  p_out may be NULL; when given it receives the decrypted integers
 * (words of this context's width). */
int flashe_decrypt_decode(flashe_ctx* ctx, uint32_t iter, const int32_t* add_idx, int na,
                          const int32_t* minus_idx, int ns, const flashe_span* span, const void* agg_in,
                          const flashe_codec* codec, double* out, void* p_out, void* stream);

/* The device noise generator made visible: out[j - begin] = u_j for j in [begin, begin+count). */
int flashe_rng_uniform(flashe_ctx* ctx, uint64_t rng_seed, uint64_t rng_stream, int resolution, uint64_t begin,
                       uint64_t count, double* out, void* stream);

/* Lane batching for int_bits up to 128 (shipped: 120) — _static_batching_padding_asymmetric /
 * _static_unbatching_padding_asymmetric (sp/jzf_quantize.py:162-185, 234-251): lane width
 * element_bits + factor, batch_size = int_bits / lane, zero padded, FIRST element most significant.
 * pack: q uint32_t[count] -> ceil(count/batch_size) words.  unpack: nwords words ->
 * uint32_t[nwords*batch_size]. */
int flashe_batch_pack(flashe_ctx* ctx, const uint32_t* q, uint64_t count, int element_bits, int factor,
                      void* words_out, void* stream);
int flashe_batch_unpack(flashe_ctx* ctx, const void* words, uint64_t nwords, int element_bits, int factor,
                        uint32_t* q_out, void* stream);

/* Lane batching of a whole model through its layer table: QuantizingClient.quantize batches
 * each layer by itself (sp/jzf_quantize.py:447-452) and flatten_weights concatenates the layers
 * (proc/jzf_aggregator.py:625-650), so layer s — elements [seg_end[s-1], seg_end[s]) of the flat quantised
 * vector — is zero-padded to a multiple of batch_size and owns words [word_end[s-1], word_end[s]).
 * flashe_batch_layout (pure host arithmetic) returns word_end[nseg]; pack reads q[seg_end[nseg-1]] and writes
 * word_end[nseg-1] words; unpack is the inverse and drops the padding lanes.  seg_end: HOST array. */
int flashe_batch_layout(int int_bits, int element_bits, int factor, const uint64_t* seg_end, int nseg,
                        uint64_t* word_end_out);
int flashe_batch_pack_layers(flashe_ctx* ctx, const uint32_t* q, const uint64_t* seg_end, int nseg,
                             int element_bits, int factor, void* words_out, void* stream);
int flashe_batch_unpack_layers(flashe_ctx* ctx, const void* words, const uint64_t* seg_end, int nseg,
                               int element_bits, int factor, uint32_t* q_out, void* stream);

/* Index-sparse path, masking scheme "single" (the only one that works in the reference).
 * expand_to_dense (proc/jzf_aggregator.py:150-165): dense[j] = zero_word for every j, then
 * dense[index[i]] = compact[i].  index: device int64_t[k], sorted unique, < total.  zero_word: HOST
 * pointer to one word. */
int flashe_sparse_expand(flashe_ctx* ctx, const void* compact, const int64_t* index, uint64_t k,
                         uint64_t total, const void* zero_word, void* dense_out, void* stream);

/* The arbiter's sum of the expanded uploads (expand_to_dense for every client, then the element-wise
 * reduce of proc/jzf_aggregator.py:421-430) without materialising the n dense vectors:
 *   dense_out[j] = sum_c (j in index_c ? compact_c[.] : zero_c)   mod 2^int_bits
 * built tile by tile of the dense vector in shared memory (start from sum_c zero_c, add compact_c - zero_c for the
 * entries of all clients, one coalesced write): O(total + sum k_c) instead of O(n * total) bytes.  compacts / indexes: HOST arrays of n device pointers (words
 * and int64_t, index sorted unique); ks: HOST uint64_t[n]; zero_words: HOST array of n words. */
int flashe_sparse_sum(flashe_ctx* ctx, const void* const* compacts, const int64_t* const* indexes,
                      const uint64_t* ks, const void* zero_words, int n_clients, uint64_t total,
                      void* dense_out, void* stream);

/* Sparse single-mask decrypt term (sp/jzf_flashe.py:315-343, 528-535): regenerate
 * sum_k sign[k]*F(iter, prf_idx[k]) over the COMPACT positions described by `span` (total_len = k) and
 * add it into dense[index[i]].  `dense` holds dense_len words; index must be sorted unique (a repeated
 * index would be a racy read-modify-write); entries outside [0, dense_len) are skipped (the reference
 * raises IndexError there; the Python mirror validates the list before launching). */
int flashe_sparse_apply_masks(flashe_ctx* ctx, uint32_t iter, const int32_t* prf_idx, const int32_t* sign,
                              int nstreams, const flashe_span* span, const int64_t* index, void* dense,
                              uint64_t dense_len, void* stream);

/* The same term for EVERY client in one call — what a client's decrypt of the sparse aggregate needs
 * (sp/jzf_flashe.py:315-343: "for every client c regenerate F(t,c) over len(mask_c) compact positions, scatter to
 * dense, sum"): dense[index_c[i]] += sign * F(iter, prf_idx[c])[i] for c < n_clients, i < ks[c], the chunk rule
 * applied to each list's own length ks[c] with n_jobs chunks.  The masks of all clients are generated into a
 * workspace (one launch per run of equal list lengths) and added tile by tile of the dense vector in shared
 * memory: two passes over `dense` instead of one read-modify-write of a sector per entry and one launch per client.
 * prf_idx, ks, indexes: HOST arrays of n_clients entries (indexes: device int64_t lists, sorted unique, entries
 * outside [0, dense_len) are skipped); sign: +1 or -1.  Uploads a small table: not capturable into a CUDA graph. */
int flashe_sparse_apply_masks_batch(flashe_ctx* ctx, uint32_t iter, const int32_t* prf_idx, int sign, int n_clients,
                                    const uint64_t* ks, uint32_t n_jobs, const int64_t* const* indexes, void* dense,
                                    uint64_t dense_len, void* stream);

/* dynamic_masking cost model (proc/jzf_flashe_block.py:89-117): overlap[i] = |mask_i ∩ mask_{i+1}| for
 * the n-1 adjacent pairs; index lists are device int64_t arrays (sorted unique); `index` and `k` are
 * HOST arrays of n entries; overlap_out is a HOST uint64_t[n-1]; synchronises `stream`. */
int flashe_sparse_overlap(flashe_ctx* ctx, const int64_t* const* index, const uint64_t* k, int n,
                          uint64_t total, uint64_t* overlap_out, void* stream);

/* ---- wire bit-packing --------------------------------------------------------------
 * _to_bytes(flatten_array, num_bits) (framework/jzf_weights.py:45-84, used by
 * JZFTransferableWeights.compress :155-198 and by Client.sparsify for the index list,
 * proc/jzf_aggregator.py:617) builds the integer s = sum_j a[j] << ((L-1-j)*bits): first element most
 * significant.  Here s is the big-endian byte string of ceil(L*bits/8) bytes (int.to_bytes(n, 'big');
 * zero bits above L*bits).  `bits` is the field width (int_bits for ciphertexts,
 * total.bit_length() for index lists), word_bytes (4, 8 or 16) the storage width of one element;
 * elements must be < 2^bits (they are masked).  `out` must be 16-byte aligned. */
int flashe_wire_nbytes(int bits, uint64_t count, uint64_t* nbytes_out);
int flashe_wire_pack(flashe_ctx* ctx, const void* words, int word_bytes, uint64_t count, int bits,
                     uint8_t* out, void* stream);
/* _from_bytes(s, l, num_bits) followed by the reverse() of decompress (framework/jzf_weights.py:98-137,
 * 224): element j = (s >> ((L-1-j)*bits)) & (2^bits - 1). */
int flashe_wire_unpack(flashe_ctx* ctx, const uint8_t* in, uint64_t count, int bits, int word_bytes,
                       void* words_out, void* stream);

/* ---- layer-wise top-k sparsification with residual accumulation -------------------
 * Client.sparsify (proc/jzf_aggregator.py:578-623).  The flat float32 vector x is a concatenation of
 * nseg layers ending at seg_end[s] (HOST array); layer s keeps k[s] = max(1, floor(sparsity*size))
 * elements (HOST array, computed by the caller as the reference does, :598).  Per layer:
 *   selection  the k[s] largest |x| (the residual is NOT part of the ranking, :594-596); elements tied
 *              with the k-th largest are taken from the highest indices first (stable-argsort order;
 *              numpy's default sort leaves exact ties unspecified);
 *   values_out float32 (x + residual_in) at the selected positions, in ascending index order (:599-600);
 *   index_out  their GLOBAL positions (location + base, :606), ascending, int64;
 *   residual_out  x + residual_in with the selected positions zeroed (:603-604).
 * residual_in may be NULL (first round: no residual yet); residual_out may be NULL or alias
 * residual_in.  values_out / index_out hold sum(k) entries, layer after layer. */
int flashe_topk_sparsify(flashe_ctx* ctx, const float* x, const float* residual_in, uint64_t total,
                         const uint64_t* seg_end, const uint64_t* k, int nseg, float* values_out,
                         int64_t* index_out, float* residual_out, void* stream);

/* ---- per-layer statistics around decode -------------------------------------------
 * QuantizingClient.unnormalize (sp/jzf_quantize.py:549-564): per layer s, w += shift[s] (the past
 * mean), then stats_out[2s] = np.mean(w), stats_out[2s+1] = np.std(w) — BIT-EXACT: the sums run in numpy's
 * own order, because this round's std becomes next round's clipping range alpha (sp/jzf_quantize.py:403-413)
 * and alpha enters every ciphertext.
 *   FLASHE_SUM_PAIRWISE    float64 ndarray semantics (the batched mode's arrays): 0.0 + numpy's pairwise
 *                          summation (blocks of <= 128 with eight accumulators, halves split at multiples of 8)
 *   FLASHE_SUM_SEQUENTIAL  object-array semantics (the un-batched mode holds Python floats): a[0] + a[1] + ...
 * mean = sum / n; std = sqrt(sum((w - mean)^2) / n), every operation correctly rounded, no FMA contraction.
 * w / w_out: device float64 [total] (w_out may be NULL = statistics only, or alias w); seg_end / shift:
 * HOST arrays (shift may be NULL = 0); stats_out: DEVICE double[2*nseg]; an empty layer yields nan, nan. */
enum { FLASHE_SUM_PAIRWISE = 0, FLASHE_SUM_SEQUENTIAL = 1 };
int flashe_segment_stats(flashe_ctx* ctx, const double* w, double* w_out, uint64_t total,
                         const uint64_t* seg_end, const double* shift, int nseg, int order, double* stats_out,
                         void* stream);

/* ---- peer buffers: the gather of a sharded result without a collective (SURVEY §8 e) --------------
 * One process per GPU.  The reference has no counterpart (its arbiter is one CPU process holding the whole
 * vector, proc/jzf_aggregator.py:404-430); here the decoded shards of G GPUs have to meet on the GPU that owns
 * the model.  flashe_peer_alloc allocates `bytes` of device memory with cudaMalloc (outside any caching
 * allocator, so that it can be exported) and fills a 64-byte handle (cudaIpcMemHandle_t) that other processes of
 * the same node pass to flashe_peer_open; the pointer that comes back addresses the OWNER's memory over NVLink
 * and can be given to any entry of this library as an output buffer — e.g. flashe_decrypt_decode(..., out =
 * peer + 8 * shard_begin) makes the decode kernel store its shard straight into the gathering GPU's vector: the
 * transfer overlaps the PRF work element by element, there is no staging copy and no all-gather afterwards.
 * The owner must not read the buffer before the writers' streams have completed (a barrier after their
 * synchronisation is enough); flashe_peer_close unmaps, flashe_peer_free releases the owner's allocation. */
#define FLASHE_PEER_HANDLE_BYTES 64
int flashe_peer_alloc(flashe_ctx* ctx, uint64_t bytes, void** ptr_out, uint8_t handle_out[FLASHE_PEER_HANDLE_BYTES]);
int flashe_peer_open(flashe_ctx* ctx, const uint8_t handle[FLASHE_PEER_HANDLE_BYTES], void** ptr_out);
int flashe_peer_close(flashe_ctx* ctx, void* ptr);
int flashe_peer_free(flashe_ctx* ctx, void* ptr);

/* Number of kernels this library has launched on the calling process since load (bench bookkeeping). */
uint64_t flashe_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FLASHE_B200_H */
