/*
 * flashe_oracle.c — CPU restatement of the FLASHE hot path (TEST INFRASTRUCTURE, NOT PRODUCT).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  The product (flashe_b200/) never links, imports or calls it; it fails loudly when
 * the CUDA library is missing.
 *
 * Parity status: the reference (SamuelGong/FLASHE) ships NO tests or golden vectors for this path
 * ("parity unpinned" by its own tests).  This oracle is therefore pinned by
 *   (1) FIPS-197 Appendix C.3 for the third-party AES primitive (pycryptodome 3.9.9,
 *       requirements.txt:144, is not vendored), and
 *   (2) tests/golden/flashe_golden.npz — outputs of the reference's own Python modules executed in
 *       the build container by tests/golden/make_golden.py (committed next to the fixtures).
 *
 * Every function cites the reference lines it restates (paths relative to the reference root,
 * federatedml/secureprotol/ unless stated otherwise).  Nothing here is copied: the reference is
 * Python; this is plain C99 written from the behaviour.
 *
 * Word layout shared with the device library: an element of int_bits b is stored little-endian in
 * 4 bytes (b <= 32), 8 bytes (b <= 64) or 16 bytes (b <= 128).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>

typedef unsigned __int128 u128;

#define FO_OK 0
#define FO_EINVAL (-1)

/* Minimal fork-join over [0, n): the reference fans its per-chunk loops out over
 * multiprocessing.Pool(cpu_count()) (jzf_flashe.py:433-441); the threads here only split index
 * ranges, the arithmetic per index is unchanged.  fo_set_threads(1) gives the scalar port. */
static int g_threads = 1;
void fo_set_threads(int t) { g_threads = t < 1 ? 1 : (t > 256 ? 256 : t); }
int fo_num_threads(void) { return g_threads; }

typedef void (*range_fn)(int64_t lo, int64_t hi, void* ctx);
typedef struct { range_fn fn; void* ctx; int64_t lo, hi; } par_task;
static void* par_entry(void* p) { par_task* t = (par_task*)p; t->fn(t->lo, t->hi, t->ctx); return NULL; }
static void parallel_for(int64_t n, range_fn fn, void* ctx) {
    int T = g_threads;
    if (T <= 1 || n < 2 * T) { fn(0, n, ctx); return; }
    pthread_t th[256]; par_task tk[256];
    for (int i = 0; i < T; ++i) {
        tk[i].fn = fn; tk[i].ctx = ctx; tk[i].lo = n * i / T; tk[i].hi = n * (i + 1) / T;
        if (i + 1 < T) pthread_create(&th[i], NULL, par_entry, &tk[i]);
    }
    par_entry(&tk[T - 1]);
    for (int i = 0; i + 1 < T; ++i) pthread_join(th[i], NULL);
}

/* ------------------------------------------------------------------------------------------------
 * AES-256 block encryption, FIPS-197, byte oriented (deliberately NOT the T-table form the device
 * code uses, so the two implementations share no structure).
 * Reference call site: jzf_aes.py:33-41 (AES.new(key, MODE_ECB).encrypt), jzf_aes_prp.py:24-30.
 * ---------------------------------------------------------------------------------------------- */
static uint8_t SBOX[256];
static int sbox_ready = 0;

static uint8_t gf_mul(uint8_t a, uint8_t b) {
    uint8_t p = 0;
    for (int i = 0; i < 8; ++i) {
        if (b & 1) p ^= a;
        uint8_t hi = a & 0x80;
        a = (uint8_t)(a << 1);
        if (hi) a ^= 0x1b;
        b >>= 1;
    }
    return p;
}

static void sbox_init(void) {
    if (sbox_ready) return;
    for (int x = 0; x < 256; ++x) {
        uint8_t inv = 0;
        if (x) {
            for (int y = 1; y < 256; ++y)
                if (gf_mul((uint8_t)x, (uint8_t)y) == 1) { inv = (uint8_t)y; break; }
        }
        uint8_t s = inv, r = inv;
        for (int i = 0; i < 4; ++i) { r = (uint8_t)((r << 1) | (r >> 7)); s ^= r; }
        SBOX[x] = s ^ 0x63;
    }
    sbox_ready = 1;
}

void fo_aes256_expand(const uint8_t key[32], uint8_t rk[240]) {
    sbox_init();
    memcpy(rk, key, 32);
    uint8_t rcon = 1;
    for (int i = 8; i < 60; ++i) {
        uint8_t t[4];
        memcpy(t, rk + 4 * (i - 1), 4);
        if (i % 8 == 0) {
            uint8_t t0 = t[0];
            t[0] = SBOX[t[1]] ^ rcon; t[1] = SBOX[t[2]]; t[2] = SBOX[t[3]]; t[3] = SBOX[t0];
            rcon = gf_mul(rcon, 2);
        } else if (i % 8 == 4) {
            for (int k = 0; k < 4; ++k) t[k] = SBOX[t[k]];
        }
        for (int k = 0; k < 4; ++k) rk[4 * i + k] = rk[4 * (i - 8) + k] ^ t[k];
    }
}

void fo_aes256_encrypt_block(const uint8_t rk[240], const uint8_t in[16], uint8_t out[16]) {
    uint8_t s[16], t[16];
    for (int i = 0; i < 16; ++i) s[i] = in[i] ^ rk[i];
    for (int round = 1; round <= 14; ++round) {
        /* SubBytes + ShiftRows: state is column major, byte (r,c) at s[4c+r] */
        for (int c = 0; c < 4; ++c)
            for (int r = 0; r < 4; ++r) t[4 * c + r] = SBOX[s[4 * ((c + r) & 3) + r]];
        if (round < 14) {
            for (int c = 0; c < 4; ++c) {
                uint8_t a0 = t[4 * c], a1 = t[4 * c + 1], a2 = t[4 * c + 2], a3 = t[4 * c + 3];
                s[4 * c + 0] = gf_mul(a0, 2) ^ gf_mul(a1, 3) ^ a2 ^ a3;
                s[4 * c + 1] = a0 ^ gf_mul(a1, 2) ^ gf_mul(a2, 3) ^ a3;
                s[4 * c + 2] = a0 ^ a1 ^ gf_mul(a2, 2) ^ gf_mul(a3, 3);
                s[4 * c + 3] = gf_mul(a0, 3) ^ a1 ^ a2 ^ gf_mul(a3, 2);
            }
        } else {
            memcpy(s, t, 16);
        }
        for (int i = 0; i < 16; ++i) s[i] ^= rk[16 * round + i];
    }
    memcpy(out, s, 16);
}

/* Key reduction: the seed handed to the PRP may be longer than 32 bytes (hosts carry a 256-byte
 * left-zero-padded copy, jzf_flashe.py:285-292); jzf_aes.py:21-28 keeps the low 32 bytes
 * (int.from_bytes(k,'big') & (256**32-1)).  Shorter seeds are left-zero-padded by to_bytes. */
void fo_reduce_key(const uint8_t* seed, size_t n, uint8_t key[32]) {
    memset(key, 0, 32);
    if (n >= 32) memcpy(key, seed + (n - 32), 32);
    else memcpy(key + (32 - n), seed, n);
}

/* ------------------------------------------------------------------------------------------------
 * chunks_idx — jzf_flashe.py:12-16.  d,r = divmod(L, n); chunk i starts at
 * (d+1)*min(i,r) + d*max(i-r,0) and has d+1 elements when i < r, else d.
 * ---------------------------------------------------------------------------------------------- */
void fo_chunk_bounds(uint64_t L, uint32_t n_jobs, uint32_t i, uint64_t* begin, uint64_t* end) {
    uint64_t d = L / n_jobs, r = L % n_jobs;
    uint64_t si = (d + 1) * (i < r ? i : r) + d * (i < r ? 0 : i - r);
    *begin = si;
    *end = si + (i < r ? d + 1 : d);
}

static int word_bytes(int b) { return b <= 32 ? 4 : (b <= 64 ? 8 : 16); }
int fo_word_bytes(int int_bits) { return (int_bits < 1 || int_bits > 128) ? FO_EINVAL : word_bytes(int_bits); }

static u128 mask_of(int b) { return b >= 128 ? ~(u128)0 : (((u128)1 << b) - 1); }

static u128 load_word(const void* p, uint64_t j, int wb) {
    if (wb == 4) return ((const uint32_t*)p)[j];
    if (wb == 8) return ((const uint64_t*)p)[j];
    u128 v; memcpy(&v, (const uint8_t*)p + 16 * j, 16); return v;
}
static void store_word(void* p, uint64_t j, int wb, u128 v) {
    if (wb == 4) ((uint32_t*)p)[j] = (uint32_t)v;
    else if (wb == 8) ((uint64_t*)p)[j] = (uint64_t)v;
    else memcpy((uint8_t*)p + 16 * j, &v, 16);
}

/* One chunk of one keystream — the loop body of _static_prepare_encrypt_single,
 * jzf_flashe.py:19-45 (and the identical add/minus halves of :48-82):
 *   merge_size = 128 // int_bits; block i of chunk [begin,end) encrypts
 *   iter(4B BE) || prf_index(4B BE) || (i+begin)(8B BE)  (:34, :304, :308-309)
 *   s = int.from_bytes(AES(block),'big'); element b..e-1 take s & mask, s >>= int_bits (:37-43).
 * Writes F[j] for j in [begin,end) ∩ [j0, j0+cnt) to out[j-j0] as u128. */
static void stream_chunk(const uint8_t rk[240], int b, uint32_t iter, uint32_t prf, uint64_t begin,
                         uint64_t end, uint64_t j0, uint64_t cnt, u128* out) {
    uint64_t len = end - begin;
    if (len == 0) return;
    int m = 128 / b;
    uint64_t nblk = (len - 1) / m + 1;
    u128 msk = mask_of(b);
    uint64_t lo = j0 > begin ? j0 : begin, hi = (j0 + cnt) < end ? (j0 + cnt) : end;
    if (lo >= hi) return;
    uint64_t i_lo = (lo - begin) / m, i_hi = (hi - 1 - begin) / m;
    (void)nblk;
    for (uint64_t i = i_lo; i <= i_hi; ++i) {
        uint8_t in[16], o[16];
        uint64_t ctr = i + begin;
        in[0] = (uint8_t)(iter >> 24); in[1] = (uint8_t)(iter >> 16); in[2] = (uint8_t)(iter >> 8); in[3] = (uint8_t)iter;
        in[4] = (uint8_t)(prf >> 24); in[5] = (uint8_t)(prf >> 16); in[6] = (uint8_t)(prf >> 8); in[7] = (uint8_t)prf;
        for (int k = 0; k < 8; ++k) in[8 + k] = (uint8_t)(ctr >> (56 - 8 * k));
        fo_aes256_encrypt_block(rk, in, o);
        u128 s = 0;
        for (int k = 0; k < 16; ++k) s = (s << 8) | o[k];
        uint64_t eb = begin + i * m;
        uint64_t ee = eb + m < end ? eb + m : end;
        for (uint64_t j = eb; j < ee; ++j) {
            if (j >= lo && j < hi) out[j - j0] = s & msk;
            s = (b >= 128) ? 0 : (s >> b);
        }
    }
}

/* F(iter, prf)[j0 .. j0+cnt) of a length-L vector chunked over n_jobs workers, as u128 per element.
 * Pool fan-out: jzf_flashe.py:433-441 / 459-466. */
typedef struct { const uint8_t* rk; int b; uint32_t n_jobs, iter, prf; uint64_t L, j0; u128* out; } stream_ctx;
static void stream_part(int64_t lo, int64_t hi, void* p) {
    stream_ctx* c = (stream_ctx*)p;
    if (hi <= lo) return;
    for (uint32_t k = 0; k < c->n_jobs; ++k) {
        uint64_t cb, ce;
        fo_chunk_bounds(c->L, c->n_jobs, k, &cb, &ce);
        stream_chunk(c->rk, c->b, c->iter, c->prf, cb, ce, c->j0 + (uint64_t)lo, (uint64_t)(hi - lo), c->out + lo);
    }
}
static int stream_range(const uint8_t rk[240], int b, uint32_t n_jobs, uint32_t iter, uint32_t prf,
                        uint64_t L, uint64_t j0, uint64_t cnt, u128* out) {
    if (b < 1 || b > 128 || n_jobs == 0 || j0 + cnt > L) return FO_EINVAL;
    stream_ctx c = { rk, b, n_jobs, iter, prf, L, j0, out };
    parallel_for((int64_t)cnt, stream_part, &c);
    return FO_OK;
}

/* out[j-j0] = sum_k sign[k] * F(iter, prf_idx[k])[j]  mod 2^b   for j in [j0, j0+cnt).
 * This is the combined mask term of encrypt (jzf_flashe.py:480-481 add - minus), of decrypt
 * (:570-571, with the per-block multi-stream sums of :115-152) and of the precompute buffers
 * (:599-666). */
int fo_masks(const uint8_t key[32], int int_bits, uint32_t n_jobs, uint32_t iter,
             const int32_t* prf_idx, const int32_t* sign, int nstreams, uint64_t L, uint64_t j0,
             uint64_t cnt, void* out) {
    if (int_bits < 1 || int_bits > 128) return FO_EINVAL;
    uint8_t rk[240];
    fo_aes256_expand(key, rk);
    int wb = word_bytes(int_bits);
    u128 msk = mask_of(int_bits);
    u128* acc = (u128*)calloc(cnt ? cnt : 1, sizeof(u128));
    u128* tmp = (u128*)calloc(cnt ? cnt : 1, sizeof(u128));
    if (!acc || !tmp) { free(acc); free(tmp); return FO_EINVAL; }
    int rc = FO_OK;
    for (int k = 0; k < nstreams && rc == FO_OK; ++k) {
        rc = stream_range(rk, int_bits, n_jobs, iter, (uint32_t)prf_idx[k], L, j0, cnt, tmp);
        if (rc) break;
        if (sign[k] >= 0) for (uint64_t j = 0; j < cnt; ++j) acc[j] = (acc[j] + tmp[j]) & msk;
        else for (uint64_t j = 0; j < cnt; ++j) acc[j] = (acc[j] - tmp[j]) & msk;
    }
    if (rc == FO_OK) for (uint64_t j = 0; j < cnt; ++j) store_word(out, j, wb, acc[j]);
    free(acc); free(tmp);
    return rc;
}

/* out = (in + combined mask) mod 2^b over [j0, j0+cnt) — encrypt (jzf_flashe.py:431-488: double
 * uses streams (idx,+),(idx+1,-) per :349-353; single uses (idx,+) per :308-309) and decrypt
 * (:506-582: double adds F(t,a) for a in A and subtracts F(t,s) for s in S; single subtracts every
 * survivor's stream).  `in`/`out` hold cnt words and may alias. */
int fo_apply_masks(const uint8_t key[32], int int_bits, uint32_t n_jobs, uint32_t iter,
                   const int32_t* prf_idx, const int32_t* sign, int nstreams, uint64_t L,
                   uint64_t j0, uint64_t cnt, const void* in, void* out) {
    int wb = word_bytes(int_bits);
    void* mk = malloc((cnt ? cnt : 1) * (size_t)wb);
    if (!mk) return FO_EINVAL;
    int rc = fo_masks(key, int_bits, n_jobs, iter, prf_idx, sign, nstreams, L, j0, cnt, mk);
    if (rc == FO_OK) {
        u128 msk = mask_of(int_bits);
        for (uint64_t j = 0; j < cnt; ++j)
            store_word(out, j, wb, (load_word(in, j, wb) + load_word(mk, j, wb)) & msk);
    }
    free(mk);
    return rc;
}

/* set_idx_list(mode="decrypt") run collapse — jzf_flashe.py:354-367: sort the surviving client
 * indices; every maximal run [a..b] contributes minus-stream a and add-stream b+1.  Duplicates are
 * handled the way the reference's loop handles them (idx == temp_add[-1] only matches idx = prev+1,
 * so a repeated idx opens a new run).  Returns the number of runs; add/minus need room for n. */
static int cmp_i32(const void* a, const void* b) {
    int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
    return (x > y) - (x < y);
}
int fo_collapse_runs(const int32_t* survivors, int n, int32_t* add, int32_t* minus) {
    if (n <= 0) return 0;
    int32_t* s = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    memcpy(s, survivors, sizeof(int32_t) * (size_t)n);
    qsort(s, (size_t)n, sizeof(int32_t), cmp_i32);
    int runs = 0;
    for (int i = 0; i < n; ++i) {
        if (runs == 0 || s[i] != add[runs - 1]) { add[runs] = s[i] + 1; minus[runs] = s[i]; ++runs; }
        else add[runs - 1] = s[i] + 1;
    }
    free(s);
    return runs;
}

/* Server sum, element-wise — framework/homo/procedure/jzf_aggregator.py:421-430:
 * reduce(lambda x, y: (x + y) % (1 << int_bits)) over n decompressed object arrays.
 * cts holds n vectors of L words, vector c at word offset c*stride. */
typedef struct { int b, n; uint64_t stride; const void* cts; void* out; } agg_ctx;
static void agg_part(int64_t lo, int64_t hi, void* p) {
    agg_ctx* a = (agg_ctx*)p;
    int wb = word_bytes(a->b);
    u128 msk = mask_of(a->b);
    for (int64_t j = lo; j < hi; ++j) {
        u128 acc = 0;
        for (int c = 0; c < a->n; ++c) acc = (acc + load_word(a->cts, (uint64_t)c * a->stride + (uint64_t)j, wb)) & msk;
        store_word(a->out, (uint64_t)j, wb, acc);
    }
}
int fo_aggregate_elementwise(int int_bits, int n, uint64_t L, uint64_t stride, const void* cts, void* out) {
    if (int_bits < 1 || int_bits > 128 || n < 1) return FO_EINVAL;
    agg_ctx a = { int_bits, n, stride, cts, out };
    parallel_for((int64_t)L, agg_part, &a);
    return FO_OK;
}

/* Server sum on the packed wire integer — jzf_aggregator.py:404-419 with the packing of
 * framework/jzf_weights.py:45-84,155-195 (vector -> sum_j ct[j] << ((L-1-j)*b), first element most
 * significant).  (x + y) % (1 << (b*L)) is an L*b-bit addition, so carries out of field j leak into
 * field j-1 and the carry out of field 0 is dropped.  Evaluated as a radix-2^b addition from the
 * last element to the first; carry_in enters at the last element (0 for a whole vector; used by
 * the shard tests), *carry_out receives the carry out of element 0. */
int fo_aggregate_packed(int int_bits, int n, uint64_t L, uint64_t stride, const void* cts, void* out,
                        uint32_t carry_in, uint32_t* carry_out) {
    if (int_bits < 1 || int_bits > 128 || n < 1) return FO_EINVAL;
    int wb = word_bytes(int_bits);
    u128 msk = mask_of(int_bits);
    u128 carry = carry_in;
    for (uint64_t jj = L; jj-- > 0;) {
        /* digit sum kept as hi*2^b + lo with lo < 2^b; for b = 128 the overflow of the u128 addition
         * IS the unit of hi */
        u128 lo = 0, hi = 0;
        for (int c = -1; c < n; ++c) {
            u128 w = c < 0 ? carry : load_word(cts, (uint64_t)c * stride + jj, wb);
            if (int_bits == 128) { u128 s = lo + w; hi += (s < lo) ? 1 : 0; lo = s; }
            else {
                hi += w >> int_bits; w &= msk;       /* (the carry may exceed 2^b only for tiny b) */
                lo += w; hi += lo >> int_bits; lo &= msk;
            }
        }
        store_word(out, jj, wb, lo);
        carry = hi;
    }
    if (carry_out) *carry_out = (uint32_t)carry;
    return FO_OK;
}

/* Encode — _static_quantize_padding_asymmetric, jzf_quantize.py:55-67, float32 layer + Python float
 * alpha (the arithmetic types that reach it: nn/backend/tf_keras/jzf_nn_model.py:144-145 hands
 * float32 layers; under the pinned numpy 1.17.2 a float32 array op a Python scalar stays float32):
 *   v = clip(x, -a, a) + a          float32, a rounded to float32
 *   v = v * (2^e - 1) / (2a)        two float32 ops, (2a) computed in double then rounded
 *   q = floor(double(v) + u)        u = np.random.random() in [0,1), float64
 * Compile with -ffp-contract=off. */
typedef struct { const float* x; const double* u; uint32_t* q; double alpha; int ebits; } q_ctx;
static void q_part(int64_t lo, int64_t hi, void* p) {
    q_ctx* c = (q_ctx*)p;
    volatile float a = (float)c->alpha, na = (float)(-c->alpha);
    volatile float scale = (float)(((int64_t)1 << c->ebits) - 1);
    volatile float two_a = (float)(2.0 * c->alpha);
    for (int64_t j = lo; j < hi; ++j) {
        float v = c->x[j];
        v = v < na ? na : v;            /* np.clip == minimum(maximum(x, lo), hi) */
        v = v > a ? a : v;
        volatile float t = v + a;
        t = t * scale;
        t = t / two_a;
        double r = floor((double)t + c->u[j]);
        c->q[j] = (uint32_t)(int64_t)r;
    }
}
int fo_quantize(const float* x, const double* u, uint64_t L, double alpha, int element_bits, uint32_t* q) {
    if (element_bits < 1 || element_bits > 24) return FO_EINVAL;
    q_ctx c = { x, u, q, alpha, element_bits };
    parallel_for((int64_t)L, q_part, &c);
    return FO_OK;
}

/* Decode — _static_unquantize_padding_asymmetric, jzf_quantize.py:102-107, on Python ints:
 *   alpha *= n;  out = v * (2*alpha) / ((2^e - 1) * n) - alpha      (float64, left to right) */
typedef struct { const void* v; double* out; int wb; double alpha; int ebits, n; } uq_ctx;
static void uq_part(int64_t lo, int64_t hi, void* p) {
    uq_ctx* c = (uq_ctx*)p;
    volatile double an = c->alpha * (double)c->n;
    volatile double two_an = 2.0 * an;
    volatile double den = (double)((((int64_t)1 << c->ebits) - 1) * (int64_t)c->n);
    for (int64_t j = lo; j < hi; ++j) {
        double val = c->wb == 4 ? (double)((const uint32_t*)c->v)[j] : (double)((const uint64_t*)c->v)[j];
        volatile double t = val * two_an;
        t = t / den;
        c->out[j] = t - an;
    }
}
int fo_unquantize(const void* v, uint64_t L, int word_bytes_, double alpha, int element_bits, int n, double* out) {
    if (word_bytes_ != 4 && word_bytes_ != 8) return FO_EINVAL;
    uq_ctx c = { v, out, word_bytes_, alpha, element_bits, n };
    parallel_for((int64_t)L, uq_part, &c);
    return FO_OK;
}

/* Lane batching — _static_batching_padding_asymmetric, jzf_quantize.py:162-185: lane width
 * element_bits + factor, batch_size = int_bits // lane, zero-pad to a multiple of batch_size, pack
 * with the FIRST element most significant (temp = temp*mod + x).  Output: ceil(L/batch_size) words
 * of 16 bytes (any int_bits; callers use 120). */
int fo_batch(const uint32_t* q, uint64_t L, int int_bits, int element_bits, int factor, void* out) {
    int lane = element_bits + factor;
    if (lane < 1 || lane > 32 || int_bits < lane || int_bits > 128) return FO_EINVAL;
    int bs = int_bits / lane;
    uint64_t nw = (L + bs - 1) / bs;
    int wb = word_bytes(int_bits);
    for (uint64_t w = 0; w < nw; ++w) {
        u128 t = 0;
        for (int i = 0; i < bs; ++i) {
            uint64_t j = w * bs + i;
            t = (t << lane) + (j < L ? q[j] : 0);
        }
        store_word(out, w, wb, t);
    }
    return FO_OK;
}

/* _static_unbatching_padding_asymmetric, jzf_quantize.py:234-251: peel batch_size lanes from the
 * least significant end, reverse; returns nw*batch_size values (caller truncates to L, :516). */
int fo_unbatch(const void* in, uint64_t nw, int int_bits, int element_bits, int factor, uint32_t* out) {
    int lane = element_bits + factor;
    if (lane < 1 || lane > 32 || int_bits < lane || int_bits > 128) return FO_EINVAL;
    int bs = int_bits / lane;
    int wb = word_bytes(int_bits);
    u128 lm = mask_of(lane);
    for (uint64_t w = 0; w < nw; ++w) {
        u128 t = load_word(in, w, wb);
        for (int i = bs - 1; i >= 0; --i) { out[w * bs + i] = (uint32_t)(t & lm); t >>= lane; }
    }
    return FO_OK;
}

/* expand_to_dense — jzf_aggregator.py:150-165: scatter a compact ciphertext to its sorted indices
 * in a `total`-long vector and fill every other position with that client's quantised zero. */
int fo_expand_to_dense(int int_bits, const void* compact, const int64_t* index, uint64_t k,
                       uint64_t total, const void* zero_word, void* dense) {
    int wb = word_bytes(int_bits);
    u128 z = load_word(zero_word, 0, wb);
    for (uint64_t j = 0; j < total; ++j) store_word(dense, j, wb, z);
    for (uint64_t i = 0; i < k; ++i) {
        if (index[i] < 0 || (uint64_t)index[i] >= total) return FO_EINVAL;
        store_word(dense, (uint64_t)index[i], wb, load_word(compact, i, wb));
    }
    return FO_OK;
}

/* Sparse single-mask decrypt term — jzf_flashe.py:315-343: for every client c regenerate F(t,c)
 * over its len(mask_c) COMPACT positions (chunked over the compact length), scatter to dense, sum
 * mod 2^b.  Adds (sign>=0) or subtracts client c's scattered stream into acc[total]. */
int fo_sparse_stream_accumulate(const uint8_t key[32], int int_bits, uint32_t n_jobs, uint32_t iter,
                                int32_t prf, int32_t sign, const int64_t* index, uint64_t k,
                                uint64_t total, void* acc) {
    int wb = word_bytes(int_bits);
    void* mk = malloc((k ? k : 1) * (size_t)wb);
    if (!mk) return FO_EINVAL;
    int32_t one = 1;
    int rc = fo_masks(key, int_bits, n_jobs, iter, &prf, &one, 1, k, 0, k, mk);
    u128 msk = mask_of(int_bits);
    for (uint64_t i = 0; rc == FO_OK && i < k; ++i) {
        if (index[i] < 0 || (uint64_t)index[i] >= total) { rc = FO_EINVAL; break; }
        u128 a = load_word(acc, (uint64_t)index[i], wb), f = load_word(mk, i, wb);
        store_word(acc, (uint64_t)index[i], wb, (sign >= 0 ? a + f : a - f) & msk);
    }
    free(mk);
    return rc;
}

/* dynamic_masking cost model — framework/homo/procedure/jzf_flashe_block.py:89-117:
 * single = 2*sum|mask_i|; double = 2*single - 2*sum_i |mask_i ∩ mask_{i+1}|; choose single when
 * single <= double.  Index lists must be sorted ascending.  Returns 0 for "single", 1 for "double". */
int fo_dynamic_masking(const int64_t* const* index, const uint64_t* k, int n, uint64_t* single_cost,
                       uint64_t* double_cost) {
    uint64_t s = 0, ov = 0;
    for (int i = 0; i < n; ++i) s += k[i];
    for (int i = 0; i + 1 < n; ++i) {
        uint64_t a = 0, b = 0;
        while (a < k[i] && b < k[i + 1]) {
            if (index[i][a] == index[i + 1][b]) { ++ov; ++a; ++b; }
            else if (index[i][a] < index[i + 1][b]) ++a; else ++b;
        }
    }
    uint64_t sc = 2 * s, dc = 2 * sc - 2 * ov;
    if (single_cost) *single_cost = sc;
    if (double_cost) *double_cost = dc;
    return sc <= dc ? 0 : 1;
}

