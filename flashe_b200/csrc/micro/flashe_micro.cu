// flashe_micro.cu — micro-benchmarks behind the PRF design decision (NOT part of libflashe_b200.so; built into
// flashe_b200/_lib/libflashe_micro.so by `python -m flashe_b200.build --micro`, driven by scripts/microbench.py).
//
//   k_lds_peak       how many conflict-free 4-byte shared-memory lookups per clock an SM really serves: the
//                    ceiling of ANY table-driven AES (one lookup per state byte and round), measured instead of
//                    quoted.  A 2-way-conflict variant checks that the counter behaves as modelled.
//   k_aes_bitslice   a bit-sliced AES-256 (no tables at all: every S-box is ~146 logic gates on 32-bit bit planes,
//                    32 blocks per thread), the alternative the T-table kernel was chosen over.  It is ALU-pipe
//                    bound by construction; its measured block rate (and the rate scaled to the best published
//                    113-gate S-box) is what the shared-memory kernel has to beat.
//
// The S-box circuit is generated and exhaustively verified by scripts/gen_bitslice_sbox.py.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "sbox_bitslice.inc"

#define MICRO_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { snprintf(g_err, sizeof(g_err), "%s: %s", #expr, cudaGetErrorString(e__)); return -1; } } while (0)
static char g_err[256];

// ------------------------------------------------------------------------------------------------ LDS peak
// 512 threads per CTA, one CTA per SM, like k_stream.  Every lane reads its own 4-byte column of a 32-column
// table (bank = lane: conflict-free) at an entry that changes every step; CONFLICT = 2 makes lane pairs share a
// bank on different rows (2 wavefronts per instruction).  16 independent loads per step keep the pipe full.
template <int CONFLICT>
__global__ void __launch_bounds__(512, 1) k_lds_peak(int steps, uint32_t* out) {
    extern __shared__ uint32_t tab[];                     // 512 rows x 32 columns x 4 B = 64 KB
    for (uint32_t i = threadIdx.x; i < 16384u; i += blockDim.x) tab[i] = i * 2654435761u;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t col = CONFLICT == 1 ? lane : (lane & ~1u);                 // 2-way: lanes 2k and 2k+1 share bank 2k ...
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab) + col * 4u + (CONFLICT == 1 ? 0u : (lane & 1u) * 128u);   // ... on different rows
    uint32_t off = (threadIdx.x * 7u & 255u) * 128u;       // row offset in [0, 32 KB); the 16 loads add k * 2 KB as immediates
    uint32_t acc = 0;
    for (int s = 0; s < steps; ++s) {
        const uint32_t a = base + off;
        uint32_t v[16];
#define LDS_K(k) asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v[k]) : "r"(a), "n"((k) * 2048))
        LDS_K(0); LDS_K(1); LDS_K(2); LDS_K(3); LDS_K(4); LDS_K(5); LDS_K(6); LDS_K(7);
        LDS_K(8); LDS_K(9); LDS_K(10); LDS_K(11); LDS_K(12); LDS_K(13); LDS_K(14); LDS_K(15);
#undef LDS_K
#pragma unroll
        for (int k = 0; k < 16; ++k) acc ^= v[k];
        off = (off + 640u) & 0x7f80u;
    }
    if (acc == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;   // keeps the loads alive
}

// ------------------------------------------------------------------------------------------------ bit-sliced AES-256
__constant__ uint32_t c_rkmask[15 * 128];   // round key r, bit plane p (byte i, bit b: p = 8 i + b): 0 or 0xffffffff

__device__ __forceinline__ void bs_sub_bytes(uint32_t (&s)[128]) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
        sbox_bitslice(s[8 * i], s[8 * i + 1], s[8 * i + 2], s[8 * i + 3], s[8 * i + 4], s[8 * i + 5], s[8 * i + 6], s[8 * i + 7]);
}
// byte i = 4*col + row; ShiftRows: new(row, col) = old(row, (col + row) mod 4)
__device__ __forceinline__ void bs_shift_rows(uint32_t (&s)[128]) {
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        uint32_t t;
        t = s[8 * 1 + b]; s[8 * 1 + b] = s[8 * 5 + b]; s[8 * 5 + b] = s[8 * 9 + b]; s[8 * 9 + b] = s[8 * 13 + b]; s[8 * 13 + b] = t;   // row 1
        t = s[8 * 2 + b]; s[8 * 2 + b] = s[8 * 10 + b]; s[8 * 10 + b] = t;                                                           // row 2
        t = s[8 * 6 + b]; s[8 * 6 + b] = s[8 * 14 + b]; s[8 * 14 + b] = t;
        t = s[8 * 15 + b]; s[8 * 15 + b] = s[8 * 11 + b]; s[8 * 11 + b] = s[8 * 7 + b]; s[8 * 7 + b] = s[8 * 3 + b]; s[8 * 3 + b] = t; // row 3
    }
}
__device__ __forceinline__ void bs_xtime(const uint32_t (&a)[8], uint32_t (&x)[8]) {
    x[0] = a[7]; x[1] = a[0] ^ a[7]; x[2] = a[1]; x[3] = a[2] ^ a[7]; x[4] = a[3] ^ a[7]; x[5] = a[4]; x[6] = a[5]; x[7] = a[6];
}
__device__ __forceinline__ void bs_mix_columns(uint32_t (&s)[128]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t a[4][8], t[4][8], x[4][8], u[8];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int b = 0; b < 8; ++b) a[r][b] = s[8 * (4 * c + r) + b];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int b = 0; b < 8; ++b) t[r][b] = a[r][b] ^ a[(r + 1) & 3][b];
#pragma unroll
        for (int b = 0; b < 8; ++b) u[b] = t[0][b] ^ t[2][b];
#pragma unroll
        for (int r = 0; r < 4; ++r) bs_xtime(t[r], x[r]);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int b = 0; b < 8; ++b) s[8 * (4 * c + r) + b] = x[r][b] ^ u[b] ^ a[r][b];     // 2 a_r + 3 a_{r+1} + a_{r+2} + a_{r+3}
    }
}
__device__ __forceinline__ void bs_add_round_key(uint32_t (&s)[128], int r) {
#pragma unroll
    for (int p = 0; p < 128; ++p) s[p] ^= c_rkmask[r * 128 + p];
}

// One thread = the 32 blocks iter || prf || (ctr0 + 32 t + k), k = 0..31 (big-endian fields, as the FLASHE PRF input).
__global__ void __launch_bounds__(128) k_aes_bitslice(uint32_t iter, uint32_t prf, uint64_t ctr0, uint64_t nthreads, int dump, uint32_t* out) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nthreads) return;
    const uint64_t c = ctr0 + 32ull * t;               // multiple of 32: the low five counter bits are the block number
    uint32_t s[128];
    uint8_t in[16];
    in[0] = iter >> 24; in[1] = iter >> 16; in[2] = iter >> 8; in[3] = iter;
    in[4] = prf >> 24; in[5] = prf >> 16; in[6] = prf >> 8; in[7] = prf;
#pragma unroll
    for (int i = 0; i < 8; ++i) in[8 + i] = (uint8_t)(c >> (8 * (7 - i)));
#pragma unroll
    for (int i = 0; i < 16; ++i)
#pragma unroll
        for (int b = 0; b < 8; ++b) s[8 * i + b] = 0u - (uint32_t)((in[i] >> b) & 1u);
    s[8 * 15 + 0] = 0xAAAAAAAAu; s[8 * 15 + 1] = 0xCCCCCCCCu; s[8 * 15 + 2] = 0xF0F0F0F0u; s[8 * 15 + 3] = 0xFF00FF00u; s[8 * 15 + 4] = 0xFFFF0000u;
    bs_add_round_key(s, 0);
#pragma unroll 1
    for (int r = 1; r < 14; ++r) {
        bs_sub_bytes(s);
        bs_shift_rows(s);
        bs_mix_columns(s);
        bs_add_round_key(s, r);
    }
    bs_sub_bytes(s);
    bs_shift_rows(s);
    bs_add_round_key(s, 14);
    if (dump) {
#pragma unroll
        for (int p = 0; p < 128; ++p) out[t * 128 + p] = s[p];
    } else {
        uint32_t x = 0;
#pragma unroll
        for (int p = 0; p < 128; ++p) x ^= s[p] + (uint32_t)p;
        out[t] = x;                                    // one word per 32 blocks: a real kernel would still have to
    }                                                  // transpose the planes back (~60 more instructions per block)
}

// ------------------------------------------------------------------------------------------------ host
static uint8_t h_sbox[256];
static void init_sbox() {
    uint8_t p = 1, q = 1;
    do {
        p = p ^ (uint8_t)(p << 1) ^ ((p & 0x80) ? 0x1b : 0);
        q ^= q << 1; q ^= q << 2; q ^= q << 4; if (q & 0x80) q ^= 0x09;
        uint8_t x = q ^ (uint8_t)((q << 1) | (q >> 7)) ^ (uint8_t)((q << 2) | (q >> 6)) ^ (uint8_t)((q << 3) | (q >> 5)) ^ (uint8_t)((q << 4) | (q >> 4));
        h_sbox[p] = x ^ 0x63;
    } while (p != 1);
    h_sbox[0] = 0x63;
}
static void expand256(const uint8_t key[32], uint8_t rk[240]) {
    init_sbox();
    memcpy(rk, key, 32);
    uint8_t rcon = 1;
    for (int i = 32; i < 240; i += 4) {
        uint8_t t[4] = {rk[i - 4], rk[i - 3], rk[i - 2], rk[i - 1]};
        if (i % 32 == 0) {
            const uint8_t t0 = t[0];
            t[0] = h_sbox[t[1]] ^ rcon; t[1] = h_sbox[t[2]]; t[2] = h_sbox[t[3]]; t[3] = h_sbox[t0];
            rcon = (uint8_t)((rcon << 1) ^ ((rcon & 0x80) ? 0x1b : 0));
        } else if (i % 32 == 16) {
            for (int k = 0; k < 4; ++k) t[k] = h_sbox[t[k]];
        }
        for (int k = 0; k < 4; ++k) rk[i + k] = rk[i - 32 + k] ^ t[k];
    }
}

extern "C" {

const char* fm_last_error(void) { return g_err; }

// conflict: 1 (conflict-free) or 2.  Returns the average milliseconds of one launch; lookups per launch =
// num_sms * 512 * steps * 16.
int fm_lds_peak(int device, int steps, int conflict, int reps, double* ms_out, int* num_sms_out) {
    MICRO_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    MICRO_TRY(cudaGetDeviceProperties(&prop, device));
    uint32_t* d = nullptr;
    MICRO_TRY(cudaMalloc((void**)&d, sizeof(uint32_t) * 512 * (size_t)prop.multiProcessorCount));
    MICRO_TRY(cudaFuncSetAttribute(k_lds_peak<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    MICRO_TRY(cudaFuncSetAttribute(k_lds_peak<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    cudaEvent_t a, b;
    MICRO_TRY(cudaEventCreate(&a)); MICRO_TRY(cudaEventCreate(&b));
    for (int w = 0; w < 2; ++w) {
        if (conflict == 1) k_lds_peak<1><<<prop.multiProcessorCount, 512, 65536>>>(steps, d);
        else k_lds_peak<2><<<prop.multiProcessorCount, 512, 65536>>>(steps, d);
    }
    MICRO_TRY(cudaEventRecord(a));
    for (int r = 0; r < reps; ++r) {
        if (conflict == 1) k_lds_peak<1><<<prop.multiProcessorCount, 512, 65536>>>(steps, d);
        else k_lds_peak<2><<<prop.multiProcessorCount, 512, 65536>>>(steps, d);
    }
    MICRO_TRY(cudaEventRecord(b));
    MICRO_TRY(cudaEventSynchronize(b));
    float ms = 0;
    MICRO_TRY(cudaEventElapsedTime(&ms, a, b));
    MICRO_TRY(cudaGetLastError());
    *ms_out = ms / reps; *num_sms_out = prop.multiProcessorCount;
    cudaFree(d); cudaEventDestroy(a); cudaEventDestroy(b);
    return 0;
}

// Bit-sliced AES-256 of nblocks (multiple of 32) consecutive counters from ctr0 (multiple of 32).
// dump = 1: planes_out (HOST, nblocks/32 * 128 words) receives the raw bit planes for verification;
// dump = 0: timing only, `reps` launches, *ms_out = average milliseconds per launch.
int fm_aes_bitslice(int device, const uint8_t key[32], uint32_t iter, uint32_t prf, uint64_t ctr0, uint64_t nblocks, int dump,
                    uint32_t* planes_out, int reps, double* ms_out) {
    MICRO_TRY(cudaSetDevice(device));
    if (nblocks % 32 || ctr0 % 32) { snprintf(g_err, sizeof(g_err), "nblocks and ctr0 must be multiples of 32"); return -1; }
    uint8_t rk[240];
    expand256(key, rk);
    static uint32_t masks[15 * 128];
    for (int r = 0; r < 15; ++r)
        for (int i = 0; i < 16; ++i)
            for (int b = 0; b < 8; ++b) masks[r * 128 + 8 * i + b] = ((rk[16 * r + i] >> b) & 1) ? 0xffffffffu : 0u;
    MICRO_TRY(cudaMemcpyToSymbol(c_rkmask, masks, sizeof(masks)));
    const uint64_t nthreads = nblocks / 32;
    uint32_t* d = nullptr;
    MICRO_TRY(cudaMalloc((void**)&d, sizeof(uint32_t) * (size_t)(dump ? nthreads * 128 : nthreads)));
    const unsigned grid = (unsigned)((nthreads + 127) / 128);
    cudaEvent_t a, b;
    MICRO_TRY(cudaEventCreate(&a)); MICRO_TRY(cudaEventCreate(&b));
    k_aes_bitslice<<<grid, 128>>>(iter, prf, ctr0, nthreads, dump, d);
    MICRO_TRY(cudaEventRecord(a));
    for (int r = 0; r < reps; ++r) k_aes_bitslice<<<grid, 128>>>(iter, prf, ctr0, nthreads, dump, d);
    MICRO_TRY(cudaEventRecord(b));
    MICRO_TRY(cudaEventSynchronize(b));
    float ms = 0;
    MICRO_TRY(cudaEventElapsedTime(&ms, a, b));
    MICRO_TRY(cudaGetLastError());
    if (ms_out) *ms_out = reps ? ms / reps : 0.0;
    if (dump && planes_out) MICRO_TRY(cudaMemcpy(planes_out, d, sizeof(uint32_t) * (size_t)nthreads * 128, cudaMemcpyDeviceToHost));
    cudaFree(d); cudaEventDestroy(a); cudaEventDestroy(b);
    return 0;
}

}  // extern "C"
