// flashe_stream_masks.cu — instantiations of k_stream for one mode (see flashe_stream.cuh).
#include "flashe_stream.cuh"

// one AES block, known-answer tests (flashe_prp_block)
__global__ void k_prp_block(const __grid_constant__ KeySched ks, const uint32_t* __restrict__ te0, const uint32_t* in, uint32_t* out) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t y = 0x00010000u | (lane << 2);
    if (smem_window_base() > TAB_BASE) { __trap(); }
    fill_tables(te0);
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


int flashe_launch_prp_block(const flashe_ctx* ctx, const uint32_t* d_in, uint32_t* d_out, cudaStream_t stream) {
    auto kern = k_prp_block;
    FLASHE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    kern<<<1, 256, SMEM_BYTES, stream>>>(ctx->ks, ctx->d_te0, d_in, d_out);
    flashe_count_launches(1);
    FLASHE_CUDA_TRY(cudaGetLastError());
    return FLASHE_OK;
}

int flashe_launch_stream_masks(const flashe_ctx* ctx, const StreamTab& st, const Geom& g, const IoDev& io, const CodecDev& cd,
        const NoiseDev& nz, cudaStream_t stream) {
    return launch_stream<M_MASKS>(ctx, st, g, io, cd, nz, stream);
}
