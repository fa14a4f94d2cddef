// flashe_stream_encode_n32.cu — k_stream<M_ENCODE> with the 32-bit-resolution noise generator (flashe_noise.resolution =
// FLASHE_NOISE_32), plain and shared-stream schedules (see flashe_stream.cuh).
#include "flashe_stream.cuh"

int flashe_launch_stream_encode_n32(const flashe_ctx* ctx, const StreamTab& st, const Geom& g, const IoDev& io, const CodecDev& cd,
        const NoiseDev& nz, cudaStream_t stream) {
    return launch_stream<M_ENCODE, false, true>(ctx, st, g, io, cd, nz, stream);
}
int flashe_launch_stream_encode_shared_n32(const flashe_ctx* ctx, const StreamTab& st, const Geom& g, const IoDev& io, const CodecDev& cd,
        const NoiseDev& nz, cudaStream_t stream) {
    return launch_stream<M_ENCODE, true, true>(ctx, st, g, io, cd, nz, stream);
}
