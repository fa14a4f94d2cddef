// flashe_stream_decode.cu — instantiations of k_stream for one mode (see flashe_stream.cuh).
#include "flashe_stream.cuh"

int flashe_launch_stream_decode(const flashe_ctx* ctx, const StreamTab& st, const Geom& g, const IoDev& io, const CodecDev& cd,
        const NoiseDev& nz, cudaStream_t stream) {
    return launch_stream<M_DECODE>(ctx, st, g, io, cd, nz, stream);
}
