// flashe_internal.h — host-side glue shared by the translation units of libflashe_b200.so
// (not part of the C ABI; include/flashe_b200.h is).
#ifndef FLASHE_INTERNAL_H
#define FLASHE_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/flashe_b200.h"

// Records `msg` as the calling thread's last error and returns `code`.
int flashe_fail(int code, const std::string& msg);
// Bench bookkeeping behind flashe_launch_count().
void flashe_count_launches(int n);
// What the other translation units need to know about a context.
struct flashe_ctx_info { int device; int int_bits; int words; int num_sms; };
int flashe_ctx_get_info(const flashe_ctx* ctx, flashe_ctx_info* out);

#define FLASHE_CUDA_TRY(expr)                                                                            \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return flashe_fail(FLASHE_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));       \
    } while (0)

struct FlasheDeviceGuard {
    int prev;
    bool ok;
    explicit FlasheDeviceGuard(int dev) : prev(-1), ok(true) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~FlasheDeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

#endif
