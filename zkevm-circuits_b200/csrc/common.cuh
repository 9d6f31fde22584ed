// common.cuh -- context, error plumbing and launch accounting shared by all translation units of libzkb200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <map>
#include <string>
#include <vector>
#include <array>
#include "../../include/zkb200.h"
#include "ff.cuh"

namespace zkb {

void set_error(const char *fmt, ...);

#define ZKB_CUDA(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            zkb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));         \
            return _e == cudaErrorMemoryAllocation ? ZKB_ERR_ALLOC : ZKB_ERR_CUDA;                       \
        }                                                                                                \
    } while (0)

#define ZKB_TRY(expr)                    \
    do {                                 \
        int32_t _r = (expr);             \
        if (_r != ZKB_OK) return _r;     \
    } while (0)

#define ZKB_ARG(cond)                                                            \
    do {                                                                         \
        if (!(cond)) {                                                           \
            zkb::set_error("%s:%d invalid argument: %s", __FILE__, __LINE__, #cond); \
            return ZKB_ERR_ARG;                                                  \
        }                                                                        \
    } while (0)

// One cached NTT plan per (log_n, omega): twiddle tables on the device.
struct NttPlan {
    uint32_t log_n = 0;
    int npass = 0;
    int bits[3] = {0, 0, 0};
    Fr *tw_lo = nullptr;   // omega^i,           i < 2^min(log_n, 12)
    Fr *tw_hi = nullptr;   // omega^(i * 2^12),  i < 2^(log_n - 12)   (log_n > 12)
    Fr *loc[3] = {nullptr, nullptr, nullptr};  // per pass: (omega_{2^a})^i, i < 2^(a-1)
};

struct DeviceBuffer {
    void *ptr = nullptr;
    size_t bytes = 0;
};

}  // namespace zkb

struct zkb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    uint64_t launches = 0;
    uint64_t msm_last_adds = 0;
    std::map<std::array<uint64_t, 5>, zkb::NttPlan> ntt_plans;
    // grow-only scratch arenas (device), keyed by purpose; avoids cudaMalloc in steady state
    zkb::DeviceBuffer scratch[8];
    void *pinned = nullptr;  // small pinned staging buffer
    size_t pinned_bytes = 0;
};

namespace zkb {
// returns a device scratch buffer of at least `bytes` (slot-indexed, grow-only)
int32_t scratch_get(zkb_ctx *ctx, int slot, size_t bytes, void **out);
inline cudaStream_t pick_stream(zkb_ctx *ctx, void *stream) { return stream ? (cudaStream_t)stream : ctx->stream; }

enum ScratchSlot { SCR_NTT = 0, SCR_MSM_A = 1, SCR_MSM_B = 2, SCR_MSM_C = 3, SCR_HOSTIO_A = 4, SCR_HOSTIO_B = 5, SCR_MISC = 6, SCR_MISC2 = 7 };
}  // namespace zkb
