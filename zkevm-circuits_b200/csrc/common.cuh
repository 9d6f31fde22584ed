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
#include "g1.cuh"

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
    // two-pass plans up to 2^25: the complete inter-pass twiddle table T[j_in * A + k] = omega^(j_in k) (x scale), n entries,
    // so the boundary costs ONE multiply per element instead of two (lo x hi combine + apply)
    Fr *tw_full = nullptr;
    Fr *tw_full_scaled = nullptr;
    Fr scaled_key;
    bool has_scaled = false;
};

struct DeviceBuffer {
    void *ptr = nullptr;
    size_t bytes = 0;
};

}  // namespace zkb

struct zkb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // H2D of witness columns overlaps the MSMs of the previous batch
    int sm_count = 148;
    // multi-GPU: NCCL communicator of this rank (comm.cu); nranks == 1 -> everything local
    void *nccl_comm = nullptr;
    int rank = 0, nranks = 1;
    uint64_t launches = 0;
    uint64_t msm_last_adds = 0;
    std::map<std::array<uint64_t, 5>, zkb::NttPlan> ntt_plans;
    // grow-only scratch arenas (device), keyed by purpose; avoids cudaMalloc in steady state
    zkb::DeviceBuffer scratch[12];
    // cached device blocks (size -> pointers) recycled between proving sessions: cudaMalloc/cudaFree of tens of GB per proof
    // costs seconds and is wildly variable; blocks go back to the driver only at zkb_destroy
    std::multimap<size_t, void *> block_cache;
    size_t block_cache_bytes = 0;
    void *pinned = nullptr;  // small pinned staging buffer
    size_t pinned_bytes = 0;
};

namespace zkb {
// returns a device scratch buffer of at least `bytes` (slot-indexed, grow-only)
int32_t scratch_get(zkb_ctx *ctx, int slot, size_t bytes, void **out);
// cached block allocator (see zkb_ctx::block_cache)
int32_t block_alloc(zkb_ctx *ctx, size_t bytes, void **out, size_t *got);
void block_free(zkb_ctx *ctx, void *p, size_t bytes);
inline cudaStream_t pick_stream(zkb_ctx *ctx, void *stream) { return stream ? (cudaStream_t)stream : ctx->stream; }

// ---- cross-translation-unit device-side services (all launch on `st`, none synchronises unless stated) ----------------
Fr host_root_of_unity(uint32_t k);
Fr host_zeta();
// dst <- NTT_omega(src * in_scale) * scale ; src == dst allowed; coset_zeta as in zkb_ntt_fr_dev
int32_t ntt_fr_device(zkb_ctx *ctx, const Fr *src, Fr *dst, uint32_t log_n, const Fr &omega, const Fr *scale_host, int coset_zeta,
                      const Fr *d_in_scale, cudaStream_t st);
// `count` transforms in one launch per pass: column y reads d_src_tbl[y], writes d_dst_tbl[y] (device pointer tables)
int32_t ntt_fr_batch_device(zkb_ctx *ctx, const Fr *src, Fr *dst, const Fr *const *d_src_tbl, Fr *const *d_dst_tbl, uint32_t count, uint32_t log_n,
                            const Fr &omega, const Fr *scale_host, int coset_zeta, const Fr *d_in_scale, cudaStream_t st);
// synchronises: the result point is returned to the host
int32_t msm_g1_device(zkb_ctx *ctx, const Fr *scalars, const G1Affine *bases, uint64_t n, G1Affine *out_affine_host, cudaStream_t st);
// `batch` MSMs over the same bases in one pass (d_scalar_cols: DEVICE array of device pointers; batch <= msm_max_batch(n))
uint32_t msm_max_batch(uint64_t n);
// window-shifted precomputed bases (copy w = 2^(c w) P_i): one bucket set per column, no Horner
uint32_t msm_shift_copies(uint64_t n);
int32_t msm_build_shifted_bases(zkb_ctx *ctx, const G1Affine *bases, uint64_t n, G1Affine *out, cudaStream_t st);
int32_t msm_g1_batch_device_ex(zkb_ctx *ctx, const Fr *const *d_scalar_cols, uint32_t batch, const G1Affine *bases, uint64_t n,
                               G1Affine *out_affine_host, bool shifted, cudaStream_t st);
int32_t msm_g1_batch_device(zkb_ctx *ctx, const Fr *const *d_scalar_cols, uint32_t batch, const G1Affine *bases, uint64_t n,
                            G1Affine *out_affine_host, cudaStream_t st);
int32_t fr_powers_device(zkb_ctx *ctx, const Fr &base, uint64_t n, Fr *out, cudaStream_t st);
int32_t poly_eval_device(zkb_ctx *ctx, const Fr *const *d_polys, uint32_t num, uint64_t n, const Fr &x, Fr *out_host, cudaStream_t st);
int32_t prefix_product_device(zkb_ctx *ctx, const Fr *in, uint64_t n, const Fr &init, Fr *out, cudaStream_t st);
int32_t prefix_sum_device(zkb_ctx *ctx, const Fr *in, uint64_t n, const Fr &init, Fr *out, cudaStream_t st);
int32_t kate_division_device(zkb_ctx *ctx, const Fr *a, uint64_t n, const Fr &u, Fr *q, cudaStream_t st);
int32_t lincomb_device(zkb_ctx *ctx, const Fr *const *d_polys, const Fr *d_coefs, uint32_t num, uint64_t n, Fr *out, bool accumulate, cudaStream_t st);
int32_t batch_invert_device(zkb_ctx *ctx, const Fr *a, Fr *out, uint64_t n, cudaStream_t st);
// in-place u64 sum across the context's ranks (exact gather when the supports are disjoint); no-op for a single rank
int32_t comm_allreduce_u64(zkb_ctx *ctx, void *dev_buf, size_t count, cudaStream_t st);

enum ScratchSlot { SCR_NTT = 0, SCR_MSM_A = 1, SCR_MSM_B = 2, SCR_MSM_C = 3, SCR_HOSTIO_A = 4, SCR_HOSTIO_B = 5, SCR_MISC = 6, SCR_MISC2 = 7, SCR_MSM_TBL = 8, SCR_COMM = 9 };
}  // namespace zkb
