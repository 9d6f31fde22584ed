// fieldops.cu -- element-wise Fr / Fq kernels and Montgomery batch inversion.
// Backs Polynomial +,-,* (halo2_proofs src/poly.rs operator impls) and halo2's batch inversion
// (`ff::BatchInvert` used by permutation/lookup provers and batch_invert_assigned, src/poly.rs / plonk/prover.rs).
#include "common.cuh"

namespace zkb {

template <class PR>
__global__ void binop_kernel(int op, const Fp<PR> *__restrict__ a, const Fp<PR> *__restrict__ b, Fp<PR> *__restrict__ out, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        Fp<PR> x = fp_load(a + i), y = fp_load(b + i), r;
        if (op == 0) r = fp_add(x, y);
        else if (op == 1) r = fp_sub(x, y);
        else r = fp_mul(x, y);
        fp_store(out + i, r);
    }
}

template <class PR>
__global__ void unop_kernel(int op, const Fp<PR> *__restrict__ a, Fp<PR> *__restrict__ out, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        Fp<PR> x = fp_load(a + i), r;
        switch (op) {
        case 0: r = fp_inv(x); break;
        case 1: r = fp_from_canonical(x); break;
        case 2: r = fp_to_canonical(x); break;
        case 3: r = fp_sqr(x); break;
        default: r = fp_neg(x); break;
        }
        fp_store(out + i, r);
    }
}

// Batch inversion: each thread owns CHUNK consecutive elements: prefix products (skipping zeros), one Fermat
// inversion per thread, then the backward sweep.  3 multiplies per element + 1 inversion (~380 mul) per CHUNK.
constexpr int BI_CHUNK = 32;
__global__ void batch_invert_kernel(const Fr *__restrict__ a, Fr *__restrict__ out, uint64_t n) {
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t start = t * BI_CHUNK;
    if (start >= n) return;
    const int len = (int)((n - start) < (uint64_t)BI_CHUNK ? (n - start) : BI_CHUNK);
    Fr acc = Fr::one();
    // forward: out[i] = product of non-zero a[start..i)
    for (int i = 0; i < len; ++i) {
        fp_store(out + start + i, acc);
        Fr v = fp_load(a + start + i);
        if (!v.is_zero()) acc = fp_mul(acc, v);
    }
    acc = fp_inv(acc);
    for (int i = len - 1; i >= 0; --i) {
        Fr v = fp_load(a + start + i);
        if (v.is_zero()) { fp_store(out + start + i, Fr::zero()); continue; }
        Fr pre = fp_load(out + start + i);
        fp_store(out + start + i, fp_mul(acc, pre));
        acc = fp_mul(acc, v);
    }
}

int32_t batch_invert_device(zkb_ctx *ctx, const Fr *a, Fr *out, uint64_t n, cudaStream_t st) {
    if (n == 0) return ZKB_OK;
    const uint64_t threads_total = (n + BI_CHUNK - 1) / BI_CHUNK;
    batch_invert_kernel<<<(unsigned)((threads_total + 127) / 128), 128, 0, st>>>(a, out, n);
    ctx->launches++;
    ZKB_CUDA(cudaGetLastError());
    return ZKB_OK;
}

}  // namespace zkb
using namespace zkb;

static unsigned grid_for(zkb_ctx *ctx, uint64_t n, int threads) {
    uint64_t blocks = (n + threads - 1) / threads;
    uint64_t cap = (uint64_t)ctx->sm_count * 16;
    return (unsigned)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

extern "C" int32_t zkb_field_binop_dev(zkb_ctx *ctx, int32_t field, int32_t op, const uint64_t *a, const uint64_t *b, uint64_t *out,
                                       uint64_t n, void *stream) {
    ZKB_ARG(ctx && a && b && out && op >= 0 && op <= 2 && (field == 0 || field == 1));
    if (n == 0) return ZKB_OK;
    cudaStream_t st = pick_stream(ctx, stream);
    if (field == 0) binop_kernel<FrParams><<<grid_for(ctx, n, 256), 256, 0, st>>>(op, (const Fr *)a, (const Fr *)b, (Fr *)out, n);
    else binop_kernel<FqParams><<<grid_for(ctx, n, 256), 256, 0, st>>>(op, (const Fq *)a, (const Fq *)b, (Fq *)out, n);
    ctx->launches++;
    ZKB_CUDA(cudaGetLastError());
    return ZKB_OK;
}

extern "C" int32_t zkb_field_unop_dev(zkb_ctx *ctx, int32_t field, int32_t op, const uint64_t *a, uint64_t *out, uint64_t n, void *stream) {
    ZKB_ARG(ctx && a && out && op >= 0 && op <= 4 && (field == 0 || field == 1));
    if (n == 0) return ZKB_OK;
    cudaStream_t st = pick_stream(ctx, stream);
    if (field == 0) unop_kernel<FrParams><<<grid_for(ctx, n, 256), 256, 0, st>>>(op, (const Fr *)a, (Fr *)out, n);
    else unop_kernel<FqParams><<<grid_for(ctx, n, 256), 256, 0, st>>>(op, (const Fq *)a, (Fq *)out, n);
    ctx->launches++;
    ZKB_CUDA(cudaGetLastError());
    return ZKB_OK;
}

extern "C" int32_t zkb_fr_batch_invert_dev(zkb_ctx *ctx, const uint64_t *a, uint64_t *out, uint64_t n, void *stream) {
    ZKB_ARG(ctx && a && out && a != out);
    if (n == 0) return ZKB_OK;
    return batch_invert_device(ctx, (const Fr *)a, (Fr *)out, n, pick_stream(ctx, stream));
}
