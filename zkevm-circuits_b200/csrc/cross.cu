// cross.cu -- multi-GPU building blocks: the size-P transform ACROSS ranks that follows the
// all-to-all of a domain-sharded NTT, and the host-side sum of per-rank MSM partial results.
//
// Sharding of one huge NTT (n = P * M, P ranks; halo2's best_fft has no distributed form, this is new):
//   input  : rank r holds the cyclic subsequence x[r + P t], t < M
//   step A : local size-M NTT with omega^P                       (existing kernels)
//   step B : multiply element t by omega^(r t)                   (fr_powers + element-wise multiply)
//   step C : ONE all-to-all: rank s receives, from every rank j, the block t in [s M/P, (s+1) M/P)
//   step D : zkb_ntt_cross_dev: out[k][t] = sum_j in[j][t] * omega_P^(j k)  (this file; log2 P radix-2 stages in registers, P <= 16)
//   output : rank s holds X[k M + s M/P + t] for k < P, t < M/P  ("strips"; zkb200.parallel.strips_to_natural documents it)
// MSM shards by point range (no exchange of points); partial sums are all-gathered as 64-byte affine points and added
// with zkb_g1_sum_affine_host (pure host code, also usable on a box without a GPU).
#include "common.cuh"
#include <string.h>

namespace zkb {

struct CrossTw { Fr w[16]; };  // omega_P^i, i < P

// log2(P) radix-2 decimation-in-frequency stages in registers: (P/2) log2 P multiplies per t (12 at P = 8, 32 at P = 16; the
// earlier direct O(P^2) form needed 56 / 240 and spilled); outputs leave in bit-reversed register order.
template <int P>
__global__ void __launch_bounds__(128) ntt_cross_kernel(const Fr *__restrict__ in, Fr *__restrict__ out, uint64_t len, CrossTw tw) {
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t >= len) return;
    Fr v[P];
#pragma unroll
    for (int j = 0; j < P; ++j) v[j] = fp_load(in + (size_t)j * len + t);
#pragma unroll
    for (int half = P / 2; half >= 1; half >>= 1) {
#pragma unroll
        for (int b = 0; b < P; b += 2 * half) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const Fr u = v[b + j], w = v[b + j + half];
                v[b + j] = fp_add(u, w);
                const Fr d = fp_sub(u, w);
                const int e = j * (P / (2 * half));   // omega_P^e
                v[b + j + half] = e == 0 ? d : fp_mul(d, tw.w[e]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < P; ++i) {
        int k = 0;
#pragma unroll
        for (int bit = 1, rb = P >> 1; bit < P; bit <<= 1, rb >>= 1) if (i & bit) k |= rb;
        fp_store(out + (size_t)k * len + t, v[i]);
    }
}

}  // namespace zkb
using namespace zkb;

extern "C" int32_t zkb_ntt_cross_dev(zkb_ctx *ctx, const uint64_t *in_dev, uint64_t *out_dev, uint32_t p, uint64_t len, const uint64_t omega_p[4],
                                     void *stream) {
    ZKB_ARG(ctx && in_dev && out_dev && omega_p && in_dev != out_dev);
    ZKB_ARG(p == 1 || p == 2 || p == 4 || p == 8 || p == 16);
    if (len == 0) return ZKB_OK;
    cudaStream_t st = pick_stream(ctx, stream);
    Fr w;
    memcpy(w.l, omega_p, 32);
    // order check: w^p == 1 and (p > 1) w^(p/2) == -1
    if (!(fp_pow_u64(w, p) == Fr::one()) || (p > 1 && !fp_add(fp_pow_u64(w, p / 2), Fr::one()).is_zero())) {
        set_error("omega_p does not have order %u", p);
        return ZKB_ERR_ARG;
    }
    CrossTw tw;
    tw.w[0] = Fr::one();
    for (uint32_t i = 1; i < 16; ++i) tw.w[i] = i < p ? fp_mul(tw.w[i - 1], w) : Fr::zero();
    const unsigned blocks = (unsigned)((len + 127) / 128);
    const Fr *in = (const Fr *)in_dev;
    Fr *out = (Fr *)out_dev;
    switch (p) {
    case 1: ZKB_CUDA(cudaMemcpyAsync(out, in, len * sizeof(Fr), cudaMemcpyDeviceToDevice, st)); break;
    case 2: ntt_cross_kernel<2><<<blocks, 128, 0, st>>>(in, out, len, tw); break;
    case 4: ntt_cross_kernel<4><<<blocks, 128, 0, st>>>(in, out, len, tw); break;
    case 8: ntt_cross_kernel<8><<<blocks, 128, 0, st>>>(in, out, len, tw); break;
    default: ntt_cross_kernel<16><<<blocks, 128, 0, st>>>(in, out, len, tw); break;
    }
    ctx->launches++;
    ZKB_CUDA(cudaGetLastError());
    return ZKB_OK;
}

// sum of `count` affine points (host buffers); pure host arithmetic, no CUDA device needed
extern "C" int32_t zkb_g1_sum_affine_host(const uint64_t *points, uint64_t count, uint64_t out_affine[8], uint8_t *out_compressed) {
    ZKB_ARG(out_affine && (points || count == 0));
    G1Xyzz acc = G1Xyzz::identity();
    for (uint64_t i = 0; i < count; ++i) {
        G1Affine p;
        memcpy(&p, points + 8 * i, 64);
        g1_add_mixed(acc, p);
    }
    const G1Affine r = g1_to_affine(acc);
    memcpy(out_affine, &r, 64);
    if (out_compressed) g1_compress(r, out_compressed);
    return ZKB_OK;
}
