// ntt.cu -- radix-2 NTT over BN254 Fr for sm_100a.
//
// Replaces halo2_proofs::arithmetic::best_fft (halo2_proofs 1.1.0 @ e5ddf67 src/arithmetic.rs; reached from
// circuit-benchmarks/src/super_circuit.rs:117-132 through EvaluationDomain::{lagrange_to_coeff, coeff_to_extended,
// extended_to_coeff}).  Same contract: in place, natural order in / natural order out, a'[k] = sum_j a[j] w^(jk).
//
// B200 design (NOT upstream's bit-reverse + log n global layers):
//   n = A1 * A2 (* A3), every factor <= 2^12.  Pass p runs all length-A_p transforms of the Cooley-Tukey index
//   splitting j = j_a * inner + j_in entirely inside one CTA's shared memory (A_p * 32 B <= 128 KB, split into two
//   16-byte planes so 128-bit LDS/STS are conflict free), decimation-in-frequency, and applies the inter-pass
//   twiddle w_n^(j_in * k_a) on the way out.  A 2^24 transform is two passes = two reads + two writes of the data,
//   the minimum for a working set larger than shared memory.  Each 32-byte element is exactly one DRAM sector, so
//   the strided column gathers of pass 1 and the digit-reversed scatter of the last pass move no wasted bytes.
//   Twiddles: a 2^(a-1)-entry local table per pass (L1/L2 resident) + a two-level table (w^lo * w^(hi*4096)) for
//   the inter-pass factor.  The 1/n of the inverse transform is folded into the two-level table (no extra multiply).
//   The kernels are bound by the integer-multiply pipe: 1 Montgomery multiply (264 IMAD) per butterfly.
#include "common.cuh"

namespace zkb {

constexpr int NTT_MAX_BITS = 12;  // largest in-CTA transform: 4096 elements = 128 KB of shared memory
constexpr int TW_LO_BITS = 12;

// ZETA = 7^((r-1)/3) and ZETA^2 (Montgomery form); EvaluationDomain::g_coset / g_coset_inv (poly/domain.rs)
__device__ __constant__ uint32_t ZETA_POW[2][8];

struct PassArgs {
    uint32_t a;          // log2 of the in-CTA transform length
    uint32_t log_inner;  // log2 of the element stride inside this pass
    uint32_t log_n;
    uint32_t tw_shift;   // boundary exponent = (j_in * k) << tw_shift
    uint32_t is_final;   // inner == 1: digit-reversed store
    uint32_t a1, a2;     // bits of the earlier passes (final store index = k1 + (k2 << a1) + (k << (a1 + a2)))
    uint32_t coset_in;   // multiply input i by ZETA^(i mod 3)        (first pass only)
    uint32_t coset_out;  // multiply output k by ZETA^(-(k mod 3))    (final pass only)
    uint32_t use_scale;  // multiply outputs by *scale                 (single-pass transforms only)
    const Fr *loc;
    const Fr *tw_lo;
    const Fr *tw_hi;
    const Fr *scale;
    const Fr *in_scale;  // optional per-element input multiplier (first pass only): coset scaling tables
    // batching over blockIdx.y: column y reads in_tbl[y] (or in + y * in_stride) and writes out_tbl[y] (or out + y * out_stride)
    const Fr *tw_full;   // optional complete inter-pass table (first pass of a two-pass plan)
    const Fr *const *in_tbl;
    Fr *const *out_tbl;
    uint64_t in_stride, out_stride;
};

// element i lives at slot i ^ ((i >> 3) & 7): keeps unit-stride runs conflict free AND makes the stride-8 accesses of the
// last radix-8 round (a thread owns 8 consecutive elements) hit 8 distinct 16-byte bank groups
__device__ __forceinline__ uint32_t swz(uint32_t i) { return i ^ ((i >> 3) & 7u); }
__device__ __forceinline__ Fr smem_ld(const uint4 *lo, const uint4 *hi, uint32_t i) {
    const uint32_t k = swz(i);
    uint4 a = lo[k], b = hi[k];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void smem_st(uint4 *lo, uint4 *hi, uint32_t i, const Fr &v) {
    const uint32_t k = swz(i);
    lo[k] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    hi[k] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
// R consecutive DIF stages (s .. s+R-1) on groups of 2^R elements held in registers: one shared-memory round trip and
// one barrier per R stages; the 2^R - 1 distinct twiddles of a group are loaded once.
template <int R>
__device__ __forceinline__ void ntt_round(uint4 *lo, uint4 *hi, uint32_t a, uint32_t s, const Fr *__restrict__ loc, uint32_t tid, uint32_t nt) {
    constexpr int E = 1 << R;
    const uint32_t A = 1u << a;
    const uint32_t h = A >> (s + 1);
    const uint32_t q = h >> (R - 1);          // spacing of the group's elements = half distance of the round's last stage
    const uint32_t lq = 31 - __clz(q);
    for (uint32_t g = tid; g < (A >> R); g += nt) {
        const uint32_t p_local = g & (q - 1);
        const uint32_t base = ((g >> lq) << (lq + R)) + p_local;
        Fr x[E];
#pragma unroll
        for (int m = 0; m < E; ++m) x[m] = smem_ld(lo, hi, base + ((uint32_t)m << lq));
#pragma unroll
        for (int t = 0; t < R; ++t) {
            constexpr int dummy = 0;
            (void)dummy;
            const int d = E >> (t + 1);
            const bool last_trivial = ((uint32_t)d << lq) == 1u;   // half distance 1: twiddle is omega^0
#pragma unroll
            for (int m = 0; m < E; ++m) {
                if ((m & d) == 0) {
                    const Fr u = x[m], v = x[m + d];
                    x[m] = fp_add(u, v);
                    Fr dif = fp_sub(u, v);
                    if (!last_trivial) {
                        const uint32_t pos = p_local + ((uint32_t)(m & (d - 1)) << lq);
                        dif = fp_mul(dif, fp_load(loc + ((size_t)pos << (s + t))));
                    }
                    x[m + d] = dif;
                }
            }
        }
#pragma unroll
        for (int m = 0; m < E; ++m) smem_st(lo, hi, base + ((uint32_t)m << lq), x[m]);
    }
}
__device__ __forceinline__ Fr zeta_pow(int i) {  // i in {1,2}
    Fr z;
#pragma unroll
    for (int k = 0; k < 8; ++k) z.l[k] = ZETA_POW[i - 1][k];
    return z;
}

__global__ void __launch_bounds__(512) ntt_pass_kernel(const Fr *__restrict__ in_base, Fr *__restrict__ out_base, PassArgs p) {
    extern __shared__ uint4 smem[];
    const Fr *__restrict__ in = p.in_tbl ? p.in_tbl[blockIdx.y] : in_base + (size_t)blockIdx.y * p.in_stride;
    Fr *__restrict__ out = p.out_tbl ? p.out_tbl[blockIdx.y] : out_base + (size_t)blockIdx.y * p.out_stride;
    const uint32_t A = 1u << p.a;
    uint4 *lo = smem, *hi = smem + A;
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const uint64_t sub = blockIdx.x;
    const uint64_t outer = sub >> p.log_inner;
    const uint64_t j_in = sub & ((1ull << p.log_inner) - 1);
    const uint64_t base = (outer << (p.a + p.log_inner)) + j_in;

    // load phase: 4 independent 32-byte loads in flight per thread before anything is consumed
    for (uint32_t j0 = tid; j0 < A; j0 += 4 * nt) {
        Fr v[4];
        uint64_t idx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t j = j0 + u * nt;
            idx[u] = base + ((uint64_t)j << p.log_inner);
            if (j < A) v[u] = fp_load_stream(in + idx[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t j = j0 + u * nt;
            if (j < A) {
                if (p.coset_in) {
                    const uint32_t m = (uint32_t)(idx[u] % 3);
                    if (m) v[u] = fp_mul(v[u], zeta_pow(m));
                }
                if (p.in_scale) v[u] = fp_mul(v[u], fp_load(p.in_scale + idx[u]));
                smem_st(lo, hi, j, v[u]);
            }
        }
    }
    __syncthreads();

    // decimation in frequency: natural order in, bit-reversed order out (inside shared memory); radix-8 rounds in registers
    {
        uint32_t s = 0;
        while (p.a - s >= 3) { ntt_round<3>(lo, hi, p.a, s, p.loc, tid, nt); __syncthreads(); s += 3; }
        if (p.a - s == 2) { ntt_round<2>(lo, hi, p.a, s, p.loc, tid, nt); __syncthreads(); }
        else if (p.a - s == 1) { ntt_round<1>(lo, hi, p.a, s, p.loc, tid, nt); __syncthreads(); }
    }

    if (p.is_final) {
        const uint64_t k1 = p.a2 ? (outer >> p.a2) : outer;
        const uint64_t k2 = p.a2 ? (outer & ((1ull << p.a2) - 1)) : 0;
        const uint64_t obase = (p.a1 ? k1 : 0) + (k2 << p.a1);
        for (uint32_t q = tid; q < A; q += nt) {
            const uint32_t k = p.a ? (__brev(q) >> (32 - p.a)) : 0;
            Fr v = smem_ld(lo, hi, q);
            const uint64_t oidx = obase + ((uint64_t)k << (p.a1 + p.a2));
            if (p.use_scale) v = fp_mul(v, fp_load(p.scale));
            if (p.coset_out) {
                const uint32_t m = (uint32_t)(oidx % 3);
                if (m) v = fp_mul(v, zeta_pow(3 - m));  // ZETA^(-m) = ZETA^(3-m)
            }
            fp_store_stream(out + oidx, v);
        }
    } else {
        // inter-pass twiddle w_n^(j_in * k) = lo[e & 4095] * hi[e >> 12]: the table reads of two elements are issued together
        const bool two_level = p.log_n > TW_LO_BITS;
        for (uint32_t q0 = tid; q0 < A; q0 += 2 * nt) {
            uint32_t k[2];
            Fr tl[2], th[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const uint32_t q = q0 + u * nt;
                k[u] = __brev(q) >> (32 - p.a);
                if (q < A) {
                    if (p.tw_full) tl[u] = fp_load_stream(p.tw_full + (j_in << p.a) + k[u]);
                    else {
                        const uint64_t e = (j_in * (uint64_t)k[u]) << p.tw_shift;
                        tl[u] = fp_load(p.tw_lo + (e & ((1u << TW_LO_BITS) - 1)));
                        if (two_level) th[u] = fp_load(p.tw_hi + (e >> TW_LO_BITS));
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const uint32_t q = q0 + u * nt;
                if (q < A) {
                    Fr tw = (two_level && !p.tw_full) ? fp_mul(tl[u], th[u]) : tl[u];
                    const Fr v = fp_mul(smem_ld(lo, hi, q), tw);
                    fp_store_stream(out + base + ((uint64_t)k[u] << p.log_inner), v);
                }
            }
        }
    }
}

// T[j_in * A + k] = lo[e & 4095] * hi[e >> 12], e = j_in * k  (lo may carry a folded scale)
__global__ void build_full_table_kernel(const Fr *__restrict__ lo, const Fr *__restrict__ hi, uint32_t log_n, uint32_t a0, Fr *__restrict__ out) {
    const uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (idx >= (1ull << log_n)) return;
    const uint64_t j_in = idx >> a0, k = idx & ((1ull << a0) - 1);
    const uint64_t e = j_in * k;
    Fr tw = fp_load(lo + (e & ((1u << TW_LO_BITS) - 1)));
    if (log_n > TW_LO_BITS) tw = fp_mul(tw, fp_load(hi + (e >> TW_LO_BITS)));
    fp_store(out + idx, tw);
}

__global__ void scale_table_kernel(const Fr *__restrict__ in, Fr *__restrict__ out, const Fr *__restrict__ scale, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) fp_store(out + i, fp_mul(fp_load(in + i), fp_load(scale)));
}

// ---------------------------------------------------------------------------------------------------------
// host side: plans
// ---------------------------------------------------------------------------------------------------------
static const uint32_t FR_ROOT_OF_UNITY_CANON[8] = {0x60c37c9cu, 0xd34f1ed9u, 0xd39329c8u, 0x3215cf6du, 0x3dd31f74u, 0x98865ea9u, 0x166d18b7u, 0x03ddb9f5u};
static const uint32_t FR_ZETA_CANON[8] = {0xb99c90ddu, 0x8b17ea66u, 0x8d8daaa7u, 0x5bfc4108u, 0x41a91758u, 0xb3c4d79du, 0x00000000u, 0x00000000u};

Fr host_root_of_unity(uint32_t k) {
    Fr c;
    for (int i = 0; i < 8; ++i) c.l[i] = FR_ROOT_OF_UNITY_CANON[i];
    Fr w = fp_from_canonical(c);
    for (uint32_t i = k; i < 28; ++i) w = fp_sqr(w);
    return w;
}
Fr host_zeta() {
    Fr c;
    for (int i = 0; i < 8; ++i) c.l[i] = FR_ZETA_CANON[i];
    return fp_from_canonical(c);
}

static int32_t upload_powers(zkb_ctx *ctx, const Fr &w, size_t count, Fr **out) {
    std::vector<Fr> h(count);
    Fr cur = Fr::one();
    for (size_t i = 0; i < count; ++i) { h[i] = cur; cur = fp_mul(cur, w); }
    ZKB_CUDA(cudaMalloc((void **)out, count * sizeof(Fr)));
    ZKB_CUDA(cudaMemcpyAsync(*out, h.data(), count * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    ZKB_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKB_OK;
}

static bool g_zeta_uploaded[64] = {false};

static int32_t get_plan(zkb_ctx *ctx, uint32_t log_n, const Fr &omega, NttPlan **out) {
    std::array<uint64_t, 5> key;
    key[0] = log_n;
    for (int i = 0; i < 4; ++i) key[1 + i] = (uint64_t)omega.l[2 * i] | ((uint64_t)omega.l[2 * i + 1] << 32);
    auto it = ctx->ntt_plans.find(key);
    if (it != ctx->ntt_plans.end()) { *out = &it->second; return ZKB_OK; }

    if (!g_zeta_uploaded[ctx->device & 63]) {
        Fr z = host_zeta(), z2 = fp_sqr(z);
        uint32_t h[2][8];
        for (int i = 0; i < 8; ++i) { h[0][i] = z.l[i]; h[1][i] = z2.l[i]; }
        ZKB_CUDA(cudaMemcpyToSymbol(ZETA_POW, h, sizeof(h)));
        g_zeta_uploaded[ctx->device & 63] = true;
    }

    // order check: omega^(2^log_n) == 1 and omega^(2^(log_n-1)) == -1
    {
        Fr t = omega;
        for (uint32_t i = 0; i + 1 < log_n; ++i) t = fp_sqr(t);
        if (log_n >= 1) {
            if (!(fp_add(t, Fr::one()).is_zero())) { set_error("omega does not have order 2^%u", log_n); return ZKB_ERR_ARG; }
        } else if (!(omega == Fr::one())) { set_error("omega must be 1 for log_n = 0"); return ZKB_ERR_ARG; }
    }

    NttPlan plan;
    plan.log_n = log_n;
    if (log_n <= NTT_MAX_BITS) { plan.npass = 1; plan.bits[0] = log_n; }
    else if (log_n <= 2 * NTT_MAX_BITS) { plan.npass = 2; plan.bits[0] = (log_n + 1) / 2; plan.bits[1] = log_n - plan.bits[0]; }
    else { plan.npass = 3; plan.bits[0] = (log_n + 2) / 3; plan.bits[1] = (log_n - plan.bits[0] + 1) / 2; plan.bits[2] = log_n - plan.bits[0] - plan.bits[1]; }

    const uint32_t lo_bits = log_n < TW_LO_BITS ? log_n : TW_LO_BITS;
    ZKB_TRY(upload_powers(ctx, omega, (size_t)1 << lo_bits, &plan.tw_lo));
    if (log_n > TW_LO_BITS) {
        Fr w = omega;
        for (int i = 0; i < TW_LO_BITS; ++i) w = fp_sqr(w);
        ZKB_TRY(upload_powers(ctx, w, (size_t)1 << (log_n - TW_LO_BITS), &plan.tw_hi));
    }
    for (int ps = 0; ps < plan.npass; ++ps) {
        const int a = plan.bits[ps];
        Fr w = omega;
        for (uint32_t i = a; i < log_n; ++i) w = fp_sqr(w);  // omega^(2^(log_n - a)) : order 2^a
        ZKB_TRY(upload_powers(ctx, w, a ? ((size_t)1 << (a - 1)) : 1, &plan.loc[ps]));
    }
    auto ins = ctx->ntt_plans.emplace(key, plan);
    *out = &ins.first->second;
    return ZKB_OK;
}

static bool g_attr_set = false;

// batch of `count` transforms: column y reads d_src_tbl[y], writes d_dst_tbl[y] (device pointer tables; tables may alias).
// With count == 1 and null tables, src/dst are used directly.
int32_t ntt_fr_batch_device(zkb_ctx *ctx, const Fr *src_data, Fr *data, const Fr *const *d_src_tbl, Fr *const *d_dst_tbl, uint32_t count,
                            uint32_t log_n, const Fr &omega, const Fr *scale_host, int coset_zeta, const Fr *d_in_scale, cudaStream_t st) {
    ZKB_ARG(log_n <= 3 * NTT_MAX_BITS && log_n <= 28 && count >= 1);
    ZKB_ARG(coset_zeta >= 0 && coset_zeta <= 2);
    NttPlan *plan = nullptr;
    ZKB_TRY(get_plan(ctx, log_n, omega, &plan));
    if (!g_attr_set) {
        ZKB_CUDA(cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << NTT_MAX_BITS) * 32));
        g_attr_set = true;
    }
    const uint64_t n = 1ull << log_n;
    Fr *scratch = nullptr;
    if (plan->npass > 1) ZKB_TRY(scratch_get(ctx, SCR_NTT, (size_t)count * n * sizeof(Fr), (void **)&scratch));

    // scale: folded into the two-level twiddle table for multi-pass transforms
    Fr *d_scale = nullptr;
    const Fr *tw_lo = plan->tw_lo;
    if (scale_host) {
        void *misc = nullptr;
        ZKB_TRY(scratch_get(ctx, SCR_MISC, ((size_t)1 << TW_LO_BITS) * sizeof(Fr) + sizeof(Fr), &misc));
        d_scale = (Fr *)misc;
        ZKB_CUDA(cudaMemcpyAsync(d_scale, scale_host, sizeof(Fr), cudaMemcpyHostToDevice, st));
        if (plan->npass > 1) {
            Fr *scaled = d_scale + 1;
            const uint32_t cnt = 1u << TW_LO_BITS;
            scale_table_kernel<<<(cnt + 255) / 256, 256, 0, st>>>(plan->tw_lo, scaled, d_scale, cnt);
            ctx->launches++;
            tw_lo = scaled;
        }
    }

    // complete inter-pass table for two-pass plans (built once per plan / scale, n multiplies)
    const Fr *tw_full = nullptr;
    if (plan->npass == 2 && log_n <= 25) {
        const unsigned bb = (unsigned)((n + 255) / 256);
        if (!scale_host) {
            if (!plan->tw_full) {
                ZKB_CUDA(cudaMalloc((void **)&plan->tw_full, n * sizeof(Fr)));
                build_full_table_kernel<<<bb, 256, 0, st>>>(plan->tw_lo, plan->tw_hi, log_n, plan->bits[0], plan->tw_full);
                ctx->launches++;
            }
            tw_full = plan->tw_full;
        } else {
            if (!plan->tw_full_scaled) ZKB_CUDA(cudaMalloc((void **)&plan->tw_full_scaled, n * sizeof(Fr)));
            if (!plan->has_scaled || !(plan->scaled_key == *scale_host)) {
                build_full_table_kernel<<<bb, 256, 0, st>>>(tw_lo, plan->tw_hi, log_n, plan->bits[0], plan->tw_full_scaled);  // tw_lo = scaled copy
                ctx->launches++;
                plan->scaled_key = *scale_host;
                plan->has_scaled = true;
            }
            tw_full = plan->tw_full_scaled;
        }
    }

    uint32_t log_inner = log_n;
    for (int ps = 0; ps < plan->npass; ++ps) {
        const uint32_t a = plan->bits[ps];
        log_inner -= a;
        PassArgs p;
        p.a = a;
        p.log_inner = log_inner;
        p.log_n = log_n;
        p.is_final = (ps == plan->npass - 1);
        uint32_t consumed = 0;
        for (int q = 0; q <= ps; ++q) consumed += plan->bits[q];
        p.tw_shift = consumed - a;  // bits consumed by the earlier passes
        p.a1 = p.a2 = 0;
        if (p.is_final) {
            if (plan->npass == 2) { p.a1 = plan->bits[0]; }
            if (plan->npass == 3) { p.a1 = plan->bits[0]; p.a2 = plan->bits[1]; }
        }
        p.coset_in = (ps == 0 && coset_zeta == 1);
        p.coset_out = (p.is_final && coset_zeta == 2);
        p.use_scale = (plan->npass == 1 && scale_host != nullptr);
        p.loc = plan->loc[ps];
        p.tw_lo = (ps == 0) ? tw_lo : plan->tw_lo;  // only the first boundary carries the folded scale
        p.tw_hi = plan->tw_hi;
        p.scale = d_scale;
        p.in_scale = (ps == 0) ? d_in_scale : nullptr;
        p.tw_full = (ps == 0) ? tw_full : nullptr;
        p.in_tbl = nullptr;
        p.out_tbl = nullptr;
        p.in_stride = p.out_stride = 0;
        const Fr *src = nullptr;
        Fr *dst = nullptr;
        const bool first = (ps == 0), last = p.is_final;
        if (first) { if (d_src_tbl) p.in_tbl = d_src_tbl; else src = src_data; }
        else { src = scratch; p.in_stride = n; }
        if (last) { if (d_dst_tbl) p.out_tbl = d_dst_tbl; else dst = data; }
        else { dst = scratch; p.out_stride = n; }
        const uint32_t A = 1u << a;
        uint32_t threads = A / 8;  // one radix-8 group per thread and round
        if (threads < 64) threads = 64;
        if (threads > 512) threads = 512;
        const uint64_t blocks = n >> a;
        ntt_pass_kernel<<<dim3((unsigned)blocks, count), threads, (size_t)A * 32, st>>>(src, dst, p);
        ctx->launches++;
    }
    ZKB_CUDA(cudaGetLastError());
    return ZKB_OK;
}

int32_t ntt_fr_device(zkb_ctx *ctx, const Fr *src_data, Fr *data, uint32_t log_n, const Fr &omega, const Fr *scale_host, int coset_zeta,
                      const Fr *d_in_scale, cudaStream_t st) {
    return ntt_fr_batch_device(ctx, src_data, data, nullptr, nullptr, 1, log_n, omega, scale_host, coset_zeta, d_in_scale, st);
}

}  // namespace zkb

using namespace zkb;

extern "C" int32_t zkb_fr_root_of_unity(uint32_t k, uint64_t omega[4], uint64_t omega_inv[4]) {
    ZKB_ARG(k <= 28 && omega != nullptr);
    Fr w = host_root_of_unity(k);
    memcpy(omega, w.l, 32);
    if (omega_inv) {
        Fr wi = fp_inv(w);
        memcpy(omega_inv, wi.l, 32);
    }
    return ZKB_OK;
}

extern "C" int32_t zkb_ntt_fr_dev(zkb_ctx *ctx, uint64_t *data_dev, uint32_t log_n, const uint64_t omega[4], const uint64_t *scale,
                                  int32_t coset_zeta, void *stream) {
    ZKB_ARG(ctx && data_dev && omega);
    ZKB_CUDA(cudaSetDevice(ctx->device));
    Fr w, sc;
    memcpy(w.l, omega, 32);
    if (scale) memcpy(sc.l, scale, 32);
    return ntt_fr_device(ctx, (const Fr *)data_dev, (Fr *)data_dev, log_n, w, scale ? &sc : nullptr, coset_zeta, nullptr, pick_stream(ctx, stream));
}

extern "C" int32_t zkb_ntt_fr_host(zkb_ctx *ctx, uint64_t *data_host, uint32_t log_n, const uint64_t omega[4], const uint64_t *scale,
                                   int32_t coset_zeta) {
    ZKB_ARG(ctx && data_host && omega && log_n <= 28);
    ZKB_CUDA(cudaSetDevice(ctx->device));
    const size_t bytes = ((size_t)1 << log_n) * 32;
    void *d = nullptr;
    ZKB_TRY(scratch_get(ctx, SCR_HOSTIO_A, bytes, &d));
    ZKB_CUDA(cudaMemcpyAsync(d, data_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ZKB_TRY(zkb_ntt_fr_dev(ctx, (uint64_t *)d, log_n, omega, scale, coset_zeta, ctx->stream));
    ZKB_CUDA(cudaMemcpyAsync(data_host, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ZKB_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKB_OK;
}
