// expr.cu -- device interpreter for constraint-expression programs (see expr.cuh).
#include "common.cuh"
#include "expr.cuh"

namespace zkb {

struct ExprLaunch {
    const Instr *code;
    uint32_t ncode;
    const Fr *const *cols;   // column slot -> device array of (1 << log_n) elements
    const Fr *consts;
    Fr *const *outs;         // output slot -> device array
    uint32_t log_n;
    uint32_t out_stride;     // output index = row * out_stride + out_offset
    uint32_t out_offset;
};

template <int NREGS>
__global__ void __launch_bounds__(128) expr_kernel(ExprLaunch L) {
    const uint32_t n = 1u << L.log_n, mask = n - 1;
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    Fr regs[NREGS];
    Fr acc = Fr::zero(), acc2 = Fr::zero();
    for (uint32_t pc = 0; pc < L.ncode; ++pc) {
        const Instr in = L.code[pc];
        switch (in.op) {
        case OP_LOADCOL: {
            const int32_t rot = (int32_t)(int16_t)(in.imm >> 16);
            const uint32_t r = (row + (uint32_t)rot) & mask;
            regs[in.dst] = fp_load(L.cols[in.imm & 0xffffu] + r);
        } break;
        case OP_LOADCONST: regs[in.dst] = fp_load(L.consts + in.imm); break;
        case OP_ADD: { Fr a = regs[in.a], b = regs[in.b]; regs[in.dst] = fp_add(a, b); } break;
        case OP_SUB: { Fr a = regs[in.a], b = regs[in.b]; regs[in.dst] = fp_sub(a, b); } break;
        case OP_MUL: { Fr a = regs[in.a], b = regs[in.b]; regs[in.dst] = fp_mul(a, b); } break;
        case OP_NEG: { Fr a = regs[in.a]; regs[in.dst] = fp_neg(a); } break;
        case OP_HORNER: acc = fp_add(fp_mul(acc, fp_load(L.consts + in.imm)), regs[in.a]); break;
        case OP_STORE: fp_store(L.outs[in.imm] + (size_t)row * L.out_stride + L.out_offset, regs[in.a]); break;
        case OP_STOREACC:
            fp_store(L.outs[in.imm & 0xffu] + (size_t)row * L.out_stride + L.out_offset, fp_mul(acc, fp_load(L.consts + (in.imm >> 8))));
            break;
        case OP_CLEARACC: acc = Fr::zero(); break;
        case OP_HORNER2: acc2 = fp_add(fp_mul(acc2, fp_load(L.consts + in.imm)), regs[in.a]); break;
        case OP_FOLD:
            acc = fp_add(fp_mul(acc, fp_load(L.consts + in.imm)), fp_mul(regs[in.a], acc2));
            acc2 = Fr::zero();
            break;
        default: break;
        }
    }
}

// d_code / d_cols / d_consts / d_outs are device pointers
int32_t expr_run_device(zkb_ctx *ctx, const Instr *d_code, uint32_t ncode, int nregs, const Fr *const *d_cols, const Fr *d_consts,
                        Fr *const *d_outs, uint32_t log_n, uint32_t out_stride, uint32_t out_offset, cudaStream_t st) {
    ExprLaunch L{d_code, ncode, d_cols, d_consts, d_outs, log_n, out_stride, out_offset};
    const uint32_t n = 1u << log_n;
    const unsigned blocks = (n + 127) / 128;
    ProfScope ps_(ctx, PROF_EXPR, st);
    if (nregs <= 8) expr_kernel<8><<<blocks, 128, 0, st>>>(L);
    else if (nregs <= 16) expr_kernel<16><<<blocks, 128, 0, st>>>(L);
    else if (nregs <= 32) expr_kernel<32><<<blocks, 128, 0, st>>>(L);
    else expr_kernel<64><<<blocks, 128, 0, st>>>(L);
    ctx->launches++;
    ZKB_CUDA(cudaGetLastError());
    return ZKB_OK;
}

}  // namespace zkb
