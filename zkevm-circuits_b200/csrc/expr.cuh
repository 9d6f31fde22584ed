// expr.cuh -- constraint-expression programs: host-side compiler (DAG -> linear register program) and the device
// interpreter that evaluates them row by row.
//
// Replaces halo2_proofs plonk/evaluation.rs (`GraphEvaluator`, `ValueSource`, `Calculation`, `Evaluator::evaluate_h`):
// upstream compiles every gate / lookup expression into a calculation graph and walks it per row on rayon threads.
// Here the whole quotient numerator -- custom gates, permutation argument terms and mv-lookup (logUp) terms, in
// upstream's Horner-in-y order -- is ONE program executed by ONE kernel: a thread owns a row, streams the column
// values it needs (rotations are index offsets inside the coset), keeps intermediates in a small register file and
// folds `acc = acc * y + term` after every constraint.  The same interpreter compresses lookup inputs / tables with
// theta on the Lagrange domain (mv_lookup/prover.rs `prepare`).
#pragma once
#include <stdint.h>
#include <map>
#include <string>
#include <tuple>
#include <vector>
#include "ff.cuh"

namespace zkb {

enum : uint8_t {
    OP_LOADCOL = 0,   // reg[dst] = cols[imm & 0xffff][(row + rot) mod n], rot = (int16)(imm >> 16)
    OP_LOADCONST = 1, // reg[dst] = consts[imm]
    OP_ADD = 2,
    OP_SUB = 3,
    OP_MUL = 4,
    OP_NEG = 5,
    OP_HORNER = 6,    // acc = acc * consts[imm] + reg[a]
    OP_STORE = 7,     // outs[imm][row] = reg[a]
    OP_STOREACC = 8,  // outs[imm][row] = acc * consts[b-as-index given in `a`]... see kernel: acc scaled by consts[a]
    OP_CLEARACC = 9,
    OP_HORNER2 = 10,  // acc2 = acc2 * consts[imm] + reg[a]            (inner Horner of a run of constraints sharing one factor)
    OP_FOLD = 11,     // acc = acc * consts[imm] + reg[a] * acc2; acc2 = 0
};

struct alignas(8) Instr {
    uint8_t op, dst, a, b;
    uint32_t imm;
};

constexpr int EXPR_MAX_REGS = 64;

// ---- host-side expression DAG with hash-consing --------------------------------------------------------------------
struct ENode {
    uint8_t kind;   // 0 col, 1 const, 2 add, 3 sub, 4 mul, 5 neg
    uint32_t a, b;  // children, or (slot, rot) for col, const index for const
};

class ExprBuilder {
public:
    std::vector<ENode> nodes;
    std::vector<Fr> consts;

    uint32_t col(uint32_t slot, int32_t rot) { return intern(0, slot, (uint32_t)rot); }
    uint32_t constant(const Fr &v) {
        // constants are deduplicated by value
        std::array<uint32_t, 8> key;
        for (int i = 0; i < 8; ++i) key[i] = v.l[i];
        auto it = const_index.find(key);
        uint32_t idx;
        if (it == const_index.end()) {
            idx = (uint32_t)consts.size();
            consts.push_back(v);
            const_index.emplace(key, idx);
        } else idx = it->second;
        return intern(1, idx, 0);
    }
    uint32_t add(uint32_t x, uint32_t y) { return intern(2, x, y); }
    uint32_t sub(uint32_t x, uint32_t y) { return intern(3, x, y); }
    uint32_t mul(uint32_t x, uint32_t y) { return intern(4, x, y); }
    uint32_t neg(uint32_t x) { return intern(5, x, 0); }
    uint32_t const_slot(const Fr &v) {  // index into consts (for HORNER / STOREACC immediates)
        uint32_t n = constant(v);
        return nodes[n].a;
    }

private:
    std::map<std::tuple<uint8_t, uint32_t, uint32_t>, uint32_t> index;
    std::map<std::array<uint32_t, 8>, uint32_t> const_index;
    uint32_t intern(uint8_t kind, uint32_t a, uint32_t b) {
        auto key = std::make_tuple(kind, a, b);
        auto it = index.find(key);
        if (it != index.end()) return it->second;
        nodes.push_back(ENode{kind, a, b});
        index.emplace(key, (uint32_t)nodes.size() - 1);
        return (uint32_t)nodes.size() - 1;
    }
};

// ---- program assembly: a sequence of "scopes"; inside a scope common subexpressions are computed once --------------
class ProgramBuilder {
public:
    explicit ProgramBuilder(ExprBuilder &eb) : eb(eb) {}
    std::vector<Instr> code;
    int max_regs_used = 0;
    std::string error;

    // evaluate `roots` (node ids) within one CSE scope and call emit_root(i, reg) after each is available
    enum RootAction { HORNER, STORE, HORNER2, FOLD };
    struct Root { uint32_t node; RootAction action; uint32_t imm; };
    bool scope(const std::vector<Root> &roots);
    void clear_acc() { code.push_back(Instr{OP_CLEARACC, 0, 0, 0, 0}); }
    void store_acc(uint32_t out_slot, uint32_t scale_const_index) {
        code.push_back(Instr{OP_STOREACC, 0, 0, 0, out_slot | (scale_const_index << 8)});
    }

private:
    ExprBuilder &eb;
};

inline bool ProgramBuilder::scope(const std::vector<Root> &roots) {
    // 1. reference counts inside the scope (number of parents + root uses)
    std::map<uint32_t, int> refs;
    std::vector<uint32_t> stack;
    std::map<uint32_t, bool> seen;
    for (auto &r : roots) {
        refs[r.node]++;
        if (!seen[r.node]) { seen[r.node] = true; stack.push_back(r.node); }
    }
    while (!stack.empty()) {
        uint32_t n = stack.back();
        stack.pop_back();
        const ENode &e = eb.nodes[n];
        if (e.kind >= 2) {
            uint32_t ch[2] = {e.a, e.b};
            int nch = e.kind == 5 ? 1 : 2;
            for (int i = 0; i < nch; ++i) {
                refs[ch[i]]++;
                if (!seen[ch[i]]) { seen[ch[i]] = true; stack.push_back(ch[i]); }
            }
        }
    }
    // 2. emit with a register pool; a node's register is released when its last use is consumed
    std::map<uint32_t, int> reg_of;
    std::vector<int> free_regs;
    for (int i = EXPR_MAX_REGS - 1; i >= 0; --i) free_regs.push_back(i);
    auto alloc = [&]() -> int {
        if (free_regs.empty()) return -1;
        int r = free_regs.back();
        free_regs.pop_back();
        if (r + 1 > max_regs_used) max_regs_used = r + 1;
        return r;
    };
    auto release_use = [&](uint32_t n) {
        if (--refs[n] == 0) {
            free_regs.push_back(reg_of[n]);
            reg_of.erase(n);
        }
    };
    // iterative post-order evaluation
    struct Frame { uint32_t node; int state; };
    for (auto &r : roots) {
        if (!reg_of.count(r.node)) {
            std::vector<Frame> st;
            st.push_back({r.node, 0});
            while (!st.empty()) {
                Frame &f = st.back();
                const ENode e = eb.nodes[f.node];
                if (reg_of.count(f.node)) { st.pop_back(); continue; }
                if (e.kind < 2) {
                    int rg = alloc();
                    if (rg < 0) { error = "expression needs more than 64 live registers"; return false; }
                    if (e.kind == 0) code.push_back(Instr{OP_LOADCOL, (uint8_t)rg, 0, 0, (e.a & 0xffffu) | ((uint32_t)(uint16_t)(int16_t)(int32_t)e.b << 16)});
                    else code.push_back(Instr{OP_LOADCONST, (uint8_t)rg, 0, 0, e.a});
                    reg_of[f.node] = rg;
                    st.pop_back();
                    continue;
                }
                const int nch = e.kind == 5 ? 1 : 2;
                if (f.state == 0) {
                    f.state = 1;
                    if (!reg_of.count(e.a)) { st.push_back({e.a, 0}); continue; }
                }
                if (f.state == 1) {
                    f.state = 2;
                    if (nch == 2 && !reg_of.count(e.b)) { st.push_back({e.b, 0}); continue; }
                }
                // children ready
                const int ra = reg_of[e.a];
                const int rb = nch == 2 ? reg_of[e.b] : 0;
                const uint32_t me = f.node;
                // consume child uses first so the destination may reuse a dying child's register
                release_use(e.a);
                if (nch == 2) release_use(e.b);
                int rg = alloc();
                if (rg < 0) { error = "expression needs more than 64 live registers"; return false; }
                uint8_t op = e.kind == 2 ? OP_ADD : e.kind == 3 ? OP_SUB : e.kind == 4 ? OP_MUL : OP_NEG;
                code.push_back(Instr{op, (uint8_t)rg, (uint8_t)ra, (uint8_t)rb, 0});
                reg_of[me] = rg;
                st.pop_back();
            }
        }
        const int rr = reg_of[r.node];
        if (r.action == HORNER) code.push_back(Instr{OP_HORNER, 0, (uint8_t)rr, 0, r.imm});
        else if (r.action == HORNER2) code.push_back(Instr{OP_HORNER2, 0, (uint8_t)rr, 0, r.imm});
        else if (r.action == FOLD) code.push_back(Instr{OP_FOLD, 0, (uint8_t)rr, 0, r.imm});
        else code.push_back(Instr{OP_STORE, 0, (uint8_t)rr, 0, r.imm});
        release_use(r.node);
    }
    return true;
}

}  // namespace zkb
