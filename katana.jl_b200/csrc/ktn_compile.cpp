// ktn_compile.cpp -- tape compiler: prefix expression arrays -> shape programs + packed chunks.
//
// Replaces, for the GPU path, what MathProgBase.initialize(oracle, [:Grad,:Jac]) and
// jac_structure do for the reference separator (src/separators.jl:88-100): build one tape per
// constraint and the per-row Jacobian column lists.  Here rows with the same expression
// STRUCTURE share one program ("shape"); constants and column indices become per-row data
// packed structure-of-arrays in warp-sized chunks (SELL-32-sigma: rows are bucketed by shape
// inside windows of sigma consecutive rows).
//
// The generated programs perform exactly the arithmetic of the ReverseDiffSparse-style
// interpreter restated in oracle/ktn_oracle.c (forward_row / reverse_row), in the same order.
#include "ktn_compile.h"
#include <algorithm>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include "../../include/ktn.h"

namespace {

struct Src { uint8_t kind; uint32_t idx; };

struct TNode {
    int op, nc, send, parent;
    bool has_var = false, is_leaf = false, need_value = false, has_val = false;
    Src val{KTN_K_NONE, 0};  // where the node's forward value can be read (leaf, or persisted slot)
    int pslot = -1;          // saved partial d parent / d this (products, divisions, general powers)
    int expclass = 0;        // CONST exponent of POW: 1 -> ==1.0, 2 -> ==2.0 (structural)
    int cslot = -1, uslot = -1;
};

static int arity(int op) {
    switch (op) {
        case KTN_OP_CONST: case KTN_OP_VAR: return 0;
        case KTN_OP_ADD: case KTN_OP_MUL: return -1;
        case KTN_OP_SUB: case KTN_OP_DIV: case KTN_OP_POW: case KTN_OP_LE: case KTN_OP_LT: case KTN_OP_GE: case KTN_OP_GT: case KTN_OP_EQ: return 2;
        case KTN_OP_NEG: case KTN_OP_EXP: case KTN_OP_LOG: case KTN_OP_SQRT: case KTN_OP_ABS: case KTN_OP_SIN: case KTN_OP_COS: return 1;
        case KTN_OP_IFELSE: return 3;
        default: return -2;
    }
}

struct ShapeCompiler {
    std::vector<TNode>& t;
    std::vector<KtnIns> code;
    uint32_t nu, next_slot;
    std::vector<uint32_t> temp_slot;  // spilled temporaries by depth (depth >= 2)
    int temp_depth = 0;
    std::vector<char> jseen;
    std::string err;

    bool alias;   // Jacobian accumulators and exp saves live in dead constant slots of the (shared-memory) blob
    ShapeCompiler(std::vector<TNode>& nodes, uint32_t n_uniq, bool alias_) : t(nodes), nu(n_uniq), next_slot(alias_ ? n_uniq : 2 * n_uniq), jseen(n_uniq, 0), alias(alias_) {}

    void emit(uint8_t op, Src s = Src{KTN_K_NONE, 0}, uint16_t n = 0) { code.push_back(KtnIns{op, s.kind, n, s.idx, 0u, 0u}); }
    void emit_terms(uint8_t op, uint8_t tk_flags, uint32_t n, uint32_t c0, uint32_t u0, uint32_t s0) { code.push_back(KtnIns{op, tk_flags, (uint16_t)n, c0, u0, s0}); }

    // ---- fused term runs inside n-ary sums ----
    struct Run { size_t pos; uint32_t n; int tk; uint32_t c0, u0, s0; };
    std::unordered_map<int, std::vector<Run>> runs;   // ADD node -> runs over its children
    static int term_cstride(int tk) { return tk == KTN_T_EXP_AFF ? 2 : (tk == KTN_T_X || tk == KTN_T_SQ) ? 0 : 1; }
    bool is_data_const(int k) const { return t[k].op == KTN_OP_CONST && t[k].cslot >= 0; }
    bool is_sq_exp(int k) const { return t[k].op == KTN_OP_CONST && t[k].expclass == 2; }
    // recognises the canonical term forms; returns the term kind or -1
    int match_term(int k, int& c0, int& u) const {
        const TNode& n = t[k];
        c0 = 0; u = -1;
        if (n.op == KTN_OP_VAR) { u = n.uslot; return KTN_T_X; }
        if (n.op == KTN_OP_MUL && n.nc == 2) {
            int a = k + 1, b = t[a].send;
            if (is_data_const(a) && t[b].op == KTN_OP_VAR) { c0 = t[a].cslot; u = t[b].uslot; return KTN_T_MULC_X; }
            if (is_data_const(a) && t[b].op == KTN_OP_POW && t[b + 1].op == KTN_OP_VAR && is_sq_exp(t[b + 1].send)) { c0 = t[a].cslot; u = t[b + 1].uslot; return KTN_T_MULC_SQ; }
            return -1;
        }
        if (n.op == KTN_OP_POW) {
            int b = k + 1, e = t[b].send;
            if (!is_sq_exp(e)) return -1;
            if (t[b].op == KTN_OP_VAR) { u = t[b].uslot; return KTN_T_SQ; }
            if (t[b].op == KTN_OP_MUL && t[b].nc == 2 && is_data_const(b + 1) && t[t[b + 1].send].op == KTN_OP_VAR) { c0 = t[b + 1].cslot; u = t[t[b + 1].send].uslot; return KTN_T_SQ_MULC; }
            return -1;
        }
        if (n.op == KTN_OP_EXP) {
            int a = k + 1;                                  // ADD(MUL(c, x), d)
            if (t[a].op != KTN_OP_ADD || t[a].nc != 2) return -1;
            int mnode = a + 1, d = t[mnode].send;
            if (t[mnode].op != KTN_OP_MUL || t[mnode].nc != 2 || !is_data_const(mnode + 1) || t[t[mnode + 1].send].op != KTN_OP_VAR || !is_data_const(d)) return -1;
            if (t[d].cslot != t[mnode + 1].cslot + 1) return -1;
            c0 = t[mnode + 1].cslot; u = t[t[mnode + 1].send].uslot; return KTN_T_EXP_AFF;
        }
        return -1;
    }
    void find_runs(int k, const std::vector<int>& ch) {
        std::vector<Run> rs;
        size_t i = 0;
        while (i < ch.size()) {
            int c0, u; int tk = match_term(ch[i], c0, u);
            if (tk < 0) { ++i; continue; }
            size_t j = i + 1;
            const int cs = term_cstride(tk);
            while (j < ch.size() && j - i < 60000) {
                int c1, u1; int tk1 = match_term(ch[j], c1, u1);
                if (tk1 != tk || u1 != u + (int)(j - i) || (cs && c1 != c0 + cs * (int)(j - i))) break;
                ++j;
            }
            const uint32_t n = (uint32_t)(j - i);
            if (n >= 2 || tk == KTN_T_EXP_AFF) {
                Run r{i, n, tk, (uint32_t)c0, (uint32_t)u, 0u};
                if (tk == KTN_T_EXP_AFF && !alias) { r.s0 = next_slot; next_slot += n; }
                rs.push_back(r);
            }
            i = j;
        }
        if (!rs.empty()) runs[k] = rs;
    }
    Src push_temp() {
        int d = temp_depth++;
        if (d == 0) return Src{KTN_K_R1, 0};
        if (d == 1) return Src{KTN_K_R2, 0};
        if ((size_t)(d - 2) >= temp_slot.size()) temp_slot.push_back(next_slot++);
        return Src{KTN_K_S, temp_slot[d - 2]};
    }
    void pop_temp() { --temp_depth; }
    Src new_slot() { return Src{KTN_K_S, next_slot++}; }
    std::vector<int> children(int k) const { std::vector<int> c; int j = k + 1; for (int i = 0; i < t[k].nc; ++i) { c.push_back(j); j = t[j].send; } return c; }

    // ---- analysis: which forward values the reverse sweep reads ----
    void analyse() {
        for (int k = (int)t.size() - 1; k >= 0; --k) {
            TNode& n = t[k];
            if (n.op == KTN_OP_VAR) n.has_var = true;
            else if (n.op != KTN_OP_CONST) for (int c : children(k)) n.has_var = n.has_var || t[c].has_var;
        }
        for (int k = 0; k < (int)t.size(); ++k) {
            TNode& n = t[k];
            if (!n.has_var || n.is_leaf) continue;
            std::vector<int> ch = children(k);
            switch (n.op) {
                case KTN_OP_EXP: case KTN_OP_SQRT: n.need_value = true; break;
                case KTN_OP_LOG: case KTN_OP_ABS: case KTN_OP_SIN: case KTN_OP_COS: t[ch[0]].need_value = true; break;
                case KTN_OP_IFELSE: t[ch[0]].need_value = true; break;      // the reverse sweep re-reads the condition
                case KTN_OP_POW: if (t[ch[1]].expclass == 2 && t[ch[0]].has_var) t[ch[0]].need_value = true; break;
                case KTN_OP_MUL:
                    if (n.nc == 2) { if (t[ch[0]].has_var) t[ch[1]].need_value = true; if (t[ch[1]].has_var) t[ch[0]].need_value = true; }
                    else if (n.nc > 2) for (int c : ch) t[c].need_value = true;  // all factors are re-read for the partials
                    break;
                default: break;
            }
        }
    }

    // ---- forward ----
    void persist(int k) {  // after the node's value is in acc
        TNode& n = t[k];
        if (n.need_value && !n.has_val) { n.val = new_slot(); n.has_val = true; emit(KF_STORE, n.val); }
    }
    // acc = acc (op) value(c) for a commutative op, value(c) possibly complex
    void combine_commutative(uint8_t fop, int c) {
        if (t[c].is_leaf) { emit(fop, t[c].val); return; }
        Src T = push_temp(); emit(KF_STORE, T);
        gen_fwd(c);
        emit(fop, T); pop_temp();
    }
    void gen_binary(int k, uint8_t fop, uint8_t rop) {  // acc = L (op) R, non-commutative
        std::vector<int> ch = children(k); int L = ch[0], R = ch[1];
        if (t[R].is_leaf) { gen_fwd(L); emit(fop, t[R].val); }
        else if (t[L].is_leaf) { gen_fwd(R); emit(rop, t[L].val); }
        else if (t[L].need_value) { gen_fwd(L); gen_fwd(R); emit(rop, t[L].val); }
        else { gen_fwd(L); Src T = push_temp(); emit(KF_STORE, T); gen_fwd(R); emit(rop, T); pop_temp(); }
    }
    void gen_fwd(int k) {
        TNode& n = t[k];
        if (!err.empty()) return;
        if (n.is_leaf) { emit(KF_LOAD, n.val); return; }
        std::vector<int> ch = children(k);
        switch (n.op) {
            case KTN_OP_ADD: {
                find_runs(k, ch);
                const std::vector<Run>* rs = runs.count(k) ? &runs[k] : nullptr;
                size_t ri = 0;
                for (size_t i = 0; i < ch.size();) {
                    if (rs && ri < rs->size() && (*rs)[ri].pos == i) {
                        const Run& r = (*rs)[ri++];
                        emit_terms(KF_TERMS, (uint8_t)(r.tk | (i == 0 ? KTN_TF_FIRST : 0) | (alias && r.tk == KTN_T_EXP_AFF ? KTN_TF_SAVEBLOB : 0)), r.n, r.c0, r.u0, r.s0);
                        i += r.n;
                    } else {
                        if (i == 0) { gen_fwd(ch[0]); emit(KF_ADDZ); }
                        else combine_commutative(KF_ADD, ch[i]);
                        ++i;
                    }
                }
                break; }
            case KTN_OP_SUB: gen_binary(k, KF_SUB, KF_RSUB); break;
            case KTN_OP_MUL:
                if (n.nc <= 2 || !n.has_var) {
                    gen_fwd(ch[0]);
                    for (size_t i = 1; i < ch.size(); ++i) combine_commutative(KF_MUL, ch[i]);
                } else {
                    // every factor is persisted (need_value), then p and the "all but one" partials
                    for (int c : ch) if (!t[c].is_leaf) gen_fwd(c);
                    emit(KF_LOAD, t[ch[0]].val);
                    for (size_t i = 1; i < ch.size(); ++i) emit(KF_MUL, t[ch[i]].val);
                    Src P = new_slot(); emit(KF_STORE, P);
                    Src A = new_slot();
                    for (size_t i = 0; i < ch.size(); ++i) {
                        if (!t[ch[i]].has_var) continue;
                        // alt = product of the others (only needed in lanes where p == 0)
                        uint16_t nalt = (uint16_t)ch.size();  // LOAD + (nc-2) MUL + STORE
                        emit(KF_SKIPNZ, P, nalt);
                        bool first = true;
                        for (size_t j = 0; j < ch.size(); ++j) { if (j == i) continue; emit(first ? KF_LOAD : KF_MUL, t[ch[j]].val); first = false; }
                        emit(KF_STORE, A);
                        emit(KF_LOAD, P); emit(KF_FDIV, t[ch[i]].val); emit(KF_LDAUX, A); emit(KF_SELZ, P);
                        Src ps = new_slot(); t[ch[i]].pslot = (int)ps.idx; emit(KF_STORE, ps);
                    }
                    emit(KF_LOAD, P);
                }
                break;
            case KTN_OP_DIV: {
                gen_binary(k, KF_DIV, KF_RDIV);
                if (t[ch[0]].has_var) { Src s = new_slot(); t[ch[0]].pslot = (int)s.idx; emit(KF_STAUX, Src{KTN_K_NONE, s.idx}); }
                if (t[ch[1]].has_var) { Src s = new_slot(); t[ch[1]].pslot = (int)s.idx; emit(KF_DENP, Src{KTN_K_NONE, s.idx}); }
                break; }
            case KTN_OP_POW: {
                int B = ch[0], E = ch[1];
                if (t[E].expclass == 2) { gen_fwd(B); emit(KF_POW2); }
                else if (t[E].expclass == 1) { gen_fwd(B); }
                else {
                    if (t[E].is_leaf) { gen_fwd(B); emit(KF_LDAUX, t[E].val); }
                    else { gen_fwd(E); Src T = push_temp(); emit(KF_STORE, T); gen_fwd(B); emit(KF_LDAUX, T); pop_temp(); }
                    if (t[B].has_var) { Src s = new_slot(); t[B].pslot = (int)s.idx; emit(KF_POWPB, Src{KTN_K_NONE, s.idx}); }
                    if (t[E].has_var) { Src s = new_slot(); t[E].pslot = (int)s.idx; emit(KF_POWPE, Src{KTN_K_NONE, s.idx}); }
                    emit(KF_POWG);
                }
                break; }
            case KTN_OP_NEG: gen_fwd(ch[0]); emit(KF_NEG); break;
            case KTN_OP_EXP: gen_fwd(ch[0]); emit(KF_EXP); break;
            case KTN_OP_LOG: gen_fwd(ch[0]); emit(KF_LOG); break;
            case KTN_OP_SQRT: gen_fwd(ch[0]); emit(KF_SQRT); break;
            case KTN_OP_ABS: gen_fwd(ch[0]); emit(KF_ABS); break;
            case KTN_OP_SIN: gen_fwd(ch[0]); emit(KF_SIN); break;
            case KTN_OP_COS: gen_fwd(ch[0]); emit(KF_COS); break;
            case KTN_OP_IFELSE: {       // both branches are evaluated (JuMP's tape is linear); acc = then, aux = else, select on cond == 1
                const int C = ch[0], A = ch[1], B = ch[2];
                Src cs; bool ctemp = false;
                if (t[C].is_leaf) cs = t[C].val;
                else { gen_fwd(C); if (t[C].has_val) cs = t[C].val; else { cs = push_temp(); emit(KF_STORE, cs); ctemp = true; } }
                if (t[B].is_leaf) { gen_fwd(A); emit(KF_LDAUX, t[B].val); }
                else { gen_fwd(B); Src T = push_temp(); emit(KF_STORE, T); gen_fwd(A); emit(KF_LDAUX, T); pop_temp(); }
                emit(KF_SEL1, cs);
                if (ctemp) pop_temp();
                break; }
            case KTN_OP_LE: case KTN_OP_LT: case KTN_OP_GE: case KTN_OP_GT: case KTN_OP_EQ: {   // acc = (L cmp R) ? 1 : 0, L as the operand, R in acc
                const int L = ch[0], R = ch[1];
                const uint16_t cmp = (uint16_t)(n.op - KTN_OP_LE);
                if (t[L].is_leaf) { gen_fwd(R); emit(KF_CMP, t[L].val, cmp); }
                else { gen_fwd(L); Src T = push_temp(); emit(KF_STORE, T); gen_fwd(R); emit(KF_CMP, T, cmp); pop_temp(); }
                break; }
            default: err = "unknown op in gen_fwd";
        }
        persist(k);
    }

    // ---- reverse: acc holds adj(k) on entry ----
    void gen_rev(int k) {
        TNode& n = t[k];
        if (n.op == KTN_OP_VAR) {
            emit(jseen[n.uslot] ? KR_JACC : KR_JSET, Src{KTN_K_NONE, (uint32_t)n.uslot});
            jseen[n.uslot] = 1; return;
        }
        if (!n.has_var) return;
        std::vector<int> ch = children(k), H;
        for (int c : ch) if (t[c].has_var) H.push_back(c);
        Src T{KTN_K_NONE, 0};
        if (H.size() > 1) { T = push_temp(); emit(KF_STORE, T); }
        if (n.op == KTN_OP_ADD && runs.count(k)) {
            const std::vector<Run>& rs = runs[k];
            size_t ri = 0; bool acc_is_adj = true;
            for (size_t i = 0; i < ch.size();) {
                if (ri < rs.size() && rs[ri].pos == i) {
                    const Run& r = rs[ri++];
                    if (!acc_is_adj) { emit(KF_LOAD, T); acc_is_adj = true; }
                    const int cs = term_cstride(r.tk);
                    for (uint32_t t0 = 0; t0 < r.n;) {      // split where set / accumulate changes
                        const bool seen = jseen[r.u0 + t0];
                        uint32_t t1 = t0 + 1;
                        while (t1 < r.n && (bool)jseen[r.u0 + t1] == seen) ++t1;
                        emit_terms(KR_TERMS, (uint8_t)(r.tk | (seen ? KTN_TF_JACC : 0) | (alias && r.tk == KTN_T_EXP_AFF ? KTN_TF_SAVEBLOB : 0)), t1 - t0, r.c0 + cs * t0, r.u0 + t0, r.s0 + t0);
                        for (uint32_t q = t0; q < t1; ++q) jseen[r.u0 + q] = 1;
                        t0 = t1;
                    }
                    i += r.n;
                } else {
                    if (t[ch[i]].has_var) {
                        if (!acc_is_adj) emit(KF_LOAD, T);
                        gen_rev(ch[i]); acc_is_adj = false;
                    }
                    ++i;
                }
            }
            if (H.size() > 1) pop_temp();
            return;
        }
        for (size_t j = 0; j < H.size(); ++j) {
            int c = H[j];
            if (j > 0) emit(KF_LOAD, T);
            size_t ci = 0; while (ch[ci] != c) ++ci;
            switch (n.op) {
                case KTN_OP_ADD: break;
                case KTN_OP_SUB: if (ci == 1) emit(KR_NEG); break;
                case KTN_OP_NEG: emit(KR_NEG); break;
                case KTN_OP_MUL:
                    if (n.nc == 2) emit(KR_MUL, t[ch[1 - ci]].val);
                    else if (n.nc > 2) emit(KR_MUL, Src{KTN_K_S, (uint32_t)t[c].pslot});
                    break;
                case KTN_OP_DIV: emit(KR_MUL, Src{KTN_K_S, (uint32_t)t[c].pslot}); break;
                case KTN_OP_POW:
                    if (t[ch[1]].expclass == 2) emit(KR_MUL2, t[ch[0]].val);
                    else if (t[ch[1]].expclass == 1) {}
                    else emit(KR_MUL, Src{KTN_K_S, (uint32_t)t[c].pslot});
                    break;
                case KTN_OP_EXP: emit(KR_MUL, n.val); break;
                case KTN_OP_LOG: emit(KR_MULRCP, t[c].val); break;
                case KTN_OP_SQRT: emit(KR_MULHRCP, n.val); break;
                case KTN_OP_ABS: emit(KR_MULSGN, t[c].val); break;
                case KTN_OP_SIN: emit(KR_MULCOS, t[c].val); break;
                case KTN_OP_COS: emit(KR_MULNSIN, t[c].val); break;
                case KTN_OP_IFELSE:
                    if (ci == 0) emit(KR_MULZERO); else emit(ci == 1 ? KR_MULEQ1 : KR_MULNE1, t[ch[0]].val);
                    break;
                case KTN_OP_LE: case KTN_OP_LT: case KTN_OP_GE: case KTN_OP_GT: case KTN_OP_EQ: emit(KR_MULZERO); break;
                default: err = "unknown op in gen_rev";
            }
            gen_rev(c);
        }
        if (H.size() > 1) pop_temp();
    }
};

static uint64_t fnv1a(const uint8_t* p, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}
static void put32(std::vector<uint8_t>& v, uint32_t x) { for (int i = 0; i < 4; ++i) v.push_back((uint8_t)(x >> (8 * i))); }
static uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

}  // namespace

void KtnProblem::reset(int64_t nvar, int64_t nconstr) {
    *this = KtnProblem();
    num_var = nvar; num_constr = nconstr;
    lb.reserve(nconstr); ub.reserve(nconstr); flags.reserve(nconstr);
    jac_ptr.assign(1, 0);
    row_const_off.assign(1, 0); row_col_off.assign(1, 0);
}

int KtnProblem::add_rows(int64_t first_row, int64_t nrows, const int64_t* eptr, const int32_t* op, const int32_t* arg,
                         const double* val, const double* lbv, const double* ubv, const uint8_t* fl) {
    if (first_row != rows_loaded || first_row + nrows > num_constr) { err = "rows must be added in ascending order"; return KTN_ERR_USAGE; }
    std::vector<TNode> t;
    std::vector<uint8_t> sig;
    std::vector<std::pair<int32_t, int32_t>> occ;  // (col, occurrence order)
    std::vector<int> stk_node, stk_rem;
    char msg[256];
    for (int64_t r = 0; r < nrows; ++r) {
        const int64_t row = first_row + r, b = eptr[r], n = eptr[r + 1] - b;
        if (n <= 0) { snprintf(msg, sizeof msg, "row %lld has an empty expression", (long long)row); err = msg; return KTN_ERR_USAGE; }
        if (n > 0x7fffff00) { err = "expression too long"; return KTN_ERR_USAGE; }
        t.assign((size_t)n, TNode());
        stk_node.clear(); stk_rem.clear();
        uint32_t nconst_wire = 0;
        for (int64_t k = 0; k < n; ++k) {
            TNode& nd = t[k];
            nd.op = op[b + k];
            int a = arity(nd.op);
            if (a == -2) { snprintf(msg, sizeof msg, "row %lld: unknown op %d", (long long)row, nd.op); err = msg; return KTN_ERR_USAGE; }
            nd.nc = a == 0 ? 0 : arg[b + k];
            if ((a > 0 && nd.nc != a) || (a == -1 && nd.nc < 1)) { snprintf(msg, sizeof msg, "row %lld: op %d has %d children", (long long)row, nd.op, nd.nc); err = msg; return KTN_ERR_USAGE; }
            nd.is_leaf = a == 0;
            if (nd.op == KTN_OP_CONST) ++nconst_wire;
            if (nd.op == KTN_OP_VAR && (arg[b + k] < 0 || arg[b + k] >= num_var)) { snprintf(msg, sizeof msg, "row %lld: variable index out of range", (long long)row); err = msg; return KTN_ERR_USAGE; }
            if (stk_node.empty()) { if (k != 0) { snprintf(msg, sizeof msg, "row %lld: more than one root", (long long)row); err = msg; return KTN_ERR_USAGE; } nd.parent = -1; }
            else { nd.parent = stk_node.back(); if (--stk_rem.back() == 0) { stk_node.pop_back(); stk_rem.pop_back(); } }
            if (nd.nc > 0) { stk_node.push_back((int)k); stk_rem.push_back(nd.nc); }
        }
        if (!stk_node.empty()) { snprintf(msg, sizeof msg, "row %lld: truncated expression", (long long)row); err = msg; return KTN_ERR_USAGE; }
        for (int64_t k = n - 1; k >= 0; --k) { int c = (int)k + 1; for (int i = 0; i < t[k].nc; ++i) c = t[c].send; t[k].send = c; }
        // structural exponent constants
        for (int64_t k = 0; k < n; ++k) if (t[k].op == KTN_OP_POW) {
            int e = t[k + 1].send;
            if (t[e].op == KTN_OP_CONST) { double v = val[b + e]; t[e].expclass = v == 2.0 ? 2 : v == 1.0 ? 1 : 0; }
        }
        // unique variable slots in first-occurrence order
        occ.clear();
        for (int64_t k = 0; k < n; ++k) if (t[k].op == KTN_OP_VAR) occ.emplace_back(arg[b + k], (int32_t)k);
        std::vector<std::pair<int32_t, int32_t>> byc = occ;
        std::sort(byc.begin(), byc.end());
        std::vector<std::pair<int32_t, int32_t>> firsts;  // (first node, col)
        for (size_t i = 0; i < byc.size(); ++i) if (i == 0 || byc[i].first != byc[i - 1].first) firsts.emplace_back(byc[i].second, byc[i].first);
        std::sort(firsts.begin(), firsts.end());
        const uint32_t nu = (uint32_t)firsts.size();
        // col -> slot via binary search on a sorted (col, slot) list
        std::vector<std::pair<int32_t, int32_t>> col2slot(nu);
        for (uint32_t s = 0; s < nu; ++s) col2slot[s] = {firsts[s].second, (int32_t)s};
        std::sort(col2slot.begin(), col2slot.end());
        uint32_t ncst = 0;
        for (int64_t k = 0; k < n; ++k) {
            if (t[k].op == KTN_OP_VAR) {
                auto it = std::lower_bound(col2slot.begin(), col2slot.end(), std::make_pair(arg[b + k], (int32_t)-1));
                t[k].uslot = it->second; t[k].val = Src{KTN_K_S, (uint32_t)t[k].uslot}; t[k].has_val = true;
            } else if (t[k].op == KTN_OP_CONST && t[k].expclass == 0) {
                t[k].cslot = (int)ncst++; t[k].val = Src{KTN_K_C, (uint32_t)t[k].cslot}; t[k].has_val = true;
            }
        }
        // signature
        sig.clear();
        sig.push_back(fl[r] & (KTN_ROW_NL | KTN_ROW_DENSE));
        for (int64_t k = 0; k < n; ++k) {
            sig.push_back((uint8_t)t[k].op);
            if (t[k].op == KTN_OP_VAR) put32(sig, (uint32_t)t[k].uslot);
            else if (t[k].op == KTN_OP_CONST) sig.push_back((uint8_t)t[k].expclass);
            else put32(sig, (uint32_t)t[k].nc);
        }
        uint64_t h = fnv1a(sig.data(), sig.size());
        uint32_t sid = UINT32_MAX;
        auto& cand = shape_by_hash[h];
        for (uint32_t s : cand) if (shape_sig[s] == sig) { sid = s; break; }
        if (sid == UINT32_MAX) {
            const bool dense = (fl[r] & KTN_ROW_DENSE) != 0;
            KtnShapeDesc sd; memset(&sd, 0, sizeof sd);
            std::vector<KtnIns> code;
            // pass 0: plain layout.  pass 1 (if eligible): J accumulators / exp saves aliased into dead constants.
            for (int pass = 0; pass < 2; ++pass) {
                std::vector<TNode> tc = t;
                ShapeCompiler sc(tc, nu, pass == 1);
                sc.analyse();
                sc.gen_fwd(0);
                const uint32_t nfwd = (uint32_t)sc.code.size();
                sc.temp_depth = 0;
                sc.emit(KR_ONE);
                sc.gen_rev(0);
                sc.emit(K_END);
                if (!sc.err.empty()) { err = sc.err; return KTN_ERR_USAGE; }
                sd.n_fwd = nfwd; sd.n_ins = (uint32_t)sc.code.size();
                sd.n_uniq = nu; sd.n_const = ncst; sd.n_scratch = sc.next_slot;
                sd.flags = ((fl[r] & KTN_ROW_NL) ? KTN_SH_NL : 0) | (dense ? (KTN_SH_DENSE | KTN_SH_BIG) : 0);
                sd.order_bytes = nu <= 256 ? 1 : nu <= 65536 ? 2 : 4;
                code = sc.code;
                if (pass == 1) break;
                sd.j_base = nu; sd.j_stride = 1; sd.j_in_blob = 0;
                // eligibility: shared-memory staged shape whose every Jacobian entry is first written by ONE fused
                // reverse run (set, not accumulate) over u = 0..nu-1 whose terms own at least one constant.
                if (dense || nu == 0 || ktn_shape_lane_bytes(sd) > lane_limit_hint) break;
                bool ok = true; int found = -1;
                for (size_t q = nfwd; q < code.size() && ok; ++q) {
                    const KtnIns& in = code[q];
                    if (in.op == KR_JSET) ok = false;
                    if (in.op == KR_TERMS && !(in.kind & KTN_TF_JACC)) {
                        const int tk = in.kind & 0xf, cs = ShapeCompiler::term_cstride(tk);
                        if (found >= 0 || cs == 0 || in.a != 0 || in.n != nu) ok = false; else found = (int)q;
                    }
                }
                if (!ok || found < 0) break;
                const int cs = ShapeCompiler::term_cstride(code[found].kind & 0xf);
                sd.j_in_blob = 1; sd.j_base = code[found].idx + cs - 1; sd.j_stride = cs;
            }
            // family detection: exact match of the program against the canonical fused forms
            {
                auto is = [&](size_t q, uint8_t op, uint8_t kind, uint32_t n, uint32_t idx, uint32_t a) {
                    return q < code.size() && code[q].op == op && code[q].kind == kind && code[q].n == n && code[q].idx == idx && code[q].a == a; };
                sd.family = KTN_FAM_GENERIC;
                const bool no_family = getenv("KTN_NO_FAMILY") != nullptr;      // A/B switch: every shape takes the interpreter
                if (!no_family && sd.j_in_blob && (sd.flags & KTN_SH_NL) && nu >= 1 && nu <= 256 && sd.n_const == 2 * nu) {
                    const uint8_t EA = KTN_T_EXP_AFF | KTN_TF_SAVEBLOB;
                    if (code.size() == 8 && sd.j_base == 1 && sd.j_stride == 2 && is(0, KF_TERMS, EA | KTN_TF_FIRST, nu, 0, 0) && code[1].op == KF_STORE && code[1].kind == KTN_K_S &&
                        code[2].op == KF_LOG && code[3].op == KR_ONE && code[4].op == KR_MULRCP && code[4].kind == KTN_K_S && code[4].idx == code[1].idx &&
                        code[5].op == KF_STORE && code[5].kind == KTN_K_R1 && is(6, KR_TERMS, EA, nu, 0, 0) && code[7].op == K_END)
                        sd.family = KTN_FAM_LSE;
                    if (code.size() == 7 && sd.j_base == 0 && sd.j_stride == 1 && is(0, KF_TERMS, KTN_T_MULC_SQ | KTN_TF_FIRST, nu, 0, 0) && is(1, KF_TERMS, KTN_T_MULC_X, nu, nu, 0) &&
                        code[2].op == KR_ONE && code[3].op == KF_STORE && code[3].kind == KTN_K_R1 && is(4, KR_TERMS, KTN_T_MULC_SQ, nu, 0, 0) &&
                        is(5, KR_TERMS, KTN_T_MULC_X | KTN_TF_JACC, nu, nu, 0) && code[6].op == K_END)
                        sd.family = KTN_FAM_QUAD;
                }
                // sqrt(sum_{u < nu-1} (s_u x_u)^2) - x_{nu-1}: the linear variable is the row's LAST unique variable and none of the squared ones
                if (!no_family && !sd.j_in_blob && (sd.flags & KTN_SH_NL) && nu >= 2 && nu <= 256 && sd.n_const == nu - 1 && code.size() == 13 &&
                    is(0, KF_TERMS, KTN_T_SQ_MULC | KTN_TF_FIRST, nu - 1, 0, 0) && code[1].op == KF_SQRT && code[2].op == KF_STORE && code[2].kind == KTN_K_S &&
                    code[3].op == KF_SUB && code[3].kind == KTN_K_S && code[3].idx == nu - 1 && code[4].op == KR_ONE && code[5].op == KF_STORE && code[5].kind == KTN_K_R1 &&
                    code[6].op == KR_MULHRCP && code[6].kind == KTN_K_S && code[6].idx == code[2].idx && code[7].op == KF_STORE && code[7].kind == KTN_K_R2 &&
                    is(8, KR_TERMS, KTN_T_SQ_MULC, nu - 1, 0, 0) && code[9].op == KF_LOAD && code[9].kind == KTN_K_R1 && code[10].op == KR_NEG &&
                    code[11].op == KR_JSET && code[11].idx == nu - 1 && code[12].op == K_END)
                    sd.family = KTN_FAM_SOC;
            }
            sd.prog_off = (uint32_t)prog.size();
            prog.insert(prog.end(), code.begin(), code.end());
            sid = (uint32_t)shapes.size();
            shapes.push_back(sd); shape_sig.push_back(sig); cand.push_back(sid);
        }
        // per-row data
        row_shape.push_back(sid);
        row_nconst_wire.push_back(nconst_wire);
        for (int64_t k = 0; k < n; ++k) if (t[k].cslot >= 0) rd_const.push_back(val[b + k]);
        row_const_off.push_back(rd_const.size());
        for (uint32_t s = 0; s < nu; ++s) rd_col.push_back(firsts[s].second);
        for (uint32_t p = 0; p < nu; ++p) rd_order.push_back((uint32_t)col2slot[p].second);
        row_col_off.push_back(rd_col.size());
        lb.push_back(lbv[r]); ub.push_back(ubv[r]); flags.push_back(fl[r]);
        if (fl[r] & KTN_ROW_DENSE) { for (int64_t c = 0; c < num_var; ++c) jac_col.push_back((int32_t)c); }
        else for (uint32_t p = 0; p < nu; ++p) jac_col.push_back(col2slot[p].first);
        jac_ptr.push_back((int64_t)jac_col.size());
    }
    rows_loaded += nrows;
    return KTN_OK;
}

void KtnProblem::repack_bounds() {
    chunk_lb.assign(chunk_rows.size(), 0.0); chunk_ub.assign(chunk_rows.size(), 0.0);
    for (size_t i = 0; i < chunk_rows.size(); ++i) if (chunk_rows[i] >= 0) { chunk_lb[i] = lb[chunk_rows[i]]; chunk_ub[i] = ub[chunk_rows[i]]; }
}

int KtnProblem::finalize(int64_t sigma, uint32_t lane_limit, size_t table_limit) {
    if (rows_loaded != num_constr) { err = "not all rows were loaded"; return KTN_ERR_USAGE; }
    if (sigma < 32) sigma = 32;
    lane_limit_hint = lane_limit;
    // classify shapes
    max_lane_bytes = 0;
    for (auto& s : shapes) {
        if (ktn_shape_lane_bytes(s) > lane_limit) s.flags |= KTN_SH_BIG;
    }
    // the regular kernel keeps every shape descriptor and the regular shapes' programs in shared memory; if that
    // table would not fit, every shape takes the global-memory kernel instead (correct, slower)
    {
        size_t tb = (shapes.size() * sizeof(KtnShapeDesc) + 15) & ~(size_t)15;
        for (auto& s : shapes) if (!(s.flags & KTN_SH_BIG)) tb += (size_t)s.n_ins * sizeof(KtnIns);
        if (tb > table_limit) for (auto& s : shapes) s.flags |= KTN_SH_BIG;
    }
    // a shape the global-scratch kernel runs is interpreted: it is no family shape (its chunks carry the interpreter's sort order)
    for (auto& s : shapes) if (s.flags & KTN_SH_BIG) s.family = KTN_FAM_GENERIC;
    for (auto& s : shapes) if (!(s.flags & KTN_SH_BIG)) max_lane_bytes = std::max(max_lane_bytes, ktn_shape_lane_bytes(s));
    chunks.clear(); blob.clear(); chunk_rows.clear(); big_scratch_doubles = 0;
    std::vector<std::pair<uint32_t, int32_t>> win;  // (shape, row)
    alg_bytes_static = 8 * num_var;
    for (int64_t i = 0; i < num_constr; ++i) if (flags[i] & KTN_ROW_NL)
        alg_bytes_static += 4 * (jac_ptr[i + 1] - jac_ptr[i]) + 8 * (int64_t)shapes[row_shape[i]].n_const + 16;   // C_i = per-row fp64 constants

    auto pack_chunk = [&](uint32_t sid, const int32_t* rows, int nr, bool isbig) {
        const KtnShapeDesc& s = shapes[sid];
        KtnChunkDesc cd; memset(&cd, 0, sizeof cd);
        const uint32_t L = (isbig && nr == 1) ? 1u : 32u;
        cd.shape = sid; cd.nrows = (uint16_t)nr; cd.stride = (uint16_t)L; cd.aux = s.n_uniq;
        uint64_t off = align_up(blob.size(), 128);
        uint64_t sec_c = 0, sec_col = align_up(sec_c + 8ull * s.n_const * L, 16), sec_ord = align_up(sec_col + 4ull * s.n_uniq * L, 16);
        const bool rankword = s.family != KTN_FAM_GENERIC && s.n_uniq <= KTN_FAM_REGS;      // family rows of <= 16 unique variables: one packed word per row
        uint64_t bytes = align_up(sec_ord + (rankword ? 8ull * L : (uint64_t)s.order_bytes * s.n_uniq * L), 16);
        if (rankword) bytes = KTN_FAM_BLOB_BYTES(s.n_uniq);      // grouped layout of the family classes (ktn_program.h)
        cd.blob_off = off; cd.blob_bytes = (uint32_t)bytes;
        blob.resize(off + bytes, 0);
        uint8_t* base = blob.data() + off;
        for (uint32_t lane = 0; lane < L; ++lane) {
            int32_t row = rows[lane < (uint32_t)nr ? lane : 0];  // idle lanes replay lane 0's (valid) data
            const double* rc = rd_const.data() + row_const_off[row];
            const int32_t* rcol = rd_col.data() + row_col_off[row];
            const uint32_t* rord = rd_order.data() + row_col_off[row];
            if (rankword) {      // family rows of <= 16 unique variables
                for (uint32_t u = 0; u < s.n_uniq; ++u) {
                    double* pr = (double*)(base + KTN_FAM_PAIR_AT(u, lane));
                    const uint32_t s0 = ktn_family_slot(s.family, 0, u, s.n_uniq), s1 = ktn_family_slot(s.family, 1, u, s.n_uniq);
                    pr[0] = s0 == KTN_FAM_NOSLOT ? 0.0 : rc[s0]; pr[1] = s1 == KTN_FAM_NOSLOT ? 0.0 : rc[s1];
                    *(int32_t*)(base + KTN_FAM_COL_AT(s.n_uniq, u, lane)) = rcol[u];
                }
                uint64_t w = 0; for (uint32_t p = 0; p < s.n_uniq; ++p) w |= (uint64_t)p << (4 * rord[p]);      // rank word: 4 bits per unique variable = its Jacobian entry index
                ((uint64_t*)(base + KTN_FAM_ORD_OFF(s.n_uniq)))[lane] = w;
                continue;
            }
            double* dc = (double*)(base + sec_c);
            for (uint32_t c = 0; c < s.n_const; ++c) dc[(uint64_t)c * L + lane] = rc[c];
            int32_t* dcol = (int32_t*)(base + sec_col);
            for (uint32_t u = 0; u < s.n_uniq; ++u) dcol[(uint64_t)u * L + lane] = rcol[u];
            uint8_t* dord = base + sec_ord;
            // long family rows carry the inverse permutation: rank[u] = Jacobian entry index of unique variable u
            if (s.family != KTN_FAM_GENERIC) { for (uint32_t p = 0; p < s.n_uniq; ++p) dord[(uint64_t)rord[p] * L + lane] = (uint8_t)p; continue; }
            for (uint32_t p = 0; p < s.n_uniq; ++p) {
                uint64_t e = (uint64_t)p * L + lane;
                if (s.order_bytes == 1) dord[e] = (uint8_t)rord[p];
                else if (s.order_bytes == 2) ((uint16_t*)dord)[e] = (uint16_t)rord[p];
                else ((uint32_t*)dord)[e] = rord[p];
            }
        }
        return cd;
    };

    // 1. form the chunks: inside every window of sigma consecutive rows, rows of one shape go together, 32 per chunk
    struct Pending { uint32_t sid; int32_t nr; uint32_t orig; int32_t rows[32]; };
    std::vector<Pending> reg, bigp;
    for (int64_t w0 = 0; w0 < num_constr; w0 += sigma) {
        int64_t w1 = std::min(num_constr, w0 + sigma);
        win.clear();
        for (int64_t i = w0; i < w1; ++i) win.emplace_back(row_shape[i], (int32_t)i);
        std::stable_sort(win.begin(), win.end(), [](const std::pair<uint32_t, int32_t>& a, const std::pair<uint32_t, int32_t>& b) { return a.first < b.first; });
        size_t i = 0;
        while (i < win.size()) {
            uint32_t sid = win[i].first; size_t j = i;
            while (j < win.size() && win[j].first == sid) ++j;
            const bool isbig = shapes[sid].flags & KTN_SH_BIG;
            for (size_t c0 = i; c0 < j; c0 += 32) {
                Pending pc; pc.sid = sid; pc.nr = (int32_t)std::min<size_t>(32, j - c0); pc.orig = (uint32_t)(reg.size() + bigp.size());
                for (int q = 0; q < 32; ++q) pc.rows[q] = q < pc.nr ? win[c0 + q].second : -1;
                (isbig ? bigp : reg).push_back(pc);
            }
            i = j;
        }
    }
    // 2. regular chunks are ordered by (family, class), window order kept inside a class, and packed in that order: the blobs
    //    of one family class are contiguous and equally sized, so the family kernel finds a chunk without a descriptor
    {
        auto key = [&](const Pending& pc) { const KtnShapeDesc& s = shapes[pc.sid]; return s.family * (uint32_t)KTN_FAM_NCLS + (s.family != KTN_FAM_GENERIC ? ktn_family_class(s.n_uniq) : ((s.flags & KTN_SH_NL) ? 0u : 1u)); };      // interpreted shapes: class 0 = rows of nlconstr_ixs, class 1 = the rest (never tested: a separation round skips them)
        std::stable_sort(reg.begin(), reg.end(), [&](const Pending& a, const Pending& b) { return key(a) < key(b); });
        std::vector<uint32_t> count((size_t)KTN_FAM__COUNT * KTN_FAM_NCLS + 1, 0u);
        memset(cls_blob_off, 0, sizeof cls_blob_off); memset(cls_blob_stride, 0, sizeof cls_blob_stride);
        for (size_t i = 0; i < reg.size(); ++i) {
            const Pending& pc = reg[i];
            KtnChunkDesc cd = pack_chunk(pc.sid, pc.rows, pc.nr, false);
            cd.row_slot = (uint32_t)chunk_rows.size();
            for (int q = 0; q < 32; ++q) chunk_rows.push_back(pc.rows[q]);
            const uint32_t k = key(pc);
            if (count[k]++ == 0) cls_blob_off[k / KTN_FAM_NCLS][k % KTN_FAM_NCLS] = cd.blob_off;
            else if (count[k] == 2) cls_blob_stride[k / KTN_FAM_NCLS][k % KTN_FAM_NCLS] = (uint32_t)(cd.blob_off - chunks.back().blob_off);
            chunks.push_back(cd);
        }
        uint32_t at = 0;
        for (int f = 0; f < KTN_FAM__COUNT; ++f) {
            fam_begin[f] = at;
            for (int k = 0; k < KTN_FAM_NCLS; ++k) { cls_begin[f][k] = at; at += count[(size_t)f * KTN_FAM_NCLS + k]; }
            cls_begin[f][KTN_FAM_NCLS] = at;
        }
        fam_begin[KTN_FAM__COUNT] = at;
    }
    n_regular_chunks = (uint32_t)chunks.size();
    for (int f = 1; f < KTN_FAM__COUNT; ++f) for (uint32_t k = 1; k < KTN_FAM_NCLS; ++k)      // the family kernels compute a class's blob addresses
        if (cls_begin[f][k + 1] - cls_begin[f][k] >= 2 && cls_blob_stride[f][k] != KTN_FAM_BLOB_BYTES(k)) { err = "family blob stride"; return KTN_ERR_USAGE; }
    for (size_t c = 0; c < bigp.size(); ++c) {
        const Pending& pc = bigp[c];
        KtnChunkDesc cd = pack_chunk(pc.sid, pc.rows, pc.nr, true);
        cd.row_slot = (uint32_t)chunk_rows.size();
        for (int q = 0; q < 32; ++q) chunk_rows.push_back(pc.rows[q]);
        cd.aux = big_scratch_doubles;
        big_scratch_doubles += (uint64_t)shapes[cd.shape].n_scratch * cd.stride;
        chunks.push_back(cd);
    }
    blob.resize(align_up(blob.size(), 128) + 128, 0);
    row_slot.assign((size_t)num_constr, -1);
    chunk_jp.assign(chunk_rows.size(), 0u);
    for (size_t i = 0; i < chunk_rows.size(); ++i) if (chunk_rows[i] >= 0) chunk_jp[i] = (uint32_t)jac_ptr[chunk_rows[i]];
    for (size_t i = 0; i < chunk_rows.size(); ++i) if (chunk_rows[i] >= 0) row_slot[chunk_rows[i]] = (int32_t)i;
    repack_bounds();
    // the ragged per-row staging is no longer needed
    std::vector<double>().swap(rd_const); std::vector<int32_t>().swap(rd_col); std::vector<uint32_t>().swap(rd_order);
    return KTN_OK;
}
