// ktn_program.h -- the tape ISA shared by the host compiler (ktn_compile.cpp) and the
// sm_100a interpreter (ktn_kernels.cu).
//
// A SHAPE is the structure of a constraint expression with constants and variable
// indices abstracted away (reference: one ReverseDiffSparse node tape per constraint,
// interpreted node by node at src/separators.jl:112-113).  All rows of a shape share one
// straight-line PROGRAM for a per-lane accumulator machine; per-row data is packed
// structure-of-arrays in CHUNKS of up to 32 rows so a warp runs 32 rows in lock step.
//
// Per-lane machine state
//   acc          forward accumulator / reverse adjoint
//   aux          second operand / by-product (reciprocal of a division)
//   R1, R2       two temporaries (running sums, saved adjoints)
//   C[c]         per-row constants               (chunk blob, read-only)
//   S[s]         scratch slots: S[0..nu) = gathered x* values XV, S[nu..2nu) = Jacobian
//                accumulators J, then saved forward values / partials / spilled temporaries
// Arithmetic order is exactly the oracle's (oracle/ktn_oracle.c); no op is fused.
#ifndef KTN_PROGRAM_H
#define KTN_PROGRAM_H
#include <stdint.h>

struct KtnIns {     // 16 bytes, fetched with one 128-bit load
    uint8_t op;
    uint8_t kind;   // operand kind (KTN_K_*); fused term ops: term kind | flags
    uint16_t n;     // small immediate (skip count, term count)
    uint32_t idx;   // operand index; fused term ops: first constant slot
    uint32_t a;     // fused term ops: first unique-variable slot
    uint32_t b;     // fused term ops: first save slot
};

enum { KTN_K_NONE = 0, KTN_K_C = 1, KTN_K_S = 2, KTN_K_R1 = 3, KTN_K_R2 = 4 };

enum {
    // ---- forward: acc machine ----
    KF_LOAD = 0,   // acc = src
    KF_ADD,        // acc = acc + src
    KF_ADDZ,       // acc = 0.0 + acc            (n-ary sum starts from zero(T))
    KF_SUB,        // acc = acc - src
    KF_RSUB,       // acc = src - acc
    KF_MUL,        // acc = acc * src
    KF_DIV,        // aux = 1/src;  acc = acc * aux      (numerator in acc)
    KF_RDIV,       // aux = 1/acc;  acc = src * aux      (denominator in acc)
    KF_FDIV,       // acc = acc / src                   (true division: n-ary product partials)
    KF_POW2,       // acc = acc * acc
    KF_NEG, KF_EXP, KF_LOG, KF_SQRT, KF_ABS, KF_SIN, KF_COS,
    KF_STORE,      // dst(kind, idx) = acc      kind in {S, R1, R2}
    KF_LDAUX,      // aux = src
    KF_STAUX,      // S[idx] = aux
    KF_DENP,       // S[idx] = (-acc) * aux             (d/d denominator after KF_DIV / KF_RDIV)
    KF_POWG,       // acc = pow(acc, aux)
    KF_POWPB,      // S[idx] = aux * pow(acc, aux - 1)  (d/d base; before KF_POWG)
    KF_POWPE,      // S[idx] = pow(acc, aux) * log(acc) (d/d exponent; before KF_POWG)
    KF_SELZ,       // acc = (src == 0) ? aux : acc
    KF_SEL1,       // acc = (src == 1) ? acc : aux          (ifelse: then-branch in acc, else-branch in aux)
    KF_CMP,        // acc = (src CMP acc) ? 1.0 : 0.0       (n: 0 <=, 1 <, 2 >=, 3 >, 4 ==; src = left operand)
    KF_SKIPNZ,     // if no lane of the warp has src == 0: skip the next n instructions
    // ---- reverse: acc holds the adjoint; revmul(a, p) = (a == 0 && !finite(p)) ? a : a * p ----
    KR_ONE,        // acc = 1.0
    KR_MUL,        // acc = revmul(acc, src)
    KR_NEG,        // acc = revmul(acc, -1.0) = -acc
    KR_MUL2,       // acc = revmul(acc, 2.0 * src)
    KR_MULRCP,     // acc = revmul(acc, 1.0 / src)
    KR_MULHRCP,    // acc = revmul(acc, 0.5 / src)
    KR_MULSGN,     // acc = revmul(acc, src >= 0 ? 1.0 : -1.0)
    KR_MULCOS,     // acc = revmul(acc, cos(src))                (d sin)
    KR_MULNSIN,    // acc = revmul(acc, -sin(src))               (d cos)
    KR_MULZERO,    // acc = revmul(acc, 0.0)                     (conditions and comparison operands)
    KR_MULEQ1,     // acc = revmul(acc, src == 1 ? 1.0 : 0.0)    (ifelse, then-branch)
    KR_MULNE1,     // acc = revmul(acc, src == 1 ? 0.0 : 1.0)    (ifelse, else-branch)
    KR_JSET,       // J[idx] = 0.0 + acc
    KR_JACC,       // J[idx] = J[idx] + acc
    // ---- fused runs of identical terms inside an n-ary sum: the same arithmetic as the primitive
    //      ops above in the same order, executed by one hand-written loop instead of 4-6 dispatches
    //      per term.  Term t of the run uses constants idx + t*cstride.., variable slot a + t, save slot b + t.
    KF_TERMS,      // acc = (FIRST ? 0.0 : acc) + term_0 + term_1 + ...
    KR_TERMS,      // J[a + t] (=|+=) d term_t / d x * acc        (acc = adjoint of the sum, unchanged)
    K_END,
    K__COUNT
};

// fused term kinds (low nibble of KtnIns.kind) and flags (high nibble)
enum {
    KTN_T_X = 0,        // x
    KTN_T_MULC_X,       // c * x
    KTN_T_SQ,           // x ^ 2
    KTN_T_MULC_SQ,      // c * x ^ 2
    KTN_T_SQ_MULC,      // (c * x) ^ 2
    KTN_T_EXP_AFF,      // exp(c * x + d)         saves exp value in S[b + t]
    KTN_T__COUNT
};
enum { KTN_TF_FIRST = 0x10, KTN_TF_JACC = 0x20, KTN_TF_SAVEBLOB = 0x40 };   // SAVEBLOB: exp values are kept in the term's dead `d` constant slot

// shape families: program patterns with a hand-written, interpreter-free fast path in ktn_kernels.cu.
// The fast path performs the same arithmetic as the shape's program (which stays the definition and is what
// the host emulator runs); everything else takes the generic interpreter.
enum {
    KTN_FAM_GENERIC = 0,
    KTN_FAM_LSE,        // log(sum_u exp(c_u x_u + d_u))                     (BASELINE.json configs[2])
    KTN_FAM_QUAD,       // sum_u a_u x_u^2 + sum_u b_u x_u                  (BASELINE.json configs[1])
    KTN_FAM_SOC,        // sqrt(sum_{u<nu-1} (s_u x_u)^2) - x_{nu-1}         (BASELINE.json configs[3], test/3d.jl:161)
    KTN_FAM__COUNT
};

// Family chunks are further split into CLASSES by unique-variable count: class k (1 <= k <= KTN_FAM_REGS) = rows of exactly
// k unique variables, which run a code path specialised for k with the whole row in registers; class 0 = more than that
// (streaming fallback).  Chunks are sorted by (family, class) so that every class is one contiguous chunk range, and the
// blobs of a class k >= 1 are contiguous and KTN_FAM_BLOB_BYTES(k) apart (constants | columns | one order word per row).
#define KTN_FAM_REGS 16
#define KTN_FAM_NCLS (KTN_FAM_REGS + 1)
static inline uint32_t ktn_family_class(uint32_t n_uniq) { return n_uniq <= KTN_FAM_REGS ? n_uniq : 0u; }
// Blob of a class k >= 1 chunk (lane stride 32 bytes: one 256-bit load per lane and group, and a SELECTED row's data sits in few
// 32-byte sectors -- the compaction kernel reads it back sector by sector from DRAM, where every sector is a row activation):
//   pair groups  [(k + 1) / 2][32 lanes] x 32 bytes: (p0, p1) of unique variables 2g and 2g + 1 (ktn_family_slot; zero padding)
//   col groups   [(k + 7) / 8][32 lanes] x 32 bytes: columns of unique variables 8g .. 8g + 7 (zero padding)
//   rank         [32 lanes] x 8 bytes: 4 bits per unique variable = its Jacobian entry index
#define KTN_FAM_PGROUPS(k) (((k) + 1u) / 2u)
#define KTN_FAM_CGROUPS(k) (((k) + 7u) / 8u)
#define KTN_FAM_COL_OFF(k) (1024u * KTN_FAM_PGROUPS(k))
#define KTN_FAM_ORD_OFF(k) (1024u * (KTN_FAM_PGROUPS(k) + KTN_FAM_CGROUPS(k)))
#define KTN_FAM_BLOB_BYTES(k) (KTN_FAM_ORD_OFF(k) + 256u)
// byte offset (inside the chunk blob) of the pair / the column of unique variable u of the row in `lane`
#define KTN_FAM_PAIR_AT(u, lane) (((u) >> 1) * 1024u + (lane) * 32u + ((u) & 1u) * 16u)
#define KTN_FAM_COL_AT(k, u, lane) (KTN_FAM_COL_OFF(k) + ((u) >> 3) * 1024u + (lane) * 32u + ((u) & 7u) * 4u)
// constants per row of a family shape (rows of more than KTN_FAM_REGS variables keep the generic section layout: constants | columns | rank bytes)
#define KTN_FAM_NCONST(family, nu) ((family) == 3u /* KTN_FAM_SOC */ ? (nu) - 1u : 2u * (nu))
#define KTN_FAM_NOSLOT 0xffffffffu      // the pair's half is not backed by a constant of the row: packed as 0.0
static inline uint32_t ktn_family_slot(uint32_t family, uint32_t which, uint32_t u, uint32_t nu) {
    if (family == 1u /* KTN_FAM_LSE */) return 2u * u + which;                           // c_u = slot 2u, d_u = slot 2u + 1
    if (family == 2u /* KTN_FAM_QUAD */) return which ? nu + u : u;                      // a_u = slot u, b_u = slot nu + u
    return (which == 0u && u + 1u < nu) ? u : KTN_FAM_NOSLOT;                            // SOC: s_u = slot u; the linear variable has none
}

// shape flags
enum { KTN_SH_NL = 1, KTN_SH_DENSE = 2, KTN_SH_BIG = 4 };

struct KtnShapeDesc {
    uint32_t prog_off;     // first instruction in the program array
    uint32_t n_fwd;        // forward instructions (then the reverse part follows)
    uint32_t n_ins;        // total instructions
    uint32_t n_uniq;       // unique variables = Jacobian entries of a (non-dense) row
    uint32_t n_const;      // per-row constants
    uint32_t n_scratch;    // scratch slots (incl. XV and J)
    uint32_t flags;        // KTN_SH_*
    uint32_t order_bytes;  // bytes per `order` entry: 1 (n_uniq <= 256), 2 or 4
    // Jacobian accumulator of unique variable u lives at slot j_base + u * j_stride of the scratch
    // area, or -- when j_in_blob -- of the chunk's constant area (a constant that is dead by then).
    uint32_t j_base, j_stride, j_in_blob;
    uint32_t family;       // KTN_FAM_*: which hand-specialised instantiation of the round kernel runs this shape
    // blob sections of one chunk (lane stride L rows):
    //   consts: n_const * L doubles | cols: n_uniq * L int32 | order: n_uniq * L * order_bytes
};

struct KtnChunkDesc {
    uint64_t blob_off;     // byte offset of the chunk's blob (16-byte aligned)
    uint64_t aux;          // BIG chunks: offset in doubles into the global scratch arena; regular chunks: n_uniq of the shape
    uint32_t shape;
    uint32_t blob_bytes;   // multiple of 16
    uint16_t nrows;        // valid lanes
    uint16_t stride;       // lane stride L of the SoA sections (32 for regular chunks)
    uint32_t row_slot;     // index of the chunk's first lane in chunk_rows[]; always chunk index * 32
};

// per-row result word `sel` written by the round kernels: 0 = not selected, else nnz | flags
#define KTN_SEL_ERRBIT 0x40000000u      // a coefficient of the row's cut is not finite (set by the kernel that built the cut)
#define KTN_SEL_DEFER  0x20000000u      // family row: K1 only evaluated and tested it; the compaction kernel builds the cut (ktn_family_cut_entries)
#define KTN_SEL_NNZ(s) ((s) & 0x1fffffffu)

// Compaction blocks: K1 counts the selected rows of every block of KTN_CROWS consecutive rows
// (cuts << KTN_BLK_SHIFT | nnz, one 64-bit atomic per warp and block), K2 turns the counts into offsets.
#define KTN_CROWS_LOG2 11
#define KTN_CROWS (1 << KTN_CROWS_LOG2)
#define KTN_BLK_SHIFT 48
#define KTN_BLK_NNZ_MASK ((1ull << KTN_BLK_SHIFT) - 1ull)

#endif
