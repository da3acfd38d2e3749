// ktn_compile.h -- host tape compiler: expression wire format -> shapes, programs, packed chunks.
// Replaces the per-row bucketing of initialize! (reference src/separators.jl:92-100) and the
// tape construction JuMP does inside MathProgBase.initialize (src/separators.jl:88).
#ifndef KTN_COMPILE_H
#define KTN_COMPILE_H
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>
#include "ktn_program.h"

struct KtnProblem {
    int64_t num_var = 0, num_constr = 0, rows_loaded = 0;
    // row-order data
    std::vector<double> lb, ub;
    std::vector<uint8_t> flags;
    std::vector<int64_t> jac_ptr;      // num_constr + 1
    std::vector<int32_t> jac_col;      // ascending unique columns per row (0..n-1 for dense rows)
    std::vector<uint32_t> row_shape;
    std::vector<uint32_t> row_nconst_wire;  // CONST nodes on the wire (algorithmic-bytes formula)
    // ragged per-row extracted data (program operand order)
    std::vector<uint64_t> row_const_off, row_col_off;
    std::vector<double> rd_const;
    std::vector<int32_t> rd_col;       // unique columns in first-occurrence order
    std::vector<uint32_t> rd_order;    // sorted position p -> unique slot
    // shapes
    std::vector<KtnShapeDesc> shapes;
    std::vector<KtnIns> prog;
    std::vector<std::vector<uint8_t>> shape_sig;
    std::unordered_map<uint64_t, std::vector<uint32_t>> shape_by_hash;
    // packed (filled by finalize)
    std::vector<KtnChunkDesc> chunks;       // regular chunks first, then BIG chunks
    uint32_t n_regular_chunks = 0;
    uint32_t fam_begin[KTN_FAM__COUNT + 1] = {0};   // regular chunks of family f: [fam_begin[f], fam_begin[f+1])
    uint32_t cls_begin[KTN_FAM__COUNT][KTN_FAM_NCLS + 1] = {{0}};   // ... of class k inside family f: [cls_begin[f][k], cls_begin[f][k+1])
    uint64_t cls_blob_off[KTN_FAM__COUNT][KTN_FAM_NCLS] = {{0}};    // blob offset of the first chunk of the class
    uint32_t cls_blob_stride[KTN_FAM__COUNT][KTN_FAM_NCLS] = {{0}}; // bytes between consecutive chunk blobs of the class (classes >= 1)
    std::vector<uint8_t> blob;
    std::vector<int32_t> chunk_rows;        // chunk * 32 + lane -> row or -1
    std::vector<int32_t> row_slot;          // row -> chunk * 32 + lane (the inverse of chunk_rows)
    std::vector<uint32_t> chunk_jp;         // chunk * 32 + lane -> jac_ptr[row]
    std::vector<double> chunk_lb, chunk_ub; // same indexing
    uint64_t big_scratch_doubles = 0;       // global scratch arena for BIG chunks
    uint32_t max_lane_bytes = 0;            // per-lane shared-memory need of the largest regular shape
    int64_t alg_bytes_static = 0;           // sum_NL (4 nnz + 8 C + 16) + 8 n
    uint32_t lane_limit_hint = 1536;       // shapes above this never alias (they run from global memory)
    std::string err;

    void reset(int64_t nvar, int64_t nconstr);
    int add_rows(int64_t first_row, int64_t nrows, const int64_t* eptr, const int32_t* op, const int32_t* arg,
                 const double* val, const double* lb, const double* ub, const uint8_t* flags);
    // sigma: rows per sorting window; lane_limit: max per-lane shared-memory bytes of a regular shape
    int finalize(int64_t sigma, uint32_t lane_limit, size_t table_limit = 32768);
    void repack_bounds();                   // chunk_lb / chunk_ub from lb / ub
};

// per-lane shared-memory bytes a regular shape needs in the round kernel
static inline uint32_t ktn_shape_blob_lane_bytes(const KtnShapeDesc& s) {
    return 8u * s.n_const + 4u * s.n_uniq + s.order_bytes * s.n_uniq;
}
static inline uint32_t ktn_shape_lane_bytes(const KtnShapeDesc& s) {
    return ktn_shape_blob_lane_bytes(s) + 8u * s.n_scratch;
}
#endif
