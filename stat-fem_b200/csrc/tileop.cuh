// Staged-tile sparse operators ("TileOp"): the storage format and kernel family behind every sparse operator of the
// multigrid-preconditioned block CG (stiffness matrix, Galerkin coarse operators, restrictions, prolongations).
//
// Rows are grouped into chunks of spatially neighbouring unknowns (the unknowns of every level are renumbered along
// lattice tiles at setup).  For each chunk the format stores the sorted list of input rows the chunk touches
// (`need_rows`: its own rows plus a thin halo) and the matrix entries with their column replaced by the 16-bit
// position ("slot") in that list.  A CTA applies one chunk to one slab of right-hand-side columns:
//   1. it stages the needed rows of the input block X[:, slab] into shared memory with fully coalesced row loads,
//      all issued before any is consumed (memory-level parallelism instead of the ptr -> idx -> X dependent chain of
//      a CSR gather);
//   2. every warp then forms its rows from shared memory (lanes = RHS columns; conflict-free 16-byte reads);
//   3. the epilogue fuses the vector work of the caller (Jacobi update, residual, CG direction update, dots).
// Each input row is thus read from L2/HBM about once per kernel instead of once per matrix entry.
#pragma once
#include "rowops.cuh"

struct TileOp {
    int64_t nrows = 0, ncols = 0, nnz = 0;
    int nchunks = 0;
    int max_rows = 0;                 // largest chunk (rows)
    int max_need = 0;                 // largest staged-row list of a chunk
    int max_ent = 0;                  // largest number of matrix entries of a chunk
    int* chunk_ptr = nullptr;         // [nchunks + 1] first row of each chunk
    int* meta = nullptr;              // [nchunks][8]: r0, rows, need offset, need count, entry offset, entries, own slot, -
    int* need_ptr = nullptr;          // [nchunks + 1]
    int* need_rows = nullptr;         // [need_ptr[nchunks]] input rows to stage, ascending inside a chunk
    int* ent_ptr = nullptr;           // [nrows + 1]
    unsigned short* ent_col = nullptr;    // [nnz] slot of the entry's column in the chunk's staged list
    double* ent_val = nullptr;            // [nnz]
    unsigned short* self_slot = nullptr;  // [nrows] slot of the row itself (square operators), 0xffff if absent
};

enum {
    TM_MULT = 0,        // Y_i  = sum_j a_ij x_j
    TM_MULT_ADD = 1,    // Y_i += sum_j a_ij x_j
    TM_PMULT_DOT = 2,   // p_j = x_j + beta p_j (staged; own rows written to Pout), Y_i = sum a_ij p_j, dot += p_i Y_i
    TM_JACOBI = 3,      // Y_i = x_i + d_i (b_i - sum a_ij x_j)
    TM_JACOBI_DOT = 4,  // as TM_JACOBI, dot += b_i Y_i
    TM_RESID = 5,       // Y_i = b_i - sum a_ij x_j
    TM_PRESMOOTH2 = 6,  // s_j = d_j x_j (staged), Y_i = s_i + d_i (x_i - sum a_ij s_j): two Jacobi sweeps from a zero guess
};

struct TileArgs {
    int k = 0;
    const double* X = nullptr; int64_t ldx = 0;     // staged input block
    double* Y = nullptr; int64_t ldy = 0;           // output block
    const double* B = nullptr; int64_t ldb = 0;     // right-hand side (Jacobi / residual modes)
    const double* dinv = nullptr;                   // per-row scaling (damping folded in)
    const double* Pin = nullptr; double* Pout = nullptr; int64_t ldp = 0;   // TM_PMULT_DOT
    const double* beta = nullptr;                   // TM_PMULT_DOT: per column
    double* partials = nullptr;                     // [nchunks][k] per-chunk dot partials
};

// Build from a CSR matrix (device arrays) whose rows and columns are already in the renumbered spaces.
// chunk_ptr: device array of nchunks + 1 row offsets (copied).  Fails with SFEM_ERR_LIMIT if a chunk has more than
// 8192 entries or needs more than 65535 rows.
int tileop_build(sfem_ctx* c, int64_t nrows, int64_t ncols, const int64_t* ptr, const int32_t* idx, const double* val,
                 const int* chunk_ptr, int nchunks, int max_rows, TileOp* out);
void tileop_free(TileOp* op);
// tag: profile family; bytes: algorithmic traffic charged to it
int tileop_apply(sfem_ctx* c, int tag, const TileOp& op, int mode, const TileArgs& a, double bytes);
