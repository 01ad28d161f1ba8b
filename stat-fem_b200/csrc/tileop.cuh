// Staged-tile sparse operators ("TileOp"): the storage format and kernel family behind every sparse operator of the
// multigrid-preconditioned block CG (stiffness matrix, Galerkin coarse operators, restrictions, prolongations).
//
// Rows are grouped into chunks of spatially neighbouring unknowns (the unknowns of every level are renumbered along
// lattice tiles at setup).  For each chunk the format stores, at a fixed stride so that every address follows from
// the chunk number alone: a descriptor (row range, local row pointers, slot of every row's own value), the sorted
// list of input rows the chunk touches (its own rows plus a thin halo) and the matrix entries with their column
// replaced by the 16-bit position ("slot") in that list.
//
// Kernel: persistent CTAs, software-pipelined with cp.async.  A CTA walks its chunks and, per chunk, the slabs of 64
// (or 32 / 16 / 4) right-hand-side columns.  While it computes work item t from shared memory, the staged input rows
// and right-hand-side rows of item t+1 are already in flight into the other stage buffer and the descriptors of the
// chunk after next into a 4-deep ring, so no dependent global-memory round trip (descriptor -> row list -> rows) is
// ever on the critical path.  Each input row is read from L2/HBM about once per kernel instead of once per matrix
// entry, and the epilogue fuses the vector work of the caller (Jacobi update, residual, dot products).
//
// Wide blocks (more than 32 right-hand sides) on operators with 32-row chunks (at most 160 staged rows and 16 k-tiles
// per row group: stiffness matrices in 2-D and 3-D, 2-D lattice operators, prolongations, 2-D restrictions) run
// tileop_dmma_kernel instead, in one of two shapes (64- or 32-column slabs, see DmCfg in tileop.cu): the chunk
// matrix is applied with FP64 tensor-core instructions (DMMA m8n8k4) from k-tiles of 4 staged slots x 8 rows, the rows
// are staged by the TMA engine (cp.async.bulk on mbarriers), work items run chunk-major with the block entries held in
// registers, and the warps are decoupled through full/empty mbarriers (see the comment above that kernel in tileop.cu).
#pragma once
#include "rowops.cuh"

constexpr int TILE_MAX_CTAS_PER_SM = 8;
constexpr int TILE_MAX_PARTS = TILE_MAX_CTAS_PER_SM * SFEM_NUM_SMS;   // upper bound of the persistent grid

struct TileOp {
    int64_t nrows = 0, ncols = 0, nnz = 0;
    int nchunks = 0;
    int max_rows = 0;                 // R: largest chunk (rows)
    int max_need = 0;                 // N: largest staged-row list of a chunk
    int max_ent = 0;                  // E: largest number of matrix entries of a chunk
    int dl = 0, n4 = 0, e8 = 0;       // strides: descriptor ints, need ints, entry slots (all 16-byte multiples)
    int* desc = nullptr;              // [nchunks][dl]: r0, rows, need, row width W, own slot, own rows, -, -, self slots (u16)
    int* need = nullptr;              // [nchunks][n4] input rows to stage, ascending
    unsigned short* ecol = nullptr;   // [nchunks][e8]  per-chunk ELL block: row r's W entries at [r*W, r*W + W), zero padded
    double* eval = nullptr;           // [nchunks][e8]
    double* eval_scaled = nullptr;    // [nchunks][e8] a_ij * d_j (square level operators: first sweep folded in), optional
    // ---- FP64 tensor-core form (chunks of 17..32 rows): every group of 8 rows is cut into k-tiles of 4 staged slots
    // with distinct (slot mod 4); a k-tile is an 8x4 block of the chunk matrix, multiplied with DMMA m8n8k4.  Only the
    // nonzeros of a block are stored (lane order), located through a 32-bit lane mask.  desc[6] = k-tile counts of the
    // four row groups (one byte each).
    int kts = 0;                      // k-tile records per chunk (stride), 0: form not built
    int max_kt = 0;                   // most k-tiles of any row group (<= 8: 64-column slabs, <= 16: 32-column slabs)
    int ap8 = 0;                      // packed block values per chunk (stride, even)
    uint4* ktmeta = nullptr;          // [nchunks][kts] {slot0 | slot1 << 16, slot2 | slot3 << 16, lane mask, first value}
    double* apack = nullptr;          // [nchunks][ap8]
    double* apack_scaled = nullptr;   // [nchunks][ap8] from eval_scaled
    ushort2* rgcnt = nullptr;         // [nchunks][4] (k-tiles, nonzeros) of every row group (builder bookkeeping)
};

enum {
    TM_MULT = 0,        // Y_i  = sum_j a_ij x_j
    TM_MULT_ADD = 1,    // Y_i += sum_j a_ij x_j
    TM_MULT_DOT = 2,    // Y_i  = sum_j a_ij x_j, dot += x_i Y_i
    TM_JACOBI = 3,      // Y_i = x_i + d_i (b_i - sum a_ij x_j)
    TM_JACOBI_DOT = 4,  // as TM_JACOBI, dot += b_i Y_i
    TM_RESID = 5,       // Y_i = b_i - sum a_ij x_j
    TM_PRESMOOTH2 = 6,  // Y_i = d_i x_i + d2_i (x_i - sum a_ij d_j x_j): two weighted-Jacobi sweeps from a zero guess, x = rhs
};

struct TileArgs {
    int k = 0;
    const double* X = nullptr; int64_t ldx = 0;     // staged input block
    double* Y = nullptr; int64_t ldy = 0;           // output block
    const double* B = nullptr; int64_t ldb = 0;     // right-hand side (Jacobi / residual modes)
    const double* dinv = nullptr;                   // per-row scaling (damping folded in)
    const double* dinv2 = nullptr;                  // TM_PRESMOOTH2: scaling of the second sweep (nullptr: same as dinv)
    double* partials = nullptr;                     // [TILE_MAX_PARTS][k] per-CTA dot partials (rows actually written: *nparts)
};

// Build from a CSR matrix (device arrays) whose rows and columns are already in the renumbered spaces.
// chunk_ptr: device array of nchunks + 1 row offsets.  Fails with SFEM_ERR_LIMIT if a chunk has more than 8192 entries.
int tileop_build(sfem_ctx* c, int64_t nrows, int64_t ncols, const int64_t* ptr, const int32_t* idx, const double* val,
                 const int* chunk_ptr, int nchunks, int max_rows, TileOp* out);
// eval_scaled[e] = eval[e] * dinv[column of e]   (dinv in the operator's own numbering; square operators)
int tileop_scale_columns(sfem_ctx* c, TileOp* op, const double* dinv);
void tileop_free(TileOp* op);
// the operator has the tensor-core form and its chunk working set fits two CTAs per SM
bool tileop_dmma_usable(const TileOp& op);
// tag: profile family; bytes: algorithmic traffic charged to it; nparts (nullable): rows of a.partials written
int tileop_apply(sfem_ctx* c, int tag, const TileOp& op, int mode, const TileArgs& a, double bytes, int* nparts = nullptr);
