// K1 / K5: CSR operator kernels on row-major multi-RHS blocks  X[n x k]  (k contiguous).
//
//   MODE_MULT     Y = S X                                   (G Z, P d, P^T v, final residual check)
//   MODE_MULT_DOT Y = S X,  partials[blk][c] = sum_i X[i][c] Y[i][c]      (q = A p and p^T A p of CG)
//   MODE_RESID    Y = B - S X                               (true-residual check at the end of a block solve)
//
// Mapping: one warp owns a row at a time, lanes run across the RHS columns (double2 per lane -> 512-byte fully
// coalesced requests per gathered row); (col, val) of the sparse row are warp-uniform broadcast loads.  A
// persistent grid (4 x 148 row chunks) walks rows in contiguous chunks so neighbouring rows share their
// gathered X rows through L1, and each lane keeps a private running dot for its columns: no shuffles or atomics
// are needed and the per-block partials are summed in a fixed order by a tiny follow-up kernel (deterministic).
//
// Replaces PETSc MatMult at ForcingCovariance.py:242 (G x), InterpolationMatrix.py:219/255 (P d, P^T v) and the
// operator applications inside the Jacobi-preconditioned KSP solves (pc_type jacobi) of solving_utils.py:51-53; the
// multigrid-preconditioned path uses the staged-tile kernels of tileop.cu instead.
#include "rowops.cuh"

enum { MODE_MULT = 0, MODE_MULT_DOT = 1, MODE_RESID = 3 };

namespace {

template <int VEC, int MODE>
__global__ void __launch_bounds__(RB_THREADS) csr_block_kernel(int64_t n, const int64_t* __restrict__ ptr,
                                                               const int32_t* __restrict__ idx,
                                                               const double* __restrict__ val, int k,
                                                               const double* __restrict__ X, int64_t ldx,
                                                               double* __restrict__ Y, int64_t ldy,
                                                               const double* __restrict__ B, int64_t ldb,
                                                               const double* __restrict__ dinv, double omega,
                                                               double* __restrict__ partials, int64_t rows_per_block) {
    __shared__ double red[RB_WARPS][64];
    const int warp = threadIdx.x >> 5;
    LaneCols<VEC> L(k);
    int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block < n ? r0 + rows_per_block : n;
    double d0 = 0.0, d1 = 0.0;
    for (int64_t row = r0 + warp; row < r1; row += RB_WARPS) {
        double a0 = 0.0, a1 = 0.0;
        csr_row_accum<VEC>(idx, val, ptr[row], ptr[row + 1], X, ldx, L, a0, a1);
        if (MODE == MODE_MULT_DOT) {
            double x0, x1;
            ld_cols<VEC>(X + row * ldx, L, x0, x1);
            d0 += x0 * a0; d1 += x1 * a1;
        } else if (MODE == MODE_RESID) {
            double b0, b1;
            ld_cols<VEC>(B + row * ldb, L, b0, b1);
            a0 = b0 - a0; a1 = b1 - a1;
        }
        st_cols<VEC>(Y + row * ldy, L, a0, a1);
    }
    if (MODE == MODE_MULT_DOT) block_partials<VEC>(d0, d1, L, k, partials, red);
}

__global__ void scatter_columns_kernel(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rows,
                                       const double* __restrict__ vals, int64_t col_begin, int k,
                                       double* __restrict__ B, int64_t ldb) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    for (int64_t p = colptr[col_begin + j]; p < colptr[col_begin + j + 1]; ++p) B[(int64_t)rows[p] * ldb + j] = vals[p];
}

template <int MODE>
int launch_mode(sfem_ctx* c, int tag, int64_t nnz, int64_t n, const int64_t* ptr, const int32_t* idx, const double* val,
                int k, const double* X, int64_t ldx, double* Y, int64_t ldy, const double* B, int64_t ldb,
                const double* dinv, double omega, double* partials) {
    RowGrid g = row_grid(n);
    // algorithmic traffic: the matrix once, each vector block once
    double vecs = (MODE == MODE_RESID) ? 24.0 : 16.0;
    ProfScope ps(c, tag, vecs * (double)n * k + 12.0 * (double)nnz + 8.0 * (double)(n + 1),
                 2.0 * (double)nnz * k);
    bool v2 = can_vec2(k, ldx, ldy, X, Y) && (B == nullptr || can_vec2(k, ldb, ldb, B, B));
    if (v2) {
        dim3 grid(g.nblk, (k + 63) / 64);
        csr_block_kernel<2, MODE><<<grid, RB_THREADS, 0, c->stream>>>(n, ptr, idx, val, k, X, ldx, Y, ldy, B, ldb, dinv,
                                                                      omega, partials, g.rows_per_block);
    } else {
        dim3 grid(g.nblk, (k + 31) / 32);
        csr_block_kernel<1, MODE><<<grid, RB_THREADS, 0, c->stream>>>(n, ptr, idx, val, k, X, ldx, Y, ldy, B, ldb, dinv,
                                                                      omega, partials, g.rows_per_block);
    }
    KCHECK(c);
    return SFEM_OK;
}

}  // namespace

int spmm_csr(sfem_ctx* c, int tag, int64_t nnz, int64_t n, const int64_t* ptr, const int32_t* idx, const double* val, int k,
             const double* X, int64_t ldx, double* Y, int64_t ldy, double* dot_partials, int* nblk_out) {
    if (nblk_out) *nblk_out = row_grid(n).nblk;
    if (n <= 0 || k <= 0) return SFEM_OK;
    if (dot_partials)
        return launch_mode<MODE_MULT_DOT>(c, tag, nnz, n, ptr, idx, val, k, X, ldx, Y, ldy, nullptr, 0, nullptr, 0.0, dot_partials);
    return launch_mode<MODE_MULT>(c, tag, nnz, n, ptr, idx, val, k, X, ldx, Y, ldy, nullptr, 0, nullptr, 0.0, nullptr);
}

int csr_residual(sfem_ctx* c, int64_t n, const int64_t* ptr, const int32_t* idx, const double* val, int k,
                 const double* X, int64_t ldx, double* R, int64_t ldr, const double* B, int64_t ldb) {
    return launch_mode<MODE_RESID>(c, TAG_CSR_RESID, c->A_nnz, n, ptr, idx, val, k, X, ldx, R, ldr, B, ldb, nullptr, 0.0, nullptr);
}

int scatter_columns(sfem_ctx* c, int64_t n, const int64_t* colptr, const int32_t* rows, const double* vals,
                    int64_t col_begin, int k, double* B, int64_t ldb) {
    ProfScope ps(c, TAG_CG_VEC, 8.0 * (double)n * k, 0.0);
    CUDA_TRY(c, cudaMemset2DAsync(B, (size_t)ldb * sizeof(double), 0, (size_t)k * sizeof(double), (size_t)n, c->stream));
    scatter_columns_kernel<<<(k + 127) / 128, 128, 0, c->stream>>>(colptr, rows, vals, col_begin, k, B, ldb);
    KCHECK(c);
    return SFEM_OK;
}
