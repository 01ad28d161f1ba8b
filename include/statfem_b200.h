/*
 * statfem_b200.h -- C ABI of libstatfem_b200.so: the B200 (sm_100a) implementation of stat-fem's data-conditioning
 * hot path.  This is the drop-in boundary: plain pointers and sizes, no torch types.
 *
 * The reference (alan-turing-institute/stat-fem) is pure Python with no FFI layer; each entry point below names the
 * reference statement(s) whose arithmetic it replaces (paths relative to stat_fem/).  INTEGRATION.md shows the ctypes
 * binding a maintainer would add inside the reference's classes.
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer owned by the caller unless its name ends in _host.  The library never frees
 *     caller memory; it keeps the pointers passed to sfem_set_stiffness / sfem_set_gcov / sfem_gcov_count until they
 *     are replaced or the context is destroyed.
 *   - Dense matrices are row-major.  Multi-RHS blocks are row-major [n x k] with leading dimension ld (k contiguous).
 *   - CSR: int64 row pointers, int32 column indices (PETSc.IntType, ForcingCovariance.py:164), double values.
 *   - Every function returns 0 on success or a negative SFEM_ERR_* code; sfem_last_error() gives the message.
 *   - One context = one GPU = one stream; a context is not thread-safe, distinct contexts are independent.
 *   - There is no CPU fallback anywhere in this library.
 */
#ifndef STATFEM_B200_H
#define STATFEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sfem_ctx sfem_ctx;

#define SFEM_OK 0
#define SFEM_ERR_CUDA (-1)
#define SFEM_ERR_ARG (-2)
#define SFEM_ERR_NOT_SPD (-3)        /* -> scipy.linalg.LinAlgError, as LinearSolver.py:272-276, 602-606 */
#define SFEM_ERR_NOT_CONVERGED (-4)  /* CG hit max_it */
#define SFEM_ERR_STATE (-5)
#define SFEM_ERR_LIMIT (-6)

#define SFEM_OP_N 0
#define SFEM_OP_T 1

/* ---- lifecycle ------------------------------------------------------------------------------------------------- */
int sfem_version(void);
/* cuda_stream: a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream) or NULL for the default stream */
int sfem_create(sfem_ctx** out, int device, void* cuda_stream);
int sfem_destroy(sfem_ctx* ctx);
const char* sfem_last_error(sfem_ctx* ctx);
long long sfem_launch_count(sfem_ctx* ctx);   /* kernels launched so far by this context */

/* ---- measurement: per-kernel-family device time (CUDA events on the launch stream) and algorithmic bytes / flops.
 *      mask bit t enables family t; sfem_profile_read fills arrays of sfem_profile_num_tags() entries (totals since
 *      the last sfem_profile_enable).  No reference counterpart (the reference has no timers). */
int sfem_profile_enable(sfem_ctx* ctx, unsigned mask);
int sfem_profile_num_tags(void);
const char* sfem_profile_tag_name(int tag);
int sfem_profile_read(sfem_ctx* ctx, double* ms_out_host, double* bytes_out_host, double* flops_out_host,
                      long long* count_out_host);

/* ---- stiffness matrix and the sparse solve (replaces firedrake LinearSolver/PETSc KSP: LinearSolver.py:136-140,
 *      solving_utils.py:49-53, LinearSolver.py:217) --------------------------------------------------------------- */
/* A: n x n CSR, symmetric positive definite, Dirichlet rows AND columns replaced by identity (as Firedrake assembles
 * with bcs=; tests/helper_funcs.py:104-128). */
int sfem_set_stiffness(sfem_ctx* ctx, int64_t n, const int64_t* indptr, const int32_t* indices, const double* vals);
/* Build the lattice multigrid preconditioner.  coords: n x dim row-major DOF coordinates; bc_mask: n bytes (1 =
 * Dirichlet DOF); lo/hi_host: bounding box; m1_host: level-1 lattice nodes per direction (2^p + 1 each);
 * nu: Jacobi sweeps before and after; omega / omega2: weights of the odd / even sweeps (equal: damped Jacobi; the
 * reciprocals of the degree-2 Chebyshev roots on [lmax/4, lmax]: polynomial smoothing; the post-smoother applies
 * them in reverse order so that the cycle stays symmetric); max_coarse: coarsen until a level has at most this many
 * nodes. */
int sfem_mg_setup(sfem_ctx* ctx, int dim, const double* coords, const uint8_t* bc_mask, const double* lo_host,
                  const double* hi_host, const int* m1_host, int nu, double omega, double omega2, int max_coarse);
int sfem_mg_num_levels(sfem_ctx* ctx);
int64_t sfem_mg_level_size(sfem_ctx* ctx, int level);
size_t sfem_solve_workspace_bytes(sfem_ctx* ctx, int k, int precond);
/* X (n x k, ldx) = A^-1 B (n x k, ldb): k independent PCG recurrences advanced in lock-step.  precond: 0 Jacobi,
 * 1 multigrid V-cycle.  Stops when ||r||_2 <= max(rtol ||b||_2, atol) for every column.  iters_out_host: iterations;
 * relres_out_host[k]: TRUE relative residuals ||b - A x|| / ||b|| recomputed at the end. */
int sfem_solve_block(sfem_ctx* ctx, int k, const double* B, int64_t ldb, double* X, int64_t ldx, int precond,
                     double rtol, double atol, int max_it, int* iters_out_host, double* relres_out_host, void* work,
                     size_t work_bytes);

/* ---- generic CSR SpMM on multi-RHS blocks: Y (nrows x k) = S X.  Used for P d (InterpolationMatrix.py:213-221), P^T v
 *      (InterpolationMatrix.py:254-259, with P's CSC arrays read as the CSR of P^T) and G x (ForcingCovariance.py:242). */
int sfem_spmm_csr(sfem_ctx* ctx, int64_t nrows, int64_t nnz, const int64_t* indptr, const int32_t* indices,
                  const double* vals, int k, const double* X, int64_t ldx, double* Y, int64_t ldy);
/* B (n x k, ldb) = dense copy of columns [col_begin, col_begin+k) of the CSC matrix P
 * (InterpolationMatrix.get_meshspace_column_vector, InterpolationMatrix.py:261-287, for a whole block of sensors). */
int sfem_scatter_columns(sfem_ctx* ctx, int64_t n, const int64_t* colptr, const int32_t* rows, const double* vals,
                         int64_t col_begin, int k, double* B, int64_t ldb);

/* ---- forcing covariance assembly (replaces ForcingCovariance._compute_G_vals + assemble, ForcingCovariance.py:130-197) */
/* Pass 1: cell list + per-row count.  exp_2sigma = exp(2 sigma), exp_m2l = exp(-2 l) evaluated by the caller (so they
 * match the reference's NumPy values bitwise).  bmax = max(int_basis).  cell_edge >= the largest distance at which
 * row[j]/row[i] > cutoff can hold (0: all pairs, required when cutoff <= 0).  indptr_out: n+1 int64 (device).
 * borderline_rows (device, capacity borderline_cap): rows holding a ratio within 1e-14 (relative) of the cutoff. */
int sfem_gcov_count(sfem_ctx* ctx, int64_t n, int dim, const double* coords, const double* int_basis, double exp_2sigma,
                    double exp_m2l, double cutoff, double regularization, double bmax, double cell_edge,
                    const double* lo_host, const double* hi_host, int64_t* indptr_out, int64_t* nnz_out_host,
                    int64_t* max_row_nnz_out_host, int32_t* borderline_rows, int borderline_cap,
                    int64_t* n_borderline_out_host);
/* Pass 2: write ascending int32 columns and values for the pattern counted by pass 1. */
int sfem_gcov_fill(sfem_ctx* ctx, const int64_t* indptr, int32_t* indices_out, double* vals_out);

/* ---- dense FP64 tensor-core primitives (replace NumPy/SciPy BLAS+LAPACK calls of LinearSolver.py) ---------------- */
/* C (M x N, ldc) = alpha op(A) op(B) + beta C.  opA = SFEM_OP_T means A is stored K x M.  Cu = W^T Z is
 * sfem_dgemm(T, N, n_obs, n_cols, n_dof, 1, W, ldw, Z, ldz, 0, Cu, ldcu)  (solving_utils.py:139-142, 164-169). */
int sfem_dgemm(sfem_ctx* ctx, int opA, int opB, int M, int N, int64_t K, double alpha, const double* A, int64_t lda,
               const double* B, int64_t ldb, double beta, double* C, int64_t ldc);
/* Symmetric split of the covariance assembly: with G = S + E (E: the entries of G whose transpose is not stored --
 * ForcingCovariance.py:160-164 drops (j, i) but keeps (i, j) only next to the boundary), (G Z)^T Z = (S Z)^T Z + (E Z)^T Z
 * where the first term is symmetric (sfem_dgemm_lower + sfem_mirror_lower: half the tiles) and the second sums over
 * the few non-empty rows of E only.  sfem_csr_unsym_count / _fill extract E as a compact-row CSR (rows_out lists the
 * non-empty rows); scratch arrays: row_cnt, row_flag n int64; eptr_full, rowpos n + 1 int64 (all device). */
int sfem_csr_unsym_count(sfem_ctx* ctx, int64_t n, const int64_t* indptr, const int32_t* indices, int64_t* row_cnt,
                         int64_t* row_flag, int64_t* eptr_full, int64_t* rowpos, int64_t* nrows_out_host,
                         int64_t* nnz_out_host);
int sfem_csr_unsym_fill(sfem_ctx* ctx, int64_t n, const int64_t* indptr, const int32_t* indices, const double* vals,
                        const int64_t* eptr_full, const int64_t* rowpos, int32_t* rows_out, int64_t* eptr_out,
                        int32_t* eidx_out, double* eval_out);
/* Y[rows[r]] += alpha X[r]  and  X[r] = Y[rows[r]]  for row-major blocks of k columns */
int sfem_rows_axpy(sfem_ctx* ctx, int64_t nrows, const int32_t* rows, int k, double alpha, const double* X, int64_t ldx,
                   double* Y, int64_t ldy);
int sfem_rows_gather(sfem_ctx* ctx, int64_t nrows, const int32_t* rows, int k, const double* Y, int64_t ldy, double* X,
                     int64_t ldx);
/* C (n x n) = alpha op(A) op(B) + beta C on the 128x128 tiles on and below the diagonal only; sfem_mirror_lower copies
 * the strict lower triangle onto the upper one. */
int sfem_dgemm_lower(sfem_ctx* ctx, int opA, int opB, int n, int64_t K, double alpha, const double* A, int64_t lda,
                     const double* B, int64_t ldb, double beta, double* C, int64_t ldc);
int sfem_mirror_lower(sfem_ctx* ctx, int n, double* A, int64_t lda);
/* Sharded form of the same split (ensemble ranks own contiguous column blocks, solving_utils.py:117-123): every
 * unordered pair of different ownership blocks of the symmetric term is computed by ONE rank (half-ring schedule), so
 * the gathered n x n array holds block (r, s) or block (s, r) and zeros in the other; sfem_sym_fold adds each
 * off-diagonal-block entry to its transpose in place (nblocks = ensemble size, PETSc's contiguous ownership rule). */
int sfem_sym_fold(sfem_ctx* ctx, int n, int nblocks, double* A, int64_t lda);
/* W = G Z on FP64 tensor cores (csrc/bsr.cu; replaces PETSc MatMult at ForcingCovariance.py:242 for blocks of
 * right-hand sides).  sfem_b84_build converts a CSR matrix with ascending column indices per row (device arrays, the
 * output of sfem_gcov_fill) into 8x4 blocks: 8 consecutive rows x the sorted union of their columns, 4 at a time, as
 * dense DMMA A operands; *ntiles_out_host receives the number of blocks (32 doubles + 4 indices each).  The handle owns
 * device memory of the library and is released with sfem_b84_free.  sfem_b84_spmm: Y (n x k, ldy, 16-byte aligned,
 * ldy even) = G X (n x k, ldx); X and Y must not alias. */
int sfem_b84_build(sfem_ctx* ctx, int64_t n, const int64_t* indptr, const int32_t* indices, const double* vals,
                   void** matrix_out, int64_t* ntiles_out_host);
int sfem_b84_spmm(sfem_ctx* ctx, const void* matrix, int k, const double* X, int64_t ldx, double* Y, int64_t ldy);
int sfem_b84_free(sfem_ctx* ctx, void* matrix);
/* Self-check hook of the staged-tile operators (csrc/tileop.cu): applies ONE operator of the multigrid hierarchy built
 * by sfem_mg_setup -- which: 0 = level operator, 1 = prolongation from level+1, 2 = restriction to level+1; mode: 0 Y=AX,
 * 1 Y+=AX, 2 Y=AX with dot(X,Y), 3 Jacobi sweep, 4 Jacobi sweep with dot(B,Y), 5 residual B-AX, 6 two sweeps from zero --
 * in that level's tile numbering, with the FP64 tensor-core (DMMA) form of the operator switched on or off, so that
 * tests can compare the two kernel families on the same data.  X: cols x k, B / Y: rows x k (leading dimension k),
 * partials: 1184 x k per-CTA dot partials of which *nparts_out_host rows are written.  With k = 0 only the shape and
 * the number of k-tile records per chunk (0: tensor form not built for this operator) are reported. */
int sfem_tile_apply_one(sfem_ctx* ctx, int level, int which, int mode, int k, const double* X, const double* B, double* Y,
                        double* partials, int* nparts_out_host, int use_tensor_form, int64_t* rows_out_host,
                        int64_t* cols_out_host, int* tensor_form_out_host);
size_t sfem_potrf_dinv_doubles(int n);
/* In-place lower Cholesky of the row-major n x n matrix A (only the lower triangle is read); dinv receives the
 * inverses of the 128x128 diagonal blocks.  info_out_host: 0 or the 1-based index of the first non-positive pivot. */
int sfem_potrf(sfem_ctx* ctx, int n, double* A, int64_t lda, double* dinv, int* info_out_host);
int sfem_potrs(sfem_ctx* ctx, int n, const double* L, int64_t ldl, const double* dinv, int k, double* B, int64_t ldb);
int sfem_potri(sfem_ctx* ctx, int n, const double* L, int64_t ldl, const double* dinv, double* work_nn, double* Kinv);
int sfem_logdet(sfem_ctx* ctx, int n, const double* L, int64_t ldl, double* out_host);

/* ---- model-discrepancy matrices and the fused dense end (ObsData.py:131-214; LinearSolver.py:351-367, 600-613, 659-681) */
/* which = 0: M = sqexp(Xs,Xs,sigma,l) [+ diag(unc^2) if unc != NULL] [+ rho2 * sym_upper(Cu) if Cu != NULL]
 * which = 1: dK/dsigma = 2K;  which = 2: dK/dl.   Xs: n_obs x dim.  unc: 1 value, or n_obs values if unc_is_vector. */
int sfem_kernel_matrix(sfem_ctx* ctx, int n_obs, int dim, const double* Xs, double sigma, double l, const double* unc,
                       int unc_is_vector, double rho2, const double* Cu, int which, double* M_out);
/* Rectangular covariance block between two point sets X1 (m x dim) and X2 (n x dim) -> M_out (m x n):
 * which = 0 sqexp, 1 d/dsigma, 2 d/dl, 3 squared distance (covariance_functions.py:4-37, 39-77, 122-145). */
int sfem_cov_matrix(sfem_ctx* ctx, int m, int n, int dim, const double* X1, const double* X2, double sigma, double l,
                    int which, double* M_out);
/* z = a x + b y on device vectors (y may be NULL): the mesh-space combinations of LinearSolver.py:287, 303-305. */
int sfem_axpby(sfem_ctx* ctx, int64_t n, double a, const double* x, double b, const double* y, double* z);
/* Y[n x ns] = X[n x k] C[k x ns] for a few columns ns (row-major, leading dimensions in elements): v = W c with the
 * resident W = G Z, so that A^-1 G A^-1 P c = A^-1 (W c) costs one mesh-space solve instead of the two of
 * solving_utils.py:49-53 in every term of the posterior mean (LinearSolver.py:283-301). */
int sfem_skinny_gemm(sfem_ctx* ctx, int64_t n, int k, int ns, const double* X, int64_t ldx, const double* C, int64_t ldc,
                     double* Y, int64_t ldy);
size_t sfem_dense_workspace_bytes(int n_obs);
/* Negative log marginal likelihood and (if grad_out_host != NULL) its gradient at theta_host = (log rho, log sigma_eta,
 * log l_eta): LinearSolver.logposterior / logpost_deriv. */
int sfem_logpost(sfem_ctx* ctx, int n_obs, int dim, const double* Xs, const double* y, const double* unc,
                 int unc_is_vector, const double* mu, const double* Cu, const double* theta_host, double* value_out_host,
                 double* grad_out_host, void* work, size_t work_bytes);
/* LinearSolver.solve_posterior_covariance: muy_out (n_obs), Cuy_out (n_obs x n_obs), unscaled. */
int sfem_posterior_cov(sfem_ctx* ctx, int n_obs, int dim, const double* Xs, const double* y, const double* unc,
                       int unc_is_vector, const double* mu, const double* Cu, const double* theta_host, double* muy_out,
                       double* Cuy_out, void* work, size_t work_bytes);
/* x = (K + diag(unc^2) + rho2 sym_upper(Cu))^-1 b : the data-space solves of LinearSolver.solve_posterior
 * (LinearSolver.py:271-277, 291-297).  Cu may be NULL when rho2 == 0. */
int sfem_dense_solve(sfem_ctx* ctx, int n_obs, int dim, const double* Xs, const double* unc, int unc_is_vector,
                     double sigma, double l, double rho2, const double* Cu, const double* b, double* x, void* work,
                     size_t work_bytes);

/* ---- The sharded covariance assembly (ensemble parallelism of solving_utils.py:117-159) over NCCL ---------------------
 * The reference splits the sensor columns over `ensemble_comm` (contiguous ownership, lines 117-123), every rank runs its
 * own solves (139-142) and the rows are gathered on ensemble rank 0 (146-159).  Here: one context per GPU, one NCCL
 * communicator per context, and ONE call per rank.  NCCL is bound at run time (dlopen of libnccl.so.2); a single rank
 * needs neither NCCL nor a communicator.
 *
 * sfem_nccl_unique_id : rank 0 fills 128 bytes (an ncclUniqueId); the caller hands them to every rank (MPI_Bcast, a file).
 * sfem_comm_init      : collective over the nranks contexts; sfem_comm_attach adopts an ncclComm_t the caller owns.
 * sfem_comm_bcast     : in-place broadcast of a device buffer (e.g. the cached prior x, mu, Cu for per-rank MAP fits).
 * sfem_prior_cov      : this rank owns sensor columns [col_begin, col_end) -- the ranks' ranges must tile [0, n_obs) in
 *                       rank order (PETSc's rule does).  Computes Z_r = A^-1 P[:, J_r] by block PCG in blocks of
 *                       `block_cols` columns (0: 384), W_r = G Z_r, and rows J_r of Cu = (G Z)^T Z, receiving the other
 *                       ranks' Z blocks round a ring (ncclSend / ncclRecv on a second stream, row chunks, double-buffered,
 *                       overlapped with the DMMA GEMM); the rows are gathered into Cu_out (n_obs x n_obs, row-major) on
 *                       rank 0.  P is CSC (device), G CSR (device, the output of sfem_gcov_fill), A the matrix of
 *                       sfem_set_stiffness (+ sfem_mg_setup when precond = 1).  Z_out / W_out: optional n x (col_end -
 *                       col_begin) device blocks that receive Z_r and W_r (NULL: temporaries).  Cu_out may be NULL on
 *                       ranks other than 0.  iters_max_out_host: largest CG iteration count of any block.
 *                       Collective: every rank of the communicator must call it. */
int sfem_nccl_unique_id(void* id_out_128_bytes_host);
int sfem_comm_init(sfem_ctx* ctx, const void* unique_id_128_bytes_host, int rank, int nranks);
int sfem_comm_attach(sfem_ctx* ctx, void* nccl_comm, int rank, int nranks);
int sfem_comm_destroy(sfem_ctx* ctx);
int sfem_comm_bcast(sfem_ctx* ctx, void* device_buffer, size_t bytes, int root);
int sfem_prior_cov(sfem_ctx* ctx, int64_t n_obs, const int64_t* P_colptr, const int32_t* P_rows, const double* P_vals,
                   const int64_t* G_indptr, const int32_t* G_indices, const double* G_vals, int64_t col_begin,
                   int64_t col_end, int block_cols, int precond, double rtol, double atol, int max_it, double* Z_out,
                   double* W_out, double* Cu_out, int* iters_max_out_host);

#ifdef __cplusplus
}
#endif
#endif /* STATFEM_B200_H */
