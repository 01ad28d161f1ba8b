// extern "C" surface of libstatfem_b200.so (see include/statfem_b200.h for the contract of every entry point).
#include "common.cuh"
#include <cstdlib>
#include "../../include/statfem_b200.h"

int stiffness_dinv(sfem_ctx* c);
int mg_setup(sfem_ctx* c, int dim, const double* coords, const unsigned char* bc_mask, const double* lo, const double* hi,
             const int* m1, int nu, double omega, int max_coarse);
int mg_num_levels(sfem_ctx* c);
int64_t mg_level_size(sfem_ctx* c, int l);
size_t cg_workspace_bytes(sfem_ctx* c, int k, int precond);
int cg_solve_block(sfem_ctx* c, int k, const double* B, int64_t ldb, double* X, int64_t ldx, int precond, double rtol,
                   double atol, int max_it, int* iters_out, double* relres_out_host, void* work, size_t work_bytes);
int scatter_columns(sfem_ctx* c, int64_t n, const int64_t* colptr, const int32_t* rows, const double* vals,
                    int64_t col_begin, int k, double* B, int64_t ldb);
void gcov_free(sfem_ctx* c);
void dense_graphs_free(sfem_ctx* c);
void prof_free(sfem_ctx* c);
void prof_collect(sfem_ctx* c);
void prof_reset(sfem_ctx* c);
int gcov_count(sfem_ctx* c, int64_t n, int d, const double* coords, const double* int_basis, double exp_2sigma,
               double exp_m2l, double cutoff, double reg, double bmax, double cell_edge, const double* lo,
               const double* hi, int64_t* indptr_out, int64_t* nnz_out, int64_t* max_row_nnz_out,
               int32_t* borderline_rows, int borderline_cap, int64_t* n_borderline_out);
int gcov_fill(sfem_ctx* c, const int64_t* indptr, int32_t* indices_out, double* vals_out);
int unsym_count(sfem_ctx* c, int64_t n, const int64_t* ptr, const int32_t* idx, int64_t* row_cnt, int64_t* row_flag,
                int64_t* eptr_full, int64_t* rowpos, int64_t* nrows_host, int64_t* nnz_host);
int unsym_fill(sfem_ctx* c, int64_t n, const int64_t* ptr, const int32_t* idx, const double* val, const int64_t* eptr_full,
               const int64_t* rowpos, int32_t* rows_out, int64_t* eptr_out, int32_t* eidx_out, double* eval_out);
int mg_tile_apply_one(sfem_ctx* c, int level, int which, int mode, int k, const double* X, const double* B, double* Y,
                      double* partials, int* nparts_host, int use_dmma, int64_t* rows_out, int64_t* cols_out, int* has_dmma_out);
struct B84Matrix;
int b84_build(sfem_ctx* c, int64_t n, const int64_t* ptr, const int32_t* idx, const double* val, B84Matrix** out);
int b84_spmm(sfem_ctx* c, const B84Matrix* m, int k, const double* X, int64_t ldx, double* Y, int64_t ldy);
void b84_free(B84Matrix* m);
int64_t b84_num_tiles(const B84Matrix* m);
int rows_axpy(sfem_ctx* c, int64_t nrows, const int32_t* rows, int k, double alpha, const double* X, int64_t ldx, double* Y, int64_t ldy);
int rows_gather(sfem_ctx* c, int64_t nrows, const int32_t* rows, int k, const double* Y, int64_t ldy, double* X, int64_t ldx);
int mirror_lower(sfem_ctx* c, int n, double* A, int64_t ld);
int sym_fold(sfem_ctx* c, int n, int nblocks, double* A, int64_t ld);
int potrf_lower(sfem_ctx* c, int n, double* A, int64_t lda, double* dinv, int* info_dev);
int potrs_lower(sfem_ctx* c, int n, const double* L, int64_t ldl, const double* dinv, int k, double* B, int64_t ldb);
int potri_lower(sfem_ctx* c, int n, const double* L, int64_t ldl, const double* dinv, double* W, double* Kinv);
int logdet_lower(sfem_ctx* c, int n, const double* L, int64_t ldl, double* out_dev);
size_t potrf_dinv_doubles(int n);
int dense_kernel_matrix(sfem_ctx* c, int n, int d, const double* Xs, double sigma, double l, const double* unc,
                        int unc_vec, double rho2, const double* Cu, int which, double* M);
size_t dense_workspace_bytes(int n);
int dense_axpby(sfem_ctx* c, int64_t n, double a, const double* x, double b, const double* y, double* z);
int dense_skinny_gemm(sfem_ctx* c, int64_t n, int k, int ns, const double* X, int64_t ldx, const double* Cm, int64_t ldc,
                      double* Y, int64_t ldy);
int dense_cov_matrix(sfem_ctx* c, int m, int n, int d, const double* X1, const double* X2, double sigma, double l, int which,
                     double* M);
int dense_logpost(sfem_ctx* c, int n, int d, const double* Xs, const double* y, const double* unc, int unc_vec,
                  const double* mu, const double* Cu, const double* theta_host, double* value_host, double* grad_host,
                  void* work, size_t work_bytes);
int dense_posterior_cov(sfem_ctx* c, int n, int d, const double* Xs, const double* y, const double* unc, int unc_vec,
                        const double* mu, const double* Cu, const double* theta_host, double* muy_out, double* Cuy_out,
                        void* work, size_t work_bytes);
int dense_solve_shifted(sfem_ctx* c, int n, int d, const double* Xs, const double* unc, int unc_vec, double sigma, double l,
                        double rho2, const double* Cu, const double* b, double* x, void* work, size_t work_bytes);

int sharded_unique_id(void* id_out);
int sharded_comm_init(sfem_ctx* c, const void* id128, int rank, int nranks);
int sharded_comm_attach(sfem_ctx* c, void* nccl_comm_handle, int rank, int nranks);
int sharded_comm_destroy(sfem_ctx* c);
int sharded_bcast(sfem_ctx* c, void* buf, size_t bytes, int root);
int sharded_prior_cov(sfem_ctx* c, int64_t n_obs, const int64_t* P_colptr, const int32_t* P_rows, const double* P_vals,
                      const int64_t* G_indptr, const int32_t* G_indices, const double* G_vals, int64_t col_begin,
                      int64_t col_end, int block_cols, int precond, double rtol, double atol, int max_it, double* Z_out,
                      double* W_out, double* Cu_out, int* iters_max_out_host);

thread_local cudaStream_t sfem_tls_stream = nullptr;

#define GUARD(ctx)                   \
    if (!(ctx)) return SFEM_ERR_ARG; \
    cudaSetDevice((ctx)->device);    \
    sfem_tls_stream = (ctx)->stream;

extern "C" {

int sfem_version(void) { return 100; }

int sfem_create(sfem_ctx** out, int device, void* cuda_stream) {
    if (!out) return SFEM_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return SFEM_ERR_CUDA;   // no GPU: fail loudly
    if (device < 0 || device >= ndev) return SFEM_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return SFEM_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SFEM_ERR_CUDA;
    if (prop.major != 10) return SFEM_ERR_CUDA;   // sm_100a binary only
    sfem_ctx* c = new sfem_ctx();
    c->device = device;
    c->stream = (cudaStream_t)cuda_stream;
    {   // keep freed blocks in the pool (see sfem_dev_malloc in common.cuh)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    if (const char* e = getenv("SFEM_TILE_ROWS")) c->tile_rows = atoi(e);     // tuning knobs of the staged-tile kernels
    if (const char* e = getenv("SFEM_TILE_SW")) c->tile_sw = atoi(e);
    if (const char* e = getenv("SFEM_TILE_DMMA")) c->tile_dmma = atoi(e);
    if (cudaMallocHost(&c->pinned, 4096) != cudaSuccess) {
        delete c;
        return SFEM_ERR_CUDA;
    }
    *out = c;
    return SFEM_OK;
}

int sfem_destroy(sfem_ctx* c) {
    if (!c) return SFEM_OK;
    cudaSetDevice(c->device);
    sfem_tls_stream = c->stream;
    cudaStreamSynchronize(c->stream);
    sharded_comm_destroy(c);
    mg_free(c);
    gcov_free(c);
    prof_free(c);
    dense_graphs_free(c);
    if (c->A_dinv) sfem_dev_free(c->A_dinv);
    if (c->scratch) sfem_dev_free(c->scratch);
    if (c->dense_tmp) sfem_dev_free(c->dense_tmp);
    if (c->aux_stream) { cudaStreamSynchronize(c->aux_stream); cudaStreamDestroy(c->aux_stream); }
    if (c->ev_main) cudaEventDestroy(c->ev_main);
    if (c->ev_aux) cudaEventDestroy(c->ev_aux);
    if (c->ev_head) cudaEventDestroy(c->ev_head);
    if (c->pinned) cudaFreeHost(c->pinned);
    delete c;
    return SFEM_OK;
}

const char* sfem_last_error(sfem_ctx* c) { return c ? c->err : "null context (no CUDA device, or not an sm_100 GPU?)"; }
long long sfem_launch_count(sfem_ctx* c) { return c ? c->launches : 0; }

int sfem_set_stiffness(sfem_ctx* c, int64_t n, const int64_t* indptr, const int32_t* indices, const double* vals) {
    GUARD(c);
    if (n <= 0 || !indptr || !indices || !vals) return sfem_fail(c, SFEM_ERR_ARG, "set_stiffness: bad arguments");
    mg_free(c);
    c->n = n; c->A_ptr = indptr; c->A_idx = indices; c->A_val = vals;
    int64_t nnz = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&nnz, indptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->A_nnz = nnz;
    return stiffness_dinv(c);
}

int sfem_mg_setup(sfem_ctx* c, int dim, const double* coords, const uint8_t* bc_mask, const double* lo_host,
                  const double* hi_host, const int* m1_host, int nu, double omega, double omega2, int max_coarse) {
    GUARD(c);
    if (!(omega > 0.0) || !(omega2 > 0.0)) return sfem_fail(c, SFEM_ERR_ARG, "mg_setup: smoothing weights must be positive");
    c->mg_omega2 = omega2;
    return mg_setup(c, dim, coords, bc_mask, lo_host, hi_host, m1_host, nu, omega, max_coarse);
}
int sfem_mg_num_levels(sfem_ctx* c) { return c ? mg_num_levels(c) : 0; }
int64_t sfem_mg_level_size(sfem_ctx* c, int level) { return c ? mg_level_size(c, level) : -1; }
size_t sfem_solve_workspace_bytes(sfem_ctx* c, int k, int precond) { return c ? cg_workspace_bytes(c, k, precond) : 0; }

int sfem_solve_block(sfem_ctx* c, int k, const double* B, int64_t ldb, double* X, int64_t ldx, int precond, double rtol,
                     double atol, int max_it, int* iters_out_host, double* relres_out_host, void* work, size_t work_bytes) {
    GUARD(c);
    return cg_solve_block(c, k, B, ldb, X, ldx, precond, rtol, atol, max_it, iters_out_host, relres_out_host, work, work_bytes);
}

int sfem_spmm_csr(sfem_ctx* c, int64_t nrows, int64_t nnz, const int64_t* indptr, const int32_t* indices, const double* vals,
                  int k, const double* X, int64_t ldx, double* Y, int64_t ldy) {
    GUARD(c);
    return spmm_csr(c, TAG_SPMM_OTHER, nnz, nrows, indptr, indices, vals, k, X, ldx, Y, ldy, nullptr, nullptr);
}

int sfem_profile_enable(sfem_ctx* c, unsigned mask) {
    GUARD(c);
    prof_reset(c);
    c->prof_mask = mask;
    return SFEM_OK;
}
int sfem_profile_num_tags(void) { return TAG_COUNT; }
const char* sfem_profile_tag_name(int tag) {
    static const char* names[TAG_COUNT] = {"csr_mult_A", "csr_jacobi_A", "csr_resid_A", "spmm_other", "cg_vector",
                                           "mg_transfer_fine", "mg_lattice", "dgemm", "potrf_panel", "gcov_count",
                                           "gcov_fill", "dense_misc", "setup", "chol_trsm", "chol_syrk", "trtri", "lauum"};
    return (tag >= 0 && tag < TAG_COUNT) ? names[tag] : "?";
}
int sfem_profile_read(sfem_ctx* c, double* ms_out_host, double* bytes_out_host, double* flops_out_host, long long* count_out_host) {
    GUARD(c);
    prof_collect(c);
    for (int t = 0; t < TAG_COUNT; ++t) {
        if (ms_out_host) ms_out_host[t] = c->prof_ms[t];
        if (bytes_out_host) bytes_out_host[t] = c->prof_bytes[t];
        if (flops_out_host) flops_out_host[t] = c->prof_flops[t];
        if (count_out_host) count_out_host[t] = c->prof_count[t];
    }
    return SFEM_OK;
}

int sfem_scatter_columns(sfem_ctx* c, int64_t n, const int64_t* colptr, const int32_t* rows, const double* vals,
                         int64_t col_begin, int k, double* B, int64_t ldb) {
    GUARD(c);
    return scatter_columns(c, n, colptr, rows, vals, col_begin, k, B, ldb);
}

int sfem_gcov_count(sfem_ctx* c, int64_t n, int dim, const double* coords, const double* int_basis, double exp_2sigma,
                    double exp_m2l, double cutoff, double regularization, double bmax, double cell_edge,
                    const double* lo_host, const double* hi_host, int64_t* indptr_out, int64_t* nnz_out_host,
                    int64_t* max_row_nnz_out_host, int32_t* borderline_rows, int borderline_cap,
                    int64_t* n_borderline_out_host) {
    GUARD(c);
    return gcov_count(c, n, dim, coords, int_basis, exp_2sigma, exp_m2l, cutoff, regularization, bmax, cell_edge, lo_host,
                      hi_host, indptr_out, nnz_out_host, max_row_nnz_out_host, borderline_rows, borderline_cap,
                      n_borderline_out_host);
}
int sfem_gcov_fill(sfem_ctx* c, const int64_t* indptr, int32_t* indices_out, double* vals_out) {
    GUARD(c);
    return gcov_fill(c, indptr, indices_out, vals_out);
}

int sfem_dgemm(sfem_ctx* c, int opA, int opB, int M, int N, int64_t K, double alpha, const double* A, int64_t lda,
               const double* B, int64_t ldb, double beta, double* C, int64_t ldc) {
    GUARD(c);
    return dgemm(c, opA, opB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, 0);
}
size_t sfem_potrf_dinv_doubles(int n) { return potrf_dinv_doubles(n); }

int sfem_potrf(sfem_ctx* c, int n, double* A, int64_t lda, double* dinv, int* info_out_host) {
    GUARD(c);
    int* info_dev = (int*)sfem_scratch(c, 256);
    if (!info_dev) return sfem_fail(c, SFEM_ERR_CUDA, "potrf: scratch allocation failed");
    int rc = potrf_lower(c, n, A, lda, dinv, info_dev);
    if (rc) return rc;
    int info = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&info, info_dev, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (info_out_host) *info_out_host = info;
    if (info) return sfem_fail(c, SFEM_ERR_NOT_SPD, "potrf: matrix not positive definite (pivot %d)", info);
    return SFEM_OK;
}
int sfem_potrs(sfem_ctx* c, int n, const double* L, int64_t ldl, const double* dinv, int k, double* B, int64_t ldb) {
    GUARD(c);
    return potrs_lower(c, n, L, ldl, dinv, k, B, ldb);
}
int sfem_potri(sfem_ctx* c, int n, const double* L, int64_t ldl, const double* dinv, double* work_nn, double* Kinv) {
    GUARD(c);
    return potri_lower(c, n, L, ldl, dinv, work_nn, Kinv);
}
int sfem_logdet(sfem_ctx* c, int n, const double* L, int64_t ldl, double* out_host) {
    GUARD(c);
    double* dev = (double*)sfem_scratch(c, 256);
    if (!dev) return sfem_fail(c, SFEM_ERR_CUDA, "logdet: scratch allocation failed");
    int rc = logdet_lower(c, n, L, ldl, dev);
    if (rc) return rc;
    CUDA_TRY(c, cudaMemcpyAsync(out_host, dev, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SFEM_OK;
}

int sfem_kernel_matrix(sfem_ctx* c, int n_obs, int dim, const double* Xs, double sigma, double l, const double* unc,
                       int unc_is_vector, double rho2, const double* Cu, int which, double* M_out) {
    GUARD(c);
    return dense_kernel_matrix(c, n_obs, dim, Xs, sigma, l, unc, unc_is_vector, rho2, Cu, which, M_out);
}
int sfem_cov_matrix(sfem_ctx* c, int m, int n, int dim, const double* X1, const double* X2, double sigma, double l, int which,
                    double* M_out) {
    GUARD(c);
    return dense_cov_matrix(c, m, n, dim, X1, X2, sigma, l, which, M_out);
}
int sfem_csr_unsym_count(sfem_ctx* c, int64_t n, const int64_t* indptr, const int32_t* indices, int64_t* row_cnt, int64_t* row_flag,
                         int64_t* eptr_full, int64_t* rowpos, int64_t* nrows_out_host, int64_t* nnz_out_host) {
    GUARD(c);
    return unsym_count(c, n, indptr, indices, row_cnt, row_flag, eptr_full, rowpos, nrows_out_host, nnz_out_host);
}
int sfem_csr_unsym_fill(sfem_ctx* c, int64_t n, const int64_t* indptr, const int32_t* indices, const double* vals,
                        const int64_t* eptr_full, const int64_t* rowpos, int32_t* rows_out, int64_t* eptr_out, int32_t* eidx_out,
                        double* eval_out) {
    GUARD(c);
    return unsym_fill(c, n, indptr, indices, vals, eptr_full, rowpos, rows_out, eptr_out, eidx_out, eval_out);
}
int sfem_rows_axpy(sfem_ctx* c, int64_t nrows, const int32_t* rows, int k, double alpha, const double* X, int64_t ldx, double* Y, int64_t ldy) {
    GUARD(c);
    return rows_axpy(c, nrows, rows, k, alpha, X, ldx, Y, ldy);
}
int sfem_rows_gather(sfem_ctx* c, int64_t nrows, const int32_t* rows, int k, const double* Y, int64_t ldy, double* X, int64_t ldx) {
    GUARD(c);
    return rows_gather(c, nrows, rows, k, Y, ldy, X, ldx);
}
int sfem_dgemm_lower(sfem_ctx* c, int opA, int opB, int n, int64_t K, double alpha, const double* A, int64_t lda, const double* B,
                     int64_t ldb, double beta, double* C, int64_t ldc) {
    GUARD(c);
    if (n < 0 || K < 0 || !A || !B || !C) return sfem_fail(c, SFEM_ERR_ARG, "dgemm_lower: bad arguments");
    return dgemm(c, opA, opB, n, n, K, alpha, A, lda, B, ldb, beta, C, ldc, 1);
}
int sfem_tile_apply_one(sfem_ctx* c, int level, int which, int mode, int k, const double* X, const double* B, double* Y,
                        double* partials, int* nparts_out_host, int use_tensor_form, int64_t* rows_out_host,
                        int64_t* cols_out_host, int* tensor_form_out_host) {
    GUARD(c);
    return mg_tile_apply_one(c, level, which, mode, k, X, B, Y, partials, nparts_out_host, use_tensor_form, rows_out_host,
                             cols_out_host, tensor_form_out_host);
}
int sfem_b84_build(sfem_ctx* c, int64_t n, const int64_t* indptr, const int32_t* indices, const double* vals, void** matrix_out,
                   int64_t* ntiles_out_host) {
    GUARD(c);
    if (!matrix_out) return sfem_fail(c, SFEM_ERR_ARG, "b84_build: no output handle");
    B84Matrix* m = nullptr;
    int rc = b84_build(c, n, indptr, indices, vals, &m);
    *matrix_out = m;
    if (ntiles_out_host) *ntiles_out_host = b84_num_tiles(m);
    return rc;
}
int sfem_b84_spmm(sfem_ctx* c, const void* matrix, int k, const double* X, int64_t ldx, double* Y, int64_t ldy) {
    GUARD(c);
    return b84_spmm(c, static_cast<const B84Matrix*>(matrix), k, X, ldx, Y, ldy);
}
int sfem_b84_free(sfem_ctx* c, void* matrix) {
    GUARD(c);
    b84_free(static_cast<B84Matrix*>(matrix));
    return SFEM_OK;
}

int sfem_mirror_lower(sfem_ctx* c, int n, double* A, int64_t lda) {
    GUARD(c);
    return mirror_lower(c, n, A, lda);
}

int sfem_nccl_unique_id(void* id_out_128_bytes_host) { return sharded_unique_id(id_out_128_bytes_host); }
int sfem_comm_init(sfem_ctx* c, const void* unique_id_128_bytes_host, int rank, int nranks) {
    GUARD(c);
    return sharded_comm_init(c, unique_id_128_bytes_host, rank, nranks);
}
int sfem_comm_attach(sfem_ctx* c, void* nccl_comm, int rank, int nranks) {
    GUARD(c);
    return sharded_comm_attach(c, nccl_comm, rank, nranks);
}
int sfem_comm_destroy(sfem_ctx* c) {
    GUARD(c);
    return sharded_comm_destroy(c);
}
int sfem_comm_bcast(sfem_ctx* c, void* device_buffer, size_t bytes, int root) {
    GUARD(c);
    return sharded_bcast(c, device_buffer, bytes, root);
}
int sfem_prior_cov(sfem_ctx* c, int64_t n_obs, const int64_t* P_colptr, const int32_t* P_rows, const double* P_vals,
                   const int64_t* G_indptr, const int32_t* G_indices, const double* G_vals, int64_t col_begin,
                   int64_t col_end, int block_cols, int precond, double rtol, double atol, int max_it, double* Z_out,
                   double* W_out, double* Cu_out, int* iters_max_out_host) {
    GUARD(c);
    return sharded_prior_cov(c, n_obs, P_colptr, P_rows, P_vals, G_indptr, G_indices, G_vals, col_begin, col_end, block_cols,
                             precond, rtol, atol, max_it, Z_out, W_out, Cu_out, iters_max_out_host);
}

int sfem_sym_fold(sfem_ctx* c, int n, int nblocks, double* A, int64_t lda) {
    GUARD(c);
    if (n < 0 || nblocks < 1 || !A) return sfem_fail(c, SFEM_ERR_ARG, "sym_fold: bad arguments");
    return sym_fold(c, n, nblocks, A, lda);
}

int sfem_axpby(sfem_ctx* c, int64_t n, double a, const double* x, double b, const double* y, double* z) {
    GUARD(c);
    return dense_axpby(c, n, a, x, b, y, z);
}
int sfem_skinny_gemm(sfem_ctx* c, int64_t n, int k, int ns, const double* X, int64_t ldx, const double* Cm, int64_t ldc,
                     double* Y, int64_t ldy) {
    GUARD(c);
    return dense_skinny_gemm(c, n, k, ns, X, ldx, Cm, ldc, Y, ldy);
}
size_t sfem_dense_workspace_bytes(int n_obs) { return dense_workspace_bytes(n_obs); }

int sfem_logpost(sfem_ctx* c, int n_obs, int dim, const double* Xs, const double* y, const double* unc, int unc_is_vector,
                 const double* mu, const double* Cu, const double* theta_host, double* value_out_host, double* grad_out_host,
                 void* work, size_t work_bytes) {
    GUARD(c);
    return dense_logpost(c, n_obs, dim, Xs, y, unc, unc_is_vector, mu, Cu, theta_host, value_out_host, grad_out_host, work,
                         work_bytes);
}
int sfem_posterior_cov(sfem_ctx* c, int n_obs, int dim, const double* Xs, const double* y, const double* unc,
                       int unc_is_vector, const double* mu, const double* Cu, const double* theta_host, double* muy_out,
                       double* Cuy_out, void* work, size_t work_bytes) {
    GUARD(c);
    return dense_posterior_cov(c, n_obs, dim, Xs, y, unc, unc_is_vector, mu, Cu, theta_host, muy_out, Cuy_out, work,
                               work_bytes);
}
int sfem_dense_solve(sfem_ctx* c, int n_obs, int dim, const double* Xs, const double* unc, int unc_is_vector, double sigma,
                     double l, double rho2, const double* Cu, const double* b, double* x, void* work, size_t work_bytes) {
    GUARD(c);
    return dense_solve_shifted(c, n_obs, dim, Xs, unc, unc_is_vector, sigma, l, rho2, Cu, b, x, work, work_bytes);
}

}  // extern "C"
