// Structure of the forcing covariance used by the covariance assembly:  G = S + E  with S the entries (i, j) whose
// transpose (j, i) is also stored -- their values b_i b_j c(r_ij) are bitwise symmetric -- and E the few entries
// without a stored transpose (the reference's row-relative threshold, ForcingCovariance.py:160-164, keeps (i, j) and
// drops (j, i) only where the lumped weights b differ, i.e. next to the boundary).  Then
//     (G Z)^T Z = (S Z)^T Z  +  (E Z)^T Z,
// the first term is symmetric (only the tiles on and below the diagonal are computed, then mirrored: half the DMMA
// work of solving_utils.py:139-142's n_obs^2 inner products) and the second only sums over the rows where E is not
// empty (a boundary layer of the mesh).  These kernels extract E in compact-row CSR form and move row subsets.
#include "rowops.cuh"

template <typename T>
int exclusive_scan(sfem_ctx* c, const T* in, int64_t n, T* out, T* sums);

namespace {

__device__ __forceinline__ bool has_entry(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx, int row, int col) {
    int64_t lo = ptr[row], hi = ptr[row + 1];
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        int v = idx[mid];
        if (v < col) lo = mid + 1;
        else hi = mid;
    }
    return lo < ptr[row + 1] && idx[lo] == col;
}

// warp per row: number of entries without a stored transpose
__global__ void __launch_bounds__(256) unsym_count_kernel(int64_t n, const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                                          int64_t* __restrict__ row_cnt, int64_t* __restrict__ row_flag) {
    int lane = threadIdx.x & 31;
    int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    int cnt = 0;
    for (int64_t p = ptr[i] + lane; p < ptr[i + 1]; p += 32) {
        int j = idx[p];
        if (j != i && !has_entry(ptr, idx, j, (int)i)) ++cnt;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) { row_cnt[i] = cnt; row_flag[i] = cnt > 0 ? 1 : 0; }
}

__global__ void __launch_bounds__(256) unsym_fill_kernel(int64_t n, const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                                         const double* __restrict__ val, const int64_t* __restrict__ eptr_full,
                                                         const int64_t* __restrict__ rowpos, int32_t* __restrict__ rows_out,
                                                         int64_t* __restrict__ eptr_out, int32_t* __restrict__ eidx_out,
                                                         double* __restrict__ eval_out) {
    int lane = threadIdx.x & 31;
    int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    int64_t base = eptr_full[i];
    if (eptr_full[i + 1] == base) {
        if (i == n - 1 && lane == 0) eptr_out[rowpos[n]] = eptr_full[n];
        return;
    }
    int64_t r = rowpos[i];
    if (lane == 0) {
        rows_out[r] = (int)i;
        eptr_out[r] = base;
        if (i == n - 1) eptr_out[rowpos[n]] = eptr_full[n];
    }
    int filled = 0;
    unsigned lt = (1u << lane) - 1u;
    int64_t p0 = ptr[i], p1 = ptr[i + 1];
    for (int64_t q = p0; q < p1; q += 32) {       // ascending order kept: ballot prefix inside each group of 32
        int64_t p = q + lane;
        bool take = false;
        int j = 0;
        if (p < p1) {
            j = idx[p];
            take = (j != i) && !has_entry(ptr, idx, j, (int)i);
        }
        unsigned m = __ballot_sync(0xffffffffu, take);
        if (take) {
            int pos = filled + __popc(m & lt);
            eidx_out[base + pos] = j;
            eval_out[base + pos] = val[p];
        }
        filled += __popc(m);
    }
}

// Y[rows[r]] += alpha X[r]   /   X[r] = Y[rows[r]]      (row-major blocks, k columns)
__global__ void rows_axpy_kernel(int64_t nrows, const int32_t* __restrict__ rows, int k, double alpha, const double* __restrict__ X,
                                 int64_t ldx, double* __restrict__ Y, int64_t ldy) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nrows * k) return;
    int64_t r = t / k;
    int cidx = (int)(t - r * k);
    Y[(int64_t)rows[r] * ldy + cidx] += alpha * X[r * ldx + cidx];
}
__global__ void rows_gather_kernel(int64_t nrows, const int32_t* __restrict__ rows, int k, const double* __restrict__ Y, int64_t ldy,
                                   double* __restrict__ X, int64_t ldx) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nrows * k) return;
    int64_t r = t / k;
    int cidx = (int)(t - r * k);
    X[r * ldx + cidx] = Y[(int64_t)rows[r] * ldy + cidx];
}
__global__ void mirror_lower_tiles_kernel(int n, double* __restrict__ A, int64_t ld) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * n) return;
    int i = (int)(t / n), j = (int)(t % n);
    if (j > i) A[(int64_t)i * ld + j] = A[(int64_t)j * ld + i];
}

// Sharded symmetric assembly: the n x n array holds every unordered pair of DIFFERENT ownership blocks exactly once
// (block (r, s) or block (s, r), the other one zero); fold it into the full symmetric matrix.  Ownership is PETSc's
// contiguous rule (the first `extra` blocks have base + 1 entries).  Diagonal blocks are left as they are.
__device__ __forceinline__ int owner_block(int i, int base, int extra) {
    int cut = extra * (base + 1);
    return i < cut ? i / (base + 1) : extra + (base > 0 ? (i - cut) / base : 0);
}
__global__ void sym_fold_kernel(int n, int base, int extra, double* __restrict__ A, int64_t ld) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * n) return;
    int i = (int)(t / n), j = (int)(t % n);
    if (j >= i || owner_block(i, base, extra) == owner_block(j, base, extra)) return;
    double v = A[(int64_t)i * ld + j] + A[(int64_t)j * ld + i];
    A[(int64_t)i * ld + j] = v;
    A[(int64_t)j * ld + i] = v;
}

}  // namespace

int sym_fold(sfem_ctx* c, int n, int nblocks, double* A, int64_t ld) {
    if (n <= 0 || nblocks <= 1) return SFEM_OK;
    ProfScope ps(c, TAG_DENSE_MISC, 16.0 * (double)n * n, 0.0);
    sym_fold_kernel<<<(unsigned)cdiv64((int64_t)n * n, 256), 256, 0, c->stream>>>(n, n / nblocks, n % nblocks, A, ld);
    KCHECK(c);
    return SFEM_OK;
}

// Pass 1.  row_cnt, row_flag: device scratch of n int64 each; eptr_full, rowpos: device, n + 1 int64 each.
int unsym_count(sfem_ctx* c, int64_t n, const int64_t* ptr, const int32_t* idx, int64_t* row_cnt, int64_t* row_flag,
                int64_t* eptr_full, int64_t* rowpos, int64_t* nrows_host, int64_t* nnz_host) {
    ProfScope ps(c, TAG_SETUP, 0.0, 0.0);
    unsym_count_kernel<<<(unsigned)cdiv64(n, 8), 256, 0, c->stream>>>(n, ptr, idx, row_cnt, row_flag);
    KCHECK(c);
    int64_t* sums = (int64_t*)sfem_scratch(c, sizeof(int64_t) * (size_t)(n / 1024 + 8));
    if (!sums) return sfem_fail(c, SFEM_ERR_CUDA, "unsym_count: scratch allocation failed");
    int rc = exclusive_scan<int64_t>(c, row_cnt, n, eptr_full, sums);
    if (rc) return rc;
    rc = exclusive_scan<int64_t>(c, row_flag, n, rowpos, sums);
    if (rc) return rc;
    CUDA_TRY(c, cudaMemcpyAsync(nnz_host, eptr_full + n, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(nrows_host, rowpos + n, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SFEM_OK;
}

int unsym_fill(sfem_ctx* c, int64_t n, const int64_t* ptr, const int32_t* idx, const double* val, const int64_t* eptr_full,
               const int64_t* rowpos, int32_t* rows_out, int64_t* eptr_out, int32_t* eidx_out, double* eval_out) {
    ProfScope ps(c, TAG_SETUP, 0.0, 0.0);
    unsym_fill_kernel<<<(unsigned)cdiv64(n, 8), 256, 0, c->stream>>>(n, ptr, idx, val, eptr_full, rowpos, rows_out, eptr_out,
                                                                    eidx_out, eval_out);
    KCHECK(c);
    return SFEM_OK;
}

int rows_axpy(sfem_ctx* c, int64_t nrows, const int32_t* rows, int k, double alpha, const double* X, int64_t ldx, double* Y, int64_t ldy) {
    if (nrows <= 0 || k <= 0) return SFEM_OK;
    ProfScope ps(c, TAG_SPMM_OTHER, 24.0 * (double)nrows * k, 0.0);
    rows_axpy_kernel<<<(unsigned)cdiv64(nrows * k, 256), 256, 0, c->stream>>>(nrows, rows, k, alpha, X, ldx, Y, ldy);
    KCHECK(c);
    return SFEM_OK;
}
int rows_gather(sfem_ctx* c, int64_t nrows, const int32_t* rows, int k, const double* Y, int64_t ldy, double* X, int64_t ldx) {
    if (nrows <= 0 || k <= 0) return SFEM_OK;
    ProfScope ps(c, TAG_SPMM_OTHER, 16.0 * (double)nrows * k, 0.0);
    rows_gather_kernel<<<(unsigned)cdiv64(nrows * k, 256), 256, 0, c->stream>>>(nrows, rows, k, Y, ldy, X, ldx);
    KCHECK(c);
    return SFEM_OK;
}
int mirror_lower(sfem_ctx* c, int n, double* A, int64_t ld) {
    ProfScope ps(c, TAG_DENSE_MISC, 8.0 * (double)n * n, 0.0);
    mirror_lower_tiles_kernel<<<(unsigned)cdiv64((int64_t)n * n, 256), 256, 0, c->stream>>>(n, A, ld);
    KCHECK(c);
    return SFEM_OK;
}
