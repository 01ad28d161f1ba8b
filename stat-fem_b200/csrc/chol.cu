// K10-K12: blocked Cholesky (lower, row-major, in place), triangular solves, inverse and log-determinant, FP64.
//
// Right-looking with 128-wide panels: the 128x128 diagonal block is factorised by one CTA in shared memory, which
// also forms the explicit inverse of the triangular block; the panel solve and every trailing update are then plain
// DMMA GEMMs (dgemm.cu), as are the block forward/backward substitutions of potrs and the two triangular products
// of potri.  Replaces SciPy/LAPACK cho_factor / cho_solve at LinearSolver.py:273-297, 353-367, 603-607, 662-677.
#include "common.cuh"

namespace {

constexpr int NB = 128;
constexpr int LDS = NB + 1;

// Factor the nb x nb (nb <= 128) lower block at A (row-major, lda) in place and write inv(L) to Dinv (ld = 128,
// explicit zeros above the diagonal and outside nb).  info: 0, or 1-based global index of the first bad pivot.
//
// One CTA, the block lives in shared memory with an odd row stride (129) so that "thread = row" accesses are
// conflict-free and column reads are broadcasts.  Two threads share a row and split the k-range by parity.
// Phase 1: right-looking Cholesky, two barriers per column.  Phase 2: in-place inversion of the triangular factor,
// columns from last to first:  X[i][j] = -(1/L[j][j]) * sum_{k=j+1..i} X[i][k] L[k][j].
__global__ void __launch_bounds__(256) potrf_panel_kernel(double* __restrict__ A, int64_t lda, int nb, int j0,
                                                          double* __restrict__ Dinv, int* __restrict__ info) {
    extern __shared__ double s[];   // [NB][LDS] + column buffer [NB] + pivots [NB]
    double* colbuf = s + NB * LDS;
    const int tid = threadIdx.x;
    const int i = tid >> 1, h = tid & 1;
    for (int t = tid; t < NB * NB; t += blockDim.x) {
        int r = t >> 7, cidx = t & (NB - 1);
        s[r * LDS + cidx] = (r < nb && cidx <= r) ? A[(int64_t)r * lda + cidx] : ((r == cidx) ? 1.0 : 0.0);
    }
    __syncthreads();
    // ---- phase 1: Cholesky.  The square roots of the pivots go to pivs[] and are put on the diagonal at the end, so
    // the unscaled s[j][j] every thread reads is never overwritten while a slower warp may still be reading it.
    double* pivs = colbuf + NB;
    for (int j = 0; j < nb; ++j) {
        __syncthreads();                       // trailing update of column j-1 complete everywhere
        double dg = s[j * LDS + j];
        if (!(dg > 0.0)) {
            if (tid == 0 && *info == 0) *info = j0 + j + 1;
            dg = 1.0;
        }
        double piv = sqrt(dg);
        double lij = (i > j) ? s[i * LDS + j] / piv : 0.0;
        __syncwarp();                          // both threads of a row have read s[i][j] before it is rescaled
        if (h == 0) {
            if (i > j) s[i * LDS + j] = lij;
            else if (i == j) pivs[j] = piv;
        }
        __syncthreads();                       // scaled column j visible to all rows
        if (i > j) {
            double* row = s + i * LDS;
            for (int k = j + 1 + h; k <= i; k += 2) row[k] -= lij * s[k * LDS + j];
        }
    }
    __syncthreads();
    if (tid < nb) s[tid * LDS + tid] = pivs[tid];
    __syncthreads();
    for (int t = tid; t < nb * nb; t += blockDim.x) {
        int r = t / nb, cidx = t % nb;
        if (cidx <= r) A[(int64_t)r * lda + cidx] = s[r * LDS + cidx];
    }
    __syncthreads();
    // ---- phase 2: X = inv(L) in place (rows/cols >= nb hold the identity, harmless)
    for (int j = NB - 1; j >= 0; --j) {
        if (tid < NB) colbuf[tid] = s[tid * LDS + j];      // old column j of L (entries k >= j)
        __syncthreads();
        double xjj = 1.0 / colbuf[j];
        double acc = 0.0;
        if (i > j) {
            const double* row = s + i * LDS;
            for (int k = j + 1 + h; k <= i; k += 2) acc += row[k] * colbuf[k];
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (h == 0) {
            if (i > j) s[i * LDS + j] = -acc * xjj;
            else if (i == j) s[j * LDS + j] = xjj;
        }
        __syncthreads();
    }
    for (int t = tid; t < NB * NB; t += blockDim.x) {
        int r = t >> 7, cidx = t & (NB - 1);
        Dinv[t] = (r < nb && cidx <= r && cidx < nb) ? s[r * LDS + cidx] : 0.0;
    }
}

__global__ void logdet_kernel(int n, const double* __restrict__ L, int64_t ldl, double* __restrict__ out) {
    __shared__ double sh[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += log(L[(int64_t)i * ldl + i]);
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = 2.0 * sh[0];
}

__global__ void set_identity_kernel(int n, double* __restrict__ W, int64_t ld) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * n) return;
    int i = (int)(t / n), j = (int)(t % n);
    W[(int64_t)i * ld + j] = (i == j) ? 1.0 : 0.0;
}

__global__ void mirror_lower_kernel(int n, double* __restrict__ A, int64_t ld) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * n) return;
    int i = (int)(t / n), j = (int)(t % n);
    if (j > i) A[(int64_t)i * ld + j] = A[(int64_t)j * ld + i];
}

}  // namespace

// dinv: device buffer of ceil(n/128) * 128*128 doubles.  info_dev: device int (zeroed here).
int potrf_lower(sfem_ctx* c, int n, double* A, int64_t lda, double* dinv, int* info_dev) {
    CUDA_TRY(c, cudaMemsetAsync(info_dev, 0, sizeof(int), c->stream));
    size_t dyn = sizeof(double) * (NB * LDS + 2 * NB);
    cudaFuncSetAttribute(potrf_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    for (int j0 = 0; j0 < n; j0 += NB) {
        int nb = imin(NB, n - j0);
        double* Ajj = A + (int64_t)j0 * lda + j0;
        double* Dj = dinv + (size_t)(j0 / NB) * NB * NB;
        {
            ProfScope ps(c, TAG_POTRF_PANEL, 24.0 * nb * nb, (2.0 / 3.0) * nb * nb * nb);
            potrf_panel_kernel<<<1, 256, dyn, c->stream>>>(Ajj, lda, nb, j0, Dj, info_dev);
            KCHECK(c);
        }
        int m = n - j0 - nb;
        if (m > 0) {
            double* A21 = A + (int64_t)(j0 + nb) * lda + j0;
            double* A22 = A + (int64_t)(j0 + nb) * lda + (j0 + nb);
            int rc = dgemm(c, OP_N, OP_T, m, nb, nb, 1.0, A21, lda, Dj, NB, 0.0, A21, lda, 0);   // A21 <- A21 inv(L)^T
            if (rc) return rc;
            rc = dgemm(c, OP_N, OP_T, m, m, nb, -1.0, A21, lda, A21, lda, 1.0, A22, lda, 1);     // SYRK, lower tiles
            if (rc) return rc;
        }
    }
    return SFEM_OK;
}

// B (n x k, ldb) <- inv(L L^T) B.  ncols_tri > 0 limits the forward pass to a lower-trapezoidal RHS (potri).
int potrs_lower(sfem_ctx* c, int n, const double* L, int64_t ldl, const double* dinv, int k, double* B, int64_t ldb) {
    // forward: L Y = B
    for (int j0 = 0; j0 < n; j0 += NB) {
        int nb = imin(NB, n - j0);
        const double* Dj = dinv + (size_t)(j0 / NB) * NB * NB;
        double* Bj = B + (int64_t)j0 * ldb;
        int rc = dgemm(c, OP_N, OP_N, nb, k, nb, 1.0, Dj, NB, Bj, ldb, 0.0, Bj, ldb, 0);
        if (rc) return rc;
        int m = n - j0 - nb;
        if (m > 0) {
            rc = dgemm(c, OP_N, OP_N, m, k, nb, -1.0, L + (int64_t)(j0 + nb) * ldl + j0, ldl, Bj, ldb, 1.0,
                       B + (int64_t)(j0 + nb) * ldb, ldb, 0);
            if (rc) return rc;
        }
    }
    // backward: L^T X = Y
    int last = ((n - 1) / NB) * NB;
    for (int j0 = last; j0 >= 0; j0 -= NB) {
        int nb = imin(NB, n - j0);
        const double* Dj = dinv + (size_t)(j0 / NB) * NB * NB;
        double* Bj = B + (int64_t)j0 * ldb;
        int rc = dgemm(c, OP_T, OP_N, nb, k, nb, 1.0, Dj, NB, Bj, ldb, 0.0, Bj, ldb, 0);
        if (rc) return rc;
        if (j0 > 0) {
            rc = dgemm(c, OP_T, OP_N, j0, k, nb, -1.0, L + (int64_t)j0 * ldl, ldl, Bj, ldb, 1.0, B, ldb, 0);
            if (rc) return rc;
        }
    }
    return SFEM_OK;
}

// Kinv (n x n, ld n) = inv(L L^T); W is an n x n workspace (receives inv(L)).
int potri_lower(sfem_ctx* c, int n, const double* L, int64_t ldl, const double* dinv, double* W, double* Kinv) {
    int64_t tot = (int64_t)n * n;
    set_identity_kernel<<<(unsigned)cdiv64(tot, 256), 256, 0, c->stream>>>(n, W, n);
    KCHECK(c);
    for (int j0 = 0; j0 < n; j0 += NB) {
        int nb = imin(NB, n - j0);
        int ncols = j0 + nb;    // block row j of inv(L) is non-zero only in its first j0+nb columns
        const double* Dj = dinv + (size_t)(j0 / NB) * NB * NB;
        double* Wj = W + (int64_t)j0 * n;
        int rc = dgemm(c, OP_N, OP_N, nb, ncols, nb, 1.0, Dj, NB, Wj, n, 0.0, Wj, n, 0);
        if (rc) return rc;
        int m = n - j0 - nb;
        if (m > 0) {
            rc = dgemm(c, OP_N, OP_N, m, ncols, nb, -1.0, L + (int64_t)(j0 + nb) * ldl + j0, ldl, Wj, n, 1.0,
                       W + (int64_t)(j0 + nb) * n, n, 0);
            if (rc) return rc;
        }
    }
    int rc = dgemm(c, OP_T, OP_N, n, n, n, 1.0, W, n, W, n, 0.0, Kinv, n, 3);   // lower tiles, k >= max(m0, n0)
    if (rc) return rc;
    mirror_lower_kernel<<<(unsigned)cdiv64(tot, 256), 256, 0, c->stream>>>(n, Kinv, n);
    KCHECK(c);
    return SFEM_OK;
}

int logdet_lower(sfem_ctx* c, int n, const double* L, int64_t ldl, double* out_dev) {
    logdet_kernel<<<1, 256, 0, c->stream>>>(n, L, ldl, out_dev);
    KCHECK(c);
    return SFEM_OK;
}

size_t potrf_dinv_doubles(int n) { return (size_t)((n + NB - 1) / NB) * NB * NB; }
