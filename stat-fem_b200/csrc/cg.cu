// K2 / K3 + driver: multi-RHS preconditioned conjugate gradients  A X = B  on row-major blocks [n x k].
//
// Every column runs its own CG recurrence (own alpha/beta) but all columns advance in lock-step through the same
// kernels, so A is streamed once per iteration for the whole block and every vector pass is a fully coalesced
// [n x k] sweep.  Replaces the 2*n_obs sequential PETSc KSP solves of solving_utils.py:49-53 / 139-142 and the mean
// solve of LinearSolver.py:217.  Preconditioner: Jacobi (precond = 0) or the lattice multigrid V-cycle of mg.cu
// (precond = 1).  All reductions are two-stage and summed in a fixed order (no floating-point atomics).
#include "mg.cuh"

int csr_residual(sfem_ctx* c, int64_t n, const int64_t* ptr, const int32_t* idx, const double* val, int k,
                 const double* X, int64_t ldx, double* R, int64_t ldr, const double* B, int64_t ldb);

namespace {

struct Scal {            // per-column scalars, each an array of k doubles
    double *rho, *alpha, *beta, *bb, *rr, *tol2;
    int* active;         // k ints
    int* n_active;       // 2 ints: [0] columns still iterating, [1] set when a norm is NaN/Inf or the true residual fails
};

// init: r = B, x = 0, partial bb = sum b^2;  JAC: z = dinv r, p = z, partial rho = sum r z
template <int VEC, bool JAC>
__global__ void __launch_bounds__(RB_THREADS) cg_init_kernel(int64_t n, int k, const double* __restrict__ B, int64_t ldb,
                                                            double* __restrict__ X, int64_t ldx, double* __restrict__ R,
                                                            double* __restrict__ P, const double* __restrict__ dinv,
                                                            double* __restrict__ part_bb, double* __restrict__ part_rho,
                                                            int64_t rows_per_block) {
    __shared__ double red[RB_WARPS][64];
    const int warp = threadIdx.x >> 5;
    LaneCols<VEC> L(k);
    int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block < n ? r0 + rows_per_block : n;
    double s0 = 0, s1 = 0, t0 = 0, t1 = 0;
    for (int64_t row = r0 + warp; row < r1; row += RB_WARPS) {
        double b0, b1;
        ld_cols<VEC>(B + row * ldb, L, b0, b1);
        st_cols<VEC>(R + row * k, L, b0, b1);
        st_cols<VEC>(X + row * ldx, L, 0.0, 0.0);
        s0 += b0 * b0; s1 += b1 * b1;
        if (JAC) {
            double di = dinv[row];
            double z0 = di * b0, z1 = di * b1;
            st_cols<VEC>(P + row * k, L, z0, z1);
            t0 += b0 * z0; t1 += b1 * z1;
        }
    }
    block_partials<VEC>(s0, s1, L, k, part_bb, red);
    if (JAC) block_partials<VEC>(t0, t1, L, k, part_rho, red);
}

// p = z, partial rho = sum r z    (after the first preconditioner application, MG path)
template <int VEC>
__global__ void __launch_bounds__(RB_THREADS) cg_pz_kernel(int64_t n, int k, const double* __restrict__ R,
                                                          const double* __restrict__ Z, double* __restrict__ P,
                                                          const double* __restrict__ beta, double* __restrict__ part_rho,
                                                          int64_t rows_per_block) {
    __shared__ double red[RB_WARPS][64];
    const int warp = threadIdx.x >> 5;
    LaneCols<VEC> L(k);
    int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block < n ? r0 + rows_per_block : n;
    double t0 = 0, t1 = 0;
    for (int64_t row = r0 + warp; row < r1; row += RB_WARPS) {
        double a0, a1, z0, z1;
        ld_cols<VEC>(R + row * k, L, a0, a1);
        ld_cols<VEC>(Z + row * k, L, z0, z1);
        t0 += a0 * z0; t1 += a1 * z1;
        if (P) st_cols<VEC>(P + row * k, L, z0, z1);
    }
    block_partials<VEC>(t0, t1, L, k, part_rho, red);
}

// x += alpha p ; r -= alpha q ; partial rr = sum r^2 ; JAC: partial rz = sum r dinv r
template <int VEC, bool JAC>
__global__ void __launch_bounds__(RB_THREADS) cg_update_kernel(int64_t n, int k, const double* __restrict__ alpha,
                                                              const double* __restrict__ P, const double* __restrict__ Q,
                                                              double* __restrict__ X, int64_t ldx, double* __restrict__ R,
                                                              const double* __restrict__ dinv, double* __restrict__ part_rr,
                                                              double* __restrict__ part_rz, int64_t rows_per_block) {
    __shared__ double red[RB_WARPS][64];
    const int warp = threadIdx.x >> 5;
    LaneCols<VEC> L(k);
    double al0 = L.ok0 ? alpha[L.c] : 0.0, al1 = L.ok1 ? alpha[L.c + 1] : 0.0;
    int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block < n ? r0 + rows_per_block : n;
    double s0 = 0, s1 = 0, t0 = 0, t1 = 0;
    for (int64_t row = r0 + warp; row < r1; row += RB_WARPS) {
        double p0, p1, q0, q1, x0, x1, a0, a1;
        ld_cols<VEC>(P + row * k, L, p0, p1);
        ld_cols<VEC>(Q + row * k, L, q0, q1);
        ld_cols<VEC>(X + row * ldx, L, x0, x1);
        ld_cols<VEC>(R + row * k, L, a0, a1);
        x0 += al0 * p0; x1 += al1 * p1;
        a0 -= al0 * q0; a1 -= al1 * q1;
        st_cols<VEC>(X + row * ldx, L, x0, x1);
        st_cols<VEC>(R + row * k, L, a0, a1);
        s0 += a0 * a0; s1 += a1 * a1;
        if (JAC) {
            double di = dinv[row];
            t0 += a0 * di * a0; t1 += a1 * di * a1;
        }
    }
    block_partials<VEC>(s0, s1, L, k, part_rr, red);
    if (JAC) block_partials<VEC>(t0, t1, L, k, part_rz, red);
}

// p = z + beta p   (JAC: z = dinv r on the fly)
template <int VEC, bool JAC>
__global__ void __launch_bounds__(RB_THREADS) cg_pupdate_kernel(int64_t n, int k, const double* __restrict__ beta,
                                                               const double* __restrict__ ZorR, const double* __restrict__ dinv,
                                                               double* __restrict__ P, int64_t rows_per_block) {
    const int warp = threadIdx.x >> 5;
    LaneCols<VEC> L(k);
    double be0 = L.ok0 ? beta[L.c] : 0.0, be1 = L.ok1 ? beta[L.c + 1] : 0.0;
    int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block < n ? r0 + rows_per_block : n;
    for (int64_t row = r0 + warp; row < r1; row += RB_WARPS) {
        double z0, z1, p0, p1;
        ld_cols<VEC>(ZorR + row * k, L, z0, z1);
        ld_cols<VEC>(P + row * k, L, p0, p1);
        if (JAC) {
            double di = dinv[row];
            z0 *= di; z1 *= di;
        }
        st_cols<VEC>(P + row * k, L, z0 + be0 * p0, z1 + be1 * p1);
    }
}

// out[c] = sum_blk part[blk][c]  (fixed order)
__device__ __forceinline__ double colsum(const double* __restrict__ part, int nblk, int k, int c) {
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += part[(int64_t)b * k + c];
    return s;
}

__global__ void cg_scal_init_kernel(int k, int nblk, const double* __restrict__ part_bb, const double* __restrict__ part_rho,
                                    Scal s, double rtol, double atol) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= k) return;
    double bb = colsum(part_bb, nblk, k, c);
    s.bb[c] = bb;
    s.rr[c] = bb;
    double t = rtol * rtol * bb;
    if (atol * atol > t) t = atol * atol;
    s.tol2[c] = t;
    int act = (bb > t) ? 1 : 0;
    if (!(fabs(bb) <= 1.7e308)) { act = 0; atomicOr(s.n_active + 1, 1); }      // NaN / Inf right-hand side: fail, never "converged"
    s.active[c] = act;
    if (part_rho) s.rho[c] = colsum(part_rho, nblk, k, c);
    if (act) atomicAdd(s.n_active, 1);
}

__global__ void cg_rho_kernel(int k, int nblk, const double* __restrict__ part_rho, Scal s) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < k) s.rho[c] = colsum(part_rho, nblk, k, c);
}

__global__ void cg_alpha_kernel(int k, int nblk, const double* __restrict__ part_sigma, Scal s) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= k) return;
    double sg = colsum(part_sigma, nblk, k, c);
    s.alpha[c] = (s.active[c] && sg > 0.0) ? s.rho[c] / sg : 0.0;
}

// rr, new rho, beta, convergence flags
__global__ void cg_beta_kernel(int k, int nblk, const double* __restrict__ part_rr, const double* __restrict__ part_rz, Scal s) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= k) return;
    int act = s.active[c];
    if (act) {
        double rr = colsum(part_rr, nblk, k, c);
        s.rr[c] = rr;
        if (!(fabs(rr) <= 1.7e308)) atomicOr(s.n_active + 1, 1);               // breakdown: reported, not taken for convergence
        if (!(rr > s.tol2[c])) act = 0;
    }
    double rho_new = colsum(part_rz, nblk, k, c);
    double rho_old = s.rho[c];
    s.beta[c] = (act && rho_old > 0.0) ? rho_new / rho_old : 0.0;
    if (act) s.rho[c] = rho_new;
    s.active[c] = act;
    if (act) atomicAdd(s.n_active, 1);
}

// true relative residual per column; flags[1] |= 2 when it is not finite or misses the requested tolerance by more
// than a factor RELRES_SLACK (the recurrence residual drifted from the true one)
#define RELRES_SLACK 100.0
__global__ void cg_relres_kernel(int k, int nblk, const double* __restrict__ part_rr, const double* __restrict__ bb,
                                 const double* __restrict__ tol2, int* __restrict__ flags, double* __restrict__ out) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= k) return;
    double rr = colsum(part_rr, nblk, k, c);
    out[c] = bb[c] > 0.0 ? sqrt(rr / bb[c]) : sqrt(rr);
    double lim = RELRES_SLACK * RELRES_SLACK * tol2[c];
    if (lim < 1.e-26 * bb[c]) lim = 1.e-26 * bb[c];                           // rounding floor of an FP64 residual
    if (!(rr <= lim)) atomicOr(flags + 1, 2);
}

// partial rr of an existing residual block
template <int VEC>
__global__ void __launch_bounds__(RB_THREADS) cg_norm_kernel(int64_t n, int k, const double* __restrict__ R,
                                                            double* __restrict__ part_rr, int64_t rows_per_block) {
    __shared__ double red[RB_WARPS][64];
    const int warp = threadIdx.x >> 5;
    LaneCols<VEC> L(k);
    int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block < n ? r0 + rows_per_block : n;
    double s0 = 0, s1 = 0;
    for (int64_t row = r0 + warp; row < r1; row += RB_WARPS) {
        double a0, a1;
        ld_cols<VEC>(R + row * k, L, a0, a1);
        s0 += a0 * a0; s1 += a1 * a1;
    }
    block_partials<VEC>(s0, s1, L, k, part_rr, red);
}

// tile path init: r = B[perm] (rows gathered into the tile numbering), x = 0, p = 0, partial bb
template <int VEC>
__global__ void __launch_bounds__(RB_THREADS) cg_init_perm_kernel(int64_t n, int k, const int* __restrict__ perm,
                                                                 const double* __restrict__ B, int64_t ldb,
                                                                 double* __restrict__ X, double* __restrict__ R,
                                                                 double* __restrict__ P, double* __restrict__ part_bb,
                                                                 int64_t rows_per_block) {
    __shared__ double red[RB_WARPS][64];
    const int warp = threadIdx.x >> 5;
    LaneCols<VEC> L(k);
    int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block < n ? r0 + rows_per_block : n;
    double s0 = 0, s1 = 0;
    for (int64_t row = r0 + warp; row < r1; row += RB_WARPS) {
        double b0, b1;
        ld_cols<VEC>(B + (int64_t)perm[row] * ldb, L, b0, b1);
        st_cols<VEC>(R + row * k, L, b0, b1);
        st_cols<VEC>(X + row * k, L, 0.0, 0.0);
        st_cols<VEC>(P + row * k, L, 0.0, 0.0);
        s0 += b0 * b0; s1 += b1 * b1;
    }
    block_partials<VEC>(s0, s1, L, k, part_bb, red);
}

// X[perm[q]] = Xq[q]  (back to the caller's numbering)
template <int VEC>
__global__ void __launch_bounds__(RB_THREADS) cg_scatter_perm_kernel(int64_t n, int k, const int* __restrict__ perm,
                                                                    const double* __restrict__ Xq, double* __restrict__ X,
                                                                    int64_t ldx, int64_t rows_per_block) {
    const int warp = threadIdx.x >> 5;
    LaneCols<VEC> L(k);
    int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block < n ? r0 + rows_per_block : n;
    for (int64_t row = r0 + warp; row < r1; row += RB_WARPS) {
        double x0, x1;
        ld_cols<VEC>(Xq + row * k, L, x0, x1);
        st_cols<VEC>(X + (int64_t)perm[row] * ldx, L, x0, x1);
    }
}

// tile path: x += alpha p (the deferred update of the previous iteration), then p = z + beta p  -- one pass over p
template <int VEC>
__global__ void __launch_bounds__(RB_THREADS) cg_pxupdate_kernel(int64_t n, int k, const double* __restrict__ alpha,
                                                                const double* __restrict__ beta, const double* __restrict__ Z,
                                                                double* __restrict__ P, double* __restrict__ X,
                                                                int64_t rows_per_block) {
    const int warp = threadIdx.x >> 5;
    LaneCols<VEC> L(k);
    double al0 = L.ok0 ? alpha[L.c] : 0.0, al1 = L.ok1 ? alpha[L.c + 1] : 0.0;
    double be0 = L.ok0 ? beta[L.c] : 0.0, be1 = L.ok1 ? beta[L.c + 1] : 0.0;
    int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block < n ? r0 + rows_per_block : n;
    for (int64_t row = r0 + warp; row < r1; row += RB_WARPS) {
        double z0, z1, p0, p1, x0, x1;
        ld_cols<VEC>(Z + row * k, L, z0, z1);
        ld_cols<VEC>(P + row * k, L, p0, p1);
        ld_cols<VEC>(X + row * k, L, x0, x1);
        st_cols<VEC>(X + row * k, L, fma(al0, p0, x0), fma(al1, p1, x1));
        st_cols<VEC>(P + row * k, L, fma(be0, p0, z0), fma(be1, p1, z1));
    }
}

// tile path: r -= alpha q ; partial rr = sum r^2
template <int VEC>
__global__ void __launch_bounds__(RB_THREADS) cg_rupdate_kernel(int64_t n, int k, const double* __restrict__ alpha,
                                                               const double* __restrict__ Q, double* __restrict__ R,
                                                               double* __restrict__ part_rr, int64_t rows_per_block) {
    __shared__ double red[RB_WARPS][64];
    const int warp = threadIdx.x >> 5;
    LaneCols<VEC> L(k);
    double al0 = L.ok0 ? alpha[L.c] : 0.0, al1 = L.ok1 ? alpha[L.c + 1] : 0.0;
    int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block < n ? r0 + rows_per_block : n;
    double s0 = 0, s1 = 0;
    for (int64_t row = r0 + warp; row < r1; row += RB_WARPS) {
        double q0, q1, a0, a1;
        ld_cols<VEC>(Q + row * k, L, q0, q1);
        ld_cols<VEC>(R + row * k, L, a0, a1);
        a0 -= al0 * q0; a1 -= al1 * q1;
        st_cols<VEC>(R + row * k, L, a0, a1);
        s0 += a0 * a0; s1 += a1 * a1;
    }
    block_partials<VEC>(s0, s1, L, k, part_rr, red);
}

// x += alpha p (final deferred update)
template <int VEC>
__global__ void __launch_bounds__(RB_THREADS) cg_xupdate_kernel(int64_t n, int k, const double* __restrict__ alpha,
                                                               const double* __restrict__ P, double* __restrict__ X,
                                                               int64_t rows_per_block) {
    const int warp = threadIdx.x >> 5;
    LaneCols<VEC> L(k);
    double al0 = L.ok0 ? alpha[L.c] : 0.0, al1 = L.ok1 ? alpha[L.c + 1] : 0.0;
    int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block < n ? r0 + rows_per_block : n;
    for (int64_t row = r0 + warp; row < r1; row += RB_WARPS) {
        double p0, p1, x0, x1;
        ld_cols<VEC>(P + row * k, L, p0, p1);
        ld_cols<VEC>(X + row * k, L, x0, x1);
        st_cols<VEC>(X + row * k, L, fma(al0, p0, x0), fma(al1, p1, x1));
    }
}

// out[c] = sum_b part[b][c], b ascending within 8 interleaved lanes, then the 8 lane sums in order (deterministic)
__global__ void __launch_bounds__(256) reduce_partials_kernel(int nparts, int k, const double* __restrict__ part, double* __restrict__ out) {
    __shared__ double sh[8][33];
    int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    int col = blockIdx.x * 32 + tx;
    double s = 0.0;
    if (col < k)
        for (int b = ty; b < nparts; b += 8) s += part[(int64_t)b * k + col];
    sh[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && col < k) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += sh[q][tx];
        out[col] = t;
    }
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

size_t mg_tiles_workspace_bytes(sfem_ctx* c, int k);
const int* mg_tiles_fine_perm(sfem_ctx* c);
int mg_apply_tiles(sfem_ctx* c, int k, const double* R, double* Z, double* T0, double* T1, void* work, double* dot_partials, int* nparts);

size_t cg_workspace_bytes(sfem_ctx* c, int k, int precond) {
    int64_t n = c->n;
    int kk = k + (k & 1);
    RowGrid g = row_grid(n);
    size_t vec = align256(sizeof(double) * (size_t)n * kk);
    size_t bytes = (precond == 1 ? 6 : 4) * vec;                  // r, p, q, z (+ second p, x in the tile numbering)
    bytes += 3 * align256(sizeof(double) * (size_t)g.nblk * kk);  // partials
    bytes += align256(sizeof(double) * 8 * (size_t)kk) + align256(sizeof(int) * ((size_t)kk + 8));
    if (precond == 1) {
        bytes += mg_tiles_workspace_bytes(c, kk);
        bytes += align256(sizeof(double) * (size_t)TILE_MAX_PARTS * kk) + align256(sizeof(double) * 2 * (size_t)kk);
    }
    return bytes + 4096;
}

// Multigrid-preconditioned path: every vector lives in the tile numbering of the fine level (mg_tiles.cu) and every
// matrix application is a staged-tile kernel.  Per iteration: the direction update, q = A p with p.q fused, the x/r
// update with r.r fused, and one V-cycle whose last smoothing sweep also returns r.z.
static int cg_solve_block_tiles(sfem_ctx* c, int k, const double* B, int64_t ldb, double* X, int64_t ldx, double rtol,
                                double atol, int max_it, int* iters_out, double* relres_out_host, void* work) {
    const int64_t n = c->n;
    RowGrid g = row_grid(n);
    const int kk = k + (k & 1);
    unsigned char* w = (unsigned char*)work;
    size_t vec = align256(sizeof(double) * (size_t)n * kk);
    double* R = (double*)w; w += vec;
    double* Pa = (double*)w; w += vec;
    double* Pb = (double*)w; w += vec;
    double* Q = (double*)w; w += vec;
    double* Z = (double*)w; w += vec;
    double* Xq = (double*)w; w += vec;
    size_t psz = align256(sizeof(double) * (size_t)g.nblk * kk);
    double* part0 = (double*)w; w += psz;
    double* part1 = (double*)w; w += psz;
    w += psz;
    Scal s;
    double* sc = (double*)w; w += align256(sizeof(double) * 8 * (size_t)kk);
    s.rho = sc; s.alpha = sc + k; s.beta = sc + 2 * k; s.bb = sc + 3 * k; s.rr = sc + 4 * k; s.tol2 = sc + 5 * k;
    double* relres_dev = sc + 6 * k;
    s.active = (int*)w; s.n_active = s.active + k;
    w += align256(sizeof(int) * ((size_t)kk + 8));
    int nch = 0;
    double* part_tile = (double*)w; w += align256(sizeof(double) * (size_t)TILE_MAX_PARTS * kk);
    double* red = (double*)w; w += align256(sizeof(double) * 2 * (size_t)kk);
    void* mgwork = w;
    const int* perm = mg_tiles_fine_perm(c);
    MGLevel* L0 = c->mg[0];

    bool v2 = (k % 2 == 0) && (ldx % 2 == 0) && (ldb % 2 == 0) &&
              ((reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(work)) % 16 == 0);
    dim3 grid(g.nblk, v2 ? (k + 63) / 64 : (k + 31) / 32);
    int sb = (k + 127) / 128;
    int rb = (k + 31) / 32;
    int* h_nact = (int*)c->pinned;
#define LAUNCH_T(passes, kern, ...)                                                    \
    do {                                                                               \
        ProfScope _ps(c, TAG_CG_VEC, (passes) * 8.0 * (double)n * k, 0.0);             \
        if (v2) kern<2> <<<grid, RB_THREADS, 0, c->stream>>>(__VA_ARGS__);              \
        else kern<1> <<<grid, RB_THREADS, 0, c->stream>>>(__VA_ARGS__);                 \
        KCHECK(c);                                                                     \
    } while (0)
    CUDA_TRY(c, cudaMemsetAsync(s.n_active, 0, 2 * sizeof(int), c->stream));
    CUDA_TRY(c, cudaMemsetAsync(sc, 0, sizeof(double) * 8 * (size_t)kk, c->stream));
    LAUNCH_T(4, cg_init_perm_kernel, n, k, perm, B, ldb, Xq, R, Pa, part0, g.rows_per_block);
    cg_scal_init_kernel<<<sb, 128, 0, c->stream>>>(k, g.nblk, part0, nullptr, s, rtol, atol);
    KCHECK(c);
    int rc = mg_apply_tiles(c, k, R, Z, Q, Pb, mgwork, part_tile, &nch);
    if (rc) return rc;
    reduce_partials_kernel<<<rb, 256, 0, c->stream>>>(nch, k, part_tile, red);
    KCHECK(c);
    cg_rho_kernel<<<sb, 128, 0, c->stream>>>(k, 1, red, s);
    KCHECK(c);
    CUDA_TRY(c, cudaMemcpyAsync(h_nact, s.n_active, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    int it = 0;
    int nact = *h_nact;
    while (nact > 0 && it < max_it && h_nact[1] == 0) {
        // x += alpha_prev p_prev (alpha = beta = 0 and p = 0 on the first pass), p = z + beta p
        LAUNCH_T(5, cg_pxupdate_kernel, n, k, s.alpha, s.beta, Z, Pa, Xq, g.rows_per_block);
        {   // q = A p ; sigma = p.q
            TileArgs a; a.k = k; a.X = Pa; a.ldx = k; a.Y = Q; a.ldy = k; a.partials = part_tile;
            rc = tileop_apply(c, TAG_CSR_MULT, L0->tA, TM_MULT_DOT, a,
                              16.0 * (double)n * k + 10.0 * (double)L0->tA.nnz + 4.0 * (double)n, &nch);
            if (rc) return rc;
        }
        reduce_partials_kernel<<<rb, 256, 0, c->stream>>>(nch, k, part_tile, red);
        KCHECK(c);
        cg_alpha_kernel<<<sb, 128, 0, c->stream>>>(k, 1, red, s);
        KCHECK(c);
        LAUNCH_T(3, cg_rupdate_kernel, n, k, s.alpha, Q, R, part1, g.rows_per_block);     // x is updated with the next direction
        rc = mg_apply_tiles(c, k, R, Z, Q, Pb, mgwork, part_tile, &nch);      // Q is free again after the update
        if (rc) return rc;
        reduce_partials_kernel<<<rb, 256, 0, c->stream>>>(nch, k, part_tile, red + kk);
        KCHECK(c);
        reduce_partials_kernel<<<rb, 256, 0, c->stream>>>(g.nblk, k, part1, red);
        KCHECK(c);
        CUDA_TRY(c, cudaMemsetAsync(s.n_active, 0, sizeof(int), c->stream));
        cg_beta_kernel<<<sb, 128, 0, c->stream>>>(k, 1, red, red + kk, s);
        KCHECK(c);
        CUDA_TRY(c, cudaMemcpyAsync(h_nact, s.n_active, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        nact = *h_nact;
        ++it;
    }
    if (iters_out) *iters_out = it;
    if (it > 0) LAUNCH_T(3, cg_xupdate_kernel, n, k, s.alpha, Pa, Xq, g.rows_per_block);     // the last deferred update
    LAUNCH_T(2, cg_scatter_perm_kernel, n, k, perm, Xq, X, ldx, g.rows_per_block);
    {   // true residual || B - A X || / || B || per column, in the caller's numbering
        rc = csr_residual(c, n, c->A_ptr, c->A_idx, c->A_val, k, X, ldx, R, k, B, ldb);
        if (rc) return rc;
        LAUNCH_T(1, cg_norm_kernel, n, k, R, part0, g.rows_per_block);
        cg_relres_kernel<<<sb, 128, 0, c->stream>>>(k, g.nblk, part0, s.bb, s.tol2, s.n_active, relres_dev);
        KCHECK(c);
        if (relres_out_host)
            CUDA_TRY(c, cudaMemcpyAsync(relres_out_host, relres_dev, sizeof(double) * k, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(h_nact, s.n_active, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
#undef LAUNCH_T
    if (h_nact[1] & 1) return sfem_fail(c, SFEM_ERR_NOT_CONVERGED, "solve_block: non-finite right-hand side or residual norm after %d iterations", it);
    if (nact > 0) return sfem_fail(c, SFEM_ERR_NOT_CONVERGED, "solve_block: %d of %d columns not converged after %d iterations", nact, k, it);
    if (h_nact[1] & 2) return sfem_fail(c, SFEM_ERR_NOT_CONVERGED, "solve_block: true residual ||B - A X|| misses the tolerance by more than %gx after %d iterations (recurrence drift or indefinite operator)", RELRES_SLACK, it);
    return SFEM_OK;
}

// Solve A X = B for k columns.  X (ldx) receives the solution.  `work` must hold cg_workspace_bytes().
int cg_solve_block(sfem_ctx* c, int k, const double* B, int64_t ldb, double* X, int64_t ldx, int precond, double rtol,
                   double atol, int max_it, int* iters_out, double* relres_out_host, void* work, size_t work_bytes) {
    if (!c->A_ptr) return sfem_fail(c, SFEM_ERR_STATE, "solve_block: stiffness matrix not set");
    if (precond == 1 && c->mg.empty()) return sfem_fail(c, SFEM_ERR_STATE, "solve_block: multigrid not set up");
    if (k <= 0) return SFEM_OK;
    int64_t n = c->n;
    if (work_bytes < cg_workspace_bytes(c, k, precond))
        return sfem_fail(c, SFEM_ERR_ARG, "solve_block: workspace too small (%zu < %zu)", work_bytes,
                         cg_workspace_bytes(c, k, precond));
    if (precond == 1) return cg_solve_block_tiles(c, k, B, ldb, X, ldx, rtol, atol, max_it, iters_out, relres_out_host, work);
    // internal blocks use leading dimension k (must be even for the double2 path: pad by requiring k even or VEC=1)
    RowGrid g = row_grid(n);
    unsigned char* w = (unsigned char*)work;
    size_t vec = align256(sizeof(double) * (size_t)n * (k + (k & 1)));
    double* R = (double*)w; w += vec;
    double* P = (double*)w; w += vec;
    double* Q = (double*)w; w += vec;
    double* Z = (double*)w; w += vec;
    size_t psz = align256(sizeof(double) * (size_t)g.nblk * (k + (k & 1)));
    double* part0 = (double*)w; w += psz;
    double* part1 = (double*)w; w += psz;
    double* part2 = (double*)w; w += psz;
    Scal s;
    double* sc = (double*)w; w += align256(sizeof(double) * 8 * (size_t)(k + (k & 1)));
    s.rho = sc; s.alpha = sc + k; s.beta = sc + 2 * k; s.bb = sc + 3 * k; s.rr = sc + 4 * k; s.tol2 = sc + 5 * k;
    double* relres_dev = sc + 6 * k;
    s.active = (int*)w; s.n_active = s.active + k;
    w += align256(sizeof(int) * ((size_t)k + 8));
    bool v2 = can_vec2(k, k, ldx, R, X) && can_vec2(k, ldb, ldb, B, B);
    dim3 grid(g.nblk, v2 ? (k + 63) / 64 : (k + 31) / 32);
    int sb = (k + 127) / 128;
    const bool jac = (precond == 0);
    int* h_nact = (int*)c->pinned;

#define LAUNCH_V(passes, kern, ...)                                                    \
    do {                                                                               \
        ProfScope _ps(c, TAG_CG_VEC, (passes) * 8.0 * (double)n * k, 0.0);             \
        if (v2) kern<2> <<<grid, RB_THREADS, 0, c->stream>>>(__VA_ARGS__);              \
        else kern<1> <<<grid, RB_THREADS, 0, c->stream>>>(__VA_ARGS__);                 \
        KCHECK(c);                                                                     \
    } while (0)
#define LAUNCH_VJ(passes, kern, ...)                                                   \
    do {                                                                               \
        ProfScope _ps(c, TAG_CG_VEC, (passes) * 8.0 * (double)n * k, 0.0);             \
        if (v2) { if (jac) kern<2, true> <<<grid, RB_THREADS, 0, c->stream>>>(__VA_ARGS__);   \
                  else kern<2, false> <<<grid, RB_THREADS, 0, c->stream>>>(__VA_ARGS__); }    \
        else { if (jac) kern<1, true> <<<grid, RB_THREADS, 0, c->stream>>>(__VA_ARGS__);      \
               else kern<1, false> <<<grid, RB_THREADS, 0, c->stream>>>(__VA_ARGS__); }       \
        KCHECK(c);                                                                     \
    } while (0)

    CUDA_TRY(c, cudaMemsetAsync(s.n_active, 0, 2 * sizeof(int), c->stream));
    LAUNCH_VJ(jac ? 4 : 3, cg_init_kernel, n, k, B, ldb, X, ldx, R, P, c->A_dinv, part0, part1, g.rows_per_block);
    cg_scal_init_kernel<<<sb, 128, 0, c->stream>>>(k, g.nblk, part0, jac ? part1 : nullptr, s, rtol, atol);
    KCHECK(c);
    CUDA_TRY(c, cudaMemcpyAsync(h_nact, s.n_active, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    int it = 0;
    int nact = *h_nact;
    while (nact > 0 && it < max_it && h_nact[1] == 0) {
        int rc = spmm_csr(c, TAG_CSR_MULT, c->A_nnz, n, c->A_ptr, c->A_idx, c->A_val, k, P, k, Q, k, part0, nullptr);
        if (rc) return rc;
        cg_alpha_kernel<<<sb, 128, 0, c->stream>>>(k, g.nblk, part0, s);
        KCHECK(c);
        LAUNCH_VJ(6, cg_update_kernel, n, k, s.alpha, P, Q, X, ldx, R, c->A_dinv, part1, part2, g.rows_per_block);
        CUDA_TRY(c, cudaMemsetAsync(s.n_active, 0, sizeof(int), c->stream));
        cg_beta_kernel<<<sb, 128, 0, c->stream>>>(k, g.nblk, part1, part2, s);
        KCHECK(c);
        CUDA_TRY(c, cudaMemcpyAsync(h_nact, s.n_active, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        LAUNCH_VJ(3, cg_pupdate_kernel, n, k, s.beta, jac ? R : Z, c->A_dinv, P, g.rows_per_block);
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        nact = *h_nact;
        ++it;
    }
    if (iters_out) *iters_out = it;
    // true residual  || B - A X || / || B ||  per column
    {
        int rc = csr_residual(c, n, c->A_ptr, c->A_idx, c->A_val, k, X, ldx, R, k, B, ldb);
        if (rc) return rc;
        LAUNCH_V(1, cg_norm_kernel, n, k, R, part0, g.rows_per_block);
        cg_relres_kernel<<<sb, 128, 0, c->stream>>>(k, g.nblk, part0, s.bb, s.tol2, s.n_active, relres_dev);
        KCHECK(c);
        if (relres_out_host) {
            CUDA_TRY(c, cudaMemcpyAsync(relres_out_host, relres_dev, sizeof(double) * k, cudaMemcpyDeviceToHost, c->stream));
        }
        CUDA_TRY(c, cudaMemcpyAsync(h_nact, s.n_active, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
#undef LAUNCH_V
#undef LAUNCH_VJ
    if (h_nact[1] & 1) return sfem_fail(c, SFEM_ERR_NOT_CONVERGED, "solve_block: non-finite right-hand side or residual norm after %d iterations", it);
    if (nact > 0) return sfem_fail(c, SFEM_ERR_NOT_CONVERGED, "solve_block: %d of %d columns not converged after %d iterations", nact, k, it);
    if (h_nact[1] & 2) return sfem_fail(c, SFEM_ERR_NOT_CONVERGED, "solve_block: true residual ||B - A X|| misses the tolerance by more than %gx after %d iterations (recurrence drift or indefinite operator)", RELRES_SLACK, it);
    return SFEM_OK;
}
