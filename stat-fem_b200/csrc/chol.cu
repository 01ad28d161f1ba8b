// K10-K12: blocked Cholesky (lower, row-major, in place), triangular solves, inverse and log-determinant, FP64.
//
// Right-looking with 128-wide panels: the 128x128 diagonal block is factorised by one CTA in shared memory, which
// also forms the explicit inverse of the triangular block; the panel solve and every trailing update are then plain
// DMMA GEMMs (dgemm.cu), as are the block forward/backward substitutions of potrs and the two triangular products
// of potri.  Replaces SciPy/LAPACK cho_factor / cho_solve at LinearSolver.py:273-297, 353-367, 603-607, 662-677.
#include "common.cuh"
#include <cstdlib>

namespace {

constexpr int NB = 128;

// Factor the nb x nb (nb <= 128) lower block at A (row-major, lda) in place and write inv(L) to Dinv (ld = 128,
// explicit zeros above the diagonal and outside nb).  info: 0, or 1-based global index of the first bad pivot.
//
// One CTA of 8 warps; the block lives in shared memory with row stride 132 (= 4 mod 16 doubles, so the
// (row = lane/4, k = lane%4) operand fragments of mma.m8n8k4.f64 hit 16 distinct 8-byte banks per half-warp).
// Phase 1, four 32-wide block steps:
//   (i)   warp 0 factorises the 32x32 diagonal block in REGISTERS (lane = row, column values exchanged by shuffles;
//         one rsqrt per column is the only long-latency operation on the chain),
//   (ii)  warp 0 inverts that triangular block (lane = column, forward substitution with broadcast reads) while
//         warps 1-7 solve the rows below it (thread = row, same substitution),
//   (iii) all warps apply the symmetric rank-32 update to the trailing lower triangle with DMMA 8x8x4 tiles.
// Phase 2: inv(L) in place, by 32x32 blocks:  M_kj = L_kj inv(D_j)  (one sweep), then block columns right to left,
// rows bottom to top:  X_ij = - sum_{k=j+1..i} X_ik M_kj  (DMMA), which only reads not-yet-overwritten blocks.
// diagnostic (build with -DSFEM_PANEL_CLOCKS, read with tools/panel_clocks.py): cycle stamps of the phases of the last panel
__device__ long long g_panel_clk[16];
#ifdef SFEM_PANEL_CLOCKS
#define PCLK(i) do { if (threadIdx.x == 0) g_panel_clk[i] = clock64(); } while (0)
#else
#define PCLK(i) do { } while (0)
#endif

constexpr int PLD = 132;       // row stride of the panel in shared memory
constexpr int DLD = 36;        // row stride of the 32x32 inverse diagonal blocks
constexpr int SB = 32;         // sub-block width

__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 1) potrf_panel_kernel(double* __restrict__ A, int64_t lda, int nb, int j0,
                                                             double* __restrict__ Dinv, int* __restrict__ info) {
    extern __shared__ __align__(16) double s[];    // [NB][PLD] | Dv[4][32][DLD] | rd[NB]
    double* Dv = s + NB * PLD;
    double* rd = Dv + 4 * SB * DLD;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lr = lane >> 2, lk = lane & 3;
    PCLK(0);
    // load the lower triangle as 16-byte pairs, 8 independent loads in flight per thread (each row is a coalesced segment);
    // the upper triangle is zero, rows / columns beyond nb are identity
    const bool vec2 = (lda % 2 == 0) && (reinterpret_cast<uintptr_t>(A) % 16 == 0);
    const bool dvec = reinterpret_cast<uintptr_t>(Dinv) % 16 == 0;
    for (int t0 = tid; t0 < NB * NB / 2; t0 += 256 * 8) {
        double2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int t = t0 + u * 256;
            const int r = t >> 6, c2 = (t & 63) * 2;
            v[u] = make_double2((r == c2) ? 1.0 : 0.0, (r == c2 + 1) ? 1.0 : 0.0);
            if (r < nb && c2 <= r) {
                const double* g = A + (int64_t)r * lda + c2;
                if (vec2) {
                    const double2 w = *reinterpret_cast<const double2*>(g);
                    v[u].x = w.x;
                    if (c2 + 1 <= r) v[u].y = w.y;
                } else {
                    v[u].x = g[0];
                    if (c2 + 1 <= r) v[u].y = g[1];
                }
                if (c2 + 1 > r) v[u].y = 0.0;
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int t = t0 + u * 256;
            *reinterpret_cast<double2*>(s + (t >> 6) * PLD + (t & 63) * 2) = v[u];
        }
    }
    __syncthreads();

    // Work that does not depend on the diagonal block being factorised (warp 0 alone, ~3 us per block step) is given to
    // the seven other warps during that window: block column pj is final after its block step, so it is written back to
    // global memory and then overwritten in place by M = L inv(D_pj), the first half of the inverse (phase 2, step A).
    auto store_L_cols = [&](int pj, int first, int nthr) {      // rows >= SB*pj of block column pj (its lower part)
        const int rows = NB - SB * pj;
        for (int t = first; t < rows * (SB / 2); t += nthr) {
            const int r = SB * pj + (t >> 4), c2 = SB * pj + (t & 15) * 2;
            if (r < nb && c2 <= r) {
                const double2 w = *reinterpret_cast<const double2*>(s + r * PLD + c2);
                double* g = A + (int64_t)r * lda + c2;
                if (vec2 && c2 + 1 <= r) *reinterpret_cast<double2*>(g) = w;
                else {
                    g[0] = w.x;
                    if (c2 + 1 <= r) g[1] = w.y;
                }
            }
        }
    };
    auto step_A_cols = [&](int pj, int wfirst, int nw) {        // M_k,pj = L_k,pj inv(D_pj) for the block rows k > pj
        int strip = 0;
        for (int kb = pj + 1; kb < NB / SB; ++kb)
            for (int ts = 0; ts < 4; ++ts, ++strip) {
                if (strip % nw != wfirst) continue;
                double* rowp = s + (SB * kb + 8 * ts + lr) * PLD + SB * pj;
                double af[8];
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) af[ks] = rowp[4 * ks + lk];
                const double* dv = Dv + pj * SB * DLD;
                double acc[4][2];
#pragma unroll
                for (int tn = 0; tn < 4; ++tn) {
                    acc[tn][0] = acc[tn][1] = 0.0;
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) dmma_8x8x4(acc[tn][0], acc[tn][1], af[ks], dv[(4 * ks + lk) * DLD + 8 * tn + lr]);
                }
                __syncwarp();
#pragma unroll
                for (int tn = 0; tn < 4; ++tn)
                    *reinterpret_cast<double2*>(rowp + 8 * tn + 2 * lk) = make_double2(acc[tn][0], acc[tn][1]);
            }
    };

    // ------------------------------------------------------------------ phase 1: Cholesky by 32-wide block steps
    PCLK(1);
    for (int jb = 0; jb < NB / SB; ++jb) {
        const int c0 = jb * SB, c1 = c0 + SB;
        if (jb == 0) PCLK(2);
        double rinvs[SB];
        if (warp == 0) {
            // 32x32 diagonal block in four 8-column sub-steps.  Within a sub-step the columns are eliminated in registers
            // (lane = row; only the 7 + 6 + ... + 1 = 28 cross-lane exchanges inside the sub-block are shuffles); the rest
            // of the block is then updated through shared memory with DMMA 8x8x4 tiles (at most 6 tiles, 2 MMAs each).
            // The all-register version spent 496 shuffle + FMA pairs on the same update: 340 cycles per column against
            // a dependency chain of ~135.
            double* blk = s + c0 * PLD + c0;
#pragma unroll
            for (int sb8 = 0; sb8 < SB / 8; ++sb8) {
                double a8[8];
                double* rowp = blk + lane * PLD + 8 * sb8;
                {
                    const double2 v0 = *reinterpret_cast<const double2*>(rowp), v1 = *reinterpret_cast<const double2*>(rowp + 2);
                    const double2 v2 = *reinterpret_cast<const double2*>(rowp + 4), v3 = *reinterpret_cast<const double2*>(rowp + 6);
                    a8[0] = v0.x; a8[1] = v0.y; a8[2] = v1.x; a8[3] = v1.y; a8[4] = v2.x; a8[5] = v2.y; a8[6] = v3.x; a8[7] = v3.y;
                }
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int j = 8 * sb8 + t;
                    double d = __shfl_sync(0xffffffffu, a8[t], j);
                    if (!(d > 0.0)) {
                        if (lane == 0 && *info == 0) *info = j0 + c0 + j + 1;
                        d = 1.0;
                    }
                    // hardware seed (relative error 2^-22) + two Newton steps: ~1 ulp, no range checks on the critical chain
                    double rinv;
                    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rinv) : "d"(d));
                    {
                        const double hd = 0.5 * d;
                        rinv = fma(rinv, fma(-hd, rinv * rinv, 0.5), rinv);
                        rinv = fma(rinv, fma(-hd, rinv * rinv, 0.5), rinv);
                    }
                    rinvs[j] = rinv;
                    double lij = a8[t] * rinv;
                    if (lane == j) { lij = d * rinv; rd[c0 + j] = rinv; }
                    if (lane < j) lij = 0.0;                    // above the diagonal
                    a8[t] = lij;
#pragma unroll
                    for (int u = t + 1; u < 8; ++u) {
                        double lkj = __shfl_sync(0xffffffffu, lij, 8 * sb8 + u);
                        a8[u] = fma(-lij, lkj, a8[u]);
                    }
                }
                *reinterpret_cast<double2*>(rowp) = make_double2(a8[0], a8[1]);
                *reinterpret_cast<double2*>(rowp + 2) = make_double2(a8[2], a8[3]);
                *reinterpret_cast<double2*>(rowp + 4) = make_double2(a8[4], a8[5]);
                *reinterpret_cast<double2*>(rowp + 6) = make_double2(a8[6], a8[7]);
                __syncwarp();
                // rows and columns 8 (sb8 + 1) .. 31 of the block: C_ij -= L_i L_j^T over the 8 columns just finished
#pragma unroll
                for (int ti = sb8 + 1; ti < SB / 8; ++ti)
#pragma unroll
                    for (int tj = sb8 + 1; tj <= ti; ++tj) {
                        const double* ap = blk + (8 * ti + lr) * PLD + 8 * sb8 + lk;
                        const double* bp = blk + (8 * tj + lr) * PLD + 8 * sb8 + lk;
                        double acc0 = 0.0, acc1 = 0.0;
                        dmma_8x8x4(acc0, acc1, ap[0], bp[0]);
                        dmma_8x8x4(acc0, acc1, ap[4], bp[4]);
                        double2* cp = reinterpret_cast<double2*>(blk + (8 * ti + lr) * PLD + 8 * tj + 2 * lk);
                        double2 cv = *cp;
                        cv.x -= acc0; cv.y -= acc1;
                        *cp = cv;
                    }
                __syncwarp();
            }
        } else if (jb > 0) {
            store_L_cols(jb - 1, tid - 32, 224);
            asm volatile("bar.sync 1, 224;\n" ::: "memory");       // the seven helper warps: all rows stored before any is overwritten
            step_A_cols(jb - 1, warp - 1, 7);
        }
        __syncthreads();
        if (jb == 0) PCLK(3);
        if (warp == 0) {
            // (ii) inverse of the diagonal block: lane = column
            double x[SB];
            double* dv = Dv + jb * SB * DLD;
#pragma unroll
            for (int i = 0; i < SB; ++i) {
                double acc0 = (i == lane) ? 1.0 : 0.0, acc1 = 0.0;
                const double* lrow = s + (c0 + i) * PLD + c0;
#pragma unroll
                for (int k = 0; k < i; ++k) {
                    if (k & 1) acc1 = fma(-lrow[k], x[k], acc1);
                    else acc0 = fma(-lrow[k], x[k], acc0);
                }
                x[i] = (acc0 + acc1) * rinvs[i];
                dv[i * DLD + lane] = x[i];
            }
        } else {
            // rows below the diagonal block: solve  x D^T = a  (thread = row)
            int r = c1 + (tid - 32);
            if (r < NB) {
                double x[SB];
                double* rowp = s + r * PLD + c0;
#pragma unroll
                for (int j = 0; j < SB; ++j) {
                    double acc0 = rowp[j], acc1 = 0.0;
                    const double* lrow = s + (c0 + j) * PLD + c0;
#pragma unroll
                    for (int k = 0; k < j; ++k) {
                        if (k & 1) acc1 = fma(-lrow[k], x[k], acc1);
                        else acc0 = fma(-lrow[k], x[k], acc0);
                    }
                    x[j] = (acc0 + acc1) * rd[c0 + j];
                }
#pragma unroll
                for (int j = 0; j < SB; ++j) rowp[j] = x[j];
            }
        }
        __syncthreads();
        if (jb == 0) PCLK(4);
        // (iii) trailing update of the lower triangle: a warp takes one 8-row strip and up to four 8x8 column tiles of it
        // at a time (four independent accumulators: the A fragment is loaded once, the MMAs of a k-step overlap)
        const int mt = (NB - c1) / 8;
        int cnt = 0;
        for (int ti = 0; ti < mt; ++ti) {
            for (int tj0 = 0; tj0 <= ti; tj0 += 4, ++cnt) {
                if ((cnt & 7) != warp) continue;
                const int ntj = min(4, ti + 1 - tj0);
                const double* ap = s + (c1 + 8 * ti + lr) * PLD + c0 + lk;
                const double* bp = s + (c1 + 8 * tj0 + lr) * PLD + c0 + lk;
                double acc[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
                for (int ks = 0; ks < SB / 4; ++ks) {
                    const double av = ap[4 * ks];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (q < ntj) dmma_8x8x4(acc[q][0], acc[q][1], av, bp[q * 8 * PLD + 4 * ks]);
                }
                double2* cp = reinterpret_cast<double2*>(s + (c1 + 8 * ti + lr) * PLD + c1 + 8 * tj0 + 2 * lk);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (q < ntj) {
                        double2 cv = cp[q * 4];
                        cv.x -= acc[q][0]; cv.y -= acc[q][1];
                        cp[q * 4] = cv;
                    }
            }
        }
        __syncthreads();
        if (jb == 0) PCLK(5);
    }
    PCLK(6);
    store_L_cols(NB / SB - 1, tid, 256);
    PCLK(7);
    // ------------------------------------------------------------------ phase 2: inv(L) in place
    // step A (M_kj = L_kj inv(D_j)) of the first three block columns was done by the helper warps above; the last block
    // column has no rows below its diagonal block.  Its diagonal block is read by the store above and overwritten below.
    __syncthreads();
    // diagonal blocks <- inv(D_j)
    for (int t = tid; t < 4 * SB * SB; t += 256) {
        int jb = t >> 10, r = (t >> 5) & 31, cidx = t & 31;
        s[(SB * jb + r) * PLD + SB * jb + cidx] = Dv[jb * SB * DLD + r * DLD + cidx];
    }
    __syncthreads();
    PCLK(8);
    // step B
    {
        const int tr = warp >> 1, tc0 = (warp & 1) * 2;
        for (int jb = 2; jb >= 0; --jb)
            for (int ib = 3; ib > jb; --ib) {
                double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
                for (int kb = jb + 1; kb <= ib; ++kb) {
                    const double* ap = s + (SB * ib + 8 * tr + lr) * PLD + SB * kb + lk;
                    const double* bp = s + (SB * kb + lk) * PLD + SB * jb + 8 * tc0 + lr;
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        double av = ap[4 * ks];
                        dmma_8x8x4(acc[0][0], acc[0][1], av, bp[4 * ks * PLD]);
                        dmma_8x8x4(acc[1][0], acc[1][1], av, bp[4 * ks * PLD + 8]);
                    }
                }
                __syncthreads();
                double* cp = s + (SB * ib + 8 * tr + lr) * PLD + SB * jb + 8 * tc0 + 2 * lk;
                *reinterpret_cast<double2*>(cp) = make_double2(-acc[0][0], -acc[0][1]);
                *reinterpret_cast<double2*>(cp + 8) = make_double2(-acc[1][0], -acc[1][1]);
                __syncthreads();
            }
    }
    PCLK(9);
    for (int t = tid; t < NB * NB / 2; t += 256) {          // Dinv is a 128 x 128 block of its own: always 16-byte aligned
        const int r = t >> 6, c2 = (t & 63) * 2;
        double2 w = make_double2(0.0, 0.0);
        if (r < nb && c2 <= r) {
            w = *reinterpret_cast<const double2*>(s + r * PLD + c2);
            if (c2 + 1 > r) w.y = 0.0;
        }
        if (dvec) *reinterpret_cast<double2*>(Dinv + 2 * t) = w;
        else { Dinv[2 * t] = w.x; Dinv[2 * t + 1] = w.y; }
    }
    PCLK(10);
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

__global__ void mirror_lower_kernel(int n, double* __restrict__ A, int64_t ld) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * n) return;
    int i = (int)(t / n), j = (int)(t % n);
    if (j > i) A[(int64_t)i * ld + j] = A[(int64_t)j * ld + i];
}

}  // namespace

static void* dense_tmp(sfem_ctx* c, size_t bytes) {
    if (bytes <= c->dense_tmp_bytes && c->dense_tmp) return c->dense_tmp;
    if (c->dense_tmp) {
        cudaStreamSynchronize(c->stream);
        sfem_dev_free(c->dense_tmp);
        c->dense_tmp = nullptr;
        c->dense_tmp_bytes = 0;
    }
    if (sfem_dev_malloc(&c->dense_tmp, bytes + 4096) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    c->dense_tmp_bytes = bytes + 4096;
    return c->dense_tmp;
}

// dinv: device buffer of ceil(n/128) * 128*128 doubles.  info_dev: device int (zeroed here).
// Two-level right-looking factorisation: 128-wide panels inside 512-wide column blocks.  Inside a column block the
// rank-128 updates touch only the block's own columns; the trailing matrix is updated once per column block with
// K = 512, where the DMMA GEMM runs at full pipeline efficiency (a K = 128 update is mostly pipeline fill and C
// traffic).  The panel solve writes to a temporary (so it may use small tiles and fill the machine) and the copy back
// into A is off the critical path.
static int potrf_lower_blocked(sfem_ctx* c, int n, double* A, int64_t lda, double* dinv, int* info_dev) {
    // column-block width: 512 keeps the big trailing updates at K = 512 (28 TFLOP/s against 25 at K = 256), which is what
    // matters from ~8000 observations up; below that the factorisation is a latency chain and the narrower block shortens
    // the in-block updates that sit on it (SFEM_POTRF_NB2 overrides, for A/B timing)
    static int nb2_env = -1;
    if (nb2_env < 0) { const char* e = getenv("SFEM_POTRF_NB2"); nb2_env = e ? atoi(e) : 0; }
    const int NB2 = nb2_env > 0 ? nb2_env : (n <= 6144 ? 256 : 512);
    CUDA_TRY(c, cudaMemsetAsync(info_dev, 0, sizeof(int), c->stream));
    size_t dyn = sizeof(double) * (NB * PLD + 4 * SB * DLD + NB);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(potrf_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        attr_set = true;
    }
    double* T = nullptr;
    if (n > NB) {
        T = (double*)dense_tmp(c, sizeof(double) * (size_t)n * NB);
        if (!T) return sfem_fail(c, SFEM_ERR_CUDA, "potrf: temporary allocation failed");
    }
    // Look-ahead: the trailing update of column block J is split into (a) the columns of block J+1, done on the main
    // stream right away, and (b) everything to the right of them, done on an auxiliary stream WHILE the main stream
    // factorises block J+1 (a chain of small, latency-bound kernels).  (b) of successive blocks accumulate into the
    // same entries, so they stay ordered on the auxiliary stream; the main stream waits for (b)(J-1) before (a)(J).
    const bool lookahead = n > 2 * NB2;
    if (lookahead && !c->aux_stream) {
        CUDA_TRY(c, cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
        CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
        CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_aux, cudaEventDisableTiming));
    }
    cudaStream_t main_stream = c->stream;
    bool aux_busy = false;
    int rc = SFEM_OK;
    for (int J0 = 0; J0 < n && rc == SFEM_OK; J0 += NB2) {
        const int Jend = imin(J0 + NB2, n);           // one past the last column of this column block
        for (int j0 = J0; j0 < Jend && rc == SFEM_OK; j0 += NB) {
            int nb = imin(NB, n - j0);
            double* Ajj = A + (int64_t)j0 * lda + j0;
            double* Dj = dinv + (size_t)(j0 / NB) * NB * NB;
            {
                ProfScope ps(c, TAG_POTRF_PANEL, 24.0 * nb * nb, (2.0 / 3.0) * nb * nb * nb);
                potrf_panel_kernel<<<1, 256, dyn, c->stream>>>(Ajj, lda, nb, j0, Dj, info_dev);
                KCHECK(c);
            }
            int m = n - j0 - nb;
            if (m <= 0) continue;
            double* A21 = A + (int64_t)(j0 + nb) * lda + j0;
            double* A22 = A + (int64_t)(j0 + nb) * lda + (j0 + nb);
            c->dgemm_tag = TAG_CHOL_TRSM;
            rc = dgemm(c, OP_N, OP_T, m, nb, nb, 1.0, A21, lda, Dj, NB, 0.0, T, NB, 0);   // T = A21 inv(L_jj)^T
            c->dgemm_tag = TAG_CHOL_SYRK;
            int ncols = Jend - (j0 + nb);              // columns of this column block still to be updated
            if (rc == SFEM_OK && ncols > 0)
                rc = dgemm(c, OP_N, OP_T, m, ncols, nb, -1.0, T, NB, T, NB, 1.0, A22, lda, 1);   // lower tiles
            c->dgemm_tag = TAG_DGEMM;
            if (rc) break;
            CUDA_TRY(c, cudaMemcpy2DAsync(A21, sizeof(double) * lda, T, sizeof(double) * NB, sizeof(double) * nb, m,
                                          cudaMemcpyDeviceToDevice, c->stream));
        }
        if (rc) break;
        int m2 = n - Jend;
        if (m2 <= 0) continue;
        const int kb = Jend - J0;
        const double* L2 = A + (int64_t)Jend * lda + J0;     // finished rows below the column block, its columns
        double* C22 = A + (int64_t)Jend * lda + Jend;
        const int na = imin(NB2, m2);                        // (a): the next column block
        c->dgemm_tag = TAG_CHOL_SYRK;
        if (lookahead) {
            CUDA_TRY(c, cudaEventRecord(c->ev_main, main_stream));          // block J is final
            if (aux_busy) CUDA_TRY(c, cudaStreamWaitEvent(main_stream, c->ev_aux, 0));   // (b)(J-1) done before (a)(J)
        }
        rc = dgemm(c, OP_N, OP_T, m2, na, kb, -1.0, L2, lda, L2, lda, 1.0, C22, lda, 1);          // (a), lower tiles
        if (rc == SFEM_OK && m2 > na) {
            const double* L3 = L2 + (int64_t)na * lda;       // rows below the next column block
            double* C33 = C22 + (int64_t)na * lda + na;
            if (lookahead) {
                CUDA_TRY(c, cudaStreamWaitEvent(c->aux_stream, c->ev_main, 0));
                c->stream = c->aux_stream;
            }
            rc = dgemm(c, OP_N, OP_T, m2 - na, m2 - na, kb, -1.0, L3, lda, L3, lda, 1.0, C33, lda, 1);   // (b)
            if (lookahead) {
                cudaEventRecord(c->ev_aux, c->aux_stream);
                c->stream = main_stream;
                aux_busy = true;
            }
        }
        c->dgemm_tag = TAG_DGEMM;
    }
    c->stream = main_stream;
    c->dgemm_tag = TAG_DGEMM;
    if (aux_busy) CUDA_TRY(c, cudaStreamWaitEvent(main_stream, c->ev_aux, 0));
    return rc;
}

// Pipelined right-looking factorisation (default).  The chain of dependent 128x128 diagonal blocks is what bounds a
// single factorisation at n_obs = 4096 (32 panels; the trailing updates total only n^3/3 = 2.3e10 flop), so the
// critical path is kept as short as the data dependence allows and everything else runs beside it:
//   main stream:  P(j) [factor block j, Dinv_j]  ->  head(j): the ONE block row the next panel needs,
//                 Th = A[j+1, j] Dinv_j^T,  A[j+1, j+1] -= Th Th^T  ->  P(j+1) ...
//   side stream:  TRSM(j): T = A[j+1.., j] Dinv_j^T for all rows; SYRK(j): A[j+2.., j+1..] -= T T^T (lower tiles, every
//                 block row except the head's); copy T back over A[j+1.., j] once the head has consumed the raw block.
// head(j) waits for SYRK(j-1) (its inputs carry the updates of all earlier panels); TRSM(j) waits for P(j).  Per panel
// the main stream executes three small kernels while the side stream streams the O(m^2) update of the previous panel.
static int potrf_lower_pipelined(sfem_ctx* c, int n, double* A, int64_t lda, double* dinv, int* info_dev) {
    CUDA_TRY(c, cudaMemsetAsync(info_dev, 0, sizeof(int), c->stream));
    size_t dyn = sizeof(double) * (NB * PLD + 4 * SB * DLD + NB);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(potrf_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        attr_set = true;
    }
    double* T = (double*)dense_tmp(c, sizeof(double) * ((size_t)n * NB + (size_t)NB * NB));
    if (!T) return sfem_fail(c, SFEM_ERR_CUDA, "potrf: temporary allocation failed");
    double* TH = T + (size_t)n * NB;                       // head block (<= 128 x 128)
    if (!c->aux_stream) {
        CUDA_TRY(c, cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
        CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
        CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_aux, cudaEventDisableTiming));
    }
    if (!c->ev_head) CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_head, cudaEventDisableTiming));
    cudaStream_t main_stream = c->stream, side = c->aux_stream;
    // the side stream starts behind everything already queued on the main stream (the matrix itself)
    CUDA_TRY(c, cudaEventRecord(c->ev_main, main_stream));
    CUDA_TRY(c, cudaStreamWaitEvent(side, c->ev_main, 0));
    int rc = SFEM_OK;
    bool side_used = false;
    for (int j0 = 0; j0 < n && rc == SFEM_OK; j0 += NB) {
        const int nb = imin(NB, n - j0);
        double* Ajj = A + (int64_t)j0 * lda + j0;
        double* Dj = dinv + (size_t)(j0 / NB) * NB * NB;
        {
            ProfScope ps(c, TAG_POTRF_PANEL, 24.0 * nb * nb, (2.0 / 3.0) * nb * nb * nb);
            potrf_panel_kernel<<<1, 256, dyn, main_stream>>>(Ajj, lda, nb, j0, Dj, info_dev);
            KCHECK(c);
        }
        const int m = n - j0 - nb;
        if (m <= 0) break;
        CUDA_TRY(c, cudaEventRecord(c->ev_main, main_stream));              // P(j) done: Dinv_j is ready
        const int j1 = j0 + nb, nbn = imin(NB, m);
        double* A21 = A + (int64_t)j1 * lda + j0;
        double* A22 = A + (int64_t)j1 * lda + j1;
        // ---- main: head(j)
        if (side_used) CUDA_TRY(c, cudaStreamWaitEvent(main_stream, c->ev_aux, 0));     // SYRK(j-1) has updated the inputs
        c->dgemm_tag = TAG_CHOL_TRSM;
        rc = dgemm(c, OP_N, OP_T, nbn, nb, nb, 1.0, A21, lda, Dj, NB, 0.0, TH, NB, 0);
        c->dgemm_tag = TAG_CHOL_SYRK;
        if (rc == SFEM_OK) rc = dgemm(c, OP_N, OP_T, nbn, nbn, nb, -1.0, TH, NB, TH, NB, 1.0, A22, lda, 1);
        if (rc) break;
        CUDA_TRY(c, cudaEventRecord(c->ev_head, main_stream));              // raw A[j+1, j] consumed
        // ---- side: TRSM(j), SYRK(j) below the head row, copy-back
        c->stream = side;
        CUDA_TRY(c, cudaStreamWaitEvent(side, c->ev_main, 0));
        c->dgemm_tag = TAG_CHOL_TRSM;
        rc = dgemm(c, OP_N, OP_T, m, nb, nb, 1.0, A21, lda, Dj, NB, 0.0, T, NB, 0);
        c->dgemm_tag = TAG_CHOL_SYRK;
        if (rc == SFEM_OK && m > nbn)
            rc = dgemm(c, OP_N, OP_T, m - nbn, m, nb, -1.0, T + (size_t)nbn * NB, NB, T, NB, 1.0, A22 + (int64_t)nbn * lda, lda,
                       1 | (nbn << 8));
        if (rc == SFEM_OK) {
            cudaEventRecord(c->ev_aux, side);                                // SYRK(j) done
            cudaStreamWaitEvent(side, c->ev_head, 0);
            cudaMemcpy2DAsync(A21, sizeof(double) * lda, T, sizeof(double) * NB, sizeof(double) * nb, m,
                              cudaMemcpyDeviceToDevice, side);
            side_used = true;
        }
        c->stream = main_stream;
        c->dgemm_tag = TAG_DGEMM;
    }
    c->stream = main_stream;
    c->dgemm_tag = TAG_DGEMM;
    if (side_used) {
        cudaEventRecord(c->ev_aux, side);
        CUDA_TRY(c, cudaStreamWaitEvent(main_stream, c->ev_aux, 0));
    }
    return rc;
}

int potrf_lower(sfem_ctx* c, int n, double* A, int64_t lda, double* dinv, int* info_dev) {
    // Default: the two-level blocked variant.  Measured on B200 (profiles/r02_potrf_ab.md): the pipelined variant is 3 %
    // faster for ONE factorisation at n = 4096 (6.59 vs 6.82 ms per value+gradient evaluation) but slower at n = 8192
    // (31.9 vs 28.9 ms) and slower when eight fits share the GPU (5.1 vs 4.0 ms per evaluation): the panel kernel is a
    // latency chain, and the GEMMs running beside it pull the SM clock down (1965 -> ~1400 MHz under tensor load), which
    // lengthens the chain by as much as the overlap saves.  SFEM_POTRF=pipelined selects it.
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("SFEM_POTRF");
        mode = (e && e[0] == 'p') ? 1 : 0;
    }
    if (mode == 0 || n <= 2 * NB) return potrf_lower_blocked(c, n, A, lda, dinv, info_dev);
    return potrf_lower_pipelined(c, n, A, lda, dinv, info_dev);
}

// B (n x k, ldb) <- inv(L L^T) B.  ncols_tri > 0 limits the forward pass to a lower-trapezoidal RHS (potri).
int potrs_lower(sfem_ctx* c, int n, const double* L, int64_t ldl, const double* dinv, int k, double* B, int64_t ldb) {
    // forward: L Y = B
    for (int j0 = 0; j0 < n; j0 += NB) {
        int nb = imin(NB, n - j0);
        const double* Dj = dinv + (size_t)(j0 / NB) * NB * NB;
        double* Bj = B + (int64_t)j0 * ldb;
        int rc = dgemm(c, OP_N, OP_N, nb, k, nb, 1.0, Dj, NB, Bj, ldb, 0.0, Bj, ldb, 4);
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
        int rc = dgemm(c, OP_T, OP_N, nb, k, nb, 1.0, Dj, NB, Bj, ldb, 0.0, Bj, ldb, 4);
        if (rc) return rc;
        if (j0 > 0) {
            rc = dgemm(c, OP_T, OP_N, j0, k, nb, -1.0, L + (int64_t)j0 * ldl, ldl, Bj, ldb, 1.0, B, ldb, 0);
            if (rc) return rc;
        }
    }
    return SFEM_OK;
}

__global__ void copy_diag_blocks_kernel(int n, const double* __restrict__ dinv, double* __restrict__ W, int64_t ldw) {
    // block b of dinv (128 x 128, explicit zeros above the diagonal) -> W[b*128.., b*128..]
    int b = blockIdx.y;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int r = t >> 7, cidx = t & (NB - 1);
    int gr = b * NB + r, gc = b * NB + cidx;
    if (r < NB && gr < n && gc < n) W[(int64_t)gr * ldw + gc] = dinv[(size_t)b * NB * NB + t];
}

// W (n x n, ld n) = inv(L) by recursive doubling over the 128-wide diagonal blocks whose inverses potrf left in
// dinv:  for block size s = 128, 256, ...:  every pair [[L11, 0], [L21, L22]] with inv(L11), inv(L22) known gets
//   X21 = - inv(L22) (L21 inv(L11)),
// two batched GEMMs per level (all pairs of a level are independent), log2(n/128) levels instead of n/128 steps.
// Strictly upper tiles of W are never written nor read.
int trtri_lower(sfem_ctx* c, int n, const double* L, int64_t ldl, const double* dinv, double* W) {
    int nblk = (n + NB - 1) / NB;
    copy_diag_blocks_kernel<<<dim3(NB * NB / 256, nblk), 256, 0, c->stream>>>(n, dinv, W, n);
    KCHECK(c);
    if (n <= NB) return SFEM_OK;
    double* T = (double*)dense_tmp(c, sizeof(double) * (size_t)n * (size_t)((n + 1) / 2 + NB));
    if (!T) return sfem_fail(c, SFEM_ERR_CUDA, "trtri: temporary allocation failed");
    for (int64_t s = NB; s < n; s *= 2) {
        int full = (int)(n / (2 * s));                 // pairs with a complete lower block
        int64_t rem = n - (int64_t)full * 2 * s;       // ragged tail: a pair exists if rem > s
        for (int pass = 0; pass < 2; ++pass) {
            int batch = pass == 0 ? full : (rem > s ? 1 : 0);
            if (batch == 0) continue;
            int64_t o = pass == 0 ? 0 : (int64_t)full * 2 * s;
            int m2 = pass == 0 ? (int)s : (int)(rem - s);
            const double* L21 = L + (o + s) * ldl + o;
            const double* X11 = W + o * n + o;
            const double* X22 = W + (o + s) * n + (o + s);
            double* X21 = W + (o + s) * n + o;
            // T = L21 X11   (m2 x s, K = s; X11 lower triangular: k >= n0)
            int rc = dgemm_batched(c, OP_N, OP_N, m2, (int)s, s, 1.0, L21, ldl, X11, n, 0.0, T, s, 8, batch,
                                   2 * s * (ldl + 1), 2 * s * ((int64_t)n + 1), s * s);
            if (rc) return rc;
            // X21 = - X22 T (m2 x s, K = m2; X22 lower triangular: k <= m0 + tile)
            rc = dgemm_batched(c, OP_N, OP_N, m2, (int)s, m2, -1.0, X22, n, T, s, 0.0, X21, n, 16, batch,
                               2 * s * ((int64_t)n + 1), s * s, 2 * s * ((int64_t)n + 1));
            if (rc) return rc;
        }
    }
    return SFEM_OK;
}

// Kinv (n x n, ld n) = inv(L L^T); W is an n x n workspace (receives inv(L)).
int potri_lower(sfem_ctx* c, int n, const double* L, int64_t ldl, const double* dinv, double* W, double* Kinv) {
    int64_t tot = (int64_t)n * n;
    c->dgemm_tag = TAG_TRTRI;
    int rc = trtri_lower(c, n, L, ldl, dinv, W);
    c->dgemm_tag = TAG_LAUUM;
    if (rc == SFEM_OK) rc = dgemm(c, OP_T, OP_N, n, n, n, 1.0, W, n, W, n, 0.0, Kinv, n, 3);   // lower tiles, k >= max(m0, n0)
    c->dgemm_tag = TAG_DGEMM;
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

// diagnostic (not part of the public header): cycle stamps of the phases of the last panel kernel
extern "C" int sfem_debug_panel_clocks(long long* out16_host) {
    return cudaMemcpyFromSymbol(out16_host, g_panel_clk, sizeof(long long) * 16) == cudaSuccess ? 0 : -1;
}
