// K9: FP64 tensor-core GEMM for sm_100a.
//
//   C[M x N] = alpha * op(A) * op(B) + beta * C          (row-major everywhere)
//
// FP64 has no tcgen05/TMEM kind on Blackwell; the FP64 tensor path is mma.sync (DMMA.8x8x4 in SASS).  The
// kernel is a 128x128x16 CTA tile, 8 warps (2 x 4), 64x32 warp tiles held in registers (64 doubles/thread),
// a 4-stage cp.async (LDGSTS) global->shared pipeline, and bank-conflict-free padded shared layouts:
//   k-major operand  (op(A)=A^T stored K x M, op(B)=B stored K x N):  tile [16][132]
//   mn-major operand (op(A)=A   stored M x K, op(B)=B^T stored N x K): tile [128][20]
// (132 = 4 mod 16 and 20 = 4 mod 16 doubles make the 16 lanes of a half-warp hit 16 distinct 8-byte banks
//  for the (row = lane>>2, k = lane&3) fragment pattern of mma.m8n8k4.f64.)
//
// Users on the hot path: Cu = W^T Z (TN, K = n_dof: SURVEY K9), blocked Cholesky/TRSM/potri (chol.cu),
// posterior covariance products (dense_ops.cu).  Reference call sites replaced: the per-sensor
// P^T (A^-1 G A^-1 p_i) loop of solving_utils.py:139-142 and NumPy/LAPACK calls of LinearSolver.py:361-367.
#include "common.cuh"

namespace {

constexpr int BK = 16, STAGES = 4, NTHREADS = 256;
constexpr int LDMN = 20;            // mn-major tile row stride (doubles)
// TB = CTA tile edge (128: 8 warps of 64x32; 64: 8 warps of 32x16, two CTAs per SM -- used when a 128-tiling would
// leave most of the 148 SMs idle, e.g. the panel-width GEMMs of the blocked Cholesky)
template <int TB> struct TileCfg {
    static constexpr int LDK = TB + 4;                      // k-major tile row stride (doubles), = 4 mod 16
    static constexpr int TILE_DOUBLES = (BK * (TB + 4) > TB * LDMN) ? BK * (TB + 4) : TB * LDMN;
    static constexpr size_t SMEM_BYTES = (size_t)STAGES * 2 * TILE_DOUBLES * sizeof(double);
    static constexpr int MI = TB / 16, NJ = TB / 32;        // 8x8 mma tiles per warp along m, n
};

struct GemmParams {
    int M, N;
    int64_t K;
    const double* A; int64_t lda;
    const double* B; int64_t ldb;
    double* C; int64_t ldc;
    double alpha, beta;
    int lower_only;      // skip tiles strictly above the block diagonal
    int diag_off;        // ... of a C whose row 0 sits diag_off rows below its column 0 (tiles with n0 > m0 + diag_off are skipped)
    int tri_k;           // 1: k starts at max(m0, n0) (both operands lower-triangular "k >= index")
    int64_t kchunk;      // split-K chunk (multiple of BK); gridDim.z chunks
    double* partial;     // split-K workspace [gridDim.z][M][N] or nullptr
    int tiles_m, tiles_n;
    int k_from_n0;       // 1: k starts at n0   (B operand lower-triangular: B[k][n] = 0 for k < n)
    int k_to_m0;         // 1: k ends at m0+TB  (A operand lower-triangular: A[m][k] = 0 for k > m)
    int64_t bsA, bsB, bsC;   // batch strides (blockIdx.y = batch index)
};

__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_8(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// Load one operand tile for k-range [k0, k0+BK) (clipped to kend) and mn-range [mn0, mn0+128) (clipped to MN).
// KM = true : source element (k, mn) at src[k*ld + mn]   -> smem[k*LDK + mn]
// KM = false: source element (mn, k) at src[mn*ld + k]   -> smem[mn*LDMN + k]
template <int TB, bool KM, bool AL16>
__device__ __forceinline__ void load_tile(double* smem, const double* __restrict__ src, int64_t ld, int64_t k0,
                                          int64_t kend, int mn0, int MN, int tid) {
    constexpr int LDK = TileCfg<TB>::LDK;
    if (KM) {
        if (AL16) {
#pragma unroll
            for (int it = 0; it < (BK * TB / 2) / NTHREADS; ++it) {
                int q = tid + it * NTHREADS;
                int row = q / (TB / 2), c2 = (q % (TB / 2)) * 2;
                int64_t k = k0 + row;
                int mn = mn0 + c2;
                int valid = (k < kend) ? (mn + 1 < MN ? 16 : (mn < MN ? 8 : 0)) : 0;
                const double* g = valid ? (src + k * ld + mn) : src;
                cp_async_16(smem + row * LDK + c2, g, valid);
            }
        } else {
#pragma unroll
            for (int it = 0; it < (BK * TB) / NTHREADS; ++it) {
                int q = tid + it * NTHREADS;
                int row = q / TB, cc = q % TB;
                int64_t k = k0 + row;
                int mn = mn0 + cc;
                int valid = (k < kend && mn < MN) ? 8 : 0;
                const double* g = valid ? (src + k * ld + mn) : src;
                cp_async_8(smem + row * LDK + cc, g, valid);
            }
        }
    } else {
        if (AL16) {
#pragma unroll
            for (int it = 0; it < (TB * 8) / NTHREADS; ++it) {
                int q = tid + it * NTHREADS;
                int row = q >> 3, c2 = (q & 7) * 2;
                int64_t k = k0 + c2;
                int mn = mn0 + row;
                int valid = (mn < MN) ? (k + 1 < kend ? 16 : (k < kend ? 8 : 0)) : 0;
                const double* g = valid ? (src + (int64_t)mn * ld + k) : src;
                cp_async_16(smem + row * LDMN + c2, g, valid);
            }
        } else {
#pragma unroll
            for (int it = 0; it < (TB * 16) / NTHREADS; ++it) {
                int q = tid + it * NTHREADS;
                int row = q >> 4, cc = q & 15;
                int64_t k = k0 + cc;
                int mn = mn0 + row;
                int valid = (mn < MN && k < kend) ? 8 : 0;
                const double* g = valid ? (src + (int64_t)mn * ld + k) : src;
                cp_async_8(smem + row * LDMN + cc, g, valid);
            }
        }
    }
}

template <int TB, bool AKM, bool BKM, bool AL16>
__global__ void __launch_bounds__(NTHREADS, (TB == 128 ? 1 : 2)) dgemm_kernel(GemmParams p) {
    extern __shared__ __align__(16) double smem[];
    constexpr int BM = TB, BN = TB, LDK = TileCfg<TB>::LDK, TILE_DOUBLES = TileCfg<TB>::TILE_DOUBLES;
    constexpr int MI = TileCfg<TB>::MI, NJ = TileCfg<TB>::NJ;
    // tile rasterisation: strips of 8 tile-rows so concurrently resident CTAs share operand panels in L2
    int tile = blockIdx.x;
    const int GROUP = 8;
    int tiles_per_group = GROUP * p.tiles_n;
    int g = tile / tiles_per_group;
    int first_m = g * GROUP;
    int gsize = min(p.tiles_m - first_m, GROUP);
    int tm = first_m + (tile % tiles_per_group) % gsize;
    int tn = (tile % tiles_per_group) / gsize;
    int m0 = tm * BM, n0 = tn * BN;
    if (p.lower_only && n0 > m0 + p.diag_off) return;
    p.A += (int64_t)blockIdx.y * p.bsA;
    p.B += (int64_t)blockIdx.y * p.bsB;
    p.C += (int64_t)blockIdx.y * p.bsC;

    int64_t kbeg = (int64_t)blockIdx.z * p.kchunk;
    int64_t kend = kbeg + p.kchunk < p.K ? kbeg + p.kchunk : p.K;
    if (p.tri_k) {
        int64_t ks = m0 > n0 ? m0 : n0;     // multiples of TB -> aligned to BK
        if (ks > kbeg) kbeg = ks;
    }
    if (p.k_from_n0 && n0 > kbeg) kbeg = n0;
    if (p.k_to_m0 && m0 + BM < kend) kend = m0 + BM;
    int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int wm0 = (warp & 1) * (TB / 2), wn0 = (warp >> 1) * (TB / 4);
    int lr = lane >> 2, lk = lane & 3;

    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    int64_t nk = kend > kbeg ? (kend - kbeg + BK - 1) / BK : 0;

    auto issue = [&](int64_t kt) {
        if (kt < nk) {
            int slot = (int)(kt % STAGES);
            double* sA = smem + (size_t)slot * 2 * TILE_DOUBLES;
            double* sB = sA + TILE_DOUBLES;
            int64_t k0 = kbeg + kt * BK;
            load_tile<TB, AKM, AL16>(sA, p.A, p.lda, k0, kend, m0, p.M, tid);
            load_tile<TB, BKM, AL16>(sB, p.B, p.ldb, k0, kend, n0, p.N, tid);
        }
        cp_async_commit();
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issue(s);

    for (int64_t kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        issue(kt + STAGES - 1);
        const double* sA = smem + (size_t)(kt % STAGES) * 2 * TILE_DOUBLES;
        const double* sB = sA + TILE_DOUBLES;
#pragma unroll
        for (int ks = 0; ks < BK / 4; ++ks) {
            double a[MI], b[NJ];
            int kk = ks * 4 + lk;
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                int m = wm0 + i * 8 + lr;
                a[i] = AKM ? sA[kk * LDK + m] : sA[m * LDMN + kk];
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                int n = wn0 + j * 8 + lr;
                b[j] = BKM ? sB[kk * LDK + n] : sB[n * LDMN + kk];
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue
    if (p.partial) {
        double* P = p.partial + (size_t)blockIdx.z * p.M * p.N;
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            int m = m0 + wm0 + i * 8 + lr;
            if (m >= p.M) continue;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                int n = n0 + wn0 + j * 8 + lk * 2;
                if ((p.N % 2 == 0) && n + 1 < p.N) {
                    *reinterpret_cast<double2*>(P + (size_t)m * p.N + n) = make_double2(acc[i][j][0], acc[i][j][1]);
                } else {
                    if (n < p.N) P[(size_t)m * p.N + n] = acc[i][j][0];
                    if (n + 1 < p.N) P[(size_t)m * p.N + n + 1] = acc[i][j][1];
                }
            }
        }
    } else {
        // C <- alpha acc + beta C.  The old values of a whole row of the warp tile are loaded BEFORE anything is stored:
        // written as one load-store pair per element, the compiler must assume that a store may alias the next load,
        // and the 64 dependent global round trips (~23 us per 128x128 tile, measured) cost more than a 128-deep k-loop.
        const bool vec = (p.ldc % 2 == 0) && (reinterpret_cast<uintptr_t>(p.C) % 16 == 0);
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            int m = m0 + wm0 + i * 8 + lr;
            if (m >= p.M) continue;
            double* crow = p.C + (int64_t)m * p.ldc + n0 + wn0 + lk * 2;
            double old[NJ][2];
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                int n = n0 + wn0 + j * 8 + lk * 2;
                old[j][0] = old[j][1] = 0.0;
                if (p.beta != 0.0) {
                    if (vec && n + 1 < p.N) {
                        double2 v = *reinterpret_cast<const double2*>(crow + j * 8);
                        old[j][0] = v.x; old[j][1] = v.y;
                    } else {
                        if (n < p.N) old[j][0] = crow[j * 8];
                        if (n + 1 < p.N) old[j][1] = crow[j * 8 + 1];
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                int n = n0 + wn0 + j * 8 + lk * 2;
                double r0 = p.alpha * acc[i][j][0] + (p.beta != 0.0 ? p.beta * old[j][0] : 0.0);
                double r1 = p.alpha * acc[i][j][1] + (p.beta != 0.0 ? p.beta * old[j][1] : 0.0);
                if (vec && n + 1 < p.N) {
                    *reinterpret_cast<double2*>(crow + j * 8) = make_double2(r0, r1);
                } else {
                    if (n < p.N) crow[j * 8] = r0;
                    if (n + 1 < p.N) crow[j * 8 + 1] = r1;
                }
            }
        }
    }
}

__global__ void splitk_reduce_kernel(int M, int N, int splits, const double* __restrict__ partial, double alpha,
                                     double beta, double* __restrict__ C, int64_t ldc, int lower_only, int tb, int diag_off) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)M * N) return;
    int m = (int)(idx / N), n = (int)(idx % N);
    if (lower_only && (n / tb) * tb > (m / tb) * tb + diag_off) return;
    double s = 0.0;
    for (int z = 0; z < splits; ++z) s += partial[(size_t)z * M * N + idx];
    double* c = C + (int64_t)m * ldc + n;
    *c = alpha * s + (beta != 0.0 ? beta * (*c) : 0.0);
}

template <int TB, bool AKM, bool BKM>
int launch(sfem_ctx* c, const GemmParams& p, bool al16, dim3 grid) {
    constexpr size_t SMEM_BYTES = TileCfg<TB>::SMEM_BYTES;
    static bool attr_set[2] = {false, false};
    if (al16) {
        auto k = dgemm_kernel<TB, AKM, BKM, true>;
        if (!attr_set[1]) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES); attr_set[1] = true; }
        k<<<grid, NTHREADS, SMEM_BYTES, c->stream>>>(p);
    } else {
        auto k = dgemm_kernel<TB, AKM, BKM, false>;
        if (!attr_set[0]) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES); attr_set[0] = true; }
        k<<<grid, NTHREADS, SMEM_BYTES, c->stream>>>(p);
    }
    KCHECK(c);
    return SFEM_OK;
}

template <int TB>
int launch_layout(sfem_ctx* c, const GemmParams& p, bool al16, dim3 grid, bool akm, bool bkm) {
    if (akm && bkm) return launch<TB, true, true>(c, p, al16, grid);
    if (akm && !bkm) return launch<TB, true, false>(c, p, al16, grid);
    if (!akm && bkm) return launch<TB, false, true>(c, p, al16, grid);
    return launch<TB, false, false>(c, p, al16, grid);
}

}  // namespace

// tri: 0 none, 1 = k starts at max(m0,n0) (used for Linv^T Linv).  Encoded in the high bits of lower_only:
//   lower_only & 1 -> only tiles on/below the block diagonal, lower_only & 2 -> tri_k.
//   lower_only & 4 -> C aliases an operand: keep 128-tiles;  & 8 -> k starts at n0;  & 16 -> k ends at m0 + tile.
//   bits 8.. (flags >> 8): diag_off in elements for the lower-only test (C's first row is that many rows below the diagonal).
int dgemm(sfem_ctx* c, int opA, int opB, int M, int N, int64_t K, double alpha, const double* A, int64_t lda,
          const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int flags) {
    return dgemm_batched(c, opA, opB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, flags, 1, 0, 0, 0);
}

int dgemm_batched(sfem_ctx* c, int opA, int opB, int M, int N, int64_t K, double alpha, const double* A, int64_t lda,
                  const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int flags, int batch,
                  int64_t strideA, int64_t strideB, int64_t strideC) {
    if (M <= 0 || N <= 0 || batch <= 0) return SFEM_OK;
    GemmParams p;
    p.M = M; p.N = N; p.K = K;
    p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc;
    p.alpha = alpha; p.beta = beta;
    p.lower_only = flags & 1;
    p.diag_off = flags >> 8;
    p.tri_k = (flags & 2) ? 1 : 0;
    p.k_from_n0 = (flags & 8) ? 1 : 0;
    p.k_to_m0 = (flags & 16) ? 1 : 0;
    p.bsA = strideA; p.bsB = strideB; p.bsC = strideC;
    // tile edge: 128 unless that tiling cannot occupy the machine (then 64, two CTAs per SM)
    int tb = 128;
    {
        int64_t tm = (M + 127) / 128, tn = (N + 127) / 128;
        int64_t eff = p.lower_only ? tm * (tn < tm ? tn : tm) - (tn < tm ? tn : tm) * ((tn < tm ? tn : tm) - 1) / 2 : tm * tn;
        eff *= batch;
        if (eff < 2 * SFEM_NUM_SMS && !(flags & 4)) tb = 64;     // flag 4: C aliases an operand (only safe when the
                                                                  // aliased dimension is a single 128-tile)
    }
    p.tiles_m = (M + tb - 1) / tb;
    p.tiles_n = (N + tb - 1) / tb;
    p.partial = nullptr;
    int tiles = p.tiles_m * p.tiles_n;
    // split-K when the tile grid cannot fill the machine and K is long (e.g. Cu = W^T Z with few sensors)
    int splits = 1;
    if (!p.tri_k && batch == 1 && !(flags & 24) && tiles < (tb == 64 ? 2 : 1) * SFEM_NUM_SMS && K >= 4096) {
        splits = (2 * SFEM_NUM_SMS + tiles - 1) / tiles;
        int64_t maxs = K / 1024;
        if (splits > maxs) splits = (int)maxs;
        if (splits < 1) splits = 1;
    }
    // lower-only products with a very long K (Cu = (S Z)^T Z: 528 tiles of 3.4e10 flop each at 4096 sensors): the last
    // wave of tiles leaves most SMs idle (528 / 148 = 3.57 -> 4 waves); cutting K into a few chunks shortens the tail
    if (splits == 1 && p.lower_only && !p.tri_k && batch == 1 && !(flags & 24) && tb == 128 && K >= 65536) {
        int64_t tmn = p.tiles_m < p.tiles_n ? p.tiles_m : p.tiles_n;
        int64_t eff = (int64_t)p.tiles_m * tmn - tmn * (tmn - 1) / 2;
        double best = (double)((eff + SFEM_NUM_SMS - 1) / SFEM_NUM_SMS);
        for (int sct = 2; sct <= 8; ++sct) {
            double waves = (double)((eff * sct + SFEM_NUM_SMS - 1) / SFEM_NUM_SMS) / sct;
            if (waves < 0.97 * best) { best = waves; splits = sct; }
        }
    }
    int64_t kchunk = K;
    if (splits > 1) {
        kchunk = ((K + splits - 1) / splits + BK - 1) / BK * BK;
        splits = (int)((K + kchunk - 1) / kchunk);
        p.partial = (double*)sfem_scratch(c, (size_t)splits * M * N * sizeof(double));
        if (!p.partial) return sfem_fail(c, SFEM_ERR_CUDA, "dgemm: split-K workspace allocation failed");
    }
    p.kchunk = kchunk > 0 ? kchunk : BK;
    bool al16 = ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) % 16 == 0) &&
                (lda % 2 == 0) && (ldb % 2 == 0);
    ProfScope ps(c, c->dgemm_tag, batch * 8.0 * ((double)M * K + (double)K * N + (double)M * N * (beta != 0.0 ? 2.0 : 1.0)),
                 batch * 2.0 * (double)M * N * (double)K *
                     ((flags & 3) == 3 ? 1.0 / 6.0 : ((flags & 1) ? 0.5 : 1.0)) * ((flags & 24) ? 0.5 : 1.0));
    dim3 grid(tiles, batch, splits);
    bool akm = (opA == OP_T), bkm = (opB == OP_N);
    int rc = tb == 128 ? launch_layout<128>(c, p, al16, grid, akm, bkm) : launch_layout<64>(c, p, al16, grid, akm, bkm);
    if (rc) return rc;
    if (splits > 1) {
        int64_t tot = (int64_t)M * N;
        splitk_reduce_kernel<<<(unsigned)cdiv64(tot, 256), 256, 0, c->stream>>>(M, N, splits, p.partial, alpha, beta,
                                                                               C, ldc, p.lower_only, tb, p.diag_off);
        KCHECK(c);
    }
    return SFEM_OK;
}
