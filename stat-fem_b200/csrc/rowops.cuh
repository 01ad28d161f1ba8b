// Shared device helpers for the row-major multi-RHS block kernels: a warp owns a row, lanes own RHS columns.
#pragma once
#include "common.cuh"

constexpr int RB_WARPS = 8;
constexpr int RB_THREADS = RB_WARPS * 32;

// Lane -> column mapping of a block kernel.  VEC = 2: lane owns columns (c, c+1) as a double2.
template <int VEC>
struct LaneCols {
    int c;
    bool ok0, ok1;
    __device__ __forceinline__ LaneCols(int k) {
        int lane = threadIdx.x & 31;
        c = blockIdx.y * (32 * VEC) + lane * VEC;
        ok0 = c < k;
        ok1 = (VEC == 2) && (c + 1 < k);
    }
};

template <int VEC>
__device__ __forceinline__ void ld_cols(const double* __restrict__ p, const LaneCols<VEC>& L, double& x0, double& x1) {
    if (VEC == 2 && L.ok1) {
        double2 v = *reinterpret_cast<const double2*>(p + L.c);
        x0 = v.x; x1 = v.y;
    } else if (L.ok0) {
        x0 = p[L.c]; x1 = 0.0;
    } else {
        x0 = 0.0; x1 = 0.0;
    }
}

template <int VEC>
__device__ __forceinline__ void st_cols(double* __restrict__ p, const LaneCols<VEC>& L, double x0, double x1) {
    if (VEC == 2 && L.ok1) {
        *reinterpret_cast<double2*>(p + L.c) = make_double2(x0, x1);
    } else if (L.ok0) {
        p[L.c] = x0;
    }
}

// a += sum_p val[p] * X[idx[p]][cols]   over the CSR row [p, pe)
template <int VEC>
__device__ __forceinline__ void csr_row_accum(const int32_t* __restrict__ idx, const double* __restrict__ val,
                                              int64_t p, int64_t pe, const double* __restrict__ X, int64_t ldx,
                                              const LaneCols<VEC>& L, double& a0, double& a1) {
    for (; p + 4 <= pe; p += 4) {
        int j0 = idx[p], j1 = idx[p + 1], j2 = idx[p + 2], j3 = idx[p + 3];
        double v0 = val[p], v1 = val[p + 1], v2 = val[p + 2], v3 = val[p + 3];
        double x00, x01, x10, x11, x20, x21, x30, x31;
        ld_cols<VEC>(X + (int64_t)j0 * ldx, L, x00, x01);
        ld_cols<VEC>(X + (int64_t)j1 * ldx, L, x10, x11);
        ld_cols<VEC>(X + (int64_t)j2 * ldx, L, x20, x21);
        ld_cols<VEC>(X + (int64_t)j3 * ldx, L, x30, x31);
        a0 += v0 * x00; a1 += v0 * x01;
        a0 += v1 * x10; a1 += v1 * x11;
        a0 += v2 * x20; a1 += v2 * x21;
        a0 += v3 * x30; a1 += v3 * x31;
    }
    for (; p < pe; ++p) {
        double x0, x1;
        ld_cols<VEC>(X + (int64_t)idx[p] * ldx, L, x0, x1);
        double v = val[p];
        a0 += v * x0; a1 += v * x1;
    }
}

// Per-block, per-column partial sums: each lane holds the running sum for its columns; summed over the block's
// warps in a fixed order and written to partials[blockIdx.x][col] (row stride = k).
template <int VEC>
__device__ __forceinline__ void block_partials(double d0, double d1, const LaneCols<VEC>& L, int k,
                                               double* __restrict__ partials, double (*red)[64]) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    red[warp][lane * VEC] = d0;
    if (VEC == 2) red[warp][lane * VEC + 1] = d1;
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < RB_WARPS; ++w) s += red[w][lane * VEC + v];
            if (L.c + v < k) partials[(int64_t)blockIdx.x * k + L.c + v] = s;
        }
    }
    __syncthreads();
}

// Persistent row-chunk decomposition shared by all block kernels (so partials of different kernels line up).
struct RowGrid {
    int nblk;
    int64_t rows_per_block;
};
static inline RowGrid row_grid(int64_t n) {
    RowGrid g;
    int64_t want = 4 * SFEM_NUM_SMS;
    g.rows_per_block = cdiv64(n, want);
    if (g.rows_per_block < RB_WARPS) g.rows_per_block = RB_WARPS;
    g.nblk = (int)cdiv64(n, g.rows_per_block);
    if (g.nblk < 1) g.nblk = 1;
    return g;
}
static inline bool can_vec2(int k, int64_t ld1, int64_t ld2, const void* p1, const void* p2) {
    return (k >= 2) && (ld1 % 2 == 0) && (ld2 % 2 == 0) &&
           ((reinterpret_cast<uintptr_t>(p1) | reinterpret_cast<uintptr_t>(p2)) % 16 == 0);
}
