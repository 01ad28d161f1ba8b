// Staged-tile sparse operators: format builder and the fused apply kernels (see tileop.cuh).
#include "tileop.cuh"
#include <climits>

template <typename T>
int exclusive_scan(sfem_ctx* c, const T* in, int64_t n, T* out, T* sums);

namespace {

constexpr int TB_MAXENT = 8192;   // matrix entries per chunk the builder can sort in shared memory
constexpr int TB_THREADS = 256;

// ------------------------------------------------------------------------------------------------ builder
// Sort the column indices of one chunk's entries and compact them to the ascending list of distinct columns.
// Returns the number of distinct columns; `uniq` holds them.  All threads of the block must call.
__device__ int chunk_distinct_columns(const int32_t* __restrict__ idx, int64_t e0, int ne, int* keys, int* uniq, int* scan) {
    const int tid = threadIdx.x;
    int npow = 32;
    while (npow < ne) npow <<= 1;
    for (int t = tid; t < npow; t += TB_THREADS) keys[t] = t < ne ? idx[e0 + t] : INT_MAX;
    for (int size = 2; size <= npow; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = tid; t < (npow >> 1); t += TB_THREADS) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool asc = (lo & size) == 0;
                int a = keys[lo], b = keys[hi];
                if ((a > b) == asc) { keys[lo] = b; keys[hi] = a; }
            }
        }
    __syncthreads();
    // flags -> positions: thread t owns a contiguous run of `per` sorted keys
    int per = (npow + TB_THREADS - 1) / TB_THREADS;
    int b0 = tid * per, b1 = b0 + per < npow ? b0 + per : npow;
    int cnt = 0;
    for (int t = b0; t < b1; ++t)
        if (keys[t] != INT_MAX && (t == 0 || keys[t] != keys[t - 1])) ++cnt;
    scan[tid] = cnt;
    __syncthreads();
    for (int o = 1; o < TB_THREADS; o <<= 1) {
        int v = tid >= o ? scan[tid - o] : 0;
        __syncthreads();
        scan[tid] += v;
        __syncthreads();
    }
    int pos = scan[tid] - cnt;
    for (int t = b0; t < b1; ++t)
        if (keys[t] != INT_MAX && (t == 0 || keys[t] != keys[t - 1])) uniq[pos++] = keys[t];
    int total = scan[TB_THREADS - 1];
    __syncthreads();
    return total;
}

__device__ __forceinline__ int lower_bound_dev(const int* a, int n, int key) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// pass 0: per-chunk counts.  stats[0] = max need, stats[1] = max entries, stats[2] = error flag
__global__ void __launch_bounds__(TB_THREADS) tile_count_kernel(const int* __restrict__ chunk_ptr, const int64_t* __restrict__ ptr,
                                                                const int32_t* __restrict__ idx, int* __restrict__ need_cnt,
                                                                int* __restrict__ stats) {
    extern __shared__ int tb_smem[];
    int* keys = tb_smem;
    int* uniq = tb_smem + TB_MAXENT;
    __shared__ int scan[TB_THREADS];
    int cidx = blockIdx.x;
    int r0 = chunk_ptr[cidx], r1 = chunk_ptr[cidx + 1];
    int64_t e0 = ptr[r0], e1 = ptr[r1];
    if (e1 - e0 > TB_MAXENT) {
        if (threadIdx.x == 0) { stats[2] = 1; need_cnt[cidx] = 0; }
        return;
    }
    int ne = (int)(e1 - e0);
    int nu = chunk_distinct_columns(idx, e0, ne, keys, uniq, scan);
    if (threadIdx.x == 0) {
        need_cnt[cidx] = nu;
        atomicMax(&stats[0], nu);
        atomicMax(&stats[1], ne);
        if (nu > 65535) stats[2] = 1;
    }
}

// pass 1: staged-row lists, slot indices, row pointers, self slots
__global__ void __launch_bounds__(TB_THREADS) tile_fill_kernel(int64_t nrows, int square, const int* __restrict__ chunk_ptr,
                                                               const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                                               const double* __restrict__ val, const int* __restrict__ need_ptr,
                                                               int* __restrict__ need_rows, int* __restrict__ ent_ptr,
                                                               unsigned short* __restrict__ ent_col, double* __restrict__ ent_val,
                                                               unsigned short* __restrict__ self_slot, int* __restrict__ meta) {
    extern __shared__ int tb_smem[];
    int* keys = tb_smem;
    int* uniq = tb_smem + TB_MAXENT;
    __shared__ int scan[TB_THREADS];
    int cidx = blockIdx.x;
    int r0 = chunk_ptr[cidx], r1 = chunk_ptr[cidx + 1];
    int64_t e0 = ptr[r0], e1 = ptr[r1];
    int ne = (int)(e1 - e0);
    int nu = chunk_distinct_columns(idx, e0, ne, keys, uniq, scan);
    int n0 = need_ptr[cidx];
    for (int t = threadIdx.x; t < nu; t += TB_THREADS) need_rows[n0 + t] = uniq[t];
    for (int t = threadIdx.x; t < ne; t += TB_THREADS) {
        ent_col[e0 + t] = (unsigned short)lower_bound_dev(uniq, nu, idx[e0 + t]);
        ent_val[e0 + t] = val[e0 + t];
    }
    for (int r = r0 + threadIdx.x; r < r1; r += TB_THREADS) {
        ent_ptr[r] = (int)ptr[r];
        unsigned short s = 0xffff;
        if (square) {
            int p = lower_bound_dev(uniq, nu, r);
            if (p < nu && uniq[p] == r) s = (unsigned short)p;
        }
        self_slot[r] = s;
    }
    if (r1 == nrows && threadIdx.x == 0) ent_ptr[nrows] = (int)ptr[nrows];
    if (threadIdx.x == 0) {
        // own rows occupy consecutive slots iff every one of them is referenced (sorted distinct integers)
        int own = -1;
        if (square && r1 > r0) {
            int p = lower_bound_dev(uniq, nu, r0);
            int nr = r1 - r0;
            if (p + nr <= nu && uniq[p] == r0 && uniq[p + nr - 1] == r1 - 1) own = p;
        }
        int* mt = meta + (size_t)cidx * 8;
        mt[0] = r0; mt[1] = r1 - r0; mt[2] = n0; mt[3] = nu; mt[4] = (int)e0; mt[5] = ne; mt[6] = own; mt[7] = 0;
    }
}

// ------------------------------------------------------------------------------------------------ apply
struct TileDev {
    const int* meta; const int* need_rows; const int* ent_ptr;
    const unsigned short* ent_col; const double* ent_val; const unsigned short* self_slot;
    // byte offsets of the shared-memory regions (after the staged input rows at offset 0)
    unsigned off_bs, off_evals, off_dneed, off_drow, off_eptr, off_ecols, off_sself, off_red;
};

template <int VEC>
__device__ __forceinline__ void ldv(const double* __restrict__ p, bool ok0, bool ok1, double& x0, double& x1) {
    if (VEC == 2 && ok1) {
        double2 v = *reinterpret_cast<const double2*>(p);
        x0 = v.x; x1 = v.y;
    } else if (ok0) {
        x0 = p[0]; x1 = 0.0;
    } else {
        x0 = 0.0; x1 = 0.0;
    }
}
template <int VEC>
__device__ __forceinline__ void stv(double* __restrict__ p, bool ok0, bool ok1, double x0, double x1) {
    if (VEC == 2 && ok1) *reinterpret_cast<double2*>(p) = make_double2(x0, x1);
    else if (ok0) p[0] = x0;
}
// asynchronous global -> shared copy of this lane's VEC columns of one row (zero fill past column k)
template <int VEC>
__device__ __forceinline__ void cp_row(double* smem_dst, const double* __restrict__ src, bool ok0, bool ok1) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    if (VEC == 2) {
        int bytes = ok1 ? 16 : (ok0 ? 8 : 0);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(src), "r"(bytes));
    } else {
        int bytes = ok0 ? 8 : 0;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(src), "r"(bytes));
    }
}

// CW lanes share one row (each owning VEC adjacent columns); a warp instruction covers 32 / CW rows.
// Phase 1: every input the chunk needs -- the staged rows of X, the chunk's own rows of B (or of Y for TM_MULT_ADD),
// its matrix entries and smoother scalings -- is requested up front (cp.async, no register round trips), so a CTA has
// its whole working set in flight at once.  Phase 2 works from shared memory only and streams the result rows out.
template <int CW, int VEC, int MODE>
__global__ void __launch_bounds__(512) tileop_kernel(TileDev op, TileArgs a) {
    extern __shared__ __align__(16) unsigned char smraw[];
    constexpr int SW = CW * VEC;
    constexpr int RPW = 32 / CW;
    constexpr bool HAS_B = (MODE == TM_JACOBI || MODE == TM_JACOBI_DOT || MODE == TM_RESID || MODE == TM_MULT_ADD);
    constexpr bool HAS_D = (MODE == TM_JACOBI || MODE == TM_JACOBI_DOT || MODE == TM_PRESMOOTH2);
    constexpr bool NEED_SELF = (MODE == TM_PMULT_DOT || MODE == TM_JACOBI || MODE == TM_JACOBI_DOT || MODE == TM_PRESMOOTH2);
    double* xs = reinterpret_cast<double*>(smraw);
    double* bs = reinterpret_cast<double*>(smraw + op.off_bs);
    double* evals = reinterpret_cast<double*>(smraw + op.off_evals);
    double* dneed = reinterpret_cast<double*>(smraw + op.off_dneed);
    double* drow = reinterpret_cast<double*>(smraw + op.off_drow);
    int* eptr = reinterpret_cast<int*>(smraw + op.off_eptr);
    unsigned short* ecols = reinterpret_cast<unsigned short*>(smraw + op.off_ecols);
    unsigned short* sself = reinterpret_cast<unsigned short*>(smraw + op.off_sself);
    double (*red)[64] = reinterpret_cast<double (*)[64]>(smraw + op.off_red);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int sub = lane / CW, cl = lane % CW;
    // slab index fastest: CTAs dispatched together cover all column slabs of the same rows (whole DRAM pages)
    const int nslab = (a.k + SW - 1) / SW;
    const int slab = blockIdx.x % nslab, cidx = blockIdx.x / nslab;
    const int col = slab * SW + cl * VEC;
    const bool ok0 = col < a.k, ok1 = (VEC == 2) && (col + 1 < a.k);
    const int4 m0 = reinterpret_cast<const int4*>(op.meta)[2 * cidx], m1 = reinterpret_cast<const int4*>(op.meta)[2 * cidx + 1];
    const int r0 = m0.x, nr = m0.y, r1 = r0 + nr, n0 = m0.z, nneed = m0.w, e0 = m1.x, ne = m1.y, own = m1.z;
    const int sstride = nwarps * RPW;

    // ---- phase 1a: staged input rows
    if (MODE == TM_PMULT_DOT) {
        // p_j = z_j + beta p_j needs two inputs per row: through registers, 8 rows in flight per lane group
        constexpr int U = 8;
        double beta0 = ok0 ? a.beta[col] : 0.0, beta1 = ok1 ? a.beta[col + 1] : 0.0;
        for (int s = warp * RPW + sub; s < nneed; s += sstride * U) {
            int rows[U];
            double v0[U], v1[U], w0[U], w1[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                int ss = s + u * sstride;
                rows[u] = ss < nneed ? op.need_rows[n0 + ss] : -1;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                v0[u] = v1[u] = w0[u] = w1[u] = 0.0;
                if (rows[u] >= 0) {
                    ldv<VEC>(a.X + (int64_t)rows[u] * a.ldx + col, ok0, ok1, v0[u], v1[u]);
                    ldv<VEC>(a.Pin + (int64_t)rows[u] * a.ldp + col, ok0, ok1, w0[u], w1[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (rows[u] < 0) continue;
                int ss = s + u * sstride;
                double x0 = fma(beta0, w0[u], v0[u]), x1 = fma(beta1, w1[u], v1[u]);
                if (rows[u] >= r0 && rows[u] < r1) stv<VEC>(a.Pout + (int64_t)rows[u] * a.ldp + col, ok0, ok1, x0, x1);
                if (VEC == 2) *reinterpret_cast<double2*>(xs + (size_t)ss * SW + cl * 2) = make_double2(x0, x1);
                else xs[(size_t)ss * SW + cl] = x0;
            }
        }
    } else {
        // own rows sit in consecutive slots [own, own + nr): no index lookup, requested first
        if (own >= 0)
            for (int r = warp * RPW + sub; r < nr; r += sstride)
                cp_row<VEC>(xs + (size_t)(own + r) * SW + cl * VEC, a.X + (int64_t)(r0 + r) * a.ldx + col, ok0, ok1);
        constexpr int U = 8;
        const int nhalo = own >= 0 ? nneed - nr : nneed;
        for (int s = warp * RPW + sub; s < nhalo; s += sstride * U) {
            int rows[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                int hh = s + u * sstride;
                int ss = (own >= 0 && hh >= own) ? hh + nr : hh;
                rows[u] = hh < nhalo ? op.need_rows[n0 + ss] : -1;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (rows[u] < 0) continue;
                int hh = s + u * sstride;
                int ss = (own >= 0 && hh >= own) ? hh + nr : hh;
                cp_row<VEC>(xs + (size_t)ss * SW + cl * VEC, a.X + (int64_t)rows[u] * a.ldx + col, ok0, ok1);
            }
        }
        if (MODE == TM_PRESMOOTH2)
            for (int t = tid; t < nneed; t += blockDim.x) dneed[t] = a.dinv[op.need_rows[n0 + t]];
    }
    // ---- phase 1b: own rows of the right-hand side (or of Y when accumulating)
    if (HAS_B) {
        const double* src = (MODE == TM_MULT_ADD) ? a.Y : a.B;
        const int64_t lds = (MODE == TM_MULT_ADD) ? a.ldy : a.ldb;
        for (int r = warp * RPW + sub; r < nr; r += sstride)
            cp_row<VEC>(bs + (size_t)r * SW + cl * VEC, src + (int64_t)(r0 + r) * lds + col, ok0, ok1);
    }
    asm volatile("cp.async.commit_group;\n" ::);
    // ---- phase 1c: matrix entries, row pointers, scalings
    for (int t = tid; t < ne; t += blockDim.x) {
        evals[t] = op.ent_val[e0 + t];
        ecols[t] = op.ent_col[e0 + t];
    }
    for (int t = tid; t <= nr; t += blockDim.x) eptr[t] = op.ent_ptr[r0 + t] - e0;
    if (NEED_SELF)
        for (int t = tid; t < nr; t += blockDim.x) sself[t] = op.self_slot[r0 + t];
    if (HAS_D)
        for (int t = tid; t < nr; t += blockDim.x) drow[t] = a.dinv[r0 + t];
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();
    if (MODE == TM_PRESMOOTH2) {
        // the staged rows hold b_j; fold the first sweep s_j = d_j b_j into the matrix entries: a_ij d_j
        for (int t = tid; t < ne; t += blockDim.x) evals[t] *= dneed[ecols[t]];
        __syncthreads();
    }

    // ---- phase 2: rows, from shared memory only
    double d0 = 0.0, d1 = 0.0;
    for (int r = warp * RPW + sub; r < nr; r += sstride) {
        const int64_t row = r0 + r;
        const int ea = eptr[r], eb = eptr[r + 1];
        double acc0 = 0.0, acc1 = 0.0;
        int e = ea;
        for (; e + 4 <= eb; e += 4) {
            double x[4][2];
            double v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                v[q] = evals[e + q];
                const double* xp = xs + (size_t)ecols[e + q] * SW + cl * VEC;
                if (VEC == 2) { double2 t2 = *reinterpret_cast<const double2*>(xp); x[q][0] = t2.x; x[q][1] = t2.y; }
                else { x[q][0] = xp[0]; x[q][1] = 0.0; }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) { acc0 = fma(v[q], x[q][0], acc0); if (VEC == 2) acc1 = fma(v[q], x[q][1], acc1); }
        }
        for (; e < eb; ++e) {
            double v = evals[e];
            const double* xp = xs + (size_t)ecols[e] * SW + cl * VEC;
            if (VEC == 2) { double2 t2 = *reinterpret_cast<const double2*>(xp); acc0 = fma(v, t2.x, acc0); acc1 = fma(v, t2.y, acc1); }
            else acc0 = fma(v, xp[0], acc0);
        }
        double s0 = 0.0, s1 = 0.0;
        if (NEED_SELF) {
            unsigned short sl = sself[r];
            if (sl != 0xffff) {
                const double* xp = xs + (size_t)sl * SW + cl * VEC;
                s0 = xp[0];
                if (VEC == 2) s1 = xp[1];
            }
        }
        double b0 = 0.0, b1 = 0.0;
        if (HAS_B) {
            const double* bp = bs + (size_t)r * SW + cl * VEC;
            b0 = bp[0];
            if (VEC == 2) b1 = bp[1];
        }
        double y0, y1;
        if (MODE == TM_MULT) { y0 = acc0; y1 = acc1; }
        else if (MODE == TM_MULT_ADD) { y0 = b0 + acc0; y1 = b1 + acc1; }
        else if (MODE == TM_PMULT_DOT) { y0 = acc0; y1 = acc1; d0 = fma(s0, acc0, d0); d1 = fma(s1, acc1, d1); }
        else if (MODE == TM_RESID) { y0 = b0 - acc0; y1 = b1 - acc1; }
        else if (MODE == TM_PRESMOOTH2) {
            // s0 = b_i (raw), acc = sum_j a_ij d_j b_j:  y = d_i b_i + d_i (b_i - acc)
            double di = drow[r];
            y0 = di * (2.0 * s0 - acc0); y1 = di * (2.0 * s1 - acc1);
        } else {   // Jacobi
            double di = drow[r];
            y0 = fma(di, b0 - acc0, s0); y1 = fma(di, b1 - acc1, s1);
            if (MODE == TM_JACOBI_DOT) { d0 = fma(b0, y0, d0); d1 = fma(b1, y1, d1); }
        }
        stv<VEC>(a.Y + row * a.ldy + col, ok0, ok1, y0, y1);
    }
    if (MODE == TM_PMULT_DOT || MODE == TM_JACOBI_DOT) {
        red[warp][lane * VEC] = d0;
        if (VEC == 2) red[warp][lane * VEC + 1] = d1;
        __syncthreads();
        if (tid < SW) {
            int ccl = tid / VEC, v = tid % VEC;
            double s = 0.0;
            for (int w = 0; w < nwarps; ++w)
                for (int sb = 0; sb < RPW; ++sb) s += red[w][(sb * CW + ccl) * VEC + v];
            int cc = slab * SW + tid;
            if (cc < a.k) a.partials[(int64_t)cidx * a.k + cc] = s;
        }
    }
}

inline size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }

// shared-memory layout for one (operator, mode, slab width); returns the total and fills the offsets
size_t tile_layout(const TileOp& op, int mode, int SW, TileDev* d) {
    bool has_b = (mode == TM_JACOBI || mode == TM_JACOBI_DOT || mode == TM_RESID || mode == TM_MULT_ADD);
    size_t off = al16((size_t)op.max_need * SW * sizeof(double));
    unsigned o_bs = (unsigned)off;
    if (has_b) off += al16((size_t)op.max_rows * SW * sizeof(double));
    unsigned o_ev = (unsigned)off; off += al16((size_t)op.max_ent * sizeof(double));
    unsigned o_dn = (unsigned)off;
    if (mode == TM_PRESMOOTH2) off += al16((size_t)op.max_need * sizeof(double));
    unsigned o_dr = (unsigned)off; off += al16((size_t)op.max_rows * sizeof(double));
    unsigned o_ep = (unsigned)off; off += al16((size_t)(op.max_rows + 1) * sizeof(int));
    unsigned o_ec = (unsigned)off; off += al16((size_t)op.max_ent * sizeof(unsigned short));
    unsigned o_ss = (unsigned)off; off += al16((size_t)op.max_rows * sizeof(unsigned short));
    unsigned o_red = (unsigned)off;
    if (mode == TM_PMULT_DOT || mode == TM_JACOBI_DOT) off += 16 * 64 * sizeof(double);
    if (d) d->off_red = o_red;
    if (d) { d->off_bs = o_bs; d->off_evals = o_ev; d->off_dneed = o_dn; d->off_drow = o_dr; d->off_eptr = o_ep; d->off_ecols = o_ec; d->off_sself = o_ss; }
    return off;
}

constexpr size_t SMEM_STATIC = 1024;                    // per-CTA reservation
constexpr size_t SMEM_CAP = 227 * 1024 - SMEM_STATIC;   // per-CTA dynamic limit

template <int CW, int VEC, int MODE>
int launch_variant(sfem_ctx* c, const TileOp& op, TileDev d, const TileArgs& a) {
    constexpr int SW = CW * VEC;
    size_t smem = tile_layout(op, MODE, SW, &d);
    if (smem > SMEM_CAP) return sfem_fail(c, SFEM_ERR_LIMIT, "tileop: chunk working set (%zu bytes) exceeds shared memory", smem);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(tileop_kernel<CW, VEC, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_CAP);
        attr_set = true;
    }
    // two (or more) CTAs of 8 warps per SM when the working set allows it, otherwise one CTA of 16 warps
    int threads = (2 * (smem + SMEM_STATIC) <= 227 * 1024) ? 256 : 512;
    unsigned grid = (unsigned)op.nchunks * (unsigned)((a.k + SW - 1) / SW);
    tileop_kernel<CW, VEC, MODE><<<grid, threads, smem, c->stream>>>(d, a);
    KCHECK(c);
    return SFEM_OK;
}

template <int MODE>
int launch_mode(sfem_ctx* c, const TileOp& op, const TileDev& d, const TileArgs& a) {
    bool al2 = (a.ldx % 2 == 0) && (a.ldy % 2 == 0) && (reinterpret_cast<uintptr_t>(a.X) % 16 == 0) &&
               (reinterpret_cast<uintptr_t>(a.Y) % 16 == 0) &&
               (a.B == nullptr || (a.ldb % 2 == 0 && reinterpret_cast<uintptr_t>(a.B) % 16 == 0)) &&
               (a.Pin == nullptr || (a.ldp % 2 == 0 && reinterpret_cast<uintptr_t>(a.Pin) % 16 == 0 &&
                                     reinterpret_cast<uintptr_t>(a.Pout) % 16 == 0));
    if (a.k <= 4) return launch_variant<4, 1, MODE>(c, op, d, a);
    if (a.k <= 16) return launch_variant<16, 1, MODE>(c, op, d, a);
    if (a.k > 32 && al2 && c->tile_sw != 32 && 2 * (tile_layout(op, MODE, 64, nullptr) + SMEM_STATIC) <= 227 * 1024) return launch_variant<32, 2, MODE>(c, op, d, a);
    if (tile_layout(op, MODE, 32, nullptr) <= SMEM_CAP) return launch_variant<32, 1, MODE>(c, op, d, a);
    return launch_variant<16, 1, MODE>(c, op, d, a);
}

}  // namespace

void tileop_free(TileOp* op) {
    if (!op) return;
    void* ptrs[] = {op->chunk_ptr, op->need_ptr, op->need_rows, op->ent_ptr, op->ent_col, op->ent_val, op->self_slot, op->meta};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    *op = TileOp();
}

int tileop_build(sfem_ctx* c, int64_t nrows, int64_t ncols, const int64_t* ptr, const int32_t* idx, const double* val,
                 const int* chunk_ptr, int nchunks, int max_rows, TileOp* out) {
    tileop_free(out);
    TileOp op;
    op.nrows = nrows; op.ncols = ncols; op.nchunks = nchunks; op.max_rows = max_rows;
    int64_t nnz = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&nnz, ptr + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (nnz >= ((int64_t)1 << 31)) return sfem_fail(c, SFEM_ERR_LIMIT, "tileop: more than 2^31 entries");
    op.nnz = nnz;
    CUDA_TRY(c, cudaMalloc(&op.chunk_ptr, sizeof(int) * ((size_t)nchunks + 1)));
    CUDA_TRY(c, cudaMemcpyAsync(op.chunk_ptr, chunk_ptr, sizeof(int) * ((size_t)nchunks + 1), cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(c, cudaMalloc(&op.need_ptr, sizeof(int) * ((size_t)nchunks + 1)));
    int *need_cnt = nullptr, *sums = nullptr, *stats = nullptr;
    CUDA_TRY(c, cudaMalloc(&need_cnt, sizeof(int) * ((size_t)nchunks + 1)));
    CUDA_TRY(c, cudaMalloc(&sums, sizeof(int) * ((size_t)nchunks / 1024 + 8)));
    CUDA_TRY(c, cudaMalloc(&stats, sizeof(int) * 4));
    CUDA_TRY(c, cudaMemsetAsync(stats, 0, sizeof(int) * 4, c->stream));
    int rc = SFEM_OK;
    {
        ProfScope ps(c, TAG_SETUP, 0.0, 0.0);
        cudaFuncSetAttribute(tile_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * TB_MAXENT * sizeof(int)));
        cudaFuncSetAttribute(tile_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * TB_MAXENT * sizeof(int)));
        tile_count_kernel<<<nchunks, TB_THREADS, 2 * TB_MAXENT * sizeof(int), c->stream>>>(op.chunk_ptr, ptr, idx, need_cnt, stats);
        c->launches++;
        rc = exclusive_scan<int>(c, need_cnt, nchunks, op.need_ptr, sums);
    }
    int hstats[4] = {0, 0, 0, 0};
    int total_need = 0;
    if (rc == SFEM_OK) {
        cudaMemcpyAsync(hstats, stats, sizeof(hstats), cudaMemcpyDeviceToHost, c->stream);
        cudaMemcpyAsync(&total_need, op.need_ptr + nchunks, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) rc = sfem_fail(c, SFEM_ERR_CUDA, "tileop: count pass failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    cudaFree(need_cnt); cudaFree(sums); cudaFree(stats);
    if (rc) { tileop_free(&op); return rc; }
    if (hstats[2]) { tileop_free(&op); return sfem_fail(c, SFEM_ERR_LIMIT, "tileop: a chunk has too many entries or columns"); }
    op.max_need = hstats[0]; op.max_ent = hstats[1];
    CUDA_TRY(c, cudaMalloc(&op.need_rows, sizeof(int) * (size_t)(total_need > 0 ? total_need : 1)));
    CUDA_TRY(c, cudaMalloc(&op.ent_ptr, sizeof(int) * ((size_t)nrows + 1)));
    CUDA_TRY(c, cudaMalloc(&op.ent_col, sizeof(unsigned short) * (size_t)(nnz > 0 ? nnz : 1)));
    CUDA_TRY(c, cudaMalloc(&op.ent_val, sizeof(double) * (size_t)(nnz > 0 ? nnz : 1)));
    CUDA_TRY(c, cudaMalloc(&op.self_slot, sizeof(unsigned short) * (size_t)nrows));
    CUDA_TRY(c, cudaMalloc(&op.meta, sizeof(int) * 8 * (size_t)nchunks));
    {
        ProfScope ps(c, TAG_SETUP, 0.0, 0.0);
        tile_fill_kernel<<<nchunks, TB_THREADS, 2 * TB_MAXENT * sizeof(int), c->stream>>>(nrows, nrows == ncols ? 1 : 0, op.chunk_ptr, ptr, idx, val, op.need_ptr,
                                                                op.need_rows, op.ent_ptr, op.ent_col, op.ent_val, op.self_slot, op.meta);
        KCHECK(c);
    }
    *out = op;
    return SFEM_OK;
}

int tileop_apply(sfem_ctx* c, int tag, const TileOp& op, int mode, const TileArgs& a, double bytes) {
    if (op.nrows <= 0 || a.k <= 0) return SFEM_OK;
    TileDev d{op.meta, op.need_rows, op.ent_ptr, op.ent_col, op.ent_val, op.self_slot, 0, 0, 0, 0, 0, 0, 0, 0};
    ProfScope ps(c, tag, bytes, 2.0 * (double)op.nnz * a.k);
    switch (mode) {
        case TM_MULT: return launch_mode<TM_MULT>(c, op, d, a);
        case TM_MULT_ADD: return launch_mode<TM_MULT_ADD>(c, op, d, a);
        case TM_PMULT_DOT: return launch_mode<TM_PMULT_DOT>(c, op, d, a);
        case TM_JACOBI: return launch_mode<TM_JACOBI>(c, op, d, a);
        case TM_JACOBI_DOT: return launch_mode<TM_JACOBI_DOT>(c, op, d, a);
        case TM_RESID: return launch_mode<TM_RESID>(c, op, d, a);
        case TM_PRESMOOTH2: return launch_mode<TM_PRESMOOTH2>(c, op, d, a);
    }
    return sfem_fail(c, SFEM_ERR_ARG, "tileop_apply: bad mode");
}
