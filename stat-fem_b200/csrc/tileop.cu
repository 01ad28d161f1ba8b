// Staged-tile sparse operators: format builder and the fused apply kernels (see tileop.cuh).
#include "tileop.cuh"
#include <climits>
#include <cstdlib>

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

// pass 0: per-chunk counts.  stats[0] = max need, stats[1] = max padded row width, stats[2] = error flag
__global__ void __launch_bounds__(TB_THREADS) tile_count_kernel(const int* __restrict__ chunk_ptr, const int64_t* __restrict__ ptr,
                                                                const int32_t* __restrict__ idx, int* __restrict__ stats) {
    extern __shared__ int tb_smem[];
    int* keys = tb_smem;
    int* uniq = tb_smem + TB_MAXENT;
    __shared__ int scan[TB_THREADS];
    __shared__ int wmax;
    int cidx = blockIdx.x;
    int r0 = chunk_ptr[cidx], r1 = chunk_ptr[cidx + 1];
    int64_t e0 = ptr[r0], e1 = ptr[r1];
    if (e1 - e0 > TB_MAXENT) {
        if (threadIdx.x == 0) stats[2] = 1;
        return;
    }
    if (threadIdx.x == 0) wmax = 0;
    __syncthreads();
    for (int r = r0 + threadIdx.x; r < r1; r += TB_THREADS) atomicMax(&wmax, (int)(ptr[r + 1] - ptr[r]));
    int ne = (int)(e1 - e0);
    int nu = chunk_distinct_columns(idx, e0, ne, keys, uniq, scan);
    if (threadIdx.x == 0) {
        atomicMax(&stats[0], nu);
        atomicMax(&stats[1], (wmax + 1) & ~1);
        if (nu > 65535) stats[2] = 1;
    }
}

// pass 1: fixed-stride chunk records: descriptor, staged-row list, and the entries as a per-chunk ELL block
// (every row padded to the chunk's even width W with zero entries pointing at slot 0)
__global__ void __launch_bounds__(TB_THREADS) tile_fill_kernel(int square, int R, int dl, int n4, int e8,
                                                               const int* __restrict__ chunk_ptr, const int64_t* __restrict__ ptr,
                                                               const int32_t* __restrict__ idx, const double* __restrict__ val,
                                                               int* __restrict__ desc, int* __restrict__ need,
                                                               unsigned short* __restrict__ ecol, double* __restrict__ eval) {
    extern __shared__ int tb_smem[];
    int* keys = tb_smem;
    int* uniq = tb_smem + TB_MAXENT;
    __shared__ int scan[TB_THREADS];
    __shared__ int wmax;
    int cidx = blockIdx.x;
    int r0 = chunk_ptr[cidx], r1 = chunk_ptr[cidx + 1], nr = r1 - r0;
    int64_t e0 = ptr[r0], e1 = ptr[r1];
    int ne = (int)(e1 - e0);
    if (threadIdx.x == 0) wmax = 0;
    __syncthreads();
    for (int r = r0 + threadIdx.x; r < r1; r += TB_THREADS) atomicMax(&wmax, (int)(ptr[r + 1] - ptr[r]));
    int nu = chunk_distinct_columns(idx, e0, ne, keys, uniq, scan);
    const int W = (wmax + 1) & ~1;
    int* dsc = desc + (size_t)cidx * dl;
    int* nd = need + (size_t)cidx * n4;
    unsigned short* ec = ecol + (size_t)cidx * e8;
    double* ev = eval + (size_t)cidx * e8;
    for (int t = threadIdx.x; t < nu; t += TB_THREADS) nd[t] = uniq[t];
    for (int t = threadIdx.x; t < nr * W; t += TB_THREADS) {
        int r = t / W, w = t - r * W;
        int64_t p = ptr[r0 + r] + w;
        if (p < ptr[r0 + r + 1]) {
            ec[t] = (unsigned short)lower_bound_dev(uniq, nu, idx[p]);
            ev[t] = val[p];
        } else {
            ec[t] = 0;
            ev[t] = 0.0;
        }
    }
    unsigned short* selfs = reinterpret_cast<unsigned short*>(dsc + 8);
    for (int r = threadIdx.x; r < nr; r += TB_THREADS) {
        unsigned short s = 0xffff;
        if (square) {
            int p = lower_bound_dev(uniq, nu, r0 + r);
            if (p < nu && uniq[p] == r0 + r) s = (unsigned short)p;
        }
        selfs[r] = s;
    }
    if (threadIdx.x == 0) {
        // own rows occupy consecutive slots iff every one of them is referenced (sorted distinct integers)
        int own = 0, nown = 0;
        if (square && nr > 0) {
            int p = lower_bound_dev(uniq, nu, r0);
            if (p + nr <= nu && uniq[p] == r0 && uniq[p + nr - 1] == r1 - 1) { own = p; nown = nr; }
        }
        dsc[0] = r0; dsc[1] = nr; dsc[2] = nu; dsc[3] = W; dsc[4] = own; dsc[5] = nown; dsc[6] = 0; dsc[7] = 0;
    }
}

__global__ void tile_scale_kernel(int nchunks, int dl, int n4, int e8, const int* __restrict__ desc, const int* __restrict__ need,
                                  const unsigned short* __restrict__ ecol, const double* __restrict__ eval,
                                  const double* __restrict__ dinv, double* __restrict__ out) {
    int cidx = blockIdx.x;
    int ne = desc[(size_t)cidx * dl + 1] * desc[(size_t)cidx * dl + 3];
    for (int t = threadIdx.x; t < e8; t += blockDim.x) {
        size_t p = (size_t)cidx * e8 + t;
        out[p] = t < ne ? eval[p] * dinv[need[(size_t)cidx * n4 + ecol[p]]] : 0.0;
    }
}

// ------------------------------------------------------------------------------------------------ apply
struct TileDev {
    const int* desc; const int* need; const unsigned short* ecol; const double* eval;
    int nchunks, R, dl, n4, e8;
    // shared-memory geometry (bytes)
    unsigned cb_bytes, st_bytes, off_bs, off_drow, off_red;
};

template <int VEC>
__device__ __forceinline__ void stv(double* __restrict__ p, bool ok0, bool ok1, double x0, double x1) {
    if (VEC == 2 && ok1) *reinterpret_cast<double2*>(p) = make_double2(x0, x1);
    else if (ok0) p[0] = x0;
}
// asynchronous global -> shared copy of this lane's VEC columns of one row (`bytes` valid, zero fill beyond)
template <int VEC>
__device__ __forceinline__ void cp_row(unsigned smem_addr, const double* __restrict__ src, int bytes) {
    if (VEC == 2) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_addr), "l"(src), "r"(bytes));
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_addr), "l"(src), "r"(bytes));
}
__device__ __forceinline__ void cp16(void* smem_dst, const void* __restrict__ src) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(src));
}
__device__ __forceinline__ void cp8(void* smem_dst, const void* __restrict__ src) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(src));
}

// CW lanes share one row (each owning VEC adjacent columns); a warp instruction covers 32 / CW rows.
// Work items of a CTA: for each column slab, its chunks blockIdx.x + j * gridDim.x  (slab-major, so the per-lane dot
// accumulators run over all chunks of a slab and are reduced once per slab).
// BREG: the right-hand-side rows of the next work item are prefetched into REGISTERS one pipeline step ahead (each lane
// group owns at most 4 rows of a chunk) instead of being staged in shared memory: a third less shared memory per
// stage (more resident CTAs) and two fewer shared-memory passes per row.
template <int CW, int VEC, int MODE, bool BREG>
__global__ void __launch_bounds__(BREG ? 256 : 512, BREG ? 3 : 1) tileop_kernel(TileDev op, TileArgs a) {
    extern __shared__ __align__(16) unsigned char sm[];
    constexpr int SW = CW * VEC;
    constexpr int SWB = SW * 8;
    constexpr int RPW = 32 / CW;
    constexpr bool HAS_B = (MODE == TM_JACOBI || MODE == TM_JACOBI_DOT || MODE == TM_RESID || MODE == TM_MULT_ADD);
    constexpr bool HAS_D = (MODE == TM_JACOBI || MODE == TM_JACOBI_DOT || MODE == TM_PRESMOOTH2);
    constexpr bool NEED_SELF = (MODE == TM_MULT_DOT || MODE == TM_JACOBI || MODE == TM_JACOBI_DOT || MODE == TM_PRESMOOTH2);
    constexpr bool HAS_DOT = (MODE == TM_MULT_DOT || MODE == TM_JACOBI_DOT);
    unsigned char* cb0 = sm;                               // 4 chunk-record buffers
    unsigned char* st0 = sm + 4 * (size_t)op.cb_bytes;     // 2 stage buffers (staged rows, rhs rows, scalings)
    double (*red)[64] = reinterpret_cast<double (*)[64]>(sm + op.off_red);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int sub = lane / CW, cl = lane % CW;
    const int wsub = warp * RPW + sub;
    const int sstride = nwarps * RPW;
    const int nslab = (a.k + SW - 1) / SW;
    const int nmine = ((int)op.nchunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int T = nmine * nslab;
    const unsigned st_base = (unsigned)__cvta_generic_to_shared(st0) + cl * VEC * 8;

    auto issue_record = [&](int t) {
        if (t >= T) return;
        const int j = t % nmine;
        const size_t cidx = (size_t)blockIdx.x + (size_t)j * gridDim.x;
        unsigned char* dst = cb0 + (size_t)(t & 3) * op.cb_bytes;
        const int4* s0 = reinterpret_cast<const int4*>(op.desc + cidx * op.dl);
        const int4* s1 = reinterpret_cast<const int4*>(op.need + cidx * op.n4);
        const int4* s2 = reinterpret_cast<const int4*>(op.ecol + cidx * op.e8);
        const int4* s3 = reinterpret_cast<const int4*>(op.eval + cidx * op.e8);
        const int n0 = op.dl >> 2, n1 = op.n4 >> 2, n2 = op.e8 >> 3, n3 = op.e8 >> 1;
        int4* d = reinterpret_cast<int4*>(dst);
        for (int i = tid; i < n0; i += blockDim.x) cp16(d + i, s0 + i);
        d += n0;
        for (int i = tid; i < n1; i += blockDim.x) cp16(d + i, s1 + i);
        d += n1;
        for (int i = tid; i < n2; i += blockDim.x) cp16(d + i, s2 + i);
        d += n2;
        for (int i = tid; i < n3; i += blockDim.x) cp16(d + i, s3 + i);
    };
    auto issue_item = [&](int t) {
        if (t >= T) return;
        const int slab = t / nmine;
        const int* dsc = reinterpret_cast<const int*>(cb0 + (size_t)(t & 3) * op.cb_bytes);
        const int* need = dsc + op.dl;
        const int r0 = dsc[0], nr = dsc[1], nneed = dsc[2], own = dsc[4], nown = dsc[5];
        const unsigned xs_a = st_base + (unsigned)(t & 1) * op.st_bytes;
        const int col = slab * SW + cl * VEC;
        const int bytes = (VEC == 2) ? (col + 1 < a.k ? 16 : (col < a.k ? 8 : 0)) : (col < a.k ? 8 : 0);
        const double* Xc = a.X + col;
        const int64_t ldx = a.ldx;
        // own rows sit in consecutive slots [own, own + nown): no list lookup
        for (int r = wsub; r < nown; r += sstride) cp_row<VEC>(xs_a + (unsigned)(own + r) * SWB, Xc + (int64_t)(r0 + r) * ldx, bytes);
        for (int ss = wsub; ss < own; ss += sstride) cp_row<VEC>(xs_a + (unsigned)ss * SWB, Xc + (int64_t)need[ss] * ldx, bytes);
        for (int ss = own + nown + wsub; ss < nneed; ss += sstride) cp_row<VEC>(xs_a + (unsigned)ss * SWB, Xc + (int64_t)need[ss] * ldx, bytes);
        if (HAS_B && !BREG) {
            const unsigned bs_a = xs_a + op.off_bs;
            const double* Bc = ((MODE == TM_MULT_ADD) ? a.Y : a.B) + col;
            const int64_t lds = (MODE == TM_MULT_ADD) ? a.ldy : a.ldb;
            for (int r = wsub; r < nr; r += sstride) cp_row<VEC>(bs_a + (unsigned)r * SWB, Bc + (int64_t)(r0 + r) * lds, bytes);
        }
        if (HAS_D) {
            double* drow = reinterpret_cast<double*>(st0 + (size_t)(t & 1) * op.st_bytes + op.off_drow);
            for (int r = tid; r < nr; r += blockDim.x) cp8(drow + r, a.dinv + r0 + r);
            if (MODE == TM_PRESMOOTH2)
                for (int r = tid; r < nr; r += blockDim.x) cp8(drow + op.R + r, a.dinv2 + r0 + r);
        }
    };

    // ---- prologue: the first three chunk records, then the first work item in flight
    issue_record(0);
    issue_record(1);
    issue_record(2);
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();
    issue_item(0);
    asm volatile("cp.async.commit_group;\n" ::);

    // register prefetch of the right-hand-side rows (BREG): rows wsub + q * sstride, q < 4, of work item t
    double bn0[4] = {0.0, 0.0, 0.0, 0.0}, bn1[4] = {0.0, 0.0, 0.0, 0.0};
    auto prefetch_b = [&](int t) {
        if (!(HAS_B && BREG) || t >= T) return;
        const int slab = t / nmine;
        const int* dsc = reinterpret_cast<const int*>(cb0 + (size_t)(t & 3) * op.cb_bytes);
        const int r0 = dsc[0], nr = dsc[1];
        const int col = slab * SW + cl * VEC;
        const bool ok0 = col < a.k, ok1 = (VEC == 2) && (col + 1 < a.k);
        const double* Bc = ((MODE == TM_MULT_ADD) ? a.Y : a.B) + col;
        const int64_t lds = (MODE == TM_MULT_ADD) ? a.ldy : a.ldb;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int r = wsub + q * sstride;
            bn0[q] = bn1[q] = 0.0;
            if (r < nr) {
                const double* p = Bc + (int64_t)(r0 + r) * lds;
                if (VEC == 2 && ok1) { const double2 v = *reinterpret_cast<const double2*>(p); bn0[q] = v.x; bn1[q] = v.y; }
                else if (ok0) bn0[q] = p[0];
            }
        }
    };
    prefetch_b(0);

    double d0 = 0.0, d1 = 0.0;
    for (int t = 0; t < T; ++t) {
        double bc0[4], bc1[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { bc0[q] = bn0[q]; bc1[q] = bn1[q]; }
        // requests of the next work item and the chunk record three items ahead (its buffer was last read by item t-1)
        issue_item(t + 1);
        prefetch_b(t + 1);
        issue_record(t + 3);
        asm volatile("cp.async.commit_group;\n" ::);
        asm volatile("cp.async.wait_group 1;\n" ::);
        __syncthreads();

        const int slab = t / nmine;
        const unsigned char* cb = cb0 + (size_t)(t & 3) * op.cb_bytes;
        const int* dsc = reinterpret_cast<const int*>(cb);
        const int r0 = dsc[0], nr = dsc[1], W = dsc[3];
        const unsigned short* sself = reinterpret_cast<const unsigned short*>(dsc + 8);
        const unsigned short* ecols = reinterpret_cast<const unsigned short*>(dsc + op.dl + op.n4);
        const double* evals = reinterpret_cast<const double*>(cb + (size_t)(op.dl + op.n4) * 4 + (size_t)op.e8 * 2);
        const unsigned char* st = st0 + (size_t)(t & 1) * op.st_bytes;
        const double* xs = reinterpret_cast<const double*>(st) + cl * VEC;
        const double* bs = reinterpret_cast<const double*>(st + op.off_bs) + cl * VEC;
        const double* drow = reinterpret_cast<const double*>(st + op.off_drow);
        const int col = slab * SW + cl * VEC;
        const bool ok0 = col < a.k, ok1 = (VEC == 2) && (col + 1 < a.k);
        double* Yc = a.Y + col;

        // epilogue of one row: fuses the caller's vector work and streams the result row out
        auto finish_row = [&](int r, double acc0, double acc1, double br0, double br1) {
            double s0 = 0.0, s1 = 0.0;
            if (NEED_SELF) {
                const unsigned sl = sself[r];
                if (sl != 0xffffu) {
                    const double* xp = xs + sl * SW;
                    s0 = xp[0];
                    if (VEC == 2) s1 = xp[1];
                }
            }
            double b0 = br0, b1 = br1;
            if (HAS_B && !BREG) {
                const double* bp = bs + r * SW;
                b0 = bp[0];
                if (VEC == 2) b1 = bp[1];
            }
            double y0, y1;
            if (MODE == TM_MULT) { y0 = acc0; y1 = acc1; }
            else if (MODE == TM_MULT_ADD) { y0 = b0 + acc0; y1 = b1 + acc1; }
            else if (MODE == TM_MULT_DOT) { y0 = acc0; y1 = acc1; d0 = fma(s0, acc0, d0); d1 = fma(s1, acc1, d1); }
            else if (MODE == TM_RESID) { y0 = b0 - acc0; y1 = b1 - acc1; }
            else if (MODE == TM_PRESMOOTH2) {
                // s = b_i, acc = sum_j (a_ij d_j) b_j:  y = d_i b_i + d2_i (b_i - acc)
                const double di = drow[r], di2 = drow[op.R + r];
                y0 = fma(di2, s0 - acc0, di * s0); y1 = fma(di2, s1 - acc1, di * s1);
            } else {   // Jacobi
                const double di = drow[r];
                y0 = fma(di, b0 - acc0, s0); y1 = fma(di, b1 - acc1, s1);
                if (MODE == TM_JACOBI_DOT) { d0 = fma(b0, y0, d0); d1 = fma(b1, y1, d1); }
            }
            stv<VEC>(Yc + (int64_t)(r0 + r) * a.ldy, ok0, ok1, y0, y1);
        };
        // two rows per pass and per lane group: twice the independent shared-memory -> DFMA chains in flight
        auto two_rows = [&](int r, double ba0, double ba1, double bb0, double bb1) {
            const int rb = r + sstride;
            const bool two = rb < nr;
            const unsigned short* eca = ecols + r * W;
            const double* eva = evals + r * W;
            const unsigned short* ecb = ecols + (two ? rb : r) * W;
            const double* evb = evals + (two ? rb : r) * W;
            double a0 = 0.0, a1 = 0.0, c0 = 0.0, c1 = 0.0;
#pragma unroll 2
            for (int w = 0; w < W; w += 2) {
                const unsigned ca = *reinterpret_cast<const unsigned*>(eca + w);
                const unsigned cb2 = *reinterpret_cast<const unsigned*>(ecb + w);
                const double2 va = *reinterpret_cast<const double2*>(eva + w);
                const double2 vb = *reinterpret_cast<const double2*>(evb + w);
                const double* xa0 = xs + (ca & 0xffffu) * SW;
                const double* xa1 = xs + (ca >> 16) * SW;
                const double* xb0 = xs + (cb2 & 0xffffu) * SW;
                const double* xb1 = xs + (cb2 >> 16) * SW;
                if (VEC == 2) {
                    const double2 ta0 = *reinterpret_cast<const double2*>(xa0), ta1 = *reinterpret_cast<const double2*>(xa1);
                    const double2 tb0 = *reinterpret_cast<const double2*>(xb0), tb1 = *reinterpret_cast<const double2*>(xb1);
                    a0 = fma(va.x, ta0.x, a0); a1 = fma(va.x, ta0.y, a1);
                    c0 = fma(vb.x, tb0.x, c0); c1 = fma(vb.x, tb0.y, c1);
                    a0 = fma(va.y, ta1.x, a0); a1 = fma(va.y, ta1.y, a1);
                    c0 = fma(vb.y, tb1.x, c0); c1 = fma(vb.y, tb1.y, c1);
                } else {
                    a0 = fma(va.x, xa0[0], a0); c0 = fma(vb.x, xb0[0], c0);
                    a0 = fma(va.y, xa1[0], a0); c0 = fma(vb.y, xb1[0], c0);
                }
            }
            finish_row(r, a0, a1, ba0, ba1);
            if (two) finish_row(rb, c0, c1, bb0, bb1);
        };
        if (BREG) {      // at most 4 rows per lane group: two passes with compile-time register indices
            if (wsub < nr) two_rows(wsub, bc0[0], bc1[0], bc0[1], bc1[1]);
            if (wsub + 2 * sstride < nr) two_rows(wsub + 2 * sstride, bc0[2], bc1[2], bc0[3], bc1[3]);
        } else {
            for (int r = wsub; r < nr; r += 2 * sstride) two_rows(r, 0.0, 0.0, 0.0, 0.0);
        }
        const bool slab_done = HAS_DOT && ((t + 1) % nmine == 0);
        if (slab_done) {
            red[warp][lane * VEC] = d0;
            if (VEC == 2) red[warp][lane * VEC + 1] = d1;
            d0 = d1 = 0.0;
        }
        __syncthreads();
        if (slab_done && tid < SW) {
            const int ccl = tid / VEC, v = tid % VEC;
            double s = 0.0;
            for (int w = 0; w < nwarps; ++w)
                for (int sb = 0; sb < RPW; ++sb) s += red[w][(sb * CW + ccl) * VEC + v];
            const int cc = slab * SW + tid;
            if (cc < a.k) a.partials[(size_t)blockIdx.x * a.k + cc] = s;
        }
    }
    asm volatile("cp.async.wait_group 0;\n" ::);
}

// Fallback for operators whose chunk working set does not fit in shared memory (e.g. wide 3-D transfer operators):
// same format, rows gathered straight from global memory (warp per row, lanes across 32 columns, slabs looped).
template <int MODE>
__global__ void __launch_bounds__(256) tileop_direct_kernel(TileDev op, TileArgs a) {
    constexpr bool HAS_B = (MODE == TM_JACOBI || MODE == TM_JACOBI_DOT || MODE == TM_RESID || MODE == TM_MULT_ADD);
    constexpr bool HAS_D = (MODE == TM_JACOBI || MODE == TM_JACOBI_DOT || MODE == TM_PRESMOOTH2);
    constexpr bool NEED_SELF = (MODE == TM_MULT_DOT || MODE == TM_JACOBI || MODE == TM_JACOBI_DOT || MODE == TM_PRESMOOTH2);
    constexpr bool HAS_DOT = (MODE == TM_MULT_DOT || MODE == TM_JACOBI_DOT);
    __shared__ double red[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nslab = (a.k + 31) / 32;
    for (int slab = 0; slab < nslab; ++slab) {
        const int col = slab * 32 + lane;
        const bool ok = col < a.k;
        double d0 = 0.0;
        for (int cidx = blockIdx.x; cidx < op.nchunks; cidx += gridDim.x) {
            const int* dsc = op.desc + (size_t)cidx * op.dl;
            const int* need = op.need + (size_t)cidx * op.n4;
            const unsigned short* ecols = op.ecol + (size_t)cidx * op.e8;
            const double* evals = op.eval + (size_t)cidx * op.e8;
            const unsigned short* sself = reinterpret_cast<const unsigned short*>(dsc + 8);
            const int r0 = dsc[0], nr = dsc[1], W = dsc[3];
            for (int r = warp; r < nr; r += 8) {
                double acc = 0.0;
                for (int w = 0; w < W; ++w) {
                    double v = evals[r * W + w];
                    if (v != 0.0 && ok) acc = fma(v, a.X[(int64_t)need[ecols[r * W + w]] * a.ldx + col], acc);
                }
                double s0 = 0.0, b0 = 0.0, di = 0.0;
                const int64_t row = r0 + r;
                if (NEED_SELF && sself[r] != 0xffff && ok) s0 = a.X[(int64_t)need[sself[r]] * a.ldx + col];
                if (HAS_B && ok) b0 = (MODE == TM_MULT_ADD) ? a.Y[row * a.ldy + col] : a.B[row * a.ldb + col];
                if (HAS_D) di = a.dinv[row];
                double y;
                if (MODE == TM_MULT) y = acc;
                else if (MODE == TM_MULT_ADD) y = b0 + acc;
                else if (MODE == TM_MULT_DOT) { y = acc; d0 = fma(s0, acc, d0); }
                else if (MODE == TM_RESID) y = b0 - acc;
                else if (MODE == TM_PRESMOOTH2) y = fma(a.dinv2[row], s0 - acc, di * s0);
                else { y = fma(di, b0 - acc, s0); if (MODE == TM_JACOBI_DOT) d0 = fma(b0, y, d0); }
                if (ok) a.Y[row * a.ldy + col] = y;
            }
        }
        if (HAS_DOT) {
            red[warp][lane] = d0;
            __syncthreads();
            if (warp == 0 && ok) {
                double s = 0.0;
                for (int w = 0; w < 8; ++w) s += red[w][lane];
                a.partials[(size_t)blockIdx.x * a.k + col] = s;
            }
            __syncthreads();
        }
    }
}

inline size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }

// shared-memory geometry for one (operator, mode, slab width); returns the total bytes
size_t tile_layout(const TileOp& op, int mode, int SW, TileDev* d, bool breg = false) {
    bool has_b = (mode == TM_JACOBI || mode == TM_JACOBI_DOT || mode == TM_RESID || mode == TM_MULT_ADD);
    bool has_d = (mode == TM_JACOBI || mode == TM_JACOBI_DOT || mode == TM_PRESMOOTH2);
    bool has_dot = (mode == TM_MULT_DOT || mode == TM_JACOBI_DOT);
    size_t cb = (size_t)op.dl * 4 + (size_t)op.n4 * 4 + (size_t)op.e8 * 2 + (size_t)op.e8 * 8;
    size_t st = al16((size_t)op.max_need * SW * sizeof(double));
    size_t o_bs = st;
    if (has_b && !breg) st += al16((size_t)op.max_rows * SW * sizeof(double));
    size_t o_dr = st;
    if (has_d) st += al16((size_t)op.max_rows * sizeof(double) * (mode == TM_PRESMOOTH2 ? 2 : 1));
    size_t tot = 4 * cb + 2 * st;
    size_t o_red = tot;
    if (has_dot) tot += 16 * 64 * sizeof(double);
    if (d) { d->cb_bytes = (unsigned)cb; d->st_bytes = (unsigned)st; d->off_bs = (unsigned)o_bs; d->off_drow = (unsigned)o_dr; d->off_red = (unsigned)o_red; }
    return tot;
}

constexpr size_t SMEM_RESERVED = 1024;                    // per-CTA reservation of the hardware
constexpr size_t SMEM_CAP = 227 * 1024 - SMEM_RESERVED;   // per-CTA dynamic limit

template <int CW, int VEC, int MODE, bool BREG>
int launch_kernel(sfem_ctx* c, const TileOp& op, TileDev d, const TileArgs& a, int* nparts) {
    constexpr int SW = CW * VEC;
    size_t smem = tile_layout(op, MODE, SW, &d, BREG);
    if (smem > SMEM_CAP) return sfem_fail(c, SFEM_ERR_LIMIT, "tileop: chunk working set (%zu bytes) exceeds shared memory", smem);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(tileop_kernel<CW, VEC, MODE, BREG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_CAP);
        attr_set = true;
    }
    // persistent grid: as many CTAs as fit on the machine (8 warps each; one 16-warp CTA per SM if only one fits)
    int per_sm = (int)((227 * 1024) / (smem + SMEM_RESERVED));
    int threads = 256;
    if (per_sm < 2) { per_sm = 1; threads = BREG ? 256 : 512; }
    if (per_sm > TILE_MAX_CTAS_PER_SM) per_sm = TILE_MAX_CTAS_PER_SM;
    int64_t items = (int64_t)op.nchunks;
    int grid = (int)(items < (int64_t)per_sm * SFEM_NUM_SMS ? items : (int64_t)per_sm * SFEM_NUM_SMS);
    if (nparts) *nparts = grid;
    tileop_kernel<CW, VEC, MODE, BREG><<<grid, threads, smem, c->stream>>>(d, a);
    KCHECK(c);
    return SFEM_OK;
}

template <int CW, int VEC, int MODE>
int launch_variant(sfem_ctx* c, const TileOp& op, const TileDev& d, const TileArgs& a, int* nparts) {
    constexpr bool has_b = (MODE == TM_JACOBI || MODE == TM_JACOBI_DOT || MODE == TM_RESID || MODE == TM_MULT_ADD);
    // register prefetch of the right-hand side needs at most 4 rows per lane group with 8 warps per CTA
    if (has_b && op.max_rows <= 4 * 8 * (32 / CW) && c->tile_sw != 33) return launch_kernel<CW, VEC, MODE, true>(c, op, d, a, nparts);
    return launch_kernel<CW, VEC, MODE, false>(c, op, d, a, nparts);
}

template <int MODE>
int launch_mode(sfem_ctx* c, const TileOp& op, const TileDev& d, const TileArgs& a, int* nparts) {
    bool al2 = (a.ldx % 2 == 0) && (a.ldy % 2 == 0) && (reinterpret_cast<uintptr_t>(a.X) % 16 == 0) &&
               (reinterpret_cast<uintptr_t>(a.Y) % 16 == 0) &&
               (a.B == nullptr || (a.ldb % 2 == 0 && reinterpret_cast<uintptr_t>(a.B) % 16 == 0));
    if (a.k <= 4 && tile_layout(op, MODE, 4, nullptr) <= SMEM_CAP) return launch_variant<4, 1, MODE>(c, op, d, a, nparts);
    if (a.k <= 16 && tile_layout(op, MODE, 16, nullptr) <= SMEM_CAP) return launch_variant<16, 1, MODE>(c, op, d, a, nparts);
    if (a.k > 32 && al2 && c->tile_sw != 32 && 2 * (tile_layout(op, MODE, 64, nullptr) + SMEM_RESERVED) <= 227 * 1024) return launch_variant<32, 2, MODE>(c, op, d, a, nparts);
    if (tile_layout(op, MODE, 32, nullptr) <= SMEM_CAP) return launch_variant<32, 1, MODE>(c, op, d, a, nparts);
    if (tile_layout(op, MODE, 16, nullptr) <= SMEM_CAP) return launch_variant<16, 1, MODE>(c, op, d, a, nparts);
    // working set too large for shared memory: gather from global memory
    int grid = op.nchunks < 4 * SFEM_NUM_SMS ? op.nchunks : 4 * SFEM_NUM_SMS;
    if (nparts) *nparts = grid;
    tileop_direct_kernel<MODE><<<grid, 256, 0, c->stream>>>(d, a);
    KCHECK(c);
    return SFEM_OK;
}

}  // namespace

void tileop_free(TileOp* op) {
    if (!op) return;
    void* ptrs[] = {op->desc, op->need, op->ecol, op->eval, op->eval_scaled};
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
    int* stats = nullptr;
    CUDA_TRY(c, cudaMalloc(&stats, sizeof(int) * 4));
    CUDA_TRY(c, cudaMemsetAsync(stats, 0, sizeof(int) * 4, c->stream));
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(tile_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * TB_MAXENT * sizeof(int)));
        cudaFuncSetAttribute(tile_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * TB_MAXENT * sizeof(int)));
        attr_set = true;
    }
    {
        ProfScope ps(c, TAG_SETUP, 0.0, 0.0);
        tile_count_kernel<<<nchunks, TB_THREADS, 2 * TB_MAXENT * sizeof(int), c->stream>>>(chunk_ptr, ptr, idx, stats);
        c->launches++;
    }
    int hstats[4] = {0, 0, 0, 0};
    cudaMemcpyAsync(hstats, stats, sizeof(hstats), cudaMemcpyDeviceToHost, c->stream);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    cudaFree(stats);
    if (e != cudaSuccess) return sfem_fail(c, SFEM_ERR_CUDA, "tileop: count pass failed: %s", cudaGetErrorString(e));
    if (hstats[2]) return sfem_fail(c, SFEM_ERR_LIMIT, "tileop: a chunk has too many entries or columns");
    op.nnz = nnz;
    op.max_need = hstats[0] > 0 ? hstats[0] : 1;
    op.max_ent = max_rows * (hstats[1] > 0 ? hstats[1] : 2);
    op.dl = (8 + (max_rows + 1) / 2 + 3) & ~3;
    op.n4 = (op.max_need + 3) & ~3;
    op.e8 = (op.max_ent + 7) & ~7;
    size_t nc = (size_t)nchunks;
    CUDA_TRY(c, cudaMalloc(&op.desc, sizeof(int) * nc * op.dl));
    CUDA_TRY(c, cudaMalloc(&op.need, sizeof(int) * nc * op.n4));
    CUDA_TRY(c, cudaMalloc(&op.ecol, sizeof(unsigned short) * nc * op.e8));
    CUDA_TRY(c, cudaMalloc(&op.eval, sizeof(double) * nc * op.e8));
    CUDA_TRY(c, cudaMemsetAsync(op.desc, 0, sizeof(int) * nc * op.dl, c->stream));
    CUDA_TRY(c, cudaMemsetAsync(op.need, 0, sizeof(int) * nc * op.n4, c->stream));
    CUDA_TRY(c, cudaMemsetAsync(op.ecol, 0, sizeof(unsigned short) * nc * op.e8, c->stream));
    CUDA_TRY(c, cudaMemsetAsync(op.eval, 0, sizeof(double) * nc * op.e8, c->stream));
    {
        ProfScope ps(c, TAG_SETUP, 0.0, 0.0);
        tile_fill_kernel<<<nchunks, TB_THREADS, 2 * TB_MAXENT * sizeof(int), c->stream>>>(nrows == ncols ? 1 : 0, max_rows, op.dl, op.n4, op.e8,
                                                                                        chunk_ptr, ptr, idx, val, op.desc, op.need, op.ecol, op.eval);
        KCHECK(c);
    }
    if (getenv("SFEM_DEBUG")) fprintf(stderr, "tileop %lld x %lld nnz %lld chunks %d R %d need %d W %d dl %d\n", (long long)nrows, (long long)ncols, (long long)nnz, nchunks, max_rows, op.max_need, hstats[1], op.dl);
    *out = op;
    return SFEM_OK;
}

int tileop_scale_columns(sfem_ctx* c, TileOp* op, const double* dinv) {
    if (op->nchunks <= 0) return SFEM_OK;
    if (!op->eval_scaled) CUDA_TRY(c, cudaMalloc(&op->eval_scaled, sizeof(double) * (size_t)op->nchunks * op->e8));
    ProfScope ps(c, TAG_SETUP, 0.0, 0.0);
    tile_scale_kernel<<<op->nchunks, 128, 0, c->stream>>>(op->nchunks, op->dl, op->n4, op->e8, op->desc, op->need, op->ecol, op->eval, dinv, op->eval_scaled);
    KCHECK(c);
    return SFEM_OK;
}

int tileop_apply(sfem_ctx* c, int tag, const TileOp& op, int mode, const TileArgs& a_in, double bytes, int* nparts) {
    if (nparts) *nparts = 0;
    if (op.nrows <= 0 || a_in.k <= 0) return SFEM_OK;
    TileArgs a = a_in;
    if (!a.dinv2) a.dinv2 = a.dinv;
    TileDev d;
    d.desc = op.desc; d.need = op.need; d.ecol = op.ecol;
    d.eval = (mode == TM_PRESMOOTH2) ? op.eval_scaled : op.eval;
    if (!d.eval) return sfem_fail(c, SFEM_ERR_STATE, "tileop_apply: scaled entries not built");
    d.nchunks = op.nchunks; d.R = op.max_rows; d.dl = op.dl; d.n4 = op.n4; d.e8 = op.e8;
    ProfScope ps(c, tag, bytes, 2.0 * (double)op.nnz * a.k);
    switch (mode) {
        case TM_MULT: return launch_mode<TM_MULT>(c, op, d, a, nparts);
        case TM_MULT_ADD: return launch_mode<TM_MULT_ADD>(c, op, d, a, nparts);
        case TM_MULT_DOT: return launch_mode<TM_MULT_DOT>(c, op, d, a, nparts);
        case TM_JACOBI: return launch_mode<TM_JACOBI>(c, op, d, a, nparts);
        case TM_JACOBI_DOT: return launch_mode<TM_JACOBI_DOT>(c, op, d, a, nparts);
        case TM_RESID: return launch_mode<TM_RESID>(c, op, d, a, nparts);
        case TM_PRESMOOTH2: return launch_mode<TM_PRESMOOTH2>(c, op, d, a, nparts);
    }
    return sfem_fail(c, SFEM_ERR_ARG, "tileop_apply: bad mode");
}
