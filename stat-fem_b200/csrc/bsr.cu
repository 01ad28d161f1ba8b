// W = G Z on FP64 tensor cores: the "8x4 block" form of a CSR matrix and its SpMM kernel.
//
// G has ~100-300 entries per row (cut-off squared-exponential covariance): W = G Z is FP64-bound, not HBM-bound
// (SURVEY 8d), and the warp-per-row CSR kernel of spmm.cu re-reads every gathered row of Z once per matrix entry
// from L1/L2.  Here 8 consecutive rows of G form a row group; the sorted union of their column indices is cut into
// k-tiles of 4 columns, and each k-tile is stored as a dense 8x4 block (explicit zeros where a row has no entry) in
// the lane order of the A operand of mma.m8n8k4.f64 plus its 4 column indices.  A warp owns one row group and 64
// columns of Z at a time (eight 8x8 accumulator tiles): per k-tile it loads its block entry (one coalesced 256-byte
// request), the column index of a B row, the four B rows straight from Z with coalesced 16-byte loads (a shuffle puts
// the pieces into the fragment layout) and issues eight DMMAs -- a gathered value of Z now serves 8 rows of G.
// Neighbouring rows of a mesh-ordered G share most of their columns (a row group of a 2-D lattice mesh with cut
// radius 5.6 h touches ~180 columns for 8 x 97 entries), so the dense blocks are 50-70 % full.
//
// Replaces PETSc MatMult at ForcingCovariance.py:242 for blocks of right-hand sides (n_obs columns at once).
#include "common.cuh"
#include <climits>

template <typename T>
int exclusive_scan(sfem_ctx* c, const T* in, int64_t n, T* out, T* sums);

struct B84Matrix {
    int64_t n = 0, nrg = 0, ntiles = 0, nnz = 0;
    int64_t* tptr = nullptr;      // [nrg + 1] first k-tile of every row group
    int32_t* cols = nullptr;      // [ntiles][4]
    double* vals = nullptr;       // [ntiles][32]  block (row lr, column q) at lr * 4 + q
};

namespace {

// 8-way merge of the (ascending) column lists of rows [8 g, 8 g + 8): calls f(position, column, row-in-group, entry)
// for every stored entry, positions counting the distinct columns.  Returns the number of distinct columns.
template <class F>
__device__ __forceinline__ int merge_group(int64_t n, int64_t g, const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx, F f) {
    int64_t hp[8], he[8];
    int ck[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int64_t row = g * 8 + r;
        hp[r] = row < n ? ptr[row] : 0;
        he[r] = row < n ? ptr[row + 1] : 0;
        ck[r] = hp[r] < he[r] ? idx[hp[r]] : INT_MAX;
    }
    int pos = 0;
    while (true) {
        int m = ck[0];
#pragma unroll
        for (int r = 1; r < 8; ++r) m = min(m, ck[r]);
        if (m == INT_MAX) break;
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if (ck[r] == m) {
                f(pos, m, r, hp[r]);
                ++hp[r];
                ck[r] = hp[r] < he[r] ? idx[hp[r]] : INT_MAX;
            }
        ++pos;
    }
    return pos;
}

// Builder: one warp per row group.  The distinct columns of the group are found with a bitmap over its column span in
// shared memory (atomicOr per entry, popcount prefix per word): every entry then knows its position in the sorted union
// without any serial merge.  Groups whose columns span more than B84_SPAN indices fall back to the merge on one lane.
constexpr int B84_SPAN = 65536;                 // bitmap bits per warp (8 KB) + as many prefix entries / 32 (8 KB)
constexpr int B84_WARPS = 2;                    // warps per builder CTA (32 KB of static shared memory)

struct GroupSpan { int lo, hi; };
__device__ __forceinline__ GroupSpan group_span(int64_t n, int64_t g, const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx, int lane) {
    int lo = INT_MAX, hi = -1;
    if (lane < 8) {
        const int64_t row = g * 8 + lane;
        if (row < n && ptr[row] < ptr[row + 1]) { lo = idx[ptr[row]]; hi = idx[ptr[row + 1] - 1]; }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    GroupSpan s;
    s.lo = __shfl_sync(0xffffffffu, lo, 0);
    s.hi = __shfl_sync(0xffffffffu, hi, 0);
    return s;
}
// marks the columns of the group in bm (words [0, nw) zeroed first); all lanes of the warp
__device__ __forceinline__ void group_bitmap(int64_t n, int64_t g, const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx, int lo, int nw,
                                             unsigned* bm, int lane) {
    for (int w = lane; w < nw; w += 32) bm[w] = 0u;
    __syncwarp();
    for (int r = 0; r < 8; ++r) {
        const int64_t row = g * 8 + r;
        if (row >= n) break;
        for (int64_t e = ptr[row] + lane; e < ptr[row + 1]; e += 32) {
            const int cb = idx[e] - lo;
            atomicOr(&bm[cb >> 5], 1u << (cb & 31));
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(32 * B84_WARPS) b84_count_kernel(int64_t n, int64_t nrg, const int64_t* __restrict__ ptr,
                                                                   const int32_t* __restrict__ idx, int64_t* __restrict__ ntile) {
    __shared__ unsigned bms[B84_WARPS][B84_SPAN / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t g = (int64_t)blockIdx.x * B84_WARPS + warp;
    if (g >= nrg) return;
    const GroupSpan sp = group_span(n, g, ptr, idx, lane);
    int u = 0;
    if (sp.hi < sp.lo) u = 0;
    else if (sp.hi - sp.lo >= B84_SPAN) {
        if (lane == 0) u = merge_group(n, g, ptr, idx, [](int, int, int, int64_t) {});
        u = __shfl_sync(0xffffffffu, u, 0);
    } else {
        const int nw = ((sp.hi - sp.lo) >> 5) + 1;
        group_bitmap(n, g, ptr, idx, sp.lo, nw, bms[warp], lane);
        for (int w = lane; w < nw; w += 32) u += __popc(bms[warp][w]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) u += __shfl_xor_sync(0xffffffffu, u, o);
    }
    if (lane == 0) ntile[g] = (u + 3) >> 2;
}

// vals must be zero on entry
__global__ void __launch_bounds__(32 * B84_WARPS) b84_fill_kernel(int64_t n, int64_t nrg, const int64_t* __restrict__ ptr,
                                                                  const int32_t* __restrict__ idx, const double* __restrict__ val,
                                                                  const int64_t* __restrict__ tptr, int32_t* __restrict__ cols,
                                                                  double* __restrict__ vals) {
    __shared__ unsigned bms[B84_WARPS][B84_SPAN / 32];
    __shared__ int pres[B84_WARPS][B84_SPAN / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t g = (int64_t)blockIdx.x * B84_WARPS + warp;
    if (g >= nrg) return;
    const int64_t t0 = tptr[g];
    const GroupSpan sp = group_span(n, g, ptr, idx, lane);
    if (sp.hi < sp.lo) return;
    if (sp.hi - sp.lo >= B84_SPAN) {          // wide group: serial merge on one lane
        if (lane == 0) {
            int lastc = 0;
            const int u = merge_group(n, g, ptr, idx, [&](int pos, int col, int r, int64_t e) {
                const int64_t t = t0 + (pos >> 2);
                const int q = pos & 3;
                cols[t * 4 + q] = col;
                vals[t * 32 + r * 4 + q] += val[e];
                lastc = col;
            });
            for (int pos = u; pos & 3; ++pos) cols[(t0 + (pos >> 2)) * 4 + (pos & 3)] = lastc;
        }
        return;
    }
    unsigned* bm = bms[warp];
    int* pre = pres[warp];
    const int nw = ((sp.hi - sp.lo) >> 5) + 1;
    group_bitmap(n, g, ptr, idx, sp.lo, nw, bm, lane);
    // exclusive prefix of the word popcounts: position of a word's first column in the sorted union
    int running = 0;
    for (int base = 0; base < nw; base += 32) {
        const int w = base + lane;
        const int cnt = w < nw ? __popc(bm[w]) : 0;
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (w < nw) pre[w] = running + incl - cnt;
        running += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
    const int u = running;
    for (int r = 0; r < 8; ++r) {
        const int64_t row = g * 8 + r;
        if (row >= n) break;
        for (int64_t e = ptr[row] + lane; e < ptr[row + 1]; e += 32) {
            const int col = idx[e];
            const int cb = col - sp.lo;
            const int pos = pre[cb >> 5] + __popc(bm[cb >> 5] & ((1u << (cb & 31)) - 1u));
            const int64_t t = t0 + (pos >> 2);
            const int q = pos & 3;
            cols[t * 4 + q] = col;                                      // every row holding this column writes the same value
            atomicAdd(&vals[t * 32 + r * 4 + q], val[e]);               // repeated columns of a row add up
        }
    }
    if (lane == 0)
        for (int pos = u; pos & 3; ++pos) cols[(t0 + (pos >> 2)) * 4 + (pos & 3)] = sp.hi;      // padding: a valid row of X, zero block column
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// Y (n x k) = G X.  Warp per row group, 64 columns per pass; FULL: k is a multiple of 64 (no column predicates).
// VEC2 (X 16-byte aligned with an even leading dimension): a lane fetches TWO adjacent columns of its B row with one
// 16-byte load and feeds them to two accumulator tiles -- tile 2 jj takes the even columns of the 16-column span jj,
// tile 2 jj + 1 the odd ones -- so a k-tile costs four full-cache-line requests instead of eight half-line ones, and
// the lane ends up with four adjacent output columns (4 lk .. 4 lk + 3 of the span) per tile pair.
template <bool FULL, bool VEC2>
__global__ void __launch_bounds__(256, 2) b84_spmm_kernel(int64_t n, int64_t nrg, const int64_t* __restrict__ tptr,
                                                          const int32_t* __restrict__ cols, const double* __restrict__ vals, int k,
                                                          const double* __restrict__ X, int64_t ldx, double* __restrict__ Y, int64_t ldy) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lr = lane >> 2, lk = lane & 3;
    // consecutive row groups per CTA.  (Dealing out spatially STACKED groups to the warps of a CTA, so that they share
    // their rows of X through L1, was measured: 6 % slower -- the kernel is bound by the L1 data pipe, not by L2.)
    const int64_t g = (int64_t)blockIdx.x * 8 + warp;
    if (g >= nrg) return;
    const int64_t t0 = tptr[g], t1 = tptr[g + 1];
    const int64_t row = g * 8 + lr;
    const int nslab = (k + 63) >> 6;
    for (int slab = 0; slab < nslab; ++slab) {
        double acc[8][2];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = 0.0;
        if (VEC2) {
            // Loads are issued in a COALESCED lane mapping -- lane L reads the 16-byte piece (L & 7) of B row (L >> 3),
            // so an LDG.128 of the warp covers four full 128-byte lines -- and a shuffle then hands every lane the piece
            // the DMMA fragment layout wants (B row lk, columns 2 lr, 2 lr + 1: held by lane 8 lk + lr).  Loading in
            // the fragment layout directly (adjacent lanes on different rows) costs 16 L1 wavefronts per request
            // instead of 4 and made the L1 data pipe the limit (95 % busy in ncu).
            const int lrow = lane >> 3, lpc = lane & 7;
            const int c0 = slab * 64 + 2 * lpc;            // first of the two columns this lane LOADS in span 0 (span jj: + 16 jj)
            const int src = 8 * lk + lr;
            auto load_b = [&](const double* xr, double2* b) {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                    b[jj] = (FULL || c0 + 16 * jj + 1 < k) ? *reinterpret_cast<const double2*>(xr + 16 * jj)
                                                           : make_double2((c0 + 16 * jj < k) ? xr[16 * jj] : 0.0, 0.0);
            };
            // software pipeline: while the DMMAs of k-tile t issue, the B pieces of t + 1 and the block entry and
            // column index of t + 2 are in flight (the gather is two dependent loads deep: index -> row of X)
            double a_cur = 0.0, a_nxt = 0.0;
            int c_nxt = 0;
            double2 b_cur[4], b_nxt[4];
            if (t0 < t1) {
                a_cur = vals[t0 * 32 + lane];
                load_b(X + (int64_t)cols[t0 * 4 + lrow] * ldx + c0, b_cur);
            }
            if (t0 + 1 < t1) { a_nxt = vals[(t0 + 1) * 32 + lane]; c_nxt = cols[(t0 + 1) * 4 + lrow]; }
            for (int64_t t = t0; t < t1; ++t) {
                double a_n2 = 0.0;
                int c_n2 = 0;
                if (t + 1 < t1) load_b(X + (int64_t)c_nxt * ldx + c0, b_nxt);
                if (t + 2 < t1) { a_n2 = vals[(t + 2) * 32 + lane]; c_n2 = cols[(t + 2) * 4 + lrow]; }
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const double bx = __shfl_sync(0xffffffffu, b_cur[jj].x, src);
                    const double by = __shfl_sync(0xffffffffu, b_cur[jj].y, src);
                    dmma884(acc[2 * jj][0], acc[2 * jj][1], a_cur, bx);
                    dmma884(acc[2 * jj + 1][0], acc[2 * jj + 1][1], a_cur, by);
                }
                a_cur = a_nxt; a_nxt = a_n2; c_nxt = c_n2;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) b_cur[jj] = b_nxt[jj];
            }
            if (row < n) {
                double* yr = Y + row * ldy + slab * 64 + 4 * lk;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int col = slab * 64 + 16 * jj + 4 * lk;
                    const double v[4] = {acc[2 * jj][0], acc[2 * jj + 1][0], acc[2 * jj][1], acc[2 * jj + 1][1]};
                    if (FULL || col + 3 < k) {
                        *reinterpret_cast<double2*>(yr + 16 * jj) = make_double2(v[0], v[1]);
                        *reinterpret_cast<double2*>(yr + 16 * jj + 2) = make_double2(v[2], v[3]);
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (col + q < k) yr[16 * jj + q] = v[q];
                    }
                }
            }
        } else {
            const int c0 = slab * 64 + lr;                 // this lane's B column inside tile 0 (tile j: + 8 j)
#pragma unroll 2
            for (int64_t t = t0; t < t1; ++t) {
                const double a = vals[t * 32 + lane];
                const double* xr = X + (int64_t)cols[t * 4 + lk] * ldx + c0;
                double b[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) b[j] = (FULL || c0 + 8 * j < k) ? xr[8 * j] : 0.0;
#pragma unroll
                for (int j = 0; j < 8; ++j) dmma884(acc[j][0], acc[j][1], a, b[j]);
            }
            if (row < n) {
                double* yr = Y + row * ldy + slab * 64 + 2 * lk;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int col = slab * 64 + 8 * j + 2 * lk;
                    if (FULL || col + 1 < k) *reinterpret_cast<double2*>(yr + 8 * j) = make_double2(acc[j][0], acc[j][1]);
                    else if (col < k) yr[8 * j] = acc[j][0];
                }
            }
        }
    }
}

}  // namespace

void b84_free(B84Matrix* m) {
    if (!m) return;
    sfem_dev_free(m->tptr);
    sfem_dev_free(m->cols);
    sfem_dev_free(m->vals);
    delete m;
}

// Build the 8x4-block form of a CSR matrix with ascending column indices in every row (device arrays).
int b84_build(sfem_ctx* c, int64_t n, const int64_t* ptr, const int32_t* idx, const double* val, B84Matrix** out) {
    *out = nullptr;
    if (n <= 0 || !ptr || !idx || !val) return sfem_fail(c, SFEM_ERR_ARG, "b84_build: bad arguments");
    B84Matrix* m = new B84Matrix();
    m->n = n;
    m->nrg = (n + 7) / 8;
    int64_t *ntile = nullptr, *sums = nullptr;
    int rc = SFEM_OK;
    auto fail = [&](int code) { sfem_dev_free(ntile); sfem_dev_free(sums); b84_free(m); return code; };
    if (sfem_dev_malloc(&ntile, sizeof(int64_t) * (size_t)(m->nrg + 1)) != cudaSuccess ||
        sfem_dev_malloc(&sums, sizeof(int64_t) * (size_t)(m->nrg / 1024 + 8)) != cudaSuccess ||
        sfem_dev_malloc(&m->tptr, sizeof(int64_t) * (size_t)(m->nrg + 1)) != cudaSuccess)
        return fail(sfem_fail(c, SFEM_ERR_CUDA, "b84_build: out of memory"));
    {
        ProfScope ps(c, TAG_SETUP, 0.0, 0.0);
        b84_count_kernel<<<(unsigned)cdiv64(m->nrg, B84_WARPS), 32 * B84_WARPS, 0, c->stream>>>(n, m->nrg, ptr, idx, ntile);
        c->launches++;
        rc = exclusive_scan<int64_t>(c, ntile, m->nrg, m->tptr, sums);
    }
    if (rc) return fail(rc);
    int64_t h[2] = {0, 0};
    cudaMemcpyAsync(&h[0], m->tptr + m->nrg, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream);
    cudaMemcpyAsync(&h[1], ptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess)
        return fail(sfem_fail(c, SFEM_ERR_CUDA, "b84_build: count pass failed: %s", cudaGetErrorString(cudaGetLastError())));
    m->ntiles = h[0];
    m->nnz = h[1];
    const size_t nt = (size_t)(m->ntiles > 0 ? m->ntiles : 1);
    if (sfem_dev_malloc(&m->cols, sizeof(int32_t) * 4 * nt) != cudaSuccess || sfem_dev_malloc(&m->vals, sizeof(double) * 32 * nt) != cudaSuccess) {
        cudaGetLastError();
        return fail(sfem_fail(c, SFEM_ERR_CUDA, "b84_build: out of memory (%lld k-tiles)", (long long)m->ntiles));
    }
    cudaMemsetAsync(m->cols, 0, sizeof(int32_t) * 4 * nt, c->stream);
    cudaMemsetAsync(m->vals, 0, sizeof(double) * 32 * nt, c->stream);
    {
        ProfScope ps(c, TAG_SETUP, 0.0, 0.0);
        b84_fill_kernel<<<(unsigned)cdiv64(m->nrg, B84_WARPS), 32 * B84_WARPS, 0, c->stream>>>(n, m->nrg, ptr, idx, val, m->tptr, m->cols, m->vals);
        c->launches++;
    }
    cudaError_t e = cudaGetLastError();
    sfem_dev_free(ntile);
    sfem_dev_free(sums);
    if (e != cudaSuccess) { b84_free(m); return sfem_fail(c, SFEM_ERR_CUDA, "b84_build: fill launch -> %s", cudaGetErrorString(e)); }
    *out = m;
    return SFEM_OK;
}

int64_t b84_num_tiles(const B84Matrix* m) { return m ? m->ntiles : 0; }

// Y (n x k, ldy) = G X (n x k, ldx).  X and Y must not alias.
int b84_spmm(sfem_ctx* c, const B84Matrix* m, int k, const double* X, int64_t ldx, double* Y, int64_t ldy) {
    if (!m) return sfem_fail(c, SFEM_ERR_ARG, "b84_spmm: no matrix");
    if (k <= 0) return SFEM_OK;
    if ((ldy % 2) || (reinterpret_cast<uintptr_t>(Y) % 16)) return sfem_fail(c, SFEM_ERR_ARG, "b84_spmm: output block must be 16-byte aligned with an even leading dimension");
    // algorithmic work: the matrix entries once per 64-column pass, each vector block once; flops on the nonzeros
    ProfScope ps(c, TAG_SPMM_OTHER, 16.0 * (double)m->n * k + 12.0 * (double)m->nnz * ((k + 63) / 64) + 8.0 * (double)(m->n + 1),
                 2.0 * (double)m->nnz * k);
    const unsigned grid = (unsigned)cdiv64(m->nrg, 8);
    const bool v2 = (ldx % 2 == 0) && (reinterpret_cast<uintptr_t>(X) % 16 == 0);
#define B84_LAUNCH(F, V) b84_spmm_kernel<F, V><<<grid, 256, 0, c->stream>>>(m->n, m->nrg, m->tptr, m->cols, m->vals, k, X, ldx, Y, ldy)
    if (k % 64 == 0) { if (v2) B84_LAUNCH(true, true); else B84_LAUNCH(true, false); }
    else { if (v2) B84_LAUNCH(false, true); else B84_LAUNCH(false, false); }
#undef B84_LAUNCH
    KCHECK(c);
    return SFEM_OK;
}
