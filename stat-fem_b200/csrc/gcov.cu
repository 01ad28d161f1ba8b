// K6-K8: forcing covariance assembly  G = CSR{ (i,j) : row[j]/row[i] > cutoff },  row = b_i * b * sqexp(x_i, X)
// with the nugget added to row[i] before the test (ForcingCovariance.py:159-167).
//
// The reference evaluates all n^2 pairs in a Python loop.  Here a cell list (cell edge >= the largest distance at
// which the ratio test can still pass) restricts every row to its 3^d neighbouring cells; one warp owns a row,
// lanes stride over the candidates.  Two passes (count -> exclusive scan -> fill) so the CSR arrays are written
// exactly once; kept entries are compacted into shared memory, bitonic-sorted by column there and written
// coalesced, so HBM sees 12 bytes per stored entry.
//
// Bit-exact pattern: the decision expression is evaluated in the reference's operation order with explicitly
// rounded intrinsics (no FMA contraction):  r = sqrt(sum_k dx_k^2) [cdist],  r2 = r*r,
// c = e2s * exp((-0.5*r2) * em2l),  row_j = (b_i*b_j) * c,  row_i = (b_i*b_i)*e2s + reg,  keep <=> row_j/row_i > cutoff.
// e2s = exp(2 sigma) and em2l = exp(-2 l) are evaluated by the caller (NumPy) so they match the reference bitwise.
// CUDA's exp() may differ from NumPy's by 1 ulp, so rows containing a ratio within 1e-14 (relative) of the cutoff
// are reported back (`borderline_rows`) for the host to re-decide with the reference's own expression.
#include "common.cuh"
#include <climits>

int build_cell_list(sfem_ctx* c, const int* cell_of, int64_t n, int ncell, int** cell_start_out, int** cell_nodes_out);
template <typename T>
int exclusive_scan(sfem_ctx* c, const T* in, int64_t n, T* out, T* sums);

namespace {

constexpr int GC_WARPS = 4;
constexpr int GC_CAP = 512;          // entries per row sorted in shared memory by the owning warp
constexpr int GC_LONG_CAP = 12288;   // entries per row sorted by a whole CTA (dynamic shared memory)
constexpr int GC_BM_MAXBITS = 65536; // widest column span of a row ranked through a per-warp bitmap (8 KB + 4 KB of prefix counts)

struct GcP {
    int64_t n;
    int d;
    const double* coords;
    const double* b;
    double e2s, em2l, cutoff, reg, bmax, two_e2l;
    double lo0, lo1, lo2, inv_edge;
    int nc0, nc1, nc2;
    const int* cell_start;
    const int* cell_nodes;
};

__device__ __forceinline__ int cell_coord(double x, double lo, double inv_edge, int nc) {
    int v = (int)floor((x - lo) * inv_edge);
    return v < 0 ? 0 : (v >= nc ? nc - 1 : v);
}

__global__ void cell_of_kernel(GcP p, int* __restrict__ cell_of) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    int c0 = cell_coord(p.coords[i * p.d], p.lo0, p.inv_edge, p.nc0);
    int c1 = p.d > 1 ? cell_coord(p.coords[i * p.d + 1], p.lo1, p.inv_edge, p.nc1) : 0;
    int c2 = p.d > 2 ? cell_coord(p.coords[i * p.d + 2], p.lo2, p.inv_edge, p.nc2) : 0;
    cell_of[i] = c0 + p.nc0 * (c1 + p.nc1 * c2);
}

// The reference's per-pair decision.  Returns the stored value (row_j, nugget included on the diagonal), sets keep
// and borderline.
struct RowCtx {
    double xi0, xi1, xi2, bi, diag, R2;
    int64_t i;
};

__device__ __forceinline__ RowCtx make_row(const GcP& p, int64_t i) {
    RowCtx r;
    r.i = i;
    r.xi0 = p.coords[i * p.d];
    r.xi1 = p.d > 1 ? p.coords[i * p.d + 1] : 0.0;
    r.xi2 = p.d > 2 ? p.coords[i * p.d + 2] : 0.0;
    r.bi = p.b[i];
    r.diag = __dadd_rn(__dmul_rn(__dmul_rn(r.bi, r.bi), p.e2s), p.reg);
    if (p.cutoff > 0.0) {
        double arg = r.bi * p.bmax * p.e2s / (p.cutoff * r.diag);
        r.R2 = arg > 1.0 ? p.two_e2l * log(arg) * (1.0 + 1.0e-6) + 1.0e-300 : -1.0;
    } else {
        r.R2 = INFINITY;
    }
    return r;
}

__device__ __forceinline__ double eval_pair(const GcP& p, const RowCtx& r, int j, bool& keep, bool& borderline) {
    borderline = false;
    if (j == r.i) {
        keep = __ddiv_rn(r.diag, r.diag) > p.cutoff;
        return r.diag;
    }
    double dx = __dsub_rn(r.xi0, p.coords[(int64_t)j * p.d]);
    double s = __dmul_rn(dx, dx);
    if (p.d > 1) {
        double dy = __dsub_rn(r.xi1, p.coords[(int64_t)j * p.d + 1]);
        s = __dadd_rn(s, __dmul_rn(dy, dy));
    }
    if (p.d > 2) {
        double dz = __dsub_rn(r.xi2, p.coords[(int64_t)j * p.d + 2]);
        s = __dadd_rn(s, __dmul_rn(dz, dz));
    }
    if (!(s <= r.R2)) {   // provably below the cutoff (conservative radius): skip the exp
        keep = false;
        return 0.0;
    }
    double rr = __dsqrt_rn(s);
    double r2 = __dmul_rn(rr, rr);
    double t = __dmul_rn(__dmul_rn(-0.5, r2), p.em2l);
    double cj = __dmul_rn(p.e2s, exp(t));
    double rowj = __dmul_rn(__dmul_rn(r.bi, p.b[j]), cj);
    double ratio = __ddiv_rn(rowj, r.diag);
    keep = ratio > p.cutoff;
    if (p.cutoff > 0.0)
        borderline = fabs(ratio - p.cutoff) <= 1.0e-14 * p.cutoff;
    else
        borderline = (t < -700.0 && t > -760.0);   // exp underflow zone decides ratio > 0
    return rowj;
}

// iterate the 3^d neighbouring cells of row i; F(j, lane_active) is called warp-synchronously
template <typename F>
__device__ __forceinline__ void for_candidates(const GcP& p, const RowCtx& r, int lane, F f) {
    int c0 = cell_coord(r.xi0, p.lo0, p.inv_edge, p.nc0);
    int c1 = p.d > 1 ? cell_coord(r.xi1, p.lo1, p.inv_edge, p.nc1) : 0;
    int c2 = p.d > 2 ? cell_coord(r.xi2, p.lo2, p.inv_edge, p.nc2) : 0;
    int z0 = max(c2 - 1, 0), z1 = min(c2 + 1, p.nc2 - 1);
    int y0 = max(c1 - 1, 0), y1 = min(c1 + 1, p.nc1 - 1);
    int x0 = max(c0 - 1, 0), x1 = min(c0 + 1, p.nc0 - 1);
    for (int z = z0; z <= z1; ++z)
        for (int y = y0; y <= y1; ++y) {
            // cells x0..x1 of one (y,z) line are contiguous in the cell list
            int cb = x0 + p.nc0 * (y + p.nc1 * z);
            int s = p.cell_start[cb], e = p.cell_start[cb + (x1 - x0) + 1];
            for (int q = s; q < e; q += 32) {
                int t = q + lane;
                bool act = t < e;
                int j = act ? p.cell_nodes[t] : -1;
                f(j, act);
            }
        }
}

__global__ void __launch_bounds__(GC_WARPS * 32) gcov_count_kernel(GcP p, int64_t* __restrict__ row_nnz,
                                                                  int* __restrict__ max_nnz,
                                                                  int* __restrict__ n_border,
                                                                  int* __restrict__ border_rows, int border_cap) {
    int lane = threadIdx.x & 31;
    int64_t i = (int64_t)blockIdx.x * GC_WARPS + (threadIdx.x >> 5);
    if (i >= p.n) return;
    RowCtx r = make_row(p, i);
    int cnt = 0;
    bool anyb = false;
    int jlo = INT_MAX, jhi = -1;
    for_candidates(p, r, lane, [&](int j, bool act) {
        bool keep = false, bl = false;
        if (act) eval_pair(p, r, j, keep, bl);
        if (keep) { jlo = min(jlo, j); jhi = max(jhi, j); }
        unsigned m = __ballot_sync(0xffffffffu, keep);
        cnt += __popc(m);
        anyb |= (__ballot_sync(0xffffffffu, bl) != 0u);
    });
    for (int o = 16; o > 0; o >>= 1) {
        jlo = min(jlo, __shfl_xor_sync(0xffffffffu, jlo, o));
        jhi = max(jhi, __shfl_xor_sync(0xffffffffu, jhi, o));
    }
    if (lane == 0) {
        row_nnz[i] = cnt;
        atomicMax(max_nnz, cnt);
        if (cnt > 0) atomicMax(max_nnz + 2, jhi - jlo + 1);      // counters[2]: widest column span of a row
        if (anyb) {
            int pos = atomicAdd(n_border, 1);
            if (pos < border_cap) border_rows[pos] = (int)i;
        }
    }
}

__device__ __forceinline__ void warp_bitonic(int* col, double* val, int S, int lane) {
    for (int k = 2; k <= S; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < S; t += 32) {
                int u = t ^ j;
                if (u > t) {
                    bool up = (t & k) == 0;
                    int ct = col[t], cu = col[u];
                    if ((ct > cu) == up) {
                        col[t] = cu; col[u] = ct;
                        double vt = val[t]; val[t] = val[u]; val[u] = vt;
                    }
                }
            }
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(GC_WARPS * 32) gcov_fill_kernel(GcP p, const int64_t* __restrict__ indptr,
                                                                 int32_t* __restrict__ indices,
                                                                 double* __restrict__ vals, int presorted, int bm_words) {
    __shared__ int scol[GC_WARPS][GC_CAP];
    __shared__ double sval[GC_WARPS][GC_CAP];
    extern __shared__ __align__(16) unsigned char gc_dyn[];      // bm_words > 0: per warp a bitmap of the row's column span + prefix counts
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t i = (int64_t)blockIdx.x * GC_WARPS + warp;
    if (i >= p.n) return;
    RowCtx r = make_row(p, i);
    int64_t base = indptr[i];
    int nnz = (int)(indptr[i + 1] - base);
    bool in_smem = nnz <= GC_CAP && !presorted;
    int filled = 0;
    unsigned lt = (1u << lane) - 1u;
    int jlo = INT_MAX, jhi = -1;
    for_candidates(p, r, lane, [&](int j, bool act) {
        bool keep = false, bl = false;
        double v = 0.0;
        if (act) v = eval_pair(p, r, j, keep, bl);
        unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            jlo = min(jlo, j); jhi = max(jhi, j);
            int pos = filled + __popc(m & lt);
            if (in_smem) {
                scol[warp][pos] = j;
                sval[warp][pos] = v;
            } else {
                indices[base + pos] = j;
                vals[base + pos] = v;
            }
        }
        filled += __popc(m);
    });
    if (in_smem && bm_words > 0 && nnz > 0) {
        // Rank by bitmap instead of sorting: one bit per column of the row's span [jlo, jhi] (the count pass measured the
        // widest span, the host sized the bitmap), a prefix count per 32-bit word, and every entry goes straight to
        // position  prefix[word] + popc(bits below)  of the row.  O(entries + span / 32) instead of O(entries log^2).
        for (int o = 16; o > 0; o >>= 1) {
            jlo = min(jlo, __shfl_xor_sync(0xffffffffu, jlo, o));
            jhi = max(jhi, __shfl_xor_sync(0xffffffffu, jhi, o));
        }
        const int stride = (bm_words * 6 + 15) & ~15;
        unsigned* bm = reinterpret_cast<unsigned*>(gc_dyn + (size_t)warp * stride);
        unsigned short* pre = reinterpret_cast<unsigned short*>(bm + bm_words);
        const int nw = ((jhi - jlo) >> 5) + 1;
        for (int w = lane; w < nw; w += 32) bm[w] = 0u;
        __syncwarp();
        for (int t = lane; t < nnz; t += 32) {
            const int o = scol[warp][t] - jlo;
            atomicOr(&bm[o >> 5], 1u << (o & 31));
        }
        __syncwarp();
        int carry = 0;
        for (int w0 = 0; w0 < nw; w0 += 32) {
            const int w = w0 + lane;
            const int cw = w < nw ? __popc(bm[w]) : 0;
            int incl = cw;
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            if (w < nw) pre[w] = (unsigned short)(carry + incl - cw);
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        __syncwarp();
        for (int t = lane; t < nnz; t += 32) {
            const int j = scol[warp][t], o = j - jlo;
            const int pos = (int)pre[o >> 5] + __popc(bm[o >> 5] & ((1u << (o & 31)) - 1u));
            indices[base + pos] = j;
            vals[base + pos] = sval[warp][t];
        }
    } else if (in_smem) {
        int S = 2;
        while (S < nnz) S <<= 1;
        for (int t = nnz + lane; t < S; t += 32) {
            scol[warp][t] = INT_MAX;
            sval[warp][t] = 0.0;
        }
        __syncwarp();
        warp_bitonic(scol[warp], sval[warp], S, lane);
        for (int t = lane; t < nnz; t += 32) {
            indices[base + t] = scol[warp][t];
            vals[base + t] = sval[warp][t];
        }
    }
}

// rows longer than GC_CAP (written unsorted by the fill kernel): one CTA sorts one row in dynamic shared memory
__global__ void __launch_bounds__(256) gcov_sort_long_kernel(int64_t n, const int64_t* __restrict__ indptr,
                                                             int32_t* __restrict__ indices, double* __restrict__ vals) {
    extern __shared__ __align__(16) unsigned char dyn[];
    double* sval = reinterpret_cast<double*>(dyn);
    int* scol = reinterpret_cast<int*>(dyn + sizeof(double) * 16384);
    for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
        int64_t base = indptr[i];
        int nnz = (int)(indptr[i + 1] - base);
        if (nnz <= GC_CAP || nnz > GC_LONG_CAP) continue;
        int S = 2;
        while (S < nnz) S <<= 1;
        for (int t = threadIdx.x; t < S; t += blockDim.x) {
            scol[t] = t < nnz ? indices[base + t] : INT_MAX;
            sval[t] = t < nnz ? vals[base + t] : 0.0;
        }
        __syncthreads();
        for (int k = 2; k <= S; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = threadIdx.x; t < S; t += blockDim.x) {
                    int u = t ^ j;
                    if (u > t) {
                        bool up = (t & k) == 0;
                        int ct = scol[t], cu = scol[u];
                        if ((ct > cu) == up) {
                            scol[t] = cu; scol[u] = ct;
                            double vt = sval[t]; sval[t] = sval[u]; sval[u] = vt;
                        }
                    }
                }
                __syncthreads();
            }
        for (int t = threadIdx.x; t < nnz; t += blockDim.x) {
            indices[base + t] = scol[t];
            vals[base + t] = sval[t];
        }
        __syncthreads();
    }
}

GcP make_params(const sfem_ctx* c) {
    GcP p;
    p.n = c->gc.n; p.d = c->gc.d; p.coords = c->gc.coords; p.b = c->gc.b;
    p.e2s = c->gc.e2s; p.em2l = c->gc.em2l; p.cutoff = c->gc.cutoff; p.reg = c->gc.reg;
    p.bmax = c->gc.bmax; p.two_e2l = c->gc.two_e2l;
    p.lo0 = c->gc.lo[0]; p.lo1 = c->gc.lo[1]; p.lo2 = c->gc.lo[2]; p.inv_edge = c->gc.inv_edge;
    p.nc0 = c->gc.nc[0]; p.nc1 = c->gc.nc[1]; p.nc2 = c->gc.nc[2];
    p.cell_start = c->gc.cell_start; p.cell_nodes = c->gc.cell_nodes;
    return p;
}

}  // namespace

void gcov_free(sfem_ctx* c) {
    if (c->gc.cell_start) sfem_dev_free(c->gc.cell_start);
    if (c->gc.cell_nodes) sfem_dev_free(c->gc.cell_nodes);
    c->gc.cell_start = nullptr;
    c->gc.cell_nodes = nullptr;
}

int gcov_count(sfem_ctx* c, int64_t n, int d, const double* coords, const double* int_basis, double exp_2sigma,
               double exp_m2l, double cutoff, double reg, double bmax, double cell_edge, const double* lo,
               const double* hi, int64_t* indptr_out, int64_t* nnz_out, int64_t* max_row_nnz_out,
               int32_t* borderline_rows, int borderline_cap, int64_t* n_borderline_out) {
    if (d < 1 || d > 3) return sfem_fail(c, SFEM_ERR_ARG, "gcov_count: dimension must be 1, 2 or 3");
    if (n <= 0 || n > INT_MAX) return sfem_fail(c, SFEM_ERR_ARG, "gcov_count: n out of range");
    gcov_free(c);
    auto& g = c->gc;
    g.n = n; g.d = d; g.coords = coords; g.b = int_basis;
    g.e2s = exp_2sigma; g.em2l = exp_m2l; g.cutoff = cutoff; g.reg = reg; g.bmax = bmax;
    g.two_e2l = 2.0 / exp_m2l;
    int64_t ncell = 1;
    for (int k = 0; k < 3; ++k) {
        g.lo[k] = k < d ? lo[k] : 0.0;
        g.nc[k] = 1;
    }
    if (cell_edge > 0.0 && cutoff > 0.0) {
        g.inv_edge = 1.0 / cell_edge;
        for (int k = 0; k < d; ++k) {
            double ext = (hi[k] - lo[k]) * g.inv_edge;
            int64_t m = (int64_t)floor(ext) + 1;
            if (m < 1) m = 1;
            if (m > 4096) m = 4096;
            g.nc[k] = (int)m;
            ncell *= m;
        }
        if (ncell > (int64_t)1 << 28) return sfem_fail(c, SFEM_ERR_LIMIT, "gcov_count: too many cells");
    } else {
        g.inv_edge = 0.0;   // every coordinate lands in cell 0: all-pairs mode
    }
    g.ncell = (int)ncell;
    // if clamping to 4096 cells per side made cells smaller than the radius the 3^d search would be wrong
    if (cell_edge > 0.0 && cutoff > 0.0)
        for (int k = 0; k < d; ++k)
            if ((hi[k] - lo[k]) * g.inv_edge >= 4096.0)
                return sfem_fail(c, SFEM_ERR_LIMIT, "gcov_count: cutoff radius too small relative to the domain");

    size_t nsum = (size_t)cdiv64(n, 1024) + 2;
    // scratch: cell_of[n] ints | row_nnz[n] int64 | sums | counters
    size_t bytes = sizeof(int) * (size_t)n + 16 + sizeof(int64_t) * ((size_t)n + nsum) + 64;
    unsigned char* s = (unsigned char*)sfem_scratch(c, bytes);
    if (!s) return sfem_fail(c, SFEM_ERR_CUDA, "gcov_count: scratch allocation failed");
    // the cell-list builder also uses the scratch arena, so cell_of lives in its own allocation
    int* cell_of = nullptr;
    CUDA_TRY(c, sfem_dev_malloc(&cell_of, sizeof(int) * (size_t)n));
    GcP p = make_params(c);
    cell_of_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, c->stream>>>(p, cell_of);
    KCHECK(c);
    int rc = build_cell_list(c, cell_of, n, g.ncell, &g.cell_start, &g.cell_nodes);
    if (rc) { sfem_dev_free(cell_of); return rc; }
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    sfem_dev_free(cell_of);

    s = (unsigned char*)sfem_scratch(c, bytes);
    if (!s) return sfem_fail(c, SFEM_ERR_CUDA, "gcov_count: scratch allocation failed");
    int64_t* row_nnz = reinterpret_cast<int64_t*>(s);
    int64_t* sums = row_nnz + n;
    int* counters = reinterpret_cast<int*>(sums + nsum);
    CUDA_TRY(c, cudaMemsetAsync(counters, 0, 3 * sizeof(int), c->stream));
    p = make_params(c);
    {
        ProfScope ps(c, TAG_GCOV_COUNT, 8.0 * (double)n * (d + 2), 0.0);
        gcov_count_kernel<<<(unsigned)cdiv64(n, GC_WARPS), GC_WARPS * 32, 0, c->stream>>>(p, row_nnz, counters, counters + 1,
                                                                                       borderline_rows, borderline_cap);
        KCHECK(c);
    }
    rc = exclusive_scan<int64_t>(c, row_nnz, n, indptr_out, sums);
    if (rc) return rc;
    int hc[3];
    CUDA_TRY(c, cudaMemcpyAsync(hc, counters, sizeof(hc), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(nnz_out, indptr_out + n, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    g.max_row_nnz = hc[0];
    g.max_span = hc[2];
    g.nnz = *nnz_out;
    if (max_row_nnz_out) *max_row_nnz_out = hc[0];
    if (n_borderline_out) *n_borderline_out = hc[1];
    if (hc[0] > GC_LONG_CAP && g.ncell > 1)
        return sfem_fail(c, SFEM_ERR_LIMIT, "gcov_count: a row keeps %d entries (> %d) in cell-list mode", hc[0], GC_LONG_CAP);
    return SFEM_OK;
}

int gcov_fill(sfem_ctx* c, const int64_t* indptr, int32_t* indices_out, double* vals_out) {
    auto& g = c->gc;
    if (!g.cell_start) return sfem_fail(c, SFEM_ERR_STATE, "gcov_fill: call gcov_count first");
    GcP p = make_params(c);
    int presorted = (g.ncell == 1) ? 1 : 0;   // single cell: candidates are visited in ascending DOF order
    {
        // algorithmic bytes: 12 per stored entry + row pointers + coordinates and basis integrals
        ProfScope ps(c, TAG_GCOV_FILL, 12.0 * (double)g.nnz + 8.0 * (double)(g.n + 1) + 8.0 * (double)g.n * (g.d + 1), 0.0);
        // rows are ordered by bitmap rank when the widest column span fits a per-warp bitmap of at most 64 Ki bits
        // (structured and bandwidth-reduced numberings); otherwise by the bitonic sort in shared memory
        int bm_words = (!presorted && g.max_span > 0 && g.max_span <= GC_BM_MAXBITS && !getenv("SFEM_GCOV_BITONIC")) ? (g.max_span + 31) / 32 : 0;
        size_t dyn = bm_words ? (size_t)GC_WARPS * (((size_t)bm_words * 6 + 15) & ~(size_t)15) : 0;
        if (dyn > 48 * 1024 - sizeof(int) * GC_WARPS * GC_CAP - sizeof(double) * GC_WARPS * GC_CAP)
            cudaFuncSetAttribute(gcov_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        gcov_fill_kernel<<<(unsigned)cdiv64(g.n, GC_WARPS), GC_WARPS * 32, dyn, c->stream>>>(p, indptr, indices_out, vals_out, presorted, bm_words);
        KCHECK(c);
    }
    if (!presorted && g.max_row_nnz > GC_CAP) {
        size_t dyn = (sizeof(double) + sizeof(int)) * 16384;
        cudaFuncSetAttribute(gcov_sort_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        gcov_sort_long_kernel<<<SFEM_NUM_SMS, 256, dyn, c->stream>>>(g.n, indptr, indices_out, vals_out);
        KCHECK(c);
    }
    gcov_free(c);
    return SFEM_OK;
}
