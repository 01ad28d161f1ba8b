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

// ------------------------------------------------------------------------------------------------ DMMA form
// One thread per chunk.  For each group of 8 consecutive rows the distinct staged slots with a nonzero entry are
// binned by (slot mod 4); position q of a k-tile takes the next slot of bin q, so that the four rows of a B fragment
// sit in different 32-byte bank groups of the (544-byte pitched) stage buffer.  When a bin has run out the position
// takes a slot of the fullest remaining bin instead -- such a tile pays a 2-way bank conflict, which is cheaper than
// the extra k-tile (4 DMMAs per warp) a strict assignment would cost: a row group needs ceil(slots / 4) k-tiles.
constexpr int TD_BIN = 48;        // slots of one residue a row group may touch

struct TdBins {
    unsigned short slot[4][TD_BIN];
    int cnt[4];
};

// bins of row group [row0, row1) of a chunk with ELL width W; returns false on overflow
__device__ bool td_bins(const unsigned short* __restrict__ ec, const double* __restrict__ pat, int W, int row0, int row1, TdBins& b) {
    b.cnt[0] = b.cnt[1] = b.cnt[2] = b.cnt[3] = 0;
    for (int r = row0; r < row1; ++r)
        for (int w = 0; w < W; ++w) {
            if (pat[r * W + w] == 0.0) continue;
            const unsigned short s = ec[r * W + w];
            const int q = s & 3;
            bool found = false;
            for (int i = 0; i < b.cnt[q]; ++i) found |= (b.slot[q][i] == s);
            if (found) continue;
            if (b.cnt[q] == TD_BIN) return false;
            b.slot[q][b.cnt[q]++] = s;
        }
    return true;
}

constexpr int TD_MAXKT = 16;      // k-tiles per row group the builder can form (the kernels keep 8 or 16 of them in registers)
constexpr int TD_MAXNEED = 160;   // staged rows per chunk of the tensor-core form (64 for the wide-slab kernel)

// k-tiles of one row group from its bins (see above); live[t] bit q: position q of tile t carries a slot of the group
__device__ int td_form(const TdBins& b, int nneed, unsigned short (*sl)[4], unsigned char* live) {
    const int KT = (b.cnt[0] + b.cnt[1] + b.cnt[2] + b.cnt[3] + 3) >> 2;
    int used[4] = {0, 0, 0, 0};
    unsigned short any = 0;
    for (int q = 0; q < 4; ++q)
        if (b.cnt[q]) any = b.slot[q][0];
    for (int t = 0; t < KT && t < TD_MAXKT; ++t) {
        unsigned lv = 0;
        for (int q = 0; q < 4; ++q)
            if (used[q] < b.cnt[q]) { sl[t][q] = b.slot[q][used[q]++]; lv |= 1u << q; }
        for (int q = 0; q < 4; ++q) {
            if (lv & (1u << q)) continue;
            int best = -1, rem = 0;
            for (int p2 = 0; p2 < 4; ++p2)
                if (b.cnt[p2] - used[p2] > rem) { rem = b.cnt[p2] - used[p2]; best = p2; }
            if (best >= 0) { sl[t][q] = b.slot[best][used[best]++]; lv |= 1u << q; }
            else sl[t][q] = (unsigned short)(q < nneed ? q : any);      // padding: a staged slot, all-zero block column
        }
        live[t] = (unsigned char)lv;
    }
    return KT;
}

// One thread per (chunk, row group).  rgcnt[chunk][group] = (k-tiles, nonzero entries);
// stats[0] = max k-tiles per chunk, stats[1] = max nonzeros per chunk, stats[2] = not representable, stats[3] = max k-tiles per row group
__global__ void tile_dmma_count_kernel(int nchunks, int dl, int e8, const int* __restrict__ desc, const unsigned short* __restrict__ ecol,
                                       const double* __restrict__ eval, ushort2* __restrict__ rgcnt, int* __restrict__ stats) {
    const int tidg = blockIdx.x * blockDim.x + threadIdx.x;
    const int cidx = tidg >> 2, g = tidg & 3;
    int kt = 0, nz = 0;
    if (cidx < nchunks) {
        const int* dsc = desc + (size_t)cidx * dl;
        const int nr = dsc[1], nneed = dsc[2], W = dsc[3];
        const unsigned short* ec = ecol + (size_t)cidx * e8;
        const double* ev = eval + (size_t)cidx * e8;
        if (nr > 32 || nneed > TD_MAXNEED) stats[2] = 1;
        const int row0 = g * 8, row1 = min(row0 + 8, nr);
        if (row0 < row1 && nr <= 32 && nneed <= TD_MAXNEED) {
            TdBins b;
            if (!td_bins(ec, ev, W, row0, row1, b)) stats[2] = 1;
            else {
                kt = (b.cnt[0] + b.cnt[1] + b.cnt[2] + b.cnt[3] + 3) >> 2;
                if (kt > TD_MAXKT) stats[2] = 1;
                for (int t = row0 * W; t < row1 * W; ++t) nz += (ev[t] != 0.0);
            }
        }
        rgcnt[tidg] = make_ushort2((unsigned short)kt, (unsigned short)nz);
        atomicMax(&stats[3], kt);
    }
    // chunk totals over the four threads of a chunk (adjacent lanes)
    int tk = kt, tn = nz;
    tk += __shfl_xor_sync(0xffffffffu, tk, 1); tn += __shfl_xor_sync(0xffffffffu, tn, 1);
    tk += __shfl_xor_sync(0xffffffffu, tk, 2); tn += __shfl_xor_sync(0xffffffffu, tn, 2);
    if (cidx < nchunks && g == 0) {
        atomicMax(&stats[0], tk);
        atomicMax(&stats[1], tn);
    }
}

// write_meta: also fills the k-tile records and desc[6]; pat decides which entries exist, src supplies their values.
// apack must be zero on entry (entries are accumulated, so repeated columns of a row add up).
__global__ void tile_dmma_fill_kernel(int nchunks, int dl, int e8, int kts, int ap8, int write_meta, int* __restrict__ desc,
                                      const unsigned short* __restrict__ ecol, const double* __restrict__ pat,
                                      const double* __restrict__ src, const ushort2* __restrict__ rgcnt, uint4* __restrict__ ktmeta,
                                      double* __restrict__ apack) {
    const int tidg = blockIdx.x * blockDim.x + threadIdx.x;
    const int cidx = tidg >> 2, g = tidg & 3;
    if (cidx >= nchunks) return;
    int* dsc = desc + (size_t)cidx * dl;
    const int nr = dsc[1], nneed = dsc[2], W = dsc[3];
    const unsigned short* ec = ecol + (size_t)cidx * e8;
    const double* pt = pat + (size_t)cidx * e8;
    const double* sv = src + (size_t)cidx * e8;
    int tile0 = 0, base0 = 0;
    unsigned counts = 0;
    for (int g2 = 0; g2 < 4; ++g2) {
        const ushort2 cn = rgcnt[cidx * 4 + g2];
        counts |= (unsigned)cn.x << (8 * g2);
        if (g2 < g) { tile0 += cn.x; base0 += cn.y; }
    }
    if (write_meta && g == 0) dsc[6] = (int)counts;
    const int row0 = g * 8, row1 = min(row0 + 8, nr);
    if (row0 >= row1) return;
    TdBins b;
    td_bins(ec, pt, W, row0, row1, b);
    unsigned short sl[TD_MAXKT][4];
    unsigned char live[TD_MAXKT];
    const int KT = td_form(b, nneed, sl, live);
    unsigned char where[TD_MAXNEED];
    for (int i = 0; i < TD_MAXNEED; ++i) where[i] = 0xff;
    for (int t = 0; t < KT; ++t)
        for (int q = 0; q < 4; ++q)
            if (live[t] & (1u << q)) where[sl[t][q]] = (unsigned char)(t * 4 + q);
    unsigned mask[TD_MAXKT];
    for (int t = 0; t < KT; ++t) mask[t] = 0;
    for (int r = row0; r < row1; ++r)
        for (int w = 0; w < W; ++w)
            if (pt[r * W + w] != 0.0) {
                const int pos = where[ec[r * W + w]];
                mask[pos >> 2] |= 1u << ((r - row0) * 4 + (pos & 3));
            }
    int first[TD_MAXKT];
    int acc = base0;
    for (int t = 0; t < KT; ++t) { first[t] = acc; acc += __popc(mask[t]); }
    double* ap = apack + (size_t)cidx * ap8;
    for (int r = row0; r < row1; ++r)
        for (int w = 0; w < W; ++w)
            if (pt[r * W + w] != 0.0) {
                const int pos = where[ec[r * W + w]];
                const int t = pos >> 2, bit = (r - row0) * 4 + (pos & 3);
                ap[first[t] + __popc(mask[t] & ((1u << bit) - 1u))] += sv[r * W + w];
            }
    if (write_meta) {
        uint4* meta = ktmeta + (size_t)cidx * kts + tile0;
        for (int t = 0; t < KT; ++t)
            meta[t] = make_uint4((unsigned)sl[t][0] | ((unsigned)sl[t][1] << 16), (unsigned)sl[t][2] | ((unsigned)sl[t][3] << 16), mask[t], (unsigned)first[t]);
    }
}

// ------------------------------------------------------------------------------------------------ apply
struct TileDev {
    const int* desc; const int* need; const unsigned short* ecol; const double* eval;
    const uint4* ktmeta; const double* apack;
    int kts, ap8, nst;               // nst: stage buffers of the tensor-core kernel (2 or 3)
    unsigned off_dacc;
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
    extern __shared__ __align__(128) unsigned char sm[];
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

// ---- FP64 tensor-core variant -----------------------------------------------------------------------------------
// The chunk matrix is applied with DMMA m8n8k4.  Warp w owns row group (w & 3) -- 8 rows -- and the 32 columns of half
// (w >> 2) of a 64-column slab, i.e. four 8x8 accumulator tiles; a staged value is read from shared memory once per
// k-tile that contains its slot (for all 8 rows of the group) instead of once per matrix entry.
//
// Work items run chunk-major: all column slabs of a chunk one after the other.  At the first slab every lane loads its
// entries of the row group's 8x4 blocks (lane mask + popcount into the packed values) and the slot offsets into
// REGISTERS, where they stay for the chunk's remaining slabs; the chunk record itself is fetched once per chunk.
// Staging is done by the TMA engine: one `cp.async.bulk` per input row (512 bytes), dealt out over the warps and
// completing on the "full" mbarrier of the stage, 2 or 3 stages deep; a stage is refilled once all eight warps have
// arrived on its "empty" mbarrier, so the loop has no block-wide barrier.  Records travel the same way through a ring
// of four buffers.  Rows are pitched 544 bytes apart, so the four rows of a B fragment --
// slots with distinct (slot mod 4) by construction -- fall into different 32-byte bank groups without any swizzle.
// Dot products (TM_MULT_DOT / TM_JACOBI_DOT): reduced over the rows of a tile with warp shuffles into one register per
// column slab (DREG, at most DM_DREG slabs), else transposed through a per-warp scratch and accumulated per column in
// shared memory; one row of partials per CTA is written at the end, as in the scalar kernel.
// Two shapes of the kernel (SLABW = right-hand-side columns per work item):
//   64: rows pitched 544 bytes, at most 64 staged rows and 8 k-tiles per row group -- 2-D meshes aligned with the lattice;
//   32: rows pitched 288 bytes, at most 160 staged rows and 16 k-tiles per row group -- 3-D meshes (a 32-row chunk of a
//       15-point stencil touches ~120 rows), 2-D meshes that do not align with the lattice, and the 2-D restrictions
//       (an 8x4 tile of coarse nodes gathers 17x9 = 153 fine rows).  A warp then owns two
//       accumulator tiles instead of four, which frees the registers for the longer k-tile list.
template <int SLABW> struct DmCfg {
    static constexpr int PITCH = SLABW * 8 + 32;           // bytes between staged rows (pitch = 32 mod 128)
    static constexpr int KTR = SLABW == 64 ? 8 : 16;       // k-tiles of a row group held in registers
    static constexpr int MAXNEED = SLABW == 64 ? 64 : 160; // staged rows per chunk (a multiple of 8: MAXNEED / 8 lanes of every warp issue row copies)
    static constexpr int NG = SLABW / 16;                  // 8-column accumulator tiles per warp
    static constexpr int CW = SLABW / 2;                   // columns per warp
};
constexpr int DM_NREC = 4;        // chunk-record ring
constexpr int DM_HDR = 128;       // mbarriers: full[3] at 0, record[4] at 24, empty[3] at 56, chunk-done[4] at 80

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// global -> shared bulk copy (bytes: multiple of 16; both addresses 16-byte aligned), completing on an mbarrier
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
template <int OFF>
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];\n" : "=d"(v) : "r"(addr), "n"(OFF));
    return v;
}
__device__ __forceinline__ double2 lds_v2f64(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

// DREG (dot modes, at most DM_DREG column slabs): the per-column sums are reduced over the 8 rows of a tile with a
// shuffle reduce-scatter (7 exchanges: afterwards every lane owns ONE of the warp's 32 columns) and accumulated in
// registers, one per slab; no shared-memory scratch, so the dot modes keep the stage depth of the plain ones.
constexpr int DM_DREG = 6;
template <int MODE, bool DREG, int SLABW>
__global__ void __launch_bounds__(256, 2) tileop_dmma_kernel(TileDev op, TileArgs a) {
    constexpr int DM_PITCH = DmCfg<SLABW>::PITCH, DM_KTR = DmCfg<SLABW>::KTR, NG = DmCfg<SLABW>::NG, CW = DmCfg<SLABW>::CW;
    constexpr int LROWS = DmCfg<SLABW>::MAXNEED / 8;        // lanes of a warp that issue row copies
    extern __shared__ __align__(128) unsigned char sm[];
    constexpr bool HAS_B = (MODE == TM_JACOBI || MODE == TM_JACOBI_DOT || MODE == TM_RESID || MODE == TM_MULT_ADD);
    constexpr bool HAS_D = (MODE == TM_JACOBI || MODE == TM_JACOBI_DOT || MODE == TM_PRESMOOTH2);
    constexpr bool NEED_SELF = (MODE == TM_MULT_DOT || MODE == TM_JACOBI || MODE == TM_JACOBI_DOT || MODE == TM_PRESMOOTH2);
    constexpr bool HAS_DOT = (MODE == TM_MULT_DOT || MODE == TM_JACOBI_DOT);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lr = lane >> 2, lk = lane & 3;
    const int rg = warp & 3, cpart = warp >> 2;
    const int NST = op.nst;
    const int nslab = (a.k + SLABW - 1) / SLABW;
    const int nmine = ((int)op.nchunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int T = nmine * nslab;
    const unsigned last_bytes = (unsigned)(a.k - (nslab - 1) * SLABW) * 8u;   // bytes of a row inside the last slab
    const unsigned sm_a = (unsigned)__cvta_generic_to_shared(sm);
    const unsigned rec_a = sm_a + DM_HDR;
    const unsigned st_a = rec_a + DM_NREC * op.cb_bytes;
    unsigned char* rec0 = sm + DM_HDR;
    const unsigned colb = (unsigned)(cpart * CW + lr) * 8u;                     // this lane's B column inside a staged row
    const unsigned cole = (unsigned)(cpart * CW + 2 * lk) * 8u;                 // first of its two C columns (tile g: + 64 g)
    double* scr = reinterpret_cast<double*>(sm + op.off_red) + warp * (8 * CW); // [8 rows][CW columns] transposition scratch
    double* dacc = reinterpret_cast<double*>(sm + op.off_dacc);                // [8 warps][nslab][CW] column sums

    if (tid == 0) {
        for (int s = 0; s < 3; ++s) { mbar_init(sm_a + 8 * s, 1); mbar_init(sm_a + 56 + 8 * s, 8); }
        for (int b = 0; b < DM_NREC; ++b) { mbar_init(sm_a + 24 + 8 * b, 1); mbar_init(sm_a + 80 + 8 * b, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (HAS_DOT && !DREG)
        for (int i = tid; i < 8 * nslab * CW; i += 256) dacc[i] = 0.0;
    double dreg[DM_DREG];
#pragma unroll
    for (int q = 0; q < DM_DREG; ++q) dreg[q] = 0.0;
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();

    // record of chunk ordinal j into ring buffer j & 3 (one thread: expect, then the four pieces)
    auto request_record = [&](int j) {
        if (j >= nmine || tid != 0) return;
        const size_t cidx = (size_t)blockIdx.x + (size_t)j * gridDim.x;
        const unsigned bar = sm_a + 24 + 8 * (j & 3);
        unsigned dst = rec_a + (unsigned)(j & 3) * op.cb_bytes;
        mbar_expect_tx(bar, op.cb_bytes);
        bulk_g2s(dst, op.desc + cidx * op.dl, (unsigned)op.dl * 4u, bar);
        dst += (unsigned)op.dl * 4u;
        bulk_g2s(dst, op.need + cidx * op.n4, (unsigned)op.n4 * 4u, bar);
        dst += (unsigned)op.n4 * 4u;
        bulk_g2s(dst, op.ktmeta + cidx * op.kts, (unsigned)op.kts * 16u, bar);
        dst += (unsigned)op.kts * 16u;
        bulk_g2s(dst, op.apack + cidx * op.ap8, (unsigned)op.ap8 * 8u, bar);
    };
    auto wait_record = [&](int j) { mbar_wait(sm_a + 24 + 8 * (j & 3), (unsigned)(j >> 2) & 1u); };
    auto record = [&](int j) { return reinterpret_cast<const int*>(rec0 + (size_t)(j & 3) * op.cb_bytes); };
    // input rows of item u = (chunk ordinal ju, slab su) into stage `stage`: one bulk copy per row.  cp.async.bulk is a
    // uniform-datapath instruction (the lanes of a warp issue theirs one after the other), so the rows are dealt out
    // over all eight warps, at most LROWS (8 or 16) lanes each.
    const double* urow = nullptr;      // producer side: this lane's input row of the chunk being requested (slab 0 position)
    unsigned udst = 0, unn = 0;
    auto issue_rows = [&](int ju, int su, int stage, unsigned par) {
        mbar_wait(sm_a + 56 + 8 * stage, par ^ 1u);      // all eight warps have released the stage's previous item
        if (su == 0) {       // every warp issues for every item, so it has observed the record since slab 0
            wait_record(ju);
            const int* dsc = record(ju);
            unn = (unsigned)dsc[2];
            const unsigned ss = (unsigned)warp + 8u * (unsigned)lane;      // the launcher guarantees at most 8 * LROWS rows per chunk
            urow = (lane < LROWS && ss < unn) ? a.X + (int64_t)dsc[op.dl + ss] * a.ldx : nullptr;
            udst = ss * DM_PITCH;
        }
        const unsigned rb = (su == nslab - 1) ? last_bytes : (unsigned)SLABW * 8u;
        const unsigned bar = sm_a + 8 * stage;
        if (tid == 0) mbar_expect_tx(bar, unn * rb);
        if (urow) bulk_g2s(st_a + (unsigned)stage * op.st_bytes + udst, urow + su * SLABW, rb, bar);
    };

    for (int j = 0; j < DM_NREC; ++j) request_record(j);
    // producer cursor: item u = t + NST - 1
    int u = 0, uj = 0, us = 0, ustage = 0;
    unsigned upar = 0;
    for (; u < NST - 1 && u < T; ++u) {
        issue_rows(uj, us, ustage, upar);
        if (++us == nslab) { us = 0; ++uj; }
        if (++ustage == NST) { ustage = 0; upar ^= 1u; }
    }
    wait_record(0);

    // right-hand side / scaling values of this lane's outputs, fetched one item ahead
    double bq[NG][2], di = 0.0, di2 = 0.0;
    const double* brow = nullptr;      // this lane's right-hand-side row of the chunk being prefetched
    double* yrow = nullptr;            // its output row of the chunk being computed
#pragma unroll
    for (int g = 0; g < NG; ++g) bq[g][0] = bq[g][1] = 0.0;
    auto prefetch_b = [&](int r0n, int nrn, int sn) {
        const int r = rg * 8 + lr;
        const bool valid = r < nrn;
        if (HAS_D) {
            di = valid ? a.dinv[r0n + r] : 0.0;
            if (MODE == TM_PRESMOOTH2) di2 = valid ? a.dinv2[r0n + r] : 0.0;
        }
        if (HAS_B) {
            const double* Bc = (MODE == TM_MULT_ADD) ? a.Y : a.B;
            const int64_t lds = (MODE == TM_MULT_ADD) ? a.ldy : a.ldb;
            if (sn == 0) brow = Bc + (int64_t)(r0n + r) * lds + cpart * CW + 2 * lk;
            const double* p = brow + sn * SLABW;
            const int col0 = sn * SLABW + cpart * CW + 2 * lk;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                bq[g][0] = bq[g][1] = 0.0;
                if (valid && col0 + g * 8 < a.k) { const double2 v = *reinterpret_cast<const double2*>(p + g * 8); bq[g][0] = v.x; bq[g][1] = v.y; }
            }
        }
    };
    {
        const int* d0 = record(0);
        prefetch_b(d0[0], d0[1], 0);
    }

    double av[DM_KTR];
    unsigned so[DM_KTR];
    int nkt = 0, r0 = 0, nr = 0;
    unsigned selfoff = 0xffffffffu;
    int j = 0, slab = 0, cstage = 0;
    unsigned cpar = 0;
    for (int t = 0; t < T; ++t) {
        if (u < T) {
            issue_rows(uj, us, ustage, upar);
            ++u;
            if (++us == nslab) { us = 0; ++uj; }
            if (++ustage == NST) { ustage = 0; upar ^= 1u; }
        }
        if (slab == 0) {        // new chunk: block entries and slot offsets of this warp's row group into registers
            const int* dsc = record(j);
            r0 = dsc[0]; nr = dsc[1];
            const unsigned counts = (unsigned)dsc[6];
            nkt = (int)((counts >> (8 * rg)) & 0xffu);
            int kt0 = 0;
#pragma unroll
            for (int g = 0; g < 3; ++g)
                if (g < rg) kt0 += (int)((counts >> (8 * g)) & 0xffu);
            const uint4* meta = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(dsc) + (size_t)(op.dl + op.n4) * 4) + kt0;
            const double* ap = reinterpret_cast<const double*>(meta - kt0 + op.kts);
            const unsigned below = (1u << lane) - 1u;
#pragma unroll
            for (int i = 0; i < DM_KTR; ++i) {
                av[i] = 0.0; so[i] = colb;
                if (i < nkt) {
                    const uint4 m = meta[i];
                    const unsigned pr = (lk & 2) ? m.y : m.x;
                    const unsigned slot = (lk & 1) ? (pr >> 16) : (pr & 0xffffu);
                    so[i] = slot * DM_PITCH + colb;
                    if ((m.z >> lane) & 1u) av[i] = ap[m.w + __popc(m.z & below)];
                }
            }
            yrow = a.Y + (int64_t)(r0 + rg * 8 + lr) * a.ldy + cpart * CW + 2 * lk;
            selfoff = 0xffffffffu;
            if (NEED_SELF && rg * 8 + lr < nr) {
                const unsigned sl = reinterpret_cast<const unsigned short*>(dsc + 8)[rg * 8 + lr];
                if (sl != 0xffffu) selfoff = sl * DM_PITCH + cole;
            }
        }
        const unsigned stg = st_a + (unsigned)cstage * op.st_bytes;
        mbar_wait(sm_a + 8 * cstage, cpar);

        double acc[NG][2];
#pragma unroll
        for (int g = 0; g < NG; ++g) acc[g][0] = acc[g][1] = 0.0;
        double bA[NG], bB[NG];     // B fragments of even / odd k-tiles (static double buffer)
        auto load_frag = [&](double (&f)[NG], unsigned ad) {
            f[0] = lds_f64<0>(ad); f[1] = lds_f64<64>(ad);
            if (NG == 4) { f[NG - 2] = lds_f64<128>(ad); f[NG - 1] = lds_f64<192>(ad); }
        };
        load_frag(bA, stg + so[0]);
#pragma unroll
        for (int i = 0; i < DM_KTR; ++i) {
            if (i >= nkt) break;
            if (i + 1 < DM_KTR) {       // next k-tile's fragments (an unused tile reads column block 0 of slot 0, harmlessly)
                const unsigned ad = stg + so[i + 1 < DM_KTR ? i + 1 : i];
                if (i & 1) load_frag(bA, ad);
                else load_frag(bB, ad);
            }
#pragma unroll
            for (int g = 0; g < NG; ++g) dmma884(acc[g][0], acc[g][1], av[i], (i & 1) ? bB[g] : bA[g]);
        }

        // epilogue: this lane holds row rg*8 + lr, columns cpart*32 + g*8 + 2*lk (+1) of the slab
        const int r = rg * 8 + lr;
        const bool rvalid = r < nr;
        double* Yr = yrow + slab * SLABW;
        double pd[NG][2];
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const int col = slab * SLABW + cpart * CW + g * 8 + 2 * lk;
            double s0 = 0.0, s1 = 0.0;
            if (NEED_SELF && selfoff != 0xffffffffu) {
                const double2 v = lds_v2f64(stg + selfoff + (unsigned)g * 64u);
                s0 = v.x; s1 = v.y;
            }
            const double a0 = acc[g][0], a1 = acc[g][1], b0 = bq[g][0], b1 = bq[g][1];
            double y0, y1, p0 = 0.0, p1 = 0.0;
            if (MODE == TM_MULT) { y0 = a0; y1 = a1; }
            else if (MODE == TM_MULT_ADD) { y0 = b0 + a0; y1 = b1 + a1; }
            else if (MODE == TM_MULT_DOT) { y0 = a0; y1 = a1; p0 = s0 * a0; p1 = s1 * a1; }
            else if (MODE == TM_RESID) { y0 = b0 - a0; y1 = b1 - a1; }
            else if (MODE == TM_PRESMOOTH2) { y0 = fma(di2, s0 - a0, di * s0); y1 = fma(di2, s1 - a1, di * s1); }
            else {
                y0 = fma(di, b0 - a0, s0); y1 = fma(di, b1 - a1, s1);
                if (MODE == TM_JACOBI_DOT) { p0 = b0 * y0; p1 = b1 * y1; }
            }
            const bool ok = rvalid && col < a.k;
            if (ok) *reinterpret_cast<double2*>(Yr + g * 8) = make_double2(y0, y1);
            if (HAS_DOT && DREG) { pd[g][0] = ok ? p0 : 0.0; pd[g][1] = ok ? p1 : 0.0; }
            if (HAS_DOT && !DREG) *reinterpret_cast<double2*>(scr + lr * CW + g * 8 + 2 * lk) = ok ? make_double2(p0, p1) : make_double2(0.0, 0.0);
        }
        if (HAS_DOT && DREG) {
            // sum over the 8 rows (lane bits 2..4) of each of this lane's 2 NG columns; every exchange halves the columns
            // a lane keeps.  NG = 4: bit 4 selects tiles {0,1} / {2,3}, bit 3 the tile of the pair, bit 2 the column of
            // the pair.  NG = 2: bit 4 the tile, bit 3 the column, and the last exchange is a plain sum (lanes with
            // bit 2 set hold a duplicate).  Additions commute, so the tree is the same for every column.
            const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
            double r2[2];
            if (NG == 4) {
                double q2[2][2];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int par = 0; par < 2; ++par) {
                        const double keep = b4 ? pd[(i + 2) % NG][par] : pd[i][par], send = b4 ? pd[i][par] : pd[(i + 2) % NG][par];
                        q2[i][par] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                    }
#pragma unroll
                for (int par = 0; par < 2; ++par) {
                    const double keep = b3 ? q2[1][par] : q2[0][par], send = b3 ? q2[0][par] : q2[1][par];
                    r2[par] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
            } else {
#pragma unroll
                for (int par = 0; par < 2; ++par) {
                    const double keep = b4 ? pd[1][par] : pd[0][par], send = b4 ? pd[0][par] : pd[1][par];
                    r2[par] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
            }
            double s;
            if (NG == 4) {
                const double keep = b2 ? r2[1] : r2[0], send = b2 ? r2[0] : r2[1];
                s = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            } else {
                const double keep = b3 ? r2[1] : r2[0], send = b3 ? r2[0] : r2[1];
                const double t = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                s = t + __shfl_xor_sync(0xffffffffu, t, 4);
            }
#pragma unroll
            for (int q = 0; q < DM_DREG; ++q)
                if (slab == q) dreg[q] += s;
        }
        if (HAS_DOT && !DREG) {
            __syncwarp();
            double s = 0.0;
            if (lane < CW) {
#pragma unroll
                for (int q = 0; q < 8; ++q) s += scr[q * CW + lane];
                dacc[(warp * nslab + slab) * CW + lane] += s;
            }
            __syncwarp();
        }

        // next item: its right-hand side rows (a new chunk's row range comes from the next record)
        int jn = j, sn = slab + 1;
        if (sn == nslab) { sn = 0; jn = j + 1; }
        if (t + 1 < T) {
            int r0n = r0, nrn = nr;
            if (sn == 0) {
                wait_record(jn);
                const int* dn = record(jn);
                r0n = dn[0]; nrn = dn[1];
            }
            prefetch_b(r0n, nrn, sn);
        }
        // release the stage (and, after a chunk's last slab, its record): one arrival per warp -- no block-wide barrier
        // in the loop, the warps drift up to NST - 1 items apart
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(sm_a + 56 + 8 * cstage);
            if (sn == 0) mbar_arrive(sm_a + 80 + 8 * (j & 3));
        }
        if (sn == 0 && j >= 1 && tid == 0) {     // one chunk later every warp has certainly left record j - 1: refill its buffer
            mbar_wait(sm_a + 80 + 8 * ((j - 1) & 3), (unsigned)((j - 1) >> 2) & 1u);
            request_record(j - 1 + DM_NREC);
        }
        j = jn; slab = sn;
        if (++cstage == NST) { cstage = 0; cpar ^= 1u; }
    }
    if (HAS_DOT) {
        __syncthreads();      // every item consumed: no bulk copy outstanding, the stage buffers are free
        if (DREG) {
            // this lane's column inside the warp's CW: NG = 4: tile 2*b4 + b3, pair lk, element b2;  NG = 2: tile b4, element b3
            const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1;
            const int cown = NG == 4 ? (b4 * 2 + b3) * 8 + 2 * lk + b2 : b4 * 8 + 2 * lk + b3;
            dacc = reinterpret_cast<double*>(sm + DM_HDR + DM_NREC * (size_t)op.cb_bytes);
            if (NG == 4 || b2 == 0) {
#pragma unroll
                for (int q = 0; q < DM_DREG; ++q)
                    if (q < nslab) dacc[(warp * nslab + q) * CW + cown] = dreg[q];
            }
            __syncthreads();
        }
        for (int i = tid; i < nslab * SLABW; i += 256) {
            const int sl = i / SLABW, c = i % SLABW, cp = c / CW, cc = c % CW;
            double s = 0.0;
#pragma unroll
            for (int g = 0; g < 4; ++g) s += dacc[((cp * 4 + g) * nslab + sl) * CW + cc];
            const int col = sl * SLABW + c;
            if (col < a.k) a.partials[(size_t)blockIdx.x * a.k + col] = s;
        }
    }
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
// slab width of the tensor-core kernel for an operator: 64 columns when its chunks fit the wide shape, else 32
// (force32: ctx->tile_dmma == 2, the A/B and parity-test switch for the narrow shape)
inline int dmma_slab_width(const TileOp& op, bool force32) {
    if (force32) return 32;
    return (op.max_need <= DmCfg<64>::MAXNEED && op.max_kt <= DmCfg<64>::KTR) ? 64 : 32;
}
inline size_t dmma_pitch(int slabw) { return slabw == 64 ? DmCfg<64>::PITCH : DmCfg<32>::PITCH; }
inline bool dmma_dot_in_registers(const TileOp& op, int k, bool force32) {
    const int sw = dmma_slab_width(op, force32), nslab = (k + sw - 1) / sw;
    return nslab <= DM_DREG && (size_t)op.max_need * dmma_pitch(sw) >= (size_t)8 * nslab * (sw / 2) * sizeof(double);
}

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

// shared-memory geometry of tileop_dmma_kernel with `nst` stage buffers for k columns; returns the total bytes
size_t tile_layout_dmma(const TileOp& op, int mode, int k, int nst, bool force32, TileDev* d) {
    // dot modes with at most DM_DREG slabs accumulate in registers (the final exchange reuses the stage buffers:
    // 8 x nslab x CW doubles <= 12 KB, less than one stage of any operator with 24 or more staged rows)
    bool has_dot = (mode == TM_MULT_DOT || mode == TM_JACOBI_DOT) && !dmma_dot_in_registers(op, k, force32);
    const int sw = dmma_slab_width(op, force32), cw = sw / 2;
    size_t cb = (size_t)op.dl * 4 + (size_t)op.n4 * 4 + (size_t)op.kts * 16 + (size_t)op.ap8 * 8;
    size_t st = (size_t)op.max_need * dmma_pitch(sw);
    size_t tot = DM_HDR + DM_NREC * cb + (size_t)nst * st;
    size_t o_red = tot;
    if (has_dot) tot += 8 * 8 * (size_t)cw * sizeof(double);
    size_t o_dacc = tot;
    if (has_dot) tot += 8 * (size_t)((k + sw - 1) / sw) * cw * sizeof(double);
    if (d) { d->cb_bytes = (unsigned)cb; d->st_bytes = (unsigned)st; d->off_bs = 0; d->off_drow = 0; d->off_red = (unsigned)o_red; d->off_dacc = (unsigned)o_dacc; d->nst = nst; }
    return tot;
}
// stage depth with which two CTAs of tileop_dmma_kernel fit on an SM (0: none)
int dmma_stages(const TileOp& op, int mode, int k, bool force32) {
    for (int nst = 3; nst >= 2; --nst)
        if (2 * (tile_layout_dmma(op, mode, k, nst, force32, nullptr) + SMEM_RESERVED) <= 227 * 1024) return nst;
    return 0;
}

template <int MODE, int SLABW>
int launch_dmma_sw(sfem_ctx* c, const TileOp& op, TileDev d, const TileArgs& a, int nst, int* nparts) {
    constexpr bool has_dot = (MODE == TM_MULT_DOT || MODE == TM_JACOBI_DOT);
    size_t smem = tile_layout_dmma(op, MODE, a.k, nst, SLABW == 32, &d);
    const bool dreg = has_dot && dmma_dot_in_registers(op, a.k, SLABW == 32);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(tileop_dmma_kernel<MODE, false, SLABW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_CAP);
        if (has_dot) cudaFuncSetAttribute(tileop_dmma_kernel<MODE, has_dot, SLABW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_CAP);
        attr_set = true;
    }
    // persistent grid: two CTAs per SM (register-limited; the caller checked the shared memory).  Three CTAs per SM at
    // 80 registers (launch bounds 256 x 3, two stages) were measured 7 % slower: spills and a shallower pipeline.
    const int64_t want = 2 * (int64_t)SFEM_NUM_SMS;
    int grid = (int)((int64_t)op.nchunks < want ? (int64_t)op.nchunks : want);
    if (nparts) *nparts = grid;
    if (dreg) tileop_dmma_kernel<MODE, has_dot, SLABW><<<grid, 256, smem, c->stream>>>(d, a);
    else tileop_dmma_kernel<MODE, false, SLABW><<<grid, 256, smem, c->stream>>>(d, a);
    KCHECK(c);
    return SFEM_OK;
}
template <int MODE>
int launch_dmma(sfem_ctx* c, const TileOp& op, const TileDev& d, const TileArgs& a, int nst, int* nparts) {
    if (dmma_slab_width(op, c->tile_dmma == 2) == 64) return launch_dmma_sw<MODE, 64>(c, op, d, a, nst, nparts);
    return launch_dmma_sw<MODE, 32>(c, op, d, a, nst, nparts);
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
    // FP64 tensor-core form: wide blocks on operators with 32-row chunks, if two CTAs fit on an SM
    if (d.kts > 0 && c->tile_dmma && a.k > 32 && a.k % 2 == 0 && al2) {
        const int nst = dmma_stages(op, MODE, a.k, c->tile_dmma == 2);
        if (nst) return launch_dmma<MODE>(c, op, d, a, nst, nparts);
    }
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

bool tileop_dmma_usable(const TileOp& op) { return op.kts > 0 && dmma_stages(op, TM_MULT, 64, false) > 0; }

void tileop_free(TileOp* op) {
    if (!op) return;
    void* ptrs[] = {op->desc, op->need, op->ecol, op->eval, op->eval_scaled, op->ktmeta, op->apack, op->apack_scaled, op->rgcnt};
    for (void* p : ptrs)
        if (p) sfem_dev_free(p);
    *op = TileOp();
}

// k-tile records and packed block values of the FP64 tensor-core form, from the ELL blocks already on the device.
// Operators that do not fit the form (a row group touching more than TD_BIN slots of one residue) keep kts = 0.
static int tileop_build_dmma(sfem_ctx* c, TileOp* op) {
    if (op->nchunks <= 0) return SFEM_OK;
    int* stats = nullptr;
    CUDA_TRY(c, sfem_dev_malloc(&stats, sizeof(int) * 4));
    CUDA_TRY(c, cudaMemsetAsync(stats, 0, sizeof(int) * 4, c->stream));
    const int tb = 128, nb = (4 * op->nchunks + tb - 1) / tb;      // one thread per (chunk, row group)
    CUDA_TRY(c, sfem_dev_malloc(&op->rgcnt, sizeof(ushort2) * 4 * (size_t)op->nchunks));
    {
        ProfScope ps(c, TAG_SETUP, 0.0, 0.0);
        tile_dmma_count_kernel<<<nb, tb, 0, c->stream>>>(op->nchunks, op->dl, op->e8, op->desc, op->ecol, op->eval, op->rgcnt, stats);
        c->launches++;
    }
    int h[4] = {0, 0, 0, 0};
    cudaMemcpyAsync(h, stats, sizeof(h), cudaMemcpyDeviceToHost, c->stream);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    sfem_dev_free(stats);
    if (e != cudaSuccess) return sfem_fail(c, SFEM_ERR_CUDA, "tileop: k-tile count failed: %s", cudaGetErrorString(e));
    // the kernels keep a row group's k-tiles in registers: 8 (64-column slabs) or 16 (32-column slabs)
    if (h[2] || h[0] <= 0 || h[1] > 65535 || h[3] > DmCfg<32>::KTR || op->max_need > DmCfg<32>::MAXNEED) return SFEM_OK;
    op->max_kt = h[3];
    const int kts = h[0], ap8 = h[1] > 0 ? (h[1] + 1) & ~1 : 2;
    size_t nc = (size_t)op->nchunks;
    CUDA_TRY(c, sfem_dev_malloc(&op->ktmeta, sizeof(uint4) * nc * kts));
    CUDA_TRY(c, sfem_dev_malloc(&op->apack, sizeof(double) * nc * ap8));
    CUDA_TRY(c, cudaMemsetAsync(op->ktmeta, 0, sizeof(uint4) * nc * kts, c->stream));
    CUDA_TRY(c, cudaMemsetAsync(op->apack, 0, sizeof(double) * nc * ap8, c->stream));
    {
        ProfScope ps(c, TAG_SETUP, 0.0, 0.0);
        tile_dmma_fill_kernel<<<nb, tb, 0, c->stream>>>(op->nchunks, op->dl, op->e8, kts, ap8, 1, op->desc, op->ecol, op->eval, op->eval,
                                                        op->rgcnt, op->ktmeta, op->apack);
        KCHECK(c);
    }
    op->kts = kts; op->ap8 = ap8;
    if (getenv("SFEM_DEBUG")) fprintf(stderr, "tileop dmma form: %d k-tiles, %d values per chunk (%d rows, need %d, at most %d k-tiles per row group: %d-column slabs)\n", kts, ap8, op->max_rows, op->max_need, op->max_kt, dmma_slab_width(*op, false));
    return SFEM_OK;
}

int tileop_build(sfem_ctx* c, int64_t nrows, int64_t ncols, const int64_t* ptr, const int32_t* idx, const double* val,
                 const int* chunk_ptr, int nchunks, int max_rows, TileOp* out) {
    tileop_free(out);
    TileOp op;
    op.nrows = nrows; op.ncols = ncols; op.nchunks = nchunks; op.max_rows = max_rows;
    int64_t nnz = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&nnz, ptr + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    int* stats = nullptr;
    CUDA_TRY(c, sfem_dev_malloc(&stats, sizeof(int) * 4));
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
    sfem_dev_free(stats);
    if (e != cudaSuccess) return sfem_fail(c, SFEM_ERR_CUDA, "tileop: count pass failed: %s", cudaGetErrorString(e));
    if (hstats[2]) return sfem_fail(c, SFEM_ERR_LIMIT, "tileop: a chunk has too many entries or columns");
    op.nnz = nnz;
    op.max_need = hstats[0] > 0 ? hstats[0] : 1;
    op.max_ent = max_rows * (hstats[1] > 0 ? hstats[1] : 2);
    op.dl = (8 + (max_rows + 1) / 2 + 3) & ~3;
    op.n4 = (op.max_need + 3) & ~3;
    op.e8 = (op.max_ent + 7) & ~7;
    size_t nc = (size_t)nchunks;
    CUDA_TRY(c, sfem_dev_malloc(&op.desc, sizeof(int) * nc * op.dl));
    CUDA_TRY(c, sfem_dev_malloc(&op.need, sizeof(int) * nc * op.n4));
    CUDA_TRY(c, sfem_dev_malloc(&op.ecol, sizeof(unsigned short) * nc * op.e8));
    CUDA_TRY(c, sfem_dev_malloc(&op.eval, sizeof(double) * nc * op.e8));
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
    if (c->tile_dmma && max_rows > 16 && max_rows <= 32) return tileop_build_dmma(c, out);
    return SFEM_OK;
}

int tileop_scale_columns(sfem_ctx* c, TileOp* op, const double* dinv) {
    if (op->nchunks <= 0) return SFEM_OK;
    if (!op->eval_scaled) CUDA_TRY(c, sfem_dev_malloc(&op->eval_scaled, sizeof(double) * (size_t)op->nchunks * op->e8));
    ProfScope ps(c, TAG_SETUP, 0.0, 0.0);
    tile_scale_kernel<<<op->nchunks, 128, 0, c->stream>>>(op->nchunks, op->dl, op->n4, op->e8, op->desc, op->need, op->ecol, op->eval, dinv, op->eval_scaled);
    KCHECK(c);
    if (op->kts > 0) {
        size_t nc = (size_t)op->nchunks;
        if (!op->apack_scaled) CUDA_TRY(c, sfem_dev_malloc(&op->apack_scaled, sizeof(double) * nc * op->ap8));
        CUDA_TRY(c, cudaMemsetAsync(op->apack_scaled, 0, sizeof(double) * nc * op->ap8, c->stream));
        tile_dmma_fill_kernel<<<(4 * op->nchunks + 127) / 128, 128, 0, c->stream>>>(op->nchunks, op->dl, op->e8, op->kts, op->ap8, 0, op->desc,
                                                                                   op->ecol, op->eval, op->eval_scaled, op->rgcnt, op->ktmeta,
                                                                                   op->apack_scaled);
        KCHECK(c);
    }
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
    d.ktmeta = op.ktmeta; d.kts = op.kts; d.ap8 = op.ap8;
    d.apack = (mode == TM_PRESMOOTH2) ? op.apack_scaled : op.apack;
    if (!d.apack) d.kts = 0;
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
