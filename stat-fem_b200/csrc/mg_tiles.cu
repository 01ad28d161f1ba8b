// Staged-tile form of the lattice multigrid hierarchy and the V-cycle built on it.
//
// mg.cu builds the hierarchy in the caller's numbering (fine CSR level + Galerkin lattice stencils).  This file
// renumbers every level along lattice tiles (so that a chunk of consecutive unknowns is a compact patch), emits each
// operator of the cycle -- A_l, the prolongation P_l and the restriction R_l = P_l^T -- as a CSR matrix in the new
// numbering, and converts it to the staged-tile format of tileop.cu.  The cycle is then a sequence of tile kernels:
//   pre-smoothing  x = S2(b)          (two damped-Jacobi sweeps from zero in ONE pass: TM_PRESMOOTH2)
//   residual       r = b - A x        (TM_RESID)        restriction  b' = R r   (TM_MULT)
//   prolongation   x += P x'          (TM_MULT_ADD)     post-smoothing          (TM_JACOBI, the last sweep on the
//                                                        fine level also returns the r.z partials for CG)
// Level 0 numbering: fine unknowns sorted by the Morton code of the level-1 lattice cell they sit in (a counting
// sort, ascending original index inside a cell); 32 consecutive unknowns ~ a 4x2 block of cells.
// Lattice levels: nodes ordered tile by tile (8x4 in 2-D, 4x4x2 in 3-D; the 2^p+1-th node of a direction forms a
// thin tile of its own), one chunk per tile.  The coarsest level (dense inverse) keeps its lexicographic numbering.
#include "mg.cuh"

int build_cell_list(sfem_ctx* c, const int* cell_of, int64_t n, int ncell, int** cell_start_out, int** cell_nodes_out);
template <typename T>
int exclusive_scan(sfem_ctx* c, const T* in, int64_t n, T* out, T* sums);

namespace {

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// ------------------------------------------------------------------------------------------------ renumbering
struct MortonCfg { int d; int mc[3]; int bits[3]; };

__global__ void morton_key_kernel(int64_t n, MortonCfg M, const int* __restrict__ cidx, int* __restrict__ key) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cc = cidx[i];
    int q[3];
    q[0] = cc % M.mc[0];
    int t = cc / M.mc[0];
    q[1] = t % M.mc[1];
    q[2] = t / M.mc[1];
    int out = 0, pos = 0;
    for (int b = 0; b < 16; ++b)
        for (int k = 0; k < M.d; ++k)
            if (b < M.bits[k]) out |= ((q[k] >> b) & 1) << pos++;
    key[i] = out;
}

__global__ void invert_perm_kernel(int64_t n, const int* __restrict__ perm, int* __restrict__ iperm) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) iperm[perm[q]] = (int)q;
}

__global__ void uniform_chunks_kernel(int nchunks, int rows, int64_t n, int* __restrict__ chunk_ptr) {
    int cidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (cidx > nchunks) return;
    int64_t v = (int64_t)cidx * rows;
    chunk_ptr[cidx] = (int)(v < n ? v : n);
}

struct TileGeom {
    int d;
    int m[3], T[3], nt[3];
    __host__ __device__ int tile_size(int k, int t) const { return t == nt[k] - 1 ? m[k] - t * T[k] : T[k]; }
    // first position of tile (t0, t1, t2) in the tile-major numbering
    __host__ __device__ int64_t tile_offset(int t0, int t1, int t2) const {
        int s1 = tile_size(1, t1), s2 = tile_size(2, t2);
        return (int64_t)m[0] * m[1] * ((int64_t)t2 * T[2]) + ((int64_t)m[0] * ((int64_t)t1 * T[1]) + (int64_t)t0 * T[0] * s1) * s2;
    }
};

TileGeom make_geom(const MGLevel* L) {
    TileGeom g;
    g.d = L->d;
    const int T2[3] = {8, 4, 1}, T3[3] = {4, 4, 2}, T1[3] = {32, 1, 1};
    for (int k = 0; k < 3; ++k) {
        g.m[k] = k < L->d ? L->m[k] : 1;
        g.T[k] = L->d == 1 ? T1[k] : (L->d == 2 ? T2[k] : T3[k]);
        if (k >= L->d) g.T[k] = 1;
        g.nt[k] = (g.m[k] + g.T[k] - 1) / g.T[k];      // the 2^p+1-th node of a direction gets a thin tile of its own
    }
    return g;
}

__global__ void lat_perm_kernel(TileGeom g, int64_t n, int* __restrict__ perm, int* __restrict__ iperm) {
    int64_t J = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (J >= n) return;
    int i[3];
    i[0] = (int)(J % g.m[0]);
    int64_t t = J / g.m[0];
    i[1] = (int)(t % g.m[1]);
    i[2] = (int)(t / g.m[1]);
    int tt[3], l[3], s[3];
    for (int k = 0; k < 3; ++k) {
        tt[k] = i[k] / g.T[k];
        l[k] = i[k] - tt[k] * g.T[k];
        s[k] = g.tile_size(k, tt[k]);
    }
    int64_t pos = g.tile_offset(tt[0], tt[1], tt[2]) + ((int64_t)l[2] * s[1] + l[1]) * s[0] + l[0];
    perm[pos] = (int)J;
    iperm[J] = (int)pos;
}

__global__ void lat_chunk_kernel(TileGeom g, int nchunks, int64_t n, int* __restrict__ chunk_ptr) {
    int cidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (cidx > nchunks) return;
    if (cidx == nchunks) { chunk_ptr[cidx] = (int)n; return; }
    int t0 = cidx % g.nt[0];
    int r = cidx / g.nt[0];
    int t1 = r % g.nt[1], t2 = r / g.nt[1];
    chunk_ptr[cidx] = (int)g.tile_offset(t0, t1, t2);
}

// split every chunk into `parts` pieces of at most `sub` consecutive rows (trailing pieces may be empty)
__global__ void refine_chunks_kernel(int nchunks, int sub, int parts, const int* __restrict__ in, int* __restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > nchunks * parts) return;
    if (t == nchunks * parts) { out[t] = in[nchunks]; return; }
    int cidx = t / parts, p = t - cidx * parts;
    int v = in[cidx] + p * sub;
    out[t] = v < in[cidx + 1] ? v : in[cidx + 1];
}

__global__ void permute_vector_kernel(int64_t n, const int* __restrict__ perm, const double* __restrict__ in, double* __restrict__ out) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) out[q] = in[perm ? perm[q] : q];
}
// second smoothing weight: ratio * dinv on free rows, dinv itself on constrained (identity) rows
__global__ void second_weight_kernel(int64_t n, const int* __restrict__ perm, const double* __restrict__ dinv,
                                     const unsigned char* __restrict__ constrained, double ratio, double* __restrict__ out) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    int64_t i = perm ? perm[q] : q;
    out[q] = constrained[i] ? dinv[i] : ratio * dinv[i];
}

// ------------------------------------------------------------------------------------------------ row generators
// Each generator enumerates the entries of one row (in the renumbered spaces) through a callback f(col, value).
struct GenA0 {     // fine stiffness matrix
    const int64_t* ptr; const int32_t* idx; const double* val; const int* perm; const int* iperm;
    template <class F> __device__ void row(int64_t q, F f) const {
        int i = perm[q];
        for (int64_t p = ptr[i]; p < ptr[i + 1]; ++p) {
            int j = idx[p];
            double v = val[p];
            if (v != 0.0 || j == i) f(iperm[j], v);
        }
    }
};
struct GenP0 {     // fine <- level-1 lattice (multilinear interpolation from the corners of the lattice cell)
    Lat L1; int d; const int* perm0; const int* cidx; const double* frac; const unsigned char* bc;
    const unsigned char* constrained1; const int* iperm1;
    template <class F> __device__ void row(int64_t q, F f) const {
        int i = perm0[q];
        if (bc[i]) return;
        int mc0 = L1.m0 - 1, mc1 = d > 1 ? L1.m1 - 1 : 1;
        int cj = cidx[i];
        int j0 = cj % mc0, tq = cj / mc0;
        int j1 = tq % mc1, j2 = tq / mc1;
        for (int g = 0; g < (1 << d); ++g) {
            double w = corner_weight(d, frac + (int64_t)i * d, g);
            if (w == 0.0) continue;
            int J = lat_index(L1, j0 + (g & 1), j1 + ((g >> 1) & 1), j2 + ((g >> 2) & 1));
            if (constrained1[J]) continue;
            f(iperm1 ? iperm1[J] : J, w);
        }
    }
};
struct GenR0 {     // level-1 lattice <- fine (transpose of GenP0)
    Lat L1; int d; const int* perm1; const int* cell_start; const int* cell_nodes; const double* frac; const unsigned char* bc;
    const unsigned char* constrained1; const int* iperm0;
    template <class F> __device__ void row(int64_t Q, F f) const {
        int J = perm1 ? perm1[Q] : (int)Q;
        if (constrained1[J]) return;
        int i0, i1, i2;
        lat_decode(L1, J, i0, i1, i2);
        int mc0 = L1.m0 - 1, mc1 = d > 1 ? L1.m1 - 1 : 1, mc2 = d > 2 ? L1.m2 - 1 : 1;
        for (int beta = 0; beta < (1 << d); ++beta) {
            int c0 = i0 - (beta & 1), c1 = i1 - ((beta >> 1) & 1), c2 = i2 - ((beta >> 2) & 1);
            if (c0 < 0 || c0 >= mc0 || c1 < 0 || c1 >= mc1 || c2 < 0 || c2 >= mc2) continue;
            int cell = c0 + mc0 * (c1 + mc1 * c2);
            for (int p = cell_start[cell]; p < cell_start[cell + 1]; ++p) {
                int i = cell_nodes[p];
                if (bc[i]) continue;
                double w = corner_weight(d, frac + (int64_t)i * d, beta);
                if (w != 0.0) f(iperm0[i], w);
            }
        }
    }
};
struct GenLatA {   // Galerkin lattice operator from its dense 5^d stencil
    Lat L; int d; const double* stencil; const int* perm; const int* iperm;
    template <class F> __device__ void row(int64_t Q, F f) const {
        int J = perm ? perm[Q] : (int)Q;
        int center = slot_of(d, 0, 0, 0);
        for (int s = 0; s < L.S; ++s) {
            double a = stencil[(int64_t)J * L.S + s];
            if (a == 0.0 && s != center) continue;
            int o0, o1, o2;
            slot_decode(d, s, o0, o1, o2);
            int nb = J + o0 + L.m0 * (o1 + L.m1 * o2);
            f(iperm ? iperm[nb] : nb, a);
        }
    }
};
struct GenLatP {   // lattice l <- lattice l+1
    LatPair P; const int* permF; const int* ipermC; const unsigned char* cf; const unsigned char* cc;
    template <class F> __device__ void row(int64_t Q, F f) const {
        int fi = permF ? permF[Q] : (int)Q;
        if (cf[fi]) return;
        int i[3];
        lat_decode(P.F, fi, i[0], i[1], i[2]);
        int base[3], cnt[3];
        double w[3];
        for (int q = 0; q < 3; ++q) {
            if (q < P.d && P.co[q]) {
                if (i[q] & 1) { base[q] = (i[q] - 1) >> 1; cnt[q] = 2; w[q] = 0.5; }
                else { base[q] = i[q] >> 1; cnt[q] = 1; w[q] = 1.0; }
            } else { base[q] = i[q]; cnt[q] = 1; w[q] = 1.0; }
        }
        for (int c2 = 0; c2 < cnt[2]; ++c2)
            for (int c1 = 0; c1 < cnt[1]; ++c1)
                for (int c0 = 0; c0 < cnt[0]; ++c0) {
                    int I = lat_index(P.C, base[0] + c0, base[1] + c1, base[2] + c2);
                    if (cc[I]) continue;
                    f(ipermC ? ipermC[I] : I, w[0] * w[1] * w[2]);
                }
    }
};
struct GenLatR {   // lattice l+1 <- lattice l (transpose of GenLatP)
    LatPair P; const int* permC; const int* ipermF; const unsigned char* cf; const unsigned char* cc;
    template <class F> __device__ void row(int64_t Q, F f) const {
        int node = permC ? permC[Q] : (int)Q;
        if (cc[node]) return;
        int d = P.d;
        int q0 = P.co[0] ? 1 : 0, q1 = (d > 1 && P.co[1]) ? 1 : 0, q2 = (d > 2 && P.co[2]) ? 1 : 0;
        int s0 = P.co[0] ? 2 : 1, s1 = P.co[1] ? 2 : 1, s2 = P.co[2] ? 2 : 1;
        int I0, I1, I2;
        lat_decode(P.C, node, I0, I1, I2);
        for (int e2 = -q2; e2 <= q2; ++e2)
            for (int e1 = -q1; e1 <= q1; ++e1)
                for (int e0 = -q0; e0 <= q0; ++e0) {
                    int i0 = s0 * I0 + e0, i1 = s1 * I1 + e1, i2 = s2 * I2 + e2;
                    if (i0 < 0 || i0 >= P.F.m0 || i1 < 0 || i1 >= P.F.m1 || i2 < 0 || i2 >= P.F.m2) continue;
                    int fi = lat_index(P.F, i0, i1, i2);
                    if (cf[fi]) continue;
                    f(ipermF ? ipermF[fi] : fi, w1d(P.co[0], e0) * w1d(P.co[1], e1) * w1d(P.co[2], e2));
                }
    }
};

template <class Gen>
__global__ void gen_count_kernel(Gen g, int64_t nrows, int64_t* __restrict__ counts) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nrows) return;
    int cnt = 0;
    g.row(q, [&](int, double) { ++cnt; });
    counts[q] = cnt;
}
template <class Gen>
__global__ void gen_fill_kernel(Gen g, int64_t nrows, const int64_t* __restrict__ ptr, int32_t* __restrict__ idx, double* __restrict__ val) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nrows) return;
    int64_t p = ptr[q];
    g.row(q, [&](int col, double v) { idx[p] = col; val[p] = v; ++p; });
}

// generator -> CSR -> staged tiles
template <class Gen>
int build_op(sfem_ctx* c, const Gen& g, int64_t nrows, int64_t ncols, const int* chunk_ptr, int nchunks, int max_rows, TileOp* out) {
    int64_t *counts = nullptr, *ptr = nullptr, *sums = nullptr;
    CUDA_TRY(c, sfem_dev_malloc(&counts, sizeof(int64_t) * (size_t)(nrows + 1)));
    CUDA_TRY(c, sfem_dev_malloc(&ptr, sizeof(int64_t) * (size_t)(nrows + 1)));
    CUDA_TRY(c, sfem_dev_malloc(&sums, sizeof(int64_t) * (size_t)(nrows / 1024 + 8)));
    int rc;
    int32_t* idx = nullptr;
    double* val = nullptr;
    {
        ProfScope ps(c, TAG_SETUP, 0.0, 0.0);
        gen_count_kernel<Gen><<<(unsigned)cdiv64(nrows, 128), 128, 0, c->stream>>>(g, nrows, counts);
        c->launches++;
        rc = exclusive_scan<int64_t>(c, counts, nrows, ptr, sums);
    }
    int64_t nnz = 0;
    if (rc == SFEM_OK) {
        cudaMemcpyAsync(&nnz, ptr + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) rc = sfem_fail(c, SFEM_ERR_CUDA, "mg tiles: operator count failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (rc == SFEM_OK && (sfem_dev_malloc(&idx, sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1)) != cudaSuccess ||
                          sfem_dev_malloc(&val, sizeof(double) * (size_t)(nnz > 0 ? nnz : 1)) != cudaSuccess))
        rc = sfem_fail(c, SFEM_ERR_CUDA, "mg tiles: out of memory");
    if (rc == SFEM_OK) {
        ProfScope ps(c, TAG_SETUP, 0.0, 0.0);
        gen_fill_kernel<Gen><<<(unsigned)cdiv64(nrows, 128), 128, 0, c->stream>>>(g, nrows, ptr, idx, val);
        c->launches++;
        rc = tileop_build(c, nrows, ncols, ptr, idx, val, chunk_ptr, nchunks, max_rows, out);
    }
    cudaStreamSynchronize(c->stream);
    sfem_dev_free(counts); sfem_dev_free(ptr); sfem_dev_free(sums);
    if (idx) sfem_dev_free(idx);
    if (val) sfem_dev_free(val);
    return rc;
}

// ------------------------------------------------------------------------------------------------ small cycle kernels
// x = dinv .* b   (single pre-smoothing sweep from a zero guess; only used when nu = 1)
__global__ void scale_rows_kernel(int64_t n, int k, const double* __restrict__ dinv, const double* __restrict__ B, double* __restrict__ X) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * k) return;
    X[t] = dinv[t / k] * B[t];
}

// x = Ainv b on the coarsest lattice (dense, lexicographic numbering); thread per (row, column)
__global__ void coarse_dense_kernel(int n, int k, const double* __restrict__ Ainv, const double* __restrict__ B, double* __restrict__ X) {
    int col = blockIdx.x * blockDim.x + threadIdx.x;
    int row = blockIdx.y;
    if (col >= k) return;
    const double* arow = Ainv + (int64_t)row * n;
    double s = 0.0;
    for (int j = 0; j < n; ++j) {
        double a = arow[j];
        if (a != 0.0) s = fma(a, B[(int64_t)j * k + col], s);
    }
    X[(int64_t)row * k + col] = s;
}

}  // namespace

void mg_tiles_free(MGLevel* L) {
    if (L->perm) sfem_dev_free(L->perm);
    if (L->iperm) sfem_dev_free(L->iperm);
    if (L->dinv_t) sfem_dev_free(L->dinv_t);
    if (L->dinv_t2) sfem_dev_free(L->dinv_t2);
    L->dinv_t2 = nullptr;
    L->perm = L->iperm = nullptr;
    L->dinv_t = nullptr;
    tileop_free(&L->tA);
    tileop_free(&L->tP);
    tileop_free(&L->tR);
}

// Called at the end of mg_setup.
int mg_build_tiles(sfem_ctx* c) {
    const int nl = (int)c->mg.size();
    if (nl < 2) return sfem_fail(c, SFEM_ERR_STATE, "mg tiles: hierarchy not built");
    MGLevel* L0 = c->mg[0];
    const int d = L0->d;
    const int64_t n = L0->n;
    std::vector<int*> chunk_ptr(nl, nullptr);
    std::vector<int> nchunks(nl, 0), max_rows(nl, 0);
    int rc = SFEM_OK;
    auto cleanup = [&]() { for (int* p : chunk_ptr) if (p) sfem_dev_free(p); };
#define TRY_RC(expr) do { rc = (expr); if (rc) { cleanup(); return rc; } } while (0)
#define TRY_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return sfem_fail(c, SFEM_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); } } while (0)

    // ---- level 0 numbering: Morton order of the level-1 lattice cells
    {
        MGLevel* L1 = c->mg[1];
        MortonCfg M;
        M.d = d;
        int total_bits = 0;
        for (int q = 0; q < 3; ++q) {
            M.mc[q] = q < d ? imax(L1->m[q] - 1, 1) : 1;
            int b = 0;
            while ((1 << b) < M.mc[q]) ++b;
            M.bits[q] = q < d ? b : 0;
            total_bits += M.bits[q];
        }
        if (total_bits > 26) return sfem_fail(c, SFEM_ERR_LIMIT, "mg tiles: level-1 lattice too fine for the Morton cell list");
        int* key = nullptr;
        TRY_CUDA(sfem_dev_malloc(&key, sizeof(int) * (size_t)n));
        morton_key_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, c->stream>>>(n, M, L0->cidx, key);
        c->launches++;
        int* start = nullptr;
        rc = build_cell_list(c, key, n, 1 << total_bits, &start, &L0->perm);
        cudaStreamSynchronize(c->stream);
        sfem_dev_free(key);
        if (start) sfem_dev_free(start);
        if (rc) { cleanup(); return rc; }
        TRY_CUDA(sfem_dev_malloc(&L0->iperm, sizeof(int) * (size_t)n));
        invert_perm_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, c->stream>>>(n, L0->perm, L0->iperm);
        c->launches++;
        const int R0 = c->tile_rows > 0 ? c->tile_rows : 32;
        nchunks[0] = (int)cdiv64(n, R0);
        max_rows[0] = R0;
        TRY_CUDA(sfem_dev_malloc(&chunk_ptr[0], sizeof(int) * ((size_t)nchunks[0] + 1)));
        uniform_chunks_kernel<<<(nchunks[0] + 256) / 256, 256, 0, c->stream>>>(nchunks[0], R0, n, chunk_ptr[0]);
        c->launches++;
    }
    // ---- lattice numberings
    for (int l = 1; l < nl; ++l) {
        MGLevel* L = c->mg[l];
        bool coarsest_dense = (l == nl - 1) && L->dense_inv;
        if (coarsest_dense) {
            const int R = 128;
            nchunks[l] = (int)cdiv64(L->n, R);
            max_rows[l] = R;
            TRY_CUDA(sfem_dev_malloc(&chunk_ptr[l], sizeof(int) * ((size_t)nchunks[l] + 1)));
            uniform_chunks_kernel<<<(nchunks[l] + 256) / 256, 256, 0, c->stream>>>(nchunks[l], R, L->n, chunk_ptr[l]);
            c->launches++;
            continue;      // identity numbering
        }
        TileGeom g = make_geom(L);
        TRY_CUDA(sfem_dev_malloc(&L->perm, sizeof(int) * (size_t)L->n));
        TRY_CUDA(sfem_dev_malloc(&L->iperm, sizeof(int) * (size_t)L->n));
        lat_perm_kernel<<<(unsigned)cdiv64(L->n, 256), 256, 0, c->stream>>>(g, L->n, L->perm, L->iperm);
        c->launches++;
        nchunks[l] = g.nt[0] * g.nt[1] * g.nt[2];
        max_rows[l] = imin(g.T[0], g.m[0]) * imin(g.T[1], g.m[1]) * imin(g.T[2], g.m[2]);
        TRY_CUDA(sfem_dev_malloc(&chunk_ptr[l], sizeof(int) * ((size_t)nchunks[l] + 1)));
        lat_chunk_kernel<<<(nchunks[l] + 256) / 256, 256, 0, c->stream>>>(g, nchunks[l], L->n, chunk_ptr[l]);
        c->launches++;
    }
    // ---- smoother scalings in the new numbering
    for (int l = 0; l < nl; ++l) {
        MGLevel* L = c->mg[l];
        TRY_CUDA(sfem_dev_malloc(&L->dinv_t, sizeof(double) * (size_t)L->n));
        permute_vector_kernel<<<(unsigned)cdiv64(L->n, 256), 256, 0, c->stream>>>(L->n, L->perm, L->dinv, L->dinv_t);
        c->launches++;
        TRY_CUDA(sfem_dev_malloc(&L->dinv_t2, sizeof(double) * (size_t)L->n));
        second_weight_kernel<<<(unsigned)cdiv64(L->n, 256), 256, 0, c->stream>>>(L->n, L->perm, L->dinv, L->constrained,
                                                                               c->mg_omega2 / c->mg_omega, L->dinv_t2);
        c->launches++;
    }
    // restrictions gather ~5 input rows per output row.  On the level's own 32-row chunks a 2-D restriction stages 153
    // rows: that fits the 32-column-slab shape of the tensor-core kernel, which is tried first.  Where it does not apply
    // (3-D: 405 rows per chunk; tensor-core form switched off) the chunks are split into pieces of 8 rows for the scalar
    // kernel, which keeps the staged working set (and with it the number of resident CTAs) like that of the other operators
    auto build_restriction = [&](auto gen, int l_rows, int64_t nrows, int64_t ncols, TileOp* out) -> int {
        const int sub = 8;
        int parts = (max_rows[l_rows] + sub - 1) / sub;
        if (parts <= 1) return build_op(c, gen, nrows, ncols, chunk_ptr[l_rows], nchunks[l_rows], max_rows[l_rows], out);
        if (c->tile_dmma && max_rows[l_rows] > 16 && max_rows[l_rows] <= 32 && !getenv("SFEM_RESTRICT_SCALAR")) {
            int r = build_op(c, gen, nrows, ncols, chunk_ptr[l_rows], nchunks[l_rows], max_rows[l_rows], out);
            if (r == SFEM_OK && tileop_dmma_usable(*out)) return SFEM_OK;
            tileop_free(out);
            if (r != SFEM_OK && r != SFEM_ERR_LIMIT) return r;
            if (r == SFEM_ERR_LIMIT) c->err[0] = 0;
        }
        int* fine = nullptr;
        int nf = nchunks[l_rows] * parts;
        if (sfem_dev_malloc(&fine, sizeof(int) * ((size_t)nf + 1)) != cudaSuccess) return sfem_fail(c, SFEM_ERR_CUDA, "mg tiles: out of memory");
        refine_chunks_kernel<<<(nf + 256) / 256, 256, 0, c->stream>>>(nchunks[l_rows], sub, parts, chunk_ptr[l_rows], fine);
        c->launches++;
        int r = build_op(c, gen, nrows, ncols, fine, nf, sub, out);
        cudaStreamSynchronize(c->stream);
        sfem_dev_free(fine);
        return r;
    };
    // ---- operators
    {
        GenA0 g{c->A_ptr, c->A_idx, c->A_val, L0->perm, L0->iperm};
        TRY_RC(build_op(c, g, n, n, chunk_ptr[0], nchunks[0], max_rows[0], &L0->tA));
        TRY_RC(tileop_scale_columns(c, &L0->tA, L0->dinv_t));
    }
    {
        MGLevel* L1 = c->mg[1];
        Lat lat1 = make_lat(L1->d, L1->m);
        GenP0 gp{lat1, d, L0->perm, L0->cidx, L0->frac, L0->constrained, L1->constrained, L1->iperm};
        TRY_RC(build_op(c, gp, n, L1->n, chunk_ptr[0], nchunks[0], max_rows[0], &L0->tP));
        GenR0 gr{lat1, d, L1->perm, L0->cell_start, L0->cell_nodes, L0->frac, L0->constrained, L1->constrained, L0->iperm};
        TRY_RC(build_restriction(gr, 1, L1->n, n, &L0->tR));
    }
    for (int l = 1; l < nl; ++l) {
        MGLevel* L = c->mg[l];
        Lat lat = make_lat(L->d, L->m);
        bool coarsest = (l == nl - 1);
        if (!(coarsest && L->dense_inv)) {
            GenLatA ga{lat, L->d, L->stencil, L->perm, L->iperm};
            TRY_RC(build_op(c, ga, L->n, L->n, chunk_ptr[l], nchunks[l], max_rows[l], &L->tA));
            TRY_RC(tileop_scale_columns(c, &L->tA, L->dinv_t));
        }
        if (!coarsest) {
            MGLevel* C = c->mg[l + 1];
            LatPair P = make_pair(L, C);
            GenLatP gp{P, L->perm, C->iperm, L->constrained, C->constrained};
            TRY_RC(build_op(c, gp, L->n, C->n, chunk_ptr[l], nchunks[l], max_rows[l], &L->tP));
            GenLatR gr{P, C->perm, L->iperm, L->constrained, C->constrained};
            TRY_RC(build_restriction(gr, l + 1, C->n, L->n, &L->tR));
        }
    }
    cudaStreamSynchronize(c->stream);
    cleanup();
#undef TRY_RC
#undef TRY_CUDA
    return SFEM_OK;
}

const int* mg_tiles_fine_perm(sfem_ctx* c) { return c->mg.empty() ? nullptr : c->mg[0]->perm; }

// Apply ONE operator of the hierarchy (which: 0 = level operator A_l, 1 = prolongation P_l, 2 = restriction R_l) in the
// level's tile numbering, with the tensor-core form switched on or off: the check that both kernel families agree
// (tests/test_gpu_kernels.py).  X: [cols x k], B and Y: [rows x k], all with leading dimension k; partials:
// [TILE_MAX_PARTS x k] (dot modes), *nparts_host rows of it are written.  rows_out / cols_out (nullable) report the shape.
int mg_tile_apply_one(sfem_ctx* c, int level, int which, int mode, int k, const double* X, const double* B, double* Y,
                      double* partials, int* nparts_host, int use_dmma, int64_t* rows_out, int64_t* cols_out, int* has_dmma_out) {
    if (level < 0 || level >= (int)c->mg.size()) return sfem_fail(c, SFEM_ERR_ARG, "tile_apply_one: no such level");
    MGLevel* L = c->mg[level];
    const TileOp& op = which == 0 ? L->tA : (which == 1 ? L->tP : L->tR);
    if (rows_out) *rows_out = op.nrows;
    if (cols_out) *cols_out = op.ncols;
    if (has_dmma_out) *has_dmma_out = op.kts > 0 ? op.kts : 0;
    if (k <= 0 || !X || !Y) return SFEM_OK;
    if (op.nrows <= 0) return sfem_fail(c, SFEM_ERR_STATE, "tile_apply_one: operator not built on this level");
    TileArgs a;
    a.k = k; a.X = X; a.ldx = k; a.Y = Y; a.ldy = k; a.B = B; a.ldb = k;
    a.dinv = L->dinv_t; a.dinv2 = L->dinv_t2; a.partials = partials;
    const int saved = c->tile_dmma;
    c->tile_dmma = use_dmma;
    int np = 0;
    int rc = tileop_apply(c, TAG_SETUP, op, mode, a, 0.0, &np);
    c->tile_dmma = saved;
    if (nparts_host) *nparts_host = np;
    return rc;
}

// workspace: for every lattice level: x0, x1, b  (each n_l x k)
size_t mg_tiles_workspace_bytes(sfem_ctx* c, int k) {
    size_t bytes = 0;
    for (size_t l = 1; l < c->mg.size(); ++l) bytes += 3 * align256(sizeof(double) * (size_t)c->mg[l]->n * k);
    return bytes + 256;
}

static double op_bytes(const TileOp& op, double vec_rows, int k) { return 8.0 * vec_rows * k + 10.0 * (double)op.nnz + 4.0 * (double)op.nrows; }

// Z = M^-1 R (one V-cycle) in the tile numbering of level 0.  R, Z, T0, T1: [n x k] with leading dimension k (k even
// or all buffers 8-byte aligned only).  dot_partials (nullable): [TILE_MAX_PARTS][k] partial sums of R .* Z, *nparts rows used.
int mg_apply_tiles(sfem_ctx* c, int k, const double* R, double* Z, double* T0, double* T1, void* work, double* dot_partials, int* nparts) {
    const int nl = (int)c->mg.size();
    if (nl < 2 || c->mg[0]->tA.nrows == 0) return sfem_fail(c, SFEM_ERR_STATE, "mg_apply: hierarchy not built");
    const int nu = c->mg_nu_pre;
    std::vector<double*> x0(nl), x1(nl), rhs(nl), xcur(nl);
    unsigned char* w = (unsigned char*)work;
    for (int l = 1; l < nl; ++l) {
        size_t sz = align256(sizeof(double) * (size_t)c->mg[l]->n * k);
        x0[l] = (double*)w; w += sz;
        x1[l] = (double*)w; w += sz;
        rhs[l] = (double*)w; w += sz;
    }
    int rc;
    // smoothing from a zero initial guess on level l: returns the buffer holding the result
    auto presmooth = [&](int l, const double* b, double* bufA, double* bufB, double** out) -> int {
        MGLevel* L = c->mg[l];
        int tagJ = l == 0 ? TAG_CSR_JACOBI : TAG_MG_LATTICE;
        double* cur = bufA;
        double* oth = bufB;
        int done;
        if (nu >= 2) {
            TileArgs a; a.k = k; a.X = b; a.ldx = k; a.Y = cur; a.ldy = k; a.dinv = L->dinv_t; a.dinv2 = L->dinv_t2;
            int r = tileop_apply(c, tagJ, L->tA, TM_PRESMOOTH2, a, op_bytes(L->tA, 2.0 * L->n, k));
            if (r) return r;
            done = 2;
        } else {
            ProfScope ps(c, tagJ, 16.0 * (double)L->n * k, 0.0);
            int64_t tot = L->n * k;
            scale_rows_kernel<<<(unsigned)cdiv64(tot, 256), 256, 0, c->stream>>>(L->n, k, L->dinv_t, b, cur);
            KCHECK(c);
            done = 1;
        }
        for (; done < nu; ++done) {
            TileArgs a; a.k = k; a.X = cur; a.ldx = k; a.Y = oth; a.ldy = k; a.B = b; a.ldb = k; a.dinv = (done & 1) ? L->dinv_t2 : L->dinv_t;
            int r = tileop_apply(c, tagJ, L->tA, TM_JACOBI, a, op_bytes(L->tA, 3.0 * L->n, k));
            if (r) return r;
            std::swap(cur, oth);
        }
        *out = cur;
        return SFEM_OK;
    };
    // ---------------- downward
    double* cur0 = nullptr;
    rc = presmooth(0, R, T0, T1, &cur0);
    if (rc) return rc;
    double* oth0 = (cur0 == T0) ? T1 : T0;
    {
        MGLevel* L = c->mg[0];
        TileArgs a; a.k = k; a.X = cur0; a.ldx = k; a.Y = oth0; a.ldy = k; a.B = R; a.ldb = k;
        rc = tileop_apply(c, TAG_CSR_RESID, L->tA, TM_RESID, a, op_bytes(L->tA, 3.0 * L->n, k));
        if (rc) return rc;
        TileArgs t; t.k = k; t.X = oth0; t.ldx = k; t.Y = rhs[1]; t.ldy = k;
        rc = tileop_apply(c, TAG_MG_XFER, L->tR, TM_MULT, t, op_bytes(L->tR, (double)L->n + c->mg[1]->n, k));
        if (rc) return rc;
    }
    for (int l = 1; l < nl; ++l) {
        MGLevel* L = c->mg[l];
        if (l == nl - 1) {
            if (L->dense_inv) {
                ProfScope ps(c, TAG_MG_LATTICE, 16.0 * (double)L->n * k, 2.0 * (double)L->n * L->n * k);
                coarse_dense_kernel<<<dim3((k + 127) / 128, (unsigned)L->n), 128, 0, c->stream>>>((int)L->n, k, L->dense_inv, rhs[l], x0[l]);
                KCHECK(c);
                xcur[l] = x0[l];
            } else {
                // no dense inverse available: smooth harder on the last level
                double* cur = nullptr;
                rc = presmooth(l, rhs[l], x0[l], x1[l], &cur);
                if (rc) return rc;
                double* oth = (cur == x0[l]) ? x1[l] : x0[l];
                for (int s = nu; s < 8 * nu; ++s) {
                    TileArgs a; a.k = k; a.X = cur; a.ldx = k; a.Y = oth; a.ldy = k; a.B = rhs[l]; a.ldb = k; a.dinv = L->dinv_t;
                    rc = tileop_apply(c, TAG_MG_LATTICE, L->tA, TM_JACOBI, a, op_bytes(L->tA, 3.0 * L->n, k));
                    if (rc) return rc;
                    std::swap(cur, oth);
                }
                xcur[l] = cur;
            }
            break;
        }
        double* cur = nullptr;
        rc = presmooth(l, rhs[l], x0[l], x1[l], &cur);
        if (rc) return rc;
        double* oth = (cur == x0[l]) ? x1[l] : x0[l];
        TileArgs a; a.k = k; a.X = cur; a.ldx = k; a.Y = oth; a.ldy = k; a.B = rhs[l]; a.ldb = k;
        rc = tileop_apply(c, TAG_MG_LATTICE, L->tA, TM_RESID, a, op_bytes(L->tA, 3.0 * L->n, k));
        if (rc) return rc;
        TileArgs t; t.k = k; t.X = oth; t.ldx = k; t.Y = rhs[l + 1]; t.ldy = k;
        rc = tileop_apply(c, TAG_MG_LATTICE, L->tR, TM_MULT, t, op_bytes(L->tR, (double)L->n + c->mg[l + 1]->n, k));
        if (rc) return rc;
        xcur[l] = cur;
    }
    // ---------------- upward
    for (int l = nl - 2; l >= 1; --l) {
        MGLevel* L = c->mg[l];
        double* cur = xcur[l];
        double* oth = (cur == x0[l]) ? x1[l] : x0[l];
        TileArgs p; p.k = k; p.X = xcur[l + 1]; p.ldx = k; p.Y = cur; p.ldy = k;
        rc = tileop_apply(c, TAG_MG_LATTICE, L->tP, TM_MULT_ADD, p, op_bytes(L->tP, 2.0 * L->n + c->mg[l + 1]->n, k));
        if (rc) return rc;
        for (int s = 0; s < nu; ++s) {
            TileArgs a; a.k = k; a.X = cur; a.ldx = k; a.Y = oth; a.ldy = k; a.B = rhs[l]; a.ldb = k;
            a.dinv = ((nu - 1 - s) & 1) ? L->dinv_t2 : L->dinv_t;      // adjoint of the pre-smoother: weights in reverse order
            rc = tileop_apply(c, TAG_MG_LATTICE, L->tA, TM_JACOBI, a, op_bytes(L->tA, 3.0 * L->n, k));
            if (rc) return rc;
            std::swap(cur, oth);
        }
        xcur[l] = cur;
    }
    {
        MGLevel* L = c->mg[0];
        TileArgs p; p.k = k; p.X = xcur[1]; p.ldx = k; p.Y = cur0; p.ldy = k;
        rc = tileop_apply(c, TAG_MG_XFER, L->tP, TM_MULT_ADD, p, op_bytes(L->tP, 2.0 * L->n + c->mg[1]->n, k));
        if (rc) return rc;
        for (int s = 0; s < nu; ++s) {
            bool last = (s == nu - 1);
            double* dst = last ? Z : oth0;
            TileArgs a; a.k = k; a.X = cur0; a.ldx = k; a.Y = dst; a.ldy = k; a.B = R; a.ldb = k;
            a.dinv = ((nu - 1 - s) & 1) ? L->dinv_t2 : L->dinv_t;
            a.partials = dot_partials;
            rc = tileop_apply(c, TAG_CSR_JACOBI, L->tA, (last && dot_partials) ? TM_JACOBI_DOT : TM_JACOBI, a, op_bytes(L->tA, 3.0 * L->n, k),
                              last ? nparts : nullptr);
            if (rc) return rc;
            std::swap(cur0, oth0);
        }
    }
    return SFEM_OK;
}
