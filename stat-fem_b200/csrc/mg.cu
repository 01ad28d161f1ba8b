// K4: lattice multigrid hierarchy for the stiffness matrix, built on the GPU from (CSR A, DOF coordinates, Dirichlet
// mask).  This file builds the levels in the caller's numbering; mg_tiles.cu renumbers them, converts every operator
// to the staged-tile format and runs the V-cycle.
//
// Hierarchy:  level 0 = the FEM matrix in CSR.  Level 1 = a regular lattice (2^p+1 nodes per direction, about half
// the fine resolution) laid over the bounding box; fine DOFs interpolate multilinearly from the 2^d corners of the
// lattice cell they sit in, whatever the mesh connectivity or DOF numbering is.  Levels 2.. = the lattice coarsened
// by two per direction (standard multilinear transfer).  Every coarse operator is the Galerkin product P^T A P,
// stored as a dense 5^d stencil per lattice node (enough for P1 meshes no coarser than the level-1 lattice).  The
// coarsest lattice is inverted densely.  Smoother: damped Jacobi, nu sweeps before and after (symmetric, so the
// V-cycle is an SPD preconditioner for CG).  All vectors are row-major multi-RHS blocks [n_level x k].
//
// This is the replacement for the sparse direct solve behind Firedrake's LinearSolver (LinearSolver.py:136-140,
// solving_utils.py:51,53): unpreconditioned CG needs ~3.7*N_side iterations on these meshes (SURVEY finding 5).
#include "mg.cuh"
#include <cmath>

int build_cell_list(sfem_ctx* c, const int* cell_of, int64_t n, int ncell, int** cell_start_out, int** cell_nodes_out);

namespace {

// ---------------------------------------------------------------------------------------------- setup kernels
struct FineXfer {
    int64_t n;
    int d;
    const double* coords;
    double lo[3], invH[3];
    Lat L1;
};

__global__ void fine_locate_kernel(FineXfer f, int* __restrict__ cidx, double* __restrict__ frac) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= f.n) return;
    int cc[3] = {0, 0, 0};
    const int mm[3] = {f.L1.m0, f.L1.m1, f.L1.m2};
    for (int k = 0; k < f.d; ++k) {
        double u = (f.coords[i * f.d + k] - f.lo[k]) * f.invH[k];
        // snap to lattice nodes and cell midpoints: the transfer weights of aligned meshes must be exactly 0, 1/2, 1
        // (a weight of 1e-16 would still count as a coupling and fill the coarse stencils)
        double u2 = rint(2.0 * u);
        if (fabs(2.0 * u - u2) < 1e-9 * (1.0 + fabs(u2))) u = 0.5 * u2;
        int cell = (int)floor(u);
        int maxc = mm[k] - 2;
        if (cell < 0) cell = 0;
        if (cell > maxc) cell = maxc;
        double t = u - (double)cell;
        t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
        cc[k] = cell;
        frac[i * f.d + k] = t;
    }
    int mc0 = f.L1.m0 - 1, mc1 = f.d > 1 ? f.L1.m1 - 1 : 1;
    cidx[i] = cc[0] + mc0 * (cc[1] + mc1 * cc[2]);
}

__global__ void csr_dinv_kernel(int64_t n, const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                const double* __restrict__ val, double* __restrict__ dinv) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double dg = 0.0;
    for (int64_t p = ptr[i]; p < ptr[i + 1]; ++p)
        if (idx[p] == i) dg += val[p];
    dinv[i] = dg != 0.0 ? 1.0 / dg : 1.0;
}

// level-1 lattice node is constrained iff its best-supported fine node is a Dirichlet node (or it has no support)
__global__ void l1_constrained_kernel(Lat L, int d, const int* __restrict__ cell_start, const int* __restrict__ cell_nodes,
                                      const double* __restrict__ frac, const unsigned char* __restrict__ bc,
                                      unsigned char* __restrict__ constrained) {
    int lane = threadIdx.x & 31;
    int node = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (node >= L.n) return;
    int i0, i1, i2;
    lat_decode(L, node, i0, i1, i2);
    int mc0 = L.m0 - 1, mc1 = d > 1 ? L.m1 - 1 : 1, mc2 = d > 2 ? L.m2 - 1 : 1;
    double wbc = 0.0, wfree = 0.0;
    for (int beta = 0; beta < (1 << d); ++beta) {
        int c0 = i0 - (beta & 1), c1 = i1 - ((beta >> 1) & 1), c2 = i2 - ((beta >> 2) & 1);
        if (c0 < 0 || c0 >= mc0 || c1 < 0 || c1 >= mc1 || c2 < 0 || c2 >= mc2) continue;
        int cell = c0 + mc0 * (c1 + mc1 * c2);
        for (int q = cell_start[cell] + lane; q < cell_start[cell + 1]; q += 32) {
            int i = cell_nodes[q];
            double w = corner_weight(d, frac + (int64_t)i * d, beta);
            if (bc[i]) wbc = fmax(wbc, w);
            else wfree = fmax(wfree, w);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        wbc = fmax(wbc, __shfl_xor_sync(0xffffffffu, wbc, o));
        wfree = fmax(wfree, __shfl_xor_sync(0xffffffffu, wfree, o));
    }
    if (lane == 0) constrained[node] = (wfree <= 0.0 || wbc >= wfree) ? 1 : 0;
}

// Galerkin product for level 1:  A1[I][J-I] = sum_{i,j} w_iI a_ij w_jJ  over free fine nodes i, j.
// One warp per lattice node; every lane accumulates into its private column of a [S][32] shared array, the columns
// are then summed in lane order (deterministic).  *err is set if a fine coupling reaches beyond the 5^d stencil.
__global__ void l1_galerkin_kernel(Lat L, int d, int warps_per_block, const int64_t* __restrict__ ptr,
                                   const int32_t* __restrict__ idx, const double* __restrict__ val,
                                   const int* __restrict__ cidx, const double* __restrict__ frac,
                                   const unsigned char* __restrict__ bc, const int* __restrict__ cell_start,
                                   const int* __restrict__ cell_nodes, const unsigned char* __restrict__ constrained,
                                   double* __restrict__ stencil, int* __restrict__ err) {
    extern __shared__ double acc_all[];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int node = blockIdx.x * warps_per_block + warp;
    if (node >= L.n) return;
    double* acc = acc_all + (size_t)warp * L.S * 32;
    for (int s = 0; s < L.S; ++s) acc[s * 32 + lane] = 0.0;
    int center = slot_of(d, 0, 0, 0);
    if (constrained[node]) {
        for (int s = lane; s < L.S; s += 32) stencil[(int64_t)node * L.S + s] = (s == center) ? 1.0 : 0.0;
        return;
    }
    int i0, i1, i2;
    lat_decode(L, node, i0, i1, i2);
    int mc0 = L.m0 - 1, mc1 = d > 1 ? L.m1 - 1 : 1, mc2 = d > 2 ? L.m2 - 1 : 1;
    for (int beta = 0; beta < (1 << d); ++beta) {
        int c0 = i0 - (beta & 1), c1 = i1 - ((beta >> 1) & 1), c2 = i2 - ((beta >> 2) & 1);
        if (c0 < 0 || c0 >= mc0 || c1 < 0 || c1 >= mc1 || c2 < 0 || c2 >= mc2) continue;
        int cell = c0 + mc0 * (c1 + mc1 * c2);
        for (int q = cell_start[cell] + lane; q < cell_start[cell + 1]; q += 32) {
            int i = cell_nodes[q];
            if (bc[i]) continue;
            double wi = corner_weight(d, frac + (int64_t)i * d, beta);
            if (wi == 0.0) continue;
            for (int64_t p = ptr[i]; p < ptr[i + 1]; ++p) {
                int j = idx[p];
                if (bc[j]) continue;
                double a = val[p] * wi;
                if (a == 0.0) continue;
                int cj = cidx[j];
                int j0 = cj % mc0, tq = cj / mc0;
                int j1 = tq % mc1, j2 = tq / mc1;
                const double* tj = frac + (int64_t)j * d;
                for (int g = 0; g < (1 << d); ++g) {
                    int J0 = j0 + (g & 1), J1 = j1 + ((g >> 1) & 1), J2 = j2 + ((g >> 2) & 1);
                    double wj = corner_weight(d, tj, g);
                    if (wj == 0.0) continue;
                    if (constrained[lat_index(L, J0, J1, J2)]) continue;
                    int o0 = J0 - i0, o1 = J1 - i1, o2 = J2 - i2;
                    if (o0 < -2 || o0 > 2 || o1 < -2 || o1 > 2 || o2 < -2 || o2 > 2) {
                        *err = 1;
                        continue;
                    }
                    acc[slot_of(d, o0, o1, o2) * 32 + lane] += a * wj;
                }
            }
        }
    }
    __syncwarp();
    for (int s = lane; s < L.S; s += 32) {
        double t = 0.0;
        for (int l = 0; l < 32; ++l) t += acc[s * 32 + l];
        stencil[(int64_t)node * L.S + s] = t;
    }
}

__global__ void lat_constrained_kernel(LatPair P, const unsigned char* __restrict__ cf, unsigned char* __restrict__ cc) {
    int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= P.C.n) return;
    int I0, I1, I2;
    lat_decode(P.C, node, I0, I1, I2);
    int f0 = P.co[0] ? 2 * I0 : I0, f1 = P.co[1] ? 2 * I1 : I1, f2 = P.co[2] ? 2 * I2 : I2;
    cc[node] = cf[lat_index(P.F, f0, f1, f2)];
}

// Galerkin for lattice levels: one thread per (coarse node, coarse offset)
__global__ void lat_galerkin_kernel(LatPair P, const double* __restrict__ sf, const unsigned char* __restrict__ cf,
                                    const unsigned char* __restrict__ cc, double* __restrict__ sc) {
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int S = P.C.S;
    if (tid >= P.C.n * S) return;
    int node = (int)(tid / S), slot = (int)(tid % S);
    int d = P.d;
    int center = slot_of(d, 0, 0, 0);
    if (cc[node]) {
        sc[tid] = (slot == center) ? 1.0 : 0.0;
        return;
    }
    int D0, D1, D2;
    slot_decode(d, slot, D0, D1, D2);
    int I0, I1, I2;
    lat_decode(P.C, node, I0, I1, I2);
    int J0 = I0 + D0, J1 = I1 + D1, J2 = I2 + D2;
    if (J0 < 0 || J0 >= P.C.m0 || J1 < 0 || J1 >= P.C.m1 || J2 < 0 || J2 >= P.C.m2 || cc[lat_index(P.C, J0, J1, J2)]) {
        sc[tid] = 0.0;
        return;
    }
    int r0 = P.co[0] ? 1 : 0, r1 = (d > 1 && P.co[1]) ? 1 : 0, r2 = (d > 2 && P.co[2]) ? 1 : 0;
    int s0 = P.co[0] ? 2 : 1, s1 = P.co[1] ? 2 : 1, s2 = P.co[2] ? 2 : 1;
    double total = 0.0;
    for (int e2 = -r2; e2 <= r2; ++e2)
        for (int e1 = -r1; e1 <= r1; ++e1)
            for (int e0 = -r0; e0 <= r0; ++e0) {
                int i0 = s0 * I0 + e0, i1 = s1 * I1 + e1, i2 = s2 * I2 + e2;
                if (i0 < 0 || i0 >= P.F.m0 || i1 < 0 || i1 >= P.F.m1 || i2 < 0 || i2 >= P.F.m2) continue;
                int fi = lat_index(P.F, i0, i1, i2);
                if (cf[fi]) continue;
                double we = w1d(P.co[0], e0) * w1d(P.co[1], e1) * w1d(P.co[2], e2);
                const double* row = sf + (int64_t)fi * P.F.S;
                for (int f2 = -r2; f2 <= r2; ++f2)
                    for (int f1 = -r1; f1 <= r1; ++f1)
                        for (int f0 = -r0; f0 <= r0; ++f0) {
                            int j0 = s0 * J0 + f0, j1 = s1 * J1 + f1, j2 = s2 * J2 + f2;
                            if (j0 < 0 || j0 >= P.F.m0 || j1 < 0 || j1 >= P.F.m1 || j2 < 0 || j2 >= P.F.m2) continue;
                            int o0 = j0 - i0, o1 = j1 - i1, o2 = j2 - i2;
                            if (o0 < -2 || o0 > 2 || o1 < -2 || o1 > 2 || o2 < -2 || o2 > 2) continue;
                            if (cf[lat_index(P.F, j0, j1, j2)]) continue;
                            double a = row[slot_of(d, o0, o1, o2)];
                            if (a != 0.0) total += we * a * w1d(P.co[0], f0) * w1d(P.co[1], f1) * w1d(P.co[2], f2);
                        }
            }
    sc[tid] = total;
}

// dinv[i] = omega / a_ii on free nodes, 1 on constrained (identity) rows so those are solved exactly by one sweep
__global__ void lat_dinv_kernel(int64_t n, int S, int center, double omega, double* __restrict__ stencil,
                                double* __restrict__ dinv, unsigned char* __restrict__ constrained) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double dg = stencil[i * S + center];
    if (!(dg > 0.0)) {   // unsupported lattice node: make it an identity row
        for (int s = 0; s < S; ++s) stencil[i * S + s] = (s == center) ? 1.0 : 0.0;
        constrained[i] = 1;
        dg = 1.0;
    }
    dinv[i] = constrained[i] ? 1.0 / dg : omega / dg;
}

__global__ void fine_wdinv_kernel(int64_t n, double omega, const double* __restrict__ dinv,
                                  const unsigned char* __restrict__ bc, double* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = bc[i] ? dinv[i] : omega * dinv[i];
}

// dense matrix of the coarsest lattice operator
__global__ void lat_to_dense_kernel(Lat L, int d, const double* __restrict__ stencil, double* __restrict__ D) {
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= L.n * L.S) return;
    int node = (int)(tid / L.S), slot = (int)(tid % L.S);
    double a = stencil[tid];
    if (a == 0.0) return;
    int o0, o1, o2, i0, i1, i2;
    slot_decode(d, slot, o0, o1, o2);
    lat_decode(L, node, i0, i1, i2);
    int j = lat_index(L, i0 + o0, i1 + o1, i2 + o2);
    D[(int64_t)node * L.n + j] = a;
}

// in-place Gauss-Jordan inverse of an SPD matrix (no pivoting), single CTA
__global__ void dense_inverse_kernel(int n, double* __restrict__ A) {
    __shared__ double piv;
    extern __shared__ double colk[];   // n doubles
    for (int k = 0; k < n; ++k) {
        if (threadIdx.x == 0) piv = 1.0 / A[(int64_t)k * n + k];
        for (int i = threadIdx.x; i < n; i += blockDim.x) colk[i] = A[(int64_t)i * n + k];
        __syncthreads();
        double p = piv;
        // row k scaled
        for (int j = threadIdx.x; j < n; j += blockDim.x) A[(int64_t)k * n + j] = (j == k) ? p : A[(int64_t)k * n + j] * p;
        __syncthreads();
        for (int64_t t = threadIdx.x; t < (int64_t)n * n; t += blockDim.x) {
            int i = (int)(t / n), j = (int)(t % n);
            if (i == k) continue;
            double f = colk[i];
            double akj = A[(int64_t)k * n + j];
            A[t] = (j == k) ? -f * p : A[t] - f * akj;
        }
        __syncthreads();
    }
}

}  // namespace

void mg_tiles_free(MGLevel* L);
int mg_build_tiles(sfem_ctx* c);

void mg_free(sfem_ctx* c) {
    for (MGLevel* L : c->mg) {
        mg_tiles_free(L);
        if (L->stencil) sfem_dev_free(L->stencil);
        if (L->dinv) sfem_dev_free(L->dinv);
        if (L->constrained) sfem_dev_free(L->constrained);
        if (L->cidx) sfem_dev_free(L->cidx);
        if (L->frac) sfem_dev_free(L->frac);
        if (L->cell_start) sfem_dev_free(L->cell_start);
        if (L->cell_nodes) sfem_dev_free(L->cell_nodes);
        if (L->dense_inv) sfem_dev_free(L->dense_inv);
        delete L;
    }
    c->mg.clear();
}

int stiffness_dinv(sfem_ctx* c) {
    if (c->A_dinv) sfem_dev_free(c->A_dinv);
    CUDA_TRY(c, sfem_dev_malloc(&c->A_dinv, sizeof(double) * (size_t)c->n));
    csr_dinv_kernel<<<(unsigned)cdiv64(c->n, 256), 256, 0, c->stream>>>(c->n, c->A_ptr, c->A_idx, c->A_val, c->A_dinv);
    KCHECK(c);
    return SFEM_OK;
}

int mg_setup(sfem_ctx* c, int dim, const double* coords, const unsigned char* bc_mask, const double* lo, const double* hi,
             const int* m1, int nu, double omega, int max_coarse) {
    if (!c->A_ptr) return sfem_fail(c, SFEM_ERR_STATE, "mg_setup: stiffness matrix not set");
    if (dim < 1 || dim > 3) return sfem_fail(c, SFEM_ERR_ARG, "mg_setup: dimension must be 1, 2 or 3");
    mg_free(c);
    c->mg_nu_pre = c->mg_nu_post = nu < 1 ? 1 : nu;
    c->mg_omega = omega;
    int64_t n = c->n;
    // ---- level 0
    MGLevel* L0 = new MGLevel();
    c->mg.push_back(L0);
    L0->kind = 0; L0->n = n; L0->d = dim;
    CUDA_TRY(c, sfem_dev_malloc(&L0->constrained, (size_t)n));
    CUDA_TRY(c, cudaMemcpyAsync(L0->constrained, bc_mask, (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(c, sfem_dev_malloc(&L0->dinv, sizeof(double) * (size_t)n));
    fine_wdinv_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, c->stream>>>(n, omega, c->A_dinv, L0->constrained, L0->dinv);
    KCHECK(c);
    CUDA_TRY(c, sfem_dev_malloc(&L0->cidx, sizeof(int) * (size_t)n));
    CUDA_TRY(c, sfem_dev_malloc(&L0->frac, sizeof(double) * (size_t)n * dim));
    // ---- level 1 lattice
    MGLevel* L1 = new MGLevel();
    c->mg.push_back(L1);
    L1->d = dim;
    for (int q = 0; q < 3; ++q) L1->m[q] = q < dim ? (m1[q] < 2 ? 2 : m1[q]) : 1;
    Lat lat1 = make_lat(dim, L1->m);
    L1->n = lat1.n; L1->S = lat1.S;
    if (lat1.n > (int64_t)1 << 30) return sfem_fail(c, SFEM_ERR_LIMIT, "mg_setup: level-1 lattice too large");
    FineXfer fx;
    fx.n = n; fx.d = dim; fx.coords = coords; fx.L1 = lat1;
    int64_t ncell = 1;
    for (int q = 0; q < 3; ++q) {
        fx.lo[q] = q < dim ? lo[q] : 0.0;
        double ext = q < dim ? hi[q] - lo[q] : 1.0;
        fx.invH[q] = (q < dim && ext > 0.0) ? (double)(L1->m[q] - 1) / ext : 0.0;
        if (q < dim) ncell *= (L1->m[q] - 1);
    }
    fine_locate_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, c->stream>>>(fx, L0->cidx, L0->frac);
    KCHECK(c);
    int rc = build_cell_list(c, L0->cidx, n, (int)ncell, &L0->cell_start, &L0->cell_nodes);
    if (rc) return rc;
    CUDA_TRY(c, sfem_dev_malloc(&L1->constrained, (size_t)lat1.n));
    CUDA_TRY(c, sfem_dev_malloc(&L1->stencil, sizeof(double) * (size_t)lat1.n * lat1.S));
    CUDA_TRY(c, sfem_dev_malloc(&L1->dinv, sizeof(double) * (size_t)lat1.n));
    l1_constrained_kernel<<<(unsigned)cdiv64(lat1.n, 4), 128, 0, c->stream>>>(lat1, dim, L0->cell_start, L0->cell_nodes,
                                                                             L0->frac, L0->constrained, L1->constrained);
    KCHECK(c);
    int* err = (int*)sfem_scratch(c, 64);
    CUDA_TRY(c, cudaMemsetAsync(err, 0, sizeof(int), c->stream));
    {
        int wpb = dim == 3 ? 4 : 8;
        size_t dyn = sizeof(double) * (size_t)wpb * lat1.S * 32;
        cudaFuncSetAttribute(l1_galerkin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        l1_galerkin_kernel<<<(unsigned)cdiv64(lat1.n, wpb), wpb * 32, dyn, c->stream>>>(
            lat1, dim, wpb, c->A_ptr, c->A_idx, c->A_val, L0->cidx, L0->frac, L0->constrained, L0->cell_start,
            L0->cell_nodes, L1->constrained, L1->stencil, err);
        KCHECK(c);
    }
    int herr = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (herr) return sfem_fail(c, SFEM_ERR_LIMIT, "mg_setup: mesh edges span more than one level-1 lattice cell; use a coarser lattice");
    int center = dim == 1 ? 2 : (dim == 2 ? 12 : 62);
    lat_dinv_kernel<<<(unsigned)cdiv64(lat1.n, 256), 256, 0, c->stream>>>(lat1.n, lat1.S, center, omega, L1->stencil, L1->dinv, L1->constrained);
    KCHECK(c);
    // ---- coarser lattices
    MGLevel* F = L1;
    while (F->n > max_coarse) {
        bool any = false;
        MGLevel* C = new MGLevel();
        C->d = dim;
        for (int q = 0; q < 3; ++q) {
            if (q < dim && F->m[q] >= 5 && (F->m[q] & 1)) {
                F->coarsen[q] = 1;
                C->m[q] = (F->m[q] + 1) / 2;
                any = true;
            } else {
                F->coarsen[q] = 0;
                C->m[q] = F->m[q];
            }
        }
        if (!any) { delete C; break; }
        c->mg.push_back(C);
        Lat lc = make_lat(dim, C->m);
        C->n = lc.n; C->S = lc.S;
        CUDA_TRY(c, sfem_dev_malloc(&C->constrained, (size_t)lc.n));
        CUDA_TRY(c, sfem_dev_malloc(&C->stencil, sizeof(double) * (size_t)lc.n * lc.S));
        CUDA_TRY(c, sfem_dev_malloc(&C->dinv, sizeof(double) * (size_t)lc.n));
        LatPair P = make_pair(F, C);
        lat_constrained_kernel<<<(unsigned)cdiv64(lc.n, 256), 256, 0, c->stream>>>(P, F->constrained, C->constrained);
        KCHECK(c);
        lat_galerkin_kernel<<<(unsigned)cdiv64(lc.n * lc.S, 128), 128, 0, c->stream>>>(P, F->stencil, F->constrained, C->constrained, C->stencil);
        KCHECK(c);
        lat_dinv_kernel<<<(unsigned)cdiv64(lc.n, 256), 256, 0, c->stream>>>(lc.n, lc.S, center, omega, C->stencil, C->dinv, C->constrained);
        KCHECK(c);
        F = C;
    }
    // ---- dense inverse on the coarsest lattice
    if (F->n <= 1200) {
        Lat lc = make_lat(dim, F->m);
        int nn = (int)lc.n;
        CUDA_TRY(c, sfem_dev_malloc(&F->dense_inv, sizeof(double) * (size_t)nn * nn));
        CUDA_TRY(c, cudaMemsetAsync(F->dense_inv, 0, sizeof(double) * (size_t)nn * nn, c->stream));
        lat_to_dense_kernel<<<(unsigned)cdiv64(lc.n * lc.S, 256), 256, 0, c->stream>>>(lc, dim, F->stencil, F->dense_inv);
        KCHECK(c);
        dense_inverse_kernel<<<1, 1024, sizeof(double) * nn, c->stream>>>(nn, F->dense_inv);
        KCHECK(c);
    }
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return mg_build_tiles(c);
}

int mg_num_levels(sfem_ctx* c) { return (int)c->mg.size(); }
int64_t mg_level_size(sfem_ctx* c, int l) { return (l >= 0 && l < (int)c->mg.size()) ? c->mg[l]->n : -1; }
