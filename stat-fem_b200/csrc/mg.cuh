// Shared declarations of the lattice multigrid (mg.cu: hierarchy setup; mg_tiles.cu: staged-tile operators and the cycle).
#pragma once
#include "tileop.cuh"

struct MGLevel {
    int kind = 1;                 // 0 = CSR fine level, 1 = lattice
    int64_t n = 0;
    int d = 1;
    int m[3] = {1, 1, 1};         // lattice nodes per direction
    int S = 5;                    // stencil entries per node (5^d)
    double* stencil = nullptr;    // [n][S]
    double* dinv = nullptr;       // [n]
    unsigned char* constrained = nullptr;   // [n]  (level 0: Dirichlet mask, owned copy)
    // level 0 -> level 1 transfer
    int* cidx = nullptr;          // [n] lattice cell of each fine node
    double* frac = nullptr;       // [n][d] position inside the cell, in [0,1]
    int* cell_start = nullptr;
    int* cell_nodes = nullptr;
    // lattice -> next lattice
    int coarsen[3] = {0, 0, 0};
    double* dense_inv = nullptr;  // coarsest: [n][n]
    // ---- staged-tile form (mg_tiles.cu): everything below lives in this level's renumbered ("tile") ordering
    int* perm = nullptr;          // [n] tile position -> original index   (nullptr: identity)
    int* iperm = nullptr;         // [n] original index -> tile position
    double* dinv_t = nullptr;     // [n] smoother scaling in tile ordering (first weight)
    double* dinv_t2 = nullptr;    // [n] second smoothing weight
    TileOp tA;                    // level operator
    TileOp tP;                    // prolongation from level l+1   (n_l x n_{l+1})
    TileOp tR;                    // restriction to level l+1      (n_{l+1} x n_l)
};

struct Lat {
    int d, m0, m1, m2, S;
    int64_t n;
};
__host__ __device__ inline Lat make_lat(int d, const int* m) {
    Lat L;
    L.d = d; L.m0 = m[0]; L.m1 = d > 1 ? m[1] : 1; L.m2 = d > 2 ? m[2] : 1;
    L.S = d == 1 ? 5 : (d == 2 ? 25 : 125);
    L.n = (int64_t)L.m0 * L.m1 * L.m2;
    return L;
}
__device__ __forceinline__ void lat_decode(const Lat& L, int node, int& i0, int& i1, int& i2) {
    i0 = node % L.m0;
    int t = node / L.m0;
    i1 = t % L.m1;
    i2 = t / L.m1;
}
__device__ __forceinline__ int lat_index(const Lat& L, int i0, int i1, int i2) { return i0 + L.m0 * (i1 + L.m1 * i2); }
// stencil slot <-> offset
__device__ __forceinline__ int slot_of(int d, int o0, int o1, int o2) {
    int s = o0 + 2;
    if (d > 1) s += 5 * (o1 + 2);
    if (d > 2) s += 25 * (o2 + 2);
    return s;
}
__device__ __forceinline__ void slot_decode(int d, int s, int& o0, int& o1, int& o2) {
    o0 = s % 5 - 2;
    o1 = d > 1 ? (s / 5) % 5 - 2 : 0;
    o2 = d > 2 ? s / 25 - 2 : 0;
}

__device__ __forceinline__ double corner_weight(int d, const double* __restrict__ t, int beta) {
    double w = (beta & 1) ? t[0] : 1.0 - t[0];
    if (d > 1) w *= (beta & 2) ? t[1] : 1.0 - t[1];
    if (d > 2) w *= (beta & 4) ? t[2] : 1.0 - t[2];
    return w;
}

// lattice -> coarser lattice
struct LatPair {
    Lat F, C;
    int d;
    int co[3];     // direction coarsened?
};

__device__ __forceinline__ double w1d(int co, int e) { return co ? (e == 0 ? 1.0 : 0.5) : 1.0; }

inline LatPair make_pair(const MGLevel* F, const MGLevel* C) {
    LatPair P;
    P.F = make_lat(F->d, F->m);
    P.C = make_lat(C->d, C->m);
    P.d = F->d;
    for (int q = 0; q < 3; ++q) P.co[q] = F->coarsen[q];
    return P;
}

