// Shared declarations for libstatfem_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#define SFEM_OK 0
#define SFEM_ERR_CUDA (-1)
#define SFEM_ERR_ARG (-2)
#define SFEM_ERR_NOT_SPD (-3)
#define SFEM_ERR_NOT_CONVERGED (-4)
#define SFEM_ERR_STATE (-5)
#define SFEM_ERR_LIMIT (-6)

struct MGLevel;   // mg.cu

struct sfem_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    char err[1024] = {0};
    long long launches = 0;          // kernels launched by this library (bench.py's gpu_launches)
    int dgemm_tag = 7;               // profile family charged for dgemm launches (chol.cu switches it per phase)

    // ---- stiffness A (caller-owned CSR; identity Dirichlet rows/cols) -------------------------------
    int64_t n = 0;
    const int64_t* A_ptr = nullptr;
    const int32_t* A_idx = nullptr;
    const double* A_val = nullptr;
    double* A_dinv = nullptr;        // owned: 1/diag(A)
    int64_t A_nnz = 0;

    // ---- forcing covariance G (caller-owned CSR) -----------------------------------------------------
    int64_t g_n = 0;
    const int64_t* G_ptr = nullptr;
    const int32_t* G_idx = nullptr;
    const double* G_val = nullptr;

    // ---- cell list kept between gcov_count and gcov_fill (owned) --------------------------------------
    struct {
        int64_t n = 0; int d = 0;
        const double* coords = nullptr; const double* b = nullptr;
        double e2s = 0, em2l = 0, cutoff = 0, reg = 0, bmax = 0, two_e2l = 0;
        double lo[3] = {0, 0, 0}; double inv_edge = 0; int nc[3] = {1, 1, 1}; int ncell = 1;
        int* cell_start = nullptr;   // ncell+1
        int* cell_nodes = nullptr;   // n, ascending node index inside each cell
        int64_t max_row_nnz = 0;
        int max_span = 0;            // largest (last - first + 1) kept column of a row (count pass)
        int64_t nnz = 0;
    } gc;

    // ---- multigrid hierarchy (owned) -------------------------------------------------------------------
    std::vector<MGLevel*> mg;
    int mg_nu_pre = 2, mg_nu_post = 2;
    double mg_omega = 0.8;
    double mg_omega2 = 0.8;          // weight of every second sweep (== mg_omega: plain damped Jacobi)
    int tile_rows = 0;               // rows per fine-level chunk of the staged-tile operators (0: default)
    int tile_sw = 0;                 // slab width override for the tile kernels (0: automatic, 32: force 32 columns)
    int tile_dmma = 1;               // FP64 tensor-core form of the staged-tile operators (0: scalar kernels only)

    // ---- optional per-kernel-family timing (CUDA events on the launch stream) + algorithmic work counters ------
    unsigned prof_mask = 0;
    struct ProfPair { int tag; cudaEvent_t e0, e1; };
    std::vector<ProfPair> prof_pairs;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[32] = {0};
    double prof_bytes[32] = {0};
    double prof_flops[32] = {0};
    long long prof_count[32] = {0};

    // ---- scratch arena (owned, grows) --------------------------------------------------------------------
    void* scratch = nullptr; size_t scratch_bytes = 0;
    void* pinned = nullptr;          // small pinned host buffer for scalar read-backs
    void* dense_tmp = nullptr; size_t dense_tmp_bytes = 0;   // panel / trtri temporaries of chol.cu (owned, grows)
    cudaStream_t aux_stream = nullptr;                        // look-ahead stream of the blocked Cholesky (owned, lazy)
    cudaEvent_t ev_main = nullptr, ev_aux = nullptr, ev_head = nullptr;

    // ---- ensemble communicator of the sharded covariance assembly (sharded.cu; NCCL bound at run time) ---------------
    void* comm = nullptr; int comm_owned = 0, comm_rank = 0, comm_size = 1;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_recv[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};

    // ---- CUDA graphs of the fused value+gradient evaluation (dense_ops.cu), keyed by the problem's device pointers ----
    struct DenseGraph { const void* key[8] = {nullptr}; void* exec = nullptr; long long launches = 0; int failed = 0; };
    std::vector<DenseGraph> dense_graphs;
};

int sfem_fail(sfem_ctx* c, int code, const char* fmt, ...);

// Device memory of the library (multigrid hierarchy, operator records, temporaries) is stream-ordered: allocated and
// freed with cudaMallocAsync / cudaFreeAsync on the stream of the context that is executing (set by every entry point
// of capi.cu), from the device's default pool with an unlimited release threshold.  Tearing down a hierarchy of a few
// GB therefore costs microseconds instead of ~100 synchronising cudaFree calls (0.2 s per solver at 1M DOF).
extern thread_local cudaStream_t sfem_tls_stream;
template <typename T>
inline cudaError_t sfem_dev_malloc(T** p, size_t bytes) {
    return cudaMallocAsync(reinterpret_cast<void**>(p), bytes > 0 ? bytes : 16, sfem_tls_stream);
}
inline cudaError_t sfem_dev_free(void* p) { return p ? cudaFreeAsync(p, sfem_tls_stream) : cudaSuccess; }
void* sfem_scratch(sfem_ctx* c, size_t bytes);   // returns device scratch of at least `bytes` (may realloc)

#define CUDA_TRY(c, expr)                                                                         \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            return sfem_fail((c), SFEM_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,     \
                             cudaGetErrorString(_e));                                             \
    } while (0)

#define KCHECK(c)                                                                                 \
    do {                                                                                          \
        (c)->launches++;                                                                          \
        cudaError_t _e = cudaGetLastError();                                                      \
        if (_e != cudaSuccess)                                                                    \
            return sfem_fail((c), SFEM_ERR_CUDA, "%s:%d kernel launch -> %s", __FILE__, __LINE__, \
                             cudaGetErrorString(_e));                                             \
    } while (0)

// kernel families (tags) for sfem_profile_*
enum {
    TAG_CSR_MULT = 0,     // csr_block_kernel MODE_MULT / MODE_MULT_DOT on the stiffness matrix (q = A p, p^T q)
    TAG_CSR_JACOBI = 1,   // csr_block_kernel MODE_JACOBI (fine-level smoother)
    TAG_CSR_RESID = 2,    // csr_block_kernel MODE_RESID
    TAG_SPMM_OTHER = 3,   // csr_block_kernel MODE_MULT on G / P / P^T
    TAG_CG_VEC = 4,       // cg_init / cg_update / cg_pupdate / cg_pz / cg_norm
    TAG_MG_XFER = 5,      // fine <-> lattice restriction / prolongation, jacobi0 on the fine level
    TAG_MG_LATTICE = 6,   // everything on lattice levels
    TAG_DGEMM = 7,
    TAG_POTRF_PANEL = 8,
    TAG_GCOV_COUNT = 9,
    TAG_GCOV_FILL = 10,
    TAG_DENSE_MISC = 11,  // kernel matrix, gradient reductions, small vector kernels
    TAG_SETUP = 12,       // multigrid setup, cell lists, scans
    TAG_CHOL_TRSM = 13,   // dgemm: panel solve of the blocked Cholesky
    TAG_CHOL_SYRK = 14,   // dgemm: trailing update of the blocked Cholesky
    TAG_TRTRI = 15,       // dgemm: recursive-doubling triangular inverse
    TAG_LAUUM = 16,       // dgemm: inv(L)^T inv(L)
    TAG_COUNT = 17
};

struct ProfScope {
    sfem_ctx* c;
    int tag;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool on;
    ProfScope(sfem_ctx* ctx, int t, double bytes, double flops) : c(ctx), tag(t) {
        on = (c->prof_mask >> t) & 1u;
        if (!on) return;
        c->prof_bytes[t] += bytes;
        c->prof_flops[t] += flops;
        c->prof_count[t] += 1;
        for (cudaEvent_t* e : {&e0, &e1}) {
            if (!c->prof_pool.empty()) { *e = c->prof_pool.back(); c->prof_pool.pop_back(); }
            else cudaEventCreate(e);
        }
        cudaEventRecord(e0, c->stream);
    }
    ~ProfScope() {
        if (!on) return;
        cudaEventRecord(e1, c->stream);
        c->prof_pairs.push_back({tag, e0, e1});
    }
};

static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

#define SFEM_NUM_SMS 148

// ---- internal entry points shared between translation units -------------------------------------------------
// spmm.cu
int spmm_csr(sfem_ctx* c, int tag, int64_t nnz, int64_t n, const int64_t* ptr, const int32_t* idx, const double* val,
             int k, const double* X, int64_t ldx, double* Y, int64_t ldy,
             double* dot_partials /*nullable: [nblk][k] of sum_i X[i][c]*Y[i][c]*/, int* nblk_out);
// dgemm.cu
enum { OP_N = 0, OP_T = 1 };
int dgemm(sfem_ctx* c, int opA, int opB, int M, int N, int64_t K, double alpha, const double* A, int64_t lda,
          const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int lower_only);
int dgemm_batched(sfem_ctx* c, int opA, int opB, int M, int N, int64_t K, double alpha, const double* A, int64_t lda,
                  const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int flags, int batch,
                  int64_t strideA, int64_t strideB, int64_t strideC);
// mg.cu
void mg_free(sfem_ctx* c);
