// The sharded covariance assembly behind the C ABI: what solving_utils.py:117-159 of the reference does over a Firedrake
// ensemble communicator -- contiguous ownership of the sensor columns, independent solves per rank, gather of the result
// rows on ensemble rank 0 -- as ONE call per rank over an NCCL communicator, so that a host program on mpi4py (or
// anything else that can move 128 bytes) reaches the multi-GPU path without torch.distributed.
//
//   sfem_nccl_unique_id  rank 0 creates the rendezvous token; the caller distributes it (MPI_Bcast, a file, ...)
//   sfem_comm_init       one communicator per context (= per GPU); sfem_comm_attach adopts an existing ncclComm_t
//   sfem_prior_cov       Z_r = A^-1 P[:, J_r]  (block PCG),  W_r = G Z_r,  rows J_r of Cu = W_r^T Z  accumulated while the
//                        other ranks' Z blocks travel round the ring (ncclSend / ncclRecv on a second stream, row chunk by
//                        row chunk, double-buffered, overlapped with the DMMA GEMM), rows gathered on rank 0
//   sfem_comm_bcast      e.g. the cached prior (x, mu, Cu) for per-rank MAP fits
//
// NCCL is bound at run time (dlopen of libnccl.so.2, the copy already in the process if there is one): the library has no
// link-time dependency on it and single-GPU users never load it.  The Python package drives the same exchange through
// torch.distributed with the half-ring schedule for the symmetric case (solving_utils.sharded_symmetric_wtz); this entry
// point uses the plain ring (every rank multiplies all size blocks), which needs no knowledge of the structure of G.
#include "common.cuh"
#include <dlfcn.h>
#include <algorithm>

int cg_solve_block(sfem_ctx* c, int k, const double* B, int64_t ldb, double* X, int64_t ldx, int precond, double rtol,
                   double atol, int max_it, int* iters_out, double* relres_out_host, void* work, size_t work_bytes);
size_t cg_workspace_bytes(sfem_ctx* c, int k, int precond);
int scatter_columns(sfem_ctx* c, int64_t n, const int64_t* colptr, const int32_t* rows, const double* vals,
                    int64_t col_begin, int k, double* B, int64_t ldb);
struct B84Matrix;
int b84_build(sfem_ctx* c, int64_t n, const int64_t* ptr, const int32_t* idx, const double* val, B84Matrix** out);
int b84_spmm(sfem_ctx* c, const B84Matrix* m, int k, const double* X, int64_t ldx, double* Y, int64_t ldy);
void b84_free(B84Matrix* m);

namespace {

// ---- the slice of the NCCL API used here (stable across NCCL 2.x) ------------------------------------------------
typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm;
enum { NCCL_INT64 = 4, NCCL_F64 = 8, NCCL_CHAR = 0 };
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(nccl_uid*) = nullptr;
    int (*CommInitRank)(nccl_comm*, int, nccl_uid, int) = nullptr;
    int (*CommDestroy)(nccl_comm) = nullptr;
    int (*Send)(const void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api;
    if (api.handle || api.ok) return api;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        api.handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);     // the copy the process already uses (torch's)
        if (api.handle) break;
    }
    if (!api.handle)
        for (const char* nm : names) {
            api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
    if (!api.handle) return api;
#define BIND(field, sym) *(void**)(&api.field) = dlsym(api.handle, sym)
    BIND(GetUniqueId, "ncclGetUniqueId");
    BIND(CommInitRank, "ncclCommInitRank");
    BIND(CommDestroy, "ncclCommDestroy");
    BIND(Send, "ncclSend");
    BIND(Recv, "ncclRecv");
    BIND(GroupStart, "ncclGroupStart");
    BIND(GroupEnd, "ncclGroupEnd");
    BIND(AllGather, "ncclAllGather");
    BIND(Broadcast, "ncclBroadcast");
    BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Send && api.Recv && api.GroupStart &&
             api.GroupEnd && api.AllGather && api.Broadcast;
    return api;
}

#define NCCL_TRY(c, expr)                                                                                          \
    do {                                                                                                           \
        int _r = (expr);                                                                                           \
        if (_r != 0)                                                                                               \
            return sfem_fail((c), SFEM_ERR_CUDA, "%s:%d %s -> NCCL error %d (%s)", __FILE__, __LINE__, #expr, _r,  \
                             nccl().GetErrorString ? nccl().GetErrorString(_r) : "?");                             \
    } while (0)


}  // namespace

int dgemm(sfem_ctx* c, int opA, int opB, int M, int N, int64_t K, double alpha, const double* A, int64_t lda,
          const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int lower_only);

int sharded_unique_id(void* id_out) {
    if (!id_out) return SFEM_ERR_ARG;
    if (!nccl().ok) return SFEM_ERR_STATE;
    nccl_uid id;
    if (nccl().GetUniqueId(&id) != 0) return SFEM_ERR_CUDA;
    memcpy(id_out, &id, sizeof(id));
    return SFEM_OK;
}

int sharded_comm_init(sfem_ctx* c, const void* id128, int rank, int nranks) {
    if (nranks < 1 || rank < 0 || rank >= nranks) return sfem_fail(c, SFEM_ERR_ARG, "comm_init: bad rank / size");
    if (c->comm && c->comm_owned) { nccl().CommDestroy((nccl_comm)c->comm); c->comm = nullptr; }
    c->comm_rank = rank; c->comm_size = nranks; c->comm_owned = 0;
    if (nranks == 1) return SFEM_OK;                       // a single rank needs no communicator
    if (!id128) return sfem_fail(c, SFEM_ERR_ARG, "comm_init: no unique id");
    if (!nccl().ok) return sfem_fail(c, SFEM_ERR_STATE, "comm_init: libnccl.so.2 could not be loaded");
    nccl_uid id;
    memcpy(&id, id128, sizeof(id));
    nccl_comm comm = nullptr;
    NCCL_TRY(c, nccl().CommInitRank(&comm, nranks, id, rank));
    c->comm = comm;
    c->comm_owned = 1;
    return SFEM_OK;
}

int sharded_comm_attach(sfem_ctx* c, void* nccl_comm_handle, int rank, int nranks) {
    if (nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !nccl_comm_handle))
        return sfem_fail(c, SFEM_ERR_ARG, "comm_attach: bad arguments");
    if (nranks > 1 && !nccl().ok) return sfem_fail(c, SFEM_ERR_STATE, "comm_attach: libnccl.so.2 could not be loaded");
    if (c->comm && c->comm_owned) nccl().CommDestroy((nccl_comm)c->comm);
    c->comm = nccl_comm_handle; c->comm_owned = 0; c->comm_rank = rank; c->comm_size = nranks;
    return SFEM_OK;
}

int sharded_comm_destroy(sfem_ctx* c) {
    if (c->comm && c->comm_owned) {
        cudaStreamSynchronize(c->stream);
        if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
        nccl().CommDestroy((nccl_comm)c->comm);
    }
    c->comm = nullptr; c->comm_owned = 0; c->comm_rank = 0; c->comm_size = 1;
    if (c->comm_stream) { cudaStreamDestroy(c->comm_stream); c->comm_stream = nullptr; }
    for (int q = 0; q < 2; ++q) {
        if (c->ev_recv[q]) { cudaEventDestroy(c->ev_recv[q]); c->ev_recv[q] = nullptr; }
        if (c->ev_free[q]) { cudaEventDestroy(c->ev_free[q]); c->ev_free[q] = nullptr; }
    }
    return SFEM_OK;
}

int sharded_bcast(sfem_ctx* c, void* buf, size_t bytes, int root) {
    if (c->comm_size <= 1) return SFEM_OK;
    NCCL_TRY(c, nccl().Broadcast(buf, buf, bytes, NCCL_CHAR, root, (nccl_comm)c->comm, c->stream));
    return SFEM_OK;
}

// See the header (include/statfem_b200.h) for the contract.
int sharded_prior_cov(sfem_ctx* c, int64_t n_obs, const int64_t* P_colptr, const int32_t* P_rows, const double* P_vals,
                      const int64_t* G_indptr, const int32_t* G_indices, const double* G_vals, int64_t col_begin,
                      int64_t col_end, int block_cols, int precond, double rtol, double atol, int max_it, double* Z_out,
                      double* W_out, double* Cu_out, int* iters_max_out_host) {
    const int R = c->comm_size, r = c->comm_rank;
    const int64_t n = c->n;
    if (!c->A_ptr) return sfem_fail(c, SFEM_ERR_STATE, "prior_cov: stiffness matrix not set");
    if (n_obs <= 0 || col_begin < 0 || col_end < col_begin || col_end > n_obs || !P_colptr || !G_indptr)
        return sfem_fail(c, SFEM_ERR_ARG, "prior_cov: bad arguments");
    if (r == 0 && !Cu_out) return sfem_fail(c, SFEM_ERR_ARG, "prior_cov: rank 0 needs Cu_out");
    const int64_t nloc = col_end - col_begin;
    if (nloc > INT32_MAX || n_obs > INT32_MAX) return sfem_fail(c, SFEM_ERR_ARG, "prior_cov: too many columns");
    // ---- column ranges of every rank (the caller's split need not be PETSc's)
    std::vector<int64_t> ranges(2 * (size_t)R);
    ranges[2 * r] = col_begin; ranges[2 * r + 1] = col_end;
    if (R > 1) {
        int64_t* dr = nullptr;
        CUDA_TRY(c, sfem_dev_malloc(&dr, sizeof(int64_t) * 2 * R));
        CUDA_TRY(c, cudaMemcpyAsync(dr + 2 * r, &ranges[2 * r], sizeof(int64_t) * 2, cudaMemcpyHostToDevice, c->stream));
        NCCL_TRY(c, nccl().AllGather(dr + 2 * r, dr, 2, NCCL_INT64, (nccl_comm)c->comm, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(ranges.data(), dr, sizeof(int64_t) * 2 * R, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        sfem_dev_free(dr);
        int64_t expect = 0;
        for (int s = 0; s < R; ++s) {
            if (ranges[2 * s] != expect || ranges[2 * s + 1] < ranges[2 * s])
                return sfem_fail(c, SFEM_ERR_ARG, "prior_cov: the ranks' column ranges must tile [0, n_obs) in rank order");
            expect = ranges[2 * s + 1];
        }
        if (expect != n_obs) return sfem_fail(c, SFEM_ERR_ARG, "prior_cov: the ranks' column ranges must tile [0, n_obs)");
    } else if (col_begin != 0 || col_end != n_obs) {
        return sfem_fail(c, SFEM_ERR_ARG, "prior_cov: a single rank owns all columns");
    }
    int64_t wmax = 0;
    for (int s = 0; s < R; ++s) wmax = std::max(wmax, ranges[2 * s + 1] - ranges[2 * s]);

    // ---- Z_r = A^-1 P[:, J_r] in blocks of columns, W_r = G Z_r
    double *Z = Z_out, *W = W_out;
    const size_t zbytes = sizeof(double) * (size_t)n * (size_t)std::max<int64_t>(nloc, 1);
    if (!Z) CUDA_TRY(c, sfem_dev_malloc(&Z, zbytes));
    if (!W) CUDA_TRY(c, sfem_dev_malloc(&W, zbytes));
    int rc = SFEM_OK, iters_max = 0;
    if (nloc > 0) {
        int kb = block_cols > 0 ? block_cols : 384;
        if (kb > nloc) kb = (int)nloc;
        size_t wsb = cg_workspace_bytes(c, kb, precond);
        unsigned char* work = nullptr;
        double* B = nullptr;
        CUDA_TRY(c, sfem_dev_malloc(&work, wsb + 256));
        CUDA_TRY(c, sfem_dev_malloc(&B, sizeof(double) * (size_t)n * (kb + (kb & 1))));
        unsigned char* wal = work + ((256 - (reinterpret_cast<uintptr_t>(work) & 255)) & 255);
        for (int64_t j0 = 0; j0 < nloc && rc == SFEM_OK; j0 += kb) {
            int k = (int)std::min<int64_t>(kb, nloc - j0);
            rc = scatter_columns(c, n, P_colptr, P_rows, P_vals, col_begin + j0, k, B, k);
            int iters = 0;
            if (rc == SFEM_OK)
                rc = cg_solve_block(c, k, B, k, Z + j0, nloc, precond, rtol, atol, max_it, &iters, nullptr, wal, wsb);
            iters_max = std::max(iters_max, iters);
        }
        sfem_dev_free(B);
        sfem_dev_free(work);
        if (rc == SFEM_OK) {
            B84Matrix* gb = nullptr;
            int64_t gnnz = 0;
            CUDA_TRY(c, cudaMemcpyAsync(&gnnz, G_indptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(c, cudaStreamSynchronize(c->stream));
            const bool block_form = nloc >= 256 && nloc % 2 == 0 && gnnz >= 16 * n;
            if (block_form) {
                rc = b84_build(c, n, G_indptr, G_indices, G_vals, &gb);
                if (rc == SFEM_OK) rc = b84_spmm(c, gb, (int)nloc, Z, nloc, W, nloc);
                CUDA_TRY(c, cudaStreamSynchronize(c->stream));
                if (gb) b84_free(gb);
            } else {
                rc = spmm_csr(c, TAG_SPMM_OTHER, gnnz, n, G_indptr, G_indices, G_vals, (int)nloc, Z, nloc, W, nloc, nullptr, nullptr);
            }
        }
    }
    if (iters_max_out_host) *iters_max_out_host = iters_max;
    // every rank must reach the exchange, or the others would wait for ever: a failed solve is reported after it
    int failed = rc;

    // ---- rows J_r of Cu = W_r^T Z, ring over the ranks
    double* res = nullptr;
    CUDA_TRY(c, sfem_dev_malloc(&res, sizeof(double) * (size_t)std::max<int64_t>(nloc, 1) * (size_t)n_obs));
    if (nloc > 0 && failed == SFEM_OK)
        rc = dgemm(c, OP_T, OP_N, (int)nloc, (int)nloc, n, 1.0, W, nloc, Z, nloc, 0.0, res + col_begin, n_obs, 0);
    if (R > 1) {
        if (!c->comm_stream) {
            CUDA_TRY(c, cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
            for (int q = 0; q < 2; ++q) {
                CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_recv[q], cudaEventDisableTiming));
                CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_free[q], cudaEventDisableTiming));
            }
        }
        int64_t chunk = std::max<int64_t>(1024, std::min<int64_t>(n, (int64_t)(1.0e9 / (8.0 * (double)std::max<int64_t>(wmax, 1)))));
        double* buf[2] = {nullptr, nullptr};
        for (int q = 0; q < 2; ++q) CUDA_TRY(c, sfem_dev_malloc(&buf[q], sizeof(double) * (size_t)chunk * (size_t)std::max<int64_t>(wmax, 1)));
        struct Step { int d; int64_t k0, k1; };
        std::vector<Step> plan;
        for (int d = 1; d < R; ++d)
            for (int64_t k0 = 0; k0 < n; k0 += chunk) plan.push_back({d, k0, std::min(n, k0 + chunk)});
        // Z_r is complete on the compute stream before the first send reads it
        CUDA_TRY(c, cudaEventRecord(c->ev_free[0], c->stream));
        CUDA_TRY(c, cudaStreamWaitEvent(c->comm_stream, c->ev_free[0], 0));
        bool used[2] = {false, false};
        auto post = [&](size_t i) -> int {
            const Step& st = plan[i];
            int q = (int)(i & 1);
            int dst = (r - st.d + R) % R, src = (r + st.d) % R;
            int64_t ws = ranges[2 * src + 1] - ranges[2 * src], wd = ranges[2 * dst + 1] - ranges[2 * dst];
            if (used[q]) CUDA_TRY(c, cudaStreamWaitEvent(c->comm_stream, c->ev_free[q], 0));     // its last GEMM is done
            NCCL_TRY(c, nccl().GroupStart());
            if (nloc > 0 && wd > 0)
                NCCL_TRY(c, nccl().Send(Z + st.k0 * nloc, (size_t)(st.k1 - st.k0) * nloc, NCCL_F64, dst, (nccl_comm)c->comm, c->comm_stream));
            if (nloc > 0 && ws > 0)
                NCCL_TRY(c, nccl().Recv(buf[q], (size_t)(st.k1 - st.k0) * ws, NCCL_F64, src, (nccl_comm)c->comm, c->comm_stream));
            NCCL_TRY(c, nccl().GroupEnd());
            CUDA_TRY(c, cudaEventRecord(c->ev_recv[q], c->comm_stream));
            return SFEM_OK;
        };
        if (!plan.empty()) { int pr = post(0); if (pr) return pr; }
        for (size_t i = 0; i < plan.size(); ++i) {
            if (i + 1 < plan.size()) { int pr = post(i + 1); if (pr) return pr; }
            const Step& st = plan[i];
            int q = (int)(i & 1);
            int src = (r + st.d) % R;
            int64_t a = ranges[2 * src], ws = ranges[2 * src + 1] - a;
            CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_recv[q], 0));
            if (nloc > 0 && ws > 0 && failed == SFEM_OK && rc == SFEM_OK)
                rc = dgemm(c, OP_T, OP_N, (int)nloc, (int)ws, st.k1 - st.k0, 1.0, W + st.k0 * nloc, nloc, buf[q], ws,
                           st.k0 == 0 ? 0.0 : 1.0, res + a, n_obs, 0);
            CUDA_TRY(c, cudaEventRecord(c->ev_free[q], c->stream));
            used[q] = true;
        }
        // ---- gather the row blocks on rank 0 (Scatter.toZero, solving_utils.py:152-155)
        CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_recv[0], 0));
        CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_recv[1], 0));
        NCCL_TRY(c, nccl().GroupStart());
        if (r == 0) {
            for (int s = 1; s < R; ++s) {
                int64_t a = ranges[2 * s], ws = ranges[2 * s + 1] - a;
                if (ws > 0) NCCL_TRY(c, nccl().Recv(Cu_out + a * n_obs, (size_t)ws * n_obs, NCCL_F64, s, (nccl_comm)c->comm, c->stream));
            }
        } else if (nloc > 0) {
            NCCL_TRY(c, nccl().Send(res, (size_t)nloc * n_obs, NCCL_F64, 0, (nccl_comm)c->comm, c->stream));
        }
        NCCL_TRY(c, nccl().GroupEnd());
        CUDA_TRY(c, cudaStreamSynchronize(c->comm_stream));
        for (int q = 0; q < 2; ++q) sfem_dev_free(buf[q]);
    }
    if (r == 0 && nloc > 0)
        CUDA_TRY(c, cudaMemcpyAsync(Cu_out + col_begin * n_obs, res, sizeof(double) * (size_t)nloc * n_obs, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    sfem_dev_free(res);
    if (!Z_out) sfem_dev_free(Z);
    if (!W_out) sfem_dev_free(W);
    return failed != SFEM_OK ? failed : rc;
}
