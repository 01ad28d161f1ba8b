// Context lifecycle, scratch arena, exclusive scans and cell lists (K6) shared by gcov.cu and mg.cu.
#include "common.cuh"
#include <cstdarg>

int sfem_fail(sfem_ctx* c, int code, const char* fmt, ...) {
    if (c) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(c->err, sizeof(c->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

void prof_collect(sfem_ctx* c) {
    if (c->prof_pairs.empty()) return;
    cudaStreamSynchronize(c->stream);
    for (auto& p : c->prof_pairs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) c->prof_ms[p.tag] += ms;
        c->prof_pool.push_back(p.e0);
        c->prof_pool.push_back(p.e1);
    }
    c->prof_pairs.clear();
}

void prof_reset(sfem_ctx* c) {
    prof_collect(c);
    for (int t = 0; t < 32; ++t) { c->prof_ms[t] = 0; c->prof_bytes[t] = 0; c->prof_flops[t] = 0; c->prof_count[t] = 0; }
}

void prof_free(sfem_ctx* c) {
    prof_collect(c);
    for (cudaEvent_t e : c->prof_pool) cudaEventDestroy(e);
    c->prof_pool.clear();
}

void* sfem_scratch(sfem_ctx* c, size_t bytes) {
    if (bytes <= c->scratch_bytes && c->scratch) return c->scratch;
    if (c->scratch) {
        cudaStreamSynchronize(c->stream);
        sfem_dev_free(c->scratch);
        c->scratch = nullptr;
        c->scratch_bytes = 0;
    }
    size_t want = bytes + bytes / 4 + 4096;
    if (sfem_dev_malloc(&c->scratch, want) != cudaSuccess) {
        cudaGetLastError();
        if (sfem_dev_malloc(&c->scratch, bytes) != cudaSuccess) {
            cudaGetLastError();
            c->scratch = nullptr;
            return nullptr;
        }
        want = bytes;
    }
    c->scratch_bytes = want;
    return c->scratch;
}

// ------------------------------------------------------------------ exclusive scan (3-phase, deterministic)
namespace {
constexpr int SCAN_B = 1024;

template <typename T>
__global__ void scan_block_sums(const T* __restrict__ in, int64_t n, T* __restrict__ sums) {
    __shared__ T sh[32];
    int64_t i = (int64_t)blockIdx.x * SCAN_B + threadIdx.x;
    T v = i < n ? in[i] : T(0);
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        T w = sh[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) w += __shfl_down_sync(0xffffffffu, w, o);
        if (threadIdx.x == 0) sums[blockIdx.x] = w;
    }
}

// single block: exclusive scan of `m` values in place (serial over chunks of 1024), total to sums[m]
template <typename T>
__global__ void scan_sums(T* __restrict__ sums, int64_t m) {
    __shared__ T sh[SCAN_B];
    __shared__ T carry;
    if (threadIdx.x == 0) carry = T(0);
    __syncthreads();
    for (int64_t base = 0; base < m; base += SCAN_B) {
        int64_t i = base + threadIdx.x;
        T v = i < m ? sums[i] : T(0);
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < SCAN_B; o <<= 1) {
            T t = threadIdx.x >= o ? sh[threadIdx.x - o] : T(0);
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        T incl = sh[threadIdx.x];
        T c0 = carry;
        if (i < m) sums[i] = c0 + incl - v;
        __syncthreads();
        if (threadIdx.x == SCAN_B - 1) carry = c0 + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[m] = carry;
}

template <typename T>
__global__ void scan_apply(const T* __restrict__ in, int64_t n, const T* __restrict__ sums, T* __restrict__ out) {
    __shared__ T sh[SCAN_B];
    int64_t i = (int64_t)blockIdx.x * SCAN_B + threadIdx.x;
    T v = i < n ? in[i] : T(0);
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < SCAN_B; o <<= 1) {
        T t = threadIdx.x >= o ? sh[threadIdx.x - o] : T(0);
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    if (i < n) out[i] = sums[blockIdx.x] + sh[threadIdx.x] - v;
    if (i == n - 1) out[n] = sums[blockIdx.x] + sh[threadIdx.x];
}
}  // namespace

// out[0..n] = exclusive scan of in[0..n-1] (out has n+1 entries; in and out may alias only if identical types
// and out == in is NOT allowed).  `sums` scratch: (ceil(n/1024)+1) elements.
template <typename T>
int exclusive_scan(sfem_ctx* c, const T* in, int64_t n, T* out, T* sums) {
    if (n <= 0) {
        CUDA_TRY(c, cudaMemsetAsync(out, 0, sizeof(T), c->stream));
        return SFEM_OK;
    }
    int64_t nb = cdiv64(n, SCAN_B);
    scan_block_sums<T><<<(unsigned)nb, SCAN_B, 0, c->stream>>>(in, n, sums);
    KCHECK(c);
    scan_sums<T><<<1, SCAN_B, 0, c->stream>>>(sums, nb);
    KCHECK(c);
    scan_apply<T><<<(unsigned)nb, SCAN_B, 0, c->stream>>>(in, n, sums, out);
    KCHECK(c);
    return SFEM_OK;
}
template int exclusive_scan<int>(sfem_ctx*, const int*, int64_t, int*, int*);
template int exclusive_scan<int64_t>(sfem_ctx*, const int64_t*, int64_t, int64_t*, int64_t*);

// ------------------------------------------------------------------ cell list (K6)
namespace {
__global__ void cell_count_kernel(const int* __restrict__ cell_of, int64_t n, int* __restrict__ counts) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&counts[cell_of[i]], 1);
}
__global__ void cell_scatter_kernel(const int* __restrict__ cell_of, int64_t n, const int* __restrict__ start,
                                    int* __restrict__ cursor, int* __restrict__ nodes) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        int cidx = cell_of[i];
        int pos = atomicAdd(&cursor[cidx], 1);
        nodes[start[cidx] + pos] = (int)i;
    }
}
// make the order inside every cell ascending (the atomic scatter order is arbitrary): insertion sort, cells are small
__global__ void cell_sort_kernel(const int* __restrict__ start, int ncell, int* __restrict__ nodes) {
    int cidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (cidx >= ncell) return;
    int s = start[cidx], e = start[cidx + 1];
    for (int i = s + 1; i < e; ++i) {
        int v = nodes[i];
        int j = i - 1;
        while (j >= s && nodes[j] > v) {
            nodes[j + 1] = nodes[j];
            --j;
        }
        nodes[j + 1] = v;
    }
}
// one big cell (all-pairs mode): identity list
__global__ void iota_kernel(int* __restrict__ nodes, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) nodes[i] = (int)i;
}
}  // namespace

// Build cell_start[ncell+1], cell_nodes[n] (both cudaMalloc'd, returned through the out pointers).
int build_cell_list(sfem_ctx* c, const int* cell_of, int64_t n, int ncell, int** cell_start_out, int** cell_nodes_out) {
    int *start = nullptr, *nodes = nullptr;
    CUDA_TRY(c, sfem_dev_malloc(&start, sizeof(int) * ((size_t)ncell + 1)));
    CUDA_TRY(c, sfem_dev_malloc(&nodes, sizeof(int) * (size_t)(n > 0 ? n : 1)));
    *cell_start_out = start;
    *cell_nodes_out = nodes;
    if (ncell == 1) {
        int se[2] = {0, (int)n};
        CUDA_TRY(c, cudaMemcpyAsync(start, se, sizeof(se), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));   // `se` is a stack buffer
        iota_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, c->stream>>>(nodes, n);
        KCHECK(c);
        return SFEM_OK;
    }
    size_t nsum = (size_t)cdiv64(ncell, SCAN_B) + 1;
    int* tmp = (int*)sfem_scratch(c, sizeof(int) * ((size_t)ncell * 2 + nsum + 16));
    if (!tmp) return sfem_fail(c, SFEM_ERR_CUDA, "cell list: scratch allocation failed");
    int* counts = tmp;
    int* cursor = tmp + ncell;
    int* sums = tmp + 2 * (size_t)ncell;
    CUDA_TRY(c, cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)ncell * 2, c->stream));
    cell_count_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, c->stream>>>(cell_of, n, counts);
    KCHECK(c);
    int rc = exclusive_scan<int>(c, counts, ncell, start, sums);
    if (rc) return rc;
    cell_scatter_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, c->stream>>>(cell_of, n, start, cursor, nodes);
    KCHECK(c);
    cell_sort_kernel<<<(unsigned)cdiv64(ncell, 128), 128, 0, c->stream>>>(start, ncell, nodes);
    KCHECK(c);
    return SFEM_OK;
}
