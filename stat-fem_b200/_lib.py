"""ctypes binding of libstatfem_b200.so (include/statfem_b200.h) + device-memory plumbing through PyTorch.

There is no CPU fallback: importing this module without the built library, or creating a context without a B200,
raises.  PyTorch is used only to own device memory (``torch.empty(..., device='cuda')``), the stream, and (in
parallel.py) ``torch.distributed``; every number is produced by the kernels in csrc/.
"""
import ctypes as C
import os
import numpy as np
from scipy.linalg import LinAlgError

_HERE = os.path.dirname(os.path.abspath(__file__))
# SFEM_LIB_PATH: a diagnostic build of the same library (e.g. with -DSFEM_PANEL_CLOCKS, tools/panel_clocks.py)
LIB_PATH = os.environ.get("SFEM_LIB_PATH") or os.path.join(_HERE, "libstatfem_b200.so")

SFEM_ERR_CUDA, SFEM_ERR_ARG, SFEM_ERR_NOT_SPD, SFEM_ERR_NOT_CONVERGED, SFEM_ERR_STATE, SFEM_ERR_LIMIT = -1, -2, -3, -4, -5, -6
OP_N, OP_T = 0, 1

EXPORTS = [
    "sfem_version", "sfem_create", "sfem_destroy", "sfem_last_error", "sfem_launch_count",
    "sfem_set_stiffness", "sfem_mg_setup", "sfem_mg_num_levels", "sfem_mg_level_size",
    "sfem_solve_workspace_bytes", "sfem_solve_block", "sfem_spmm_csr", "sfem_scatter_columns",
    "sfem_gcov_count", "sfem_gcov_fill", "sfem_dgemm", "sfem_potrf_dinv_doubles", "sfem_potrf", "sfem_potrs",
    "sfem_potri", "sfem_logdet", "sfem_kernel_matrix", "sfem_dense_workspace_bytes", "sfem_logpost",
    "sfem_posterior_cov", "sfem_dense_solve", "sfem_cov_matrix", "sfem_axpby", "sfem_profile_enable", "sfem_profile_num_tags",
    "sfem_profile_tag_name", "sfem_profile_read", "sfem_csr_unsym_count", "sfem_csr_unsym_fill", "sfem_rows_axpy",
    "sfem_rows_gather", "sfem_dgemm_lower", "sfem_mirror_lower", "sfem_tile_apply_one",
    "sfem_b84_build", "sfem_b84_spmm", "sfem_b84_free", "sfem_skinny_gemm", "sfem_sym_fold",
    "sfem_nccl_unique_id", "sfem_comm_init", "sfem_comm_attach", "sfem_comm_destroy", "sfem_comm_bcast", "sfem_prior_cov",
]


class NotConvergedError(RuntimeError):
    pass


def load_library():
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libstatfem_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "or `make -C stat-fem_b200/csrc`); stat-fem_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_size_t
    sig = {
        "sfem_version": (i32, []),
        "sfem_create": (i32, [C.POINTER(vp), i32, vp]),
        "sfem_destroy": (i32, [vp]),
        "sfem_last_error": (C.c_char_p, [vp]),
        "sfem_launch_count": (C.c_longlong, [vp]),
        "sfem_set_stiffness": (i32, [vp, i64, vp, vp, vp]),
        "sfem_mg_setup": (i32, [vp, i32, vp, vp, C.POINTER(f64), C.POINTER(f64), C.POINTER(i32), i32, f64, f64, i32]),
        "sfem_mg_num_levels": (i32, [vp]),
        "sfem_mg_level_size": (i64, [vp, i32]),
        "sfem_solve_workspace_bytes": (sz, [vp, i32, i32]),
        "sfem_solve_block": (i32, [vp, i32, vp, i64, vp, i64, i32, f64, f64, i32, C.POINTER(i32), C.POINTER(f64), vp, sz]),
        "sfem_spmm_csr": (i32, [vp, i64, i64, vp, vp, vp, i32, vp, i64, vp, i64]),
        "sfem_profile_enable": (i32, [vp, C.c_uint]),
        "sfem_profile_num_tags": (i32, []),
        "sfem_profile_tag_name": (C.c_char_p, [i32]),
        "sfem_profile_read": (i32, [vp, C.POINTER(f64), C.POINTER(f64), C.POINTER(f64), C.POINTER(C.c_longlong)]),
        "sfem_scatter_columns": (i32, [vp, i64, vp, vp, vp, i64, i32, vp, i64]),
        "sfem_gcov_count": (i32, [vp, i64, i32, vp, vp, f64, f64, f64, f64, f64, f64, C.POINTER(f64), C.POINTER(f64), vp,
                                  C.POINTER(i64), C.POINTER(i64), vp, i32, C.POINTER(i64)]),
        "sfem_gcov_fill": (i32, [vp, vp, vp, vp]),
        "sfem_dgemm": (i32, [vp, i32, i32, i32, i32, i64, f64, vp, i64, vp, i64, f64, vp, i64]),
        "sfem_potrf_dinv_doubles": (sz, [i32]),
        "sfem_potrf": (i32, [vp, i32, vp, i64, vp, C.POINTER(i32)]),
        "sfem_potrs": (i32, [vp, i32, vp, i64, vp, i32, vp, i64]),
        "sfem_potri": (i32, [vp, i32, vp, i64, vp, vp, vp]),
        "sfem_logdet": (i32, [vp, i32, vp, i64, C.POINTER(f64)]),
        "sfem_kernel_matrix": (i32, [vp, i32, i32, vp, f64, f64, vp, i32, f64, vp, i32, vp]),
        "sfem_dense_workspace_bytes": (sz, [i32]),
        "sfem_logpost": (i32, [vp, i32, i32, vp, vp, vp, i32, vp, vp, C.POINTER(f64), C.POINTER(f64), C.POINTER(f64), vp, sz]),
        "sfem_posterior_cov": (i32, [vp, i32, i32, vp, vp, vp, i32, vp, vp, C.POINTER(f64), vp, vp, vp, sz]),
        "sfem_cov_matrix": (i32, [vp, i32, i32, i32, vp, vp, f64, f64, i32, vp]),
        "sfem_axpby": (i32, [vp, i64, f64, vp, f64, vp, vp]),
        "sfem_csr_unsym_count": (i32, [vp, i64, vp, vp, vp, vp, vp, vp, C.POINTER(i64), C.POINTER(i64)]),
        "sfem_csr_unsym_fill": (i32, [vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "sfem_rows_axpy": (i32, [vp, i64, vp, i32, f64, vp, i64, vp, i64]),
        "sfem_rows_gather": (i32, [vp, i64, vp, i32, vp, i64, vp, i64]),
        "sfem_dgemm_lower": (i32, [vp, i32, i32, i32, i64, f64, vp, i64, vp, i64, f64, vp, i64]),
        "sfem_mirror_lower": (i32, [vp, i32, vp, i64]),
        "sfem_b84_build": (i32, [vp, i64, vp, vp, vp, C.POINTER(vp), C.POINTER(i64)]),
        "sfem_b84_spmm": (i32, [vp, vp, i32, vp, i64, vp, i64]),
        "sfem_b84_free": (i32, [vp, vp]),
        "sfem_tile_apply_one": (i32, [vp, i32, i32, i32, i32, vp, vp, vp, vp, C.POINTER(i32), i32, C.POINTER(i64), C.POINTER(i64),
                                      C.POINTER(i32)]),
        "sfem_nccl_unique_id": (i32, [vp]),
        "sfem_comm_init": (i32, [vp, vp, i32, i32]),
        "sfem_comm_attach": (i32, [vp, vp, i32, i32]),
        "sfem_comm_destroy": (i32, [vp]),
        "sfem_comm_bcast": (i32, [vp, vp, sz, i32]),
        "sfem_prior_cov": (i32, [vp, i64, vp, vp, vp, vp, vp, vp, i64, i64, i32, i32, f64, f64, i32, vp, vp, vp, C.POINTER(i32)]),
        "sfem_sym_fold": (i32, [vp, i32, i32, vp, i64]),
        "sfem_skinny_gemm": (i32, [vp, i64, i32, i32, vp, i64, vp, i64, vp, i64]),
        "sfem_dense_solve": (i32, [vp, i32, i32, vp, vp, i32, f64, f64, f64, vp, vp, vp, vp, sz]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None
_ctx = {}
_retired_launches = [0]
_live = []
USE_BLOCK_FORM = [os.environ.get("SFEM_BLOCK_FORM", "1") != "0"]     # W = G Z through the DMMA block form (csrc/bsr.cu)
DEVICE_CACHE = [False]     # bench.py's device-resident arm: keep uploaded inputs (A, coords, b, P) in HBM between steps


def lib():
    global _lib
    if _lib is None:
        _lib = load_library()
    return _lib


def torch_mod():
    import torch
    return torch


class Context(object):
    """One sfem_ctx per CUDA device, bound to torch's current stream on that device."""

    def __init__(self, device=None, stream=None):
        torch = torch_mod()
        if not torch.cuda.is_available():
            raise RuntimeError("stat-fem_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.tdev = torch.device("cuda", self.device)
        self.handle = C.c_void_p()
        self.tstream = stream            # a torch.cuda.Stream this context is bound to (None: the stream current now)
        stream = (stream if stream is not None else torch.cuda.current_stream(self.tdev)).cuda_stream
        rc = lib().sfem_create(C.byref(self.handle), self.device, C.c_void_p(stream))
        if rc != 0:
            raise RuntimeError("sfem_create failed (rc=%d): %s" % (rc, lib().sfem_last_error(None).decode()))
        self._keep = {}
        _live.append(self)

    def check(self, rc):
        if rc == 0:
            return
        msg = lib().sfem_last_error(self.handle).decode()
        if rc == SFEM_ERR_NOT_SPD:
            raise LinAlgError(msg)
        if rc == SFEM_ERR_NOT_CONVERGED:
            raise NotConvergedError(msg)
        if rc == SFEM_ERR_ARG:
            raise ValueError(msg)
        raise RuntimeError("libstatfem_b200 error %d: %s" % (rc, msg))

    def launch_count(self):
        return int(lib().sfem_launch_count(self.handle))

    def profile_enable(self, mask=0xffffffff):
        self.check(lib().sfem_profile_enable(self.handle, mask))

    def profile_read(self):
        "dict tag-name -> dict(ms, bytes, flops, count) accumulated since profile_enable"
        nt = lib().sfem_profile_num_tags()
        ms, by, fl = (C.c_double * nt)(), (C.c_double * nt)(), (C.c_double * nt)()
        cnt = (C.c_longlong * nt)()
        self.check(lib().sfem_profile_read(self.handle, ms, by, fl, cnt))
        return {lib().sfem_profile_tag_name(t).decode(): dict(ms=ms[t], bytes=by[t], flops=fl[t], count=int(cnt[t]))
                for t in range(nt)}

    # ---- device memory helpers ------------------------------------------------------------------------
    def to_device(self, arr, dtype):
        torch = torch_mod()
        a = np.ascontiguousarray(arr, dtype=dtype)
        return torch.from_numpy(a).to(self.tdev)

    def empty(self, shape, dtype=None):
        torch = torch_mod()
        return torch.empty(shape, dtype=dtype or torch.float64, device=self.tdev)

    def zeros(self, shape, dtype=None):
        torch = torch_mod()
        return torch.zeros(shape, dtype=dtype or torch.float64, device=self.tdev)

    def workspace(self, nbytes):
        torch = torch_mod()
        return torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=self.tdev)

    def sync(self):
        torch_mod().cuda.synchronize(self.tdev)


def ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def get_context(device=None):
    torch = torch_mod()
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("stat-fem_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        device = torch.cuda.current_device()
    if device not in _ctx:
        _ctx[device] = Context(device)
    return _ctx[device]


class DeviceCSR(object):
    """CSR (or, read the other way, CSC) matrix resident in HBM: int64 pointers, int32 indices, float64 values."""

    def __init__(self, ctx, shape, indptr, indices, data, on_device=False):
        torch = torch_mod()
        self.ctx = ctx
        self.shape = tuple(int(s) for s in shape)
        if on_device:
            self.indptr, self.indices, self.data = indptr, indices, data
        else:
            self.indptr = ctx.to_device(indptr, np.int64)
            self.indices = ctx.to_device(indices, np.int32)
            self.data = ctx.to_device(data, np.float64)
        assert self.indptr.dtype == torch.int64 and self.indices.dtype == torch.int32 and self.data.dtype == torch.float64
        self.nnz = int(self.indices.shape[0])

    # ---- FP64 tensor-core form (csrc/bsr.cu): worth its set-up for wide blocks on matrices with long rows (G)
    BLOCK_FORM_MIN_COLS = 256
    BLOCK_FORM_MIN_ROW_NNZ = 16

    def block_form(self):
        "handle of the 8x4-block form, built on first use"
        if self.__dict__.get("_b84") is None:
            h, nt = C.c_void_p(), C.c_int64(0)
            self.ctx.check(lib().sfem_b84_build(self.ctx.handle, self.shape[0], ptr(self.indptr), ptr(self.indices), ptr(self.data),
                                                C.byref(h), C.byref(nt)))
            self._b84, self.block_tiles = h, int(nt.value)
        return self._b84

    def release_block_form(self):
        h = self.__dict__.get("_b84")
        if h is not None and self.ctx.handle:
            lib().sfem_b84_free(self.ctx.handle, h)
        self._b84 = None

    def __del__(self):
        try:
            self.release_block_form()
        except Exception:
            pass

    def matmul(self, X, out=None):
        "Y = S X for a device block X (nrows_in x k) or vector"
        ctx = self.ctx
        vec = (X.dim() == 1)
        X2 = X.reshape(-1, 1) if vec else X
        assert X2.shape[0] == self.shape[1] and X2.stride(1) == 1
        k = int(X2.shape[1])
        Y = out if out is not None else ctx.empty((self.shape[0], k))
        Y2 = Y.reshape(-1, 1) if Y.dim() == 1 else Y
        if (USE_BLOCK_FORM[0] and k >= self.BLOCK_FORM_MIN_COLS and self.shape[0] == self.shape[1]
                and self.nnz >= self.BLOCK_FORM_MIN_ROW_NNZ * self.shape[0] and Y2.stride(0) % 2 == 0 and Y2.data_ptr() % 16 == 0):
            ctx.check(lib().sfem_b84_spmm(ctx.handle, self.block_form(), k, ptr(X2), int(X2.stride(0)), ptr(Y2), int(Y2.stride(0))))
            return Y
        ctx.check(lib().sfem_spmm_csr(ctx.handle, self.shape[0], self.nnz, ptr(self.indptr), ptr(self.indices), ptr(self.data), k,
                                      ptr(X2), int(X2.stride(0)), ptr(Y2), int(Y2.stride(0))))
        return Y.reshape(-1) if vec and out is None else Y

    def to_host(self):
        return (self.indptr.cpu().numpy(), self.indices.cpu().numpy(), self.data.cpu().numpy())


def new_context(device=None):
    "a private context (own stiffness matrix + multigrid hierarchy), e.g. one per SparseSolver"
    return Context(device)


_stream_pool = {}


def acquire_stream_context(device=None):
    """a stream context from a per-device pool (created on demand).  Pooling keeps the CUDA streams, the library's
    temporaries and the torch allocations made on those streams alive between uses: torch's caching allocator keeps one
    pool of blocks per stream, so fresh streams every time would mean fresh cudaMallocs (and synchronising cudaFrees
    under memory pressure) every time."""
    torch = torch_mod()
    dev = torch.cuda.current_device() if device is None else int(device)
    free = _stream_pool.setdefault(dev, [])
    return free.pop() if free else new_stream_context(dev)


def release_stream_context(ctx):
    if ctx is not None and ctx.handle:
        _stream_pool.setdefault(ctx.device, []).append(ctx)


def new_stream_context(device=None):
    """a private context on its OWN CUDA stream: independent dense work (e.g. the MAP fits of several data snapshots)
    submitted from different host threads overlaps on the GPU.  Callers make that stream current in their thread
    (``torch.cuda.stream(ctx.tstream)``) so that torch allocations and copies are ordered with the library's kernels."""
    torch = torch_mod()
    dev = torch.cuda.current_device() if device is None else int(device)
    # The private stream gets a high priority, so that the latency chain of a factorisation (panel kernel, panel solve,
    # in-block update) is scheduled ahead of the bulk trailing updates that the fits queue on their default-priority
    # side streams: 8 concurrent fits at 4096 sensors 1.72 -> 1.65 s on the same box with identical evaluation counts
    # (gpurun_out/r2b_t10_prio*.err; SFEM_FIT_PRIORITY=0 switches it off)
    prio = -1 if os.environ.get("SFEM_FIT_PRIORITY", "1") == "1" else 0
    return Context(dev, stream=torch.cuda.Stream(device=dev, priority=prio))


def destroy_context(ctx):
    if ctx is not None and ctx.handle:
        _retired_launches[0] += ctx.launch_count()
        lib().sfem_destroy(ctx.handle)
        ctx.handle = C.c_void_p()
        if ctx in _live:
            _live.remove(ctx)


def total_launches():
    "kernels launched by this library in this process (all contexts, live or destroyed)"
    return _retired_launches[0] + sum(c.launch_count() for c in _live if c.handle)


def cached_upload(holder, key, ctx, build):
    """device copy of a host input; kept on ``holder`` only while DEVICE_CACHE is on"""
    if not DEVICE_CACHE[0]:
        return build()
    cache = holder.__dict__.setdefault("_dev_cache", {})
    k = (key, ctx.device)
    if k not in cache:
        cache[k] = build()
    return cache[k]
