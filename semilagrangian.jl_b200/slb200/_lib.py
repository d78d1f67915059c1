"""ctypes binding of libslb200.so (include/slb200.h).  The product path has no CPU
fallback: a missing library or a missing CUDA device raises immediately."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libslb200.so")

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)
c_void_pp = C.POINTER(C.c_void_p)

class SlbHalo(C.Structure):
    """struct slb_halo (include/slb200.h)"""
    _fields_ = [("mode", C.c_int), ("halo", C.c_int), ("shard_dim", C.c_int), ("push_lo", C.c_void_p), ("push_hi", C.c_void_p),
                ("err_flag", C.c_void_p)]


SLB_HALO_MARCH, SLB_HALO_PASSIVE = 1, 2

# every symbol include/slb200.h declares: (name, restype, argtypes)
SIGNATURES = [
    ("slb_ctx_create", C.c_int, [C.c_int, C.c_void_p, c_void_pp]),
    ("slb_ctx_destroy", None, [C.c_void_p]),
    ("slb_last_error", C.c_char_p, []),
    ("slb_sync", C.c_int, [C.c_void_p]),
    ("slb_device_count", C.c_int, []),
    ("slb_launch_count", C.c_int64, [C.c_void_p]),
    ("slb_timer_start", C.c_int, [C.c_void_p]),
    ("slb_timer_stop", C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    ("slb_event_create", C.c_int, [C.c_void_p, c_void_pp]),
    ("slb_event_record", C.c_int, [C.c_void_p, C.c_void_p]),
    ("slb_stream_wait_event", C.c_int, [C.c_void_p, C.c_void_p]),
    ("slb_event_elapsed_ms", C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]),
    ("slb_event_destroy", C.c_int, [C.c_void_p]),
    ("slb_malloc", C.c_int, [C.c_void_p, C.c_int64, c_void_pp]),
    ("slb_free", C.c_int, [C.c_void_p, C.c_void_p]),
    ("slb_host_alloc", C.c_int, [C.c_int64, c_void_pp]),
    ("slb_host_free", C.c_int, [C.c_void_p]),
    ("slb_memcpy_h2d", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    ("slb_memcpy_d2h", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    ("slb_grid_create", C.c_int, [C.c_void_p, C.c_int, c_int64_p, c_void_pp]),
    ("slb_grid_create_external", C.c_int, [C.c_void_p, C.c_int, c_int64_p, C.c_void_p, C.c_void_p, c_void_pp]),
    ("slb_grid_destroy", None, [C.c_void_p]),
    ("slb_grid_upload", C.c_int, [C.c_void_p, C.c_void_p]),
    ("slb_grid_download", C.c_int, [C.c_void_p, C.c_void_p]),
    ("slb_grid_front", C.c_void_p, [C.c_void_p]),
    ("slb_grid_back", C.c_void_p, [C.c_void_p]),
    ("slb_grid_swap", C.c_int, [C.c_void_p]),
    ("slb_interp_create", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, c_double_p, C.c_int, c_double_p, c_void_pp]),
    ("slb_interp_destroy", None, [C.c_void_p]),
    ("slb_sweep", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, c_int64_p, C.c_double, C.c_int, C.c_int]),
    ("slb_sweep_ex", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, c_int64_p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    ("slb_sweep_peer", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, c_int64_p, C.c_double, C.c_int, C.c_int, C.c_int, c_void_pp]),
    ("slb_sweep_pair", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, c_int64_p, C.c_double,
                                 C.c_int, C.c_void_p, C.c_void_p, C.c_int64, c_int64_p, C.c_double, C.c_int, C.c_int]),
    ("slb_sweep_pair_ex", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, c_int64_p, C.c_double,
                                    C.c_int, C.c_void_p, C.c_void_p, C.c_int64, c_int64_p, C.c_double, C.c_int, C.c_int,
                                    C.c_int, C.c_int, c_void_pp, C.c_int]),
    ("slb_sweep_pair_halo", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, c_int64_p, C.c_double,
                                      C.c_int, C.c_void_p, C.c_void_p, C.c_int64, c_int64_p, C.c_double, C.c_int, C.c_int,
                                      C.POINTER(SlbHalo)]),
    ("slb_halo_error", C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    ("slb_comm_create", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, c_void_pp]),
    ("slb_comm_destroy", None, [C.c_void_p]),
    ("slb_comm_export", C.c_int, [C.c_void_p, C.c_void_p]),
    ("slb_comm_connect", C.c_int, [C.c_void_p, C.c_void_p]),
    ("slb_comm_export_buffer", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("slb_comm_open_buffer", C.c_int, [C.c_void_p, C.c_void_p, c_void_pp]),
    ("slb_comm_close_buffer", C.c_int, [C.c_void_p, C.c_void_p]),
    ("slb_comm_barrier", C.c_int, [C.c_void_p]),
    ("slb_comm_signal", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    ("slb_comm_wait", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    ("slb_comm_allgather", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, c_void_pp]),
    ("slb_poisson_solve_partial", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p, c_void_pp]),
    ("slb_ipc_get_handle", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("slb_ipc_open_handle", C.c_int, [C.c_void_p, C.c_void_p, c_void_pp]),
    ("slb_ipc_close_handle", C.c_int, [C.c_void_p, C.c_void_p]),
    ("slb_presolve", C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    ("slb_charge_density", C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_void_p]),
    ("slb_charge_density_raw", C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_void_p]),
    ("slb_grid_set_linesum", C.c_int, [C.c_void_p, C.c_void_p]),
    ("slb_grid_set_rhopart", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    ("slb_grid_rhopart_planes", C.c_int64, [C.c_void_p]),
    ("slb_charge_density_from", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.c_void_p, C.c_int]),
    ("slb_subtract_mean", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    ("slb_poisson_create", C.c_int, [C.c_void_p, C.c_int, c_int64_p, C.POINTER(c_double_p), c_void_pp]),
    ("slb_poisson_destroy", None, [C.c_void_p]),
    ("slb_poisson_solve", C.c_int, [C.c_void_p, C.c_void_p, c_void_pp]),
    ("slb_vp_field_solve", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_void_p, c_void_pp]),
    ("slb_poisson_solve_raw", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, c_void_pp]),
    ("slb_reduce_sumsq", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, c_double_p]),
    ("slb_reduce_sumsq_async", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_void_p]),
    ("slb_capture_begin", C.c_int, [C.c_void_p]),
    ("slb_capture_end", C.c_int, [C.c_void_p, c_void_pp]),
    ("slb_graph_launch", C.c_int, [C.c_void_p]),
    ("slb_graph_destroy", None, [C.c_void_p]),
    ("slb_program_begin", C.c_int, [C.c_void_p]),
    ("slb_program_end", C.c_int, [C.c_void_p, c_void_pp]),
    ("slb_program_launch", C.c_int, [C.c_void_p, C.c_int, C.c_int64]),
    ("slb_program_info", C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ("slb_program_profile", C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_int), c_double_p, c_double_p]),
    ("slb_program_destroy", None, [C.c_void_p]),
    ("slb_reduce_sum", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, c_double_p]),
    ("slb_kinetic_energy", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_double, c_double_p]),
    ("slb_interp2d_points", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_int]),
    ("slb_fill_dec2d", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_double, C.c_void_p, C.c_double]),
    ("slb_lincomb", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, c_double_p, c_void_pp, C.c_int64]),
    ("slb_memcpy_d2d", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
]

SLB_SWEEP_EXACT = 1
SLB_SWEEP_INSIDE_EDGE = 2
SLB_E_UNSUPPORTED = -4
SLB_RESHARD_NONE, SLB_RESHARD_OUT_BLOCKED, SLB_RESHARD_IN_BLOCKED = 0, 1, 2


class SlbError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libslb200.so; raise (never fall back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SlbError(
                f"{LIB_PATH} is missing: build it with semilagrangian.jl_b200/build.sh "
                "(or __graft_entry__.build()).  There is no CPU fallback."
            )
        L = C.CDLL(LIB_PATH)
        for name, res, args in SIGNATURES:
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().slb_last_error().decode(errors="replace")
        if rc == -1:
            raise ValueError(msg)  # ArgumentError / DomainError of the reference
        raise SlbError(f"libslb200 error {rc}: {msg}")


def i64(seq):
    return (C.c_int64 * len(seq))(*[int(x) for x in seq])


def dptr(a):
    assert isinstance(a, np.ndarray) and a.dtype == np.float64
    return a.ctypes.data_as(c_double_p)


class Context:
    """One per process and GPU (slb_ctx)."""

    LEGACY_DEFAULT_STREAM = 1  # cudaStreamLegacy: adopt the default stream explicitly
    _next_serial = 1

    def __init__(self, device=0, stream=None):
        """stream: None -> the library creates a private non-blocking stream; an integer
        cudaStream_t handle -> adopt it (0, torch's default stream, is passed as cudaStreamLegacy)."""
        h = C.c_void_p()
        if stream is not None and int(stream) == 0:
            stream = self.LEGACY_DEFAULT_STREAM
        check(lib().slb_ctx_create(int(device), C.c_void_p(int(stream)) if stream is not None else None, C.byref(h)))
        self.h = h
        self.device = device
        # caches elsewhere (interpolation handles, pooled device fields) are keyed by this serial number, never by
        # id(ctx): ids are reused once a Context has been collected.  Cached handles register a release callback here.
        self.serial = Context._next_serial
        Context._next_serial += 1
        self._on_close = []

    def sync(self):
        check(lib().slb_sync(self.h))

    def launch_count(self):
        return int(lib().slb_launch_count(self.h))

    def timer_start(self):
        check(lib().slb_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        check(lib().slb_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def event(self):
        e = C.c_void_p()
        check(lib().slb_event_create(self.h, C.byref(e)))
        return e

    def record(self, ev):
        check(lib().slb_event_record(self.h, ev))

    def wait_event(self, ev):
        """work enqueued on this context's stream from now on waits for `ev` (recorded on another context)"""
        check(lib().slb_stream_wait_event(self.h, ev))

    @staticmethod
    def elapsed_ms(e0, e1):
        ms = C.c_float()
        check(lib().slb_event_elapsed_ms(e0, e1, C.byref(ms)))
        return ms.value

    def malloc(self, nbytes):
        p = C.c_void_p()
        check(lib().slb_malloc(self.h, int(nbytes), C.byref(p)))
        return p

    def free(self, p):
        check(lib().slb_free(self.h, p))

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        p = self.malloc(arr.nbytes)
        check(lib().slb_memcpy_h2d(self.h, p, arr.ctypes.data_as(C.c_void_p), arr.nbytes))
        self.sync()
        return p

    def to_host(self, p, n):
        out = np.empty(int(n), dtype=np.float64)
        check(lib().slb_memcpy_d2h(self.h, out.ctypes.data_as(C.c_void_p), p, out.nbytes))
        self.sync()
        return out

    def on_close(self, fn):
        self._on_close.append(fn)

    def close(self):
        if self.h:
            for fn in reversed(self._on_close):
                try:
                    fn()
                except Exception:
                    pass
            self._on_close = []
            lib().slb_ctx_destroy(self.h)
            self.h = None


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        dev = int(os.environ.get("LOCAL_RANK", "0"))
        n = lib().slb_device_count()
        if n == 0:
            raise SlbError("no CUDA device visible: libslb200 has no CPU fallback")
        _default_ctx = Context(dev % n)
    return _default_ctx


def pinned_empty(shape, order="F"):
    """float64 numpy array backed by pinned host memory (cudaHostAlloc)."""
    n = int(np.prod(shape))
    p = C.c_void_p()
    check(lib().slb_host_alloc(n * 8, C.byref(p)))
    buf = (C.c_double * n).from_address(p.value)
    arr = np.frombuffer(buf, dtype=np.float64).reshape(shape, order=order)
    return arr, p
