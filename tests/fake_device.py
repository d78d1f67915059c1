"""TEST DOUBLE (tests/ only) of the part of libslb200's C ABI that the unsplit 2-D host layer calls,
so that the HOST sequencing logic of slb200/unsplit2d.py (Adams-Bashforth start-up procedures,
buffer rotation, provider protocol) can be exercised by `-m "not gpu"` tests in a container without
a GPU.  "Device" memory is host memory; the per-point interpolation runs the CUDA kernel's own
per-thread body compiled for the host (lib/libslb200_hosttest.so: slbt_points_host,
slbt_bspline_solve_host); the small array operations are restated in numpy.

This is not a fallback: it is installed by a pytest fixture through monkeypatching and nothing in
the product can reach it.  The GPU tests (tests/test_gpu_unsplit2d.py) run the real library."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_HOSTTEST = os.path.join(ROOT, "semilagrangian.jl_b200", "lib", "libslb200_hosttest.so")
dp = C.POINTER(C.c_double)


def _addr(x):
    if x is None:
        return 0
    if isinstance(x, int):
        return x
    if isinstance(x, C.c_void_p):
        return x.value or 0
    if isinstance(x, C.Array):
        return C.addressof(x)
    return C.cast(x, C.c_void_p).value or 0


def _arr(addr, n):
    return np.ctypeslib.as_array((C.c_double * int(n)).from_address(_addr(addr)))


def _set(byref_obj, value):
    byref_obj._obj.value = value


class FakeLib:
    def __init__(self):
        self.host = C.CDLL(_HOSTTEST)
        self.host.slbt_points_host.argtypes = [dp, dp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, dp, C.c_int, C.c_int, dp, C.c_int, C.c_int]
        self.host.slbt_bspline_solve_host.argtypes = [C.c_int, C.c_longlong, dp, dp, dp]
        self.mem = {}      # address -> numpy buffer (keeps it alive)
        self.grids = {}    # id -> dict
        self.interps = {}  # id -> dict
        self.plans = {}
        self.next_id = 1000
        self.calls = []

    def _new_id(self):
        self.next_id += 8
        return self.next_id

    def _alloc(self, nbytes):
        buf = np.zeros(max(int(nbytes) // 8, 1))
        self.mem[buf.ctypes.data] = buf
        return buf.ctypes.data

    # -- context / memory ------------------------------------------------------------------
    def slb_device_count(self):
        return 1

    def slb_last_error(self):
        return b"fake device"

    def slb_ctx_create(self, dev, stream, out):
        _set(out, self._new_id())
        return 0

    def slb_ctx_destroy(self, h):
        return None

    def slb_sync(self, h):
        return 0

    def slb_launch_count(self, h):
        return len(self.calls)

    def slb_malloc(self, ctx, nbytes, out):
        _set(out, self._alloc(nbytes))
        return 0

    def slb_free(self, ctx, p):
        self.mem.pop(_addr(p), None)
        return 0

    def slb_memcpy_h2d(self, ctx, dst, src, nbytes):
        _arr(dst, nbytes // 8)[:] = _arr(src, nbytes // 8)
        return 0

    slb_memcpy_d2h = slb_memcpy_h2d
    slb_memcpy_d2d = slb_memcpy_h2d

    # -- grid ------------------------------------------------------------------------------
    def slb_grid_create(self, ctx, nd, ext, out):
        ext = [int(ext[i]) for i in range(nd)]
        numel = int(np.prod(ext))
        gid = self._new_id()
        self.grids[gid] = {"ext": ext, "numel": numel, "front": self._alloc(numel * 8), "back": self._alloc(numel * 8)}
        _set(out, gid)
        return 0

    def _g(self, g):
        return self.grids[_addr(g)]

    def slb_grid_destroy(self, g):
        self.grids.pop(_addr(g), None)

    def slb_grid_upload(self, g, host):
        gr = self._g(g)
        _arr(gr["front"], gr["numel"])[:] = _arr(host, gr["numel"])
        return 0

    def slb_grid_download(self, g, host):
        gr = self._g(g)
        _arr(host, gr["numel"])[:] = _arr(gr["front"], gr["numel"])
        return 0

    def slb_grid_front(self, g):
        return self._g(g)["front"]

    def slb_grid_back(self, g):
        return self._g(g)["back"]

    def slb_grid_swap(self, g):
        gr = self._g(g)
        gr["front"], gr["back"] = gr["back"], gr["front"]
        return 0

    def slb_grid_set_rhopart(self, g, p, cap):
        return 0

    def slb_grid_rhopart_planes(self, g):
        return 0  # the double never fuses pairs: no partial planes

    def slb_grid_set_linesum(self, g, p):
        self._g(g)["linesum"] = _addr(p)
        return 0

    # -- interpolation objects -------------------------------------------------------------
    def slb_interp_create(self, ctx, kind, order, n, coef, nc, nodes, out):
        iid = self._new_id()
        self.interps[iid] = {
            "kind": kind, "order": order, "n": int(n), "nc": nc,
            "coef": np.array(_arr(coef, (order + 1) * nc), copy=True),
            "nodes": np.array(_arr(nodes, order), copy=True) if _addr(nodes) else None,
        }
        _set(out, iid)
        return 0

    def slb_interp_destroy(self, h):
        return None

    def _presolve(self, it, a, axis):
        """sol(interp, .) along `axis` of the [n1, n2, ncomp] array a (in place)"""
        n = a.shape[axis]
        assert it["n"] == n
        mv = np.moveaxis(a, axis, 0)
        flat = mv.reshape(n, -1)
        for k in range(flat.shape[1]):
            b = np.ascontiguousarray(flat[:, k])
            x = np.empty(n)
            rc = self.host.slbt_bspline_solve_host(it["order"], n, it["nodes"].ctypes.data_as(dp), b.ctypes.data_as(dp), x.ctypes.data_as(dp))
            assert rc == 0
            flat[:, k] = x
        mv[...] = flat.reshape(mv.shape)

    def slb_interp2d_points(self, ctx, h1, h2, n1, n2, ncomp, inp, dec, out, work, flags):
        self.calls.append("interp2d_points")
        it1, it2 = self.interps[_addr(h1)], self.interps[_addr(h2)]
        assert _addr(inp) != _addr(out)
        numel = n1 * n2 * ncomp
        res = np.array(_arr(inp, numel), copy=True).reshape((n1, n2, ncomp), order="F")
        if it1["kind"] in (1, 2):
            assert _addr(work)
            self._presolve(it1, res, 0)
        if it2["kind"] in (1, 2):
            assert _addr(work)
            self._presolve(it2, res, 1)
        if it1["kind"] in (1, 2) or it2["kind"] in (1, 2):
            _arr(inp, numel)[:] = np.nan  # the real library overwrites its input: callers must not rely on it
            _arr(work, numel)[:] = np.nan
        res = np.asfortranarray(res)
        o = np.empty(numel)
        d = np.array(_arr(dec, n1 * n2 * 2), copy=True)
        templ = 1 if (it1["order"] == it2["order"] and it1["order"] + 1 <= 14 and it1["nc"] <= 14 and it2["nc"] <= 14) else 0
        rc = self.host.slbt_points_host(res.ctypes.data_as(dp), d.ctypes.data_as(dp), o.ctypes.data_as(dp), n1, n2, ncomp,
                                        it1["order"] + 1, it1["nc"], it1["coef"].ctypes.data_as(dp),
                                        it2["order"] + 1, it2["nc"], it2["coef"].ctypes.data_as(dp), 1 if flags & 1 else 0, templ)
        assert rc == 0
        _arr(out, numel)[:] = o
        return 0

    def slb_fill_dec2d(self, ctx, dec, n1, n2, tab_j, scale_j, tab_i, scale_i):
        self.calls.append("fill_dec2d")
        d = _arr(dec, n1 * n2 * 2).reshape((n1, n2, 2), order="F")
        d[:, :, 0] = (scale_j * _arr(tab_j, n2))[None, :]
        d[:, :, 1] = (scale_i * _arr(tab_i, n1))[:, None]
        return 0

    def slb_lincomb(self, ctx, out, nterms, coefs, ptrs, n):
        self.calls.append("lincomb")
        acc = coefs[0] * _arr(ptrs[0], n)
        for k in range(1, nterms):
            acc = acc + coefs[k] * _arr(ptrs[k], n)
        _arr(out, n)[:] = acc
        return 0

    # -- the sweep itself, through the oracle's C line algorithms (dry-runs of GPU test code only) ----------
    def slb_sweep(self, g, dim, h, alpha, alpha_len, strides, scale, on_device, flags):
        from oracle import clib

        self.calls.append("sweep")
        gr, it = self._g(g), self.interps[_addr(h)]
        ext, nd = gr["ext"], len(gr["ext"])
        n = ext[dim]
        if it["kind"] in (1, 2) and it["n"] != n:
            return -1
        L = clib.lib()
        oh = L.orc_interp_create(it["kind"], it["order"], n, it["coef"].ctypes.data_as(dp), it["nc"],
                                 it["nodes"].ctypes.data_as(dp) if it["nodes"] is not None else None)
        tab = scale * np.array(_arr(alpha, alpha_len), copy=True)
        data = _arr(gr["front"], gr["numel"])
        if flags & 2:  # InsideEdge: line by line
            a = data.reshape(ext, order="F")
            mv = np.moveaxis(a, dim, 0)
            other = [e for d, e in enumerate(ext) if d != dim]
            ostr = [int(strides[d]) for d in range(nd) if d != dim]
            for idx in np.ndindex(*other):
                al = float(tab[sum(i * s_ for i, s_ in zip(idx, ostr))])
                line = np.ascontiguousarray(mv[(slice(None),) + idx])
                out = np.empty(n)
                rc = L.orc_interpolate_inside(out.ctypes.data_as(dp), line.ctypes.data_as(dp), n, int(np.floor(al)), al - np.floor(al),
                                              it["coef"].ctypes.data_as(dp), it["order"], it["nc"])
                if rc != 0:
                    return -1
                mv[(slice(None),) + idx] = out
            return 0
        scratch = np.empty(gr["numel"])
        rc = L.orc_sweep(data.ctypes.data_as(dp), scratch.ctypes.data_as(dp), nd, clib.lp(ext), dim, oh, tab.ctypes.data_as(dp),
                         clib.lp([int(strides[d]) for d in range(nd)]), 1)
        L.orc_interp_destroy(oh)
        if rc == 0 and gr.get("linesum"):  # per-line sums of the outputs (line index: the other dims, Fortran order)
            ls = data.reshape(ext, order="F").sum(axis=dim).reshape(-1, order="F")
            _arr(gr["linesum"], ls.size)[:] = ls
        return rc

    # -- Vlasov-Poisson pieces (numpy restatement, 1-D space only) -----------------------------
    def slb_poisson_create(self, ctx, nsp, ext, fctv, out):
        assert nsp in (1, 2)
        shape = tuple(int(ext[d]) for d in range(nsp))
        n = int(np.prod(shape))
        pid = self._new_id()
        self.plans[pid] = {"n": n, "shape": shape, "mult": np.array(_arr(fctv[0], n), copy=True),
                           "mults": [np.array(_arr(fctv[x], n), copy=True).reshape(shape, order="F") for x in range(nsp)]}
        _set(out, pid)
        return 0

    def slb_poisson_destroy(self, p):
        return None

    def _solve(self, plan, rho, E):
        buf = np.fft.fft(rho)
        _arr(E[0], plan["n"])[:] = np.real(np.fft.ifft((1j * plan["mult"]) * buf))

    def slb_vp_field_solve(self, plan, f, nv, dv, rho, E):
        self.calls.append("field_solve")
        pl = self.plans[_addr(plan)]
        n = pl["n"]
        r = dv * _arr(f, n * nv).reshape((n, nv), order="F").sum(axis=1)
        r = r - r.sum() / n
        _arr(rho, n)[:] = r
        if len(pl["shape"]) == 2:
            return self.slb_poisson_solve(plan, rho, E)
        self._solve(pl, r, E)
        return 0

    # -- pieces bench.py touches (dry-runs of its host logic on the CPU) ---------------------------
    def slb_sweep_pair(self, *a):
        return -4  # SLB_E_UNSUPPORTED: callers issue two sweeps

    def slb_timer_start(self, ctx):
        import time
        self._t0 = time.perf_counter()
        return 0

    def slb_timer_stop(self, ctx, ms):
        import time
        _set(ms, 1e3 * (time.perf_counter() - self._t0))
        return 0

    def slb_event_create(self, ctx, out):
        eid = self._new_id()
        self.mem[("ev", eid)] = 0.0
        _set(out, eid)
        return 0

    def slb_event_record(self, ctx, ev):
        import time
        self.mem[("ev", _addr(ev))] = time.perf_counter()
        return 0

    def slb_stream_wait_event(self, ctx, ev):
        return 0

    def slb_event_elapsed_ms(self, e0, e1, ms):
        _set(ms, max(1e-6, 1e3 * (self.mem[("ev", _addr(e1))] - self.mem[("ev", _addr(e0))])))
        return 0

    # CUDA-graph entry points: the double executes eagerly, so a "recording" has already run once and a replay is
    # a no-op -- enough to dry-run bench.py's control flow, not to check graph semantics (GPU test does that)
    def slb_capture_begin(self, ctx):
        return 0

    def slb_capture_end(self, ctx, out):
        _set(out, self._new_id())
        return 0

    def slb_graph_launch(self, g):
        return 0

    def slb_graph_destroy(self, g):
        return None

    # step programs: same convention as the graph entry points (eager execution, launch is a no-op)
    def slb_program_begin(self, ctx):
        return 0

    def slb_program_end(self, ctx, out):
        _set(out, self._new_id())
        return 0

    def slb_program_launch(self, p, nrep, out_stride):
        return 0

    def slb_program_info(self, p, nops, nbarriers, nblocks):
        for q in (nops, nbarriers, nblocks):
            if q:
                _set(q, 0)
        return 0

    def slb_program_profile(self, p, nrep, out_stride, cap, kinds, wait, run):
        return 0

    def slb_program_destroy(self, p):
        return None

    def slb_reduce_sumsq_async(self, ctx, p, n, scale, out):
        _arr(out, 1)[0] = scale * float(np.sum(_arr(p, n) ** 2))
        return 0

    def slb_host_alloc(self, nbytes, out):
        _set(out, self._alloc(nbytes))
        return 0

    def slb_host_free(self, p):
        self.mem.pop(_addr(p), None)
        return 0

    def slb_charge_density(self, g, nsp, dv, rho):
        gr = self._g(g)
        n = gr["ext"][0]
        r = dv * _arr(gr["front"], gr["numel"]).reshape((n, -1), order="F").sum(axis=1)
        _arr(rho, n)[:] = r - r.sum() / n
        return 0

    def slb_poisson_solve(self, plan, rho, E):
        pl = self.plans[_addr(plan)]
        if len(pl["shape"]) == 2:  # E_x = real(ifft2(i m_x fft2(rho)))
            self.calls.append("poisson_solve_2d")
            buf = np.fft.fft2(np.array(_arr(rho, pl["n"]), copy=True).reshape(pl["shape"], order="F"))
            for x in range(2):
                _arr(E[x], pl["n"])[:] = np.real(np.fft.ifft2((1j * pl["mults"][x]) * buf)).reshape(-1, order="F")
            return 0
        self._solve(pl, np.array(_arr(rho, pl["n"]), copy=True), E)
        return 0

    def slb_reduce_sumsq(self, ctx, p, n, out):
        _set(out, float(np.sum(_arr(p, n) ** 2)))
        return 0

    def slb_kinetic_energy(self, g, nsp, vsq, scale, out):
        gr = self._g(g)
        n = gr["ext"][0]
        f = _arr(gr["front"], gr["numel"]).reshape((n, -1), order="F")
        _set(out, scale * float(np.sum(_arr(vsq, f.shape[1]) * f.sum(axis=0))))
        return 0


def install(monkeypatch):
    """route slb200._lib.lib() to a FakeLib for the duration of a test"""
    from slb200 import _lib

    fake = FakeLib()
    monkeypatch.setattr(_lib, "lib", lambda: fake)
    monkeypatch.setattr(_lib, "_default_ctx", None)
    from slb200.unsplit2d import DeviceField

    monkeypatch.setattr(DeviceField, "_pool", {})
    return fake
