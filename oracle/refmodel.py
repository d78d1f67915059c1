"""ORACLE (test infrastructure) -- Python/numpy restatement of the reference's driver layer.

Mirrors, with file:line citations (relative to /root/reference), the Julia objects on the
hot path: UniformMesh, the interpolation types, the splitting tables, Advection /
AdvectionData / advection! / nextstate!, and the Poisson / rotation / translation
displacement plugins.  The per-line arithmetic is done by oracle.c (liboracle.so).

Arrays are numpy float64 in Fortran order with Julia's shapes, so index i here is index
i+1 there.  `states` keep the reference's 1-based permutations.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may
import this module.
"""
from fractions import Fraction
import math

import numpy as np

from . import clib, tables

LAGRANGE, BSPLINE_LU, BSPLINE_FFT, HERMITE = 0, 1, 2, 3


# ---------------------------------------------------------------------------------------
# src/mesh.jl:21-33, :48-98, :110-115
# ---------------------------------------------------------------------------------------
class UniformMesh:
    """src/mesh.jl:21-33.  Julia builds the nodes with a twice-precision `range`
    (start + i*step evaluated in double-double); exact rational arithmetic rounded once
    reproduces those values (ulp-level agreement is unpinned: no Julia here)."""

    def __init__(self, start, stop, length):
        start, stop = float(start), float(stop)
        fs, fe = Fraction(start), Fraction(stop)
        st = (fe - fs) / length
        self.points = np.array([float(fs + i * st) for i in range(length)], dtype=np.float64)
        self.step = float(st)
        self.width = stop - start
        self.length = length

    def __len__(self):
        return self.length

    @property
    def start(self):  # src/mesh.jl:53
        return self.points[0]

    @property
    def stop(self):  # src/mesh.jl:58
        return self.points[-1] + self.step


def vec_k_fft(mesh):
    """src/mesh.jl:110-115: 2pi/width .* fftfreq(nx, nx)"""
    nx = len(mesh)
    k = 2 * math.pi / mesh.width
    freq = np.array([i if i < (nx + 1) // 2 else i - nx for i in range(nx)], dtype=np.float64)
    return k * freq


# ---------------------------------------------------------------------------------------
# interpolation types: src/lagrange.jl:58-72, src/bsplinelu.jl:253-270,
# src/bsplinefft.jl:25-45, src/hermite.jl:99-132
# ---------------------------------------------------------------------------------------
# src/interpolation.jl:3  @enum EdgeType
CircEdge, InsideEdge = 1, 2


class Interp:
    def __init__(self, kind, order, n=0, flbis=False, edge=CircEdge):
        self.kind, self.order, self.n, self.edge = kind, order, n, edge
        if kind == LAGRANGE:
            rat = tables.lagrange_tabfct_rat(order)
            nodes = None
        elif kind in (BSPLINE_LU, BSPLINE_FFT):
            if kind == BSPLINE_LU and order % 2 == 0:
                raise ValueError(f"order={order} BSplineLU for even  order is not implemented n={n}")
            if kind == BSPLINE_FFT and (n & (n - 1)) != 0:
                raise ValueError("n must be a power of two")  # src/fftbig.jl:57
            rat = tables.bspline_tabfct_rat(order)
            nodes = np.array([float(x) for x in tables.bspline_node_values_rat(order)])
        elif kind == HERMITE:
            rat = tables.hermite_tabfct_rat(order, flbis=flbis)
            nodes = None
        else:
            raise ValueError("unknown kind")
        self.tabfct = np.array(tables.to_float64_table(rat), dtype=np.float64)  # (order+1, nc)
        self.nodes = nodes
        L = clib.lib()
        self._h = L.orc_interp_create(
            kind, order, n, clib.dp(self.tabfct), self.tabfct.shape[1], clib.dp(nodes) if nodes is not None else None
        )
        if not self._h:
            raise RuntimeError("orc_interp_create failed")

    def __del__(self):
        try:
            clib.lib().orc_interp_destroy(self._h)
        except Exception:
            pass

    # src/interpolation.jl:96-98
    def getprecal(self, decf):
        w = np.empty(self.order + 1)
        clib.lib().orc_getprecal(clib.dp(self.tabfct), self.order + 1, self.tabfct.shape[1], float(decf), clib.dp(w))
        return w

    # sol(interp, b): src/interpolation.jl:40, src/bsplinelu.jl:282-284, src/bsplinefft.jl:49-51
    def sol(self, b):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty_like(b)
        if self.kind in (BSPLINE_LU, BSPLINE_FFT):
            assert len(b) == self.n
            clib.lib().orc_sol_line(self._h, clib.dp(x), clib.dp(b))
            return x
        return b.copy()


def Lagrange(order, edge=CircEdge):
    return Interp(LAGRANGE, order, edge=edge)


def BSplineLU(order, n):
    return Interp(BSPLINE_LU, order, n)


def BSplineFFT(order, n):
    return Interp(BSPLINE_FFT, order, n)


def Hermite(order, flbis=False):
    return Interp(HERMITE, order, flbis=flbis)


def interpolate(fp, fi, dec, interp):
    """src/interpolation.jl:302-315 (CircEdge): decint = floor(dec); decfloat = dec - decint;
    interpolate!(fp, fi, decint, getprecal(interp, decfloat), interp)."""
    fi = np.ascontiguousarray(fi, dtype=np.float64)
    assert fp.flags.c_contiguous and fp is not fi
    if interp.edge == InsideEdge:
        # :308-314: interpolate!(fp, fi, decint, get_allprecal(interp, decint, decfloat), interp)
        decint = int(np.floor(dec))
        res = interp.sol(fi)
        rc = clib.lib().orc_interpolate_inside(clib.dp(fp), clib.dp(res), len(fi), decint, float(dec) - decint, clib.dp(interp.tabfct),
                                               interp.order, interp.tabfct.shape[1])
        if rc != 0:
            raise ValueError("InsideEdge: the shift moves the stencil window outside the array")
        return None
    clib.lib().orc_interpolate_alpha(interp._h, clib.dp(fp), clib.dp(fi), len(fi), float(dec))
    return None


def interpolate_precal(fp, fi, decint, precal, interp):
    """src/interpolation.jl:175-193 with explicit (decint, precal)."""
    res = interp.sol(fi)
    precal = np.ascontiguousarray(precal, dtype=np.float64)
    clib.lib().orc_interpolate_circ(clib.dp(fp), clib.dp(res), len(fi), int(decint), clib.dp(precal), interp.order)


# ---------------------------------------------------------------------------------------
# src/splitting.jl:4-104
# ---------------------------------------------------------------------------------------
def nosplit(dt):
    return [dt * 1]


def standardsplit(dt):
    return [dt * 1, dt * 1]


def strangsplit(dt):
    # dt * [1//2, 1//1, 1//2]: Float64 * Rational -> dt*num/den
    return [dt * 1 / 2, dt * 1 / 1, dt * 1 / 2]


def magicsplit(dt):
    return [math.tan(dt / 2), math.sin(dt), math.tan(dt / 2)]


def triplejumpsplit(dt):
    c = 2.0 ** (1.0 / 3.0)
    c1 = 1 / (2 * (2 - c))
    c2 = (1 - c) / (2 * (2 - c))
    d1 = 1 / (2 - c)
    d2 = -c / (2 - c)
    return [dt * x for x in (c1, d1, c2, d2, c2, d1, c1)]


# ---------------------------------------------------------------------------------------
# src/advection.jl:10-20, :73-141, :152-163
# ---------------------------------------------------------------------------------------
def modone(ind, n):  # src/util.jl:73
    return (ind - 1) % n + 1


def invperm(p):
    q = [0] * len(p)
    for i, v in enumerate(p):
        q[v - 1] = i + 1
    return q


# src/advection.jl:2  @enum TimeAlgorithm
NoTimeAlg, ABTimeAlg_ip, ABTimeAlg_new, ABTimeAlg_init = 1, 2, 3, 4


class StateAdv:
    def __init__(self, ind, perm, ndims, stcoef, isconstdec):
        self.ind, self.perm, self.invp = ind, list(perm), invperm(perm)
        self.ndims, self.stcoef, self.isconstdec = ndims, stcoef, isconstdec


class Advection:
    """src/advection.jl:73-141"""

    def __init__(self, t_mesh, t_interp, dt_base, states, tab_coef=None, nthreads=1, timealg=NoTimeAlg, ordalg=None):
        N = len(t_mesh)
        # :96-97 timealg::TimeAlgorithm = NoTimeAlg, ordalg::Int = timealg != NoTimeAlg ? 4 : 0
        self.timealg = timealg
        self.ordalg = (4 if timealg != NoTimeAlg else 0) if ordalg is None else int(ordalg)
        self.abcoef = tables.abcoef_rat(self.ordalg + 1)  # :136 ABcoef(ordalg + 1)
        if len(t_interp) != N:
            raise ValueError(f"size of vector of Interpolation must be equal to N={N}")
        self.sizeall = tuple(len(m) for m in t_mesh)
        self.t_mesh, self.t_interp, self.dt_base = tuple(t_mesh), list(t_interp), float(dt_base)
        self.states = [StateAdv(i + 1, *s) for i, s in enumerate(states)]
        self.tab_coef = list(strangsplit(dt_base) if tab_coef is None else tab_coef)
        self.maxcoef = max(s.stcoef for s in self.states)
        restcoef = len(self.tab_coef) % self.maxcoef
        # `x.stcoef in restcoef` with an Int on the right == (x.stcoef == restcoef)
        nbstatesplus = len([s for s in self.states if s.stcoef == restcoef])
        self.nbstates = (len(self.tab_coef) // self.maxcoef) * len(self.states) + nbstatesplus
        self.nthreads = nthreads  # NoTimeOpt (1) or SimpleThreadsOpt (>1)
        self.N = N

    def getst(self, x):  # :152
        return self.states[modone(x, len(self.states)) - 1]

    def getstcoef(self, x):  # :154-156
        return ((x - 1) // len(self.states)) * self.maxcoef + self.getst(x).stcoef

    def getcur_t(self, x):  # :158
        return self.tab_coef[self.getstcoef(x) - 1]

    def getinterp(self, x):  # :160-163
        st = self.getst(x)
        return [self.t_interp[d - 1] for d in st.perm[: st.ndims]]


class AdvectionData:
    """src/advection.jl:229-313 (state + data); advection! is `advection(advd)` below."""

    def __init__(self, adv, data, parext, time_init=0.0, initdatas=None):
        if tuple(data.shape) != adv.sizeall:
            raise ValueError(f"size(data)={data.shape} it must be {adv.sizeall}")
        self.adv = adv
        self.bufcur = None          # :241 bufcur (missing): per-point displacement field [sizeall..., N]
        self.t_bufc = []            # :242 t_bufc: history of displacement fields (AB time algorithms)
        self.initdatas = initdatas  # :243
        self.state_gen = 1
        self.time_cur = float(time_init)
        self.data = np.array(data, dtype=np.float64, order="F", copy=True)  # :264-267
        self.bufdata = np.empty(self.data.size, dtype=np.float64)
        self.parext = parext

    # accessors :315-336
    def getst(self):
        return self.adv.getst(self.state_gen)

    def getcur_t(self):
        return self.adv.getcur_t(self.state_gen)

    def getstcoef(self):
        return self.adv.getstcoef(self.state_gen)

    def getinterp(self):
        return self.adv.getinterp(self.state_gen)

    def getdata(self):
        return self.data

    def _getcurrentindice(self):
        return self.getst().perm[0]

    def nextstate(self):  # :358-367
        if self.state_gen < self.adv.nbstates:
            self.state_gen += 1
            return True
        self.state_gen = 1
        self.time_cur += self.adv.dt_base
        return False


def sol_nd(interps, b):
    """sol(interp_t, b) for an N-D slice (src/interpolation.jl:48-94): the 1-D solve of each
    non-identity interpolation applied along its dimension."""
    out = np.array(b, dtype=np.float64, copy=True)
    for d, it in enumerate(interps):
        if it.kind in (BSPLINE_LU, BSPLINE_FFT):
            out = np.apply_along_axis(it.sol, d, out)
    return out


def interpolate_nd_const(fi, dec, interps):
    """interpolate!(fp, fi, decint::NTuple, precal::Array, interp::Vector) -- the N-D constant-shift
    tensor stencil of src/interpolation.jl:212-231, with precal = dotprod of the 1-D weights
    (src/interpolation.jl:112-118) and (decint, decfloat) = floor split per dim (:381-389):
        fp[ind] = sum(res[window(ind)] .* precal),  window_d = ind_d + decint_d - order_d/2 + (0..order_d)
    The products res .* precal are rounded, then summed (Julia's sum over < 1024 elements is a
    plain loop; its @simd reassociation is not pinned by any test -- see oracle.c's header)."""
    nd = fi.ndim
    res = sol_nd(interps, fi)
    decint = [int(np.floor(a)) for a in dec]
    ws = [interps[d].getprecal(float(dec[d]) - decint[d]) for d in range(nd)]
    precal = ws[0]
    for d in range(1, nd):
        precal = np.multiply.outer(precal, ws[d])     # dotprod: rounded products of the 1-D weights
    out = np.zeros_like(res)
    for idx in np.ndindex(*precal.shape):             # column-major order of Julia's sum would be
        sh = res                                      # reversed(idx); the set of terms is the same
        for d in range(nd):
            off = decint[d] - interps[d].order // 2 + idx[d]
            sh = np.roll(sh, -off, axis=d)
        out = out + sh * precal[idx]
    return out


def _advection_nd(advd):
    """advection! for a const-shift state with ndims > 1 (src/advection.jl:622-632 with the N-D
    interpolate!): every slice over the first ndims permuted dims is shifted by getalpha(indext)."""
    adv = advd.adv
    st = advd.getst()
    nd = st.ndims
    ext = advd.parext
    ext.initcoef(advd)
    dims = [p - 1 for p in st.perm[:nd]]
    rest = [p - 1 for p in st.perm[nd:]]
    interps = [adv.t_interp[d] for d in dims]
    f = np.transpose(advd.data, dims + rest)          # getformdata: permuted view
    out = np.empty_like(f)
    for ind in np.ndindex(*[adv.sizeall[d] for d in rest]):
        alpha = ext.getalpha_nd(advd, ind)            # tuple of ndims shifts (grid units)
        sl = (slice(None),) * nd + ind
        out[sl] = interpolate_nd_const(np.ascontiguousarray(f[sl]), alpha, interps)
    advd.data[...] = np.transpose(out, np.argsort(dims + rest))   # copydata!
    return advd.nextstate()


def advection(advd):
    """advection!(self)  src/advection.jl:594-704, const-shift states
    (NoTimeOpt :622-632 / SimpleThreadsOpt :647-657)."""
    adv = advd.adv
    st = advd.getst()
    if not st.isconstdec:
        from . import unsplit2d

        return unsplit2d.advection_single_state(advd)
    if st.ndims != 1:
        return _advection_nd(advd)
    interp = advd.getinterp()[0]
    ext = advd.parext
    ext.initcoef(advd)  # :407-408
    dim = st.perm[0] - 1
    tab, astride = ext.alpha_table(advd)  # getalpha for every line, as a strided table
    tab = np.ascontiguousarray(tab, dtype=np.float64)
    rc = clib.lib().orc_sweep(
        advd.data.ctypes.data_as(clib.c_double_p),
        clib.dp(advd.bufdata),
        adv.N,
        clib.lp(adv.sizeall),
        dim,
        interp._h,
        clib.dp(tab),
        clib.lp(astride),
        adv.nthreads,
    )
    if rc != 0:
        raise RuntimeError(f"orc_sweep failed rc={rc}")
    return advd.nextstate()


# ---------------------------------------------------------------------------------------
# translation plugin -- src/translation.jl:5-35
# ---------------------------------------------------------------------------------------
class TranslationVar:
    def __init__(self, values):
        self.values = tuple(float(v) for v in values)
        self.valok = None

    def initcoef(self, advd):
        st = advd.getst()
        self.valok = tuple(self.values[st.perm[i] - 1] * advd.getcur_t() for i in range(st.ndims))

    def alpha_table(self, advd):
        return np.array([self.valok[0]]), [0] * advd.adv.N

    def getalpha_nd(self, advd, ind):  # getalpha(pv::TranslationVar, self, ind) = pv.valok  (src/translation.jl:33-35)
        return self.valok


def gettranslationvar(v):
    return TranslationVar(v)


# ---------------------------------------------------------------------------------------
# rotation plugin -- src/rotation.jl:4-31, :59-71
# ---------------------------------------------------------------------------------------
class RotationVar:
    def __init__(self, adv):
        self.decfl = None

    def initcoef(self, advd):
        st_cur, st_other = advd.getst().perm
        mesh_cur = advd.adv.t_mesh[st_cur - 1]
        mesh_other = advd.adv.t_mesh[st_other - 1]
        sign = -1 if st_cur == 1 else 1
        self.decfl = sign * advd.getcur_t() / mesh_cur.step * mesh_other.points
        self._other = st_other

    def alpha_table(self, advd):
        astride = [0, 0]
        astride[self._other - 1] = 1
        return self.decfl, astride


def getrotationvar(adv):
    return RotationVar(adv)


# ---------------------------------------------------------------------------------------
# Poisson plugin -- src/poisson.jl:7-15, :35-95, :119-125, :139-144, :155-224;
# src/util_poisson.jl:15-53, :68-79, :133-183
# ---------------------------------------------------------------------------------------
def _get_fctv_k(adv):
    """src/poisson.jl:7-15: fctv_k[x] = k_x .* (im ./ |k|^2), zero mode 0."""
    N = adv.N
    Nsp = N // 2
    v_k = [vec_k_fft(m) for m in adv.t_mesh[:Nsp]]
    sz = [len(m) for m in adv.t_mesh[:Nsp]]
    s = np.zeros(sz, order="F")
    for x in range(Nsp):  # sum(v .^ 2) left to right
        shape = [1] * Nsp
        shape[x] = sz[x]
        s = s + (v_k[x] ** 2).reshape(shape)
    with np.errstate(divide="ignore"):
        inv = 1.0 / s  # im / r = Complex(0/r, 1/r)
    inv.reshape(-1, order="F")[0] = 0.0
    out = []
    for x in range(Nsp):
        shape = [1] * Nsp
        shape[x] = sz[x]
        out.append(np.asfortranarray(v_k[x].reshape(shape) * inv))  # imaginary parts
    return out


class PoissonVar:
    """PoissonConst + PoissonVar (src/poisson.jl:35-95), StdPoisson."""

    def __init__(self, adv):
        N = adv.N
        if N % 2 != 0:
            raise ValueError(f"N={N} must be a multiple of 2")
        self.adv = adv
        self.Nsp = self.Nv = N // 2
        self.fctv_k_imag = _get_fctv_k(adv)
        self.v_square = dotprod([m.points for m in adv.t_mesh[self.Nsp:]]) ** 2  # src/poisson.jl:51
        self.rho = np.empty([len(m) for m in adv.t_mesh[: self.Nsp]], order="F")
        self.t_elfield = None
        self.bufcur_sp = None
        self.bufcur_v = None
        self.tupleind = None

    def compute_charge(self, advd):  # src/poisson.jl:119-125 -> src/util_poisson.jl:68-79
        adv = self.adv
        dv = 1.0
        for m in adv.t_mesh[self.Nsp:]:  # prod(step, t_mesh_v)
            dv = dv * m.step
        nsp = int(np.prod(adv.sizeall[: self.Nsp]))
        nv = int(np.prod(adv.sizeall[self.Nsp:]))
        clib.lib().orc_compute_charge(
            self.rho.ctypes.data_as(clib.c_double_p), advd.data.ctypes.data_as(clib.c_double_p), nsp, nv, dv, adv.nthreads
        )

    def compute_elfield(self):  # src/poisson.jl:139-144
        buf = np.fft.fftn(self.rho)
        self.t_elfield = tuple(
            np.asfortranarray(np.real(np.fft.ifftn((1j * self.fctv_k_imag[x]) * buf))) for x in range(self.Nsp)
        )

    def isvelocity(self, advd):  # :155-158
        return advd.getst().perm[0] > self.Nsp

    def initcoef(self, advd):  # :164-205
        st = advd.getst()
        adv = advd.adv
        dt = advd.getcur_t()
        Nsp = self.Nsp
        if self.isvelocity(advd):
            if (Nsp + 1) in st.perm[: st.ndims]:
                self.compute_charge(advd)
                self.compute_elfield()
            self.bufcur_v = tuple(
                (dt / adv.t_mesh[st.perm[x] - 1].step) * self.t_elfield[st.perm[x] - 1 - Nsp] for x in range(st.ndims)
            )
        else:
            self.tupleind = tuple(st.perm[st.invp[x] + Nsp - 1] - st.ndims for x in range(st.ndims))
            self.bufcur_sp = tuple(
                (-dt / adv.t_mesh[st.invp[x] - 1].step) * adv.t_mesh[st.invp[x] + Nsp - 1].points for x in range(st.ndims)
            )

    def alpha_table(self, advd):
        """getalpha (src/poisson.jl:210-224) for every trailing index, as (table, strides per
        original dim).  Velocity: bufcur_v[1][ind.I[end-Nsp+1:end]]; the trailing index runs
        over dims perm[2:], so E's axis i is indexed by dim perm[N-Nsp+i].  Space:
        bufcur_sp[1][ind.I[tupleind[1]]] = index along dim perm[1+tupleind[1]]."""
        st = advd.getst()
        N, Nsp = advd.adv.N, self.Nsp
        astride = [0] * N
        if self.isvelocity(advd):
            tab = self.bufcur_v[0]
            stride = 1
            for i in range(Nsp):
                d = st.perm[N - Nsp + i]  # 1-based dim supplying E's axis i
                astride[d - 1] = stride
                stride *= tab.shape[i]
            return tab.reshape(-1, order="F"), astride
        tab = self.bufcur_sp[0]
        d = st.perm[st.ndims + self.tupleind[0] - 1]
        astride[d - 1] = 1
        return tab, astride


def _poisson_getalpha_nd(self, advd, ind):
    """getalpha (src/poisson.jl:210-224) for a state with ndims > 1: `ind` is the 0-based trailing
    index over dims perm[ndims:]."""
    st = advd.getst()
    Nsp = self.Nsp
    if self.isvelocity(advd):
        sub = ind[len(ind) - Nsp:]                       # ind.I[end-Nsp+1:end]
        return tuple(b[sub] for b in self.bufcur_v)
    return tuple(self.bufcur_sp[x][ind[self.tupleind[x] - 1]] for x in range(st.ndims))


PoissonVar.getalpha_nd = _poisson_getalpha_nd


def getpoissonvar(adv):
    return PoissonVar(adv)


def compute_ee(advd):
    """src/util_poisson.jl:156-162: dx * sum(map(x -> sum(x .^ 2), t_elfield))"""
    pv = advd.parext
    dx = 1.0
    for m in advd.adv.t_mesh[: pv.Nsp]:
        dx = dx * m.step
    tot = 0.0
    for e in pv.t_elfield:
        tot = tot + float(np.sum(e**2))
    return dx * tot


def compute_ke(advd):
    """src/util_poisson.jl:41-53"""
    pv = advd.parext
    adv = advd.adv
    dsp = 1.0
    for m in adv.t_mesh[: pv.Nsp]:
        dsp *= m.step
    dv = 1.0
    for m in adv.t_mesh[pv.Nsp:]:
        dv *= m.step
    sum_sp = np.sum(advd.data, axis=tuple(range(pv.Nsp)))
    return (dsp * dv) * float(np.sum(pv.v_square * sum_sp))


def getenergy(advd):
    """src/util_poisson.jl:167-175"""
    pv = advd.parext
    pv.compute_charge(advd)
    pv.compute_elfield()
    ee = compute_ee(advd)
    ke = compute_ke(advd)
    return ee, ke, ee + ke


def dotprod(vs):
    """src/util.jl:59-67: outer product of vectors as an N-D (Fortran-order) array."""
    N = len(vs)
    res = np.ones([1] * N)
    for i, v in enumerate(vs):
        shape = [1] * N
        shape[i] = len(v)
        res = res * np.asarray(v).reshape(shape)
    return np.asfortranarray(res)
