"""ORACLE (test infrastructure) -- the unsplit 2-D solvers of the reference:
per-point N-D interpolation, the Adams-Bashforth time algorithms built on it and the displacement
providers that fill `bufcur`.  numpy restatement; the per-point arithmetic is oracle.c's
orc_interpolate_points2d.  Citations are relative to /root/reference.

Conventions: a field of OpTuple{2,T} (src/util.jl:3-23) of size (n1, n2) is a Fortran-ordered
numpy array [n1, n2, 2] (component planes); a scalar field is [n1, n2].

Only tests/ and __graft_entry__.smoke() may import this module.
"""
import numpy as np

from . import clib
from . import refmodel as R


def _as_planes(a):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 2:
        return np.asfortranarray(a).reshape(a.shape + (1,), order="F"), True
    return np.asfortranarray(a), False


def interpolate_points(fi, dec, interps, nthreads=1):
    """interpolate!(fp, fi, bufdec, interp_t) -- src/interpolation.jl:561-621 (N = 2); returns fp.
    fi: [n1, n2] or [n1, n2, ncomp]; dec: [n1, n2, 2] displacements in grid units."""
    planes, scalar = _as_planes(fi)
    n1, n2, nc = planes.shape
    dec = np.asfortranarray(dec, dtype=np.float64)
    assert dec.shape == (n1, n2, 2) and len(interps) == 2
    res = np.asfortranarray(R.sol_nd(interps, planes))  # sol(interp_t, fi), :48-94 (component-wise)
    out = np.empty_like(res, order="F")
    clib.lib().orc_interpolate_points2d(
        interps[0]._h, interps[1]._h, out.ctypes.data_as(clib.c_double_p), res.ctypes.data_as(clib.c_double_p),
        dec.ctypes.data_as(clib.c_double_p), n1, n2, nc, int(nthreads))
    return np.asfortranarray(out[:, :, 0]) if scalar else out


def interpolate_fct(fi, decfct, interps, nthreads=1):
    """interpolate!(fp, fi, dec::Function, interp_t) -- src/interpolation.jl:401-429: dec(ind) gives
    the tuple of shifts of point ind (0-based here)."""
    n1, n2 = fi.shape[:2]
    dec = np.empty((n1, n2, 2), order="F")
    for j in range(n2):
        for i in range(n1):
            dec[i, j, :] = decfct((i, j))
    return interpolate_points(fi, dec, interps, nthreads)


def autointerp(to, frm, nb, interps, nthreads=1):
    """autointerp!(to, from, nb, interp_t) -- src/interpolation.jl:626-655 (mutates `to`)."""
    if nb < 1:
        to[...] = frm
    fmr = frm.copy(order="F")
    for i in range(1, nb + 1):
        to[...] = interpolate_points(frm, fmr, interps, nthreads)
        if i != nb:
            fmr[...] = to


def interpbufc(t_buf, bufdec, interps, nb=None, nthreads=1):
    """interpbufc!(t_buf, bufdec, interp_t, nb = length(t_buf)) -- src/interpolation.jl:661-682"""
    nb = len(t_buf) if nb is None else nb
    for i in range(nb):
        buf = t_buf[len(t_buf) - 1 - i]
        buf[...] = interpolate_points(buf.copy(order="F"), bufdec, interps, nthreads)


def _c(adv, k, n):
    """c(st::ABcoef, k, n) = st.tab[k, n] (src/lagrange.jl:88), 1-based; Rational * Float64 promotes
    the rational to Float64 first"""
    fr = adv.abcoef[k - 1][n - 1]
    return fr.numerator / fr.denominator


def _lincomb(adv, arrs, n, ord_):
    """sum(map(k -> c(abcoef, k, ord) * arrs[k], 1:n)): rounded products, summed left to right"""
    acc = _c(adv, 1, ord_) * arrs[0]
    for k in range(2, n + 1):
        acc = acc + _c(adv, k, ord_) * arrs[k - 1]
    return np.asfortranarray(acc)


def decbegin(t_trv, t_cal, interps, nthreads=1):
    """decbegin!(t_trv, t_cal, t_interp) -- src/advection.jl:391-399"""
    indice = len(t_trv)
    for i in range(1, indice):
        buf = t_cal[-1]
        autointerp(buf, buf.copy(order="F"), indice - 1, interps, nthreads)
        interpbufc(t_trv, buf, interps, i, nthreads)
        t_cal.pop()


def initcoef(advd):
    """initcoef!(self::AdvectionData) -- src/advection.jl:404-580: the provider's initcoef! plus the
    Adams-Bashforth extrapolation of the displacement field."""
    nbtours = 3
    adv = advd.adv
    timealg, ordalg = adv.timealg, adv.ordalg
    interps = adv.t_interp
    nt = adv.nthreads
    isbegin = advd.bufcur is None
    ext = advd.parext
    ext.initcoef(advd)

    if timealg == R.ABTimeAlg_new and isbegin:  # :414-481
        t_ref, t_cal = [], []
        t_ref.append(advd.bufcur.copy(order="F"))
        svdata = advd.data.copy(order="F")
        svbufcur = advd.bufcur.copy(order="F")
        sens = 1 if (ordalg * nbtours) % 2 == 1 else -1
        for indice in range(1, ordalg + 1):
            for nb in range(1, (nbtours - 1 if indice == ordalg else nbtours) + 1):
                t_ref = t_ref[::-1]
                t_trv = [np.asfortranarray(sens * a) for a in t_ref]
                decbegin(t_trv, t_cal, interps, nt)
                t_cal = []
                advd.data[...] = svdata
                for i in range(1, indice + 1):
                    fmrdec = _lincomb(adv, t_trv, indice, indice)
                    if i != 1:
                        t_cal.append(np.asfortranarray(0.0 - fmrdec))
                    if i != 1 or nb != 1:
                        t_trv.pop()
                        t_ref.pop()
                    autointerp(advd.bufcur, fmrdec, indice, interps, nt)
                    interpbufc(advd.t_bufc, advd.bufcur, interps, None, nt)
                    f = interpolate_points(advd.data, advd.bufcur, interps, nt)
                    advd.data[...] = f
                    ext.initcoef(advd)
                    t_trv.insert(0, advd.bufcur.copy(order="F"))
                    t_ref.insert(0, advd.bufcur.copy(order="F"))
                fmrdec = np.asfortranarray(0.0 - _lincomb(adv, t_trv, indice + 1, indice + 1))
                t_cal.append(fmrdec)
                sens = -sens
        assert sens == 1, "sens must be positive at this place"
        t_ref = t_ref[::-1]
        t_trv = [np.asfortranarray(sens * a) for a in t_ref]
        del t_trv[0]
        decbegin(t_trv, t_cal, interps, nt)
        advd.t_bufc = t_trv
        advd.data[...] = svdata
        advd.bufcur[...] = svbufcur

    if timealg == R.ABTimeAlg_ip and isbegin:  # :484-508
        for indice in range(1, ordalg):
            advd.t_bufc.insert(0, advd.bufcur.copy(order="F"))
            fmrdec = _lincomb(adv, advd.t_bufc, indice, indice)
            autointerp(fmrdec, fmrdec.copy(order="F"), indice - 1, interps, nt)
            interpbufc(advd.t_bufc, fmrdec, interps, None, nt)

    if timealg == R.ABTimeAlg_init and isbegin:  # :509-543
        for indice in range(1, len(advd.initdatas) + 1):
            advd.t_bufc.insert(0, advd.bufcur.copy(order="F"))
            ord_ = min(indice, ordalg)
            fmrdec = _lincomb(adv, advd.t_bufc, ord_, ord_)
            if ord_ == ordalg:
                advd.t_bufc.pop()
            autointerp(fmrdec, fmrdec.copy(order="F"), ordalg - 1, interps, nt)
            interpbufc(advd.t_bufc, fmrdec, interps, None, nt)
            advd.data[...] = advd.initdatas[indice - 1]
            advd.time_cur += advd.getcur_t()
            ext.initcoef(advd)

    if timealg in (R.ABTimeAlg_ip, R.ABTimeAlg_new, R.ABTimeAlg_init):  # :545-578
        advd.t_bufc.insert(0, advd.bufcur.copy(order="F"))
        bufc = _lincomb(adv, advd.t_bufc, ordalg, ordalg)
        autointerp(advd.bufcur, bufc, ordalg - 1, interps, nt)
        advd.t_bufc.pop()
        interpbufc(advd.t_bufc, advd.bufcur, interps, None, nt)


def advection_single_state(advd):
    """advection! when the Advection has ONE state with per-point shifts
    (src/advection.jl:594-619, :703-704): f = interpolate(data, bufcur); data = f; nextstate!."""
    adv = advd.adv
    st = advd.getst()
    if len(adv.states) != 1 and st.ndims == 1:
        return advection_split_points(advd)
    if len(adv.states) != 1 or st.ndims != 2 or adv.N != 2 or st.perm != [1, 2]:
        raise NotImplementedError("oracle: per-point shifts are covered for one 2-D state ([1, 2], 2, 1, false)")
    initcoef(advd)
    advd.data[...] = interpolate_points(advd.data, advd.bufcur, adv.t_interp, adv.nthreads)
    return advd.nextstate()


def interpolate_line_points(fi, alphas, interp):
    """interpolate!(fp, fi, dec::Function, interp_t) for N = 1 (src/interpolation.jl:401-429): every point of the
    line has its own shift; (dint, tab) = getprecal(cache, dec(ind)) is the floor split of :381-389 and
    fp[ind] = sum(res[window] .* tab) -- rounded products summed left to right."""
    n = len(fi)
    res = interp.sol(np.ascontiguousarray(fi, dtype=np.float64)) if interp.kind in (R.BSPLINE_LU, R.BSPLINE_FFT) else fi
    p = interp.order
    out = np.empty(n)
    for i in range(n):
        a = float(alphas[i])
        dint = int(np.floor(a))
        w = interp.getprecal(a - dint)
        acc = res[(i + dint - p // 2) % n] * w[0]
        for j in range(1, p + 1):
            acc = acc + res[(i + dint - p // 2 + j) % n] * w[j]
        out[i] = acc
    return out


def advection_split_points(advd):
    """advection! for a SPLIT state with per-point shifts, e.g. [([1, 2], 1, 1, false), ([2, 1], 1, 2, false)] -- the
    split form of the quasi-geostrophic driver (src/advection.jl:633-645 with getalpha(parext, self, indext, indbuf),
    src/quasigeostrophic.jl:126-135): every line along dim perm[1] is interpolated with the shifts
    bufcur[ind][invp[1]] of its own points.  2-D grids, NoTimeAlg."""
    adv = advd.adv
    st = advd.getst()
    if adv.N != 2 or st.ndims != 1 or adv.timealg != R.NoTimeAlg:
        raise NotImplementedError("oracle: split per-point states are covered for 2-D grids without a time algorithm")
    advd.parext.initcoef(advd)
    d, comp = st.perm[0] - 1, st.invp[0] - 1
    interp = adv.t_interp[d]
    out = np.empty_like(advd.data)
    if d == 0:
        for j in range(adv.sizeall[1]):
            out[:, j] = interpolate_line_points(np.ascontiguousarray(advd.data[:, j]), advd.bufcur[:, j, comp], interp)
    else:
        for i in range(adv.sizeall[0]):
            out[i, :] = interpolate_line_points(np.ascontiguousarray(advd.data[i, :]), advd.bufcur[i, :, comp], interp)
    advd.data[...] = out
    return advd.nextstate()


# ---------------------------------------------------------------------------------------
# providers
# ---------------------------------------------------------------------------------------
class PoissonVar2d(R.PoissonVar):
    """PoissonVar{..., StdPoisson2d}: initcoef! of src/poisson.jl:229-247"""

    def initcoef(self, advd):
        adv = advd.adv
        self.compute_charge(advd)
        self.compute_elfield()
        bufc_v = (advd.getcur_t() / adv.t_mesh[1].step) * self.t_elfield[0]
        bufc_sp = (-advd.getcur_t() / adv.t_mesh[0].step) * adv.t_mesh[1].points
        if advd.bufcur is None:
            advd.bufcur = np.zeros(adv.sizeall + (2,), order="F")
        advd.bufcur[:, :, 0] = bufc_sp[None, :]
        advd.bufcur[:, :, 1] = bufc_v[:, None]


def getpoissonvar2d(adv):
    """getpoissonvar(adv; type = StdPoisson2d) -- src/poisson.jl:100-103"""
    return PoissonVar2d(adv)


class RotationVar2d:
    """RotationVar with the ABTimeAlg initcoef! of src/rotation.jl:36-54"""

    def __init__(self, adv):
        if adv.timealg not in (R.ABTimeAlg_ip, R.ABTimeAlg_new):
            raise ValueError("the unsplit rotation is defined for ABTimeAlg_ip / ABTimeAlg_new (src/rotation.jl:36-42)")

    def initcoef(self, advd):
        adv = advd.adv
        buf1 = -advd.getcur_t() / adv.t_mesh[0].step * adv.t_mesh[1].points
        buf2 = advd.getcur_t() / adv.t_mesh[1].step * adv.t_mesh[0].points
        if advd.bufcur is None:
            advd.bufcur = np.zeros(adv.sizeall + (2,), order="F")
        advd.bufcur[:, :, 0] = buf1[None, :]
        advd.bufcur[:, :, 1] = buf2[:, None]


def getrotationvar2d(adv):
    return RotationVar2d(adv)


# ---------------------------------------------------------------------------------------
# surface quasi-geostrophic provider -- src/quasigeostrophic.jl:1-135 (single unsplit 2-D state)
# ---------------------------------------------------------------------------------------
class GeoVar:
    """GeoConst + GeoVar (src/quasigeostrophic.jl:6-74): u = real(ifft(coefrsqk .* fft(b))) with
    coefrsqk[1] = i k_y / |k|, coefrsqk[2] = -i k_x / |k| (zero mode 0); bufcur = dt * (u_1, u_2)."""

    def __init__(self, adv, odg_b=1e-3):
        if adv.N != 2:
            raise ValueError("the number of dimension must be 2")
        self.odg_b = float(odg_b)
        kx, ky = R.vec_k_fft(adv.t_mesh[0])[:, None], R.vec_k_fft(adv.t_mesh[1])[None, :]
        k = np.sqrt(kx**2 + ky**2)       # sqrt(sum(tp .^ 2)): kx^2 + ky^2 summed left to right
        with np.errstate(divide="ignore"):
            v = np.where(k == 0, 0.0, 1.0 / k)
        self.coef_imag = (np.asfortranarray(v * ky), np.asfortranarray(-(v * kx)))   # imaginary parts (:37-42)

    def initdata(self, advd):
        """initdata! -- src/quasigeostrophic.jl:79-104"""
        mx, my = advd.adv.t_mesh
        lx, ly = mx.stop - mx.start, my.stop - my.start
        x, y = mx.points, my.points
        ee = 4
        sig = lx / 15

        def anticyclone(cx, cy):
            return np.exp(-ee * (x - cx) ** 2 / (2 * sig**2))[:, None] * np.exp(-((y - cy) ** 2) / (2 * sig**2))[None, :]

        d = anticyclone(lx / 4, ly / 4)
        d = d + anticyclone(3 * lx / 4, ly / 4)
        d = d - anticyclone(lx / 4, 3 * ly / 4)
        d = d - anticyclone(3 * lx / 4, 3 * ly / 4)
        advd.data[...] = d * self.odg_b

    def initcoef(self, advd):
        """initcoef! -- src/quasigeostrophic.jl:109-121"""
        dt = advd.getcur_t()
        buf = np.fft.fft2(advd.data)
        if advd.bufcur is None:
            advd.bufcur = np.zeros(advd.adv.sizeall + (2,), order="F")
        for x in range(2):
            advd.bufcur[:, :, x] = dt * np.real(np.fft.ifft2((1j * self.coef_imag[x]) * buf))


def getgeovar(adv, **kw):
    return GeoVar(adv, **kw)
