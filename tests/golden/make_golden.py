"""Generates tests/golden/oracle_vectors.npz: small seeded input/output vectors of the ORACLE for every
interpolation kind on the hot path and for the unsplit 2-D driver.

The reference (Julia) cannot run in this container or on the GPU box and holds no stored golden vectors
(SURVEY.md 8c), so these fixtures do not pin the oracle to the reference -- the known-answer tests in
tests/test_oracle_*.py do that.  They freeze the oracle's OUTPUTS: a later change to oracle.c / refmodel.py
that moves any value by one ulp, or a GPU kernel that drifts from the 1e-12 bar, is caught against numbers
that do not depend on the oracle being rebuilt the same way.

    python tests/golden/make_golden.py        (from the repo root; rewrites the .npz)
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refmodel as R, unsplit2d as U  # noqa: E402

CASES = [("lagrange", 3, 24), ("lagrange", 7, 32), ("lagrange", 9, 40), ("lagrange", 11, 32), ("hermite", 5, 24), ("hermite", 9, 32),
         ("bspline_lu", 3, 24), ("bspline_lu", 5, 32), ("bspline_lu", 11, 32), ("bspline_fft", 7, 32), ("bspline_fft", 11, 64)]
ALPHAS = np.array([0.345141526199181716, -0.3859416191876155, -1.2856139011444161, 4.98766514566778, -5.678513256790099, 3.0, 0.0, 17.25])


def make_interp(M, kind, order, n):
    return {"lagrange": lambda: M.Lagrange(order), "hermite": lambda: M.Hermite(order), "bspline_lu": lambda: M.BSplineLU(order, n),
            "bspline_fft": lambda: M.BSplineFFT(order, n)}[kind]()


def line_inputs(kind, order, n):
    rng = np.random.default_rng(20240611 + 100 * order + n)
    return np.asfortranarray(rng.random((n, len(ALPHAS))))


def main():
    out = {"alphas": ALPHAS}
    for kind, order, n in CASES:
        f = line_inputs(kind, order, n)
        it = make_interp(R, kind, order, n)
        res = np.empty_like(f)
        for k, a in enumerate(ALPHAS):
            col = np.empty(n)
            R.interpolate(col, np.ascontiguousarray(f[:, k]), float(a), it)
            res[:, k] = col
        out[f"line_{kind}_{order}_{n}"] = res
    # per-point 2-D interpolation (Lagrange 5 x Lagrange 5 and B-spline LU 5 x 5), one scalar field
    rng = np.random.default_rng(7)
    f2 = np.asfortranarray(rng.random((32, 24)))
    dec = np.asfortranarray(rng.uniform(-5, 5, (32, 24, 2)))
    out["points_in"], out["points_dec"] = f2, dec
    out["points_lagrange5"] = U.interpolate_points(f2, dec, [R.Lagrange(5), R.Lagrange(5)])
    out["points_bsplinelu5"] = U.interpolate_points(f2, dec, [R.BSplineLU(5, 32), R.BSplineLU(5, 24)])
    # unsplit Vlasov-Poisson (StdPoisson2d, ABTimeAlg_ip order 3), 3 steps on 32 x 40: final data and energies
    mesh_sp, mesh_v = R.UniformMesh(0.0, 4 * math.pi, 32), R.UniformMesh(-9.0, 9.0, 40)
    dt = 0.02
    adv = R.Advection((mesh_sp, mesh_v), [R.Lagrange(5), R.Lagrange(5)], dt, [([1, 2], 2, 1, False)], tab_coef=R.nosplit(dt),
                      timealg=R.ABTimeAlg_ip, ordalg=3)
    x, y = mesh_sp.points[:, None], mesh_v.points[None, :]
    data = np.asfortranarray(1 / math.sqrt(2 * math.pi) * np.exp(-0.5 * y**2) * (1 + 0.5 * np.cos(x / 2)))
    advd = R.AdvectionData(adv, data, U.getpoissonvar2d(adv))
    en = []
    for _ in range(3):
        while R.advection(advd):
            pass
        en.append(R.getenergy(advd))
    out["vp2d_ab3_data"] = np.array(advd.data)
    out["vp2d_ab3_energies"] = np.array(en)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
