"""Times the pieces of one field solve (charge partial sums, one-kernel Poisson solve: cluster-FFT vs cooperative-DFT
form) with CUDA events on the library's stream.  One JSON line."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "semilagrangian.jl_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import slb200 as S  # noqa: E402
from slb200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ctx = S.default_context()
L = _lib.lib()
adv, vecs = bench.vp2d2v_setup(S, n, 7, "lagrange")
pv = S.getpoissonvar(adv)
rng = np.random.default_rng(0)
ls = ctx.to_device(rng.random(n * n * n))      # line sums: [n1 n2, n3]
e0, e1 = ctx.event(), ctx.event()
arr = (C.c_void_p * 2)(*[p.value for p in pv.E_dev])


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    ctx.sync()
    ctx.record(e0)
    for _ in range(reps):
        fn()
    ctx.record(e1)
    return _lib.Context.elapsed_ms(e0, e1) / reps * 1e3


out = {}
for fft in ("1", "0"):
    os.environ["SLB_FIELD_FFT"] = fft
    tag = "fft" if fft == "1" else "coop_dft"
    out[f"vp_field_solve_from_linesums_us[{tag}]"] = timed(lambda: _lib.check(L.slb_vp_field_solve(pv.plan, ls, n, 1.0, pv.rho_dev, arr)))
    out[f"poisson_solve_raw_us[{tag}]"] = timed(lambda: _lib.check(L.slb_poisson_solve_raw(pv.plan, ls, 1, arr)))
os.environ["SLB_FIELD_FFT"] = "1"
out["charge_density_from_linesums_us"] = timed(lambda: _lib.check(L.slb_charge_density_from(ctx.h, ls, n * n, n, 1.0, pv.rho_dev, 0)))
out["empty_launch_pair_us"] = timed(lambda: _lib.check(L.slb_subtract_mean(ctx.h, pv.rho_dev, 16)))
print(json.dumps(out))
