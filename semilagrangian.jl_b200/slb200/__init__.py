"""slb200 -- Python host layer of the B200 semi-Lagrangian sweep library (libslb200.so).

Mirrors the exported surface of SemiLagrangian.jl (src/SemiLagrangian.jl:44-64) for the hot
path: UniformMesh, Lagrange / BSplineLU / BSplineFFT / Hermite, the splitting tables,
Advection / AdvectionData / advection (== advection!), and the Poisson / rotation /
translation displacement providers.  Everything numerical runs in the CUDA library through
its C ABI (include/slb200.h); there is no CPU fallback.
"""
from ._lib import Context, SlbError, default_context, LIB_PATH, SLB_SWEEP_EXACT, SLB_SWEEP_INSIDE_EDGE
from .mesh import UniformMesh, step, points, width, start, stop, vec_k_fft
from .interp import (AbstractInterpolation, Lagrange, BSplineLU, BSplineFFT, Hermite, get_order, get_kl_ku,
                     LAGRANGE, BSPLINE_LU, BSPLINE_FFT, HERMITE, CircEdge, InsideEdge)
from .splitting import (nosplit, standardsplit, strangsplit, magicsplit, triplejumpsplit, order6split,
                        hamsplit_3_11)
from .advection import (Advection, AdvectionData, AbstractExtDataAdv, StateAdv, advection, getdata, sizeall,
                        sweep, sweep_pair, modone, invperm, StepGraph, StepProgram)
from .sharded import HaloShardedAdvectionData, HaloUnsupported, local_group
from .poisson import (PoissonVar, getpoissonvar, compute_ee, compute_ke, getenergy, getenergyall, dotprod,
                      StdPoisson, StdPoisson2d)
from .unsplit2d import (NoTimeAlg, ABTimeAlg_ip, ABTimeAlg_new, ABTimeAlg_init, DeviceField, interpolate_points,
                        autointerp, interpbufc, abcoef)
from .rotation import RotationVar, getrotationvar
from .translation import TranslationVar, gettranslationvar
from .quasigeostrophic import GeoVar, getgeovar
from .interpolate import interpolate, interpolate_lines, interpolate_nd, sol
