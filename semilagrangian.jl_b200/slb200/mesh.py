"""UniformMesh -- host mirror of src/mesh.jl:21-115 (paths relative to the reference root)."""
from fractions import Fraction
import math

import numpy as np


class UniformMesh:
    """UniformMesh(start, stop, length): `length` nodes start + i*(stop-start)/length
    (src/mesh.jl:26-32).  Nodes are the exact rational values rounded once to Float64, which
    is what Julia's twice-precision `range` delivers."""

    def __init__(self, start, stop, length):
        start, stop, length = float(start), float(stop), int(length)
        if length < 1:
            raise ValueError("length must be positive")
        lo, hi = Fraction(start), Fraction(stop)
        h = (hi - lo) / length
        self.points = np.fromiter((float(lo + i * h) for i in range(length)), dtype=np.float64, count=length)
        self.step = float(h)
        self.width = stop - start

    def __len__(self):
        return self.points.shape[0]


def step(mesh):
    return mesh.step


def points(mesh):
    return mesh.points


def width(mesh):
    return mesh.width


def start(mesh):
    return mesh.points[0]


def stop(mesh):
    return mesh.points[-1] + mesh.step


def vec_k_fft(mesh):
    """src/mesh.jl:110-115: (2 pi / width) .* fftfreq(n, n)"""
    n = len(mesh)
    k = 2 * math.pi / mesh.width
    idx = np.arange(n, dtype=np.float64)
    idx[(n + 1) // 2:] -= n
    return k * idx
