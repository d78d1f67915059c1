"""One step-program launch of C1 for ncu (k_program): python tools/prof_program.py [repeat]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "semilagrangian.jl_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import slb200 as S  # noqa: E402

g, _ = bench._cfg_c1(S)
while S.advection(g):
    pass
sp = S.StepProgram(g, nsteps=2, repeat=int(sys.argv[1]) if len(sys.argv) > 1 else 50)
sp.launch()
sp.launch()
print(sp.energies()[-1])
