"""C1 only (1D1V 128 x 256, Lagrange 9, Strang): stepwise, CUDA-graph and step-program timings through bench.run_configs.
One JSON line.  Usage: python tools/bench_c1.py [blocks ...]  (SLB_PROGRAM_BLOCKS values to try)"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "semilagrangian.jl_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import slb200 as S  # noqa: E402

ctx = S.default_context()
args = argparse.Namespace(size=128)
res = {}
for spec in (sys.argv[1:] or ["32"]):   # "32" or "32cg" (cooperative_groups grid.sync instead of the counter barrier)
    os.environ["SLB_PROGRAM_BLOCKS"] = spec.rstrip("cg")
    os.environ["SLB_PROGRAM_CGSYNC"] = "1" if spec.endswith("cg") else "0"
    out = bench.run_configs(S, ctx, args, 6541.5, only=["C1"], with_oracle=False)["C1_vp1d1v_128x256_L9_strang"]
    res[f"blocks_{spec}"] = {k: out.get(k) for k in ("ms_per_step", "ms_per_step_graph", "ms_per_step_program", "program_equals_stepwise_bitwise",
                                                      "program", "program_error", "error")}
print(json.dumps(res))
