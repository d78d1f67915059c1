"""Runs the single-GPU (in-process ranks) halo checks of tests/test_gpu_halo.py over a list of cases and reports
every outcome instead of stopping at the first failure (debug aid)."""
import os
import sys
import traceback

os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "semilagrangian.jl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import test_gpu_halo as T

cases = [(8, 7, (64, 64, 64, 64)), (4, 7, (64, 64, 64, 64)), (8, 7, (32, 8, 16, 64)), (8, 7, (64, 8, 16, 64)), (4, 7, (64, 8, 16, 32)),
         (2, 7, (64, 8, 16, 32)), (8, 7, (128, 4, 16, 128))]
for P, order, sz in cases:
    class _Env:   # stands in for pytest's monkeypatch
        @staticmethod
        def setenv(k, v):
            os.environ[k] = v

    for name, fn, extra in (("passes", T.test_halo_passes_bitwise_equal_unsharded, ()),
                            ("steps/split", T.test_halo_sharded_steps_match_single_grid_and_oracle, (2, "1", _Env)),
                            ("steps/in-pass", T.test_halo_sharded_steps_match_single_grid_and_oracle, (2, "0", _Env))):
        try:
            fn(P, order, sz, *extra)
            print(f"P={P} order={order} sz={sz} {name}: OK", flush=True)
        except AssertionError:
            tb = traceback.extract_tb(sys.exc_info()[2])[-1]
            print(f"P={P} order={order} sz={sz} {name}: FAIL at line {tb.lineno}: {tb.line}", flush=True)
        except Exception as exc:
            print(f"P={P} order={order} sz={sz} {name}: ERROR {type(exc).__name__}: {exc}", flush=True)
