"""bench.py's host logic (timed loop, roofline choice, e2e legs, parity and configs objects) dry-run on the CPU
over the test double of the C ABI (tests/fake_device.py): a tiny 2D2V grid, the oracle's line algorithms
standing in for the kernels.  Guards the JSON contract, not any performance number."""
import argparse
import json

import fake_device


def test_bench_line_contract(monkeypatch, capsys):
    fake_device.install(monkeypatch)
    import bench

    args = argparse.Namespace(gpus=1, steps=2, warmup=3, impl="ours", size=8, order=7, interp="lagrange", no_cpu=False,
                              no_configs=True, no_e2e=False, exchange="p2p")
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)
    assert bench.run_ours(args) == 0
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
                "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "parity", "ee_after_timed", "steps_done"):
        assert key in line, key
    assert line["steps_done"] == 5 and line["parity"]["steps"] == 5
    # the test double IS the oracle's arithmetic: the histories agree exactly and the resident run equals a fresh one
    assert line["parity"]["ee_hist_rel_vs_oracle"] <= 1e-12
    assert line["parity"]["ee_after_timed_equals_fresh_run"] is True
    assert line["roofline"]["kernel"].startswith("sweep/")  # no pair fusion on the double: slowest single sweep
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}


def test_reference_arm_line(capsys):
    import bench

    args = argparse.Namespace(gpus=1, steps=1, warmup=3, impl="reference", size=8, order=7, interp="lagrange")
    assert bench.run_reference(args) == 0
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["steps_done"] == 4 and line["ee_after_timed"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_configs_object(monkeypatch):
    fake = fake_device.install(monkeypatch)
    import bench
    import slb200 as S

    args = argparse.Namespace(size=8)
    out = bench.run_configs(S, S.default_context(), args, 6541.5, only=("C1",))
    c1 = out["C1_vp1d1v_128x256_L9_strang"]
    assert "error" not in c1, c1
    assert c1["parity"]["f_rel_maxabs"] <= 1e-12 and c1["parity"]["ee_rel"] <= 1e-12 and c1["ms_per_step"] > 0
