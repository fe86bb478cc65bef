"""Host logic of bench.py on CPU: the per-class / per-shape tables and the roofline object built from profiled records,
the reference arm's rank gating, and the size of the JSON line (the driver keeps a 2 kB tail).  No GPU, no kernels."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class Ev:
    """Stand-in for torch.cuda.Event: elapsed_time in milliseconds between two stamps."""

    def __init__(self, t):
        self.t = t

    def elapsed_time(self, other):
        return other.t - self.t


def rec(name, tflop, ms, t0=0.0, executed=None):
    fl = tflop * 1e12
    return (name, fl, Ev(t0), Ev(t0 + ms), fl if executed is None else executed * 1e12)


def test_tables_split_the_launch_classes_and_compute_fractions(bench):
    pk = {"tflops": 1000.0, "hbm": 6000.0}
    prof = [
        rec("conv3x3 M=1 N=192 K=1728 act=0 res=0 rs=0 rope=0", 1.0, 1.0, executed=0.9),          # 1000 TFLOP/s
        rec("linear M=1 N=384 K=384 act=0 res=1 rs=0 rope=0", 0.5, 1.0),                           #  500
        rec("conv3x3_dgrad M=1 N=192 K=1728 act=0 res=0 rs=0 rope=0 +gn_bwd_reduce", 1.0, 2.0),    #  500, own class
        rec("wgrad linear M=1 N=384 K=384", 0.4, 1.0),
        rec("attn_fwd", 0.3, 1.0),
        rec("attn_bwd", 0.75, 3.0),
    ]
    ctab, stab = bench.tensor_tables(prof, 1, pk)
    by = {r["kernel"]: r for r in ctab}
    assert set(by) == {"mtgemm (fwd + dgrad)", "mtgemm dgrad + GroupNorm-backward reduce pass (fused epilogue)", "wgrad",
                       "attn_fwd", "attn_bwd"}
    assert by["mtgemm (fwd + dgrad)"]["launches"] == 2 and abs(by["mtgemm (fwd + dgrad)"]["frac_of_tensor_peak"] - 0.75) < 1e-6
    assert abs(by["attn_bwd"]["achieved_tflops"] - 250.0) < 1e-6 and abs(by["wgrad"]["ms"] - 1.0) < 1e-9
    assert len(stab) == 6 and stab[0]["launch"] == "attn_bwd"                # sorted by time
    roof = bench.roofline_of(prof, 1, 10.0, pk)
    assert roof["bound"] == "tensor" and roof["unit"] == "TFLOP/s" and roof["peak"] == 1000.0
    assert abs(roof["frac"] - 0.75) < 1e-6                                   # plain launches only: 1.5 TFLOP in 2 ms
    assert abs(roof["frac_executed"] - 0.70) < 1e-6                          # (0.9 + 0.5) TFLOP in 2 ms
    assert abs(roof["frac_with_fused_gn_bwd"] - 0.625) < 1e-6                # 2.5 TFLOP in 4 ms
    assert abs(roof["share_of_step"] - 0.4) < 1e-9                           # all mtgemm launches: 4 of 10 ms
    # without fused launches the extra key is absent
    assert "frac_with_fused_gn_bwd" not in bench.roofline_of(prof[:2], 1, 10.0, pk)


def test_hbm_table(bench):
    pk = {"tflops": 1000.0, "hbm": 6000.0}
    tab = bench.hbm_table([("gn_apply_silu", 3e9, Ev(0.0), Ev(1.0)), ("gn_apply_silu", 3e9, Ev(1.0), Ev(2.0)),
                           ("adamw", 1e9, Ev(0.0), Ev(0.5))], 2, pk)
    by = {r["kernel"]: r for r in tab}
    assert by["gn_apply_silu"]["launches"] == 1 and abs(by["gn_apply_silu"]["achieved_gbs"] - 3000.0) < 1e-6
    assert abs(by["gn_apply_silu"]["frac_of_hbm_peak"] - 0.5) < 1e-6 and abs(by["adamw"]["ms"] - 0.25) < 1e-9


def test_workload_strings_name_the_baseline_configs(bench):
    for name, cfg in bench.CONFIGS.items():
        s = bench.workload_string(name, cfg)
        assert f"configs[{cfg['baseline_cfg']}]" in s and f"@{cfg['res']}^2" in s and len(s) < 100
        assert bench.metric_name(cfg).startswith("images_per_sec_")


def test_reference_arm_is_silent_on_other_ranks():
    """Under torchrun rank 0 alone runs the CPU arm; every other rank exits 0 without output and without work."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_committed_bench_line_fits_the_drivers_tail():
    """The final line of the round (profiles/r2fin_bench.json) carries every contract key and stays under 2 kB."""
    p = os.path.join(ROOT, "profiles", "r2fin_bench.json")
    if not os.path.exists(p):
        pytest.skip("no committed bench line")
    line = [l for l in open(p).read().splitlines() if l.startswith("{")][-1]
    assert len(line) < 2000, len(line)
    d = json.loads(line)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["gpu_launches"] > 0 and d["vs_baseline"] is None
