"""bench.py's bookkeeping (no GPU): the algorithmic-bytes figures of SURVEY.md 8(d), the workload table, the CLI."""
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _w(B, H, W, K, kpt):
    return dict(B=B, C=3, H=H, W=W, K=K, kpt=kpt)


def test_algorithmic_bytes_match_survey_8d():
    # SURVEY.md 8(d) "A per config (fp32)": main-only and main + 9 keypoint planes
    assert bench.algorithmic_bytes_per_image(_w(32, 96, 320, 50, 0))[0] == 402_444
    assert bench.algorithmic_bytes_per_image(_w(32, 96, 320, 50, 9))[0] == 1_542_564
    assert bench.algorithmic_bytes_per_image(_w(256, 96, 320, 100, 0))[0] == 436_244
    assert bench.algorithmic_bytes_per_image(_w(256, 96, 320, 100, 9))[0] == 1_610_564
    assert bench.algorithmic_bytes_per_image(_w(128, 192, 640, 100, 0))[0] == 1_542_164
    assert bench.algorithmic_bytes_per_image(_w(128, 192, 640, 100, 9))[0] == 6_034_244
    # bf16 hand-off halves only the heat-map term
    full, main, kpt = bench.algorithmic_bytes_per_image(_w(256, 96, 320, 100, 9), elem=2)
    assert full == 1_610_564 - 12 * 96 * 320 * 2 and main + kpt == full


def test_workloads_are_the_baseline_configs():
    assert bench.WORKLOADS["cfg4"]["B"] == 256 and bench.WORKLOADS["cfg4"]["kpt"] == 9 and bench.WORKLOADS["cfg4"]["K"] == 100
    assert (bench.WORKLOADS["cfg5"]["H"], bench.WORKLOADS["cfg5"]["W"], bench.WORKLOADS["cfg5"]["B"]) == (192, 640, 128)
    assert bench.WORKLOADS["cfg2"]["B"] == 32 and bench.WORKLOADS["cfg2"]["K"] == 50
    assert bench.WORKLOADS["cfg3"]["kpt"] == 0 and bench.WORKLOADS["cfg3"]["reg"] == 8            # configs[2] as written: 8 regression channels
    assert bench.WORKLOADS["cfg2x"]["B"] == 64 and bench.WORKLOADS["cfg2x"]["kpt"] == 9            # north_star's batch >= 64 point


def test_per_kernel_bytes_add_up_to_the_step():
    # scan kernel = the heat-maps, select + post kernel = gathers + outputs: together the algorithmic bytes of the step
    w = _w(256, 96, 320, 100, 9)
    kb = bench.kernel_bytes_per_image(w)
    assert kb["scan"] == 12 * 96 * 320 * 4 and kb["scan"] + kb["post"] == bench.algorithmic_bytes_per_image(w)[0]


def test_cli_lists_the_contract_flags():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--workload", "--dtype", "--no-e2e", "--max-ctas", "--scaling", "--gather", "--graph", "--verify"):
        assert flag in out.stdout, flag
