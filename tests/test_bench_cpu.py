"""CPU: the roofline accounting of bench.py reproduces the algorithmic-byte budgets of SURVEY.md section 8(d), and the
reference arm / argument handling work without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_algorithmic_bytes_match_the_survey():
    sys.path.insert(0, ROOT)
    import bench
    b = bench.algorithmic_bytes_per_px
    # headline: fwd + bwd = 36 + 24 * n_src = 84 B/px at n_src = 2, one scale
    assert abs(b(bench.WORKLOADS["headline"]) - 84.0) < 1e-9
    # reference-live composition (2 directions), S = 1: 144 B/px-step
    assert abs(b(bench.WORKLOADS["c1"]) - 144.0) < 1e-9
    # ... fused over scales at S = 4: "about 152 B/px-step"
    assert abs(b(bench.WORKLOADS["c2"]) - 151.875) < 1e-9
    # 3 sources (C3): 108 B/px for the 3-source direction at one scale, + 60 for the single-source one
    c3 = dict(bench.WORKLOADS["c3"], n_scales=1)
    assert abs(b(c3) - (108.0 + 60.0)) < 1e-9


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--steps", "3", "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mpix/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["gpu_launches"] == 0 and line["higher_is_better"] is True
