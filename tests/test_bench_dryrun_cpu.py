"""bench.py's own Python, end to end, without a GPU: tests/emu/bench_dryrun.py runs the real run_ours() over the
host-emulated kernel library on a shrunken workload; here the ONE JSON line it prints is checked against the
bench contract (keys, types, the roofline / cpu_baseline / e2e objects) and against the canned autotune report
(the verified-and-faster K1 variant is switched on, the K3 variant that failed its parity check is not)."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_json_contract_over_emulator():
    r = subprocess.run([sys.executable, os.path.join(REPO, "tests", "emu", "bench_dryrun.py")], capture_output=True,
                       text=True, timeout=600, cwd=REPO)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]  # exactly ONE JSON line
    out = json.loads(lines[0])
    for key, typ in (("metric", str), ("value", float), ("unit", str), ("n_gpus", int), ("steps", int), ("warmup", int),
                     ("ms_per_step", float), ("higher_is_better", bool), ("scaling", str), ("dtype", str),
                     ("data", str), ("config", dict), ("e2e", dict), ("gpu_launches", int), ("clocks", dict),
                     ("roofline", dict), ("cpu_baseline", dict)):
        assert isinstance(out[key], typ), (key, out.get(key))
    assert out["vs_baseline"] is None and out["n_gpus"] == 1 and out["steps"] == 2 and out["warmup"] == 3
    assert out["metric"] == "train samples/sec EigenWorms Log-NCDE" and out["unit"] == "samples/s"
    assert out["scaling"] == "weak" and out["dtype"] == "f32" and out["data"] == "synthetic"
    assert "workload" in out["config"] and "model" not in out["config"]
    assert out["gpu_launches"] == 14 * out["steps"]  # K1, K5 (5), solve, sweep + 3 GEMM groups + reduce, Adam (2)
    # value = whole-job samples over the timed region
    assert abs(out["value"] - out["config"]["global_batch"] / (out["ms_per_step"] * 1e-3)) < 1e-6 * out["value"]
    e2e = out["e2e"]
    assert e2e["unit"] == "samples/s" and e2e["value"] > 0 and e2e["d2h_bytes_per_step"] == 4
    assert e2e["h2d_bytes_per_step"] == 4 * (2 * 120 * 7 + 2 * 5)  # the batch of paths + its labels, counted from the tensors
    rf = out["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and rf["peak"] > 0
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    K, D = 11, 29  # zero row prepended: ceil(121 / 12) windows
    assert rf["algorithmic_bytes_per_launch"] == 4 * (4 * 120 * 7 + 4 * K * D)  # 4LC + 4KD per sample
    assert abs(rf["achieved"] - rf["algorithmic_bytes_per_launch"] / (rf["launch_ms"] * 1e-3) / 1e9) < 1e-9
    cb = out["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and cb["unit"] == "samples/s" and cb["sample"]
    rd = out["roofline_dominant"]
    assert rd["peak_source"].startswith("measured") and rd["peak"] > 0 and abs(rd["frac"] - rd["achieved"] / rd["peak"]) < 1e-12
    bd = out["step_breakdown_ms"]
    assert set(bd) == {"logsig_K1", "forward_solve_K3+readout", "backward_K3b+gemm", "adam_K4", "whole_step"}
    assert "kernels" in out["impl_detail"] and "autotune" not in out  # kernels are chosen by shape inside the library
    assert "TMA" in rf["kernel"]


def test_reference_arm_line():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference"], capture_output=True,
                       text=True, timeout=120, cwd=REPO, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""  # ranks other than 0 exit 0 without work


def test_lncde_sweep_over_emulator():
    """scripts/lncde_sweep.py (BASELINE configs[4]: DE vs D LNCDE on depth-3 rows): one JSON line per point + a summary."""
    r = subprocess.run([sys.executable, os.path.join(REPO, "tests", "emu", "sweep_dryrun.py")], capture_output=True,
                       text=True, timeout=600, cwd=REPO)
    assert r.returncode == 0, r.stderr[-3000:]
    recs = [json.loads(ln) for ln in r.stdout.splitlines() if ln.startswith("{")]
    points, summary = recs[:-1], recs[-1]
    assert [(p["model"], p["T"]) for p in points] == [("diagonal_linear_ncde", 9), ("dense_linear_ncde", 9),
                                                      ("diagonal_linear_ncde", 70), ("dense_linear_ncde", 70)]
    for p in points:
        assert "error" not in p and "skipped" not in p and p["depth"] == 3 and p["ms_per_step"] > 0
        assert abs(p["samples_per_s"] - p["B_per_gpu"] / (p["ms_per_step"] * 1e-3)) < 1e-6 * p["samples_per_s"]
        assert p["hbm"]["algorithmic_bytes"] > 0 and ("flow_build" in p) == (p["model"] == "dense_linear_ncde")
    assert points[0]["logsig_dim"] == 15 and points[2]["logsig_dim"] == 6  # Hall dimensions of (3, 3) and (2, 3)
    assert set(summary["ratio"]) == {"B2_T9_H16_C3", "B2_T70_H8_C2"}
