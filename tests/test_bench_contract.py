"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the agreed keys, and the
static workload description matches BASELINE.json's headline config."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "waveglow_infer_audio_samples_per_sec" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "64 x" in d["config"]["workload"]
    # the line says what it really ran: a bounded sample (one 80x100 mel per step), and it honours --steps / --warmup
    assert "bounded sample" in d["config"]["workload"] and "1 x 80x100" in d["config"]["workload"]
    assert d["config"]["sample_per_step"] == {"batch": 1, "frames": 100, "samples": 25600}
    assert d["steps"] == 1 and d["warmup"] == 1 and d["steps_requested"] == 1


def test_ncu_traffic_is_read_from_the_newest_profile():
    sys.path.insert(0, ROOT)
    import bench
    traffic, src = bench.ncu_traffic("pair_kernel<3, 0, 0>", 64)
    assert traffic and 2e9 < traffic < 8e9 and src.startswith("profiles/") and "_ncu_full_summary.csv" in src
    half, _ = bench.ncu_traffic("pair_kernel<3, 0, 0>", 32)
    assert abs(half - traffic / 2) < 1.0
    assert bench.ncu_traffic("no_such_kernel", 64)[0] is None


def test_reference_arm_non_zero_ranks_exit_silently():
    env = dict(os.environ, RANK="3", WORLD_SIZE="8", LOCAL_RANK="3")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "8"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_workload_matches_baseline_json():
    sys.path.insert(0, ROOT)
    import bench
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        base = json.load(f)
    assert "64 × 80×860" in base["configs"][2].replace("\\u00d7", "×")
    assert bench.GLOBAL_BATCH == 64 and bench.FRAMES == 860 and abs(bench.SIGMA - 0.666) < 1e-9
    assert bench.GATE_FLOP_PER_STEP == 2 * 2176 * 1024
