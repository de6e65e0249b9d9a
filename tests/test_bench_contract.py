"""bench.py's reference arm runs without a GPU (it times the oracle): its one JSON line must carry the contract's keys and
say what it ran.  (The b200 arm needs a B200; its line is checked by the driver.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--frames-per-link", "4"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                            # ONE JSON line on stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Msamples/s" and d["higher_is_better"] is True and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"]
    # the line states what was run: the links and frames of this run, not the b200 arm's
    cfg = d["config"]
    assert cfg["links_per_gpu"] == cb["cores"] and cfg["frames_per_link"] == 4 and "bounded_sample_of" in cfg
    n_samples = cfg["links_per_gpu"] * (128 + 4 * (4961 + 1100))
    assert cfg["samples_per_gpu"] == n_samples and str(n_samples) in cb["sample"]
    assert d["crc_ok_per_step"] >= d["frames_per_step"] - 2           # 64-QAM 3/4 at 30 dB


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
