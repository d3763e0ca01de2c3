"""bench.py's reference arm runs without a GPU: check the JSON line it prints against the contract
(keys, metric / unit / config identical to the GPU arm's, e2e = value, no device bytes)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "subspace_detector_template_samples_per_sec"
    assert d["unit"] == "template*samples/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    cfg = d["config"]
    assert d["value"] > 1e5 and abs(d["ms_per_step"] * d["value"] / 1e3 - cfg["lags_per_chunk"] *
                                    cfg["subspaces_per_step"] * cfg["chunks_per_step"]) < 1e-3 * d["value"]
    # the unmodified reference where it is available (oracle/_ref or /root/reference), else the oracle port
    sys.path.insert(0, ROOT)
    from oracle import ref_shim
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_shim.available() else "port")
    assert 1 <= d["cpu_baseline"]["cores"] <= (os.cpu_count() or 1)
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert 0 < d["cpu_baseline"]["one_process_value"] <= d["value"] * 1.5
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert cfg["subspaces"] == 256 and cfg["n"] == 9000 and "workload" in cfg and "tcgen05" not in json.dumps(cfg)


def test_other_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]
