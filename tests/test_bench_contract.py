"""bench.py's reference arm (--impl reference: the oracle port of the reference step on the host cores) prints ONE JSON
line with the keys the driver reads.  Runs on CPU; the GPU arm is exercised by the driver itself."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "base_64",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines                         # stdout carries exactly the JSON line
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_images_per_sec" and d["unit"] == "images/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("vae-gan base 64x64")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_workloads_name_the_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    assert set(bench.WORKLOADS) >= {"v2_128", "base_64", "unet_256", "unet_256_z512", "oldv_64x448"}
    assert set(bench.STEP_GFLOP_PER_IMG) == set(bench.WORKLOADS)
    wl = bench.WORKLOADS["v2_128"]
    assert (wl["h"], wl["w"], wl["batch"], wl["family"]) == (128, 128, 64, "v2")      # BASELINE configs[1]
