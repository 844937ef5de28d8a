"""The bench line contract, checked without a GPU: the committed lines of the GPU arm (profiles/) and a live run of the
reference arm carry every key the driver reads, with consistent arithmetic between them."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e")
N_COMPOSITED = 97


def _last_json_line(path):
    lines = [ln for ln in open(path).read().splitlines() if ln.strip().startswith("{")]
    assert lines, path
    return json.loads(lines[-1])


def _check_common(d):
    for k in BASE_KEYS:
        assert k in d, k
    assert d["unit"] == "ray-samples/s" and d["higher_is_better"] is True and d["data"] == "synthetic"
    assert d["vs_baseline"] is None                       # BASELINE.md publishes no number for this metric
    assert "workload" in d["config"] and "model" not in d["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    if d["n_gpus"] == 1 or "cpu_baseline" in d:          # timed on rank 0 at N = 1 only
        cb = d["cpu_baseline"]
        for k in ("value", "unit", "cores", "kind", "sample"):
            assert k in cb, k
        assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1


def test_committed_gpu_bench_lines_keep_the_contract():
    sys.path.insert(0, ROOT)
    import bench
    from i2sdf_b200 import configs
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02*_bench_train*.json")))
    assert files, "no round-2 bench line under profiles/"
    for f in files:
        d = _last_json_line(f)
        _check_common(d)
        assert "impl" not in d or d["impl"] != "reference"
        assert d["gpu_launches"] > 0
        assert d["warmup"] >= 3
        # value = rays of all ranks x 97 composited samples per timed second (whole job, not per GPU)
        rays = d["config"]["rays_per_gpu"] * d["n_gpus"]
        assert abs(d["value"] - rays * N_COMPOSITED / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"], f
        assert d["config"]["global_rays"] == rays
        e2e = d["e2e"]
        assert e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0 and e2e["value"] != d["value"]
        r = d["roofline"]
        for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
            assert k in r, (f, k)
        assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        # achieved = ALGORITHMIC flops per launch / measured launch time (never the 3x the tensor cores execute)
        assert abs(r["achieved"] - r["flop_per_launch"] / (r["ms_per_launch"] * 1e-3) / 1e12) < 1e-6 * r["achieved"]
        sdf_eval = bench.flop_model(configs.model_conf(d["config"]["network_config"]))["sdf_eval"]      # 918 016 for synthetic.yml
        assert r["flop_per_launch"] == d["config"]["rays_per_gpu"] * 128 * sdf_eval
        c = d["clocks"]
        assert c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"]
        assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"]), f
        assert "l2" in d["config"] and "flush" in d["config"]["l2"]
        if d["n_gpus"] > 1:
            s = d["strong_scaling"]
            assert s["global_rays"] == d["config"]["rays_per_gpu"] and s["rays_per_gpu"] * d["n_gpus"] == s["global_rays"]
            assert abs(s["value"] - s["global_rays"] * N_COMPOSITED / (s["ms_per_step"] * 1e-3)) < 1e-6 * s["value"]


def test_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` on a bounded sample (8 rays, one step): one JSON line, the reference-arm keys, zero-byte e2e."""
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--cpu-rays", "8", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    _check_common(d)
    assert d["impl"] == "reference" and d["dtype"] == "f32" and d["config"]["mode"] == "train"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port"
    assert d["config"]["same_config"] is False            # 8 of the 1024 rays: the sample is named as such
    assert abs(d["value"] - 8 * N_COMPOSITED / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    # the other ranks of a torchrun launch exit at once without output
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                        capture_output=True, text=True, timeout=300, env=dict(env, RANK="1", WORLD_SIZE="2"), cwd=ROOT)
    assert r1.returncode == 0 and r1.stdout.strip() == ""
