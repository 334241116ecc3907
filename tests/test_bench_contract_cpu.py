"""The bench line of record (profiles/r2_bench.json, written by `python bench.py` on a B200) carries every key the
measurement contract asks for, with self-consistent values.  Guards the JSON contract against edits of bench.py; no GPU."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_line_of_record_has_the_contract_keys():
    d = _line("r2_bench.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "wgan_gd_train_steps_per_s" and d["unit"] == "steps/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    # value is steps / time of the timed region
    assert abs(d["value"] - d["n_gpus"] * 1000.0 / d["ms_per_step"]) / d["value"] < 1e-6
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 50e6 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"] * 1.02          # host copies inside the timed region cannot make it faster
    assert d["gpu_launches"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] > 0.8 * c["sm_max_mhz"]
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"])
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6 and 0.0 < r["frac"] < 1.0
    assert r["traffic"] is None or r["traffic"] > 0
    # the dominant family's launches fit inside the step
    fam = r["families"]["fwd_dgrad"]
    assert fam["ms_per_step"] < d["ms_per_step"]
    b = d["cpu_baseline"]
    assert b["kind"] in ("port", "reference") and b["cores"] >= 1 and b["value"] > 0 and b["sample"]


def test_data_parallel_line_carries_dp_check():
    d = _line("r2_bench_n2_ce.json")
    assert d["n_gpus"] == 2 and d["scaling"] == "weak"
    chk = d["dp_check"]
    assert chk["ok"] is True and chk["weights_identical"] is True and chk["bench_weights_identical_after_timed_steps"] is True
    # tolerance of the check itself: at most twice torch-bf16's own deviation
    assert chk["worst_grad_cos"] >= 1.0 - 2.0 * (1.0 - chk["bf16_cos"]) - 0.01
