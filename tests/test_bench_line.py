"""The committed bench line (profiles/r02_bench_n1_c3.json, the output of `python bench.py --steps 20 --warmup 3` on a B200) carries
every key the measurement contract names, with the types and the internal consistency the driver relies on."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    rows = [x for x in open(os.path.join(ROOT, "profiles", name)) if x.startswith("{")]
    assert len(rows) == 1, "bench.py prints ONE JSON line"
    return json.loads(rows[0])


def test_single_gpu_line_carries_the_contract():
    l = _line("r02_bench_n1_c3.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "configs_other", "parity_basis"):
        assert k in l, k
    assert l["n_gpus"] == 1 and l["higher_is_better"] is True and l["scaling"] == "weak" and l["vs_baseline"] is None
    assert l["unit"] == "Gsamples/s" and l["dtype"] == "f32" and l["data"] == "synthetic" and "workload" in l["config"]
    assert l["warmup"] >= 3 and l["gpu_launches"] > 0
    # value is samples / time of exactly `steps` steps
    s = l["config"]["active_samples_per_step_per_gpu"]
    assert abs(l["value"] - s / (l["ms_per_step"] * 1e-3) / 1e9) <= 1e-6 * l["value"]
    r = l["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "l2", "issue") and abs(r["frac"] - r["achieved"] / r["peak"]) <= 1e-6
    assert abs(r["frac"] - r["roof_ms"] / r["kernel_ms"]) <= 1e-6 and r["roof_ms"] == max(r["roof_times_ms"].values())
    assert "STALE" not in r["traffic_source"]                 # the committed capture is of the kernels that produced the line
    c = l["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == l["unit"] and c["sample"]
    e = l["e2e"]
    assert e["unit"] == l["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= 1.02 * l["value"]
    k = l["clocks"]
    assert k["sm_mhz"] and k["sm_max_mhz"] and not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert {c["config"] for c in l["configs_other"]} >= {"c1", "c2", "c3gray", "c4", "c5"}


def test_multi_gpu_lines_scale():
    one = _line("r02_bench_n1_c3.json")
    for n in (2, 8):
        l = _line(f"r02_bench_n{n}_c3.json")
        assert l["n_gpus"] == n and l["scaling"] == "weak" and l["metric"] == one["metric"]
        assert 0.9 * n * one["value"] <= l["value"] <= 1.02 * n * one["value"]          # whole-job aggregate
        assert l["allreduce"]["ms_isolated"] > 0 and l["strong_scaling"]
