"""The bench line committed under profiles/ carries every key the measurement contract names (CPU test)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


import pytest


@pytest.mark.parametrize("line", ["r01_bench_c2_end.json", "r02_bench_default_final.json", "r02b_bench_default_final.json"])
def test_committed_bench_line_has_the_contract_keys(line):
    d = json.load(open(os.path.join(ROOT, "profiles", line)))
    if line.startswith("r02"):
        # round 2: the DRAM traffic is labelled as the constant it is, every rank reports its own sweep time and clocks, and
        # the end-to-end leg says what the host link gave inside the call
        assert d["roofline"]["traffic_source"].startswith("constant from profiles/") and d["kernel"] == "tc4"
        assert d["ranks"][0]["kernel_ms_median"] > 0 and "sm_mhz" in d["ranks"][0]
        e = d["e2e"]
        assert e["h2d_gbps_in_call"] > 0 and e["h2d_gbps_link_alone"] > 0 and e["slowest_stage"] in e["stage_seconds"]
        assert set(d["config"]) == {"workload", "samples", "variants_per_gpu", "phenotypes", "covariates", "missing_rate", "groups",
                                    "parallelism", "l2"}
    if line.startswith("r02b"):
        # second session: what the sweeps issued and the tensor roof next to the HBM roof; the host phases of every repetition
        t = d["roofline"]["tensor"]
        assert t["bound"] == "tensor" and t["unit"] == "TFLOP/s" and (t["sweep_launches"], t["mma_columns"]) == (1, 80)
        assert abs(t["frac"] - t["achieved"] / t["peak"]) < 1e-3 and d["roofline"]["binding"] == "hbm"
        # multiply-adds issued = variants x padded samples x MMA columns; achieved = 2 x that / the sweep's own duration
        assert t["flop_per_launch"] == 2.0 * 1_000_000 * 400_384 * 80
        assert abs(t["achieved"] - t["flop_per_launch"] / (d["roofline"]["kernel_ms"] / 1e3) / 1e12) < 0.1
        assert len(d["e2e"]["host_phase_ms_per_rep"]) == d["e2e"]["reps"] == len(d["e2e"]["rep_seconds"])
        assert max(d["e2e"]["rep_seconds"]) < 1.05 * min(d["e2e"]["rep_seconds"])     # no heavy tail left
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "gpu_launches", "roofline", "clocks", "e2e", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "genotypes/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["config"]["workload"].startswith("C2") and d["vs_baseline"] is None
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert r["traffic"] is None or r["traffic"] >= r["algorithmic_bytes_per_launch"]
    # achieved = algorithmic bytes / the sweep kernel's own duration
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["kernel_ms"] / 1e3) / 1e9) < 1.0
    # value = genotypes of one step / its duration
    assert abs(d["value"] - 400000 * 1e6 / (d["ms_per_step"] / 1e3)) / d["value"] < 1e-6
    assert d["gpu_launches"] >= d["steps"]
    e = d["e2e"]
    assert e["unit"] == "genotypes/s" and e["h2d_bytes_per_step"] > 1e9 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in d["clocks"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_algorithmic_bytes_formula_matches_survey():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    # M ceil(N/4) + M G (4 + 8 + 40 P) + G 8 (nK + nP + KP + P)
    assert b.algorithmic_bytes(1_000_000, 400_000, 1, 10, 1, 400_000) == 100_087_200_088
