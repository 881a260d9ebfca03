"""bench.py's contract pieces that need no GPU: the algorithmic-FLOP model, the `config` object shared by both
arms, the roofline denominators and the ncu traffic lookup from the committed summary."""
import importlib.util
import json
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_module", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_algorithmic_flops_per_view_match_the_survey(bench):
    # totals per view 2.03 T (V=2), 2.45 T (V=8), 3.55 T (V=24), 8.80 T (V=100), 71.0 T (V=1000)
    for v, want in ((2, 2.03), (8, 2.45), (24, 3.55), (100, 8.80), (1000, 71.0)):
        assert abs(bench.gflop_per_view(v) / 1e3 - want) < 0.02 * want, (v, bench.gflop_per_view(v))


def test_config_object_is_shared_by_both_arms_and_names_the_baseline_config(bench):
    c1 = bench.bench_config(8, 1, "shard", False, False)
    assert c1 == bench.bench_config(8, 1, "shard", False, False)
    assert c1["views"] == 8 and c1["scene_views"] == 8 and c1["parallelism"] == "single" and "model" not in c1
    assert "8 views" in c1["workload"] and "config[1]" in c1["workload"]
    c8 = bench.bench_config(8, 8, "shard", False, True)
    assert c8["views"] == 64 and c8["scene_views"] == 64 and "all-gather" in c8["parallelism"]
    assert abs(c8["tflop_per_view"] - bench.gflop_per_view(64) / 1e3) < 1e-9
    rep = bench.bench_config(8, 4, "replicas", False, False)
    assert rep["scene_views"] == 8 and rep["views"] == 32 and rep["parallelism"] == "replicas x4"
    mm = bench.bench_config(24, 1, "shard", True, False)
    assert "multi-modal" in mm["workload"] and "config[2]" in mm["workload"]
    json.dumps(c8)   # plain JSON


def test_roofline_denominators_and_traffic_lookup(bench):
    tf, hbm, src = bench.measured_peaks()
    assert src in ("measured", "fallback") and 1000 < tf < 2300 and 5000 < hbm < 8100
    traffic, note = bench.ncu_gemm_traffic()
    # the committed ncu summary of the dominant GEMM launch: within 1.5x of its algorithmic 120.6 MB
    assert traffic is not None and 0.5 * 120.6e6 < traffic < 1.5 * 120.6e6, (traffic, note)
    assert "profiles/" in note
