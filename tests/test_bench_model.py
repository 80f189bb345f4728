"""CPU: the algorithmic-work model bench.py reports rooflines against."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def test_flops_per_point_matches_survey():
    # 3 584 + 131 400 (NV+1) + 158 900 + 6 688 + 2 928 NV
    for nv, total in ((3, 703_600), (5, 972_200), (10, 1_643_900)):
        f = bench.flops_per_point(nv)
        assert abs(f["view"] - 131_200 * (nv + 1)) <= 400 * (nv + 1)      # 2 MAC (8 d^2 + attention) per token
        assert abs(f["total"] - total) / total < 5e-3, (nv, f["total"])


def test_tap_bytes_match_survey():
    assert bench.tap_bytes_per_point(3) == 7392
    assert bench.tap_bytes_per_point(5) == 17440
    assert bench.tap_bytes_per_point(10) == 60480
