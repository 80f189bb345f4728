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


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line with the contract keys,
    at a reduced ray grid so that the CPU suite stays fast."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ref-chunks", "1", "--width", "416", "--height", "320"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "rays_per_sec" and d["unit"] == "rays/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]


def test_roofline_entries_bookkeeping():
    """tensor kernels: algorithmic FLOPs per launch; the gather: DRAM bytes of the ncu capture (only when the configuration is
    the captured one), tap bytes reported beside it - never tap bytes against the HBM peak under the name of a roofline."""
    import argparse
    import bench
    n_rays = 1000
    prof = [("k_view_tc<3, 0>", 4, 8.0), ("k_ray_tc<128, 0>", 4, 6.0), ("k_gather_tc<3, 0>", 4, 6.0), ("k_render<kNS>", 2, 0.1)]
    pk = {"tf_sustained": 1394.5, "hbm_gbs": 6546.6, "src": "test"}
    t = bench._ncu_traffic()
    for cfg_matches in (True, False):
        args = argparse.Namespace(nv=3, steps=2, mode="tc16", width=1600 if cfg_matches else 416, height=1216, rays=0)
        roofs = bench.roofline_entries(prof, args, n_rays, pk)
        by = {r["kernel"].split("<")[0]: r for r in roofs}
        assert set(by) == {"k_view_tc", "k_ray_tc", "k_gather_tc"} and roofs[0]["kernel"].startswith("k_view_tc")
        fl = bench.flops_per_point(3)
        assert abs(by["k_view_tc"]["algorithmic_work_per_launch"] - n_rays * 128 * 2 * (fl["view"] + fl["radiance"]) / 4) < 1
        assert abs(by["k_ray_tc"]["algorithmic_work_per_launch"] - n_rays * 192 * 2 * (fl["ray"] + fl["density"]) / 4) < 1
        assert by["k_gather_tc"]["bound"] == "hbm" and by["k_view_tc"]["bound"] == "tensor"
        assert abs(by["k_gather_tc"]["tap_bytes_per_launch"] - n_rays * 128 * 2 * bench.tap_bytes_per_point(3) / 4) < 1
        if cfg_matches and "k_gather_tc" in t:
            assert by["k_gather_tc"]["traffic"] == t["k_gather_tc"]["bytes_per_launch"]
            assert by["k_gather_tc"]["algorithmic_work_per_launch"] == t["k_gather_tc"]["bytes_per_launch"]
            assert by["k_ray_tc"]["traffic"] == t["k_ray_tc"]["bytes_per_launch"] and "ncu" in by["k_view_tc"]
        else:
                # no capture for this configuration: the bytes the launch must write (a lower bound of its DRAM traffic), never tap bytes
                assert by["k_gather_tc"]["traffic"] is None and "lower bound" in by["k_gather_tc"]["note"]
                assert abs(by["k_gather_tc"]["algorithmic_work_per_launch"] - n_rays * 128 * 2 * 3 * 192 / 4) < 1
                assert by["k_gather_tc"]["frac"] < 1.0
        assert abs(sum(r["share_of_step"] for r in roofs) - 20.0 / 20.1) < 1e-9
