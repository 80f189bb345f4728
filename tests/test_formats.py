"""CPU: on-disk formats (row N4): cam.txt round trip (and, in the build container, through the reference's own parser),
depth result npy as extract_geometry writes / save_tsdf reads it."""
import os
import sys
import types

import numpy as np
import pytest

from uforecon_b200 import formats, synthetic

REF = "/root/reference"


def test_cam_file_round_trip(tmp_path):
    rig = synthetic.make_rig()
    p = str(tmp_path / "cameras" / "00000016_cam.txt")
    formats.write_cam_file(p, rig[16], synthetic.DTU_K, synthetic.DTU_DEPTH_MIN, synthetic.DTU_DEPTH_INTERVAL)
    c = formats.read_cam_file(p)
    assert np.array_equal(c["extrinsic"], rig[16].astype(np.float32))
    assert np.array_equal(c["intrinsic"], synthetic.DTU_K.astype(np.float32))
    assert c["depth_min"] == 425.0 and c["depth_interval"] == 2.5 and c["depth_max"] == 425.0 + 2.5 * 192
    k4 = np.eye(4, dtype=np.float32)
    k4[:3, :3] = c["intrinsic"]
    assert np.array_equal(c["P"], k4 @ c["extrinsic"])


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout only exists in the build container")
def test_cam_file_read_by_reference_parser(tmp_path):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import ref_shim
    ref_shim.install()
    from code1.dataset.dtu_test_sparse import DtuFitSparse        # the reference's parser, unmodified
    rig = synthetic.make_rig()
    p = str(tmp_path / "00000001_cam.txt")
    formats.write_cam_file(p, rig[1], synthetic.DTU_K)
    holder = types.SimpleNamespace()
    P = DtuFitSparse.read_cam_file(holder, p)
    mine = formats.read_cam_file(p)
    assert np.array_equal(P, mine["P"])
    assert holder.depth_min == mine["depth_min"] and holder.depth_interval == mine["depth_interval_scaled"]


def test_depth_result_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    depth = (500 + 100 * rng.random((12, 16))).astype(np.float32)
    rgb = rng.random((12, 16, 3)).astype(np.float32)
    E = synthetic.make_rig()[3].astype(np.float32)
    K = synthetic.DTU_K.astype(np.float32)
    path = formats.save_depth_result(str(tmp_path), "scan24", "00000003", depth, rgb, E, K, npy_name="refview{view}.npy")
    assert path.endswith(os.path.join("depth", "scan24", "refview00000003.npy"))
    d, k, pose = formats.load_depth_result(path)
    assert np.array_equal(d, depth) and np.array_equal(k, K) and np.allclose(pose @ E, np.eye(4), atol=1e-4)
    from PIL import Image
    png = np.array(Image.open(tmp_path / "scan24" / "depth" / "00000003.png"))
    assert png.shape == (12, 16) and png.max() == 255
    assert np.array(Image.open(tmp_path / "rgb" / "scan24" / "00000003.jpg")).shape == (12, 16, 3)


def test_pair_file_roundtrip_and_reference_layout(tmp_path):
    """the two-lines-per-viewpoint layout of dtu_pairs.txt (dtu_train.py:171-178), incl. the trailing blank of each line"""
    p = tmp_path / "pair.txt"
    p.write_text("3\n0\n3 10 2346.410000 1 2036.530000 9 1243.890000 \n1\n2 9 2850.870000 10 2583.940000 \n7\n0 \n")
    pairs = formats.read_pair_file(str(p))
    assert pairs == {0: [10, 1, 9], 1: [9, 10], 7: []}
    q = tmp_path / "pair2.txt"
    formats.write_pair_file(str(q), pairs, {0: [2346.41, 2036.53, 1243.89], 1: [2850.87, 2583.94], 7: []})
    assert formats.read_pair_file(str(q)) == pairs
    assert q.read_text().splitlines()[2] == "3 10 2346.410000 1 2036.530000 9 1243.890000 "
