"""CPU: the marching-cubes oracle (generated case table, numpy restatement) and the PLY writers.

scikit-image - what the reference calls - is not available: the oracle is held to the properties any correct
iso-surface extraction has (closed, consistently oriented, one vertex per sign-changing grid edge, converging area and
volume on an analytic shape), the committed CUDA table header to the generator, and the PLY writers to the bytes the
reference's own writers produce (tests/golden/ply_case.npz, tools/make_golden_ply.py)."""
import os
import sys

import numpy as np

from conftest import GOLDEN, ROOT
from oracle import mc_oracle as mc
from uforecon_b200 import formats

sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_case_table_properties():
    T = mc.build_tables()
    assert int(T["max_tri"]) == 5 and T["ntri"][0] == 0 and T["ntri"][255] == 0
    for k in range(256):
        used = {int(e) for e in T["tri"][k] if e >= 0}
        cross = {e for e in range(12) if ((k >> mc.edge_corners(e)[0]) & 1) != ((k >> mc.edge_corners(e)[1]) & 1)}
        assert used == cross, k                                # every crossing edge carries a vertex, no other edge does


def test_committed_cuda_table_matches_generator():
    import gen_mc_table
    with open(os.path.join(ROOT, "uforecon_b200", "csrc", "ufo_mc_table.cuh")) as f:
        assert f.read() == gen_mc_table.render()


def test_random_sign_volumes_give_closed_oriented_surfaces():
    """white noise visits all 256 cases, ambiguous faces included; exact zeros and flat regions too"""
    rng = np.random.default_rng(0)
    for shape in ((9, 8, 10), (5, 12, 7), (2, 2, 2), (6, 1, 6)):
        vol = rng.standard_normal(shape).astype(np.float32)
        vol[rng.random(shape) < 0.05] = 0.0
        v, f, n = mc.marching_cubes(vol)
        inside = vol < 0
        n_cross = sum(int((np.take(inside, range(0, s - 1), a) != np.take(inside, range(1, s), a)).sum()) for a, s in enumerate(shape))
        assert len(v) == n_cross
        assert mc.mesh_is_closed(f, v, shape)
        assert np.isfinite(v).all() and np.isfinite(n).all()
        assert (v >= 0).all() and (v <= np.array(shape) - 1).all()


def test_sphere_area_volume_orientation():
    N, c, r = 48, 23.3, 15.7
    g = np.mgrid[0:N, 0:N, 0:N].astype(np.float32)
    vol = np.sqrt(((g - c) ** 2).sum(0)) - r
    v, f, n = mc.marching_cubes(vol)
    tri = v[f] - c
    cr = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    assert (np.einsum("ij,ij->i", cr, tri.mean(1)) > 0).all()                       # wound towards larger f (outwards)
    assert abs(0.5 * np.linalg.norm(cr, axis=1).sum() / (4 * np.pi * r * r) - 1) < 5e-3
    assert abs(np.einsum("ij,ij->i", tri[:, 0], np.cross(tri[:, 1], tri[:, 2])).sum() / 6 / (4 / 3 * np.pi * r ** 3) - 1) < 5e-3
    rad = (v - c) / np.linalg.norm(v - c, axis=1, keepdims=True)
    assert np.einsum("ij,ij->i", n, rad).min() > 0.999                              # gradient normals
    assert np.abs(np.linalg.norm(v - c, axis=1) - r).max() < 0.02                   # linear interpolation error
    assert mc.mesh_is_closed(f, v, vol.shape)


def test_ply_writers_match_reference_bytes(tmp_path):
    from make_golden_ply import ply_case_inputs
    g = np.load(os.path.join(GOLDEN, "ply_case.npz"))
    verts, faces, norms, colors = ply_case_inputs()
    formats.meshwrite(str(tmp_path / "m.ply"), verts, faces, norms, colors)
    formats.pcwrite(str(tmp_path / "p.ply"), np.hstack([verts, colors.astype(np.float32)]))
    assert (tmp_path / "m.ply").read_bytes() == g["mesh"].tobytes()
    assert (tmp_path / "p.ply").read_bytes() == g["cloud"].tobytes()
    formats.meshwrite(str(tmp_path / "e.ply"), np.zeros((0, 3)), np.zeros((0, 3), int), np.zeros((0, 3)), np.zeros((0, 3), np.uint8))
    assert b"element vertex 0" in (tmp_path / "e.ply").read_bytes()
