"""GPU: device marching cubes (ufo_tsdf_mesh_*) - bit-exact against oracle/mc_oracle.py (positions, normals, indices
and their order), and through the TSDFVolume mirror after a fusion."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import mc_oracle as mc
from uforecon_b200 import _lib

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _volumes():
    rng = np.random.default_rng(3)
    out = {}
    # rows shorter than, equal to and longer than one 32-voxel group; single planes and single rows
    for shape in ((9, 8, 10), (33, 17, 40), (2, 2, 2), (6, 1, 6), (1, 1, 5), (5, 7, 132), (3, 3, 4), (6, 5, 256), (4, 1, 8)):
        vol = rng.standard_normal(shape).astype(np.float32)
        vol[rng.random(shape) < 0.05] = 0.0
        out["noise" + "x".join(map(str, shape))] = vol
    g = np.mgrid[0:70, 0:64, 0:61].astype(np.float32)
    out["sphere"] = (np.sqrt(((g - 30.3) ** 2).sum(0)) - 21.7).astype(np.float32)
    g = np.mgrid[0:40, 0:44, 0:160].astype(np.float32)
    out["sphere4"] = (np.sqrt(((g - np.array([19.2, 21.1, 80.4], np.float32)[:, None, None, None]) ** 2).sum(0)) - 17.3).astype(np.float32)
    out["empty"] = np.ones((8, 8, 8), dtype=np.float32)
    return out


@pytest.mark.parametrize("name", list(_volumes().keys()))
def test_mesh_bit_exact_vs_oracle(name):
    from uforecon_b200.tsdf import marching_cubes
    vol = _volumes()[name]
    level = 0.0 if not name.startswith("sphere") else 0.25
    v0, f0, n0 = mc.marching_cubes(vol, level)
    v, f, n = marching_cubes(torch.from_numpy(vol).cuda(), level)
    assert tuple(v.shape) == v0.shape and tuple(f.shape) == f0.shape
    assert np.array_equal(v.cpu().numpy(), v0)
    assert np.array_equal(f.cpu().numpy(), f0)
    assert np.array_equal(n.cpu().numpy(), n0)
    v2, f2, n2 = marching_cubes(torch.from_numpy(vol).cuda(), level, normals=False, faces=False)
    assert f2 is None and n2 is None and torch.equal(v2, v)


def test_fused_volume_to_mesh_and_ply(tmp_path):
    from make_golden_tsdf import tsdf_case_inputs
    from uforecon_b200 import formats
    from uforecon_b200.tsdf import TSDFVolume
    depths, intrs, poses, vol_bnds, voxel_size, margin = tsdf_case_inputs()
    tv = TSDFVolume(vol_bnds, voxel_size, margin=margin)
    tv.integrate_many(depths, intrs, poses)
    verts, faces, norms, colors = tv.get_mesh()
    t, _, _ = tv.get_volume()
    v0, f0, n0 = mc.marching_cubes(t, 0.0)
    assert len(verts) > 100 and np.array_equal(faces, f0) and np.array_equal(norms, n0)
    assert np.array_equal(verts, v0 * np.float32(voxel_size) + tv._vol_origin)       # tsdf_fusion.py:347
    assert colors.shape == (len(verts), 3) and colors.dtype == np.uint8 and not colors.any()
    # the fused surface is the unit sphere the depth maps were rendered from (2 cm noise, 8 cm voxels)
    seen = np.abs(np.linalg.norm(verts, axis=1) - 1.0)
    assert np.median(seen) < 0.05
    pc = tv.get_point_cloud()
    assert pc.shape == (len(verts), 6) and np.array_equal(pc[:, :3], verts)
    formats.meshwrite(str(tmp_path / "m.ply"), verts, faces, norms, colors)
    formats.pcwrite(str(tmp_path / "p.ply"), pc)
    head = (tmp_path / "m.ply").read_text().splitlines()
    assert head[2] == f"element vertex {len(verts)}" and head[12] == f"element face {len(faces)}"


def test_mesh_errors():
    lib = _lib.load()
    import ctypes as C
    g = _lib.UfoTsdfGrid()
    g.dim[:] = [4, 4, 4]
    mesh, nv, nf = C.c_void_p(), C.c_int64(), C.c_int64()
    assert lib.ufo_tsdf_mesh_begin(C.byref(g), None, 0.0, C.byref(mesh), C.byref(nv), C.byref(nf), None) != 0
    g.dim[:] = [0, 4, 4]
    t = torch.ones(4, 4, 4, device="cuda")
    assert lib.ufo_tsdf_mesh_begin(C.byref(g), t.data_ptr(), 0.0, C.byref(mesh), C.byref(nv), C.byref(nf), None) != 0
    assert lib.ufo_tsdf_mesh_emit(None, None, None, None, None) != 0
    lib.ufo_tsdf_mesh_destroy(None)
