"""GPU: TSDF integration through the C ABI (ufo_tsdf_integrate / uforecon_b200.tsdf.TSDFVolume) against the oracle.
fp32 arithmetic in the reference's order with one rounding per operation: the bar is bit-exact."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT
from oracle import tsdf_oracle as orc

sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_golden_tsdf import tsdf_case_inputs  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_views,hw", [(3, (48, 64)), (20, (40, 56))])      # 20 views: two launches (16 + 4)
def test_tsdf_bit_exact_vs_oracle(n_views, hw):
    from uforecon_b200.tsdf import TSDFVolume
    depths, intrs, poses, vol_bnds, voxel_size, margin = tsdf_case_inputs(seed=1, n_views=n_views, hw=hw)
    vol = TSDFVolume(vol_bnds, voxel_size, margin=margin)
    dim, origin, trunc = orc.volume_from_bounds(vol_bnds, voxel_size, margin)
    assert tuple(vol._vol_dim) == tuple(dim) and np.array_equal(vol._vol_origin, origin)
    t, w = np.ones(dim, np.float32), np.zeros(dim, np.float32)
    # first view through the reference-shaped single call, the rest fused
    vol.integrate(None, depths[0], intrs[0], poses[0], obs_weight=1.0)
    t, w = orc.integrate(t, w, origin, voxel_size, trunc, depths[0], intrs[0], poses[0])
    tv, cv, wv = vol.get_volume()
    assert np.array_equal(tv, t) and np.array_equal(wv, w) and not cv.any()
    vol.integrate_many(depths[1:], intrs[1:], poses[1:])
    t, w = orc.integrate_views(t, w, origin, voxel_size, trunc, depths[1:], intrs[1:], poses[1:])
    tv, _, wv = vol.get_volume()
    assert np.array_equal(wv, w)
    assert np.array_equal(tv, t)


def test_tsdf_golden_and_properties():
    """reference CPU-mode golden (same bound as the oracle test) + size-independent properties on a larger volume."""
    from uforecon_b200.tsdf import TSDFVolume
    g = np.load(os.path.join(GOLDEN, "tsdf_case.npz"))
    depths, intrs, poses, vol_bnds, voxel_size, margin = tsdf_case_inputs()
    vol = TSDFVolume(vol_bnds, voxel_size, margin=margin)
    vol.integrate_many(depths, intrs, poses)
    tv, _, wv = vol.get_volume()
    differ = (wv != g["weight"][-1]) | (np.abs(tv - g["tsdf"][-1]) > 1e-5)
    assert differ.mean() < 5e-3
    # larger volume (1.6 M voxels): weights are integer observation counts <= n_views, tsdf in [-1, 1], untouched voxels keep 1
    big = TSDFVolume(vol_bnds, voxel_size / 4, margin=margin)
    big.integrate_many(depths, intrs, poses)
    t, w = big.device_volumes()
    assert float(w.max()) <= len(depths) and torch.equal(w, w.round())
    assert float(t.max()) <= 1.0 and float(t[w > 0].min()) >= -1.0 - 1e-6
    assert torch.all(t[w == 0] == 1.0)
    # integrating nothing is a no-op; a second identical pass keeps tsdf (running average of equal values) within 1 ulp
    t0 = t.clone()
    big.integrate_many([], [], [])
    assert torch.equal(big.device_volumes()[0], t0)


def test_tsdf_errors():
    from uforecon_b200 import _lib
    from uforecon_b200.tsdf import TSDFVolume
    with pytest.raises(_lib.UfoError):
        TSDFVolume(np.array([[0, 1], [0, 1], [0, 1.0]]), 0.1, use_gpu=False)
    vol = TSDFVolume(np.array([[0, 1], [0, 1], [0, 1.0]]), 0.1)
    vol._grid.trunc_margin = 0.0
    with pytest.raises(_lib.UfoError):
        vol.integrate(None, np.ones((4, 4), np.float32), np.eye(3), np.eye(4))
