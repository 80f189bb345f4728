"""Generate tests/golden/tsdf_case.npz from the UNMODIFIED reference's CPU mode of TSDF integration.

Build-container only.  ``TSDFVolume.__init__`` hard-imports PyCUDA (tsdf_fusion.py:33-35), which is not available
offline, so the object is created without running ``__init__`` and given exactly the attributes the constructor
would set in CPU mode (:44-71, :167-179); ``TSDFVolume.integrate`` itself (:221-306) and its numba helpers
(:181-218) run unmodified.  Inputs are regenerated from the seed by the tests (``tsdf_case_inputs``).
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("UFO_REFERENCE_ROOT", "/root/reference")


def tsdf_case_inputs(seed=0, n_views=3, hw=(48, 64)):
    """A small inward-looking rig around a bumpy sphere: depth maps, intrinsics, camera-to-world poses, bounds."""
    rng = np.random.default_rng(seed)
    H, W = hw
    K = np.array([[60.0, 0, W / 2 - 0.5], [0, 60.0, H / 2 - 0.5], [0, 0, 1]], dtype=np.float64)
    depths, poses = [], []
    for v in range(n_views):
        th = 0.5 * v - 0.4
        eye = 3.0 * np.array([np.sin(th), 0.15 * v, -np.cos(th)])
        z = -eye / np.linalg.norm(eye)
        x = np.cross(np.array([0, 1.0, 0]), z); x /= np.linalg.norm(x)
        y = np.cross(z, x)
        c2w = np.eye(4); c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = x, y, z, eye
        ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
        dirs = np.stack([(xs - K[0, 2]) / K[0, 0], (ys - K[1, 2]) / K[1, 1], np.ones_like(xs, dtype=float)], -1)
        dw = dirs @ c2w[:3, :3].T
        # ray / unit-sphere intersection (first hit), depth = z in the camera frame; misses -> 0 (invalid)
        b = dw @ eye; a = (dw * dw).sum(-1); c = eye @ eye - 1.0
        disc = b * b - a * c
        t = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / a, 0.0)
        depth = np.where(disc > 0, t, 0.0) + np.where(disc > 0, 0.02 * rng.standard_normal((H, W)), 0.0)
        depths.append(depth.astype(np.float32))
        poses.append(c2w.astype(np.float32))
    vol_bnds = np.array([[-1.3, 1.3], [-1.2, 1.2], [-1.3, 1.1]])
    return depths, [K.astype(np.float32)] * n_views, poses, vol_bnds, 0.08, 3


def main():
    sys.modules.setdefault("skimage", types.ModuleType("skimage"))
    sk_measure = types.ModuleType("skimage.measure")
    sys.modules["skimage"].measure = sk_measure
    sys.modules["skimage.measure"] = sk_measure
    sys.path.insert(0, REF)
    import tsdf_fusion as ref                         # the reference module

    depths, intrs, poses, vol_bnds, voxel_size, margin = tsdf_case_inputs()
    vol = object.__new__(ref.TSDFVolume)              # skip __init__ (hard PyCUDA import); CPU-mode attributes:
    vb = np.asarray(vol_bnds, dtype=np.float64).copy()
    vol._vol_bnds = vb
    vol._voxel_size = float(voxel_size)
    vol._trunc_margin = margin * vol._voxel_size
    vol._color_const = 256 * 256
    vol._vol_dim = np.round((vb[:, 1] - vb[:, 0]) / vol._voxel_size).copy(order="C").astype(int)
    vol._vol_bnds[:, 1] = vol._vol_bnds[:, 0] + vol._vol_dim * vol._voxel_size
    vol._vol_origin = vol._vol_bnds[:, 0].copy(order="C").astype(np.float32)
    vol._tsdf_vol_cpu = np.ones(vol._vol_dim).astype(np.float32)
    vol._weight_vol_cpu = np.zeros(vol._vol_dim).astype(np.float32)
    vol._color_vol_cpu = np.zeros(vol._vol_dim).astype(np.float32)
    vol.gpu_mode = 0
    xv, yv, zv = np.meshgrid(range(vol._vol_dim[0]), range(vol._vol_dim[1]), range(vol._vol_dim[2]), indexing="ij")
    vol.vox_coords = np.concatenate([xv.reshape(1, -1), yv.reshape(1, -1), zv.reshape(1, -1)], axis=0).astype(int).T
    snaps = []
    for d, K, P in zip(depths, intrs, poses):
        color = np.zeros(d.shape + (3,), dtype=np.float32)
        try:
            vol.integrate(color, d, K, P, obs_weight=1.0)
        except IndexError:
            # the CPU mode's colour update indexes the flattened colour image with two indices (tsdf_fusion.py:236 vs
            # :303) and raises; the TSDF and weight volumes were already updated (:293-294), which is all we take
            pass
        snaps.append((vol._tsdf_vol_cpu.copy(), vol._weight_vol_cpu.copy()))
    out = os.path.join(ROOT, "tests", "golden", "tsdf_case.npz")
    np.savez_compressed(out, tsdf=np.stack([s[0] for s in snaps]).astype(np.float32),
                        weight=np.stack([s[1] for s in snaps]).astype(np.float32), vol_dim=vol._vol_dim,
                        vol_origin=vol._vol_origin, trunc=np.float32(vol._trunc_margin))
    from oracle import tsdf_oracle as orc
    t = np.ones(vol._vol_dim, np.float32); w = np.zeros(vol._vol_dim, np.float32)
    for i, (d, K, P) in enumerate(zip(depths, intrs, poses)):
        t, w = orc.integrate(t, w, vol._vol_origin, voxel_size, vol._trunc_margin, d, K, P)
        dt = np.abs(t - snaps[i][0])
        print(f"view {i}: voxels {t.size}, updated {int((w > 0).sum())}, |dtsdf| > 1e-5 on {int((dt > 1e-5).sum())} voxels,"
              f" weight mismatches {int((w != snaps[i][1]).sum())}, max {dt.max():.3e}")
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
