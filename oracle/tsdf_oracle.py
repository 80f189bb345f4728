"""ORACLE - CPU restatement of the reference's TSDF ``integrate`` CUDA kernel.  TEST INFRASTRUCTURE ONLY.

Only ``tests/`` and ``bench.py``'s CPU-baseline leg may import this module.

Follows the CUDA-C source string in ``tsdf_fusion.py:77-152`` of the reference line by line, in numpy float32
with one rounding per arithmetic operation (no FMA contraction): voxel index -> grid coordinates (:96-99) ->
world (:101-104) -> camera (:106-111, ``cam_pose`` is camera-to-world; the kernel applies R^T (p - t)) -> pixel
with ``roundf`` = round half away from zero (:113-114) -> frustum / depth tests (:116-127) -> truncated distance
and running average (:128-134).  The colour branch is dead code in the reference (``return`` at :137) and the
colour volume stays zero.  The reference's guard ``voxel_idx > X*Y*Z`` (:93) lets one out-of-range thread through;
the restatement (and the library) stop at ``>=``.

Parity pinning: PyCUDA is not available offline, so the CUDA string itself cannot be executed here.  The oracle is
pinned against the reference's own CPU mode - the ``else`` branch of ``TSDFVolume.integrate`` (:267-306) with its
numba helpers (:181-218) - run unmodified by ``tools/make_golden_tsdf.py`` (fixture ``tests/golden/tsdf_case.npz``).
That branch computes the camera transform in float64 and rounds pixels half-to-even, so it agrees with the kernel
arithmetic everywhere except at voxels that project within rounding distance of a pixel boundary or of the
truncation threshold; ``tests/test_tsdf_oracle.py`` bounds that set (< 0.5 % of the voxels) and requires 1e-5
agreement elsewhere.
"""
from __future__ import annotations

import numpy as np

F = np.float32


def _roundf(x: np.ndarray) -> np.ndarray:
    """C roundf: half away from zero, evaluated in float32."""
    x = x.astype(F)
    r = np.rint(x)                                        # half to even ...
    tr = np.trunc(x)
    tie = np.abs(x - tr) == F(0.5)                        # ... except exact ties, which go away from zero
    return np.where(tie, tr + np.sign(x), r).astype(F)


def integrate(tsdf: np.ndarray, weight: np.ndarray, vol_origin, voxel_size: float, trunc_margin: float,
              depth_im: np.ndarray, cam_intr: np.ndarray, cam_pose: np.ndarray, obs_weight: float = 1.0):
    """One call of the reference kernel over the whole volume.  tsdf, weight [X,Y,Z] float32 (updated copies are
    returned); depth_im [H,W]; cam_intr [3,3]; cam_pose [4,4] camera-to-world."""
    X, Y, Z = tsdf.shape
    im_h, im_w = depth_im.shape
    org = np.asarray(vol_origin, dtype=F)
    K = np.asarray(cam_intr, dtype=F).reshape(3, 3)
    P = np.asarray(cam_pose, dtype=F).reshape(4, 4)
    vs, tm, ow = F(voxel_size), F(trunc_margin), F(obs_weight)
    d = np.asarray(depth_im, dtype=F).reshape(-1)
    vx, vy, vz = np.meshgrid(np.arange(X, dtype=F), np.arange(Y, dtype=F), np.arange(Z, dtype=F), indexing="ij")
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        px = org[0] + vx * vs
        py = org[1] + vy * vs
        pz = org[2] + vz * vs
        tx, ty, tz = px - P[0, 3], py - P[1, 3], pz - P[2, 3]
        cx = (P[0, 0] * tx + P[1, 0] * ty) + P[2, 0] * tz            # tsdf_fusion.py:109-111, left to right
        cy = (P[0, 1] * tx + P[1, 1] * ty) + P[2, 1] * tz
        cz = (P[0, 2] * tx + P[1, 2] * ty) + P[2, 2] * tz
        fx = _roundf(K[0, 0] * (cx / cz) + K[0, 2])
        fy = _roundf(K[1, 1] * (cy / cz) + K[1, 2])
        ok = np.isfinite(fx) & np.isfinite(fy) & (fx >= 0) & (fx < im_w) & (fy >= 0) & (fy < im_h) & ~(cz < 0)
        ix = np.where(ok, fx, 0).astype(np.int64)
        iy = np.where(ok, fy, 0).astype(np.int64)
        dv = d[iy * im_w + ix]
        ok &= dv != 0
        diff = dv - cz
        ok &= ~(diff < -tm)
        dist = np.minimum(F(1.0), diff / tm).astype(F)
        w_new = (weight + ow).astype(F)
        t_new = (((tsdf * weight).astype(F) + (ow * dist).astype(F)).astype(F) / w_new).astype(F)
    out_t = np.where(ok, t_new, tsdf).astype(F)
    out_w = np.where(ok, w_new, weight).astype(F)
    return out_t, out_w


def integrate_views(tsdf, weight, vol_origin, voxel_size, trunc_margin, depths, intrs, poses, obs_weight=1.0):
    """Sequential integration of several depth maps (the loop of save_tsdf, tsdf_fusion.py:486-502)."""
    for dpt, K, P in zip(depths, intrs, poses):
        tsdf, weight = integrate(tsdf, weight, vol_origin, voxel_size, trunc_margin, dpt, K, P, obs_weight)
    return tsdf, weight


def volume_from_bounds(vol_bnds, voxel_size: float, margin: int = 5):
    """Grid parameters as ``TSDFVolume.__init__`` derives them (tsdf_fusion.py:44-60)."""
    vol_bnds = np.asarray(vol_bnds, dtype=np.float64).copy()
    vol_dim = np.round((vol_bnds[:, 1] - vol_bnds[:, 0]) / float(voxel_size)).astype(int)
    origin = vol_bnds[:, 0].astype(np.float32)
    return vol_dim, origin, float(margin) * float(voxel_size)
