"""Host-side mirror of the reference's ``TSDFVolume`` (tsdf_fusion.py:21-356) on top of ``ufo_tsdf_integrate``.

Same constructor arguments, attribute names and ``integrate`` / ``get_volume`` signatures as the reference class;
volumes live on the device.  ``integrate_many`` fuses the per-view loop of ``save_tsdf`` (tsdf_fusion.py:486-502) into
launches of up to 16 views.  ``get_mesh`` / ``get_point_cloud`` (tsdf_fusion.py:319-356) run marching cubes on the
device (``ufo_tsdf_mesh_*``) instead of copying the volumes to scikit-image; ``formats.meshwrite`` / ``pcwrite`` write the
reference's PLY files.  There is no CPU fallback: without the CUDA library this module raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib


def marching_cubes(volume: torch.Tensor, level: float = 0.0, normals: bool = True, faces: bool = True):
    """``ufo_tsdf_mesh_begin`` / ``_emit`` on a device volume [X,Y,Z] fp32: (verts [Nv,3] f32 in voxel coordinates,
    faces [Nf,3] i32 | None, normals [Nv,3] f32 | None) as device tensors - what skimage's
    ``marching_cubes_lewiner(volume, level)`` returns, minus its ``values``."""
    lib = _lib.load()
    if not (volume.is_cuda and volume.dtype == torch.float32 and volume.dim() == 3):
        raise ValueError("marching_cubes: volume must be a 3-D float32 CUDA tensor")
    vol = volume.contiguous()
    grid = _lib.UfoTsdfGrid()
    grid.dim[:] = [int(x) for x in vol.shape]
    grid.voxel_size, grid.trunc_margin = 1.0, 1.0
    mesh = C.c_void_p()
    nv, nf = C.c_int64(), C.c_int64()
    with torch.cuda.device(vol.device):
        st = torch.cuda.current_stream(vol.device)
        _lib.check(lib.ufo_tsdf_mesh_begin(C.byref(grid), vol.data_ptr(), float(level), C.byref(mesh), C.byref(nv), C.byref(nf),
                                           st.cuda_stream))
        try:
            v = torch.empty(nv.value, 3, dtype=torch.float32, device=vol.device)
            n = torch.empty(nv.value, 3, dtype=torch.float32, device=vol.device) if normals else None
            f = torch.empty(nf.value, 3, dtype=torch.int32, device=vol.device) if faces else None
            _lib.check(lib.ufo_tsdf_mesh_emit(mesh, v.data_ptr(), n.data_ptr() if normals else None, f.data_ptr() if faces else None,
                                              st.cuda_stream))
            st.synchronize()
        finally:
            lib.ufo_tsdf_mesh_destroy(mesh)
    return v, f, n


class TSDFVolume:
    def __init__(self, vol_bnds, voxel_size: float, use_gpu: bool = True, margin: int = 5, device=None):
        if not use_gpu:
            raise _lib.UfoError("uforecon_b200.tsdf.TSDFVolume has no CPU mode")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda")
        vol_bnds = np.asarray(vol_bnds, dtype=np.float64).copy()
        assert vol_bnds.shape == (3, 2), "[!] `vol_bnds` should be of shape (3, 2)."
        self._vol_bnds = vol_bnds
        self._voxel_size = float(voxel_size)
        self._trunc_margin = margin * self._voxel_size                               # tsdf_fusion.py:49
        self._vol_dim = np.round((vol_bnds[:, 1] - vol_bnds[:, 0]) / self._voxel_size).copy(order="C").astype(int)
        self._vol_bnds[:, 1] = self._vol_bnds[:, 0] + self._vol_dim * self._voxel_size
        self._vol_origin = self._vol_bnds[:, 0].copy(order="C").astype(np.float32)
        dims = tuple(int(x) for x in self._vol_dim)
        self._tsdf_vol = torch.ones(dims, dtype=torch.float32, device=self.device)   # :57
        self._weight_vol = torch.zeros(dims, dtype=torch.float32, device=self.device)
        self._grid = _lib.UfoTsdfGrid()
        self._grid.dim[:] = dims
        self._grid.origin[:] = [float(x) for x in self._vol_origin]
        self._grid.voxel_size = self._voxel_size
        self._grid.trunc_margin = self._trunc_margin

    def _view(self, depth_im, cam_intr, cam_pose, keep):
        d = torch.as_tensor(depth_im).to(self.device, torch.float32).contiguous()
        keep.append(d)
        v = _lib.UfoTsdfView()
        v.depth, v.im_h, v.im_w = d.data_ptr(), int(d.shape[0]), int(d.shape[1])
        v.intr[:] = [float(x) for x in np.asarray(cam_intr, dtype=np.float32).reshape(-1)[:9]]
        v.pose[:] = [float(x) for x in np.asarray(cam_pose, dtype=np.float32).reshape(-1)[:16]]
        return v

    def integrate_many(self, depth_ims: Sequence, cam_intrs: Sequence, cam_poses: Sequence, obs_weight: float = 1.0):
        """Integrate the depth maps in order (identical to calling ``integrate`` once per view)."""
        keep: list = []
        n = len(depth_ims)
        arr = (_lib.UfoTsdfView * n)(*[self._view(d, k, p, keep) for d, k, p in zip(depth_ims, cam_intrs, cam_poses)])
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device)
            _lib.check(self.lib.ufo_tsdf_integrate(C.byref(self._grid), self._tsdf_vol.data_ptr(), self._weight_vol.data_ptr(), arr, n,
                                                   float(obs_weight), st.cuda_stream))
            st.synchronize()          # `keep` (device copies of the depth maps) may be released after this

    def integrate(self, color_im, depth_im, cam_intr, cam_pose, obs_weight: float = 1.0):
        """``TSDFVolume.integrate`` (tsdf_fusion.py:221).  ``color_im`` is accepted and ignored: the reference's kernel
        returns before its colour update (:137), so its colour volume stays zero."""
        self.integrate_many([depth_im], [cam_intr], [cam_pose], obs_weight)

    def get_volume(self):
        """(tsdf, color, weight) as numpy arrays, like the reference (:308-313); colour is all zeros."""
        t = self._tsdf_vol.cpu().numpy()
        return t, np.zeros_like(t), self._weight_vol.cpu().numpy()

    def device_volumes(self):
        return self._tsdf_vol, self._weight_vol

    def extract_surface(self, level: float = 0.0, normals: bool = True, faces: bool = True):
        """Marching cubes of the fused volume on the device, see :func:`marching_cubes`."""
        return marching_cubes(self._tsdf_vol, level, normals, faces)

    def _vertex_colors(self, verts_vox: torch.Tensor) -> np.ndarray:
        # tsdf_fusion.py:346-354 decode color_vol at round(verts); the reference's kernel never writes color_vol
        # (`return` before the colour branch, :137), so every decoded colour is 0
        return np.zeros((verts_vox.shape[0], 3), dtype=np.uint8)

    def get_mesh(self):
        """``TSDFVolume.get_mesh`` (tsdf_fusion.py:340-356): (verts [Nv,3] world, faces [Nf,3], norms [Nv,3], colors
        [Nv,3] uint8) as numpy arrays."""
        v, f, n = self.extract_surface()
        origin = torch.from_numpy(self._vol_origin).to(self.device)
        verts = v * self._voxel_size + origin                                       # :347
        return verts.cpu().numpy(), f.cpu().numpy(), n.cpu().numpy(), self._vertex_colors(v)

    def get_point_cloud(self):
        """``TSDFVolume.get_point_cloud`` (tsdf_fusion.py:319-338): [Nv, 6] = xyz (world) | rgb."""
        v, _, _ = self.extract_surface(normals=False, faces=False)
        origin = torch.from_numpy(self._vol_origin).to(self.device)
        verts = (v * self._voxel_size + origin).cpu().numpy()
        return np.hstack([verts, self._vertex_colors(v).astype(verts.dtype)])
