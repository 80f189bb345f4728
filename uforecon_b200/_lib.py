"""ctypes binding of ``libuforecon_b200.so`` (the C ABI in ``include/uforecon_b200.h``).

The library is built in-tree by ``uforecon_b200.build`` / ``__graft_entry__.build()``.  There is no
Python or CPU fallback: if the shared object is missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UFO_LIB_PATH", os.path.join(_HERE, "libuforecon_b200.so"))

UFO_MAX_VIEWS = 10
UFO_N_STAGES = 3
UFO_N_COARSE = 64
UFO_N_FINE = 64
UFO_N_SAMPLES = 128
UFO_MODE_FP32 = 0
UFO_MODE_TC = 1
UFO_MODE_TC_F16 = 2

c_float_p = C.POINTER(C.c_float)


class UfoSceneDesc(C.Structure):
    _fields_ = [
        ("n_views", C.c_int32), ("img_h", C.c_int32), ("img_w", C.c_int32), ("feat_h", C.c_int32), ("feat_w", C.c_int32),
        ("source_imgs", C.c_void_p), ("img_feats", C.c_void_p), ("depth_info", C.c_void_p), ("match_feats", C.c_void_p),
        ("vol_feat", C.c_void_p * UFO_N_STAGES), ("vol_weight", C.c_void_p * UFO_N_STAGES),
        ("vol_d", C.c_int32 * UFO_N_STAGES), ("vol_h", C.c_int32 * UFO_N_STAGES), ("vol_w", C.c_int32 * UFO_N_STAGES),
        ("source_poses", C.c_void_p), ("source_poses_inv", C.c_void_p), ("ref_pose_inv", C.c_void_p),
        ("w2cs", C.c_void_p), ("near_fars", C.c_void_p), ("ray_o", C.c_void_p),
        ("ray_d", C.c_void_p), ("cam_ray_d", C.c_void_p), ("match_pairs", C.c_void_p),
    ]


class UfoLoftrLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("q", "k", "v", "merge", "mlp0", "mlp2", "norm1_w", "norm1_b", "norm2_w", "norm2_b")]


class UfoMlp3(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w0", "b0", "w2", "b2", "w4", "b4")]


class UfoWeightsDesc(C.Structure):
    _fields_ = [
        ("view", UfoLoftrLayer), ("ray", UfoLoftrLayer),
        ("pre_sim", UfoMlp3), ("density", UfoMlp3), ("radiance", UfoMlp3),
        ("view_token", C.c_void_p), ("depth_freqs", C.c_void_p), ("depth_phases", C.c_void_p),
        ("variance", C.c_float),
    ]


class UfoDebugTaps(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("z_coarse", "weight_coarse", "srdf_coarse", "z_fine", "sim8", "vol24", "tokens",
                                           "view_tok0", "ray_out", "radiance", "weight")]


class UfoRenderOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("depth", "depth_z", "rgb", "srdf", "z", "points")]


class UfoPixelwiseNet(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("conv0_w", "bn0_w", "bn0_b", "bn0_mean", "bn0_var", "conv1_w", "bn1_w", "bn1_b",
                                           "bn1_mean", "bn1_var", "conv2_w")] + [("conv2_b", C.c_float)]


class UfoTsdfGrid(C.Structure):
    _fields_ = [("dim", C.c_int32 * 3), ("origin", C.c_float * 3), ("voxel_size", C.c_float), ("trunc_margin", C.c_float)]


class UfoTsdfView(C.Structure):
    _fields_ = [("depth", C.c_void_p), ("im_h", C.c_int32), ("im_w", C.c_int32), ("intr", C.c_float * 9), ("pose", C.c_float * 16)]


class UfoProfileEntry(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("launches", C.c_int64), ("ms", C.c_double)]


#: every symbol the header declares (tests check the built library exports all of them)
EXPORTS = (
    "ufo_abi_version", "ufo_last_error", "ufo_device_info", "ufo_weights_create", "ufo_weights_destroy",
    "ufo_scene_create", "ufo_scene_destroy", "ufo_scene_device_bytes", "ufo_render_rays", "ufo_render_rays_host",
    "ufo_launch_count", "ufo_costvolume_stage", "ufo_costvolume_stage_rt", "ufo_debug_umma_selftest", "ufo_profile_begin", "ufo_profile_end",
    "ufo_tsdf_integrate", "ufo_feature_grid", "ufo_tsdf_mesh_begin", "ufo_tsdf_mesh_emit", "ufo_tsdf_mesh_destroy",
    "ufo_trim_pool",
)

_lib = None


class UfoError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library; raises if it has not been built (no fallback by design)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise UfoError(f"{LIB_PATH} not found: build it with `python -m uforecon_b200.build` "
                       f"(there is no CPU/PyTorch fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    lib.ufo_abi_version.restype = C.c_int
    lib.ufo_last_error.restype = C.c_char_p
    lib.ufo_device_info.argtypes = [C.POINTER(C.c_int32)] * 3
    lib.ufo_weights_create.argtypes = [C.POINTER(UfoWeightsDesc), C.POINTER(C.c_void_p), C.c_void_p]
    lib.ufo_weights_destroy.argtypes = [C.c_void_p]
    lib.ufo_weights_destroy.restype = None
    lib.ufo_scene_create.argtypes = [C.POINTER(UfoSceneDesc), C.POINTER(C.c_void_p), C.c_void_p]
    lib.ufo_scene_destroy.argtypes = [C.c_void_p]
    lib.ufo_scene_destroy.restype = None
    lib.ufo_scene_device_bytes.argtypes = [C.c_void_p]
    lib.ufo_scene_device_bytes.restype = C.c_int64
    lib.ufo_render_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                    C.c_int64, C.c_int32, C.POINTER(UfoRenderOut), C.POINTER(UfoDebugTaps), C.c_void_p]
    lib.ufo_render_rays_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ufo_launch_count.restype = C.c_int64
    lib.ufo_costvolume_stage.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(UfoPixelwiseNet),
                                         C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ufo_costvolume_stage_rt.argtypes = lib.ufo_costvolume_stage.argtypes
    lib.ufo_debug_umma_selftest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                            C.c_void_p]
    lib.ufo_tsdf_integrate.argtypes = [C.POINTER(UfoTsdfGrid), C.c_void_p, C.c_void_p, C.POINTER(UfoTsdfView), C.c_int32, C.c_float,
                                       C.c_void_p]
    lib.ufo_feature_grid.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(UfoMlp3), C.c_void_p,
                                     C.c_void_p]
    lib.ufo_tsdf_mesh_begin.argtypes = [C.POINTER(UfoTsdfGrid), C.c_void_p, C.c_float, C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                                        C.POINTER(C.c_int64), C.c_void_p]
    lib.ufo_tsdf_mesh_emit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ufo_tsdf_mesh_destroy.argtypes = [C.c_void_p]
    lib.ufo_tsdf_mesh_destroy.restype = None
    lib.ufo_profile_end.argtypes = [C.POINTER(UfoProfileEntry), C.c_int32, C.POINTER(C.c_int32)]
    if lib.ufo_abi_version() != 2:
        raise UfoError(f"ABI version mismatch: library {lib.ufo_abi_version()} != binding 1")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().ufo_last_error().decode(errors="replace")
        raise UfoError(f"libuforecon_b200 error {rc}: {msg}")


def profile_begin() -> None:
    check(load().ufo_profile_begin())


def profile_end(cap: int = 256):
    """-> list of (kernel name, launches, summed device ms) since ``profile_begin``."""
    arr = (UfoProfileEntry * cap)()
    n = C.c_int32()
    check(load().ufo_profile_end(arr, cap, C.byref(n)))
    return [(arr[i].name.decode(), int(arr[i].launches), float(arr[i].ms)) for i in range(n.value)]
