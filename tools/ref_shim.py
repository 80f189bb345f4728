"""Import shim for running the UNMODIFIED reference from /root/reference in the build container.

Container-only tooling (the reference does not exist on the GPU box): used by ``make_golden.py`` to
produce the committed fixtures under ``tests/golden/``.  Installs tiny ``sys.modules`` stubs for the
reference's imports that are missing offline and the ``torch.from_numpy`` tensor
pass-through needed by dtu_test_sparse.py:389 under torch 2.x.
"""
import argparse
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("UFO_REFERENCE_ROOT", "/root/reference")


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT} (this tool only runs in the build container)")

    class LightningModule(torch.nn.Module):
        def log(self, *a, **k):
            pass

    _mod("pytorch_lightning", LightningModule=LightningModule)
    _mod("piq", psnr=lambda a, b: torch.tensor(0.0))
    _mod("mcubes")

    def create_meshgrid(h, w, normalized_coordinates=False, device=None):
        ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32, device=device),
                                torch.arange(w, dtype=torch.float32, device=device), indexing="ij")
        return torch.stack([xs, ys], -1)[None]

    k = _mod("kornia")
    k.utils = _mod("kornia.utils", create_meshgrid=create_meshgrid)

    class EasyDict(dict):
        def __init__(self, **kw):
            super().__init__(**kw)
            self.__dict__ = self

    _mod("easydict", EasyDict=EasyDict)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _fn = torch.from_numpy
    if not getattr(torch.from_numpy, "_ufo_patched", False):
        def from_numpy(x):
            return x if torch.is_tensor(x) else _fn(x)
        from_numpy._ufo_patched = True
        torch.from_numpy = from_numpy


def canonical_args(n_view=3, **over):
    """Flag set of script/eval_dtu_unfavorable.sh:7-11 plus main.py defaults."""
    a = dict(patch_size=48, sW=1, sH=1, train_ray_num=1024, extract_geometry=True,
             test_sample_coarse=64, test_sample_fine=64, coarse_sample=64, fine_sample=64,
             ndepths="48,32,8", depth_inter_r="4,2,1", share_cr=False, cr_base_chs="8,8,8", grad_method="detach",
             volume_type="correlation", volume_reso=96, mvs_depth_guide=1, depth_pos_encoding=True,
             explicit_similarity=True, use_dir_srdf=False, only_reference_frustum=False, test_coarse_only=False,
             test_ray_num=800, test_n_view=n_view, train_n_view=5, uforecon_lr=1e-4,
             out_dir="/tmp/ufo_ref_out", logdir="/tmp/ufo_ref_log")
    a.update(over)
    return argparse.Namespace(**a)
