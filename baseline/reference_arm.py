"""The reference arm: the UNMODIFIED reference (``code1/`` of Youngju-Na/UFORecon) staged under ``baseline/_ref``.

``baseline/_ref`` is git-ignored (the reference is not product source) but travels to the GPU box with the
repository snapshot.  ``stage()`` copies the reference's own files from ``/root/reference`` byte for byte - it runs in
the build container only (``__graft_entry__.build()`` calls it when the reference is mounted).  Nothing under
``uforecon_b200/`` imports this module: it is used by ``bench.py --impl reference`` (the reference's ``UFORecon.infer``
on the host cores, code1/model.py:393-478), by bench.py's same-device extra (the same call with the model on
``cuda:0`` - the reference's ATen-op sequence on the B200) and by ``tools/make_golden.py``.

The reference imports five packages that are absent offline; ``install_stubs`` provides the same tiny
``sys.modules`` stand-ins the survey probe used (none of them is on the per-ray path) and the ``torch.from_numpy``
pass-through dtu_test_sparse.py:389 needs under torch 2.x.
"""
from __future__ import annotations

import argparse
import os
import shutil
import sys
import time
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
SOURCE_ROOT = os.environ.get("UFO_REFERENCE_ROOT", "/root/reference")


def stage(force: bool = False) -> str | None:
    """Copy ``code1/`` (+ main.py, scripts: the flag contract) from the mounted reference into baseline/_ref."""
    src = os.path.join(SOURCE_ROOT, "code1")
    if not os.path.isdir(src):
        return REF_DIR if available() else None
    dst = os.path.join(REF_DIR, "code1")
    if os.path.isdir(dst) and not force:
        return REF_DIR
    if os.path.isdir(REF_DIR):
        shutil.rmtree(REF_DIR)
    os.makedirs(REF_DIR)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for f in ("main.py", "tsdf_fusion.py", "License.txt"):
        p = os.path.join(SOURCE_ROOT, f)
        if os.path.exists(p):
            shutil.copy2(p, os.path.join(REF_DIR, f))
    if os.path.isdir(os.path.join(SOURCE_ROOT, "script")):
        shutil.copytree(os.path.join(SOURCE_ROOT, "script"), os.path.join(REF_DIR, "script"))
    with open(os.path.join(REF_DIR, "STAGED_FROM"), "w") as f:
        f.write(f"{SOURCE_ROOT} (byte-for-byte copy by baseline/reference_arm.py:stage)\n")
    return REF_DIR


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "code1", "model.py"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs(root: str) -> None:
    """sys.modules stand-ins for pytorch_lightning / piq / mcubes / kornia.utils / easydict, then ``root`` on sys.path."""
    class LightningModule(torch.nn.Module):
        def log(self, *a, **k):
            pass

    if "pytorch_lightning" not in sys.modules:
        _mod("pytorch_lightning", LightningModule=LightningModule)
    if "piq" not in sys.modules:
        _mod("piq", psnr=lambda a, b: torch.tensor(0.0))
    if "mcubes" not in sys.modules:
        _mod("mcubes")

    def create_meshgrid(h, w, normalized_coordinates=False, device=None):
        ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32, device=device),
                                torch.arange(w, dtype=torch.float32, device=device), indexing="ij")
        return torch.stack([xs, ys], -1)[None]

    if "kornia" not in sys.modules:
        k = _mod("kornia")
        k.utils = _mod("kornia.utils", create_meshgrid=create_meshgrid)

    class EasyDict(dict):
        def __init__(self, **kw):
            super().__init__(**kw)
            self.__dict__ = self

    if "easydict" not in sys.modules:
        _mod("easydict", EasyDict=EasyDict)
    if root not in sys.path:
        sys.path.insert(0, root)
    _fn = torch.from_numpy
    if not getattr(torch.from_numpy, "_ufo_patched", False):
        def from_numpy(x):
            return x if torch.is_tensor(x) else _fn(x)
        from_numpy._ufo_patched = True
        torch.from_numpy = from_numpy


def canonical_args(n_view=3, **over):
    """Flag set of script/eval_dtu_unfavorable.sh:7-11 plus main.py defaults (main.py:37-104)."""
    a = dict(patch_size=48, sW=1, sH=1, train_ray_num=1024, extract_geometry=True,
             test_sample_coarse=64, test_sample_fine=64, coarse_sample=64, fine_sample=64,
             ndepths="48,32,8", depth_inter_r="4,2,1", share_cr=False, cr_base_chs="8,8,8", grad_method="detach",
             volume_type="correlation", volume_reso=96, mvs_depth_guide=1, depth_pos_encoding=True,
             explicit_similarity=True, use_dir_srdf=False, only_reference_frustum=False, test_coarse_only=False,
             test_ray_num=800, test_n_view=n_view, train_n_view=5, uforecon_lr=1e-4,
             out_dir="/tmp/ufo_ref_out", logdir="/tmp/ufo_ref_log")
    a.update(over)
    return argparse.Namespace(**a)


def load_model(nv: int, state_dict: dict, device="cpu"):
    """``UFORecon(args)`` of the staged reference in eval mode with the hot-path tensors of ``state_dict`` loaded
    (the encoder keeps its seeded initialisation: it is not on the timed path)."""
    if not available():
        raise RuntimeError("baseline/_ref is not staged (run __graft_entry__.build() in the build container)")
    install_stubs(REF_DIR)
    import warnings
    warnings.filterwarnings("ignore")
    from code1.model import UFORecon          # the reference, unmodified
    torch.manual_seed(0)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):     # the reference's constructors print banners; stdout carries bench.py's JSON line
        m = UFORecon(canonical_args(n_view=nv)).eval()
    missing, unexpected = m.load_state_dict(state_dict, strict=False)
    assert not unexpected, unexpected
    hot = [k for k in missing if k.startswith("ray_transformer") or k.startswith("deviation")]
    assert not hot, f"hot-path keys missing from the state dict: {hot}"
    return m.to(device)


def to_device(obj, device):
    if torch.is_tensor(obj):
        return obj.to(device)
    if isinstance(obj, dict):
        return {k: to_device(v, device) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(to_device(v, device) for v in obj)
    return obj


def infer_chunks(model, batch, scene, n_chunks: int, chunk: int = 800, warm: int = 1, device="cpu", seed0: int = 100):
    """Times ``n_chunks`` calls of the reference's own ``UFORecon.infer(extract_geometry=True)`` on ``chunk`` rays each
    (its own chunking: --test_ray_num 800).  ``batch`` / ``scene`` must already live on ``device``.  Returns
    (rays/s, seconds, rays, last (depth, rgb))."""
    H, W = batch["source_imgs"].shape[-2:]
    total = H * W
    cuda = torch.device(device).type == "cuda"
    times, last = [], None
    with torch.no_grad():
        for i in range(warm + n_chunks):
            begin = (total // (warm + n_chunks + 1)) * (i + 1)
            ray_idx = torch.arange(begin, min(begin + chunk, total), device=device)
            torch.manual_seed(seed0 + i)       # the reference draws its sampler uniforms from the CPU generator
            if cuda:
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            srdf, pts, depth, rgb = model.infer(batch=batch, ray_idx=ray_idx[None], source_imgs_feat=scene["source_imgs_feat"],
                                                feature_volume=scene["feature_volume"], match_feature=scene["match_feature"],
                                                extract_geometry=True, is_train=False)
            if cuda:
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            last = (depth, rgb)
            if i >= warm:
                times.append((dt, int(ray_idx.numel())))
    secs = sum(t for t, _ in times)
    rays = sum(n for _, n in times)
    return rays / secs, secs, rays, last


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
