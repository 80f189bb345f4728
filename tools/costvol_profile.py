"""Kernel 1 (cost-volume build) alone at the BASELINE size, for timing and for ncu.

    python tools/costvol_profile.py [--nv 3] [--wh 1600 1216] [--noisy] [--iters 3]

Cascade shapes of the reference (TransMVSNet.py:49-121): stage 1 (C=32, D=48, 1/4 res) with the SAME depth planes for every
pixel, stages 2 / 3 (C=16, D=32, 1/2 res; C=8, D=8, full res) with per-pixel hypotheses centred on the previous stage's depth
map.  By default that depth map is smooth (what a regularised depth map is); --noisy draws it per pixel like bench.py's
worst case, which scatters the source footprint of a reference tile over hundreds of pixels.
Prints one JSON line: per stage kernel ms (CUDA events inside the library), tap GB/s and compulsory GB/s.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from uforecon_b200 import _lib, checkpoint, synthetic  # noqa: E402
from uforecon_b200.costvolume import similarity_volume  # noqa: E402


def stage_inputs(nv, W, H, noisy, dev, seed=3):
    views = synthetic.UNFAVORABLE_VIEWS if nv == 3 else synthetic.TEN_VIEW_LIST[:nv]
    batch = synthetic.make_batch(views, (W, H))
    comb = [list(range(i, nv)) + list(range(0, i)) for i in range(nv)]
    g = torch.Generator(device=dev).manual_seed(seed)
    gc = torch.Generator().manual_seed(seed)
    out = []
    for si, (stage, D, C) in enumerate((("stage1", 48, 32), ("stage2", 32, 16), ("stage3", 8, 8))):
        sc = synthetic.STAGE_SCALE[stage]
        hs, ws = H // sc, W // sc
        feats = [torch.randn(nv, C, hs, ws, device=dev, generator=g) for _ in range(nv)]
        proj = batch["proj_matrices"][stage][0][torch.tensor(comb)].contiguous()
        if noisy:
            base = 425.0 + 2.65 * 192 * (0.3 + 0.4 * torch.rand(nv, 1, hs, ws, device=dev, generator=g))
        elif si == 0:
            base = torch.full((nv, 1, hs, ws), 425.0 + 2.65 * 96, device=dev)
        else:
            f = synthetic._smooth_field(gc, (nv, 1, hs, ws), coarse=16).to(dev)
            f = (f - f.amin()) / (f.amax() - f.amin() + 1e-8)
            base = 425.0 + 2.65 * 192 * (0.3 + 0.4 * f)
        step = 2.65 * (4 / (si + 1)) * (4.0 if si == 0 else 1.0)
        hyp = (base + (torch.arange(D, device=dev).view(1, D, 1, 1) - D / 2) * step).contiguous()
        out.append((stage, D, C, hs, ws, feats, proj, hyp))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nv", type=int, default=3)
    ap.add_argument("--wh", type=int, nargs=2, default=[1600, 1216])
    ap.add_argument("--noisy", action="store_true")
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda")
    sd = checkpoint.synthetic_state_dict(0)
    nv = a.nv
    res, vw = [], None
    for stage, D, C, hs, ws, feats, proj, hyp in stage_inputs(nv, a.wh[0], a.wh[1], a.noisy, dev):
        if vw is not None:
            vw = torch.nn.functional.interpolate(vw, scale_factor=2, mode="nearest").contiguous()
        similarity_volume(feats, proj, hyp, sd, view_weights=vw, device=dev)
        _lib.profile_begin()
        for _ in range(a.iters):
            sim, vw_new = similarity_volume(feats, proj, hyp, sd, view_weights=vw, device=dev)
        prof = _lib.profile_end(64)
        ms = sum(m for n, c, m in prof if n.startswith("k_costvol")) / a.iters
        ms_repack = sum(m for n, c, m in prof if n.startswith("k_nchw")) / a.iters
        vox = nv * D * hs * ws
        tap = vox * (nv - 1) * (4 * C * 4 + C * 4 / (nv - 1)) + vox * 4
        compulsory = nv * nv * C * hs * ws * 4 + vox * 4 * 2
        res.append({"stage": stage, "C": C, "D": D, "hw": [hs, ws], "voxels": vox, "kernel_ms": round(ms, 4), "repack_ms": round(ms_repack, 4),
                    "kernels": sorted({n for n, c, m in prof}), "tap_gbs": round(tap / (ms * 1e-3) / 1e9, 1),
                    "compulsory_bytes": compulsory, "compulsory_gbs": round(compulsory / (ms * 1e-3) / 1e9, 1),
                    "checksum": float(sim.double().sum())})
        vw = vw_new
        del feats, sim
    print(json.dumps({"n_views": nv, "wh": a.wh, "hypotheses": "noisy" if a.noisy else "smooth", "stages": res}))


if __name__ == "__main__":
    main()
