"""Shared generator of cost-volume test inputs (same draws as tools/make_golden.py::costvol_case)."""
import numpy as np
import torch

from uforecon_b200 import checkpoint, synthetic


def build_pairs(proj_matrices, n):
    """All cyclic rotations of the view list (reference: UFORecon.build_pairs, code1/model.py:139-160)."""
    comb = np.array([list(range(i, n)) + list(range(0, i)) for i in range(n)])
    out = {}
    for st in ("stage1", "stage2", "stage3"):
        out[st] = proj_matrices[st][0][comb]          # [N, V, 2, 4, 4]
    return out


def costvol_inputs(views, wh):
    nv = len(views)
    W, H = wh
    sd = checkpoint.synthetic_state_dict(0)
    batch = synthetic.make_batch(views, wh, seed=0)
    pm = build_pairs(batch["proj_matrices"], nv)
    gen = torch.Generator().manual_seed(7)
    stages = []
    for si, (stage, D, C) in enumerate((("stage1", 48, 32), ("stage2", 32, 16), ("stage3", 8, 8))):
        s = synthetic.STAGE_SCALE[stage]
        hs, ws = H // s, W // s
        feats = [synthetic._smooth_field(gen, (nv, C, hs, ws), coarse=4) for _ in range(nv)]
        base = 425.0 + 2.65 * 192 * (0.3 + 0.4 * torch.rand(nv, 1, hs, ws, generator=gen))
        hyp = base + (torch.arange(D).view(1, D, 1, 1) - D / 2) * 2.65 * (4 / (si + 1)) * (4.0 if si == 0 else 1.0)
        stages.append((stage, feats, pm[stage].contiguous(), hyp.contiguous()))
    return batch, sd, stages


def test_build_pairs_shape():
    batch = synthetic.make_batch([1, 16, 36], (96, 64))
    pm = build_pairs(batch["proj_matrices"], 3)
    assert tuple(pm["stage1"].shape) == (3, 3, 2, 4, 4)
    assert torch.equal(pm["stage1"][1, 0], batch["proj_matrices"]["stage1"][0, 1])
