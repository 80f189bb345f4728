"""PyTorch-side helpers around the reference's encoder (SURVEY.md "next" row N1): the encoder itself stays the reference's
PyTorch code; what lives here removes work it repeats, with bit-identical results.

``dedup_feature_passes``  ``TransMVSNet.forward`` runs its FPN (``FeatureNet``, 9 deformable convs) once per view INDEX on a
batch of all N cyclic rotations of the view list (code1/encoder_utils/fmt/TransMVSNet.py:174-177 after
``UFORecon.build_pairs``, code1/model.py:139-160): ``imgs[r, j] = image[(r + j) % N]``, so the j-th call sees the same N
images as the first one, rotated by j along the batch axis - N^2 FPN passes for N distinct images.  Inside the context
manager the first call runs the FPN; the others return ``torch.roll`` of its outputs (new tensors, because
``FMT_with_pathway`` updates the per-view dicts in place, FMT.py:282-315).  Survey probe: the encoder is 4.7 / 16.1 / 64.6 s on
8 CPU cores for N = 3 / 5 / 10 - the wall-clock floor of a depth map once the render is sub-second.

``compact_match_features``  (N2, producer side) ``TransMVSNet.get_match_feat`` (TransMVSNet.py:341-374) stacks the cross-view
maps of ``FMT_with_pathway.extract_cross_features`` into ``[B, NV, (NV-1)*32, h, w]``: every unordered pair's map appears
twice, because ``FMT.forward(feat="cross")`` returns the same tensor for both images of a pair (FMT.py:197).
This function returns the ``[NV(NV-1)/2, 32, h, w]`` tensor the FMT produced, without the redundant stack, for
``Scene(..., pair_maps=...)`` / ``UfoSceneDesc.match_pairs`` (1.4 GB less at NV = 10, 1600x1216).
"""
from __future__ import annotations

import contextlib
from typing import Dict, Optional

import torch


class _RolledFeature(torch.nn.Module):
    def __init__(self, feature: torch.nn.Module, check: bool = True):
        super().__init__()
        self.feature = feature
        self.check = check
        self.calls = 0
        self.fpn_passes = 0
        self._img0: Optional[torch.Tensor] = None
        self._out0: Optional[Dict[str, torch.Tensor]] = None

    def forward(self, img: torch.Tensor):
        j = self.calls
        self.calls += 1
        n = img.shape[0]
        if j == 0 or self._out0 is None or j >= n:
            self._img0, self._out0 = img, self.feature(img)
            self.fpn_passes += 1
            return dict(self._out0)
        if self.check and not torch.equal(img, torch.roll(self._img0, -j, 0)):
            # not the cyclic-rotation batch build_pairs makes: fall back to the real pass
            self.fpn_passes += 1
            return self.feature(img)
        return {k: torch.roll(v, -j, 0) for k, v in self._out0.items()}


@contextlib.contextmanager
def dedup_feature_passes(transmvsnet: torch.nn.Module, check: bool = True):
    """Within the block ``transmvsnet.feature`` is evaluated once per ``forward`` instead of once per view index.

        with dedup_feature_passes(model.transmvsnet) as stats:
            feats, outputs = model.transmvsnet(imgs, proj_matrices, depth_values)     # model.py:781
        stats.fpn_passes == 1 (the reference runs imgs.size(1) passes over the same images)
    """
    orig = transmvsnet.feature
    wrapped = _RolledFeature(orig, check)
    transmvsnet.feature = wrapped
    try:
        yield wrapped
    finally:
        transmvsnet.feature = orig


def compact_match_features(transmvsnet: torch.nn.Module, features, check: bool = True) -> Optional[torch.Tensor]:
    """The cross-view match maps with every view pair stored once: ``[NV(NV-1)/2, 32, h, w]``, pairs in the reference's
    enumeration order ``(a, b) for a in range(NV-1) for b in range(a+1, NV)`` (model.py:273-276).

    ``features`` is what ``TransMVSNet.forward`` returns first (one dict of FPN stages per view, stage 1 already cut to the first
    rotation as ``extract_geometry`` does, model.py:782-783).  Returns ``None`` when the two tensors of the FMT differ (an encoder
    whose cross attention is not symmetric): the caller then keeps the reference layout of ``get_match_feat``."""
    out = transmvsnet.FMT_with_pathway.extract_cross_features(features)          # FMT.py:282-315
    f0, f1 = out["aug_feat0s"][0], out["aug_feat1s"][0]                          # [B, nC2, 32, h, w] each
    if check and not (f0 is f1 or torch.equal(f0, f1)):
        return None
    if f0.shape[0] != 1:
        raise ValueError(f"expected batch 1 (stage-1 features of the first rotation), got {f0.shape[0]}")
    nv = len(features)
    n_pairs = nv * (nv - 1) // 2
    # the FMT's cross output carries 2 nC2 maps (both attention directions, FMT.py:308-309); get_match_feat indexes the first nC2
    # of them for BOTH views of a pair (TransMVSNet.py:362-366) - those are the maps the hot path samples
    return f0[0, :n_pairs].contiguous()
