"""GPU: kernel 1 inside the reference's own cascade (TransMVSNet.forward with DepthNet's warp loop swapped for the fused
kernel) against the unmodified reference run on the same device, and the FPN dedup on CUDA."""
import copy
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from baseline import reference_arm  # noqa: E402
from uforecon_b200 import checkpoint, synthetic  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not reference_arm.available(), reason="baseline/_ref not staged")]


def _run(m, batch, fused=False, dedup=False):
    from uforecon_b200.costvolume import fused_cost_volume
    from uforecon_b200.encoder import dedup_feature_passes
    import contextlib
    imgs = batch["source_imgs"]
    pm = copy.deepcopy(batch["proj_matrices"])
    imgs_p, pm, dv = m.build_pairs(imgs, pm, batch["depth_values_org_scale"])
    with torch.no_grad(), (fused_cost_volume(m.transmvsnet) if fused else contextlib.nullcontext()), \
            (dedup_feature_passes(m.transmvsnet) if dedup else contextlib.nullcontext()):
        feats, out = m.transmvsnet(imgs_p, pm, dv)
    return out


def test_fused_cost_volume_inside_the_reference_cascade():
    sd = checkpoint.synthetic_state_dict(0)
    batch = reference_arm.to_device(synthetic.make_batch(synthetic.UNFAVORABLE_VIEWS, (160, 128)), "cuda")
    m = reference_arm.load_model(3, sd, "cuda")
    ref = _run(m, batch)
    got = _run(m, batch, fused=True)
    for st in ("stage1", "stage2", "stage3"):
        pv_r, pv_g = ref[st]["prob_volume"], got[st]["prob_volume"]
        pix_err = (pv_r - pv_g).abs().amax(dim=1)                                     # [N,h,w] worst probability error per pixel
        close = float((pix_err <= 1e-4).float().mean())
        same_depth = float((ref[st]["depth"] == got[st]["depth"]).float().mean())
        print(f"{st}: pixels with prob volume within 1e-4: {close:.5f}, max {float(pix_err.max()):.2e}; identical WTA depth {same_depth:.5f}")
        # The fused similarity volume is 1e-5 from the reference's CPU arithmetic (test_gpu_costvol.py); the comparator here is the
        # reference's CUDA path, whose own homographies (cuSOLVER inverse) and warps differ from its CPU path at that level.  Stage 1
        # sees identical inputs: every probability agrees.  A winner-take-all flip in one stage (a near tie) moves that pixel's
        # depth hypotheses in the next stage, so later stages are held to the fraction of pixels that agree.
        if st == "stage1":
            assert float(pix_err.max()) <= 1e-5 and same_depth >= 0.9999
        else:
            assert close >= 0.99 and same_depth >= 0.99
    # FPN dedup on CUDA: bit-identical on the CPU (tests/test_encoder_dedup.py); cuDNN's result for a sample can depend on its
    # position in the batch (tile mapping of the implicit GEMM), and the dedup rolls the batch, so CUDA is held to rounding level
    dd = _run(m, batch, dedup=True)
    ref2 = _run(m, batch)
    for st in ("stage1", "stage2", "stage3"):
        e = float((dd[st]["prob_volume"] - ref[st]["prob_volume"]).abs().max())
        e_self = float((ref2[st]["prob_volume"] - ref[st]["prob_volume"]).abs().max())
        same = float((dd[st]["depth"] == ref[st]["depth"]).float().mean())
        print(f"{st}: FPN dedup on CUDA: prob volume max abs err {e:.2e} (reference vs itself, second run: {e_self:.2e}), identical WTA depth {same:.5f}")
        if st == "stage1":
            assert e <= 1e-5 and same >= 0.9999
        else:
            assert same >= 0.99
