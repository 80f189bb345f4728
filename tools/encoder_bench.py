"""Seconds of the reference's encoder (unmodified, staged under baseline/_ref) on cuda:0 next to the render, with and without
the FPN dedup (uforecon_b200/encoder.py) and with kernel 1 swapped into DepthNet (uforecon_b200/costvolume.py):

    python tools/encoder_bench.py [W H] > profiles/r02_encoder_seconds.json

What is timed is what extract_geometry runs before its chunk loop (model.py:780-802): build_pairs, TransMVSNet.forward (FPN x N^2,
FMT, three cascade stages with 3-D U-Nets), get_match_feat, MVSVolume x 3.  Random images, synthetic checkpoint."""
import contextlib
import copy
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import reference_arm  # noqa: E402
from uforecon_b200 import checkpoint, synthetic  # noqa: E402
from uforecon_b200.costvolume import fused_cost_volume  # noqa: E402
from uforecon_b200.encoder import dedup_feature_passes  # noqa: E402


def encode(m, batch, dedup, fused):
    imgs = batch["source_imgs"]
    pm = copy.deepcopy(batch["proj_matrices"])
    imgs_p, pm, dv = m.build_pairs(imgs, pm, batch["depth_values_org_scale"])
    N = imgs.shape[1]
    with torch.no_grad(), (dedup_feature_passes(m.transmvsnet) if dedup else contextlib.nullcontext()), \
            (fused_cost_volume(m.transmvsnet) if fused else contextlib.nullcontext()):
        feats, out = m.transmvsnet(imgs_p, pm, dv)                                      # model.py:781
        for i in range(len(feats)):
            feats[i]["stage1"] = feats[i]["stage1"][0:1]                                # :782-783
        m.transmvsnet.get_match_feat(feats, cur_n_src_views=N)                         # :785
        for st in ("stage1", "stage2", "stage3"):
            m.build_mvs_volume(batch, out[st]["cost_volume"])                          # :796-798
    return out


def _grid_sample_without_cudnn():
    """torch 2.11's cuDNN spatial-transformer path rejects the 1600x1216 warps of homo_warping_trans (fmt/module.py:363:
    CUDNN_STATUS_NOT_SUPPORTED); route F.grid_sample to ATen's native kernel - convolutions keep cuDNN."""
    import torch.nn.functional as F
    if getattr(F.grid_sample, "_ufo_patched", False):
        return
    orig = F.grid_sample

    def grid_sample(*a, **k):
        with torch.backends.cudnn.flags(enabled=False):
            return orig(*a, **k)
    grid_sample._ufo_patched = True
    F.grid_sample = grid_sample


def main():
    _grid_sample_without_cudnn()
    W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1600, 1216)
    sd = checkpoint.synthetic_state_dict(0)
    res = {"what": __doc__.split("\n\n")[0], "wh": [W, H], "rows": []}
    for nv in (3, 5, 10):
        views = synthetic.UNFAVORABLE_VIEWS if nv == 3 else synthetic.TEN_VIEW_LIST[:nv]
        batch = reference_arm.to_device(synthetic.make_batch(views, (W, H)), "cuda")
        m = reference_arm.load_model(nv, sd, "cuda")
        row = {"n_views": nv}
        for name, (dd, ff) in (("reference", (False, False)), ("fpn_dedup", (True, False)), ("fpn_dedup+kernel1", (True, True))):
            try:
                for it in range(2):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    encode(m, batch, dd, ff)
                    torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
                row[name + "_s"] = dt
                row[name + "_peak_gb"] = torch.cuda.max_memory_allocated() / 1e9
            except torch.OutOfMemoryError:
                row[name + "_s"] = "out of memory"
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats()
        res["rows"].append(row)
        del m, batch
        torch.cuda.empty_cache()
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
