"""Generate tests/golden/feature_grid.npz from the UNMODIFIED reference's ``FeatureVolume`` (code1/feature_volume.py),
the alternative ``--volume_type featuregrid`` (row a19).  Build container only.  The 3-D regulariser behind it is
replaced by a capture hook, so the stored tensor is what ``volume_regularization`` receives (:92-95).  The weights
of ``FeatureVolume.linear`` are not in the shipped checkpoint, so they are seeded here and stored."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_shim  # noqa: E402

ref_shim.install()
from code1.feature_volume import FeatureVolume  # noqa: E402  (the reference)
from oracle import uforecon_oracle as orc  # noqa: E402
from uforecon_b200 import synthetic  # noqa: E402

RESO = 20


def fgrid_inputs(reso=RESO):
    batch = synthetic.make_batch(synthetic.UNFAVORABLE_VIEWS, (96, 64))
    gen = torch.Generator().manual_seed(11)
    feats = synthetic._smooth_field(gen, (3, 32, 16, 24), coarse=4)
    lin = {}
    for name, shp in (("0.weight", (32, 32)), ("0.bias", (32,)), ("2.weight", (16, 32)), ("2.bias", (16,)),
                      ("4.weight", (8, 16)), ("4.bias", (8,))):
        lin[name] = (torch.rand(shp, generator=gen) * 2 - 1) * (0.3 if name.endswith("weight") else 0.1)
    return batch, feats, lin


def main():
    batch, feats, lin = fgrid_inputs()
    torch.manual_seed(0)
    m = FeatureVolume(RESO).eval()
    m.linear.load_state_dict(lin)
    cap = {}

    class Capture(torch.nn.Module):
        def forward(self, v):
            cap["v"] = v
            return v

    m.volume_regularization = Capture()
    with torch.no_grad():
        m(feats[None], batch)
        ref = cap["v"][0]
        o = orc.feature_grid_meanvar(feats, batch["source_poses"][0], lin, RESO)
    print("reference vs oracle: max |d|", float((ref - o).abs().max()), "scale", float(ref.abs().max()), tuple(ref.shape))
    out = os.path.join(ROOT, "tests", "golden", "feature_grid.npz")
    np.savez_compressed(out, meanvar=ref.numpy().astype(np.float32))
    print("wrote", out, os.path.getsize(out))


if __name__ == "__main__":
    main()
