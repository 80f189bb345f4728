"""The staged, unmodified reference (baseline/_ref) against the oracle restatement on the same seeded inputs.

Runs wherever baseline/_ref exists (the build container stages it from /root/reference; it travels to the GPU box
with the snapshot).  This is the pin of oracle/uforecon_oracle.py to the reference's own ``UFORecon.infer``
(code1/model.py:393-478) beyond the committed golden files."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from baseline import reference_arm  # noqa: E402
from oracle import uforecon_oracle as orc  # noqa: E402
from uforecon_b200 import checkpoint, synthetic  # noqa: E402

pytestmark = pytest.mark.skipif(not reference_arm.available(), reason="baseline/_ref not staged")


@pytest.mark.parametrize("views", [synthetic.UNFAVORABLE_VIEWS, synthetic.TEN_VIEW_LIST[:5]])
def test_staged_reference_infer_matches_oracle(views):
    sd = checkpoint.synthetic_state_dict(0)
    batch = synthetic.make_batch(views, (96, 64))
    scene = synthetic.make_scene(batch)
    batch["depth_info"] = scene["depth_info"]
    m = reference_arm.load_model(len(views), sd)
    ray_idx = torch.arange(7, 96 * 64, 96 * 64 // 24)[:24]
    torch.manual_seed(5)
    with torch.no_grad():
        srdf, pts, depth, rgb = m.infer(batch=batch, ray_idx=ray_idx[None], source_imgs_feat=scene["source_imgs_feat"],
                                        feature_volume=scene["feature_volume"], match_feature=scene["match_feature"],
                                        extract_geometry=True, is_train=False)
    u_c, u_f = synthetic.sampler_uniforms(len(ray_idx), seed=5)
    with torch.no_grad():
        o = orc.infer(batch, scene, sd, ray_idx, u_c, u_f)
    assert float((o["depth"] - depth[0]).abs().max()) < 2e-6
    assert float((o["rgb"] - rgb[0]).abs().max()) < 2e-6
    assert float((o["srdf"] - srdf[0]).abs().max()) < 2e-5 * max(1.0, float(srdf.abs().max()))


def test_infer_chunks_timer_runs():
    sd = checkpoint.synthetic_state_dict(0)
    batch = synthetic.make_batch(synthetic.UNFAVORABLE_VIEWS, (96, 64))
    scene = synthetic.make_scene(batch)
    batch["depth_info"] = scene["depth_info"]
    m = reference_arm.load_model(3, sd)
    v, secs, rays, last = reference_arm.infer_chunks(m, batch, scene, 1, chunk=32, warm=0)
    assert rays == 32 and v > 0 and last[0].shape == (1, 32)
