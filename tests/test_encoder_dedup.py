"""N1 (PyTorch side): the FPN passes TransMVSNet.forward repeats are removed with bit-identical encoder outputs.
Runs the UNMODIFIED reference staged under baseline/_ref on CPU."""
import copy
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from baseline import reference_arm  # noqa: E402
from uforecon_b200 import checkpoint, synthetic  # noqa: E402
from uforecon_b200.encoder import dedup_feature_passes  # noqa: E402

pytestmark = pytest.mark.skipif(not reference_arm.available(), reason="baseline/_ref not staged")


def _encode(m, batch, dedup):
    imgs = batch["source_imgs"]
    pm = copy.deepcopy(batch["proj_matrices"])                    # build_pairs rewrites the dict in place (model.py:154-155)
    imgs_p, pm, dv = m.build_pairs(imgs, pm, batch["depth_values_org_scale"])
    with torch.no_grad():
        if dedup:
            with dedup_feature_passes(m.transmvsnet) as st:
                feats, out = m.transmvsnet(imgs_p, pm, dv)
            assert st.fpn_passes == 1 and st.calls == imgs.shape[1]
        else:
            feats, out = m.transmvsnet(imgs_p, pm, dv)
    return feats, out


@pytest.mark.parametrize("nv", [3, 4])
def test_dedup_encoder_is_bit_identical(nv):
    sd = checkpoint.synthetic_state_dict(0)
    views = synthetic.UNFAVORABLE_VIEWS if nv == 3 else synthetic.TEN_VIEW_LIST[:nv]
    batch = synthetic.make_batch(views, (160, 128))
    m = reference_arm.load_model(nv, sd)
    fa, oa = _encode(m, batch, False)
    fb, ob = _encode(m, batch, True)
    for j in range(nv):
        for k in fa[j]:
            assert torch.equal(fa[j][k], fb[j][k]), (j, k)
    for st in ("stage1", "stage2", "stage3"):
        for k in ("depth", "prob_volume", "depth_values"):
            assert torch.equal(oa[st][k], ob[st][k]), (st, k)
    assert m.transmvsnet.feature.__class__.__name__ == "FeatureNet"     # restored


@pytest.mark.parametrize("nv", [3, 4])
def test_compact_match_features_are_the_reference_maps_stored_once(nv):
    """N2 producer side: the compact pair maps equal every slot of the reference's redundant get_match_feat layout."""
    from uforecon_b200.encoder import compact_match_features
    from uforecon_b200.synthetic import expand_pair_maps
    sd = checkpoint.synthetic_state_dict(0)
    views = synthetic.UNFAVORABLE_VIEWS if nv == 3 else synthetic.TEN_VIEW_LIST[:nv]
    batch = synthetic.make_batch(views, (160, 128))
    m = reference_arm.load_model(nv, sd)
    feats, _ = _encode(m, batch, True)
    for i in range(len(feats)):
        feats[i]["stage1"] = feats[i]["stage1"][0:1]                                  # model.py:782-783
    with torch.no_grad():
        ref = m.transmvsnet.get_match_feat(feats, cur_n_src_views=nv)[0]              # [1, NV, (NV-1)*32, h, w]  (model.py:785)
        pairs = compact_match_features(m.transmvsnet, feats)
    assert pairs is not None and tuple(pairs.shape) == (nv * (nv - 1) // 2, 32) + tuple(ref.shape[-2:])
    assert torch.equal(expand_pair_maps(pairs, nv), ref)
