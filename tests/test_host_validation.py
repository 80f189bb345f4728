"""CPU: argument validation of the host-side mirror happens before any CUDA call (same error behaviour as the reference's
shape asserts), and the sharding helpers cover the BASELINE configs."""
import pytest
import torch

from conftest import make_case
from uforecon_b200 import dist as ufodist
from uforecon_b200 import synthetic
from uforecon_b200.renderer import TAP_SHAPES, Scene


@pytest.fixture(scope="module")
def small():
    return make_case(synthetic.UNFAVORABLE_VIEWS, (96, 64))


def test_scene_rejects_bad_inputs_before_touching_the_gpu(small):
    batch, scene, sd = small
    with pytest.raises(ValueError):                                     # batch size must be 1 (main.py:159-162)
        Scene(batch, scene["source_imgs_feat"].repeat(2, 1, 1, 1, 1), scene["feature_volume"], scene["match_feature"], device="cpu")
    with pytest.raises(ValueError):                                     # match maps must be [1, NV, (NV-1)*32, h, w]
        Scene(batch, scene["source_imgs_feat"], scene["feature_volume"], [scene["match_feature"][0][:, :, :32]], device="cpu")
    with pytest.raises(ValueError):                                     # 32-channel FPN features
        Scene(batch, scene["source_imgs_feat"][:, :, :16], scene["feature_volume"], scene["match_feature"], device="cpu")
    b2 = {k: v for k, v in batch.items() if k != "depth_info"}
    with pytest.raises(KeyError):                                       # depth_info is set by extract_geometry (model.py:806-808)
        Scene(b2, scene["source_imgs_feat"], scene["feature_volume"], scene["match_feature"], device="cpu")


def test_tap_shapes_follow_the_header():
    assert TAP_SHAPES["tokens"](7, 5) == (7, 128, 5, 80)
    assert TAP_SHAPES["vol24"](7, 5) == (7, 128, 24) and TAP_SHAPES["z_fine"](7, 5) == (7, 64)


def test_sharding_covers_baseline_configs():
    # config 3: 1600x1216 rows over 2/4/8 GPUs; config 5: 49 depth maps over 8 GPUs
    for world in (1, 2, 4, 8):
        spans = [ufodist.shard_rows(1216, 1600, world, r) for r in range(world)]
        assert sum(n for _, n in spans) == 1216 * 1600 and all(n % 1600 == 0 for _, n in spans)
    per_rank = [len(ufodist.shard_images(49, 8, r)) for r in range(8)]
    assert sum(per_rank) == 49 and max(per_rank) == 7
    with pytest.raises(ValueError):
        ufodist.shard_rows(10, 10, 4, 4)
