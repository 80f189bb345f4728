"""GPU: the host-side mirror of the reference's method interface (uforecon_b200/renderer.py).

``UFOReconRenderer.infer`` must be callable exactly like ``UFORecon.infer(extract_geometry=True)``
(code1/model.py:393-478): same arguments, same return tuple, sampler uniforms drawn from torch's global CPU
generator in the reference's order, so that ``torch.manual_seed`` reproduces the reference's sample positions.
``render_depth_map`` is the chunk loop of ``extract_geometry`` (code1/model.py:814-826).
"""
import pytest
import torch

from conftest import make_case, rel_err
from oracle import uforecon_oracle as orc
from uforecon_b200 import synthetic
from uforecon_b200._lib import UFO_MODE_FP32, UFO_MODE_TC_F16 as UFO_MODE_TC
from uforecon_b200.renderer import UFOReconRenderer, draw_uniforms

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small():
    batch, scene, sd = make_case(synthetic.UNFAVORABLE_VIEWS, (96, 64))
    return batch, scene, sd


def test_infer_signature_and_seeded_parity(small):
    batch, scene, sd = small
    ren = UFOReconRenderer(sd, mode=UFO_MODE_FP32)
    ray_idx = torch.arange(1000, 1000 + 200)[None]                       # [1, RN] like torch.split(ray_idx_all, ...)
    torch.manual_seed(11)
    srdf, pts, depth, rgb = ren.infer(batch, ray_idx, scene["source_imgs_feat"], feature_volume=scene["feature_volume"],
                                      extract_geometry=True, match_feature=scene["match_feature"])
    assert tuple(srdf.shape) == (1, 200, 128) and tuple(pts.shape) == (1, 200, 128, 3)
    assert tuple(depth.shape) == (1, 200) and tuple(rgb.shape) == (1, 200, 3)
    torch.manual_seed(11)                                                 # the reference's draws: coarse, then fine
    u_c, u_f = draw_uniforms(200)
    with torch.no_grad():
        o = orc.infer(batch, scene, sd, ray_idx[0], u_c, u_f)
    assert rel_err(depth[0].cpu(), o["depth"]) <= 1e-4
    assert rel_err(srdf[0].cpu(), o["srdf"]) <= 5e-4      # end to end through the importance sampler
    assert rel_err(pts[0].cpu(), o["points"]) <= 5e-4       # fine samples move with the coarse weights (inverse CDF)
    with pytest.raises(NotImplementedError):
        ren.infer(batch, ray_idx, scene["source_imgs_feat"], feature_volume=scene["feature_volume"],
                  extract_geometry=False, match_feature=scene["match_feature"])
    ren.close()


@pytest.mark.parametrize("mode", [UFO_MODE_FP32, UFO_MODE_TC])
def test_render_depth_map_matches_chunked_infer(small, mode):
    """One library call over the whole ray grid == the reference's loop over 800-ray chunks (same uniform draws)."""
    batch, scene, sd = small
    ren = UFOReconRenderer(sd, mode=mode, test_ray_num=800)
    H, W = 64, 96
    torch.manual_seed(5)
    depth_mm, rgb = ren.render_depth_map(batch, scene["source_imgs_feat"], scene["feature_volume"], scene["match_feature"])
    assert tuple(depth_mm.shape) == (H, W) and tuple(rgb.shape) == (H, W, 3)
    torch.manual_seed(5)
    cz = batch["cam_ray_d"][0][2].to(depth_mm.device)
    scale = float(batch["scale_mat"][0][0, 0])
    got_d, got_c = [], []
    for ray_idx in torch.split(torch.arange(H * W), 800):                # extract_geometry's loop, model.py:814-823
        _, _, d, c = ren.infer(batch, ray_idx[None], scene["source_imgs_feat"], feature_volume=scene["feature_volume"],
                               extract_geometry=True, match_feature=scene["match_feature"])
        got_d.append(d[0] * cz[ray_idx.to(cz.device)] * scale)
        got_c.append(c[0])
    ref_d, ref_c = torch.cat(got_d).view(H, W), torch.cat(got_c).view(H, W, 3)
    assert torch.equal(depth_mm, ref_d) and torch.equal(rgb, ref_c)      # tiling never changes a bit
    ren.close()


@pytest.mark.parametrize("mode", [UFO_MODE_FP32, UFO_MODE_TC])
def test_host_buffer_entry_point_equals_device_call(small, mode, monkeypatch):
    """ufo_render_rays_host (pinned host uniforms in, host depth/rgb out, uploads pipelined in column blocks) must
    return exactly what ufo_render_rays returns for the same ray range."""
    import ctypes as C
    from uforecon_b200 import _lib
    from uforecon_b200.renderer import HotPathWeights, Scene, render_rays
    batch, scene, sd = small
    monkeypatch.setenv("UFO_TC_CHUNK", "96")        # several upload blocks (4 chunks each) for this small ray count
    monkeypatch.setenv("UFO_FP32_CHUNK", "96")
    lib = _lib.load()
    w = HotPathWeights(sd)
    sc = Scene(batch, scene["source_imgs_feat"], scene["feature_volume"], scene["match_feature"])
    n, begin = 1000, 321
    u_c, u_f = synthetic.sampler_uniforms(n, seed=9)
    ref = render_rays(sc, w, None, n, u_c, u_f, mode, ray_begin=begin, want=("depth_z", "rgb"))
    u_ch, u_fh = u_c.contiguous().pin_memory(), u_f.contiguous().pin_memory()
    dz, rgb = torch.empty(n).pin_memory(), torch.empty(n, 3).pin_memory()
    _lib.check(lib.ufo_render_rays_host(sc.handle, w.handle, begin, n, u_ch.data_ptr(), u_fh.data_ptr(), mode, dz.data_ptr(),
                                        rgb.data_ptr(), torch.cuda.current_stream().cuda_stream))
    assert torch.equal(dz, ref["depth_z"].cpu()) and torch.equal(rgb, ref["rgb"].cpu())
    sc.close()
    w.close()


def test_render_depth_map_device_rng(small):
    """device-side uniforms: same distribution of jitter, so the depth map agrees with the CPU-seeded one up to the
    sampling noise of the quadrature (a loose statistical bound), and two seeds differ."""
    batch, scene, sd = small
    ren = UFOReconRenderer(sd, mode=UFO_MODE_TC)
    torch.manual_seed(1)
    d_ref, _ = ren.render_depth_map(batch, scene["source_imgs_feat"], scene["feature_volume"], scene["match_feature"])
    torch.cuda.manual_seed(1)
    d_a, c_a = ren.render_depth_map(batch, scene["source_imgs_feat"], scene["feature_volume"], scene["match_feature"], device_rng=True)
    torch.cuda.manual_seed(2)
    d_b, _ = ren.render_depth_map(batch, scene["source_imgs_feat"], scene["feature_volume"], scene["match_feature"], device_rng=True)
    assert torch.isfinite(d_a).all() and torch.isfinite(c_a).all()
    assert not torch.equal(d_a, d_b)
    rel = float(((d_a - d_ref).abs() / d_ref.abs().clamp_min(1e-6)).median())
    assert rel < 2e-2, rel
    ren.close()


def test_one_renderer_two_render_views_and_in_place_updates(small):
    """The cached Scene must follow its inputs: another render view of the same source set (only rays / poses / MVS depth
    change - the encoder tensors are the SAME objects), an in-place update of an encoder output, and a freed-and-reallocated
    batch must each give the result of a fresh renderer."""
    batch, scene, sd = small
    ren = UFOReconRenderer(sd, mode=UFO_MODE_FP32)
    ray_idx = torch.arange(100, 400)[None]

    def run(r, b, sc_):
        torch.manual_seed(3)
        return r.infer(b, ray_idx, sc_["source_imgs_feat"], feature_volume=sc_["feature_volume"], extract_geometry=True,
                       match_feature=sc_["match_feature"])

    first = run(ren, batch, scene)
    # render view 2: same encoder tensor OBJECTS, other rays / reference pose (what DtuFitSparse yields for the next batch)
    b2 = dict(batch)
    b2["ray_d"] = torch.roll(batch["ray_d"], 37, dims=2).clone()
    b2["cam_ray_d"] = torch.roll(batch["cam_ray_d"], 37, dims=2).clone()
    b2["ref_pose_inv"] = batch["ref_pose_inv"].clone()
    b2["ref_pose_inv"][0, :3, 3] += 0.05
    second = run(ren, b2, scene)
    fresh = UFOReconRenderer(sd, mode=UFO_MODE_FP32)
    want = run(fresh, b2, scene)
    fresh.close()
    assert not torch.equal(first[2], second[2])
    for a, b in zip(second, want):
        assert torch.equal(a, b)
    # in-place update of an encoder output (same object, same address): must be noticed through the version counter
    scene["source_imgs_feat"].mul_(1.25)
    third = run(ren, b2, scene)
    fresh = UFOReconRenderer(sd, mode=UFO_MODE_FP32)
    want = run(fresh, b2, scene)
    fresh.close()
    scene["source_imgs_feat"].div_(1.25)
    for a, b in zip(third, want):
        assert torch.equal(a, b)
    assert not torch.equal(third[2], second[2])
    # explicit scope
    ren.begin_scene(batch, scene["source_imgs_feat"], scene["feature_volume"], scene["match_feature"])
    again = run(ren, batch, scene)
    ren.end_scene()
    assert torch.allclose(again[2], first[2], atol=1e-5)
    ren.close()


def test_scene_rejects_mismatched_shapes_and_bad_ray_indices(small):
    from uforecon_b200.renderer import HotPathWeights, Scene, render_rays
    batch, scene, sd = small
    for key, bad in (("depth_info", batch["depth_info"][:, :, ::2, ::2]), ("ray_d", batch["ray_d"][:, :, :100]),
                     ("cam_ray_d", batch["cam_ray_d"][:, :2]), ("w2cs", batch["w2cs"][:, :2]), ("source_poses", batch["source_poses"][:, :2])):
        b = dict(batch)
        b[key] = bad
        with pytest.raises(ValueError):
            Scene(b, scene["source_imgs_feat"], scene["feature_volume"], scene["match_feature"])
    fv = {k: dict(v) for k, v in scene["feature_volume"].items()}
    fv["stage2"]["weight_volume"] = fv["stage2"]["weight_volume"][:, :, :-1]
    with pytest.raises(ValueError):
        Scene(batch, scene["source_imgs_feat"], fv, scene["match_feature"])
    b = dict(batch)
    b["start_idx"] = 1                                    # training-style batch: w2cs would need NV + 1 rows
    with pytest.raises(ValueError):
        Scene(b, scene["source_imgs_feat"], scene["feature_volume"], scene["match_feature"])
    w = HotPathWeights(sd)
    sc = Scene(batch, scene["source_imgs_feat"], scene["feature_volume"], scene["match_feature"])
    u_c, u_f = synthetic.sampler_uniforms(4, seed=1)
    with pytest.raises(ValueError):
        render_rays(sc, w, torch.tensor([0, 5, 64 * 96, 7]), 4, u_c, u_f, UFO_MODE_FP32)
    with pytest.raises(ValueError):
        render_rays(sc, w, torch.tensor([0, -1, 3, 7]), 4, u_c, u_f, UFO_MODE_FP32)
    sc.close()
    w.close()


def test_start_idx_selects_the_w2c_rows(small):
    """ray_transformer.py:182,240: w2cs[:, s_idx:] - a batch with a leading reference-view row and start_idx = 1 must render
    like the inference batch without it."""
    from uforecon_b200.renderer import HotPathWeights, Scene, render_rays
    batch, scene, sd = small
    b = dict(batch)
    b["start_idx"] = 1
    b["w2cs"] = torch.cat([torch.eye(4)[None, None] * 3.0, batch["w2cs"]], 1)
    w = HotPathWeights(sd)
    u_c, u_f = synthetic.sampler_uniforms(64, seed=2)
    out = []
    for bb in (batch, b):
        sc = Scene(bb, scene["source_imgs_feat"], scene["feature_volume"], scene["match_feature"])
        out.append(render_rays(sc, w, None, 64, u_c, u_f, UFO_MODE_FP32, ray_begin=1000)["depth"].clone())
        sc.close()
    w.close()
    assert torch.equal(out[0], out[1])


@pytest.mark.parametrize("nv", [3, 5])
def test_compact_pair_maps_equal_reference_layout(nv):
    """N2: the match maps given once per pair ([NV(NV-1)/2,32,h,w]) render bit-identically to the reference's 2x redundant
    [NV,(NV-1)*32,h,w] stack built from the same maps."""
    from uforecon_b200.renderer import HotPathWeights, Scene, render_rays
    views = synthetic.UNFAVORABLE_VIEWS if nv == 3 else synthetic.TEN_VIEW_LIST[:nv]
    batch, scene, sd = make_case(views, (96, 64))
    w = HotPathWeights(sd)
    u_c, u_f = synthetic.sampler_uniforms(300, seed=4)
    outs = []
    for kw in (dict(), dict(pair_maps=scene["pair_maps"])):
        sc = Scene(batch, scene["source_imgs_feat"], scene["feature_volume"], None if kw else scene["match_feature"], **kw) if kw else \
            Scene(batch, scene["source_imgs_feat"], scene["feature_volume"], scene["match_feature"])
        for mode in (UFO_MODE_FP32, UFO_MODE_TC):
            r = render_rays(sc, w, None, 300, u_c, u_f, mode, ray_begin=900, want=("depth", "rgb"), taps=("sim8",))
            outs.append({k: v.clone() for k, v in r.items()})
        if kw:
            assert sc.device_bytes < ref_bytes
        else:
            ref_bytes = sc.device_bytes
        sc.close()
    w.close()
    for a, b in ((outs[0], outs[2]), (outs[1], outs[3])):
        for k in a:
            assert torch.equal(a[k], b[k]), k
