"""GPU: the CUDA path through the C ABI against the oracle and the committed reference goldens.

Tolerances (north star): fp32 gather / sampler / compositing kernels within 1e-5 relative
(max |a-b| / max |ref|); each kernel is isolated by feeding the oracle the CUDA path's own upstream
tensors (taps), so that a tolerance is a statement about ONE kernel.  End to end (fp32 mode) the bar is
1e-4 on depth / rgb because the importance sampler and the NeuS alpha amplify upstream rounding.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, make_case, rel_err
from oracle import uforecon_oracle as orc
from uforecon_b200 import synthetic
from uforecon_b200._lib import UFO_MODE_FP32, UFO_MODE_TC_F16 as UFO_MODE_TC

pytestmark = pytest.mark.gpu

ALL_TAPS = ("z_coarse", "weight_coarse", "srdf_coarse", "z_fine", "sim8", "vol24", "tokens", "view_tok0", "ray_out",
            "radiance", "weight")


def run_cuda(batch, scene, sd, ray_idx, u_c, u_f, mode, taps=ALL_TAPS):
    from uforecon_b200.renderer import HotPathWeights, Scene, render_rays
    w = HotPathWeights(sd)
    sc = Scene(batch, scene["source_imgs_feat"], scene["feature_volume"], scene["match_feature"])
    r = render_rays(sc, w, ray_idx, len(ray_idx), u_c, u_f, mode, want=("depth", "depth_z", "rgb", "srdf", "z", "points"),
                    taps=taps)
    torch.cuda.synchronize()
    r = {k: v.cpu() for k, v in r.items()}
    sc.close()
    w.close()
    return r


@pytest.fixture(scope="module", params=["nv3", "nv5"])
def case(request):
    g = load_golden(f"infer_{request.param}.npz")
    W, H, seed, dr = [int(x) for x in g["meta"][:4]]
    views = [int(x) for x in g["meta"][4:]]
    batch, scene, sd = make_case(views, (W, H))
    ray_idx = torch.from_numpy(g["ray_idx"])
    u_c, u_f = synthetic.sampler_uniforms(len(ray_idx), seed=seed)
    with torch.no_grad():
        o = orc.infer(batch, scene, sd, ray_idx, u_c, u_f, detail=True)
    r = run_cuda(batch, scene, sd, ray_idx, u_c, u_f, UFO_MODE_FP32)
    return dict(g=g, batch=batch, scene=scene, sd=sd, ray_idx=ray_idx, u_c=u_c, u_f=u_f, o=o, r=r, nv=len(views), dr=dr)


def test_coarse_sampler(case):
    assert rel_err(case["r"]["z_coarse"], case["o"]["z_coarse"]) <= 1e-6
    assert rel_err(case["r"]["z_coarse"], case["g"]["z_coarse"]) <= 1e-6


def _pts_from_z(case, z):
    batch, ray_idx = case["batch"], case["ray_idx"]
    d = batch["ray_d"][0][:, ray_idx].t()
    return (batch["ray_o"][0][None, None] + z[:, :, None] * d[:, None, :]).float()


def mask_ambiguous(uv, eps=2e-5):
    """Points whose in-image test |u|<=1, |v|<=1 (grid_sample.py:12-15) is decided by the last bits of the projection:
    e.g. every sample of a bottom-row ray of the render view has v = 1.0000000 in source view 0 (the render camera is
    source camera 0 shifted along x).  The reference's own CPU and CUDA builds flip these differently, so the masked
    softmax over views (ray_transformer.py:316-317) is not comparable there.  uv [NV,RN,SN,2] -> [RN,SN] bool."""
    a = ((uv.abs() - 1).abs() < eps).any(-1)
    return a.any(0)


def test_gather_kernels_isolated(case):
    """projection + bilinear/trilinear gathers + similarity prior + volume blend + depth PE at the CUDA path's own z."""
    batch, scene, sd, r = case["batch"], case["scene"], case["sd"], case["r"]
    z = r["z"]
    pts = _pts_from_z(case, z)
    assert rel_err(r["points"], pts) <= 1e-6
    with torch.no_grad():
        o = orc.sample2rgb(batch, scene, sd, pts, z, detail=True)
    RN = z.shape[0]
    assert rel_err(r["sim8"], o["sim8"]) <= 1e-5
    assert rel_err(r["vol24"], o["vol24"]) <= 1e-5
    tok = o["tokens"].view(RN, 128, case["nv"], 80)
    # sim16 columns (56:72) go through pre_sim_mlp: same 1e-5 bar
    assert rel_err(r["tokens"][..., :32], tok[..., :32]) <= 1e-5
    assert rel_err(r["tokens"][..., 32:56], tok[..., 32:56]) <= 1e-5
    assert rel_err(r["tokens"][..., 56:72], tok[..., 56:72]) <= 1e-5
    # depth PE = sin(8*pi*2^k/8 * delta): away from the zero-padding cliff at the image border the bar is 2e-5;
    # on the cliff (|u| or |v| within two pixels of +-1) the MVS-depth sample falls from ~2 to 0 within one pixel,
    # so a 1e-7 difference in uv is amplified by (depth/pixel) * 8*pi - the reference on another device differs as much
    uv = o["uv"].permute(1, 2, 0, 3)                                        # [RN,SN,NV,2]
    W, H = case["batch"]["source_imgs"].shape[-1], case["batch"]["source_imgs"].shape[-2]
    cliff = ((uv[..., 0].abs() - 1).abs() < 4.0 / W) | ((uv[..., 1].abs() - 1).abs() < 4.0 / H)
    d_pe = (r["tokens"][..., 72:] - tok[..., 72:]).abs()
    assert float(d_pe[~cliff].max()) <= 2e-5
    assert float(d_pe.max()) <= 5e-3 and float(d_pe.mean()) <= 1e-5


def test_transformer_fp32_isolated(case):
    batch, scene, sd, r = case["batch"], case["scene"], case["sd"], case["r"]
    z = r["z"]
    pts = _pts_from_z(case, z)
    RN = z.shape[0]
    with torch.no_grad():
        o = orc.sample2rgb(batch, scene, sd, pts, z, detail=True)
    errs = {
        "view_tok0": rel_err(r["view_tok0"], o["view_out"].view(RN, 128, case["nv"] + 1, 80)[:, :, 0]),
        "ray_out": rel_err(r["ray_out"], o["ray_out"]),
        "srdf": rel_err(r["srdf"], o["srdf"]),
    }
    amb = mask_ambiguous(o["uv"])
    errs["radiance"] = rel_err(r["radiance"][~amb], o["radiance"][~amb])
    assert float(amb.float().mean()) < 0.1
    # fp32 CUDA-core GEMMs with a different summation order than ATen's: 1e-4 of the tensor scale
    assert all(v <= 1e-4 for v in errs.values()), errs


def test_compositing_isolated(case):
    r, sd = case["r"], case["sd"]
    with torch.no_grad():
        rgb, depth, opacity, weight = orc.render(r["z"], r["radiance"], r["srdf"], sd["deviation_network.variance"])
        _, _, _, weight_c = orc.render(r["z_coarse"], torch.zeros(r["z_coarse"].shape + (3,)), r["srdf_coarse"],
                                       sd["deviation_network.variance"])
    assert rel_err(r["weight"], weight) <= 1e-5
    assert rel_err(r["depth"], depth) <= 1e-5
    assert rel_err(r["rgb"], rgb) <= 1e-5
    assert rel_err(r["weight_coarse"], weight_c) <= 1e-5
    cz = case["batch"]["cam_ray_d"][0][2, case["ray_idx"]]
    assert rel_err(r["depth_z"], r["depth"] * cz) <= 1e-6


def test_importance_sampler_isolated(case):
    r, batch = case["r"], case["batch"]
    ray_idx = case["ray_idx"]
    d = batch["ray_d"][0][:, ray_idx].t()
    o = batch["ray_o"][0][None].expand_as(d)
    w, zc, u = r["weight_coarse"], r["z_coarse"], case["u_f"].t()
    _, z2 = orc.importance_sampler(o, d, w, zc, u)
    # conditioning of the inverse CDF: z = (u-lc)/(rc-lc+1e-6)*(zr-zl)+zl amplifies a 1e-7 difference in the cumsum by
    # (zr-zl)/(rc-lc+1e-6) in flat regions of the CDF; the bound is 1e-5 relative plus that term
    cdf = torch.cumsum(w, 1) / (w.sum(1, keepdim=True) + 1e-6)
    s = torch.minimum(torch.maximum(u, cdf[:, :1]), cdf[:, -1:])
    ri = torch.searchsorted(cdf, s.contiguous()).clamp(1, 63)
    gain = (torch.gather(zc, 1, ri) - torch.gather(zc, 1, ri - 1)) / (torch.gather(cdf, 1, ri) - torch.gather(cdf, 1, ri - 1) + 1e-6)
    bound = 1e-5 * zc.abs().max() + 4e-7 * torch.sort(gain, dim=1)[0]      # both outputs are sorted by z
    err = (r["z_fine"] - z2).abs()
    ok = err <= torch.maximum(bound, torch.sort(4e-7 * gain + 1e-5 * zc.abs().max(), dim=1, descending=True)[0])
    assert float(err.mean()) <= 2e-6
    assert float(err.max()) <= 1e-5 * float(zc.abs().max()) + 4e-7 * float(gain.max())
    z_all = torch.sort(torch.cat([r["z_coarse"], r["z_fine"]], 1), dim=1)[0]
    assert torch.equal(r["z"], z_all)          # merge is exact


def test_end_to_end_fp32_vs_oracle_and_golden(case):
    r, o, g = case["r"], case["o"], case["g"]
    clean = ~mask_ambiguous(o["uv"]).any(1)        # rays without a mask-ambiguous sample (colour only)
    for ref in (o, g):
        assert rel_err(r["depth"], ref["depth"]) <= 1e-4
        assert rel_err(r["rgb"][clean], torch.as_tensor(ref["rgb"])[clean]) <= 1e-4
        assert rel_err(r["z"], ref["z"]) <= 1e-4
        assert rel_err(r["srdf"], ref["srdf"]) <= 2e-4
    # taps at the CUDA path's own fine samples vs the golden's: the sample positions agree to 1e-4 (asserted above)
    # and the synthetic feature fields vary by O(1) over ~1e-1, so 2e-3 here; the kernels themselves are held to
    # 1e-5 at identical positions in test_gather_kernels_isolated
    dr = case["dr"]
    assert rel_err(r["sim8"][:dr], g["sim8_f"]) <= 2e-3
    assert rel_err(r["vol24"][:dr], g["vol24_f"]) <= 2e-3


def test_chunking_and_ray_range_invariance():
    """n shards == 1 shard: rays are independent, so any tiling gives identical results."""
    import os
    batch, scene, sd = make_case(synthetic.UNFAVORABLE_VIEWS, (96, 64))
    n = 300
    u_c, u_f = synthetic.sampler_uniforms(n, seed=5)
    from uforecon_b200.renderer import HotPathWeights, Scene, render_rays
    w = HotPathWeights(sd)
    sc = Scene(batch, scene["source_imgs_feat"], scene["feature_volume"], scene["match_feature"])
    full = render_rays(sc, w, None, n, u_c, u_f, UFO_MODE_FP32, ray_begin=1000)
    os.environ["UFO_FP32_CHUNK"] = "64"
    try:
        tiled = render_rays(sc, w, None, n, u_c, u_f, UFO_MODE_FP32, ray_begin=1000)
    finally:
        del os.environ["UFO_FP32_CHUNK"]
    idx = torch.arange(1000, 1000 + n)
    by_idx = render_rays(sc, w, idx, n, u_c, u_f, UFO_MODE_FP32)
    part = render_rays(sc, w, None, 100, u_c[:, 100:200].contiguous(), u_f[:, 100:200].contiguous(), UFO_MODE_FP32, ray_begin=1100)
    torch.cuda.synchronize()
    for k in ("depth", "rgb", "depth_z"):
        assert torch.equal(full[k], tiled[k]), k
        assert torch.equal(full[k], by_idx[k]), k
        assert torch.equal(full[k][100:200], part[k]), k
    sc.close()
    w.close()


def test_api_errors():
    from uforecon_b200 import _lib
    from uforecon_b200.renderer import HotPathWeights, Scene, render_rays
    batch, scene, sd = make_case(synthetic.UNFAVORABLE_VIEWS, (96, 64))
    w = HotPathWeights(sd)
    sc = Scene(batch, scene["source_imgs_feat"], scene["feature_volume"], scene["match_feature"])
    u_c, u_f = synthetic.sampler_uniforms(8, seed=5)
    with pytest.raises(_lib.UfoError):
        render_rays(sc, w, None, 8, u_c, u_f, UFO_MODE_FP32, ray_begin=96 * 64 - 4)   # range outside grid
    with pytest.raises(_lib.UfoError):
        render_rays(sc, w, None, 8, u_c, u_f, 7)                                        # unknown mode
    bad = dict(sd)
    del bad["ray_transformer.viewToken.view_token"]
    with pytest.raises(KeyError):
        HotPathWeights(bad)
    r = render_rays(sc, w, None, 0, u_c, u_f, UFO_MODE_FP32)                            # empty ray set is a no-op
    assert r["depth"].numel() == 0
    sc.close()
    w.close()


def test_asymmetric_match_maps_use_both_slots():
    """The reference samples view a's slot (b-1) at uv_a and view b's slot a at uv_b (model.py:273-285).  Its encoder
    stores the same map in both slots, which the library detects and exploits; when a caller passes
    different data in the two slots the result must still follow the reference."""
    from uforecon_b200.renderer import HotPathWeights, Scene, render_rays
    batch, scene, sd = make_case(synthetic.UNFAVORABLE_VIEWS, (96, 64))
    scene = dict(scene)
    m = scene["match_feature"][0].clone()
    m[0, 2, :32] += 0.25 * torch.randn(m[0, 2, :32].shape, generator=torch.Generator().manual_seed(3))   # view 2, slot 0 = pair (0,2)
    scene["match_feature"] = [m]
    n = 96
    ray_idx = torch.arange(0, 96 * 64, 64)[:n]
    u_c, u_f = synthetic.sampler_uniforms(n, seed=8)
    w = HotPathWeights(sd)
    sc = Scene(batch, scene["source_imgs_feat"], scene["feature_volume"], scene["match_feature"])
    r = render_rays(sc, w, ray_idx, n, u_c, u_f, UFO_MODE_FP32, want=("depth", "z"), taps=("sim8",))
    r = {k: v.cpu() for k, v in r.items()}
    d = batch["ray_d"][0][:, ray_idx].t()
    pts = (batch["ray_o"][0][None, None] + r["z"][:, :, None] * d[:, None, :]).float()
    with torch.no_grad():
        o = orc.sample2rgb(batch, scene, sd, pts, r["z"], detail=True)
    assert rel_err(r["sim8"], o["sim8"]) <= 2e-5
    sc.close()
    w.close()
