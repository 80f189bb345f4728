"""GPU: the BASELINE configuration itself (1600x1216 ray grid, NV=3, 4.3 GB scene) - parity on a sample of rays against
the oracle, and the size-independent properties of the path (rays are independent; results do not depend on which
rays share a call) on the full map.  The small-size suites hold every kernel to its tolerance; this file is about
what only shows at full size: > 2^31-byte volumes, 1.9 M-ray index ranges, the last rows of every map.
"""
import math

import pytest
import torch

from conftest import make_case, rel_err
from oracle import uforecon_oracle as orc
from uforecon_b200 import synthetic
from uforecon_b200._lib import UFO_MODE_FP32, UFO_MODE_TC_F16

pytestmark = pytest.mark.gpu

W, H = 1600, 1216
# north star: p99 depth error <= 0.5 % of the depth interval, colour PSNR >= 50 dB, for the tensor-core mode (fp16 operands).
# bf16 operands measured 5.6e-3 / 46 dB on this scene in round 1 - outside the bound - and were retired (ufo_render_rays
# rejects UFO_MODE_TC), so there is no second bound here.
P99_BOUND = 5e-3
PSNR_BOUND = 50.0
# strict reading of "depth interval" (SURVEY.md D3): cam.txt DEPTH_INTERVAL = 2.5 mm; reported next to the lenient one
DEPTH_INTERVAL_MM = 2.5


def to64(o, device=None):
    """float tensors of a nested batch / scene / state dict as float64 (optionally on `device`): the fp64 referee"""
    if torch.is_tensor(o):
        o = o.double() if o.is_floating_point() else o
        return o.to(device) if device is not None else o
    if isinstance(o, dict):
        return {k: to64(v, device) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return type(o)(to64(v, device) for v in o)
    return o


@pytest.fixture(scope="module")
def full():
    from uforecon_b200.renderer import HotPathWeights, Scene
    batch, scene, sd = make_case(synthetic.UNFAVORABLE_VIEWS, (W, H))
    w = HotPathWeights(sd)
    sc = Scene(batch, scene["source_imgs_feat"], scene["feature_volume"], scene["match_feature"])
    yield dict(batch=batch, scene=scene, sd=sd, w=w, sc=sc)
    sc.close()
    w.close()


def test_fullsize_fp32_sample_vs_oracle(full):
    """192 rays spread over the grid (corners and the very last pixel included), fp32 mode vs the oracle: every kernel
    isolated at the CUDA path's own sample positions (same bars as test_gpu_parity.py), then end to end."""
    from uforecon_b200.renderer import render_rays
    batch, scene, sd = full["batch"], full["scene"], full["sd"]
    g = torch.Generator().manual_seed(11)
    ray_idx = torch.cat([torch.tensor([0, W - 1, (H - 1) * W, H * W - 1]), torch.randint(0, H * W, (188,), generator=g)])
    n = len(ray_idx)
    u_c, u_f = synthetic.sampler_uniforms(n, seed=13)
    taps = ("sim8", "vol24", "tokens", "view_tok0", "ray_out", "radiance", "weight")
    r = render_rays(full["sc"], full["w"], ray_idx, n, u_c, u_f, UFO_MODE_FP32, want=("depth", "depth_z", "rgb", "srdf", "z", "points"),
                    taps=taps)
    torch.cuda.synchronize()
    r = {k: v.cpu() for k, v in r.items()}
    d = batch["ray_d"][0][:, ray_idx].t()
    pts = (batch["ray_o"][0][None, None] + r["z"][:, :, None] * d[:, None, :]).float()
    assert rel_err(r["points"], pts) <= 1e-6
    with torch.no_grad():
        k = orc.sample2rgb(batch, scene, sd, pts, r["z"], detail=True)          # oracle at the CUDA path's own samples
        o = orc.infer(batch, scene, sd, ray_idx, u_c, u_f, detail=True)         # oracle end to end
        # fp64 referee: the same restatement evaluated in float64 (on the GPU: the 4.3 GB scene as doubles) at the SAME fp32
        # sample positions.  It says how far the fp32 oracle (= the reference's own arithmetic) is from the true value
        # of its formulas at this resolution - one ulp of u is 1e-7 * 800 pixels here - and the kernel must be as close
        # to that true value as the reference is: |kernel - fp64| <= 2 |oracle_fp32 - fp64| + 1e-5 per tensor.
        k64 = orc.sample2rgb(to64(batch, "cuda"), to64(scene, "cuda"), to64(sd, "cuda"), pts.double().cuda(), r["z"].double().cuda(), detail=True)
        k64 = {a: (v.cpu() if torch.is_tensor(v) else v) for a, v in k64.items()}

    def split(t):
        t = t.view(n, 128, 3, 80)
        return {"feat": t[..., :32], "vol_tok": t[..., 32:56], "sim16": t[..., 56:72]}

    amb = ((k["uv"].abs() - 1).abs() < 2e-5).any(-1).any(0)                     # [RN,SN] mask-ambiguous samples
    ours = {"sim8": r["sim8"], "vol24": r["vol24"], **split(r["tokens"]), "view_tok0": r["view_tok0"], "ray_out": r["ray_out"],
            "srdf": r["srdf"], "radiance": r["radiance"][~amb]}

    def ref_of(d):
        return {"sim8": d["sim8"], "vol24": d["vol24"], **split(d["tokens"]), "view_tok0": d["view_out"].view(n, 128, 4, 80)[:, :, 0],
                "ray_out": d["ray_out"], "srdf": d["srdf"], "radiance": d["radiance"][~amb]}

    ka, k8 = ref_of(k), ref_of(k64)
    err = {a: rel_err(ours[a], ka[a]) for a in ours}            # kernel vs fp32 oracle
    err64 = {a: rel_err(ours[a], k8[a]) for a in ours}          # kernel vs fp64 referee
    orc64 = {a: rel_err(ka[a], k8[a]) for a in ours}            # fp32 oracle vs fp64 referee
    base = {a: (1e-5 if a in ("sim8", "vol24", "feat", "vol_tok", "sim16") else 1e-4) for a in ours}      # test_gpu_parity.py bars
    print("fullsize fp32 kernel vs fp32 oracle :", {a: f"{b:.1e}" for a, b in err.items()})
    print("fullsize fp32 kernel vs fp64 referee:", {a: f"{b:.1e}" for a, b in err64.items()})
    print("fullsize fp32 oracle vs fp64 referee:", {a: f"{b:.1e}" for a, b in orc64.items()})
    for a in ours:
        # (1) directly against the reference's arithmetic, same bars as the small cases: the kernels mirror the order of
        #     the reference's roundings in the projection / warps (sgemm k-order, unfused 3-D grid_sample), so no
        #     resolution-dependent allowance is needed
        assert err[a] <= base[a], (a, err[a])
        # (2) and as close to the true value of the formulas as the reference itself is
        assert err64[a] <= 2 * orc64[a] + base[a], (a, err64[a], orc64[a])
    with torch.no_grad():
        rgb, depth, _, weight = orc.render(r["z"], r["radiance"], r["srdf"], sd["deviation_network.variance"])
    comp = {"weight": rel_err(r["weight"], weight), "depth": rel_err(r["depth"], depth), "rgb": rel_err(r["rgb"], rgb)}
    clean = ~amb.any(1)
    e2e = {"z": rel_err(r["z"], o["z"]), "depth": rel_err(r["depth"], o["depth"]), "depth_z": rel_err(r["depth_z"], o["depth_z"]),
           "rgb": rel_err(r["rgb"][clean], o["rgb"][clean])}
    print("fullsize fp32 isolated compositing:", {a: f"{b:.1e}" for a, b in comp.items()})
    print("fullsize fp32 end to end:", {a: f"{b:.1e}" for a, b in e2e.items()}, "ambiguous rays", int((~clean).sum()))
    assert all(v <= 1e-5 for v in comp.values()), comp
    # tensor-core mode DIRECTLY against the oracle on the same rays and uniforms (not through the fp32 mode)
    t = render_rays(full["sc"], full["w"], ray_idx, n, u_c, u_f, UFO_MODE_TC_F16, want=("depth", "rgb"))
    torch.cuda.synchronize()
    span = float(batch["near_fars"][0, 0, 1] - batch["near_fars"][0, 0, 0])
    de = (t["depth"].cpu() - o["depth"]).abs() / span
    mse = float(((t["rgb"].cpu() - o["rgb"])[~amb.any(1)] ** 2).mean())
    print(f"fullsize tc16 vs oracle on {n} rays: depth err/interval p99 {float(de.quantile(0.99)):.2e} max {float(de.max()):.2e}, "
          f"colour PSNR {10 * math.log10(1.0 / max(mse, 1e-20)):.1f} dB")
    assert float(de.quantile(0.99)) <= P99_BOUND and 10 * math.log10(1.0 / max(mse, 1e-20)) >= PSNR_BOUND
    # end to end the inverse-CDF sampler divides by the coarse weight of the hit bin, so a fine sample in a bin of
    # near-zero weight moves by far more than the rounding upstream (bound in test_importance_sampler_isolated); the
    # rendered depth is insensitive to it, the colour of a ray whose moved sample carries weight is not
    assert e2e["depth"] <= 1e-4 and e2e["depth_z"] <= 1e-4 and e2e["z"] <= 1e-3 and e2e["rgb"] <= 1e-3, e2e


@pytest.mark.parametrize("mode", [UFO_MODE_TC_F16])
def test_fullsize_map_properties(full, mode):
    """Full 1.9 M-ray map in tensor-core mode: finite, inside the sampled range, bit-identical when row bands are
    rendered on their own, and within the north-star tolerance of the fp32 path on those bands."""
    from uforecon_b200.renderer import render_rays
    n = H * W
    g = torch.Generator(device="cuda").manual_seed(5)
    u_c = torch.rand(64, n, device="cuda", generator=g)
    u_f = torch.rand(64, n, device="cuda", generator=g)
    m = render_rays(full["sc"], full["w"], None, n, u_c, u_f, mode, ray_begin=0, want=("depth", "depth_z", "rgb"))
    torch.cuda.synchronize()
    assert bool(torch.isfinite(m["depth"]).all()) and bool(torch.isfinite(m["rgb"]).all())
    batch = full["batch"]
    cz = batch["cam_ray_d"][0][2].cuda()
    near, far = batch["near_fars"][0, 0, 0].item(), batch["near_fars"][0, 0, 1].item()
    # depth = sum w z with sum w <= 1 and z inside [near, far] / cam_ray_d.z (plus the half-bin jitter)
    assert float((m["depth"] * cz).max()) <= far * (1 + 1e-2) and float(m["depth"].min()) >= 0.0
    assert float(m["rgb"].min()) >= -1e-4 and float(m["rgb"].max()) <= 1 + 1e-4

    span = far - near
    de_all, se, cnt = [], 0.0, 0
    for row0, rows in ((0, 8), (601, 16), (H - 8, 8)):                        # first rows, an odd band, the last rows
        b, k = row0 * W, rows * W
        uc, uf = u_c[:, b:b + k].contiguous(), u_f[:, b:b + k].contiguous()
        band = render_rays(full["sc"], full["w"], None, k, uc, uf, mode, ray_begin=b, want=("depth", "depth_z", "rgb"))
        idx = torch.arange(b, b + k)
        listed = render_rays(full["sc"], full["w"], idx, k, uc, uf, mode, want=("depth", "rgb"))
        ref = render_rays(full["sc"], full["w"], None, k, uc, uf, UFO_MODE_FP32, ray_begin=b, want=("depth", "rgb", "z"))
        torch.cuda.synchronize()
        for key in ("depth", "depth_z", "rgb"):
            assert torch.equal(band[key], m[key][b:b + k]), (row0, key)
        assert torch.equal(listed["depth"], band["depth"]) and torch.equal(listed["rgb"], band["rgb"])
        # mask-ambiguous rays (see test_gpu_tc.py) are left out of the colour error only
        d = batch["ray_d"][0][:, b:b + k].t()
        pts = batch["ray_o"][0][None, None] + ref["z"].cpu()[:, :, None] * d[:, None, :]
        uv, _, _ = orc.project(batch["source_poses"][0], pts.float())
        amb = ((uv.abs() - 1).abs() < 2e-5).any(-1).any(0).any(1)
        de_all.append(((band["depth"] - ref["depth"]).abs() / span).cpu())
        diff = (band["rgb"] - ref["rgb"]).cpu()[~amb]
        se += float((diff ** 2).sum())
        cnt += diff.numel()
        print(f"fullsize mode {mode} band rows {row0}+{rows}: depth err/interval p50 {float(de_all[-1].median()):.2e} p99 {float(de_all[-1].quantile(0.99)):.2e} "
              f"max {float(de_all[-1].max()):.2e}; ambiguous {float(amb.float().mean()):.3f}; rgb mse {float((diff ** 2).mean()):.2e}")
    de = torch.cat(de_all)
    p99, psnr = float(de.quantile(0.99)), 10 * math.log10(1.0 / max(se / cnt, 1e-20))
    print(f"fullsize mode {mode}: depth err/interval p99 {p99:.2e}, colour PSNR {psnr:.1f} dB over {len(de)} rays")
    # strict reading of the tolerance ("report both"): the error in mm against 0.5 % of cam.txt's 2.5 mm
    # DEPTH_INTERVAL = 0.0125 mm - below fp16 operand rounding by construction; printed, not asserted
    mm = p99 * span * float(batch["scale_mat"][0, 0, 0]) if "scale_mat" in batch else float("nan")
    print(f"fullsize mode {mode}: strict reading: p99 depth error {mm:.3f} mm = {mm / DEPTH_INTERVAL_MM:.3f} of the 2.5 mm DEPTH_INTERVAL "
          f"(north-star bound read strictly: 0.005)")
    assert p99 <= P99_BOUND, p99
    assert psnr >= PSNR_BOUND, psnr
