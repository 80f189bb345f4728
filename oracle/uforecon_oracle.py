"""ORACLE - CPU restatement of UFORecon's per-ray rendering hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
module; the product path (``uforecon_b200``) never does and fails loudly without its CUDA library.

What it is: a plain fp32 PyTorch-on-CPU restatement of the reference algorithm, function by function,
each citing the reference lines it follows (paths relative to the reference repository root).  The
reference is itself pure PyTorch, so the third-party arithmetic (``F.grid_sample``, ``layer_norm``,
``elu``, ``searchsorted`` ...) is the *same library* the reference calls; what is restated here is
the reference's own control flow and tensor algebra, without its einops/Lightning scaffolding.

Parity pinning: the reference ships no tests and no golden vectors, so the oracle is
pinned against outputs of the reference itself: ``tools/make_golden.py`` imports the unmodified
reference from /root/reference in the build container, runs ``UFORecon.infer``, ``query_cond_info``,
``query_depth_from_volume``, ``RayTransformer.forward``, ``VolumeRenderer.render``, both samplers and
``DepthNet.forward`` on the seeded synthetic inputs of ``uforecon_b200.synthetic`` and commits the
outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file against them.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

RT = "ray_transformer."


# ----------------------------------------------------------------------------------------------
# a1  FixedSampler.sample_ray            code1/encoder_utils/sampler.py:15-50
# ----------------------------------------------------------------------------------------------
def fixed_sampler(ray_o: torch.Tensor, ray_d: torch.Tensor, near: torch.Tensor, far: torch.Tensor,
                  u: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """ray_o, ray_d [RN,3]; near, far [RN]; u [SN,RN] uniforms -> points [RN,SN,3], z [RN,SN]."""
    SN = u.shape[0]
    lin = torch.from_numpy(np.linspace(0, 1, SN).astype("float32"))
    z = lin[:, None] * (far - near) + near                       # sampler.py:38  [SN,RN]
    interval = 1 / (SN - 1)
    z = z + (u - 0.5) * interval * (far - near)                  # sampler.py:42-43
    z = z.transpose(0, 1)
    pts = ray_o[:, None, :] + z[:, :, None] * ray_d[:, None, :]  # sampler.py:47
    return pts, z.clone()


# ----------------------------------------------------------------------------------------------
# a2  ImportanceSampler.sample_ray       code1/encoder_utils/sampler.py:74-108
# ----------------------------------------------------------------------------------------------
def importance_sampler(ray_o, ray_d, weight, z_val, u) -> Tuple[torch.Tensor, torch.Tensor]:
    """weight, z_val [RN,SN]; u [RN,SNf] uniforms -> sorted points [RN,SNf,3], z [RN,SNf]."""
    RN, SN = z_val.shape
    cdf = torch.cumsum(weight, dim=1) / (weight.sum(dim=1)[:, None] + 1e-6)
    s = torch.clamp(u, min=cdf[:, 0].view(-1, 1), max=cdf[:, -1].view(-1, 1))
    right = torch.searchsorted(cdf, s.contiguous())
    right[right == 0] = 1
    right[right > SN - 1] = SN - 1
    lc, rc = torch.gather(cdf, 1, right - 1), torch.gather(cdf, 1, right)
    zl, zr = torch.gather(z_val, 1, right - 1), torch.gather(z_val, 1, right)
    z_new = (s - lc) / (rc - lc + 1e-6) * (zr - zl) + zl
    pts = ray_o[:, None, :] + z_new[:, :, None] * ray_d[:, None, :]
    idx = torch.sort(z_new, dim=1)[1]
    return torch.gather(pts, 1, idx[..., None].expand(-1, -1, 3)), torch.gather(z_new, 1, idx)


# ----------------------------------------------------------------------------------------------
# a4  camera.get_coord_ref_ndc           code1/misc/camera.py:378-407
# ----------------------------------------------------------------------------------------------
def project(poses: torch.Tensor, pts: torch.Tensor, near_far: Optional[Sequence[float]] = None):
    """poses [NV,4,4] world->NDC; pts [RN,SN,3] -> uv [NV,RN,SN,2], z [NV,RN,SN] (camera depth, or
    near/far-normalised to [-1,1] when ``near_far`` is given), mask_z [NV,RN,SN] (q_z > 0)."""
    RN, SN, _ = pts.shape
    xh = torch.cat([pts, torch.ones_like(pts[..., :1])], -1).reshape(-1, 4)
    q = torch.matmul(poses, xh.t()[None])[:, :3]                # bmm(P, [x;1])  camera.py:387-388
    mask = (q[:, 2] > 0).float()
    uv = q[:, :2] / q[:, 2:3]
    z = q[:, 2]
    if near_far is not None:
        near, far = near_far
        z = (z - near) / (far - near)
        z = z * 2 - 1.0
    NV = poses.shape[0]
    return (uv.permute(0, 2, 1).reshape(NV, RN, SN, 2), z.reshape(NV, RN, SN), mask.reshape(NV, RN, SN))


def _bil(img: torch.Tensor, grid: torch.Tensor, align_corners: bool, pad: str) -> torch.Tensor:
    """img [C,H,W], grid [RN,SN,2] -> [RN,SN,C] through torch's grid_sample (the reference's call)."""
    return F.grid_sample(img[None], grid[None], mode="bilinear", align_corners=align_corners,
                         padding_mode=pad)[0].permute(1, 2, 0)


# ----------------------------------------------------------------------------------------------
# a5  UFORecon.query_cond_info           code1/model.py:218-305 (+ utils/gmflow_utils.py:80-83)
# ----------------------------------------------------------------------------------------------
def similarity_prior(match: torch.Tensor, uv: torch.Tensor) -> torch.Tensor:
    """match [NV,(NV-1)*32,h,w] (reference layout), uv [NV,RN,SN,2] -> feat_info [RN,SN,8]."""
    NV = match.shape[0]
    C = match.shape[1] // (NV - 1)
    sampled = [_bil(match[v], uv[v], True, "border") for v in range(NV)]      # model.py:251
    split = [torch.split(s, C, dim=-1) for s in sampled]                       # model.py:271
    sims = []
    for i in range(NV - 1):
        for j in range(i, NV - 1):                                             # model.py:273-276
            a = split[i][j].reshape(*split[i][j].shape[:-1], 8, C // 8)
            b = split[j + 1][i].reshape(*a.shape)
            sims.append(F.cosine_similarity(a, b, dim=-1))                     # model.py:280
    return torch.stack(sims, 0).mean(0)                                        # model.py:283


# ----------------------------------------------------------------------------------------------
# a6  UFORecon.query_depth_from_volume   code1/model.py:350-390
# ----------------------------------------------------------------------------------------------
def volume_blend(volumes: Dict[str, Dict[str, torch.Tensor]], poses: torch.Tensor, pts: torch.Tensor,
                 near_far: Sequence[float], stages=("stage1", "stage2", "stage3")) -> torch.Tensor:
    """volumes[stage]{feature_volume [NV,8,D,h,w], weight_volume [NV,1,D,h,w]} -> [RN,SN,24]."""
    NV = poses.shape[0]
    RN, SN, _ = pts.shape
    G_all = W_all = None
    for n in range(NV):
        uv, z, _ = project(poses[n:n + 1], pts, near_far)
        grid = torch.cat([uv[0], z[0][..., None]], -1).view(1, 1, RN, SN, 3)
        feats, w_l = [], None
        for st in stages:
            f = F.grid_sample(volumes[st]["feature_volume"][n:n + 1], grid, mode="bilinear",
                              align_corners=True, padding_mode="zeros")[0, :, 0].permute(1, 2, 0)
            w = F.grid_sample(volumes[st]["weight_volume"][n:n + 1], grid, mode="bilinear",
                              align_corners=True, padding_mode="zeros")[0, :, 0].permute(1, 2, 0)
            feats.append(f)
            w_l = w if w_l is None else w_l + w
        f_l = torch.cat(feats, -1)
        if n == 0:
            G_all, W_all = f_l * w_l, w_l
        else:
            G_all, W_all = G_all + f_l * w_l, W_all + w_l
    return G_all / (W_all + 1e-8)


# ----------------------------------------------------------------------------------------------
# a12/a13  LoFTREncoderLayer + LinearAttention   code1/attention/transformer.py:35-58,
#                                                code1/attention/linear_attention.py:20-47
# ----------------------------------------------------------------------------------------------
def loftr_layer(x: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str, nhead: int = 8) -> torch.Tensor:
    """x [N,L,d] -> [N,L,d]; bias-free projections, elu+1 linear attention, post-LN, concat-MLP."""
    N, L, d = x.shape
    D = d // nhead
    q = F.linear(x, sd[prefix + "q_proj.weight"]).view(N, L, nhead, D)
    k = F.linear(x, sd[prefix + "k_proj.weight"]).view(N, L, nhead, D)
    v = F.linear(x, sd[prefix + "v_proj.weight"]).view(N, L, nhead, D)
    Q, K = F.elu(q) + 1, F.elu(k) + 1
    v = v / L
    KV = torch.einsum("nshd,nshv->nhdv", K, v)
    Z = 1 / (torch.einsum("nlhd,nhd->nlh", Q, K.sum(dim=1)) + 1e-6)
    msg = torch.einsum("nlhd,nhdv,nlh->nlhv", Q, KV, Z) * L
    msg = F.linear(msg.reshape(N, L, d), sd[prefix + "merge.weight"])
    msg = F.layer_norm(msg, (d,), sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"], 1e-5)
    y = F.linear(F.relu(F.linear(torch.cat([x, msg], -1), sd[prefix + "mlp.0.weight"])), sd[prefix + "mlp.2.weight"])
    y = F.layer_norm(y, (d,), sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"], 1e-5)
    return x + y


def _mlp3(x, sd, prefix):
    x = F.relu(F.linear(x, sd[prefix + "0.weight"], sd[prefix + "0.bias"]))
    x = F.relu(F.linear(x, sd[prefix + "2.weight"], sd[prefix + "2.bias"]))
    return F.linear(x, sd[prefix + "4.weight"], sd[prefix + "4.bias"])


def order_posenc(d_hid: int, n: int) -> torch.Tensor:
    """Sample-order sinusoid table, float64 like the reference (ray_transformer.py:165-173)."""
    pos = np.arange(n, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    ang = pos / np.power(10000, 2 * (j // 2) / d_hid)
    tab = ang.copy()
    tab[:, 0::2] = np.sin(ang[:, 0::2])
    tab[:, 1::2] = np.cos(ang[:, 1::2])
    return torch.from_numpy(tab)


# ----------------------------------------------------------------------------------------------
# a7-a15  RayTransformer.forward         code1/ray_transformer.py:175-322
# ----------------------------------------------------------------------------------------------
def build_tokens(batch, feats, depth_info, pts, uv, mask_z, vol24, sim8, sd):
    """Per-view tokens and side inputs.  Returns dict with
    tokens [P,NV,80], rgb [RN,SN,NV,3], mask [RN,SN,NV], dir [RN,SN,NV,3]."""
    NV = feats.shape[0]
    RN, SN, _ = pts.shape
    imgs = batch["source_imgs"][0]
    w2cs = batch["w2cs"][0]
    o_ref = batch["ref_pose_inv"][0, :3, -1]
    sim16 = _mlp3(sim8, sd, RT + "pre_sim_mlp.")                                   # :268
    freqs = sd[RT + "depthcode._freqs"].view(-1)
    phases = sd[RT + "depthcode._phases"].view(-1)
    toks, rgbs, masks, dirs = [], [], [], []
    v1 = pts - o_ref
    v1 = v1 / torch.linalg.norm(v1, dim=-1, keepdim=True)                          # :185-188
    for v in range(NV):
        f = _bil(feats[v], uv[v], False, "zeros")                                  # :222
        rgbs.append(_bil(imgs[v], uv[v], False, "zeros"))                          # :224
        dm = _bil(depth_info[v][None], uv[v], False, "zeros")[..., 0]              # :236
        zc = (pts @ w2cs[v, :3, :3].t() + w2cs[v, :3, 3])[..., 2]                  # :240-243
        delta = (dm - zc)[..., None]
        pe = torch.sin(torch.addcmul(phases, delta, freqs))                        # :65-66
        toks.append(torch.cat([f, vol24, sim16, pe], -1))                          # :258-275
        g = uv[v]
        inb = ((g[..., 0] <= 1.) & (g[..., 0] >= -1.) & (g[..., 1] <= 1.) & (g[..., 1] >= -1.)).float()
        masks.append(inb * mask_z[v])                                              # :251-252
        o_v = batch["source_poses_inv"][0, v, :3, -1]
        v2 = pts - o_v
        v2 = v2 / torch.linalg.norm(v2, dim=-1, keepdim=True)
        dirs.append(v1 - v2)                                                       # :190
    return {"tokens": torch.stack(toks, 2).reshape(RN * SN, NV, -1), "rgb": torch.stack(rgbs, 2),
            "mask": torch.stack(masks, 2), "dir": torch.stack(dirs, 2)}


def ray_transformer(tok: Dict[str, torch.Tensor], sd: Dict[str, torch.Tensor], RN: int, SN: int):
    """tokens -> (radiance [RN,SN,3], srdf [RN,SN], view_out [P,NV+1,80], ray_out [RN,SN,88])."""
    x = tok["tokens"]
    P, NV, d = x.shape
    vt = sd[RT + "viewToken.view_token"].expand(P, 1, d)                           # :286-288
    x = loftr_layer(torch.cat([vt, x], 1), sd, RT + "density_view_transformer.layers.0.")
    a = x[:, 0].reshape(RN, SN, d)
    view_feat = x[:, 1:].reshape(RN, SN, NV, d)
    pe = order_posenc(8, SN).to(a)                                                 # :302 (.type_as)
    r = loftr_layer(torch.cat([a, pe[None].expand(RN, SN, 8)], -1), sd, RT + "density_ray_transformer.layers.0.")
    srdf = _mlp3(r, sd, RT + "DensityMLP.")[..., 0]                                # :307
    om = _mlp3(torch.cat([view_feat, tok["dir"]], -1), sd, RT + "linear_radianceweight_1_softmax.")[..., 0]
    om = torch.where(tok["mask"] == 0, torch.full_like(om, -1e9), om)              # :316
    p = torch.softmax(om, dim=-1)                                                  # :317
    radiance = (p[..., None] * tok["rgb"]).sum(2)                                  # :319
    return radiance, srdf, x, r


# ----------------------------------------------------------------------------------------------
# a16  VolumeRenderer.render             code1/encoder_utils/renderer.py:7-48
# ----------------------------------------------------------------------------------------------
def render(z: torch.Tensor, radiance: torch.Tensor, srdf: torch.Tensor, variance: torch.Tensor):
    """z, srdf [RN,SN]; radiance [RN,SN,3] -> rgb [RN,3], depth [RN], opacity [RN], weight [RN,SN]."""
    RN, SN = z.shape
    dz = z[:, 1:] - z[:, :-1]
    dz = torch.cat([dz[:, 0:1], dz, dz[:, -1:]], dim=1)
    interval = (dz[:, :-1] + dz[:, 1:]) / 2
    inv_s = torch.exp(variance * 10.0).clip(1e-6, 1e6)                             # single_variance_network.py:11
    iter_cos = -(0.5 + 0.0 + 1.0)                                                  # renderer.py:28-29  (= -1.5, F10)
    nxt = srdf + iter_cos * interval * 0.5
    prv = srdf - iter_cos * interval * 0.5
    prev_cdf, next_cdf = torch.sigmoid(prv * inv_s), torch.sigmoid(nxt * inv_s)
    alpha = ((prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)).clip(0.0, 1.0)
    T = torch.cumprod(torch.cat([torch.ones(RN, 1).to(alpha), 1. - alpha + 1e-7], -1), -1)[:, :-1]
    w = alpha * T
    return (radiance * w[:, :, None]).sum(1), (w * z).sum(1), w.sum(1), w


# ----------------------------------------------------------------------------------------------
# sample2rgb + infer                     code1/model.py:308-348, 393-482
# ----------------------------------------------------------------------------------------------
def sample2rgb(batch, scene, sd, pts, z, detail: bool = False):
    poses = batch["source_poses"][0]
    uv, _, mask_z = project(poses, pts)
    sim8 = similarity_prior(scene["match_feature"][0][0], uv)
    nf = batch["near_fars"][0][0]
    vol24 = volume_blend(scene["feature_volume"], poses, pts, (nf[0], nf[1]))
    tok = build_tokens(batch, scene["source_imgs_feat"][0], scene["depth_info"][0], pts, uv, mask_z, vol24, sim8, sd)
    RN, SN = z.shape
    radiance, srdf, view_out, ray_out = ray_transformer(tok, sd, RN, SN)
    rgb, depth, opacity, weight = render(z, radiance, srdf, sd["deviation_network.variance"])
    out = {"rgb": rgb, "depth": depth, "opacity": opacity, "weight": weight, "srdf": srdf, "radiance": radiance}
    if detail:
        out.update({"uv": uv, "mask_z": mask_z, "sim8": sim8, "vol24": vol24, "tokens": tok["tokens"],
                    "mask": tok["mask"], "dir": tok["dir"], "rgb_s": tok["rgb"], "view_out": view_out, "ray_out": ray_out})
    return out


def infer(batch, scene, sd, ray_idx: torch.Tensor, u_coarse: torch.Tensor, u_fine: torch.Tensor,
          detail: bool = False) -> Dict[str, torch.Tensor]:
    """``UFORecon.infer(..., extract_geometry=True)`` for rays ``ray_idx`` [RN] (model.py:393-478).

    u_coarse [SNc,RN], u_fine [SNf,RN] are the sampler uniforms in the reference's draw order.
    Returns srdf [RN,SN], z [RN,SN], points [RN,SN,3], depth [RN] (ray distance), rgb [RN,3],
    depth_z [RN] (= depth * cam_ray_d.z, model.py:818-821), plus the coarse pass under 'coarse'.
    """
    ray_d = batch["ray_d"][0][:, ray_idx].t()
    ray_o = batch["ray_o"][0][None].expand_as(ray_d)
    cz = batch["cam_ray_d"][0][2, ray_idx]
    near = batch["near_fars"][0, 0, 0] / cz                                         # model.py:416-427
    far = batch["near_fars"][0, 0, 1] / cz
    pts, z = fixed_sampler(ray_o, ray_d, near, far, u_coarse)
    c = sample2rgb(batch, scene, sd, pts.float(), z, detail)
    pts2, z2 = importance_sampler(ray_o, ray_d, c["weight"], z, u_fine.transpose(0, 1))
    pts_all = torch.cat([pts, pts2], 1)
    z_all = torch.cat([z, z2], 1)
    idx = torch.sort(z_all, dim=1)[1]                                               # model.py:468
    z_all = torch.gather(z_all, 1, idx)
    pts_all = torch.gather(pts_all, 1, idx[..., None].expand(-1, -1, 3))
    f = sample2rgb(batch, scene, sd, pts_all.float(), z_all, detail)
    f.update({"z": z_all, "points": pts_all, "depth_z": f["depth"] * cz, "coarse": c, "z_coarse": z, "z_fine": z2})
    return f


# ----------------------------------------------------------------------------------------------
# a18  cost-volume build   code1/encoder_utils/fmt/TransMVSNet.py:49-100, fmt/module.py:329-367
# ----------------------------------------------------------------------------------------------
def _pixelwise_net(sim: torch.Tensor, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """sim [B,1,D,H,W] -> view weight [B,1,H,W]  (PixelwiseNet, TransMVSNet.py:23-41; BN in eval mode)."""
    pw = "transmvsnet.DepthNet.pixel_wise_net."
    x = sim
    for name in ("conv0", "conv1"):
        x = F.conv3d(x, sd[pw + name + ".conv.weight"])
        x = F.batch_norm(x, sd[pw + name + ".bn.running_mean"], sd[pw + name + ".bn.running_var"],
                         sd[pw + name + ".bn.weight"], sd[pw + name + ".bn.bias"], False, 0.1, 1e-5)
        x = F.relu(x)
    x = F.conv3d(x, sd[pw + "conv2.weight"], sd[pw + "conv2.bias"]).squeeze(1)
    return torch.sigmoid(x).max(dim=1, keepdim=True)[0]


def homo_warp(src_fea, src_proj, ref_proj, depth_values):
    """fmt/module.py:329-367.  src_fea [B,C,H,W]; projs [B,4,4]; depth_values [B,D,H,W] -> [B,C,D,H,W]."""
    B, C, H, W = src_fea.shape
    D = depth_values.shape[1]
    proj = torch.matmul(src_proj, torch.inverse(ref_proj))
    rot, trans = proj[:, :3, :3], proj[:, :3, 3:4]
    y, x = torch.meshgrid(torch.arange(0, H, dtype=torch.float32), torch.arange(0, W, dtype=torch.float32), indexing="ij")
    xyz = torch.stack((x.reshape(-1), y.reshape(-1), torch.ones(H * W)))[None].repeat(B, 1, 1)
    rot_xyz = torch.matmul(rot, xyz)
    p = rot_xyz.unsqueeze(2).repeat(1, 1, D, 1) * depth_values.view(B, 1, D, -1) + trans.view(B, 3, 1, 1)
    invalid = (p[:, 2:3] < 1e-6).squeeze(1)
    xy = p[:, :2] / p[:, 2:3]
    gx = xy[:, 0] / ((W - 1) / 2) - 1
    gy = xy[:, 1] / ((H - 1) / 2) - 1
    gx[invalid] = -99.
    gy[invalid] = -99.
    grid = torch.stack((gx, gy), dim=3)
    out = F.grid_sample(src_fea, grid.view(B, D * H, W, 2), mode="bilinear", padding_mode="zeros", align_corners=True)
    return out.view(B, C, D, H, W)


def cost_volume_stage(features: List[torch.Tensor], proj_matrices: torch.Tensor, depth_values: torch.Tensor,
                      sd: Dict[str, torch.Tensor], view_weights: Optional[torch.Tensor] = None):
    """``DepthNet.forward`` up to (not including) ``cost_regularization`` (TransMVSNet.py:61-100).

    features: V tensors [B,C,H,W] (slot 0 = reference); proj_matrices [B,V,2,4,4];
    depth_values [B,D,H,W]; view_weights [B,V-1,H,W] or None (stage 1: computed by PixelwiseNet).
    Returns (similarity [B,1,D,H,W], view_weights [B,V-1,H,W]).
    """
    projs = torch.unbind(proj_matrices, 1)
    ref_fea, src_feas = features[0], features[1:]
    ref_proj = projs[0]
    ref_new = ref_proj[:, 0].clone()
    ref_new[:, :3, :4] = torch.matmul(ref_proj[:, 1, :3, :3], ref_proj[:, 0, :3, :4])
    sim_sum, w_sum, vws = 0, 1e-5, []
    for i, (src_fea, src_proj) in enumerate(zip(src_feas, projs[1:])):
        src_new = src_proj[:, 0].clone()
        src_new[:, :3, :4] = torch.matmul(src_proj[:, 1, :3, :3], src_proj[:, 0, :3, :4])
        warped = homo_warp(src_fea, src_new, ref_new, depth_values)
        sim = (warped * ref_fea.unsqueeze(2)).mean(1, keepdim=True)
        if view_weights is None:
            vw = _pixelwise_net(sim, sd)
            vws.append(vw)
        else:
            vw = view_weights[:, i:i + 1]
        sim_sum = sim_sum + sim * vw.unsqueeze(1)
        w_sum = w_sum + vw.unsqueeze(1)
    out_vw = torch.cat(vws, 1) if view_weights is None else view_weights
    return sim_sum / w_sum, out_vw


# ----------------------------------------------------------------------------------------------
# a19 (alt)  FeatureVolume.forward up to the 3-D regulariser      code1/feature_volume.py:40-92
# ----------------------------------------------------------------------------------------------
def feature_grid_meanvar(feats: torch.Tensor, source_poses: torch.Tensor, lin: Dict[str, torch.Tensor], reso: int) -> torch.Tensor:
    """feats [NV,32,h,w]; source_poses [NV,4,4] world->NDC; lin = {'0.weight','0.bias','2.weight',...,'4.bias'} of
    ``FeatureVolume.linear`` (32->32->16->8).  Returns the tensor handed to ``volume_regularization``:
    [16, Z, Y, X] = (masked mean | masked variance over views) of the compressed features on the reso^3 grid."""
    NV = feats.shape[0]
    line = np.linspace(0, reso - 1, reso) * 2 / (reso - 1) - 1                       # feature_volume.py:23-25 (float64)
    x, y, z = np.meshgrid(line, line, line, indexing="ij")
    xyz = torch.tensor(np.stack([x, y, z])).type_as(source_poses).reshape(3, -1)     # :48-49
    homo = torch.cat([xyz, torch.ones_like(xyz[0:1])], 0)                            # [4, XYZ]
    q = (source_poses @ homo[None].expand(NV, 4, -1))[:, :3]                         # :55-56
    mask_z = (q[:, 2] > 0).float()                                                   # :57-58
    uv = (q / q[:, 2:3])[:, :2].permute(0, 2, 1)                                     # [NV, XYZ, 2]   :61-63
    inb = ((uv[..., 0] <= 1.) & (uv[..., 0] >= -1.) & (uv[..., 1] <= 1.) & (uv[..., 1] >= -1.)).float()
    f = F.grid_sample(feats, uv[:, :, None, :], mode="bilinear", padding_mode="zeros", align_corners=False)[..., 0]  # [NV,32,XYZ]
    mask = inb * mask_z                                                              # :74
    weight = (mask / (mask.sum(0, keepdim=True) + 1e-8))[..., None]                  # [NV, XYZ, 1]  :79-80
    c = f.permute(0, 2, 1)                                                           # [NV, XYZ, 32]
    c = F.relu(F.linear(c, lin["0.weight"], lin["0.bias"]))
    c = F.relu(F.linear(c, lin["2.weight"], lin["2.bias"]))
    c = F.linear(c, lin["4.weight"], lin["4.bias"])                                  # [NV, XYZ, 8]  :83
    mean = (c * weight).sum(0, keepdim=True)                                         # :86
    var = (weight * (c - mean) ** 2).sum(0)                                          # :87
    mv = torch.cat([mean[0], var], -1).view(reso, reso, reso, 16)                    # [X,Y,Z,C]     :91
    return mv.permute(3, 2, 1, 0).contiguous()                                       # [C,Z,Y,X]     :92
