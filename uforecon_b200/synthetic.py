"""Synthetic DTU-shaped inputs for the per-ray rendering hot path.

There is no dataset and no checkpoint on the GPU box, so tests, ``bench.py`` and the
golden-vector tool all feed the hot path from this one seeded generator:

* ``make_rig``      - 49 inward-looking cameras with DTU intrinsics.
* ``make_batch``    - the ``batch`` dict that the reference's eval dataset hands to
                      ``UFORecon.infer`` (reference: code1/dataset/dtu_test_sparse.py:382-436,
                      collated with batch size 1).  The camera normalisation follows
                      ``cal_scale_mat``/``scale_cam_info`` (dtu_test_sparse.py:281-360) and
                      ``get_boundingbox`` (code1/dataset/scene_transform.py:60-107) in closed form
                      (the reference goes through cv2.decomposeProjectionMatrix; K and [R|t] are
                      known here so no decomposition is needed).
* ``make_scene``    - the encoder-side tensors the hot path reads: FPN features, cascade
                      feature/weight frustum volumes, cross-view match maps (with the reference's
                      pair duplication) and the MVS depth maps.  Smooth random
                      fields of the right shape, range and layout - the encoder itself stays in
                      PyTorch and is not part of this repository.

Everything is CPU torch / numpy with explicit generators, so the same seed gives the same
tensors in the build container and on the GPU box.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

DTU_K = np.array([[2892.33, 0.0, 823.205], [0.0, 2883.18, 619.071], [0.0, 0.0, 1.0]], dtype=np.float64)
DTU_WH = (1600, 1200)
DTU_DEPTH_MIN = 425.0
DTU_DEPTH_INTERVAL = 2.5
#: canonical source-view orderings of the reference's eval scripts (main.py:78, script/eval_dtu_*.sh)
UNFAVORABLE_VIEWS = [1, 16, 36]
FAVORABLE_VIEWS = [23, 24, 33]
TEN_VIEW_LIST = [23, 24, 33, 22, 15, 34, 14, 32, 16, 35]
STAGE_DEPTHS = {"stage1": 48, "stage2": 32, "stage3": 8}
STAGE_SCALE = {"stage1": 4, "stage2": 2, "stage3": 1}


def _look_at(eye: np.ndarray) -> np.ndarray:
    z = -eye / np.linalg.norm(eye)
    up = np.array([0.0, 0.0, 1.0])
    x = np.cross(z, up)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z], 0)
    E = np.eye(4)
    E[:3, :3] = R
    E[:3, 3] = -R @ eye
    return E


def make_rig(n_views: int = 49) -> List[np.ndarray]:
    """World-to-camera 4x4 matrices (mm) of an inward-looking spherical-cap rig of radius 650 mm."""
    out = []
    for v in range(n_views):
        th = 2 * np.pi * (v % 7) / 28.0 + 0.1 * (v // 7)
        ph = 0.5 + 0.08 * (v // 7)
        eye = 650.0 * np.array([np.cos(th) * np.sin(ph), np.sin(th) * np.sin(ph), np.cos(ph)])
        out.append(_look_at(eye))
    return out


def _smooth_field(gen: torch.Generator, shape: Sequence[int], coarse: int = 8, fine_amp: float = 0.15) -> torch.Tensor:
    """Band-limited random field [..., H, W]: coarse noise upsampled bicubically plus a little white noise."""
    *lead, H, W = shape
    n = int(np.prod(lead)) if lead else 1
    h, w = max(2, H // coarse), max(2, W // coarse)
    lo = torch.randn(n, 1, h, w, generator=gen)
    up = F.interpolate(lo, size=(H, W), mode="bicubic", align_corners=True)
    up = up + fine_amp * torch.randn(n, 1, H, W, generator=gen)
    return up.reshape(*lead, H, W).contiguous()


def _frustum_corners(H: int, W: int, K: np.ndarray, c2w: np.ndarray, near: float, far: float) -> np.ndarray:
    xs = np.array([0, 0, W, W, 0, 0, W, W], dtype=np.float64)
    ys = np.array([0, H, 0, H, 0, H, 0, H], dtype=np.float64)
    ds = np.array([near] * 4 + [far] * 4, dtype=np.float64)
    pts = np.stack([(xs - K[0, 2]) * ds / K[0, 0], (ys - K[1, 2]) * ds / K[1, 1], ds, np.ones(8)], 0)
    return (c2w @ pts)[:3]


def make_batch(view_ids: Sequence[int], img_wh: Tuple[int, int], seed: int = 0, render_idx: int = 0,
               near: float = 425.0, far: float = 900.0, ndepths: int = 192) -> Dict[str, torch.Tensor]:
    """The eval ``batch`` dict (B=1) for source views ``view_ids`` rendered at ``img_wh`` = (W, H).

    Mirrors dtu_test_sparse.py: intrinsics rescaled non-uniformly to ``img_wh`` (:249-278), all
    extrinsics re-expressed in the first source camera's frame (:267), the scene normalised to the
    unit box of the joint view frusta x1.1 (:281-289), world->NDC projection matrices with the
    align_corners=True pixel->NDC map baked in (:405-416), rays of the render view (:418-427).
    """
    W, H = img_wh
    assert W % 32 == 0 and H % 32 == 0, "encoder needs H, W divisible by 32 (SURVEY.md F5)"
    rig = make_rig()
    NV = len(view_ids)
    sx, sy = W / DTU_WH[0], H / DTU_WH[1]
    K = DTU_K.copy()
    K[0] *= sx
    K[1] *= sy
    ref_w2c = rig[view_ids[0]]
    offset_dist = 25.0

    w2cs, render_w2cs = [], []
    for vid in view_ids:
        w2c = rig[vid]
        c2w = np.linalg.inv(w2c)
        rc2w = c2w.copy()
        rc2w[:3, 3] += rc2w[:3, 0] * offset_dist
        w2cs.append(w2c @ np.linalg.inv(ref_w2c))
        render_w2cs.append(np.linalg.inv(rc2w) @ np.linalg.inv(ref_w2c))
    w2cs32 = [m.astype(np.float32) for m in w2cs]
    K32 = K.astype(np.float32)

    # scale_mat: centre/radius of the joint frusta bounding box, radius x 1.1
    lo = np.full(3, np.inf)
    hi = np.full(3, -np.inf)
    for m in w2cs32:
        c = _frustum_corners(H, W, K32.astype(np.float64), np.linalg.inv(m.astype(np.float64)), near, far)
        lo = np.minimum(lo, c.min(1))
        hi = np.maximum(hi, c.max(1))
    center = (lo + hi) / 2
    radius = float((hi - lo).max() / 2 * 1.1)
    scale_mat = np.diag([radius, radius, radius, 1.0])
    scale_mat[:3, 3] = center
    scale_mat = scale_mat.astype(np.float32)
    scale_factor = np.float32(1.0 / radius)

    def _scaled(m: np.ndarray) -> np.ndarray:
        """w2c after the world is normalised by scale_mat (rotation kept, centre moved and scaled)."""
        m = m.astype(np.float64)
        R = m[:3, :3]
        C = -R.T @ m[:3, 3]
        Cn = (C - center) / radius
        out = np.eye(4)
        out[:3, :3] = R
        out[:3, 3] = -R @ Cn
        return out

    s_w2cs = np.stack([_scaled(m) for m in w2cs32])
    s_render = np.stack([_scaled(m) for m in render_w2cs])
    near_fars = []
    for m in s_w2cs:
        cam_o = -m[:3, :3].T @ m[:3, 3]
        dist = float(np.sqrt((cam_o ** 2).sum()))
        near_fars.append([0.95 * (dist - 1.0), 1.05 * (dist + 1.0)])

    w2cs_t = torch.from_numpy(np.float32(s_w2cs))
    render_t = torch.from_numpy(np.float32(s_render))
    intr = torch.from_numpy(K32)[None].repeat(NV, 1, 1)
    intr_pad = torch.eye(4)[None].repeat(NV, 1, 1)
    intr_pad[:, :3, :3] = intr
    nrm = torch.tensor([[2.0 / (W - 1), 0, -1, 0], [0, 2.0 / (H - 1), -1, 0], [0, 0, 1, 0], [0, 0, 0, 1]],
                       dtype=torch.float32)
    ref_pose = nrm @ (intr_pad @ render_t)[render_idx]
    source_poses = nrm @ (intr_pad @ w2cs_t)
    ref_pose_inv = torch.inverse(ref_pose)
    source_poses_inv = torch.inverse(source_poses)
    ray_o = ref_pose_inv[:3, -1]

    h_line = np.linspace(0, H - 1, H) * 2 / (H - 1) - 1
    w_line = np.linspace(0, W - 1, W) * 2 / (W - 1) - 1
    hm, wm = np.meshgrid(h_line, w_line, indexing="ij")
    homo = np.stack([wm.reshape(-1), hm.reshape(-1), np.ones(H * W), np.ones(H * W)])
    homo_t = torch.from_numpy(homo)
    tmp = (ref_pose_inv.double() @ homo_t)[:3] - ray_o.double()[:, None]
    ray_d = (tmp / torch.norm(tmp, dim=0)).float()
    cam = (torch.inverse(nrm @ intr_pad[0]).double() @ homo_t)[:3]
    cam_ray_d = (cam / torch.norm(cam, dim=0)).float()

    # multi-stage MVS projection matrices (mm world, relative to first source camera), :330-356
    proj = np.zeros((NV, 2, 4, 4), dtype=np.float32)
    for i, m in enumerate(w2cs32):
        k = K32.copy()
        k[:2] /= 4
        proj[i, 0] = m
        proj[i, 1, :3, :3] = k
    proj2, proj3 = proj.copy(), proj.copy()
    proj2[:, 1, :2, :] *= 2
    proj3[:, 1, :2, :] *= 4
    depth_interval = DTU_DEPTH_INTERVAL * 1.06
    depth_values = np.arange(DTU_DEPTH_MIN, depth_interval * ndepths + DTU_DEPTH_MIN, depth_interval, dtype=np.float32)

    gen = torch.Generator().manual_seed(seed)
    imgs = torch.sigmoid(1.2 * _smooth_field(gen, (NV, 3, H, W), coarse=16, fine_amp=0.05))

    batch = {
        "scale_mat": torch.from_numpy(scale_mat)[None],
        "scale_factor": torch.tensor([float(scale_factor)]),
        "w2cs": w2cs_t[None],
        "intrinsics": intr[None],
        "near_fars": torch.tensor(near_fars, dtype=torch.float32)[None],
        "source_imgs": imgs[None].contiguous(),
        "ref_img": imgs[render_idx][None].contiguous(),
        "ref_pose": ref_pose[None],
        "source_poses": source_poses[None],
        "ref_pose_inv": ref_pose_inv[None],
        "source_poses_inv": source_poses_inv[None],
        "ray_o": ray_o[None].contiguous(),
        "ray_d": ray_d[None].contiguous(),
        "cam_ray_d": cam_ray_d[None].contiguous(),
        "proj_matrices": {"stage1": torch.from_numpy(proj)[None], "stage2": torch.from_numpy(proj2)[None],
                          "stage3": torch.from_numpy(proj3)[None]},
        "depth_values_org_scale": torch.from_numpy(depth_values)[None],
        "start_idx": 0,
        "meta": ["synthetic-scan0-%08d" % render_idx],
    }
    return batch


def pair_list(n_views: int) -> List[Tuple[int, int]]:
    """Unordered view pairs in the reference's enumeration order (model.py:273-276 -> pair (a, b+1))."""
    return [(a, b + 1) for a in range(n_views - 1) for b in range(a, n_views - 1)]


def expand_pair_maps(pair_maps: torch.Tensor, n_views: int) -> torch.Tensor:
    """Compact ``[nC2, 32, h, w]`` pair maps -> the reference's ``[1, NV, (NV-1)*32, h, w]`` layout.

    Slot ``j`` of view ``v`` holds the map of the pair {v, other} where ``other`` is the j-th view
    different from ``v`` in ascending order (TransMVSNet.get_match_feat, TransMVSNet.py:341-374 and
    SURVEY.md F8: the same tensor is stored on both sides of a pair).
    """
    nC2, C, h, w = pair_maps.shape
    pairs = [(a, b) for a in range(n_views - 1) for b in range(a + 1, n_views)]
    assert len(pairs) == nC2
    per_view: List[List[torch.Tensor]] = [[] for _ in range(n_views)]
    for i, (a, b) in enumerate(pairs):
        per_view[a].append(pair_maps[i])
        per_view[b].append(pair_maps[i])
    return torch.stack([torch.cat(v, 0) for v in per_view], 0)[None].contiguous()


def make_scene(batch: Dict[str, torch.Tensor], seed: int = 1) -> Dict[str, object]:
    """Encoder-side inputs of ``infer`` for ``batch`` (shapes of SURVEY.md section 8b)."""
    _, NV, _, H, W = batch["source_imgs"].shape
    h, w = H // 4, W // 4
    gen = torch.Generator().manual_seed(seed)
    feats = _smooth_field(gen, (NV, 32, h, w), coarse=4)
    volumes = {}
    for stage, D in STAGE_DEPTHS.items():
        s = STAGE_SCALE[stage]
        hs, ws = H // s, W // s
        fv = _smooth_field(gen, (NV, 8, D, hs, ws), coarse=8 // s if s < 8 else 1, fine_amp=0.1)
        # decorrelate along depth a little so the z interpolation matters
        fv = fv + 0.5 * torch.randn(NV, 8, D, 1, 1, generator=gen)
        wv = torch.sigmoid(1.5 * _smooth_field(gen, (NV, 1, D, hs, ws), coarse=8 // s if s < 8 else 1, fine_amp=0.1))
        volumes[stage] = {"feature_volume": fv.contiguous(), "weight_volume": wv.contiguous()}
    npairs = NV * (NV - 1) // 2
    pair_maps = _smooth_field(gen, (npairs, 32, h, w), coarse=4)
    match = expand_pair_maps(pair_maps, NV)
    # MVS depth in normalised units: distance of each camera to the unit-box centre, wobbling +-0.15
    w2cs = batch["w2cs"][0]
    depth = []
    for v in range(NV):
        cam_o = -w2cs[v, :3, :3].T @ w2cs[v, :3, 3]
        depth.append(float(cam_o.norm()) + 0.15 * _smooth_field(gen, (H, W), coarse=16, fine_amp=0.02))
    depth_info = torch.stack(depth, 0)[None].contiguous()
    return {
        "source_imgs_feat": feats[None].contiguous(),
        "feature_volume": volumes,
        "match_feature": [match],
        "pair_maps": pair_maps.contiguous(),
        "depth_info": depth_info,
    }


def sampler_uniforms(n_rays: int, n_coarse: int = 64, n_fine: int = 64, seed: int = 1) -> Tuple[torch.Tensor, torch.Tensor]:
    """The two uniform draws of one ``infer`` call in the reference's order (sampler.py:42 then :86):
    ``torch.rand(n_coarse, RN)`` and ``torch.rand(n_fine, RN)`` from the global CPU generator."""
    gen = torch.Generator().manual_seed(seed)
    u_c = torch.rand(n_coarse, n_rays, generator=gen)
    u_f = torch.rand(n_fine, n_rays, generator=gen)
    return u_c, u_f
