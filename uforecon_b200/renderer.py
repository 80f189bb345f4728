"""Host-side mirror of the reference's per-ray rendering interface on top of the C ABI.

Same names, argument meaning and return values as the reference methods this replaces:

* ``UFOReconRenderer.infer(batch, ray_idx, source_imgs_feat, feature_volume, extract_geometry=True,
  match_feature=...)`` -> ``(srdf, points_x_all, depth, rgb)``       code1/model.py:393-478
* ``UFOReconRenderer.render_depth_map(...)`` - the chunk loop of ``extract_geometry``
  (code1/model.py:814-826): depth map in mm ``[H, W]`` and colour ``[H, W, 3]``.

PyTorch is only plumbing here (device memory, streams); all arithmetic of the path runs in
``libuforecon_b200.so``.  Nothing in this module falls back to PyTorch ops or to ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import UFO_MODE_FP32, UFO_MODE_TC, UFO_MODE_TC_F16, UFO_N_COARSE, UFO_N_FINE, UFO_N_SAMPLES

RT = "ray_transformer."
STAGES = ("stage1", "stage2", "stage3")


def _host_f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to("cpu", torch.float32).contiguous()


def _dev_f32(t: torch.Tensor, device: torch.device) -> torch.Tensor:
    return t.detach().to(device, torch.float32).contiguous()


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class HotPathWeights:
    """``ray_transformer.*`` + ``deviation_network.variance`` packed on the device."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device: Optional[torch.device] = None):
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda")
        keep: List[torch.Tensor] = []

        self.max_abs_weight = 0.0

        def p(name: str) -> int:
            if name not in state_dict:
                raise KeyError(f"state dict lacks hot-path key {name!r}")
            t = _host_f32(state_dict[name])
            if not bool(torch.isfinite(t).all()):
                raise ValueError(f"hot-path tensor {name!r} holds non-finite values")
            if name.endswith("weight") and t.numel():
                self.max_abs_weight = max(self.max_abs_weight, float(t.abs().max()))
            keep.append(t)
            return t.data_ptr()

        def loftr(prefix: str) -> _lib.UfoLoftrLayer:
            l = _lib.UfoLoftrLayer()
            l.q, l.k, l.v = p(prefix + "q_proj.weight"), p(prefix + "k_proj.weight"), p(prefix + "v_proj.weight")
            l.merge, l.mlp0, l.mlp2 = p(prefix + "merge.weight"), p(prefix + "mlp.0.weight"), p(prefix + "mlp.2.weight")
            l.norm1_w, l.norm1_b = p(prefix + "norm1.weight"), p(prefix + "norm1.bias")
            l.norm2_w, l.norm2_b = p(prefix + "norm2.weight"), p(prefix + "norm2.bias")
            return l

        def mlp3(prefix: str) -> _lib.UfoMlp3:
            m = _lib.UfoMlp3()
            m.w0, m.b0 = p(prefix + "0.weight"), p(prefix + "0.bias")
            m.w2, m.b2 = p(prefix + "2.weight"), p(prefix + "2.bias")
            m.w4, m.b4 = p(prefix + "4.weight"), p(prefix + "4.bias")
            return m

        d = _lib.UfoWeightsDesc()
        d.view = loftr(RT + "density_view_transformer.layers.0.")
        d.ray = loftr(RT + "density_ray_transformer.layers.0.")
        d.pre_sim = mlp3(RT + "pre_sim_mlp.")
        d.density = mlp3(RT + "DensityMLP.")
        d.radiance = mlp3(RT + "linear_radianceweight_1_softmax.")
        d.view_token = p(RT + "viewToken.view_token")
        d.depth_freqs = p(RT + "depthcode._freqs")
        d.depth_phases = p(RT + "depthcode._phases")
        d.variance = float(state_dict["deviation_network.variance"])
        # the tensor-core mode packs operands in fp16 (saturating at +-65504, 11-bit mantissa): weights far outside the
        # xavier scale it was validated on would clip or lose their small entries - say so instead of doing it silently
        if self.max_abs_weight > 1024.0:
            import warnings
            warnings.warn(f"hot-path weights reach |w| = {self.max_abs_weight:.3g}: UFO_MODE_TC_F16 packs operands in fp16 "
                          f"(validated on |w| <~ 1 only); check the result against UFO_MODE_FP32", RuntimeWarning)
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ufo_weights_create(C.byref(d), C.byref(self.handle), _stream_ptr(self.device)))

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self.lib.ufo_weights_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Scene:
    """One view set: what ``extract_geometry`` assembles before its chunk loop (model.py:777-808)."""

    def __init__(self, batch: Dict[str, torch.Tensor], source_imgs_feat: torch.Tensor,
                 feature_volume: Dict[str, Dict[str, torch.Tensor]], match_feature: Sequence[torch.Tensor],
                 device: Optional[torch.device] = None, pair_maps: Optional[torch.Tensor] = None):
        """``pair_maps`` [NV(NV-1)/2, 32, h, w]: the cross-view match maps with every pair stored once (pairs in the reference's
        order, model.py:273-276) - what an encoder without the reference's 2x redundant stack emits; when
        given, ``match_feature`` may be None."""
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda")
        if source_imgs_feat.shape[0] != 1:
            raise ValueError("batch size must be 1 (reference eval DataLoader, main.py:159-162)")
        B, NV, _, H, W = batch["source_imgs"].shape
        _, _, FC, h, w = source_imgs_feat.shape
        if FC != 32:
            raise ValueError(f"feature maps must have 32 channels, got {FC}")
        if pair_maps is not None:
            if tuple(pair_maps.shape) != (NV * (NV - 1) // 2, 32, h, w):
                raise ValueError(f"pair_maps shape {tuple(pair_maps.shape)} != {(NV * (NV - 1) // 2, 32, h, w)}")
        elif match_feature is None or tuple(match_feature[0].shape) != (1, NV, (NV - 1) * 32, h, w):
            raise ValueError(f"match_feature[0] must be {(1, NV, (NV - 1) * 32, h, w)}")
        if "depth_info" not in batch:
            raise KeyError("batch['depth_info'] missing (set by extract_geometry, model.py:806-808)")
        # every tensor the kernels index with NV / H*W strides is checked here: the library trusts these shapes
        # (the reference's grid_sample would resample a mismatched map; the fused kernels would read out of bounds)
        s_idx = batch.get("start_idx", 1)                     # ray_transformer.py:182: 0 in inference batches, 1 in training ones
        s_idx = int(s_idx.item()) if torch.is_tensor(s_idx) else int(s_idx)

        def need(name, t, shape):
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"{name} shape {tuple(t.shape)} != {tuple(shape)}")

        need("batch['source_imgs']", batch["source_imgs"], (1, NV, 3, H, W))
        need("source_imgs_feat", source_imgs_feat, (1, NV, 32, h, w))
        need("batch['depth_info']", batch["depth_info"], (1, NV, H, W))
        need("batch['ray_d']", batch["ray_d"], (1, 3, H * W))
        need("batch['cam_ray_d']", batch["cam_ray_d"], (1, 3, H * W))
        need("batch['ray_o']", batch["ray_o"], (1, 3))
        need("batch['ref_pose_inv']", batch["ref_pose_inv"], (1, 4, 4))
        need("batch['source_poses']", batch["source_poses"], (1, NV, 4, 4))
        need("batch['source_poses_inv']", batch["source_poses_inv"], (1, NV, 4, 4))
        if batch["w2cs"].dim() != 4 or batch["w2cs"].shape[0] != 1 or batch["w2cs"].shape[1] < s_idx + NV or tuple(batch["w2cs"].shape[2:]) != (4, 4):
            raise ValueError(f"batch['w2cs'] shape {tuple(batch['w2cs'].shape)}: need [1, >= start_idx + {NV}, 4, 4] (start_idx = {s_idx})")
        if batch["near_fars"].dim() != 3 or batch["near_fars"].shape[0] != 1 or batch["near_fars"].shape[1] < 1 or batch["near_fars"].shape[2] != 2:
            raise ValueError(f"batch['near_fars'] shape {tuple(batch['near_fars'].shape)}: need [1, >= 1, 2]")
        dev = self.device
        keep = []

        def dv(t):
            t = _dev_f32(t, dev)
            keep.append(t)
            return t.data_ptr()

        def hv(t):
            t = _host_f32(t)
            keep.append(t)
            return t.data_ptr()

        d = _lib.UfoSceneDesc()
        d.n_views, d.img_h, d.img_w, d.feat_h, d.feat_w = NV, H, W, h, w
        d.source_imgs = dv(batch["source_imgs"][0])
        d.img_feats = dv(source_imgs_feat[0])
        d.depth_info = dv(batch["depth_info"][0])
        if pair_maps is not None:
            d.match_pairs = dv(pair_maps)
        else:
            d.match_feats = dv(match_feature[0][0])
        for i, st in enumerate(STAGES):
            fv, wv = feature_volume[st]["feature_volume"], feature_volume[st]["weight_volume"]
            if fv.dim() != 5 or wv.dim() != 5 or fv.shape[0] != NV or fv.shape[1] != 8 or tuple(wv.shape) != (NV, 1) + tuple(fv.shape[2:]):
                raise ValueError(f"{st}: feature/weight volume shapes {tuple(fv.shape)} / {tuple(wv.shape)}: need [{NV},8,D,h,w] / [{NV},1,D,h,w]")
            d.vol_feat[i], d.vol_weight[i] = dv(fv), dv(wv)
            d.vol_d[i], d.vol_h[i], d.vol_w[i] = fv.shape[2], fv.shape[3], fv.shape[4]
        d.source_poses = hv(batch["source_poses"][0])
        d.source_poses_inv = hv(batch["source_poses_inv"][0])
        d.ref_pose_inv = hv(batch["ref_pose_inv"][0])
        d.w2cs = hv(batch["w2cs"][0][s_idx:s_idx + NV])      # ray_transformer.py:240: w2cs[:, s_idx:]
        d.near_fars = hv(batch["near_fars"][0])
        d.ray_o = hv(batch["ray_o"][0])
        d.ray_d = dv(batch["ray_d"][0])
        d.cam_ray_d = dv(batch["cam_ray_d"][0])
        self.n_views, self.H, self.W = NV, H, W
        self.scale = float(batch["scale_mat"][0][0, 0]) if "scale_mat" in batch else 1.0
        self.handle = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(self.lib.ufo_scene_create(C.byref(d), C.byref(self.handle), _stream_ptr(dev)))
            torch.cuda.current_stream(dev).synchronize()   # inputs in `keep` may be freed after this

    @property
    def device_bytes(self) -> int:
        return int(self.lib.ufo_scene_device_bytes(self.handle))

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self.lib.ufo_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


TAP_SHAPES = {
    "z_coarse": lambda n, nv: (n, 64), "weight_coarse": lambda n, nv: (n, 64), "srdf_coarse": lambda n, nv: (n, 64),
    "z_fine": lambda n, nv: (n, 64), "sim8": lambda n, nv: (n, 128, 8), "vol24": lambda n, nv: (n, 128, 24),
    "tokens": lambda n, nv: (n, 128, nv, 80), "view_tok0": lambda n, nv: (n, 128, 80),
    "ray_out": lambda n, nv: (n, 128, 88), "radiance": lambda n, nv: (n, 128, 3), "weight": lambda n, nv: (n, 128),
}


def draw_uniforms(n_rays: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """The sampler draws of one reference ``infer`` call, from torch's global CPU generator and in the
    reference's order: ``torch.rand(64, RN)`` (sampler.py:42) then ``torch.rand(64, RN)`` (sampler.py:86)."""
    return torch.rand(UFO_N_COARSE, n_rays), torch.rand(UFO_N_FINE, n_rays)


def render_rays(scene: Scene, weights: HotPathWeights, ray_idx: Optional[torch.Tensor], n_rays: int,
                u_coarse: torch.Tensor, u_fine: torch.Tensor, mode: int = UFO_MODE_FP32, ray_begin: int = 0,
                want: Sequence[str] = ("depth", "depth_z", "rgb"), taps: Sequence[str] = ()) -> Dict[str, torch.Tensor]:
    """Thin wrapper of ``ufo_render_rays`` with device tensors; returns the requested outputs/taps."""
    dev = scene.device
    lib = scene.lib
    u_c, u_f = _dev_f32(u_coarse, dev), _dev_f32(u_fine, dev)
    if u_c.shape != (UFO_N_COARSE, u_c.shape[1]) or u_f.shape != u_c.shape or u_c.shape[1] < n_rays:
        raise ValueError("u_coarse/u_fine must be [64, >=n_rays]")
    shapes = {"depth": (n_rays,), "depth_z": (n_rays,), "rgb": (n_rays, 3), "srdf": (n_rays, UFO_N_SAMPLES),
              "z": (n_rays, UFO_N_SAMPLES), "points": (n_rays, UFO_N_SAMPLES, 3)}
    res: Dict[str, torch.Tensor] = {}
    out = _lib.UfoRenderOut()
    for k in want:
        res[k] = torch.empty(shapes[k], dtype=torch.float32, device=dev)
        setattr(out, k, res[k].data_ptr())
    tp = _lib.UfoDebugTaps()
    for k in taps:
        res[k] = torch.empty(TAP_SHAPES[k](n_rays, scene.n_views), dtype=torch.float32, device=dev)
        setattr(tp, k, res[k].data_ptr())
    idx_ptr = None
    if ray_idx is not None:
        ray_idx = ray_idx.detach().to(dev, torch.int64).contiguous()
        if ray_idx.numel() != n_rays:
            raise ValueError("ray_idx size != n_rays")
        if n_rays > 0:
            lo, hi = torch.aminmax(ray_idx)
            if int(lo) < 0 or int(hi) >= scene.H * scene.W:
                raise ValueError(f"ray_idx values [{int(lo)}, {int(hi)}] outside the {scene.H}x{scene.W} ray grid")
        idx_ptr = ray_idx.data_ptr()
    with torch.cuda.device(dev):
        _lib.check(lib.ufo_render_rays(scene.handle, weights.handle, idx_ptr, ray_begin, n_rays, u_c.data_ptr(), u_f.data_ptr(),
                                       u_c.shape[1], mode, C.byref(out), C.byref(tp) if taps else None, _stream_ptr(dev)))
    return res


class UFOReconRenderer:
    """Drop-in for the reference model's ray-rendering methods (see module docstring)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device=None, mode: int = UFO_MODE_TC_F16,
                 test_ray_num: int = 800):
        # default = tensor cores with fp16 operands: inside the north-star tolerance at the full 1600x1216 size
        # (tests/test_gpu_fullsize.py); operand packing saturates at +-65504 (validated on synthetic xavier-scale weights
        # only - the pretrained checkpoint is not available offline; HotPathWeights warns about weights outside fp16 range)
        self.device = torch.device(device if device is not None else "cuda")
        self.weights = HotPathWeights(state_dict, self.device)
        self.mode = mode
        self.test_ray_num = test_ray_num
        self._scene_key = None
        self._scene: Optional[Scene] = None
        self._scene_refs = None
        self._pinned_scene: Optional[Scene] = None

    @staticmethod
    def _scene_inputs(batch, source_imgs_feat, feature_volume, match_feature):
        """Every tensor a Scene bakes in (not only the encoder outputs: the render view's rays, poses and MVS depth too)."""
        ts = [source_imgs_feat, match_feature[0]]
        for st in STAGES:
            ts += [feature_volume[st]["feature_volume"], feature_volume[st]["weight_volume"]]
        for k in ("source_imgs", "depth_info", "source_poses", "source_poses_inv", "ref_pose_inv", "w2cs", "near_fars", "ray_o",
                  "ray_d", "cam_ray_d", "scale_mat"):
            if k in batch and torch.is_tensor(batch[k]):
                ts.append(batch[k])
        return ts

    def _scene_for(self, batch, source_imgs_feat, feature_volume, match_feature) -> Scene:
        """The Scene of this call's inputs, rebuilt whenever ANY consumed tensor is another object or was written in place.

        Identity is (object id, in-place version counter) per tensor, and the cached entry keeps the tensors alive, so neither
        an address handed out again by the caching allocator nor an in-place update of the encoder outputs can alias a
        stale Scene (consecutive ``extract_geometry`` batches of one scan share the source views and differ in the render
        view's rays / poses only).  ``begin_scene`` / ``end_scene`` give the caller explicit control instead.
        """
        if self._pinned_scene is not None:
            return self._pinned_scene
        ts = self._scene_inputs(batch, source_imgs_feat, feature_volume, match_feature)
        s_idx = batch.get("start_idx", 1)
        key = tuple((id(t), t._version) for t in ts) + (int(s_idx.item()) if torch.is_tensor(s_idx) else int(s_idx),)
        if self._scene is None or key != self._scene_key:
            if self._scene is not None:
                self._scene.close()
            self._scene = Scene(batch, source_imgs_feat, feature_volume, match_feature, self.device)
            self._scene_key = key
            self._scene_refs = ts                      # keeps ids / addresses from being reused while cached
        return self._scene

    def begin_scene(self, batch, source_imgs_feat, feature_volume, match_feature) -> Scene:
        """Explicit scope: build the Scene of one ``extract_geometry`` call; ``infer`` / ``render_depth_map`` use it until ``end_scene``."""
        self.end_scene()
        self._pinned_scene = Scene(batch, source_imgs_feat, feature_volume, match_feature, self.device)
        return self._pinned_scene

    def end_scene(self):
        if self._pinned_scene is not None:
            self._pinned_scene.close()
            self._pinned_scene = None

    def infer(self, batch, ray_idx, source_imgs_feat, feature_volume=None, extract_geometry=False, match_feature=None,
              ray_idx_all=None, is_train=True):
        """``UFORecon.infer`` in its ``extract_geometry=True`` form (code1/model.py:393-478).

        ray_idx [1, RN] int64.  Uniforms are drawn here from torch's global CPU generator exactly as the
        reference's samplers do, so ``torch.manual_seed(s)`` gives the same sample positions.
        Returns (srdf [1,RN,128], points_x_all [1,RN,128,3], depth [1,RN], rgb [1,RN,3]).
        """
        if not extract_geometry:
            raise NotImplementedError("only the extract_geometry=True form of infer is on the hot path")
        if feature_volume is None or match_feature is None:
            raise ValueError("feature_volume and match_feature are required (canonical flag set, SURVEY.md section 5)")
        scene = self._scene_for(batch, source_imgs_feat, feature_volume, match_feature)
        RN = ray_idx.shape[1]
        u_c, u_f = draw_uniforms(RN)
        r = render_rays(scene, self.weights, ray_idx[0], RN, u_c, u_f, self.mode, want=("depth", "rgb", "srdf", "points"))
        return r["srdf"][None], r["points"][None], r["depth"][None], r["rgb"][None]

    def render_depth_map(self, batch, source_imgs_feat, feature_volume, match_feature, chunk: Optional[int] = None,
                         device_rng: bool = False):
        """The chunk loop of ``extract_geometry`` (model.py:814-826): depth [H,W] in mm, rgb [H,W,3].

        The whole ray grid goes through one library call.  By default the sampler uniforms are still drawn per
        reference chunk (``test_ray_num`` rays, coarse then fine) from torch's CPU generator, so the result is the one
        the reference's loop produces for the same seed and does not depend on how the library tiles the work.
        ``device_rng=True`` draws them on the GPU instead (torch's CUDA generator): statistically the same jitter, not
        the reference's bit stream - the 250 M CPU draws per 1600x1216 map otherwise cost about as much as the render.
        """
        scene = self._scene_for(batch, source_imgs_feat, feature_volume, match_feature)
        H, W = scene.H, scene.W
        n = H * W
        if device_rng:
            u_c = torch.rand(UFO_N_COARSE, n, device=self.device)
            u_f = torch.rand(UFO_N_FINE, n, device=self.device)
        else:
            step = self.test_ray_num
            u_c = torch.empty(UFO_N_COARSE, n, pin_memory=True)
            u_f = torch.empty(UFO_N_FINE, n, pin_memory=True)
            for s in range(0, n, step):                 # reference draw order: per 800-ray chunk, coarse then fine
                e = min(n, s + step)
                a, b = draw_uniforms(e - s)
                u_c[:, s:e] = a
                u_f[:, s:e] = b
        r = render_rays(scene, self.weights, None, n, u_c, u_f, self.mode, ray_begin=0, want=("depth_z", "rgb"))
        depth_mm = (r["depth_z"] * scene.scale).view(H, W)
        return depth_mm, r["rgb"].view(H, W, 3)

    def close(self):
        self.end_scene()
        if self._scene is not None:
            self._scene.close()
            self._scene = None
            self._scene_refs = None
        self.weights.close()
