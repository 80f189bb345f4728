"""Host-side mirror of the cost-volume build inside ``DepthNet.forward`` (kernel 1).

``similarity_volume`` takes what ``DepthNet.forward`` takes (code1/encoder_utils/fmt/TransMVSNet.py:49)
and returns the tensor it hands to ``cost_regularization`` (TransMVSNet.py:100-103) plus the view
weights of stage 1 (TransMVSNet.py:118-119).  The 3-D CNN regulariser stays in PyTorch.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from .renderer import _dev_f32, _host_f32, _stream_ptr

PW = "transmvsnet.DepthNet.pixel_wise_net."


def _pixelwise_desc(state_dict: Dict[str, torch.Tensor], keep: list) -> _lib.UfoPixelwiseNet:
    def p(name):
        t = _host_f32(state_dict[PW + name]).reshape(-1).contiguous()
        keep.append(t)
        return t.data_ptr()
    d = _lib.UfoPixelwiseNet()
    d.conv0_w = p("conv0.conv.weight")
    d.bn0_w, d.bn0_b, d.bn0_mean, d.bn0_var = p("conv0.bn.weight"), p("conv0.bn.bias"), p("conv0.bn.running_mean"), p("conv0.bn.running_var")
    d.conv1_w = p("conv1.conv.weight")
    d.bn1_w, d.bn1_b, d.bn1_mean, d.bn1_var = p("conv1.bn.weight"), p("conv1.bn.bias"), p("conv1.bn.running_mean"), p("conv1.bn.running_var")
    d.conv2_w = p("conv2.weight")
    d.conv2_b = float(state_dict[PW + "conv2.bias"].reshape(-1)[0])
    return d


def similarity_volume(features: Sequence[torch.Tensor], proj_matrices: torch.Tensor, depth_values: torch.Tensor,
                      state_dict: Optional[Dict[str, torch.Tensor]] = None, view_weights: Optional[torch.Tensor] = None,
                      device=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """features: V tensors [N,C,h,w] (slot 0 = reference); proj_matrices [N,V,2,4,4];
    depth_values [N,D,h,w]; view_weights [N,V-1,h,w] or None (stage 1).
    Returns (similarity [N,1,D,h,w], view_weights [N,V-1,h,w])."""
    lib = _lib.load()
    dev = torch.device(device if device is not None else "cuda")
    V = len(features)
    N, Cc, h, w = features[0].shape
    D = depth_values.shape[1]
    if tuple(proj_matrices.shape) != (N, V, 2, 4, 4):
        raise ValueError(f"proj_matrices shape {tuple(proj_matrices.shape)} != {(N, V, 2, 4, 4)}")
    if tuple(depth_values.shape) != (N, D, h, w):
        raise ValueError("depth_values must be [N,D,h,w]")
    keep: list = []
    fe = [_dev_f32(f, dev) for f in features]
    ptrs = (C.c_void_p * V)(*[f.data_ptr() for f in fe])
    # homographies exactly as the reference builds them (fp32 on the host, TransMVSNet.py:77-81 + fmt/module.py:340-342):
    # K[:3,:3] @ E[:3,:4] folded into E, then src . inverse(ref); rot = [:3,:3], trans = [:3,3]
    pm = proj_matrices.detach().float().cpu()
    ref_new = pm[:, 0, 0].clone()
    ref_new[:, :3, :4] = torch.matmul(pm[:, 0, 1, :3, :3], pm[:, 0, 0, :3, :4])
    ref_inv = torch.inverse(ref_new)
    rt = torch.empty(N, V - 1, 12)
    for i in range(1, V):
        src_new = pm[:, i, 0].clone()
        src_new[:, :3, :4] = torch.matmul(pm[:, i, 1, :3, :3], pm[:, i, 0, :3, :4])
        T = torch.matmul(src_new, ref_inv)
        rt[:, i - 1, :9] = T[:, :3, :3].reshape(N, 9)
        rt[:, i - 1, 9:] = T[:, :3, 3]
    proj = rt.contiguous()
    hyp = _dev_f32(depth_values, dev)
    sim = torch.empty(N, 1, D, h, w, dtype=torch.float32, device=dev)
    if view_weights is None:
        if state_dict is None:
            raise ValueError("stage 1 needs the pixel-wise net weights (state_dict)")
        pw = _pixelwise_desc(state_dict, keep)
        vw_out = torch.empty(N, V - 1, h, w, dtype=torch.float32, device=dev)
        vw_in_ptr, pw_ref = None, C.byref(pw)
    else:
        vw_in = _dev_f32(view_weights, dev)
        vw_out = vw_in
        vw_in_ptr, pw_ref = vw_in.data_ptr(), None
    with torch.cuda.device(dev):
        _lib.check(lib.ufo_costvolume_stage_rt(ptrs, N, V, Cc, h, w, D, proj.data_ptr(), hyp.data_ptr(), vw_in_ptr, pw_ref,
                                            sim.data_ptr(), vw_out.data_ptr() if view_weights is None else None,
                                            _stream_ptr(dev)))
        torch.cuda.current_stream(dev).synchronize()
    return sim, vw_out


def feature_grid(feats: torch.Tensor, batch: Dict[str, torch.Tensor], linear: Dict[str, torch.Tensor], volume_reso: int,
                 device=None) -> torch.Tensor:
    """``FeatureVolume.forward(feats, batch)`` up to its 3-D regulariser (code1/feature_volume.py:40-92), the alternative
    ``--volume_type featuregrid``.  feats [1,NV,32,h,w]; ``linear`` = state dict of ``FeatureVolume.linear``
    (keys '0.weight' ... '4.bias').  Returns ``volume_mean_var`` [1,16,Z,Y,X], the input of ``volume_regularization``."""
    lib = _lib.load()
    dev = torch.device(device if device is not None else "cuda")
    B, NV, Cc, h, w = feats.shape
    if B != 1 or Cc != 32:
        raise ValueError("feats must be [1, NV, 32, h, w]")
    keep = []

    def hp(name):
        t = _host_f32(linear[name]).reshape(-1).contiguous()
        keep.append(t)
        return t.data_ptr()

    m = _lib.UfoMlp3()
    m.w0, m.b0, m.w2, m.b2, m.w4, m.b4 = hp("0.weight"), hp("0.bias"), hp("2.weight"), hp("2.bias"), hp("4.weight"), hp("4.bias")
    f = _dev_f32(feats[0], dev)
    poses = _host_f32(batch["source_poses"][0])
    out = torch.empty(1, 16, volume_reso, volume_reso, volume_reso, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ufo_feature_grid(f.data_ptr(), NV, h, w, poses.data_ptr(), volume_reso, C.byref(m), out.data_ptr(), _stream_ptr(dev)))
        torch.cuda.current_stream(dev).synchronize()
    return out


class FusedDepthNet(torch.nn.Module):
    """``DepthNet`` with its warp / correlate / view-weight loop (TransMVSNet.py:69-100) replaced by kernel 1; everything
    after the similarity volume - the 3-D regulariser, softmax, winner-take-all depth, confidence - is the reference's own code
    path restated line by line (TransMVSNet.py:102-121).  Same call signature and return values as ``DepthNet.forward``."""

    def __init__(self, depthnet: torch.nn.Module):
        super().__init__()
        self.depthnet = depthnet

    def forward(self, features, proj_matrices, depth_values, num_depth, cost_regularization, prob_volume_init=None,
                view_weights=None, mvs_volume_only=False):
        assert depth_values.shape[1] == num_depth
        sd = {PW + k: v for k, v in self.depthnet.pixel_wise_net.state_dict().items()}
        dev = features[0].device
        sim, vw = similarity_volume(list(features), proj_matrices, depth_values, sd, view_weights=view_weights, device=dev)
        cost_reg = cost_regularization(sim)                                    # TransMVSNet.py:103
        if mvs_volume_only:
            return None
        prob_volume_pre = cost_reg.squeeze(1)
        if prob_volume_init is not None:
            prob_volume_pre = prob_volume_pre + prob_volume_init
        prob_volume = torch.exp(torch.nn.functional.log_softmax(prob_volume_pre, dim=1))
        idx = torch.argmax(prob_volume, dim=1, keepdim=True).type(torch.long)  # depth_wta, fmt/module.py:561-565
        depth = torch.gather(depth_values, 1, idx).squeeze(1)
        with torch.no_grad():
            conf = torch.max(prob_volume, dim=1)[0]
        out = {"depth": depth, "photometric_confidence": conf, "prob_volume": prob_volume, "depth_values": depth_values,
               "cost_volume": cost_reg}
        if view_weights is None:
            return out, vw.detach()
        return out


import contextlib  # noqa: E402


@contextlib.contextmanager
def fused_cost_volume(transmvsnet: torch.nn.Module):
    """Within the block the cascade of ``TransMVSNet.forward`` builds its three cost volumes with kernel 1."""
    orig = transmvsnet.DepthNet
    transmvsnet.DepthNet = FusedDepthNet(orig)
    try:
        yield transmvsnet.DepthNet
    finally:
        transmvsnet.DepthNet = orig
