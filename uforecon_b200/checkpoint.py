"""State-dict layout contract of the hot path and a seed-deterministic synthetic checkpoint.

The reference loads ``pretrained/uforecon.ckpt`` (a Lightning file: ``{'state_dict': {...}}``) with
``strict=True`` (main.py:187).  That file is not available offline, so parity runs use
``synthetic_state_dict`` - identical keys and shapes for every tensor the hot path reads - and the
real file is picked up by ``load_hot_path_state`` whenever it exists.

Key list follows SURVEY.md appendix A.7 (module tree of code1/ray_transformer.py:86-163,
code1/attention/transformer.py:7-33, code1/encoder_utils/single_variance_network.py:5-11 and
code1/encoder_utils/fmt/TransMVSNet.py:23-33).
"""
from __future__ import annotations

import math
import os
from typing import Dict, Tuple

import torch

D_VIEW = 80   # 32 img feat + 24 volume + 16 similarity + 8 depth PE (ray_transformer.py:135)
D_RAY = 88    # D_VIEW + 8 sample-order PE (ray_transformer.py:138)
N_HEAD = 8


def _loftr_keys(prefix: str, d: int) -> Dict[str, Tuple[int, ...]]:
    p = prefix + ".layers.0."
    return {
        p + "q_proj.weight": (d, d), p + "k_proj.weight": (d, d), p + "v_proj.weight": (d, d),
        p + "merge.weight": (d, d),
        p + "mlp.0.weight": (2 * d, 2 * d), p + "mlp.2.weight": (d, 2 * d),
        p + "norm1.weight": (d,), p + "norm1.bias": (d,),
        p + "norm2.weight": (d,), p + "norm2.bias": (d,),
    }


def hot_path_layout() -> Dict[str, Tuple[int, ...]]:
    """name -> shape of every state-dict entry the hot path consumes (fp32 unless noted)."""
    rt = "ray_transformer."
    lay: Dict[str, Tuple[int, ...]] = {
        rt + "depthcode._freqs": (1, 8, 1), rt + "depthcode._phases": (1, 8, 1),
        rt + "pre_sim_mlp.0.weight": (32, 8), rt + "pre_sim_mlp.0.bias": (32,),
        rt + "pre_sim_mlp.2.weight": (32, 32), rt + "pre_sim_mlp.2.bias": (32,),
        rt + "pre_sim_mlp.4.weight": (16, 32), rt + "pre_sim_mlp.4.bias": (16,),
    }
    lay.update(_loftr_keys(rt + "density_view_transformer", D_VIEW))
    lay.update(_loftr_keys(rt + "density_ray_transformer", D_RAY))
    lay.update({
        rt + "DensityMLP.0.weight": (32, D_RAY), rt + "DensityMLP.0.bias": (32,),
        rt + "DensityMLP.2.weight": (16, 32), rt + "DensityMLP.2.bias": (16,),
        rt + "DensityMLP.4.weight": (1, 16), rt + "DensityMLP.4.bias": (1,),
        rt + "viewToken.view_token": (1, D_VIEW),
        rt + "linear_radianceweight_1_softmax.0.weight": (16, D_VIEW + 3),
        rt + "linear_radianceweight_1_softmax.0.bias": (16,),
        rt + "linear_radianceweight_1_softmax.2.weight": (8, 16),
        rt + "linear_radianceweight_1_softmax.2.bias": (8,),
        rt + "linear_radianceweight_1_softmax.4.weight": (1, 8),
        rt + "linear_radianceweight_1_softmax.4.bias": (1,),
        "deviation_network.variance": (),
    })
    # pixel-wise view-weight net of the cost-volume build (kernel 1); BN3d in eval mode
    pw = "transmvsnet.DepthNet.pixel_wise_net."
    for name, cin, cout in (("conv0", 1, 16), ("conv1", 16, 8)):
        lay[pw + name + ".conv.weight"] = (cout, cin, 1, 1, 1)
        lay[pw + name + ".bn.weight"] = (cout,)
        lay[pw + name + ".bn.bias"] = (cout,)
        lay[pw + name + ".bn.running_mean"] = (cout,)
        lay[pw + name + ".bn.running_var"] = (cout,)
        lay[pw + name + ".bn.num_batches_tracked"] = ()  # int64
    lay[pw + "conv2.weight"] = (1, 8, 1, 1, 1)
    lay[pw + "conv2.bias"] = (1,)
    return lay


def _shape_surface_head(sd: Dict[str, torch.Tensor], gen: torch.Generator) -> None:
    """Make the random SRDF head behave like a trained one: positive in front of a surface, a single zero
    crossing somewhere along the ray, negative behind, slope ~ -1 per unit of ray distance.

    With purely random weights the SRDF is negative at the first sample of most rays, NeuS gives that sample
    alpha = 1 and the rendered depth collapses onto ``near`` - a depth that does not depend on the network at all
    would make every depth tolerance in the tests vacuous.  The construction uses the one input the head sees
    un-mixed: the sample-order encoding ``sin(i/1000)`` that ``RayTransformer`` concatenates to the ray tokens
    (ray_transformer.py:165-173, 301-303; column 86 of the residual stream), plus a small random projection of
    the other features so that the crossing moves from ray to ray and the profile is not perfectly smooth.
    """
    rt = "ray_transformer."
    # keep the residual stream's order-encoding columns (80..87) close to the encoding itself
    sd[rt + "density_ray_transformer.layers.0.norm2.weight"][80:] *= 1e-3
    sd[rt + "density_ray_transformer.layers.0.norm2.bias"][80:] = 0.0
    w0, b0 = sd[rt + "DensityMLP.0.weight"], sd[rt + "DensityMLP.0.bias"]
    w2, b2 = sd[rt + "DensityMLP.2.weight"], sd[rt + "DensityMLP.2.bias"]
    w4, b4 = sd[rt + "DensityMLP.4.weight"], sd[rt + "DensityMLP.4.bias"]
    alpha, c, beta, s0 = 20.0, 0.2, 0.905, 1.0
    w0[0].zero_()
    w0[0, :80] = (torch.rand(80, generator=gen) * 2 - 1) * (0.05 / math.sqrt(80.0 / 3.0))
    w0[0, 86] = -alpha
    b0[0] = alpha * c                       # h0 = 4 - 20 sin(i/1000) + noise   (always > 0)
    w2[0].zero_()
    w2[0, 0] = 1.0
    b2[0] = 0.0                             # g0 = h0
    w2[1:, 0] = 0.0                         # the other hidden units do not see the ramp
    w4.mul_(0.1)
    w4[0, 0] = beta                         # srdf = 0.905 g0 - 2.62 + small random part
    b4[0] = s0 - beta * alpha * c


def synthetic_state_dict(seed: int = 0, inv_s: float = 64.0, surface: bool = True) -> Dict[str, torch.Tensor]:
    """Random but well-conditioned weights with the checkpoint's keys/shapes.

    Linear weights ~ U(+-sqrt(6/(fan_in+fan_out))) (the reference's xavier init, transformer.py:73-76),
    LayerNorm/bias/BN statistics perturbed away from their trivial defaults so they are exercised.
    ``variance`` is set so that ``inv_s = exp(10*variance)`` equals ``inv_s``.
    """
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for k, shp in hot_path_layout().items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.tensor(100, dtype=torch.int64)
        elif k.endswith("_freqs"):
            sd[k] = torch.repeat_interleave(math.pi * 2.0 ** torch.arange(0, 4), 2).view(1, -1, 1).float()
        elif k.endswith("_phases"):
            ph = torch.zeros(8)
            ph[1::2] = math.pi * 0.5
            sd[k] = ph.view(1, -1, 1)
        elif k == "deviation_network.variance":
            sd[k] = torch.tensor(math.log(inv_s) / 10.0, dtype=torch.float32)
        elif k.endswith("running_var"):
            sd[k] = 0.5 + torch.rand(shp, generator=gen)
        elif k.endswith("running_mean"):
            sd[k] = 0.1 * torch.randn(shp, generator=gen)
        elif "norm" in k and k.endswith("weight") or k.endswith("bn.weight"):
            sd[k] = 1.0 + 0.1 * torch.randn(shp, generator=gen)
        elif k.endswith("bias"):
            sd[k] = 0.1 * torch.randn(shp, generator=gen)
        elif k.endswith("view_token"):
            sd[k] = torch.randn(shp, generator=gen)
        else:  # linear / 1x1x1 conv weights
            fan_out = shp[0]
            fan_in = int(torch.tensor(shp[1:]).prod()) if len(shp) > 1 else 1
            bound = math.sqrt(6.0 / (fan_in + fan_out))
            sd[k] = (torch.rand(shp, generator=gen) * 2 - 1) * bound
    if surface:
        _shape_surface_head(sd, gen)
    return sd


def load_hot_path_state(path: str = "pretrained/uforecon.ckpt", seed: int = 0) -> Tuple[Dict[str, torch.Tensor], str]:
    """Hot-path tensors from the real checkpoint if present, else the synthetic one.

    Returns ``(state, source)`` with ``source`` in {"checkpoint", "synthetic"}.  A checkpoint that lacks
    a key or has a wrong shape raises - same failure mode as the reference's strict load.
    """
    if os.path.exists(path):
        ck = torch.load(path, map_location="cpu", weights_only=False)
        full = ck["state_dict"] if "state_dict" in ck else ck
        out = {}
        for k, shp in hot_path_layout().items():
            if k not in full:
                raise KeyError(f"checkpoint {path} lacks hot-path key {k}")
            if tuple(full[k].shape) != tuple(shp):
                raise ValueError(f"checkpoint key {k}: shape {tuple(full[k].shape)} != {shp}")
            out[k] = full[k].detach().clone()
        return out, "checkpoint"
    return synthetic_state_dict(seed), "synthetic"
