import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from conftest import load_golden, make_case
from oracle import uforecon_oracle as orc
from uforecon_b200 import synthetic
from uforecon_b200._lib import UFO_MODE_FP32
from test_gpu_parity import run_cuda, ALL_TAPS
g = load_golden("infer_nv3.npz")
W, H, seed, dr = [int(x) for x in g["meta"][:4]]; views = [int(x) for x in g["meta"][4:]]
batch, scene, sd = make_case(views, (W, H))
ray_idx = torch.from_numpy(g["ray_idx"]); u_c, u_f = synthetic.sampler_uniforms(len(ray_idx), seed=seed)
r = run_cuda(batch, scene, sd, ray_idx, u_c, u_f, UFO_MODE_FP32)
z = r["z"]; d = batch["ray_d"][0][:, ray_idx].t()
pts = (batch["ray_o"][0][None, None] + z[:, :, None] * d[:, None, :]).float()
with torch.no_grad():
    o = orc.sample2rgb(batch, scene, sd, pts, z, detail=True)
RN = z.shape[0]
tok = o["tokens"].view(RN, 128, 3, 80)
diff = (r["tokens"][..., 72:] - tok[..., 72:]).abs()
print("max diff", diff.max().item())
flat = diff.flatten().topk(10)
for v, i in zip(flat.values, flat.indices):
    i = int(i); c = i % 8; n = (i // 8) % 3; s = (i // 24) % 128; ray = i // (24 * 128)
    uv = o["uv"][n, ray, s]
    p = pts[ray, s]
    w2c = batch["w2cs"][0, n]
    zc = float(w2c[2, :3] @ p + w2c[2, 3])
    ours, ref = float(r["tokens"][ray, s, n, 72 + c]), float(tok[ray, s, n, 72 + c])
    # infer delta from ours/ref at lowest freq component
    print(f"ray {ray} s {s} view {n} comp {c}: ours {ours:.6f} ref {ref:.6f} diff {float(v):.2e} uv ({uv[0]:.5f},{uv[1]:.5f}) zc {zc:.5f} mask {float(o['mask'][ray,s,n])}")
# histogram of diffs by component
for c in range(8):
    print("comp", c, "max", diff[..., c].max().item(), "mean", diff[..., c].mean().item())
# check rgb path
print("rgb_s diff", (o["rgb_s"] - 0).abs().max().item())
