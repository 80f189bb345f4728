import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from conftest import load_golden, make_case, rel_err
from oracle import uforecon_oracle as orc
from uforecon_b200 import synthetic
from uforecon_b200._lib import UFO_MODE_FP32
from test_gpu_parity import run_cuda
g = load_golden("infer_nv5.npz")
W, H, seed, dr = [int(x) for x in g["meta"][:4]]; views = [int(x) for x in g["meta"][4:]]
batch, scene, sd = make_case(views, (W, H))
ray_idx = torch.from_numpy(g["ray_idx"]); u_c, u_f = synthetic.sampler_uniforms(len(ray_idx), seed=seed)
r = run_cuda(batch, scene, sd, ray_idx, u_c, u_f, UFO_MODE_FP32)
z = r["z"]; d = batch["ray_d"][0][:, ray_idx].t()
pts = (batch["ray_o"][0][None, None] + z[:, :, None] * d[:, None, :]).float()
with torch.no_grad():
    o = orc.sample2rgb(batch, scene, sd, pts, z, detail=True)
diff = (r["radiance"] - o["radiance"]).abs().max(-1)[0]
print("n > 1e-4:", int((diff > 1e-4).sum()), "of", diff.numel())
for i in torch.nonzero(diff > 1e-4)[:10]:
    ray, s = int(i[0]), int(i[1])
    print("ray", ray, "s", s, "diff", float(diff[ray, s]), "ours", r["radiance"][ray, s], "ref", o["radiance"][ray, s])
    print("   uv", o["uv"][:, ray, s], "mask", o["mask"][ray, s], "z", float(z[ray, s]))
    print("   rgb_s", o["rgb_s"][ray, s])
