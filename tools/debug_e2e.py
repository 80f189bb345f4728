import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from conftest import load_golden, make_case, rel_err
from oracle import uforecon_oracle as orc
from uforecon_b200 import synthetic
from uforecon_b200._lib import UFO_MODE_FP32
from test_gpu_parity import run_cuda
g = load_golden("infer_nv3.npz")
W, H, seed, dr = [int(x) for x in g["meta"][:4]]; views = [int(x) for x in g["meta"][4:]]
batch, scene, sd = make_case(views, (W, H))
ray_idx = torch.from_numpy(g["ray_idx"]); u_c, u_f = synthetic.sampler_uniforms(len(ray_idx), seed=seed)
r = run_cuda(batch, scene, sd, ray_idx, u_c, u_f, UFO_MODE_FP32)
with torch.no_grad():
    o = orc.infer(batch, scene, sd, ray_idx, u_c, u_f, detail=True)
c = o["coarse"]
print("z_coarse", rel_err(r["z_coarse"], o["z_coarse"]))
print("srdf_coarse", rel_err(r["srdf_coarse"], c["srdf"]))
print("weight_coarse", rel_err(r["weight_coarse"], c["weight"]))
print("z_fine", rel_err(r["z_fine"], o["z_fine"]))
print("z", rel_err(r["z"], o["z"]))
print("srdf", rel_err(r["srdf"], o["srdf"]))
print("weight", rel_err(r["weight"], o["weight"]))
print("radiance", rel_err(r["radiance"], o["radiance"]))
print("depth", rel_err(r["depth"], o["depth"]), r["depth"][:4], o["depth"][:4])
print("depth from our weight*z", (r["weight"] * r["z"]).sum(1)[:4], " oracle weight*z", (o["weight"] * o["z"]).sum(1)[:4])
print("rgb", rel_err(r["rgb"], o["rgb"]))
print("---- density head")
import torch.nn.functional as F
RT = "ray_transformer."
x = r["ray_out"]
h = F.relu(F.linear(x, sd[RT+"DensityMLP.0.weight"], sd[RT+"DensityMLP.0.bias"]))
h2 = F.relu(F.linear(h, sd[RT+"DensityMLP.2.weight"], sd[RT+"DensityMLP.2.bias"]))
s = F.linear(h2, sd[RT+"DensityMLP.4.weight"], sd[RT+"DensityMLP.4.bias"])[..., 0]
print("torch on our ray_out:", s[0, :6])
print("ours               :", r["srdf"][0, :6])
print("ratio", (r["srdf"][0, :6] / s[0, :6]))
print("bias4", sd[RT+"DensityMLP.4.bias"], "w4", sd[RT+"DensityMLP.4.weight"])
