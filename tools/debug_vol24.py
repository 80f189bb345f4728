"""Debug aid: locate the largest full-size vol24 difference between the fp32 CUDA path and the oracle, and print the
oracle's intermediate quantities at that point (fp32 and fp64)."""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import uforecon_oracle as orc
from uforecon_b200 import synthetic, checkpoint
from uforecon_b200._lib import UFO_MODE_FP32
from uforecon_b200.renderer import HotPathWeights, Scene, render_rays
import torch.nn.functional as F

W, H = 1600, 1216
sd = checkpoint.synthetic_state_dict(0)
batch = synthetic.make_batch(synthetic.UNFAVORABLE_VIEWS, (W, H)); scene = synthetic.make_scene(batch); batch["depth_info"] = scene["depth_info"]
w = HotPathWeights(sd); sc = Scene(batch, scene["source_imgs_feat"], scene["feature_volume"], scene["match_feature"])
g = torch.Generator().manual_seed(11)
ray_idx = torch.cat([torch.tensor([0, W - 1, (H - 1) * W, H * W - 1]), torch.randint(0, H * W, (188,), generator=g)])
n = len(ray_idx)
u_c, u_f = synthetic.sampler_uniforms(n, seed=13)
r = render_rays(sc, w, ray_idx, n, u_c, u_f, UFO_MODE_FP32, want=("z", "points"), taps=("vol24",))
torch.cuda.synchronize()
r = {k: v.cpu() for k, v in r.items()}
d = batch["ray_d"][0][:, ray_idx].t()
pts = (batch["ray_o"][0][None, None] + r["z"][:, :, None] * d[:, None, :]).float()
print("points bit-equal:", torch.equal(pts, r["points"]))
P = batch["source_poses"][0]; nf = batch["near_fars"][0][0]
ref = orc.volume_blend(scene["feature_volume"], P, pts, (nf[0], nf[1]))
diff = (r["vol24"] - ref).abs()
print("rel err", float(diff.max() / ref.abs().max()), "max|ref|", float(ref.abs().max()))
idx = torch.nonzero(diff == diff.max())[0]
ri, si, ci = [int(v) for v in idx]
print("worst at ray", ri, "sample", si, "channel", ci, "ours", float(r["vol24"][ri, si, ci]), "oracle", float(ref[ri, si, ci]))
# histogram of errors
rel = diff / ref.abs().max()
for thr in (1e-7, 1e-6, 5e-6, 1e-5):
    print(f"frac > {thr:g}: {float((rel > thr).float().mean()):.2e}")
p1 = pts[ri:ri + 1, si:si + 1]
for dt in (torch.float32, torch.float64):
    G = Wn = None
    for nview in range(3):
        uv, z, _ = orc.project(P[nview:nview + 1].to(dt), p1.to(dt), (nf[0].to(dt), nf[1].to(dt)))
        grid = torch.cat([uv[0], z[0][..., None]], -1).view(1, 1, 1, 1, 3)
        fl, wl = [], None
        for st in ("stage1", "stage2", "stage3"):
            f = F.grid_sample(scene["feature_volume"][st]["feature_volume"][nview:nview + 1].to(dt), grid, mode="bilinear", align_corners=True, padding_mode="zeros")[0, :, 0, 0, 0]
            ww = F.grid_sample(scene["feature_volume"][st]["weight_volume"][nview:nview + 1].to(dt), grid, mode="bilinear", align_corners=True, padding_mode="zeros")[0, 0, 0, 0, 0]
            fl.append(f); wl = ww if wl is None else wl + ww
        f = torch.cat(fl)
        print(dt, "view", nview, "uvz", [float(v) for v in grid.view(-1)], "w_l", float(wl), "f[c]", float(f[ci]))
        G = f * wl if G is None else G + f * wl
        Wn = wl if Wn is None else Wn + wl
    print(dt, "G[c]", float(G[ci]), "W", float(Wn), "out", float(G[ci] / (Wn + 1e-8)))
sc.close(); w.close()
