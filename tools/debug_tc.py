"""GPU debug: tensor-core modes vs the oracle at the CUDA path's own sample positions."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from conftest import make_case, rel_err
from oracle import uforecon_oracle as orc
from uforecon_b200 import synthetic
from uforecon_b200._lib import UFO_MODE_FP32, UFO_MODE_TC_F16
from uforecon_b200.renderer import HotPathWeights, Scene, render_rays

TAPS = ("z_coarse", "weight_coarse", "srdf_coarse", "z_fine", "sim8", "vol24", "tokens", "view_tok0", "ray_out", "radiance", "weight")

def main():
    views = synthetic.UNFAVORABLE_VIEWS if len(sys.argv) < 2 else synthetic.TEN_VIEW_LIST[:int(sys.argv[1])]
    wh = (160, 128)
    batch, scene, sd = make_case(views, wh)
    n = 333
    ray_idx = torch.randperm(wh[0] * wh[1], generator=torch.Generator().manual_seed(0))[:n]
    u_c, u_f = synthetic.sampler_uniforms(n, seed=7)
    w = HotPathWeights(sd)
    sc = Scene(batch, scene["source_imgs_feat"], scene["feature_volume"], scene["match_feature"])
    span = float(batch["near_fars"][0, 0, 1] - batch["near_fars"][0, 0, 0])
    ref = render_rays(sc, w, ray_idx, n, u_c, u_f, UFO_MODE_FP32, want=("depth", "rgb", "srdf", "z"))
    torch.cuda.synchronize()
    for name, mode in (("fp16", UFO_MODE_TC_F16),):
        r = render_rays(sc, w, ray_idx, n, u_c, u_f, mode, want=("depth", "depth_z", "rgb", "srdf", "z", "points"), taps=TAPS)
        torch.cuda.synchronize()
        r = {k: v.cpu() for k, v in r.items()}
        z = r["z"]
        d = batch["ray_d"][0][:, ray_idx].t()
        pts = (batch["ray_o"][0][None, None] + z[:, :, None] * d[:, None, :]).float()
        with torch.no_grad():
            o = orc.sample2rgb(batch, scene, sd, pts, z, detail=True)
        nv = len(views)
        tok = o["tokens"].view(n, 128, nv, 80)
        print(f"== {name} nv={nv}")
        print("  tokens feat ", rel_err(r["tokens"][..., :32], tok[..., :32]), " vol", rel_err(r["tokens"][..., 32:56], tok[..., 32:56]),
              " sim16", rel_err(r["tokens"][..., 56:72], tok[..., 56:72]), " pe", float((r["tokens"][..., 72:] - tok[..., 72:]).abs().mean()))
        vo = o["view_out"].view(n, 128, nv + 1, 80)[:, :, 0]
        print("  view_tok0   ", rel_err(r["view_tok0"], vo), " mean abs", float((r["view_tok0"] - vo).abs().mean()), "scale", float(vo.abs().mean()))
        print("  ray_out     ", rel_err(r["ray_out"], o["ray_out"]), " mean abs", float((r["ray_out"] - o["ray_out"]).abs().mean()))
        print("  srdf        ", rel_err(r["srdf"], o["srdf"]), " mean abs", float((r["srdf"] - o["srdf"]).abs().mean()), "scale", float(o["srdf"].abs().mean()))
        print("  radiance    ", rel_err(r["radiance"], o["radiance"]), " mean abs", float((r["radiance"] - o["radiance"]).abs().mean()))
        rgb, depth, _, weight = orc.render(r["z"], r["radiance"], r["srdf"], sd["deviation_network.variance"])
        print("  compositing ", rel_err(r["depth"], depth), rel_err(r["rgb"], rgb))
        # end to end against the fp32 CUDA path (same uniforms)
        import math
        de = (r["depth"] - ref["depth"].cpu()).abs() / span
        mse = float(((r["rgb"] - ref["rgb"].cpu()) ** 2).mean())
        amb = ((o["uv"].abs() - 1).abs() < 2e-5).any(-1).any(0).any(1)          # rays with a mask-ambiguous sample
        with torch.no_grad():
            o32 = orc.sample2rgb(batch, scene, sd, (batch["ray_o"][0][None, None] + ref["z"].cpu()[:, :, None] * d[:, None, :]).float(), ref["z"].cpu(), detail=True)
        amb |= ((o32["uv"].abs() - 1).abs() < 2e-5).any(-1).any(0).any(1)
        er = (r["rgb"] - ref["rgb"].cpu()).abs().max(1)[0]
        top = er.argsort(descending=True)[:6]
        print("  top rgb err", [(int(i), round(float(er[i]), 5), bool(amb[i]), round(float(de[i]), 6)) for i in top], "n_amb", int(amb.sum()))
        mse_c = float(((r["rgb"] - ref["rgb"].cpu())[~amb] ** 2).mean())
        print(f"  PSNR without ambiguous rays {10 * math.log10(1.0 / max(mse_c, 1e-20)):.1f} dB")
        print(f"  e2e depth err/interval: p50 {float(de.median()):.3e} p99 {float(de.quantile(0.99)):.3e} max {float(de.max()):.3e};"
              f" rgb PSNR {10 * math.log10(1.0 / max(mse, 1e-20)):.1f} dB")
    sc.close(); w.close()

if __name__ == "__main__":
    main()
