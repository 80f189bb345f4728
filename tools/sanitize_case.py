"""Tiny end-to-end render for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
    compute-sanitizer --tool memcheck python tools/sanitize_case.py [nv] [mode]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from conftest import make_case  # noqa: E402
from uforecon_b200 import synthetic  # noqa: E402
from uforecon_b200.renderer import HotPathWeights, Scene, render_rays  # noqa: E402


def main():
    nv = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    mode = int(sys.argv[2]) if len(sys.argv) > 2 else 2      # UFO_MODE_TC_F16 (0 = fp32; 1 = bf16 is retired)
    views = synthetic.UNFAVORABLE_VIEWS if nv == 3 else synthetic.TEN_VIEW_LIST[:nv]
    batch, scene, sd = make_case(views, (96, 64))
    n = 67
    u_c, u_f = synthetic.sampler_uniforms(n, seed=2)
    w = HotPathWeights(sd)
    sc = Scene(batch, scene["source_imgs_feat"], scene["feature_volume"], scene["match_feature"])
    r = render_rays(sc, w, None, n, u_c, u_f, mode, ray_begin=500, want=("depth", "rgb"))
    torch.cuda.synchronize()
    print("ok", float(r["depth"].mean()), float(r["rgb"].mean()))
    sc.close()
    w.close()


if __name__ == "__main__":
    main()
