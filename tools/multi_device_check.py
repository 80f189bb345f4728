"""One process, two devices: handles are per device and kernel attributes must be set on each (run on a >= 2-GPU box)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from conftest import make_case  # noqa: E402
from uforecon_b200 import synthetic  # noqa: E402
from uforecon_b200.renderer import HotPathWeights, Scene, render_rays  # noqa: E402

batch, scene, sd = make_case(synthetic.UNFAVORABLE_VIEWS, (96, 64))
n = 300
u_c, u_f = synthetic.sampler_uniforms(n, seed=4)
res = []
for d in range(min(2, torch.cuda.device_count())):
    dev = torch.device("cuda", d)
    with torch.cuda.device(dev):
        w = HotPathWeights(sd, dev)
        sc = Scene(batch, scene["source_imgs_feat"], scene["feature_volume"], scene["match_feature"], dev)
        out = {}
        for mode in (0, 1):
            r = render_rays(sc, w, None, n, u_c, u_f, mode, ray_begin=100)
            torch.cuda.synchronize(dev)
            out[mode] = (r["depth"].cpu(), r["rgb"].cpu())
        res.append(out)
        sc.close()
        w.close()
if len(res) == 2:
    for mode in (0, 1):
        assert torch.equal(res[0][mode][0], res[1][mode][0]) and torch.equal(res[0][mode][1], res[1][mode][1]), mode
    print("multi-device OK: identical results on cuda:0 and cuda:1 in one process")
else:
    print("single device only")
