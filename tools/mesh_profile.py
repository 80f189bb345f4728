"""Profiling aid: marching cubes of a 512^3 bumpy-sphere distance field on cuda:0 (run under ncu -k regex:k_mc)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uforecon_b200 import _lib  # noqa: E402
from uforecon_b200.tsdf import marching_cubes  # noqa: E402


def bumpy_sphere(n=512, device="cuda"):
    ax = torch.arange(n, device=device, dtype=torch.float32) - (n - 1) / 2
    x, y, z = ax[:, None, None], ax[None, :, None], ax[None, None, :]
    r = torch.sqrt(x * x + y * y + z * z)
    bumps = 6.0 * torch.sin(x * 0.11) * torch.sin(y * 0.13) * torch.sin(z * 0.09)
    return ((r - 0.37 * n + bumps) / 3.0).clamp(-1, 1).contiguous()          # truncated like a TSDF


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    vol = bumpy_sphere(n)
    for it in range(3):
        _lib.profile_begin()
        t0 = time.time()
        v, f, nrm = marching_cubes(vol)
        torch.cuda.synchronize()
        dt = time.time() - t0
        prof = _lib.profile_end(16)
        print(f"{n}^3: {v.shape[0]} verts {f.shape[0]} faces, wall {dt * 1e3:.2f} ms;", ", ".join(f"{a} {m:.3f} ms" for a, c, m in prof))
