"""Generate tests/golden/ply_case.npz: the bytes the UNMODIFIED reference's ``meshwrite`` / ``pcwrite``
(tsdf_fusion.py:384-446) produce for a small seeded mesh.  Build-container only (imports /root/reference with a
stub for the absent scikit-image, which these two functions do not use)."""
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("UFO_REFERENCE_ROOT", "/root/reference")


def ply_case_inputs(seed=0, nv=37, nf=51):
    rng = np.random.default_rng(seed)
    verts = (rng.standard_normal((nv, 3)) * 123.456).astype(np.float32)
    verts[0] = [0.0, -0.0000004, 1e6]                                      # %f rounding / sign / width cases
    norms = rng.standard_normal((nv, 3)).astype(np.float32)
    norms /= np.linalg.norm(norms, axis=1, keepdims=True)
    colors = rng.integers(0, 256, (nv, 3)).astype(np.uint8)
    faces = rng.integers(0, nv, (nf, 3)).astype(np.int32)
    return verts, faces, norms, colors


def main():
    sys.modules.setdefault("skimage", types.ModuleType("skimage"))
    sk_measure = types.ModuleType("skimage.measure")
    sys.modules["skimage"].measure = sk_measure
    sys.modules["skimage.measure"] = sk_measure
    sys.path.insert(0, REF)
    import tsdf_fusion as ref                                              # the reference module

    verts, faces, norms, colors = ply_case_inputs()
    d = tempfile.mkdtemp()
    ref.meshwrite(os.path.join(d, "m.ply"), verts, faces, norms, colors)
    ref.pcwrite(os.path.join(d, "p.ply"), np.hstack([verts, colors.astype(np.float32)]))
    out = os.path.join(ROOT, "tests", "golden", "ply_case.npz")
    np.savez_compressed(out, mesh=np.frombuffer(open(os.path.join(d, "m.ply"), "rb").read(), dtype=np.uint8),
                        cloud=np.frombuffer(open(os.path.join(d, "p.ply"), "rb").read(), dtype=np.uint8))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
