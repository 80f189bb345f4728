"""On-disk formats either side of the hot path ("next" row N4 of SURVEY.md section 8f).

* DTU ``cameras/%08d_cam.txt``: read/write in the layout ``DtuFitSparse.read_cam_file`` parses
  (code1/dataset/dtu_test_sparse.py:184-206; README.md:67-81 of the reference): line 0 ``extrinsic``, lines 1-4 the
  4x4 world-to-camera matrix, line 6 ``intrinsic``, lines 7-9 the 3x3 K, line 11 ``DEPTH_MIN DEPTH_INTERVAL``.
* the MVSNet view-pair list ``dtu_pairs.txt`` (``MVSDataset.build_metas``, code1/dataset/dtu_train.py:171-178): first line
  the number of viewpoints, then per viewpoint one line with its id and one line ``N id score id score ...``.
* depth-map results: ``depth/<scan>/<view>.npy`` holding the pickled dict ``{"depth", "extrinsic", "intrinsic"}`` that
  ``extract_geometry`` saves (code1/model.py:839-842) and ``save_tsdf`` loads (tsdf_fusion.py:459-467; it reads
  ``refview{id}.npy`` - the producer/consumer names disagree in the reference, SURVEY F14, so the name is a parameter).
* the 8-bit previews next to them (model.py:834-836).
* the ASCII PLY files of ``save_tsdf``: ``meshwrite`` / ``pcwrite`` (tsdf_fusion.py:384-446), same header and the same
  ``%f`` / ``%d`` records, written in one vectorised pass instead of a Python loop per vertex.

Pure numpy / PIL host code; no part of the hot path.
"""
from __future__ import annotations

import os
from typing import Dict, Tuple

import numpy as np


def read_cam_file(filename: str) -> Dict[str, np.ndarray]:
    """-> extrinsic [4,4], intrinsic [3,3] (float32), depth_min, depth_interval; plus ``P = K4 @ E`` and the
    ``depth_max`` / scaled interval the reference derives (dtu_test_sparse.py:199-204)."""
    with open(filename) as f:
        lines = [line.rstrip() for line in f.readlines()]
    extr = np.array(" ".join(lines[1:5]).split(), dtype=np.float32).reshape(4, 4)
    intr = np.array(" ".join(lines[7:10]).split(), dtype=np.float32).reshape(3, 3)
    k4 = np.float32(np.diag([1, 1, 1, 1]))
    k4[:3, :3] = intr
    depth_min = float(lines[11].split()[0])
    interval = float(lines[11].split()[1])
    return {"extrinsic": extr, "intrinsic": intr, "P": k4 @ extr, "depth_min": depth_min, "depth_interval": interval,
            "depth_max": depth_min + interval * 192, "depth_interval_scaled": interval * 1.06}


def write_cam_file(filename: str, extrinsic: np.ndarray, intrinsic: np.ndarray, depth_min: float = 425.0,
                   depth_interval: float = 2.5) -> None:
    os.makedirs(os.path.dirname(os.path.abspath(filename)), exist_ok=True)
    with open(filename, "w") as f:
        f.write("extrinsic\n")
        for r in np.asarray(extrinsic, dtype=np.float64).reshape(4, 4):
            f.write(" ".join(repr(float(np.float32(x))) for x in r) + "\n")
        f.write("\nintrinsic\n")
        for r in np.asarray(intrinsic, dtype=np.float64).reshape(3, 3):
            f.write(" ".join(repr(float(np.float32(x))) for x in r) + "\n")
        f.write(f"\n{depth_min} {depth_interval}\n")


def read_pair_file(filename: str) -> Dict[int, list]:
    """``ref_src_pairs`` of ``build_metas`` (dtu_train.py:171-178): reference view -> source views, best first (the scores
    at the odd positions of the line are dropped exactly like the reference's ``split()[1::2]``)."""
    pairs: Dict[int, list] = {}
    with open(filename) as f:
        n = int(f.readline())
        for _ in range(n):
            ref = int(f.readline().rstrip())
            pairs[ref] = [int(x) for x in f.readline().rstrip().split()[1::2]]
    return pairs


def write_pair_file(filename: str, pairs: Dict[int, list], scores: Dict[int, list] = None) -> None:
    """Inverse of :func:`read_pair_file` (scores default to 0)."""
    with open(filename, "w") as f:
        f.write("%d\n" % len(pairs))
        for ref, srcs in pairs.items():
            sc = scores[ref] if scores is not None else [0.0] * len(srcs)
            f.write("%d\n%d " % (ref, len(srcs)) + " ".join("%d %f" % (s, c) for s, c in zip(srcs, sc)) + " \n")


def save_depth_result(out_dir: str, scan: str, view: str, depth_mm: np.ndarray, rgb: np.ndarray, extrinsic: np.ndarray,
                      intrinsic: np.ndarray, npy_name: str = "{view}.npy", previews: bool = True) -> str:
    """What ``extract_geometry`` writes for one rendered view (model.py:825-842).  depth_mm [H,W] float, rgb [H,W,3] in
    [0,1].  ``npy_name="refview{view}.npy"`` gives the name ``tsdf_fusion.save_tsdf`` looks for.  Returns the npy path."""
    depth_mm = np.asarray(depth_mm, dtype=np.float32)
    d = os.path.join(out_dir, "depth", scan)
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, npy_name.format(view=view))
    np.save(path, {"depth": depth_mm, "extrinsic": np.asarray(extrinsic), "intrinsic": np.asarray(intrinsic)})
    if previews:
        from PIL import Image
        rgb8 = (np.asarray(rgb, dtype=np.float32) * 255).astype(np.uint8)                       # model.py:831
        depth8 = ((depth_mm / np.max(depth_mm)).astype(np.float32) * 255).astype(np.uint8)     # model.py:834
        os.makedirs(os.path.join(out_dir, scan, "depth"), exist_ok=True)
        os.makedirs(os.path.join(out_dir, "rgb", scan), exist_ok=True)
        Image.fromarray(depth8).save(os.path.join(out_dir, scan, "depth", f"{view}.png"))
        Image.fromarray(rgb8).save(os.path.join(out_dir, "rgb", scan, f"{view}.jpg"))
    return path


def load_depth_result(path: str) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """-> (depth, intrinsic, cam_pose = inv(extrinsic)) exactly as ``save_tsdf`` unpacks the file (tsdf_fusion.py:463-467)."""
    data = np.load(path, allow_pickle=True).item()
    return data["depth"], data["intrinsic"], np.linalg.inv(data["extrinsic"])


def meshwrite(filename: str, verts: np.ndarray, faces: np.ndarray, norms: np.ndarray, colors: np.ndarray) -> None:
    """``meshwrite`` (tsdf_fusion.py:384-417): ASCII PLY with per-vertex normal and colour, ``3 i j k`` faces."""
    verts, norms = np.asarray(verts, dtype=np.float64), np.asarray(norms, dtype=np.float64)
    colors, faces = np.asarray(colors).astype(np.int64), np.asarray(faces).astype(np.int64)
    with open(filename, "w") as f:
        f.write("ply\nformat ascii 1.0\n")
        f.write("element vertex %d\n" % verts.shape[0])
        f.write("property float x\nproperty float y\nproperty float z\n")
        f.write("property float nx\nproperty float ny\nproperty float nz\n")
        f.write("property uchar red\nproperty uchar green\nproperty uchar blue\n")
        f.write("element face %d\n" % faces.shape[0])
        f.write("property list uchar int vertex_index\nend_header\n")
        if verts.shape[0]:
            np.savetxt(f, np.hstack([verts, norms, colors]), fmt=["%f"] * 6 + ["%d"] * 3)
        if faces.shape[0]:
            np.savetxt(f, np.hstack([np.full((faces.shape[0], 1), 3, dtype=np.int64), faces]), fmt="%d")


def pcwrite(filename: str, xyzrgb: np.ndarray) -> None:
    """``pcwrite`` (tsdf_fusion.py:420-446): ASCII PLY point cloud, xyz as ``%f`` and rgb as uchar."""
    xyzrgb = np.asarray(xyzrgb)
    xyz = xyzrgb[:, :3].astype(np.float64)
    rgb = xyzrgb[:, 3:].astype(np.uint8).astype(np.int64)
    with open(filename, "w") as f:
        f.write("ply\nformat ascii 1.0\n")
        f.write("element vertex %d\n" % xyz.shape[0])
        f.write("property float x\nproperty float y\nproperty float z\n")
        f.write("property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n")
        if xyz.shape[0]:
            np.savetxt(f, np.hstack([xyz, rgb]), fmt=["%f"] * 3 + ["%d"] * 3)
