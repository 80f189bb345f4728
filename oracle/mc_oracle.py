"""TEST INFRASTRUCTURE - CPU oracle for the iso-surface extraction that follows TSDF fusion ("next" row N3).

The reference calls ``skimage.measure.marching_cubes_lewiner(tsdf_vol, level=0)`` (tsdf_fusion.py:325,345): a
THIRD-PARTY dependency (scikit-image, unpinned in the reference's requirements; absent from /root/reference and from
this image), so **parity with the reference's mesh is unpinned**.  What is restated here is the published algorithm
(Lorensen & Cline marching cubes on the cell grid, one vertex per sign-changing grid edge placed by linear
interpolation - Lewiner's variant places its vertices identically up to FLT_EPSILON and differs only in how it
triangulates topologically ambiguous cells), with a case table *generated* below rather than copied:

* corner c = dx + 2 dy + 4 dz of a cell is "inside" when f < level; edge e = 4*axis + (b0 + 2 b1) joins the corners
  that differ along `axis` (b0, b1 = the other two coordinates in increasing axis order);
* on every cell face the crossing edges are joined pairwise; a face with four crossings (inside corners on a diagonal)
  is resolved by cutting off each inside corner - a rule that only looks at the face's own signs, so the two cells
  sharing a face agree and the surface is closed;
* the segments close into loops, each loop is oriented so that its normal points towards larger f (outside, where
  the TSDF is positive) and fan-triangulated from its smallest edge.

Only tests/, __graft_entry__.smoke() and tools/gen_mc_table.py (which prints the table as a CUDA header at
development time) may import this module; the product reads the committed header.
"""
from __future__ import annotations

from functools import lru_cache
from typing import Dict, List, Tuple

import numpy as np

AXES = np.eye(3, dtype=np.int64)


def corner_xyz(c: int) -> Tuple[int, int, int]:
    return (c & 1, (c >> 1) & 1, (c >> 2) & 1)


def edge_corners(e: int) -> Tuple[int, int]:
    """(lower corner, upper corner) of edge e."""
    a, b = divmod(e, 4)
    others = [i for i in range(3) if i != a]
    xyz = [0, 0, 0]
    xyz[others[0]], xyz[others[1]] = b & 1, (b >> 1) & 1
    c0 = xyz[0] + 2 * xyz[1] + 4 * xyz[2]
    return c0, c0 + (1 << a)


def _face_edges(a: int, v: int) -> List[int]:
    """the four edges lying in the face {coordinate a == v}"""
    out = []
    for e in range(12):
        ax = e // 4
        if ax == a:
            continue
        if corner_xyz(edge_corners(e)[0])[a] == v:
            out.append(e)
    return out


def _edge_faces(e: int):
    p0, p1 = corner_xyz(edge_corners(e)[0]), corner_xyz(edge_corners(e)[1])
    return {(a, p0[a]) for a in range(3) if p0[a] == p1[a]}


def _triangulate(loop: List[int]) -> List[Tuple[int, int, int]]:
    """First triangulation (fan-like recursion, lexicographic in the split index) of the oriented loop none of whose
    diagonals lies in a cube face: a diagonal inside a face would coincide with the face's own segments or cross the
    neighbour cell's, making the surface non-manifold there."""
    n = len(loop)

    def ok(lp, i, j):                                        # is (i,j) a polygon side, or a diagonal off every face?
        if (j - i) % n in (1, n - 1):
            return True
        return not (_edge_faces(lp[i]) & _edge_faces(lp[j]))

    def rec(lp, i, j):                                       # first valid triangulation of the sub-polygon i..j, or None
        if j - i < 2:
            return []
        for k in range(i + 1, j):
            if ok(lp, i, k) and ok(lp, k, j):
                a, b = rec(lp, i, k), rec(lp, k, j)
                if a is not None and b is not None:
                    return a + [(lp[i], lp[k], lp[j])] + b
        return None

    for rot in range(n):                                     # (0, n-1) is a side for every rotation
        lp = loop[rot:] + loop[:rot]
        r = rec(lp, 0, n - 1)
        if r is not None:
            return r
    raise AssertionError(("no face-free triangulation", loop))


@lru_cache(maxsize=None)
def build_tables() -> Dict[str, np.ndarray]:
    """ntri [256] uint8, tri [256, 3*max_tri] int8 (edge ids, -1 padded), edge_owner [12,4] = (dx,dy,dz,axis)."""
    mid = {e: (np.array(corner_xyz(edge_corners(e)[0]), float) + np.array(corner_xyz(edge_corners(e)[1]), float)) / 2 for e in range(12)}
    cases: List[List[Tuple[int, int, int]]] = []
    for case in range(256):
        inside = [(case >> c) & 1 for c in range(8)]
        cross = [e for e in range(12) if inside[edge_corners(e)[0]] != inside[edge_corners(e)[1]]]
        link: Dict[int, List[int]] = {e: [] for e in cross}
        for a in range(3):
            for v in (0, 1):
                fe = [e for e in _face_edges(a, v) if e in link]
                if len(fe) == 2:
                    link[fe[0]].append(fe[1]); link[fe[1]].append(fe[0])
                elif len(fe) == 4:
                    for c in range(8):                       # cut off each inside corner of the face
                        if corner_xyz(c)[a] == v and inside[c]:
                            pair = [e for e in fe if c in edge_corners(e)]
                            link[pair[0]].append(pair[1]); link[pair[1]].append(pair[0])
                else:
                    assert len(fe) == 0
        assert all(len(v) == 2 for v in link.values())
        tris: List[Tuple[int, int, int]] = []
        seen = set()
        for start in cross:                                   # ascending edge id
            if start in seen:
                continue
            loop, prev, cur = [start], None, start
            seen.add(start)
            while True:
                a, b = link[cur]
                nx = min(a, b) if prev is None else (a if a != prev else b)
                if nx == start:
                    break
                loop.append(nx)
                seen.add(nx)
                prev, cur = cur, nx
            assert len(loop) >= 3
            pts = [mid[e] for e in loop]
            nrm = sum(np.cross(pts[i], pts[(i + 1) % len(pts)]) for i in range(len(pts)))
            d = np.zeros(3)
            for e in loop:
                c0, c1 = edge_corners(e)
                p0, p1 = np.array(corner_xyz(c0), float), np.array(corner_xyz(c1), float)
                d += (p0 - p1) if inside[c1] else (p1 - p0)   # from the inside end to the outside end
            s = float(np.dot(nrm, d))
            assert abs(s) > 1e-9, (case, loop)
            if s < 0:
                loop = [loop[0]] + loop[:0:-1]
            tris.extend(_triangulate(loop))
        cases.append(tris)
    max_tri = max(len(t) for t in cases)
    ntri = np.array([len(t) for t in cases], dtype=np.uint8)
    tri = -np.ones((256, 3 * max_tri), dtype=np.int8)
    for k, t in enumerate(cases):
        flat = [e for tr in t for e in tr]
        tri[k, :len(flat)] = flat
    owner = np.array([list(corner_xyz(edge_corners(e)[0])) + [e // 4] for e in range(12)], dtype=np.int32)
    return {"ntri": ntri, "tri": tri, "edge_owner": owner, "max_tri": np.int32(max_tri)}


def gradient(f: np.ndarray) -> np.ndarray:
    """[X,Y,Z,3] fp32: central differences (f[i+1]-f[i-1])*0.5, one-sided f[1]-f[0] / f[-1]-f[-2] at the borders."""
    g = np.zeros(f.shape + (3,), dtype=np.float32)
    for a in range(3):
        fa = np.moveaxis(f, a, 0)
        ga = np.moveaxis(g[..., a], a, 0)
        if fa.shape[0] > 2:
            ga[1:-1] = (fa[2:] - fa[:-2]) * np.float32(0.5)
        if fa.shape[0] > 1:
            ga[0] = fa[1] - fa[0]
            ga[-1] = fa[-1] - fa[-2]
    return g


def marching_cubes(vol: np.ndarray, level: float = 0.0):
    """verts [Nv,3] fp32 (voxel coordinates), faces [Nf,3] int32, normals [Nv,3] fp32 (unit, towards larger f).

    Vertex order: owner voxel in C order (x slowest, z fastest), then edge axis.  Face order: cell in C order, then the
    case table's triangle order.  All arithmetic fp32, one rounding per operation."""
    T = build_tables()
    f = np.ascontiguousarray(vol, dtype=np.float32)
    X, Y, Z = f.shape
    lv = np.float32(level)
    inside = f < lv
    cross = np.zeros((3, X, Y, Z), dtype=bool)
    cross[0, :-1] = inside[:-1] != inside[1:]
    cross[1, :, :-1] = inside[:, :-1] != inside[:, 1:]
    cross[2, :, :, :-1] = inside[:, :, :-1] != inside[:, :, 1:]
    cnt = cross.sum(0).astype(np.int64)
    base = np.cumsum(cnt.ravel()) - cnt.ravel()
    base = base.reshape(X, Y, Z)
    vid = np.stack([base, base + cross[0], base + cross[0].astype(np.int64) + cross[1]], 0)
    nv = int(cnt.sum())
    verts = np.zeros((nv, 3), dtype=np.float32)
    normals = np.zeros((nv, 3), dtype=np.float32)
    g = gradient(f)
    for a in range(3):
        idx = np.argwhere(cross[a])
        if len(idx) == 0:
            continue
        i0 = tuple(idx.T)
        i1 = tuple((idx + AXES[a]).T)
        f0, f1 = f[i0], f[i1]
        t = ((lv - f0) / (f1 - f0)).astype(np.float32)
        p = idx.astype(np.float32)
        p[:, a] = p[:, a] + t
        n = g[i0] + t[:, None] * (g[i1] - g[i0])
        ln = np.sqrt(n[:, 0] * n[:, 0] + n[:, 1] * n[:, 1] + n[:, 2] * n[:, 2]).astype(np.float32)
        n = n / np.maximum(ln, np.float32(1e-20))[:, None]
        k = vid[a][i0]
        verts[k] = p
        normals[k] = n.astype(np.float32)
    if min(X, Y, Z) < 2:
        return verts, np.zeros((0, 3), np.int32), normals
    case = np.zeros((X - 1, Y - 1, Z - 1), dtype=np.int64)
    for c in range(8):
        dx, dy, dz = corner_xyz(c)
        case |= inside[dx:X - 1 + dx, dy:Y - 1 + dy, dz:Z - 1 + dz].astype(np.int64) << c
    ntri = T["ntri"][case]
    cells = np.argwhere(ntri > 0)                            # C order
    faces = []
    own = T["edge_owner"]
    for cell in cells:
        k = case[tuple(cell)]
        for j in range(int(T["ntri"][k])):
            tri = []
            for e in T["tri"][k, 3 * j:3 * j + 3]:
                o = cell + own[e, :3]
                tri.append(vid[own[e, 3]][o[0], o[1], o[2]])
            faces.append(tri)
    return verts, np.array(faces, dtype=np.int32).reshape(-1, 3), normals


def mesh_is_closed(faces: np.ndarray, verts: np.ndarray, shape) -> bool:
    """every directed edge (a,b) whose vertices are not both on the volume boundary is matched by exactly one (b,a)."""
    from collections import Counter
    d = Counter()
    for a, b, c in faces:
        for u, v in ((a, b), (b, c), (c, a)):
            d[(int(u), int(v))] += 1
    hi = np.array(shape, dtype=np.float32) - 1
    on_b = ((verts <= 0) | (verts >= hi)).any(1)
    for (u, v), n in d.items():
        if n != 1:
            return False
        if d.get((v, u), 0) != 1 and not (on_b[u] and on_b[v]):
            return False
    return True
