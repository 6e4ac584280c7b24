"""Mesh files of the reference pipeline, without trimesh.

The reference reads its quadrature mesh with `trimesh.load(path, force='mesh', process=False)`
(mesh_utils.py:193) and, for the baked path, the UV-unwrapped mesh with `trimesh.load(path, process=False)`
followed by `mesh.visual.uv` (test_baking_texture_images.py:323-328).  The marching-cubes stage exports binary
little-endian PLY (trimesh's default exporter), the xatlas stage an OBJ with one `vt` per vertex.  This module reads
and writes those two formats with numpy only and hands back the attributes the hot path touches:
`vertices` (V,3) f64, `faces` (F,3) i64, `face_normals` (F,3) f64, `triangles` (F,3,3) f64, `visual.uv` (V,2) f64.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np


class _Visual:
    """`mesh.visual` with the one attribute the reference reads (`uv`)."""

    def __init__(self, uv: Optional[np.ndarray] = None):
        self.uv = None if uv is None else np.asarray(uv, dtype=np.float64)


class Mesh:
    """Stand-in for the `trimesh.Trimesh` attributes the reference touches."""

    def __init__(self, vertices, faces, uv=None):
        self.vertices = np.asarray(vertices, dtype=np.float64).reshape(-1, 3)
        self.faces = np.asarray(faces, dtype=np.int64).reshape(-1, 3)
        self.visual = _Visual(uv)
        if self.faces.size and (self.faces.min() < 0 or self.faces.max() >= len(self.vertices)):
            raise ValueError("mesh faces index outside the vertex array")

    @property
    def triangles(self) -> np.ndarray:
        return self.vertices[self.faces]

    @property
    def face_normals(self) -> np.ndarray:
        """Unit normals cross(v1-v0, v2-v0) in fp64; zero for degenerate faces (trimesh pads those with zeros)."""
        t = self.triangles
        n = np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0])
        ln = np.linalg.norm(n, axis=1)
        ok = ln > np.finfo(np.float64).eps
        out = np.zeros_like(n)
        out[ok] = n[ok] / ln[ok, None]
        return out

    @property
    def scale(self) -> float:
        """Length of the bounding-box diagonal (trimesh `Trimesh.scale`; the Embree restart epsilon derives from it)."""
        if not len(self.vertices):
            return 1.0
        return float(np.linalg.norm(self.vertices.max(0) - self.vertices.min(0)))


_PLY_TYPES = {
    "char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
    "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8",
}


def _triangulate(polys: List[np.ndarray]) -> np.ndarray:
    """Fan triangulation, the order trimesh uses for quads / polygons."""
    out = []
    for p in polys:
        for k in range(1, len(p) - 1):
            out.append((p[0], p[k], p[k + 1]))
    return np.asarray(out, dtype=np.int64).reshape(-1, 3)


def _parse_ply_header(f):
    magic = f.readline().strip()
    if magic != b"ply":
        raise ValueError("not a PLY file")
    fmt, elements = None, []
    while True:
        line = f.readline()
        if not line:
            raise ValueError("PLY header is not terminated")
        p = line.decode("ascii", "replace").split()
        if not p or p[0] in ("comment", "obj_info"):
            continue
        if p[0] == "format":
            fmt = p[1]
        elif p[0] == "element":
            elements.append({"name": p[1], "count": int(p[2]), "props": []})
        elif p[0] == "property":
            if p[1] == "list":
                elements[-1]["props"].append(("list", _PLY_TYPES[p[2]], _PLY_TYPES[p[3]], p[4]))
            else:
                elements[-1]["props"].append(("scalar", _PLY_TYPES[p[1]], None, p[2]))
        elif p[0] == "end_header":
            break
    if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
        raise ValueError(f"unknown PLY format {fmt!r}")
    return fmt, elements


def _read_ply_element_binary(f, el, endian):
    """-> dict name -> array (scalars) or list of arrays (lists)."""
    n, props = el["count"], el["props"]
    if all(k == "scalar" for k, *_ in props):
        dt = np.dtype([(name, endian + t) for _, t, _, name in props])
        return np.frombuffer(f.read(dt.itemsize * n), dtype=dt, count=n)
    if len(props) == 1 and n > 0:
        # one list property (the usual face element): try the fixed-width fast path when every face has the same arity
        _, ct, it, name = props[0]
        pos = f.tell()
        first = int(np.frombuffer(f.read(np.dtype(ct).itemsize), dtype=endian + ct, count=1)[0])
        f.seek(pos)
        dt = np.dtype([("n", endian + ct), ("i", endian + it, (first,))])
        raw = f.read(dt.itemsize * n)
        if len(raw) == dt.itemsize * n:
            rec = np.frombuffer(raw, dtype=dt, count=n)
            if np.all(rec["n"] == first):
                return {name: rec["i"]}
        f.seek(pos)
    cols = {name: [] for *_, name in props}
    for _ in range(n):
        for kind, t, it, name in props:
            if kind == "scalar":
                cols[name].append(np.frombuffer(f.read(np.dtype(t).itemsize), dtype=endian + t, count=1)[0])
            else:
                c = int(np.frombuffer(f.read(np.dtype(t).itemsize), dtype=endian + t, count=1)[0])
                cols[name].append(np.frombuffer(f.read(np.dtype(it).itemsize * c), dtype=endian + it, count=c))
    return cols


def _read_ply_element_ascii(f, el):
    cols = {name: [] for *_, name in el["props"]}
    for _ in range(el["count"]):
        tok = f.readline().split()
        j = 0
        for kind, t, it, name in el["props"]:
            if kind == "scalar":
                cols[name].append(float(tok[j])); j += 1
            else:
                c = int(tok[j]); j += 1
                cols[name].append(np.asarray([int(float(x)) for x in tok[j:j + c]], dtype=np.int64)); j += c
    return cols


def load_ply(path: str) -> Mesh:
    """ASCII, binary little-endian and binary big-endian PLY; extra vertex properties (normals, colours) are skipped,
    per-vertex `s`/`t` (or `u`/`v`, `texture_u`/`texture_v`) become `visual.uv`."""
    with open(path, "rb") as f:
        fmt, elements = _parse_ply_header(f)
        data = {}
        for el in elements:
            if fmt == "ascii":
                data[el["name"]] = _read_ply_element_ascii(f, el)
            else:
                data[el["name"]] = _read_ply_element_binary(f, el, "<" if fmt == "binary_little_endian" else ">")
    if "vertex" not in data:
        raise ValueError("PLY file has no vertex element")
    v = data["vertex"]
    names = v.dtype.names if isinstance(v, np.ndarray) else tuple(v.keys())
    col = lambda n: np.asarray(v[n], dtype=np.float64)
    vertices = np.stack([col("x"), col("y"), col("z")], axis=1) if len(col("x")) else np.zeros((0, 3))
    uv = None
    for a, b in (("s", "t"), ("u", "v"), ("texture_u", "texture_v")):
        if a in names and b in names:
            uv = np.stack([col(a), col(b)], axis=1)
            break
    faces = np.zeros((0, 3), dtype=np.int64)
    if "face" in data:
        fe = data["face"]
        key = next((k for k in ("vertex_indices", "vertex_index") if k in (fe.keys() if isinstance(fe, dict) else ())), None)
        if key is not None:
            idx = fe[key]
            if isinstance(idx, np.ndarray) and idx.ndim == 2 and idx.shape[1] == 3:
                faces = idx.astype(np.int64)
            else:
                faces = _triangulate([np.asarray(p, dtype=np.int64) for p in idx])
    return Mesh(vertices, faces, uv)


def load_obj(path: str) -> Mesh:
    """Wavefront OBJ: `v`, `vt`, `f` with `v`, `v/vt`, `v//vn` or `v/vt/vn` corners, negative (relative) indices,
    polygons fan-triangulated.  When every corner uses the same index for position and texture coordinate (what the
    xatlas stage writes) the vertex array is kept as is and `visual.uv[i]` belongs to vertex i; otherwise vertices are
    split per distinct (v, vt) pair in order of first appearance, like trimesh's un-merge with `process=False`."""
    verts, uvs, corners, sizes = [], [], [], []
    with open(path) as f:
        for line in f:
            p = line.split()
            if not p:
                continue
            if p[0] == "v":
                verts.append((float(p[1]), float(p[2]), float(p[3])))
            elif p[0] == "vt":
                uvs.append((float(p[1]), float(p[2]) if len(p) > 2 else 0.0))
            elif p[0] == "f":
                poly = []
                for q in p[1:]:
                    s = q.split("/")
                    vi = int(s[0])
                    ti = int(s[1]) if len(s) > 1 and s[1] else 0
                    vi = vi - 1 if vi > 0 else len(verts) + vi
                    ti = (ti - 1 if ti > 0 else len(uvs) + ti) if ti != 0 else -1
                    poly.append((vi, ti))
                corners.extend(poly)
                sizes.append(len(poly))
    vertices = np.asarray(verts, dtype=np.float64).reshape(-1, 3)
    c = np.asarray(corners, dtype=np.int64).reshape(-1, 2)
    starts = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    polys = [np.arange(starts[i], starts[i + 1]) for i in range(len(sizes))]
    tri_corner = _triangulate(polys)                                          # (F,3) indices into the corner list
    if not uvs or not len(c) or np.all(c[:, 1] < 0):
        return Mesh(vertices, c[:, 0][tri_corner] if len(c) else np.zeros((0, 3), np.int64))
    uv_all = np.asarray(uvs, dtype=np.float64)
    if np.all(c[:, 0] == c[:, 1]) and len(uv_all) == len(vertices):
        return Mesh(vertices, c[:, 0][tri_corner], uv_all)
    # split: one vertex per distinct (v, vt) pair, first appearance first
    key = c[:, 0] * (len(uv_all) + 1) + (c[:, 1] + 1)
    _, first, inverse = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    new_of_corner = rank[inverse]
    pair = c[first[order]]
    uv = np.where(pair[:, 1:2] >= 0, uv_all[np.clip(pair[:, 1], 0, None)], 0.0)
    return Mesh(vertices[pair[:, 0]], new_of_corner[tri_corner], uv)


def load_mesh(path: str) -> Mesh:
    """`trimesh.load(path, force='mesh', process=False)` for the two formats the pipeline writes."""
    low = path.lower()
    if low.endswith(".obj"):
        return load_obj(path)
    if low.endswith(".ply"):
        return load_ply(path)
    raise NotImplementedError(f"unsupported mesh format: {path}")


def save_ply(path: str, vertices, faces, binary: bool = True) -> None:
    """Binary little-endian (trimesh's default export, `mesh.export('x.ply')`) or ASCII PLY, float32 vertices."""
    v = np.asarray(vertices, dtype="<f4").reshape(-1, 3)
    f = np.asarray(faces, dtype="<i4").reshape(-1, 3)
    header = ("ply\nformat {} 1.0\nelement vertex {}\nproperty float x\nproperty float y\nproperty float z\n"
              "element face {}\nproperty list uchar int vertex_indices\nend_header\n").format(
                  "binary_little_endian" if binary else "ascii", len(v), len(f))
    with open(path, "wb") as out:
        out.write(header.encode("ascii"))
        if binary:
            out.write(v.tobytes())
            rec = np.empty(len(f), dtype=[("n", "u1"), ("i", "<i4", (3,))])
            rec["n"], rec["i"] = 3, f
            out.write(rec.tobytes())
        else:
            for r in v:
                out.write("{:.9g} {:.9g} {:.9g}\n".format(*r).encode("ascii"))
            for r in f:
                out.write("3 {} {} {}\n".format(*r).encode("ascii"))


def save_obj(path: str, vertices, faces, uv=None) -> None:
    """OBJ with one `vt` per vertex when `uv` is given (`f a/a b/b c/c`), the layout of the xatlas stage."""
    v = np.asarray(vertices, dtype=np.float64).reshape(-1, 3)
    f = np.asarray(faces, dtype=np.int64).reshape(-1, 3) + 1
    with open(path, "w") as out:
        for r in v:
            out.write("v {:.17g} {:.17g} {:.17g}\n".format(*r))
        if uv is not None:
            for r in np.asarray(uv, dtype=np.float64).reshape(-1, 2):
                out.write("vt {:.17g} {:.17g}\n".format(*r))
            for r in f:
                out.write("f {0}/{0} {1}/{1} {2}/{2}\n".format(*r))
        else:
            for r in f:
                out.write("f {} {} {}\n".format(*r))
