"""TEST INFRASTRUCTURE ONLY — CPU oracle (numpy / PyTorch fp32) for the Quadfield render hot path.

Standalone restatement of the reference algorithm, row by row of SURVEY.md §8(a).  It never
touches ``/root/reference`` at run time (that tree does not exist on the GPU box); it is pinned
against the reference by ``tests/test_oracle_golden.py`` using fixtures that
``oracle/make_golden.py`` produced from the unmodified reference modules.

All ``file:line`` citations are into ``/root/reference/examples/``.

Numeric conventions shared with the CUDA kernels:
  * geometry is fp32 with one rounding per operation, no fused multiply-add, in the operand
    order spelled out below (numpy evaluates ``a*b-c*d`` as three separately rounded ops;
    the kernel uses ``__fmul_rn/__fadd_rn``), so hit ids / counts are bit-exact;
  * tcnn module boundaries round to fp16 (hash-grid output, SH output, head input) because
    tinycudann's encodings emit ``__half``; the hash-grid interpolation accumulates in half like
    tcnn's ``kernel_grid`` (one half FMA per corner); MLP arithmetic itself is restated in fp32.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

F32 = np.float32

# --------------------------------------------------------------------------------------
# a1  ray generation — datasets/nerf_synthetic.py:219-226, 289-378
# --------------------------------------------------------------------------------------


def pinhole_intrinsics(width: int, height: int, camera_angle_x: float, upsample: int = 1):
    """K as built at nerf_synthetic.py:101-102 (focal) and :214-226 (upsample)."""
    focal = 0.5 * width / math.tan(0.5 * camera_angle_x)
    focal = focal * upsample
    W, H = int(width * upsample), int(height * upsample)
    return F32(focal), F32(W / 2.0), F32(H / 2.0), W, H


def generate_rays(c2w: np.ndarray, W: int, H: int, focal, cx, cy, opengl: bool = True):
    """Eval-mode rays of one camera: nerf_synthetic.py:310-317 (row-major ``meshgrid(indexing="xy")``),
    :341-360 (camera dirs, rotate, normalise).  Returns origins (N,3), viewdirs (N,3) fp32."""
    c2w = torch.as_tensor(np.asarray(c2w, dtype=np.float32))
    x, y = torch.meshgrid(torch.arange(W), torch.arange(H), indexing="xy")
    x = x.flatten()
    y = y.flatten()
    K00 = torch.tensor(focal, dtype=torch.float32)
    cxt = torch.tensor(cx, dtype=torch.float32)
    cyt = torch.tensor(cy, dtype=torch.float32)
    sgn = -1.0 if opengl else 1.0
    camera_dirs = torch.nn.functional.pad(
        torch.stack([(x - cxt + 0.5) / K00, (y - cyt + 0.5) / K00 * sgn], dim=-1), (0, 1), value=sgn
    )
    directions = (camera_dirs[:, None, :] * c2w[None, :3, :3]).sum(dim=-1)
    origins = torch.broadcast_to(c2w[:3, -1], directions.shape)
    viewdirs = directions / torch.linalg.norm(directions, dim=-1, keepdims=True)
    return origins.contiguous().numpy().copy(), viewdirs.contiguous().numpy().copy()


def generate_rays_indexed(camtoworlds, image_id, x, y, focal, cx, cy, opengl: bool = True):
    """Training-branch rays: nerf_synthetic.py:341-370 with one camera matrix per ray (`c2w = camtoworlds[image_id]`,
    :337).  x, y: pixel indices (int) or index + noise (float).  -> origins (n,3), viewdirs (n,3) fp32."""
    c2w = torch.as_tensor(np.asarray(camtoworlds, dtype=np.float32))[torch.as_tensor(np.asarray(image_id, dtype=np.int64))]
    x, y = torch.as_tensor(np.asarray(x)), torch.as_tensor(np.asarray(y))
    K00, cxt, cyt = (torch.tensor(v, dtype=torch.float32) for v in (focal, cx, cy))
    sgn = -1.0 if opengl else 1.0
    camera_dirs = torch.nn.functional.pad(
        torch.stack([(x - cxt + 0.5) / K00, (y - cyt + 0.5) / K00 * sgn], dim=-1), (0, 1), value=sgn
    )
    directions = (camera_dirs[:, None, :] * c2w[:, :3, :3]).sum(dim=-1)
    origins = torch.broadcast_to(c2w[:, :3, -1], directions.shape)
    viewdirs = directions / torch.linalg.norm(directions, dim=-1, keepdims=True)
    return origins.contiguous().numpy().copy(), viewdirs.contiguous().numpy().copy()


def subject_pixels(images, image_id, x, y, upsample: int, color_bkgd):
    """Target colours of a batch: nerf_synthetic.py:331 (rgba at floor(y/upsample), floor(x/upsample)) and :266-281
    (alpha-composite over the background colour).  images (n,h,w,4) uint8 -> (n_rays,3) fp32."""
    img = torch.as_tensor(np.asarray(images))
    iid = torch.as_tensor(np.asarray(image_id, dtype=np.int64))
    xi, yi = torch.as_tensor(np.asarray(x)), torch.as_tensor(np.asarray(y))
    rgba = img[iid, torch.floor(yi / upsample).long(), torch.floor(xi / upsample).long()] / 255.0
    pixels, alpha = torch.split(rgba, [3, 1], dim=-1)
    bk = torch.as_tensor(np.asarray(color_bkgd, dtype=np.float32))
    return (pixels * alpha + bk * (1.0 - alpha)).numpy()


def look_at_c2w(eye, target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0)) -> np.ndarray:
    """OpenGL-convention camera-to-world (camera looks down −z), used only to make synthetic poses."""
    eye = np.asarray(eye, dtype=np.float64)
    f = np.asarray(target, dtype=np.float64) - eye
    f /= np.linalg.norm(f)
    r = np.cross(f, np.asarray(up, dtype=np.float64))
    r /= np.linalg.norm(r)
    u = np.cross(r, f)
    m = np.zeros((3, 4), dtype=np.float64)
    m[:, 0], m[:, 1], m[:, 2], m[:, 3] = r, u, -f, eye
    return m.astype(np.float32)


# --------------------------------------------------------------------------------------
# synthetic scene: jittered concentric icospheres
# --------------------------------------------------------------------------------------


def icosphere(subdivisions: int) -> Tuple[np.ndarray, np.ndarray]:
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array(
        [[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
         [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    f = np.array(
        [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
         [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
         [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    for _ in range(subdivisions):
        edges = {}
        verts = list(v)
        new_f = []

        def mid(a, b):
            key = (a, b) if a < b else (b, a)
            if key not in edges:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                edges[key] = len(verts) - 1
            return edges[key]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            new_f += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        v = np.array(verts)
        f = np.array(new_f, dtype=np.int64)
    return v, f


def shell_mesh(radii: Sequence[float], subdivisions: int, jitter: float = 1e-3, seed: int = 42):
    """Concentric jittered icospheres → (vertices f32 (V,3), faces i32 (F,3))."""
    rng = np.random.RandomState(seed)
    v0, f0 = icosphere(subdivisions)
    vs, fs, base = [], [], 0
    for r in radii:
        vs.append(v0 * r + rng.normal(0.0, jitter, size=v0.shape))
        fs.append(f0 + base)
        base += v0.shape[0]
    return np.concatenate(vs).astype(np.float32), np.concatenate(fs).astype(np.int32)


# --------------------------------------------------------------------------------------
# a2  ray–mesh intersection, first K hits by t
#     replaces trimesh/Embree `intersects_id(..., multiple_hits=True, max_hits=K)`
#     (mesh_utils.py:223,350-354) and the OptiX `Intersector.find_intersections`
#     (mesh_utils.py:77-96).  PARITY UNPINNED: neither native library is available; the
#     predicate below is the definition both the oracle and the kernel implement.
# --------------------------------------------------------------------------------------

BOX_PAD_REL = F32(1e-4)


def mesh_box_pad(vertices: np.ndarray) -> np.float32:
    """pad = 1e-4 · (largest scene extent), one fp32 multiply."""
    v = np.asarray(vertices, dtype=np.float32)
    ext = (v.max(axis=0) - v.min(axis=0)).max()
    return F32(F32(ext) * BOX_PAD_REL)


def triangle_arrays(vertices: np.ndarray, faces: np.ndarray):
    v = np.asarray(vertices, dtype=np.float32)[np.asarray(faces, dtype=np.int64)]
    v0, v1, v2 = v[:, 0], v[:, 1], v[:, 2]
    pad = mesh_box_pad(vertices)
    lo = np.minimum(np.minimum(v0, v1), v2) - pad
    hi = np.maximum(np.maximum(v0, v1), v2) + pad
    return v0, v1 - v0, v2 - v0, lo, hi


def face_normals(vertices: np.ndarray, faces: np.ndarray) -> np.ndarray:
    """trimesh 3.23.5 ``Trimesh.face_normals`` restated (absent dependency, requirements.txt:5):
    fp64 cross(v1−v0, v2−v1), divided by its length, zero for degenerate faces; the reference
    casts it to fp32 at mesh_utils.py:104."""
    v = np.asarray(vertices, dtype=np.float32).astype(np.float64)[np.asarray(faces, dtype=np.int64)]
    a = v[:, 1] - v[:, 0]
    b = v[:, 2] - v[:, 1]
    n = np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1],
                  a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                  a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], axis=1)
    ln = np.sqrt((n[:, 0] * n[:, 0] + n[:, 1] * n[:, 1]) + n[:, 2] * n[:, 2])
    ok = ln > 0.0
    out = np.zeros_like(n)
    out[ok] = n[ok] / ln[ok, None]
    return out.astype(np.float32)


def _ray_tri_block(o, d, v0, e1, e2, lo, hi):
    """All pairs of R rays × F triangles. o,d: (R,1,3); triangle arrays (1,F,3). Returns hit (R,F) bool, t (R,F) f32.

    Möller–Trumbore in fp32, op order fixed; plus the ray-vs-padded-triangle-box slab test that
    makes BVH culling exact by monotonicity."""
    ox, oy, oz = o[..., 0], o[..., 1], o[..., 2]
    dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]
    e1x, e1y, e1z = e1[..., 0], e1[..., 1], e1[..., 2]
    e2x, e2y, e2z = e2[..., 0], e2[..., 1], e2[..., 2]
    px = dy * e2z - dz * e2y
    py = dz * e2x - dx * e2z
    pz = dx * e2y - dy * e2x
    det = (e1x * px + e1y * py) + e1z * pz
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        inv = F32(1.0) / det
        tx, ty, tz = ox - v0[..., 0], oy - v0[..., 1], oz - v0[..., 2]
        u = ((tx * px + ty * py) + tz * pz) * inv
        qx = ty * e1z - tz * e1y
        qy = tz * e1x - tx * e1z
        qz = tx * e1y - ty * e1x
        v = ((dx * qx + dy * qy) + dz * qz) * inv
        t = ((e2x * qx + e2y * qy) + e2z * qz) * inv
        hit = (det != 0) & (u >= 0) & (v >= 0) & ((u + v) <= 1) & (t > 0)
        # slab test against the triangle's own padded box (fminf/fmaxf semantics = np.fmin/np.fmax)
        idx, idy, idz = F32(1.0) / dx, F32(1.0) / dy, F32(1.0) / dz
        ax0, ax1 = (lo[..., 0] - ox) * idx, (hi[..., 0] - ox) * idx
        ay0, ay1 = (lo[..., 1] - oy) * idy, (hi[..., 1] - oy) * idy
        az0, az1 = (lo[..., 2] - oz) * idz, (hi[..., 2] - oz) * idz
        tn = np.fmax(np.fmax(np.fmin(ax0, ax1), np.fmin(ay0, ay1)), np.fmax(np.fmin(az0, az1), F32(0.0)))
        tf = np.fmin(np.fmin(np.fmax(ax0, ax1), np.fmax(ay0, ay1)), np.fmax(az0, az1))
        hit &= (tn <= tf) & (t >= tn)
    return hit, t


def intersect_firstk(origins, dirs, vertices, faces, K: int, pair_budget: int = 1 << 21, threads: int = 1):
    """Brute force over all triangles.  Per ray: every hit, sorted by (t, triangle id), first K.

    Returns tri (N,K) int32 (−1 padded), t (N,K) f32 (+inf padded), count (N,) int32 = min(total,K),
    total (N,) int32 = untruncated hit count.  `threads` > 1 runs ray chunks on a thread pool (numpy
    releases the GIL inside ufuncs); the result does not depend on it."""
    o_all = np.ascontiguousarray(origins, dtype=np.float32)
    d_all = np.ascontiguousarray(dirs, dtype=np.float32)
    N = o_all.shape[0]
    v0, e1, e2, lo, hi = (a[None] for a in triangle_arrays(vertices, faces))
    Fn = v0.shape[1]
    tri = np.full((N, K), -1, dtype=np.int32)
    tt = np.full((N, K), np.inf, dtype=np.float32)
    count = np.zeros(N, dtype=np.int32)
    total = np.zeros(N, dtype=np.int32)
    R = max(1, pair_budget // max(Fn, 1))

    def chunk(s):
        o = o_all[s:s + R, None, :]
        d = d_all[s:s + R, None, :]
        hit, t = _ray_tri_block(o, d, v0, e1, e2, lo, hi)
        rr, ff = np.nonzero(hit)
        if rr.size == 0:
            return
        th = t[rr, ff]
        order = np.lexsort((ff, th, rr))  # ray, then t, then triangle id
        rr, ff, th = rr[order], ff[order], th[order]
        first = np.searchsorted(rr, rr, side="left")
        rank = np.arange(rr.size) - first
        np.add.at(total, s + rr, 1)       # chunks own disjoint ray ranges
        keep = rank < K
        tri[s + rr[keep], rank[keep]] = ff[keep]
        tt[s + rr[keep], rank[keep]] = th[keep]

    starts = list(range(0, N, R))
    if threads > 1 and len(starts) > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(chunk, starts))
    else:
        for s in starts:
            chunk(s)
    count[:] = np.minimum(total, K)
    return tri, tt, count, total


_C_LIB = None


def _c_oracle():
    """oracle/_ref/libqf_oracle.so (oracle/bruteforce.c, built by __graft_entry__.build()) or None."""
    global _C_LIB
    if _C_LIB is None:
        import ctypes
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libqf_oracle.so")
        _C_LIB = ctypes.CDLL(path) if os.path.exists(path) else False
    return _C_LIB or None


def set_c_threads(n: int) -> None:
    """OpenMP thread count of the C restatement (torchrun pins OMP_NUM_THREADS=1 in every rank's environment)."""
    lib = _c_oracle()
    if lib is not None and hasattr(lib, "qf_oracle_set_threads"):
        lib.qf_oracle_set_threads(int(n))


def intersect_firstk_c(origins, dirs, vertices, faces, K: int):
    """Same contract and bit-identical results as `intersect_firstk`, through the C restatement (OpenMP)."""
    import ctypes as C
    lib = _c_oracle()
    if lib is None:
        raise RuntimeError("oracle/_ref/libqf_oracle.so missing: run __graft_entry__.build()")
    o = np.ascontiguousarray(origins, dtype=np.float32)
    d = np.ascontiguousarray(dirs, dtype=np.float32)
    v = np.ascontiguousarray(vertices, dtype=np.float32)
    f = np.ascontiguousarray(faces, dtype=np.int32)
    N, Fn = o.shape[0], f.shape[0]
    tri = np.empty((N, K), dtype=np.int32)
    tt = np.empty((N, K), dtype=np.float32)
    count = np.empty(N, dtype=np.int32)
    total = np.empty(N, dtype=np.int32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.qf_oracle_intersect_firstk.restype = None
    lib.qf_oracle_intersect_firstk(P(o), P(d), C.c_int64(N), P(v), P(f), C.c_int64(Fn), C.c_int(K),
                                   C.c_float(float(mesh_box_pad(v))), P(tri), P(tt), P(count), P(total))
    return tri, tt, count, total


def intersect_firstk_bvh_c(origins, dirs, vertices, faces, K: int, want_total: bool = True):
    """`intersect_firstk` through a CPU bounding-volume hierarchy (oracle/bruteforce.c: qf_oracle_intersect_firstk_bvh):
    the shape of the reference's shipped CPU path (Embree behind trimesh, mesh_utils.py:223,350-354), bit-identical to
    the brute force (tests/test_oracle_golden.py::test_c_bvh_matches_bruteforce).  `want_total=False` lets the traversal
    cull behind the K-th hit (total = -1)."""
    import ctypes as C
    lib = _c_oracle()
    if lib is None or not hasattr(lib, "qf_oracle_intersect_firstk_bvh"):
        raise RuntimeError("oracle/_ref/libqf_oracle.so missing or stale: run __graft_entry__.build()")
    o = np.ascontiguousarray(origins, dtype=np.float32)
    d = np.ascontiguousarray(dirs, dtype=np.float32)
    v = np.ascontiguousarray(vertices, dtype=np.float32)
    f = np.ascontiguousarray(faces, dtype=np.int32)
    N, Fn = o.shape[0], f.shape[0]
    tri = np.empty((N, K), dtype=np.int32)
    tt = np.empty((N, K), dtype=np.float32)
    count = np.empty(N, dtype=np.int32)
    total = np.empty(N, dtype=np.int32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.qf_oracle_intersect_firstk_bvh.restype = None
    lib.qf_oracle_intersect_firstk_bvh(P(o), P(d), C.c_int64(N), P(v), P(f), C.c_int64(Fn), C.c_int(K),
                                       C.c_float(float(mesh_box_pad(v))), C.c_int(1 if want_total else 0), P(tri), P(tt),
                                       P(count), P(total))
    return tri, tt, count, total


def embree_restart_firstk(t_all, tri_all, count_all, K: int, eps: float):
    """What the reference's SHIPPED intersector would keep (SURVEY §8 row a2', recalled trimesh 3.23.5
    `ray_pyembree.intersects_id(multiple_hits=True, max_hits=K)`): a first-hit query repeated up to K times, each time
    restarted `eps` beyond the previous hit along the ray — so hits come out front to back and a hit closer than `eps`
    to the previously KEPT one is never returned.  Input: all hits per ray sorted by (t, id) (`intersect_firstk` with a
    large K).  -> tri (N,K) int32 (-1 padded), count (N,).  Used only to quantify how far the two definitions differ."""
    t_all, tri_all = np.asarray(t_all, dtype=np.float32), np.asarray(tri_all)
    N = t_all.shape[0]
    tri = np.full((N, K), -1, dtype=np.int32)
    count = np.zeros(N, dtype=np.int32)
    eps32 = np.float32(eps)
    for i in range(N):
        last = np.float32(-np.inf)
        c = 0
        for j in range(int(count_all[i])):
            if c == K:
                break
            if np.float32(t_all[i, j] - last) > eps32:          # fp32 difference against the fp32 epsilon (the kernel's test)
                tri[i, c] = tri_all[i, j]
                last = t_all[i, j]
                c += 1
        count[i] = c
    return tri, count


def embree_restart_eps(vertices) -> float:
    """trimesh 3.23.5 `ray_pyembree` (recalled): offset = clip(1e-4 * 100 / mesh.scale, 1e-8, inf) in world units,
    mesh.scale = bounding-box diagonal."""
    v = np.asarray(vertices, dtype=np.float64)
    return float(np.clip(1e-4 * 100.0 / float(np.linalg.norm(v.max(0) - v.min(0))), 1e-8, np.inf))


def plane_hit_points(o, r, n, v):
    """mesh_utils.py:33-40 ``ray_triangle_intersection``: d=−(n·v); t=−((n·o)+d)/(n·r); t←|t|; ψ=o+t r.  fp32."""
    o, r, n, v = (np.asarray(a, dtype=np.float32) for a in (o, r, n, v))
    dot = lambda a, b: (a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1]) + a[:, 2] * b[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        dd = -dot(n, v)
        t = -(dot(n, o) + dd) / dot(n, r)
        t = np.abs(t)
    return o + t[:, None] * r


PREFER_C = False   # bench.py's CPU legs set this: the OpenMP C restatement (identical results) is the faster CPU path
PREFER_BVH = False  # bench.py's CPU legs set this too: CPU BVH traversal (identical results), the reference's Embree shape


def intersects_id(origins, vectors, vertices, faces, max_hits: int, threads: int = 1):
    """The `RayIntersector.intersects_id` contract (mesh_utils.py:87-109): flat
    (triangle_indices, ray_indices, psi) over all kept hits, ray-major in slot order."""
    if PREFER_BVH and _c_oracle() is not None and hasattr(_c_oracle(), "qf_oracle_intersect_firstk_bvh"):
        tri, _, count, _ = intersect_firstk_bvh_c(origins, vectors, vertices, faces, max_hits, want_total=False)
    elif (PREFER_C or np.asarray(faces).shape[0] > 100000) and _c_oracle() is not None:   # the C restatement (identical results)
        tri, _, count, _ = intersect_firstk_c(origins, vectors, vertices, faces, max_hits)
    else:
        tri, _, count, _ = intersect_firstk(origins, vectors, vertices, faces, max_hits, threads=threads)
    slot = np.arange(max_hits)[None, :] < count[:, None]
    ray_indices, _ = np.nonzero(slot)
    triangle_indices = tri[slot].astype(np.int64)
    o = np.asarray(origins, dtype=np.float32)[ray_indices]
    r = np.asarray(vectors, dtype=np.float32)[ray_indices]
    n = face_normals(vertices, faces)[triangle_indices]
    v = np.asarray(vertices, dtype=np.float32)[np.asarray(faces, dtype=np.int64)[triangle_indices, 0]]
    psi = plane_hit_points(o, r, n, v)
    return triangle_indices, ray_indices.astype(np.int64), psi


# --------------------------------------------------------------------------------------
# a3 / a4  hit tuple and re-sort — mesh_utils.py:343-387, 389-412, 225-231
# --------------------------------------------------------------------------------------


def sampling_raytrace(vectors, origins, vertices, faces, max_hits: int, threads: int = 1):
    """`MeshIntersection.sampling_raytrace_numpy` (mesh_utils.py:343-387).  Returns the 7-tuple
    (points, vectors, index_ray, depth, index_tri, 0, origins) or None when nothing is hit.
    The reference's `np.argsort` (:359) is not a stable sort; ties keep (t, triangle id) order here."""
    vectors = np.asarray(vectors, dtype=np.float32)
    origins = np.asarray(origins, dtype=np.float32)
    index_tri, index_ray, points = intersects_id(origins, vectors, vertices, faces, max_hits, threads=threads)
    if index_tri.shape[0] == 0:
        return None
    indices = np.argsort(index_ray, kind="stable")
    index_tri, index_ray, points = index_tri[indices], index_ray[indices], points[indices]
    vectors = vectors[index_ray]
    origins = origins[index_ray]
    norm = _norm3(vectors) + F32(1e-7)
    vectors = vectors / norm[:, None]
    depth = _norm3(points - origins)
    new_indices = np.lexsort((depth, index_ray))
    return (points[new_indices], vectors[new_indices], index_ray[new_indices], depth[new_indices],
            index_tri[new_indices], 0, origins)


def _norm3(a):
    return np.sqrt((a[:, 0] * a[:, 0] + a[:, 1] * a[:, 1]) + a[:, 2] * a[:, 2])


def mark_pack_boundaries(ids: torch.Tensor) -> torch.Tensor:
    """kaolin 0.14.0 `render.spc.mark_pack_boundaries` restated (absent dependency): True where a new pack starts."""
    b = torch.ones_like(ids, dtype=torch.bool)
    if ids.numel() > 1:
        b[1:] = ids[1:] != ids[:-1]
    return b


def sampling_indexing(points, origins, vectors, index_ray, depth, index_tri, render_step_size: float = 0.005):
    """`MeshIntersection.sampling_indexing` (mesh_utils.py:389-412) + `find_deltas` (:225-231)."""
    new_indices = torch.from_numpy(np.lexsort((depth.numpy(), index_ray.numpy())))
    index_tri, index_ray = index_tri[new_indices], index_ray[new_indices]
    points, depth = points[new_indices], depth[new_indices]
    origins, vectors = origins[new_indices], vectors[new_indices]
    boundary = mark_pack_boundaries(index_ray)
    deltas = torch.full((depth.shape[0],), render_step_size, dtype=torch.float32)
    return points, deltas, boundary, vectors, index_ray, depth, index_tri, origins


# --------------------------------------------------------------------------------------
# a5 / a6 / a7  Instant-NGP field — radiance_fields/ngp.py:146-159, 657-809
#   tinycudann (un-pinned git dependency, ngp.py:17-21) is absent: HashGrid, FullyFusedMLP and
#   SphericalHarmonics are restated from its published algorithm.  PARITY UNPINNED for those.
# --------------------------------------------------------------------------------------


@dataclass
class GridMeta:
    n_levels: int
    n_features: int
    scale: np.ndarray       # (L,) f32   grid_scale(level)
    resolution: np.ndarray  # (L,) u32   ceil(scale)+1
    offset: np.ndarray      # (L+1,) i64 entry offsets
    size: np.ndarray        # (L,) i64   entries in level ("hashmap_size")
    hashed: np.ndarray      # (L,) bool  index goes through the spatial hash

    @property
    def n_entries(self) -> int:
        return int(self.offset[-1])


def make_grid_meta(n_levels: int = 16, base_resolution: int = 16, max_resolution: int = 4096,
                   log2_hashmap_size: int = 19, n_features: int = 2,
                   per_level_scale: Optional[float] = None) -> GridMeta:
    """tcnn GridEncoding level table: scale_l = exp2f(l·log2(b))·base − 1; res_l = ceil(scale_l)+1;
    size_l = min(align8(res_l³), 2^log2_T); dense indexing while res³ ≤ size_l.  b as at ngp.py:689-691."""
    if per_level_scale is None:
        per_level_scale = float(np.exp((np.log(max_resolution) - np.log(base_resolution)) / (n_levels - 1)))
    log2_pls = F32(np.log2(F32(per_level_scale)))
    scale = np.zeros(n_levels, dtype=np.float32)
    res = np.zeros(n_levels, dtype=np.uint32)
    size = np.zeros(n_levels, dtype=np.int64)
    hashed = np.zeros(n_levels, dtype=bool)
    offset = np.zeros(n_levels + 1, dtype=np.int64)
    T = 1 << log2_hashmap_size
    for l in range(n_levels):
        s = F32(F32(np.exp2(F32(F32(l) * log2_pls))) * F32(base_resolution)) - F32(1.0)
        scale[l] = s
        r = int(np.ceil(s)) + 1
        res[l] = r
        dense = r ** 3
        n = min((dense + 7) // 8 * 8, T)
        size[l] = n
        hashed[l] = dense > n
        offset[l + 1] = offset[l] + n
    return GridMeta(n_levels, n_features, scale, res, offset, size, hashed)


def _round_h(x: torch.Tensor) -> torch.Tensor:
    """Round to fp16 with a straight-through gradient (tcnn's __half module boundaries)."""
    return x + (x.half().float() - x).detach()


_PRIMES = (1, 2654435761, 805459861)
_U32 = 0xFFFFFFFF


def hashgrid_encode(x01: torch.Tensor, table: torch.Tensor, meta: GridMeta) -> torch.Tensor:
    """tcnn `kernel_grid` forward restated (recalled from tiny-cuda-nn `encodings/grid.h`, the un-pinned master the
    reference installs, ngp.py:17-21).  x01 (M,3) fp32, table (n_entries, F) fp32 holding fp16-representable values.
    Output (M, L·F) fp32 holding fp16 values, level-major.

    Interpolation arithmetic is tcnn's: the grid and the result are `T = __half`; per corner the weight
    ((1·wx)·wy)·wz is formed in fp32, cast to half, and `result = fma((T)weight, grid_val, result)` accumulates with one
    fused half multiply-add (single rounding) in corner order 0..7.  The chain is evaluated exactly here (float64
    product + sum of half values, one correctly rounded conversion to half per corner — numpy's double->half).
    Gradients (table, positions) follow the un-rounded trilinear form, straight through the half roundings."""
    M = x01.shape[0]
    Fdim = meta.n_features
    if (PREFER_C and Fdim == 2 and x01.shape[1] == 3 and not x01.requires_grad and not table.requires_grad
            and _c_oracle() is not None and hasattr(_c_oracle(), "qf_oracle_hashgrid_encode")):
        return hashgrid_encode_c(x01, table, meta)       # bench.py's CPU legs: the OpenMP C restatement, identical results
    outs = []
    xd = x01.double()
    for l in range(meta.n_levels):
        scale = float(meta.scale[l])
        res = int(meta.resolution[l])
        size = int(meta.size[l])
        pos = (xd * scale + 0.5).float()              # fmaf(scale, x, 0.5f)
        cell = torch.floor(pos)
        frac = pos - cell
        cu = cell.to(torch.int64) & _U32              # (uint32_t)(int)floorf(pos)
        lin = torch.zeros((M, Fdim), dtype=torch.float32)
        acc16 = np.zeros((M, Fdim), dtype=np.float16)
        for corner in range(8):
            w = torch.ones(M, dtype=torch.float32)
            g = []
            for dim in range(3):
                if (corner >> dim) & 1:
                    w = w * frac[:, dim]
                    g.append((cu[:, dim] + 1) & _U32)
                else:
                    w = w * (1.0 - frac[:, dim])
                    g.append(cu[:, dim])
            if meta.hashed[l]:
                idx = ((g[0] * _PRIMES[0]) & _U32) ^ ((g[1] * _PRIMES[1]) & _U32) ^ ((g[2] * _PRIMES[2]) & _U32)
            else:
                idx = (g[0] + ((g[1] * res) & _U32) + ((g[2] * ((res * res) & _U32)) & _U32)) & _U32
            idx = idx % size + int(meta.offset[l])
            tv = table[idx]
            lin = lin + w[:, None] * tv
            w16 = w.detach().numpy().astype(np.float16).astype(np.float64)
            acc16 = (w16[:, None] * tv.detach().numpy().astype(np.float64) + acc16.astype(np.float64)).astype(np.float16)
        exact = torch.from_numpy(acc16.astype(np.float32))
        # forward value = the half chain; under autograd the un-rounded trilinear form carries the gradient
        outs.append(lin + (exact - lin).detach() if lin.requires_grad else exact)
    return torch.cat(outs, dim=1)


def hashgrid_encode_c(x01: torch.Tensor, table: torch.Tensor, meta: GridMeta) -> torch.Tensor:
    """`hashgrid_encode` through oracle/bruteforce.c (OpenMP): same arithmetic, same bits, no gradients."""
    import ctypes as C
    lib = _c_oracle()
    x = np.ascontiguousarray(x01.detach().numpy(), dtype=np.float32)
    t = np.ascontiguousarray(table.detach().numpy(), dtype=np.float32)
    L = int(meta.n_levels)
    scale = np.ascontiguousarray(meta.scale, dtype=np.float32)
    res = np.ascontiguousarray(meta.resolution, dtype=np.uint32)
    offset = np.ascontiguousarray(meta.offset, dtype=np.int64)
    size = np.ascontiguousarray(meta.size, dtype=np.int64)
    hashed = np.ascontiguousarray(meta.hashed, dtype=np.uint8)
    out = np.empty((x.shape[0], 2 * L), dtype=np.float32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.qf_oracle_hashgrid_encode.restype = None
    lib.qf_oracle_hashgrid_encode(P(x), C.c_int64(x.shape[0]), P(t), C.c_int(L), P(scale), P(res), P(offset), P(size), P(hashed), P(out))
    return torch.from_numpy(out)


def sh4(d: torch.Tensor) -> torch.Tensor:
    """tcnn SphericalHarmonics degree 4 on unit-ish direction d (already mapped back from [0,1]); (M,16) fp32."""
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    o = torch.empty((d.shape[0], 16), dtype=torch.float32)
    o[:, 0] = 0.28209479177387814
    o[:, 1] = -0.48860251190291987 * y
    o[:, 2] = 0.48860251190291987 * z
    o[:, 3] = -0.48860251190291987 * x
    o[:, 4] = 1.0925484305920792 * xy
    o[:, 5] = -1.0925484305920792 * yz
    o[:, 6] = 0.94617469575755997 * z2 - 0.31539156525251999
    o[:, 7] = -1.0925484305920792 * xz
    o[:, 8] = 0.54627421529603959 * x2 - 0.54627421529603959 * y2
    o[:, 9] = 0.59004358992664352 * y * (-3.0 * x2 + y2)
    o[:, 10] = 2.8906114426405538 * xy * z
    o[:, 11] = 0.45704579946446572 * y * (1.0 - 5.0 * z2)
    o[:, 12] = 0.3731763325901154 * z * (5.0 * z2 - 3.0)
    o[:, 13] = 0.45704579946446572 * x * (1.0 - 5.0 * z2)
    o[:, 14] = 1.4453057213202769 * z * (x2 - y2)
    o[:, 15] = 0.59004358992664352 * x * (-x2 + 3.0 * y2)
    return o


ROUND_HIDDEN = False  # tcnn keeps hidden activations in __half; the oracle's default is the stricter fp32 restatement


def mlp_forward(x: torch.Tensor, weights: Sequence[torch.Tensor]) -> torch.Tensor:
    """tcnn FullyFusedMLP semantics: bias-free, ReLU hidden, linear output; weights[i] is (out,in).  fp32 math
    (hidden activations optionally rounded to fp16 with a straight-through gradient, see ROUND_HIDDEN)."""
    h = x
    for W in weights[:-1]:
        h = torch.relu(h @ W.t())
        if ROUND_HIDDEN:
            h = _round_h(h)
    return h @ weights[-1].t()


@dataclass
class NGPParams:
    """Explicit parameters of `NGPRadianceField` (ngp.py:660-746) in tcnn's layout.

    table   (n_entries, 2) fp32, fp16-representable
    base_w  [(64,32), (16,64)]            mlp_base  32→64→16
    head_w  [(64,32), (64,64), (16,64)]   mlp_head  31(+1 pad)→64→64→3(+13 pad)
    """
    aabb: torch.Tensor
    meta: GridMeta
    table: torch.Tensor
    base_w: List[torch.Tensor]
    head_w: List[torch.Tensor]


HEAD_PAD_VALUE = 1.0  # tcnn pads the 31-wide head input to 32 with ones (Identity-encoding alignment padding)


def make_ngp_params(seed: int = 42, log2_hashmap_size: int = 19, table_scale: float = 1e3,
                    aabb=(-1.5, -1.5, -1.5, 1.5, 1.5, 1.5)) -> NGPParams:
    """Random-init field of SURVEY §8d C1: table U(−1e-4,1e-4)·table_scale, Xavier-uniform MLPs, rounded to fp16."""
    g = torch.Generator().manual_seed(seed)
    meta = make_grid_meta(log2_hashmap_size=log2_hashmap_size)
    table = ((torch.rand((meta.n_entries, 2), generator=g) * 2 - 1) * 1e-4 * table_scale).half().float()

    def xavier(o, i):
        b = math.sqrt(6.0 / (i + o))
        return ((torch.rand((o, i), generator=g) * 2 - 1) * b).half().float()

    base_w = [xavier(64, 32), xavier(16, 64)]
    head_w = [xavier(64, 32), xavier(64, 64), xavier(16, 64)]
    return NGPParams(torch.tensor(aabb, dtype=torch.float32), meta, table, base_w, head_w)


def ngp_normalize(x: torch.Tensor, aabb: torch.Tensor):
    """ngp.py:748-755."""
    aabb_min, aabb_max = aabb[:3], aabb[3:]
    x = (x - aabb_min) / (aabb_max - aabb_min)
    selector = ((x > 0.0) & (x < 1.0)).all(dim=-1)
    return selector, x


def ngp_query_density(x: torch.Tensor, p: NGPParams):
    """ngp.py:757-779: σ = trunc_exp(h0 − 1)·selector, feat = h[1:16]."""
    selector, x01 = ngp_normalize(x, p.aabb)
    enc = hashgrid_encode(x01, p.table, p.meta)
    h = mlp_forward(enc, p.base_w)
    density = torch.exp(h[:, 0:1] - 1.0) * selector[:, None]
    return density, h[:, 1:16]


def ngp_query_rgb(dirs: torch.Tensor, embedding: torch.Tensor, p: NGPParams, apply_act: bool = True):
    """ngp.py:781-796: SH4((d+1)/2) ⊕ feat → head MLP → sigmoid."""
    d01 = (dirs + 1.0) / 2.0
    sh = sh4(d01 * 2.0 - 1.0).half().float()
    pad = torch.full((dirs.shape[0], 1), HEAD_PAD_VALUE, dtype=torch.float32)
    h = torch.cat([sh, _round_h(embedding), pad], dim=-1)
    rgb = mlp_forward(h, p.head_w)[:, :3]
    return torch.sigmoid(rgb) if apply_act else rgb


def ngp_forward(positions: torch.Tensor, directions: torch.Tensor, p: NGPParams):
    """ngp.py:798-809 → (rgb (M,3), density (M,1))."""
    assert positions.shape == directions.shape, f"{positions.shape} v.s. {directions.shape}"
    density, emb = ngp_query_density(positions, p)
    return ngp_query_rgb(directions, emb, p), density


def ngp_relu_margin(positions: torch.Tensor, directions: torch.Tensor, p: NGPParams) -> torch.Tensor:
    """Per sample, the smallest |pre-activation| over the 192 hidden units of `ngp_forward` (tcnn precision: hidden
    activations rounded to half).  The gradient of the field is discontinuous where a pre-activation crosses zero, so two
    correct implementations whose accumulation orders differ disagree by a whole unit's gradient on samples with a
    margin below their rounding noise; gradient parity tests leave those samples out."""
    with torch.no_grad():
        _, x01 = ngp_normalize(positions, p.aabb)
        enc = hashgrid_encode(x01, p.table, p.meta)
        pre1 = enc @ p.base_w[0].t()
        out = _round_h(torch.relu(pre1)) @ p.base_w[1].t()
        sh = sh4((directions + 1.0) / 2.0 * 2.0 - 1.0).half().float()
        pad = torch.full((directions.shape[0], 1), HEAD_PAD_VALUE, dtype=torch.float32)
        pre3 = torch.cat([sh, _round_h(out[:, 1:16]), pad], dim=-1) @ p.head_w[0].t()
        pre4 = _round_h(torch.relu(pre3)) @ p.head_w[1].t()
        return torch.minimum(torch.minimum(pre1.abs().min(dim=1).values, pre3.abs().min(dim=1).values), pre4.abs().min(dim=1).values)


# --------------------------------------------------------------------------------------
# a8  spherical-Gaussian head — ngp.py:371-393, 456-461 ;  texture_utils.py:126-147
# --------------------------------------------------------------------------------------


def sg_features_to_rgb(features: torch.Tensor, dirs: torch.Tensor, num_lobes: int) -> torch.Tensor:
    """rgb = sigmoid(diffuse + Σ_l c_l·exp(|λ_l|·(â_l·d − 1))), lobe layout [a(3), λ, c(3)] (ngp.py:371-393,456-461)."""
    rgb = features[:, :3].clone()
    x = features[:, 3:3 + 7 * num_lobes]
    for l in range(num_lobes):
        lobe = x[:, 7 * l:7 * l + 7]
        axis = lobe[:, :3]
        axis = axis / torch.linalg.norm(axis, dim=-1, keepdim=True)
        lam = torch.abs(lobe[:, 3])
        rgb = rgb + lobe[:, 4:7] * torch.exp(lam * (torch.sum(axis * dirs, -1) - 1))[:, None]
    return torch.sigmoid(rgb)


# --------------------------------------------------------------------------------------
# a9  baked texture decode — texture_utils.py:149-175, 61-65 ; ngp.py:245-281
# --------------------------------------------------------------------------------------


@dataclass
class TextureSet:
    alpha: torch.Tensor            # (S,S) u8
    diffuse: torch.Tensor          # (S,S,3) u8
    sg_colors: List[torch.Tensor]  # L × (S,S,3) u8
    lambdas: List[torch.Tensor]    # L × (S,S,3) u8  [λ, azimuth, elevation]
    compression_type: str = "linear"
    lambda_thres: float = 7.5

    @property
    def num_lobes(self):
        return len(self.sg_colors)

    @property
    def texture_size(self):
        return self.alpha.shape[0]


def make_texture_set(size: int, num_lobes: int, seed: int = 42, compression_type="linear", lambda_thres=7.5):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randint(0, 256, s, generator=g, dtype=torch.int32).to(torch.uint8)
    return TextureSet(r(size, size), r(size, size, 3), [r(size, size, 3) for _ in range(num_lobes)],
                      [r(size, size, 3) for _ in range(num_lobes)], compression_type, lambda_thres)


def _inv_colors(c_u8: torch.Tensor, compress_type: str, thres: float = 12.0):
    """ngp.py:275-281.  Quirk Q5: only the literal "sigma" takes the logit branch."""
    c = c_u8.to(torch.float32) / 255.0
    if compress_type == "sigma":
        return torch.log(torch.clip(c / (1 - c), 1e-8, 1e37))
    return c * 2 * thres - thres


def texture_decode(indices: torch.Tensor, tex: TextureSet) -> torch.Tensor:
    """`FeatureCompression.get_features_from_texture_map` (texture_utils.py:149-175) → (M, 3+7L+1) fp32
    laid out [diffuse(3), L×(axis3, λ, c3), σ]."""
    i0, i1 = indices[:, 0], indices[:, 1]
    a = tex.alpha[i0, i1].to(torch.float32) / 255.0
    sigma = -torch.log(torch.clip(1 - a, 1e-6)) / 0.005                       # texture_utils.py:61-65
    diffuse = _inv_colors(tex.diffuse[i0, i1], tex.compression_type)
    cols = [diffuse]
    for l in range(tex.num_lobes):
        sh = tex.lambdas[l][i0, i1]
        lam = torch.exp(sh[:, 0] * tex.lambda_thres / 255 - 2.5)               # ngp.py:260-262
        az = (sh[:, 1] - 128) / 128 * np.pi                                   # ngp.py:245-252, uint8 wrap (Q6)
        el = sh[:, 2] / 256 * np.pi
        axis = torch.stack([torch.cos(az) * torch.sin(el), torch.sin(az) * torch.sin(el), torch.cos(el)], dim=-1)
        cols += [axis, lam[:, None], _inv_colors(tex.sg_colors[l][i0, i1], tex.compression_type)]
    cols.append(sigma[:, None])
    return torch.cat(cols, dim=-1)


def texture_compress(features: torch.Tensor, num_lobes: int, compression_type: str = "linear", lambda_thres: float = 7.5):
    """SURVEY §8 row f-4 — the bake writer `FeatureCompression.compress` (texture_utils.py:67-98) with the quantisers
    of ngp.py:239-273 and texture_utils.py:51-55.  features (M, 3+7L+1) -> dict(alpha (M), diffuse (M,3),
    lambdas L x (M,3) [lambda, azimuth, elevation], colors L x (M,3)) uint8."""
    def colors(c):
        if compression_type == "sigma":
            c = torch.sigmoid(c)
        else:
            c = (torch.clip(c, -12, 12) + 12) / 2 / 12
        return (c * 255).to(torch.uint8)
    M = features.shape[0]
    sigma = features[:, -1]
    alpha = torch.clip((1 - torch.exp(-sigma * 0.005)) * 255, 0, 255).to(torch.uint8)
    lobes = features[:, 3:-1].reshape(M, num_lobes, 7)
    v = lobes[..., :3] / (torch.norm(lobes[..., :3], dim=-1, keepdim=True) + 1e-6)
    az = (torch.atan2(v[..., 1], v[..., 0]) * 128 / np.pi + 128).to(torch.uint8)
    el = (torch.acos(v[..., 2]) * 256 / np.pi).to(torch.uint8)
    lam = torch.clamp((torch.log(torch.clamp(torch.abs(lobes[..., 3]), 1e-5, np.inf)) + 2.5) / lambda_thres, 0.0, 1.0)
    lam = (255 * lam).to(torch.uint8)
    return dict(alpha=alpha, diffuse=colors(features[:, :3]),
                lambdas=[torch.stack([lam[:, i], az[:, i], el[:, i]], dim=-1) for i in range(num_lobes)],
                colors=[colors(lobes[:, i, 4:]) for i in range(num_lobes)])


# --------------------------------------------------------------------------------------
# a10  barycentric → texel — utils.py:1055-1063 (trimesh.triangles.points_to_barycentric, fp64)
# --------------------------------------------------------------------------------------


def points_to_barycentric(triangles: np.ndarray, points: np.ndarray) -> np.ndarray:
    """trimesh 3.23.5 `triangles.points_to_barycentric(method="cramer")` restated, fp64 (absent dependency)."""
    tri = np.asarray(triangles, dtype=np.float64)
    pts = np.asarray(points, dtype=np.float64)
    ev = tri[:, 1:] - tri[:, :1]
    w = pts - tri[:, 0]
    dot = lambda a, b: (a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1]) + a[:, 2] * b[:, 2]
    d00, d01, d11 = dot(ev[:, 0], ev[:, 0]), dot(ev[:, 0], ev[:, 1]), dot(ev[:, 1], ev[:, 1])
    d02, d12 = dot(ev[:, 0], w), dot(ev[:, 1], w)
    inv = 1.0 / (d00 * d11 - d01 * d01)
    b = np.zeros((tri.shape[0], 3), dtype=np.float64)
    b[:, 2] = (d00 * d12 - d01 * d02) * inv
    b[:, 1] = (d11 * d02 - d01 * d12) * inv
    b[:, 0] = 1.0 - b[:, 1] - b[:, 2]
    return b


def hit_texels(points: torch.Tensor, index_tri: torch.Tensor, vertices: np.ndarray, faces: np.ndarray,
               uv_scaled: torch.Tensor, texture_size: int) -> torch.Tensor:
    """utils.py:1055-1063: b = clamp(bary,0,1); b/=Σb; uv = Σ b_k uv_k; floor; clip to [0,S−1].  (M,2) int64."""
    f = np.asarray(faces, dtype=np.int64)[index_tri.numpy()]
    uv_ = uv_scaled[torch.from_numpy(f)]
    b = points_to_barycentric(np.asarray(vertices, dtype=np.float64)[f], points.numpy())
    b = torch.clamp(torch.from_numpy(b.astype(np.float32)), 0, 1)
    b = b / ((b[:, 0] + b[:, 1]) + b[:, 2])[:, None]
    uv_points = (uv_[:, 0] * b[:, 0:1] + uv_[:, 1] * b[:, 1:2]) + uv_[:, 2] * b[:, 2:3]
    return torch.clip(torch.floor(uv_points).long(), 0, texture_size - 1)


def scale_uv(uv: np.ndarray, size: int) -> torch.Tensor:
    """test_baking_texture_images.py:325-328."""
    uv = np.asarray(uv) - 1e-7
    uv = np.array(uv).astype(np.float32) * size
    return torch.from_numpy(np.clip(uv, 0, size - 1))


# --------------------------------------------------------------------------------------
# a11  mesh-path compositing — utils.py:863-898 (kaolin exponential_integration / sum_reduce restated)
# --------------------------------------------------------------------------------------


def _segment_ids(boundary: torch.Tensor) -> torch.Tensor:
    return torch.cumsum(boundary.to(torch.int64), 0) - 1


def segmented_exclusive_cumsum(x: torch.Tensor, boundary: torch.Tensor) -> torch.Tensor:
    """kaolin `cumsum(..., exclusive=True)` per pack; fp64 accumulate so the oracle is order-independent."""
    seg = _segment_ids(boundary)
    xd = x.double()
    inc = torch.cumsum(xd, 0)
    exc = inc - xd
    starts = torch.nonzero(boundary).flatten()
    base = exc[starts]
    return (exc - base[seg]).to(x.dtype)


def segmented_sum(x: torch.Tensor, boundary: torch.Tensor) -> torch.Tensor:
    seg = _segment_ids(boundary)
    n = int(seg[-1]) + 1 if seg.numel() else 0
    out = torch.zeros((n,) + x.shape[1:], dtype=torch.float64)
    out.index_add_(0, seg, x.double())
    return out.to(x.dtype)


def exponential_integration(feats, tau, boundary):
    """kaolin 0.14.0 `render.spc.exponential_integration(exclusive=True)` restated: returns
    (Σ_pack w·feat, w) with w = exp(−excl-cumsum τ)·(1 − e^{−τ})."""
    alpha = 1.0 - torch.exp(-tau)
    w = torch.exp(-segmented_exclusive_cumsum(tau, boundary)) * alpha
    return segmented_sum(w * feats, boundary), w


def derive_properties(color, density, depths, deltas, boundary, index_ray, render_bkgd=None,
                      bg_color="white", N=0):
    """utils.py:863-898 including quirks Q1 (α multiplies the already-weighted colour) and Q2 (fills)."""
    color = color.reshape(-1, 3)
    tau = (density * deltas).reshape(-1, 1)
    ray_colors, transmittance = exponential_integration(color, tau, boundary)
    depths, _ = exponential_integration(depths.reshape(-1, 1), tau, boundary)
    alpha = segmented_sum(transmittance, boundary)
    out_alpha = torch.zeros(N, 1)
    Depth = torch.zeros(N, 1)
    if bg_color == "white":
        rgb = torch.ones(N, 3)
        color = (1.0 - alpha) + alpha * ray_colors
    elif bg_color == "black":
        rgb = torch.zeros(N, 3)
        color = alpha * ray_colors
    else:
        rgb = torch.ones(N, 3)
        color = alpha * ray_colors + (1.0 - alpha) * render_bkgd
    ids = index_ray[boundary]
    Depth[ids] = depths.float()
    rgb[ids] = color.float()
    out_alpha[ids] = alpha.float()
    return rgb, out_alpha, ids, Depth, transmittance


# --------------------------------------------------------------------------------------
# a12–a15  nerfacc-style compositing — field_rendering.py (nerfacc 0.5.3 pack/scan restated)
# --------------------------------------------------------------------------------------


def pack_info(ray_indices: torch.Tensor, n_rays: Optional[int] = None) -> torch.Tensor:
    """nerfacc 0.5.3 `pack.pack_info`: counts by index_add, starts by cumsum; assumes ascending ids (Q3)."""
    if n_rays is None:
        n_rays = int(ray_indices.max()) + 1 if ray_indices.numel() else 0
    cnt = torch.zeros((n_rays,), dtype=torch.long)
    cnt.index_add_(0, ray_indices.long(), torch.ones_like(ray_indices, dtype=torch.long))
    start = cnt.cumsum(0) - cnt
    return torch.stack([start, cnt], dim=-1)


def _packed_scan(x: torch.Tensor, packed_info: Optional[torch.Tensor], prod: bool) -> torch.Tensor:
    if packed_info is None:  # batched (n_rays, S)
        if prod:
            return torch.cumprod(torch.cat([torch.ones_like(x[..., :1]), x[..., :-1]], dim=-1), dim=-1)
        return torch.cumsum(torch.cat([torch.zeros_like(x[..., :1]), x[..., :-1]], dim=-1), dim=-1)
    out = torch.empty_like(x)
    for s, c in packed_info.tolist():
        seg = x[s:s + c]
        if c == 0:
            continue
        if prod:
            out[s:s + c] = torch.cumprod(torch.cat([torch.ones(1, dtype=x.dtype), seg[:-1]]), 0)
        else:
            out[s:s + c] = torch.cumsum(torch.cat([torch.zeros(1, dtype=x.dtype), seg[:-1]]), 0)
    return out


def exclusive_sum(x, packed_info=None):
    return _packed_scan(x, packed_info, prod=False)


def exclusive_prod(x, packed_info=None):
    return _packed_scan(x, packed_info, prod=True)


def render_transmittance_from_alpha(alphas, packed_info=None, ray_indices=None, n_rays=None, prefix_trans=None):
    """field_rendering.py:161-206."""
    if ray_indices is not None and packed_info is None:
        packed_info = pack_info(ray_indices, n_rays)
    trans = exclusive_prod(1 - alphas, packed_info)
    if prefix_trans is not None:
        trans = trans * prefix_trans
    return trans


def render_transmittance_from_density(t_starts, t_ends, sigmas, packed_info=None, ray_indices=None,
                                      n_rays=None, prefix_trans=None):
    """field_rendering.py:209-264."""
    if ray_indices is not None and packed_info is None:
        packed_info = pack_info(ray_indices, n_rays)
    sigmas_dt = sigmas * (t_ends - t_starts)
    alphas = 1.0 - torch.exp(-sigmas_dt)
    trans = torch.exp(-exclusive_sum(sigmas_dt, packed_info))
    if prefix_trans is not None:
        trans = trans * prefix_trans
    return trans, alphas


def render_weight_from_alpha(alphas, packed_info=None, ray_indices=None, n_rays=None, prefix_trans=None):
    """field_rendering.py:267-309."""
    trans = render_transmittance_from_alpha(alphas, packed_info, ray_indices, n_rays, prefix_trans)
    return trans * alphas, trans


def render_weight_from_density(t_starts, t_ends, sigmas, packed_info=None, ray_indices=None, n_rays=None,
                               prefix_trans=None):
    """field_rendering.py:312-362."""
    trans, alphas = render_transmittance_from_density(t_starts, t_ends, sigmas, packed_info, ray_indices,
                                                      n_rays, prefix_trans)
    return trans * alphas, trans, alphas


def render_visibility_from_alpha(alphas, packed_info=None, ray_indices=None, n_rays=None,
                                 early_stop_eps=1e-4, alpha_thre=0.0, prefix_trans=None):
    """field_rendering.py:365-418."""
    trans = render_transmittance_from_alpha(alphas, packed_info, ray_indices, n_rays, prefix_trans)
    vis = trans >= early_stop_eps
    if alpha_thre > 0:
        vis = vis & (alphas >= alpha_thre)
    return vis


def render_visibility_from_density(t_starts, t_ends, sigmas, packed_info=None, ray_indices=None, n_rays=None,
                                   early_stop_eps=1e-4, alpha_thre=0.0, prefix_trans=None):
    """field_rendering.py:421-480."""
    trans, alphas = render_transmittance_from_density(t_starts, t_ends, sigmas, packed_info, ray_indices,
                                                      n_rays, prefix_trans)
    vis = trans >= early_stop_eps
    if alpha_thre > 0:
        vis = vis & (alphas >= alpha_thre)
    return vis


def accumulate_along_rays(weights, values=None, ray_indices=None, n_rays=None):
    """field_rendering.py:483-547."""
    src = weights[..., None] if values is None else weights[..., None] * values
    if ray_indices is not None:
        out = torch.zeros((n_rays, src.shape[-1]), dtype=src.dtype)
        out.index_add_(0, ray_indices, src)
        return out
    return torch.sum(src, dim=-2)


def rendering(t_starts, t_ends, ray_indices=None, n_rays=None, rgbs=None, sigmas=None, alphas=None,
              render_bkgd=None):
    """field_rendering.py:14-158 with the callback already evaluated (rgbs + sigmas, or rgbs + alphas)."""
    if sigmas is not None:
        weights, trans, alphas = render_weight_from_density(t_starts, t_ends, sigmas, ray_indices=ray_indices,
                                                            n_rays=n_rays)
    else:
        weights, trans = render_weight_from_alpha(alphas, ray_indices=ray_indices, n_rays=n_rays)
    colors = accumulate_along_rays(weights, rgbs, ray_indices, n_rays)
    opac = accumulate_along_rays(weights, None, ray_indices, n_rays)
    depths = accumulate_along_rays(weights, (t_starts + t_ends)[..., None] / 2.0, ray_indices, n_rays)
    depths = depths / opac.clamp_min(torch.finfo(rgbs.dtype).eps)
    if render_bkgd is not None:
        colors = colors + render_bkgd * (1.0 - opac)
    return colors, opac, depths, dict(weights=weights, trans=trans, alphas=alphas)


def reversed_weights(t_starts, t_ends, sigmas, ray_indices, n_rays):
    """field_rendering.py:719-731 (quirk Q3: pack_info of the *flipped* descending ray ids)."""
    max_val = torch.max(t_starts) + torch.max(t_ends)
    ts = torch.flip(max_val - t_starts, dims=[0])
    te = torch.flip(max_val - t_ends, dims=[0])
    sg = torch.flip(sigmas, dims=[0])
    w_rev, _, _ = render_weight_from_density(te, ts, sg, ray_indices=torch.flip(ray_indices, dims=[0]), n_rays=n_rays)
    return torch.flip(w_rev, dims=[0])


# --------------------------------------------------------------------------------------
# end-to-end drivers (what `bench.py`'s cpu_baseline and the parity tests run)
# --------------------------------------------------------------------------------------


def render_mesh_ngp(origins, viewdirs, vertices, faces, params: NGPParams, K: int, render_step_size=0.005,
                    bg_color="white", render_bkgd=None, timings: Optional[dict] = None, threads: int = 1):
    """utils.py:465-607 with scaling=0 (no deformation): intersect → sort → field(points, viewdirs[index_ray])
    → derive_properties.  Returns dict(rgb (N,3), opacity (N,1), depth (N,1), + per-hit arrays)."""
    import time
    N = origins.shape[0]
    t0 = time.perf_counter()
    tup = sampling_raytrace(viewdirs, origins, vertices, faces, K, threads=threads)
    t1 = time.perf_counter()
    if tup is None:  # Q9 → Q2 fill
        fill = 0.0 if bg_color == "black" else 1.0
        return dict(rgb=torch.full((N, 3), fill), opacity=torch.zeros(N, 1), depth=torch.zeros(N, 1),
                    index_ray=torch.zeros(0, dtype=torch.long), index_tri=torch.zeros(0, dtype=torch.long),
                    weights=torch.zeros(0, 1), points=torch.zeros(0, 3), ts=torch.zeros(0))
    points, dirs, index_ray, depth, index_tri, _, _ = tup
    points_t = torch.from_numpy(points.astype(np.float32))
    index_ray_t = torch.from_numpy(index_ray.astype(np.int64))
    depth_t = torch.from_numpy(depth.astype(np.float32))
    t_dirs = torch.from_numpy(np.asarray(viewdirs, dtype=np.float32))[index_ray_t]      # Q7: original viewdirs
    rgbs, sigmas = ngp_forward(points_t, t_dirs, params)
    t2 = time.perf_counter()
    boundary = mark_pack_boundaries(index_ray_t)
    deltas = torch.full((points_t.shape[0],), render_step_size, dtype=torch.float32)
    rgb, opacity, _, Depth, weights = derive_properties(rgbs, sigmas.squeeze(-1), depth_t, deltas, boundary,
                                                        index_ray_t, render_bkgd=render_bkgd, bg_color=bg_color, N=N)
    t3 = time.perf_counter()
    if timings is not None:
        timings.update(intersect_s=t1 - t0, field_s=t2 - t1, composite_s=t3 - t2)
    return dict(rgb=rgb, opacity=opacity, depth=Depth, index_ray=index_ray_t,
                index_tri=torch.from_numpy(index_tri.astype(np.int64)), weights=weights, points=points_t,
                ts=depth_t, rgbs=rgbs, sigmas=sigmas)


def render_mesh_baked(origins, viewdirs, vertices, faces, uv_scaled, tex: TextureSet, K: int,
                      render_step_size=0.005, bg_color="white"):
    """utils.py:998-1095: intersect → sort → texel lookup → decode → SG (tuple dirs, Q7) → derive_properties."""
    N = origins.shape[0]
    tup = sampling_raytrace(viewdirs, origins, vertices, faces, K)
    if tup is None:
        fill = 0.0 if bg_color == "black" else 1.0
        return dict(rgb=torch.full((N, 3), fill), opacity=torch.zeros(N, 1), depth=torch.zeros(N, 1))
    points, dirs, index_ray, depth, index_tri, _, _ = tup
    points_t = torch.from_numpy(points.astype(np.float32))
    dirs_t = torch.from_numpy(dirs.astype(np.float32))
    index_ray_t = torch.from_numpy(index_ray.astype(np.int64))
    index_tri_t = torch.from_numpy(index_tri.astype(np.int64))
    depth_t = torch.from_numpy(depth.astype(np.float32))
    texels = hit_texels(points_t, index_tri_t, vertices, faces, uv_scaled, tex.texture_size)
    feats = texture_decode(texels, tex)
    sigmas = feats[:, -1]
    rgbs = sg_features_to_rgb(feats[:, :-1], dirs_t, tex.num_lobes)
    boundary = mark_pack_boundaries(index_ray_t)
    deltas = torch.full((points_t.shape[0],), render_step_size, dtype=torch.float32)
    rgb, opacity, _, Depth, weights = derive_properties(rgbs, sigmas, depth_t, deltas, boundary, index_ray_t,
                                                        bg_color=bg_color, N=N)
    return dict(rgb=rgb, opacity=opacity, depth=Depth, index_ray=index_ray_t, index_tri=index_tri_t,
                weights=weights, texels=texels, rgbs=rgbs, sigmas=sigmas, points=points_t)


# ---------------------------------------------------------------------------------------------------------------------
# f-3: mesh finetuning accumulators (mesh_utils.py:112-156; prune_mesh_after_finetuning.py:354-357)
# ---------------------------------------------------------------------------------------------------------------------
def mesh_finetune_update_d(cache_d, cache_w, d, w, index_tri):
    """MeshFinetune.update_d (mesh_utils.py:126-133): scatter d*w and w per triangle, add to the caches (fp64 sums so the
    oracle does not depend on accumulation order; the kernel's fp32 atomics are compared with a tolerance)."""
    cache_d, cache_w = np.asarray(cache_d, np.float32), np.asarray(cache_w, np.float32)
    d, w, index_tri = np.asarray(d, np.float32), np.asarray(w, np.float32), np.asarray(index_tri, np.int64)
    acc_d = np.zeros(cache_d.shape, np.float64)
    acc_w = np.zeros(cache_w.shape, np.float64)
    np.add.at(acc_d, index_tri, (d * w[:, None]).astype(np.float64))
    np.add.at(acc_w, index_tri, w.astype(np.float64))
    return (cache_d + acc_d).astype(np.float32), (cache_w + acc_w).astype(np.float32)


def mesh_finetune_update_faces(vertices, faces, cache_d, cache_w, scaling):
    """MeshFinetune.update_faces (mesh_utils.py:135-144): clip(cache_d / cache_w, +-scaling) per triangle, scatter_mean over
    the face corners (count clamped to >= 1), vertices += mean."""
    vertices, faces = np.asarray(vertices, np.float32), np.asarray(faces, np.int64)
    deform = np.clip(np.asarray(cache_d, np.float32) / np.asarray(cache_w, np.float32)[:, None], -np.float32(scaling), np.float32(scaling))
    sums = np.zeros(vertices.shape, np.float64)
    cnt = np.zeros(vertices.shape[0], np.float64)
    np.add.at(sums, faces.reshape(-1), np.repeat(deform, 3, axis=0).astype(np.float64))
    np.add.at(cnt, faces.reshape(-1), 1.0)
    dv = (sums / np.maximum(cnt, 1.0)[:, None]).astype(np.float32)
    return vertices + dv


def triangle_weight_max(tri_w, weights, index_tri):
    """prune pass (prune_mesh_after_finetuning.py:354-357): scatter_max into zeros, then the running torch.maximum."""
    cur = np.zeros_like(np.asarray(tri_w, np.float32))
    np.maximum.at(cur, np.asarray(index_tri, np.int64), np.asarray(weights, np.float32))
    return np.maximum(np.asarray(tri_w, np.float32), cur)


# ---------------------------------------------------------------------------------------------------------------------
# f-2: the quadrature Field net (field.py:130-259), back_prop=False
# ---------------------------------------------------------------------------------------------------------------------
class _HalfBoundary(torch.autograd.Function):
    """A tensor handed over as `__half`: values AND the gradient coming back are rounded to fp16 (the cast that
    torch.cat([x, h]) inserts for the half encoder output, field.py:193, rounds the incoming gradient the same way)."""

    @staticmethod
    def forward(ctx, v):
        return v.half().float()

    @staticmethod
    def backward(ctx, g):
        return g.half().float()


def field_net_forward(x, table, meta, w1, b1, w2, b2, w3, b3, xyz_min, xyz_max, activation="elu", return_grad=True):
    """Field.forward (field.py:203-221): x (M,3) -> field (M,out_dim), field_grad (M,3).  The grid sees x.detach()
    (field.py:196-199), so field_grad = d(sum field)/dx flows through the raw-xyz inputs of the MLP only; it is built with
    create_graph=True like the reference, so losses on it back-propagate into every parameter passed with requires_grad.
    `table` (n_entries,2) fp32 master parameters (rounded to fp16 with a straight-through gradient, tcnn's working copy)."""
    act = torch.nn.functional.elu if activation == "elu" else torch.relu
    x = x.detach().clone().requires_grad_(True)
    lo, hi = torch.as_tensor(xyz_min, dtype=torch.float32), torch.as_tensor(xyz_max, dtype=torch.float32)
    x01 = (x - lo) / (hi - lo)
    h = _HalfBoundary.apply(hashgrid_encode(x01.detach(), _round_h(table), meta))     # tcnn emits / receives __half
    inp = torch.cat([x01, h], dim=1)
    a1 = act(torch.nn.functional.linear(inp, w1, b1))
    a2 = act(torch.nn.functional.linear(a1, w2, b2))
    field = torch.nn.functional.linear(a2, w3, b3)
    if not return_grad:
        return field, None
    grad = torch.autograd.grad(field.flatten(), [x], grad_outputs=torch.ones_like(field.flatten()), create_graph=True,
                               retain_graph=True)[0]
    return field, grad


def compute_field_loss(weights, weights_rev, field_norm, view_dirs):
    """Field.compute_field_loss (field.py:253-259)."""
    view_dirs = view_dirs / torch.norm(view_dirs, dim=1, keepdim=True)
    loss = torch.abs(torch.maximum(weights.detach(), weights_rev.detach()) - torch.abs(torch.sum(field_norm * view_dirs.detach(), 1)))
    return loss.mean()


# ---------------------------------------------------------------------------------------------------------------------
# f-1: occupancy-grid ray marcher (nerfacc 0.5.3 `OccGridEstimator.sampling` -> `traverse_grids`; utils.py:137-148,
# 241-285, 422-433).  PARITY UNPINNED: nerfacc is an absent third-party dependency; this restates its published
# algorithm as recalled — samples sit on a per-ray grid t_{k+1} = t_k + dt that starts where the ray enters the outermost
# box (clipped to [near, far]), dt = clamp(t * cone_angle, step, 1e10); sample k is kept when its MIDPOINT lies before
# the exit point and in an occupied cell of the finest grid level whose box contains it ("march until t_mid is right
# after t_traverse").  Ties of a midpoint with a cell face follow floor() of the normalised position here (nerfacc walks
# the cells with a DDA), which can differ on a set of measure zero.
# ---------------------------------------------------------------------------------------------------------------------
def occgrid_aabbs(roi_aabb, levels: int) -> np.ndarray:
    """nerfacc `OccGridEstimator.__init__`: level l is the region of interest scaled by 2^l about its centre."""
    roi = np.asarray(roi_aabb, dtype=np.float32)
    c, h = (roi[:3] + roi[3:]) / F32(2), (roi[3:] - roi[:3]) / F32(2)
    return np.stack([np.concatenate([c - h * F32(2 ** l), c + h * F32(2 ** l)]) for l in range(levels)]).astype(np.float32)


def occgrid_cell_scale(aabbs: np.ndarray, resolution) -> np.ndarray:
    """(L,3) fp32: cells per unit length, fl(R / fl(hi - lo)), shared by oracle and kernel."""
    res = np.asarray(resolution, dtype=np.float32).reshape(1, 3)
    return (res / (aabbs[:, 3:] - aabbs[:, :3])).astype(np.float32)


def occgrid_march(origins, dirs, binaries, aabbs, near_planes, far_plane, step_size, cone_angle=0.0, max_samples=0,
                  ray_mask=None, return_termination=False):
    """-> (ray_indices int64 (M,), t_starts (M,), t_ends (M,), counts int32 (N,)), ray-major.  fp32, one rounding per
    operation in the order written (the kernel file is compiled without FMA contraction).
    `max_samples` > 0: a ray stops after that many samples; `ray_mask` (N,) bool: retired rays emit nothing;
    `return_termination`: a fifth value, the plane (N,) where every ray stopped (utils.py:254-330's chunked marching)."""
    o, d = np.asarray(origins, dtype=np.float32), np.asarray(dirs, dtype=np.float32)
    N = o.shape[0]
    B = np.asarray(binaries).astype(bool)
    L, R = B.shape[0], np.asarray(B.shape[1:], dtype=np.int64)
    aabbs = np.asarray(aabbs, dtype=np.float32)
    scale = occgrid_cell_scale(aabbs, R)
    near = np.broadcast_to(np.asarray(near_planes, dtype=np.float32), (N,)).copy()
    far, step, cone = F32(far_plane), F32(step_size), F32(cone_angle)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        inv = F32(1.0) / d
        lo, hi = aabbs[L - 1, :3], aabbs[L - 1, 3:]
        t0, t1 = (lo[None] - o) * inv, (hi[None] - o) * inv
        tmin = np.fmax(np.fmax(np.fmin(t0[:, 0], t1[:, 0]), np.fmin(t0[:, 1], t1[:, 1])), np.fmin(t0[:, 2], t1[:, 2]))
        tmax = np.fmin(np.fmin(np.fmax(t0[:, 0], t1[:, 0]), np.fmax(t0[:, 1], t1[:, 1])), np.fmax(t0[:, 2], t1[:, 2]))
    t = np.fmax(tmin, near).astype(np.float32)
    t_exit = np.fmin(tmax, far).astype(np.float32)
    alive = (tmin <= tmax) & (t < t_exit)
    term_keep = None
    if ray_mask is not None:
        m = np.asarray(ray_mask).astype(bool)
        term_keep = (~m, near.copy())
        alive &= m
    emitted = np.zeros(N, dtype=np.int64)
    rays, starts, ends = [], [], []
    ids = np.arange(N)
    while alive.any():
        idx = ids[alive]
        tt = t[idx]
        dt = np.minimum(np.maximum(tt * cone, step), F32(1e10)).astype(np.float32)
        tm = (tt + F32(0.5) * dt).astype(np.float32)
        go = tm < t_exit[idx]
        alive[idx[~go]] = False
        idx, tt, dt, tm = idx[go], tt[go], dt[go], tm[go]
        p = (o[idx] + tm[:, None] * d[idx]).astype(np.float32)
        occ = np.zeros(idx.shape[0], dtype=bool)
        found = np.zeros(idx.shape[0], dtype=bool)
        for l in range(L):
            inside = ~found & np.all((p >= aabbs[l, :3]) & (p < aabbs[l, 3:]), axis=1)
            if inside.any():
                u = ((p[inside] - aabbs[l, :3]) * scale[l]).astype(np.float32)
                c = np.clip(np.floor(u).astype(np.int64), 0, R - 1)
                occ[inside] = B[l, c[:, 0], c[:, 1], c[:, 2]]
            found |= inside
        rays.append(idx[occ]); starts.append(tt[occ]); ends.append((tt[occ] + dt[occ]).astype(np.float32))
        t[idx] = (tt + dt).astype(np.float32)
        if max_samples > 0:
            emitted[idx[occ]] += 1
            alive[idx[occ][emitted[idx[occ]] == max_samples]] = False
    if rays:
        rays, starts, ends = np.concatenate(rays), np.concatenate(starts), np.concatenate(ends)
        order = np.lexsort((starts, rays))
        rays, starts, ends = rays[order], starts[order], ends[order]
    else:
        rays, starts, ends = np.zeros(0, np.int64), np.zeros(0, np.float32), np.zeros(0, np.float32)
    counts = np.bincount(rays, minlength=N).astype(np.int32)
    if return_termination:
        term = t.copy()
        if term_keep is not None:
            term[term_keep[0]] = term_keep[1][term_keep[0]]
        return rays.astype(np.int64), starts, ends, counts, term
    return rays.astype(np.int64), starts, ends, counts


def render_image_with_occgrid_test(max_samples, origins, viewdirs, binaries, aabbs, params: NGPParams, near_plane=0.0,
                                   far_plane=1e10, render_step_size=1e-3, render_bkgd=None, cone_angle=0.0, alpha_thre=0.0,
                                   early_stop_eps=1e-4):
    """utils.py:175-350 restated over the oracle's marcher / field / compositing (PARITY UNPINNED like the marcher itself):
    rounds of at most n_samples = max(min(num_rays // n_alive, 64), min_samples) samples per ray, composited on top of the
    accumulated opacity, rays retired at opacity > 1 - early_stop_eps or when a round returned fewer than n_samples."""
    o, d = np.asarray(origins, dtype=np.float32), np.asarray(viewdirs, dtype=np.float32)
    N = o.shape[0]
    ot, dt_ = torch.from_numpy(o), torch.from_numpy(d)
    opacity, depth, rgb = torch.zeros(N, 1), torch.zeros(N, 1), torch.zeros(N, 3)
    ray_mask = np.ones(N, dtype=bool)
    min_samples = 1 if cone_angle == 0 else 4
    iter_samples = total = 0
    near = np.full(N, near_plane, dtype=np.float32)
    positions_all = []
    while iter_samples < max_samples:
        n_alive = int(ray_mask.sum())
        if n_alive == 0:
            break
        n_samples = max(min(N // n_alive, 64), min_samples)
        iter_samples += n_samples
        with np.errstate(all="ignore"):
            ri, ts, te, counts, term = occgrid_march(o, d, binaries, aabbs, near, far_plane, render_step_size, cone_angle,
                                                     max_samples=n_samples, ray_mask=ray_mask, return_termination=True)
        ri_t, ts_t, te_t = torch.from_numpy(ri), torch.from_numpy(ts), torch.from_numpy(te)
        pos = ot[ri_t] + dt_[ri_t] * (ts_t + te_t)[:, None] / 2.0
        positions_all.append(pos)
        if pos.shape[0]:
            rgbs, sig = ngp_forward(pos, dt_[ri_t], params)
            w, _, al = render_weight_from_density(ts_t, te_t, sig.squeeze(-1), ray_indices=ri_t, n_rays=N,
                                                  prefix_trans=1 - opacity[ri_t].squeeze(-1))
            if alpha_thre > 0:
                vis = al >= alpha_thre
                ri_t, rgbs, w, ts_t, te_t = ri_t[vis], rgbs[vis], w[vis], ts_t[vis], te_t[vis]
            rgb = rgb + accumulate_along_rays(w, rgbs, ri_t, N)
            opacity = opacity + accumulate_along_rays(w, None, ri_t, N)
            depth = depth + accumulate_along_rays(w, (ts_t + te_t)[..., None] / 2.0, ri_t, N)
            total += int(ri_t.shape[0])
        near = term
        ray_mask = (opacity.view(-1).numpy() <= 1 - early_stop_eps) & (counts == n_samples)
    bk = torch.zeros(3) if render_bkgd is None else torch.as_tensor(render_bkgd, dtype=torch.float32)
    rgb = rgb + bk * (1.0 - opacity)
    return rgb, opacity, depth, total, (torch.cat(positions_all) if positions_all else torch.zeros(0, 3))
