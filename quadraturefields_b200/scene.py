"""Synthetic scenes of BASELINE.json / SURVEY.md §8(d): jittered concentric icosphere "quadrature meshes", a
random-init NGP field, NeRF-synthetic pinhole cameras on a sphere, and a random baked texture set.
Used by bench.py, the parity tests and smoke(); holds everything both as numpy/CPU arrays (so a checker
can be fed the very same inputs) and as resident device objects."""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch

from .datasets import ray_gen
from .mesh_utils import MeshIntersection
from .radiance_fields.ngp import NGPRadianceField, NGPRadianceFieldSGNew
from .texture_utils import FeatureCompression
from .utils import MeshRenderer

CAMERA_ANGLE_X = 0.6911112070083618  # NeRF-synthetic transforms_*.json


def icosphere(subdivisions: int):
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11],
                  [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    for _ in range(subdivisions):
        nv = v.shape[0]
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
        e.sort(axis=1)
        ue, inv = np.unique(e, axis=0, return_inverse=True)
        mid = v[ue[:, 0]] + v[ue[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        v = np.concatenate([v, mid])
        F = f.shape[0]
        ab, bc, ca = nv + inv[:F], nv + inv[F:2 * F], nv + inv[2 * F:]
        a, b, c = f[:, 0], f[:, 1], f[:, 2]
        f = np.concatenate([np.stack([a, ab, ca], 1), np.stack([b, bc, ab], 1), np.stack([c, ca, bc], 1),
                            np.stack([ab, bc, ca], 1)])
    return v, f


def shell_mesh(radii, subdivisions: int, jitter: float = 1e-3, seed: int = 42):
    rng = np.random.RandomState(seed)
    v0, f0 = icosphere(subdivisions)
    vs, fs, base = [], [], 0
    for r in radii:
        vs.append(v0 * r + rng.normal(0.0, jitter, size=v0.shape))
        fs.append(f0 + base)
        base += v0.shape[0]
    return np.concatenate(vs).astype(np.float32), np.concatenate(fs).astype(np.int32)


def look_at_c2w(eye, target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0)) -> np.ndarray:
    eye = np.asarray(eye, dtype=np.float64)
    f = np.asarray(target, dtype=np.float64) - eye
    f /= np.linalg.norm(f)
    r = np.cross(f, np.asarray(up, dtype=np.float64))
    r /= np.linalg.norm(r)
    u = np.cross(r, f)
    m = np.zeros((3, 4), dtype=np.float64)
    m[:, 0], m[:, 1], m[:, 2], m[:, 3] = r, u, -f, eye
    return m.astype(np.float32)


def spiral_poses(n: int, radius: float = 4.03) -> np.ndarray:
    """n cameras on a sphere of `radius` looking at the origin (NeRF-synthetic test-set shaped)."""
    out = []
    for i in range(n):
        th = 2 * math.pi * i / max(n, 1)
        ph = math.radians(30.0) + math.radians(20.0) * math.sin(2 * th)
        out.append(look_at_c2w((radius * math.cos(th) * math.cos(ph), radius * math.sin(th) * math.cos(ph), radius * math.sin(ph))))
    return np.stack(out)


def random_field_params(seed: int, n_entries: int, table_scale: float = 1e4, density_gain: float = 8.0,
                        rgb_gain: float = 4.0):
    """Random-init field of SURVEY §8d: table U(-1e-4,1e-4)*table_scale, Xavier-uniform bias-free MLPs, all rounded to
    fp16-representable values.  The density row is made positive and scaled (`density_gain`) and the colour rows
    scaled (`rgb_gain`) so opacity spans (0,1) and colours vary — raw init renders a near-constant image."""
    g = torch.Generator().manual_seed(seed)
    table = ((torch.rand((n_entries, 2), generator=g) * 2 - 1) * 1e-4 * table_scale)

    def xavier(o, i):
        b = math.sqrt(6.0 / (i + o))
        return (torch.rand((o, i), generator=g) * 2 - 1) * b

    base_w = [xavier(64, 32), xavier(16, 64)]
    head_w = [xavier(64, 32), xavier(64, 64), xavier(16, 64)]
    base_w[1][0] = base_w[1][0].abs() * density_gain
    head_w[2][:3] = head_w[2][:3] * rgb_gain
    r = lambda t: t.half().float()
    return r(table), [r(w) for w in base_w], [r(w) for w in head_w]


def random_uv(n_vertices: int, seed: int = 7) -> np.ndarray:
    """A synthetic per-vertex atlas in [0,1)^2 (the real one comes from xatlas, out of scope)."""
    rng = np.random.RandomState(seed)
    return rng.uniform(0.02, 0.98, size=(n_vertices, 2)).astype(np.float32)


def scale_uv(uv: np.ndarray, size: int) -> torch.Tensor:
    """test_baking_texture_images.py:325-328."""
    uv = np.asarray(uv) - 1e-7
    uv = np.array(uv).astype(np.float32) * size
    return torch.from_numpy(np.clip(uv, 0, size - 1))


CONFIGS = {
    # name: radii, subdivisions, W, H, K, log2_T, n_views, baked(lobes, texture size)
    "smoke": dict(radii=[0.5, 0.8, 1.0], sub=2, W=48, H=48, K=8, log2_T=14, views=2),
    "c1": dict(radii=[0.4, 0.6, 0.8, 1.0], sub=4, W=100, H=100, K=8, log2_T=19, views=1, cam_radius=4.0),
    "c2": dict(radii=[0.4, 0.6, 0.8, 1.0], sub=4, W=800, H=800, K=8, log2_T=19, views=200),
    "c4": dict(radii=[0.30 + 0.05 * i for i in range(14)], sub=6, W=1920, H=1080, K=32, log2_T=21, views=8),
    "c5": dict(radii=[0.30 + 0.05 * i for i in range(14)], sub=6, W=3840, H=2160, K=32, log2_T=14, views=4,
               lobes=3, tex=8192, lambda_thres=5.0),
    "c5_small": dict(radii=[0.5, 0.8, 1.0], sub=3, W=160, H=90, K=8, log2_T=14, views=2, lobes=3, tex=256,
                     lambda_thres=5.0),
}


@dataclass
class Scene:
    name: str
    cfg: dict
    vertices_np: np.ndarray
    faces_np: np.ndarray
    poses: np.ndarray
    K: int
    W: int
    H: int
    focal: float
    cx: float
    cy: float
    table: torch.Tensor
    base_w: List[torch.Tensor]
    head_w: List[torch.Tensor]
    aabb: List[float]
    log2_T: int
    device: torch.device
    mesh_intersect: Optional[MeshIntersection] = None
    radiance_field: Optional[NGPRadianceField] = None
    renderer: Optional[MeshRenderer] = None
    baked_renderer: Optional[MeshRenderer] = None
    compressor: Optional[FeatureCompression] = None
    uv_scaled: Optional[torch.Tensor] = None
    planes: Optional[dict] = None
    extras: dict = field(default_factory=dict)

    @property
    def n_rays(self) -> int:
        return self.W * self.H

    def rays(self, view: int):
        r = ray_gen.generate_rays(self.poses[view % len(self.poses)], self.W, self.H, self.focal, self.cx, self.cy,
                                  device=self.device)
        return r.origins, r.viewdirs

    def render(self, origins, viewdirs, bg_color="white", render_bkgd=None, out=None, hits_out=None, image_width=None):
        return self.renderer.render(origins, viewdirs, bg_color=bg_color, render_bkgd=render_bkgd, out=out, hits_out=hits_out,
                                    image_width=image_width)

    def render_baked(self, origins, viewdirs, bg_color="white", out=None, image_width=None):
        return self.baked_renderer.render(origins, viewdirs, bg_color=bg_color, out=out, image_width=image_width)


def make_scene(name: str = "c2", device="cuda", seed: int = 42, build_field: bool = True, **overrides) -> Scene:
    cfg = dict(CONFIGS[name])
    cfg.update(overrides)
    dev = torch.device(device)
    vertices, faces = shell_mesh(cfg["radii"], cfg["sub"], jitter=1e-3, seed=seed)
    focal, cx, cy, W, H = ray_gen.intrinsics(cfg["W"], cfg["H"], CAMERA_ANGLE_X)
    poses = spiral_poses(cfg["views"], cfg.get("cam_radius", 4.03))
    aabb = [-1.5, -1.5, -1.5, 1.5, 1.5, 1.5]
    sc = Scene(name=name, cfg=cfg, vertices_np=vertices, faces_np=faces, poses=poses, K=cfg["K"], W=W, H=H, focal=focal,
               cx=cx, cy=cy, table=None, base_w=None, head_w=None, aabb=aabb, log2_T=cfg["log2_T"], device=dev)
    sc.mesh_intersect = MeshIntersection((vertices, faces), simplify_mesh=False, scale=1.0, num_intersections=cfg["K"],
                                         render_step_size=0.005, device=dev)
    if build_field:
        rf = NGPRadianceField(aabb=aabb, log2_hashmap_size=cfg["log2_T"])
        sc.table, sc.base_w, sc.head_w = random_field_params(seed, rf._n_entries)
        rf.load_arrays(sc.table, sc.base_w, sc.head_w)
        sc.radiance_field = rf.to(dev)
        sc.renderer = MeshRenderer(sc.mesh_intersect, radiance_field=sc.radiance_field)
    if "lobes" in cfg:
        L, S = cfg["lobes"], cfg["tex"]
        g = torch.Generator().manual_seed(seed + 1)
        r = lambda *s: torch.randint(0, 256, s, generator=g, dtype=torch.int32).to(torch.uint8)
        sc.planes = dict(alpha=r(S, S), diffuse=r(S, S, 3), sg_colors=[r(S, S, 3) for _ in range(L)],
                         lambdas=[r(S, S, 3) for _ in range(L)])
        sc.compressor = FeatureCompression(L, planes=sc.planes, compression_type="linear",
                                           lambda_thres=cfg.get("lambda_thres", 7.5), device=dev)
        sc.extras["uv"] = random_uv(vertices.shape[0], seed + 2)
        sc.uv_scaled = scale_uv(sc.extras["uv"], S)
        sc.extras["sg_field"] = NGPRadianceFieldSGNew(num_g_lobes=L, log2_hashmap_size=12)
        sc.baked_renderer = MeshRenderer(sc.mesh_intersect, compressor=sc.compressor, uv=sc.uv_scaled)
    return sc
