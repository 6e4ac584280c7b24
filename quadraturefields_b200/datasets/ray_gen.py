"""Eval-mode pinhole ray generation of the reference loaders (examples/datasets/nerf_synthetic.py:219-226,
289-378) as one kernel: pixel (x+0.5-cx)/f, -(y+0.5-cy)/f, -1 (OpenGL) -> rotate by c2w -> unit viewdirs."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from .. import _lib
from .utils import Rays


def intrinsics(width: int, height: int, camera_angle_x: float, upsample: int = 1):
    """focal = 0.5*W/tan(0.5*camera_angle_x) (nerf_synthetic.py:101-102), scaled by `upsample` (:214-217)."""
    focal = 0.5 * width / math.tan(0.5 * camera_angle_x) * upsample
    W, H = int(width * upsample), int(height * upsample)
    return float(np.float32(focal)), float(np.float32(W / 2.0)), float(np.float32(H / 2.0)), W, H


@torch.no_grad()
def generate_rays(c2w, W: int, H: int, focal: float, cx: float, cy: float, opengl: bool = True, device="cuda") -> Rays:
    lib = _lib.load()
    dev = torch.device(device)
    m = np.ascontiguousarray(np.asarray(c2w.cpu() if isinstance(c2w, torch.Tensor) else c2w, dtype=np.float32)[:3, :4])
    origins = torch.empty((W * H, 3), dtype=torch.float32, device=dev)
    viewdirs = torch.empty((W * H, 3), dtype=torch.float32, device=dev)
    _lib.check(lib.qf_generate_rays(m.ctypes.data_as(C.POINTER(C.c_float)), W, H, focal, cx, cy, 1 if opengl else 0,
                                    _lib.ptr(origins), _lib.ptr(viewdirs), _lib.stream(dev)), "qf_generate_rays")
    return Rays(origins=origins, viewdirs=viewdirs)


@torch.no_grad()
def generate_rays_indexed(camtoworlds: torch.Tensor, image_id, x: torch.Tensor, y: torch.Tensor, focal: float, cx: float,
                          cy: float, opengl: bool = True) -> Rays:
    """Training-branch rays (nerf_synthetic.py:293-309, 341-370): ray i through pixel (x[i], y[i]) of camera
    image_id[i] (None: camera 0 of `camtoworlds`).  camtoworlds (V,3|4,4) CUDA fp32; x, y integer or float tensors."""
    lib = _lib.load()
    dev = camtoworlds.device
    m = _lib.f32(camtoworlds[:, :3, :4])
    n = x.shape[0]
    ids = _lib.i64(image_id.to(dev)) if image_id is not None else None
    xs, ys = _lib.f32(x.to(dev)), _lib.f32(y.to(dev))
    origins = torch.empty((n, 3), dtype=torch.float32, device=dev)
    viewdirs = torch.empty((n, 3), dtype=torch.float32, device=dev)
    _lib.check(lib.qf_generate_rays_indexed(_lib.ptr(m), m.shape[0], _lib.ptr(ids), _lib.ptr(xs), _lib.ptr(ys), n, focal, cx, cy,
                                            1 if opengl else 0, _lib.ptr(origins), _lib.ptr(viewdirs), _lib.stream(dev)),
               "qf_generate_rays_indexed")
    return Rays(origins=origins, viewdirs=viewdirs)
