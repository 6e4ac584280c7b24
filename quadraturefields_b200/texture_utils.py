"""Drop-in for the decode half of the reference's `examples/texture_utils.py` (`FeatureCompression`)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from .radiance_fields.ngp import spherical_gaussian_features_to_rgb


class FeatureCompression:
    """texture_utils.py:17-49: the uint8 atlas set (alpha, diffuse, per-lobe colour and [lambda, azimuth, elevation]).

    `path` loads `alpha.png, diffuse.png, color_i.png, lambda_axis_i.png` with PIL (the reference uses imageio);
    `planes=dict(alpha=, diffuse=, sg_colors=[...], lambdas=[...])` takes uint8 arrays directly."""

    def __init__(self, num_lobes, initialize=False, texture_size=None, path=None, compression_type="sigmoid",
                 lambda_thres=7.5, planes: Optional[Dict] = None, device="cuda"):
        self.num_lobes = num_lobes
        self.texture_size = texture_size
        self.compression_type = compression_type
        self.lambda_thres = lambda_thres
        self.device = torch.device(device)
        if initialize:
            S = texture_size
            z = lambda *s: torch.zeros(s, dtype=torch.uint8, device=self.device)
            self.alpha, self.diffuse = z(S, S), z(S, S, 3)
            self.sg_colors = {i: z(S, S, 3) for i in range(num_lobes)}
            self.lambdas = {i: z(S, S, 3) for i in range(num_lobes)}
        elif planes is not None:
            t = lambda a: torch.as_tensor(np.asarray(a) if not isinstance(a, torch.Tensor) else a).to(self.device, torch.uint8).contiguous()
            self.alpha, self.diffuse = t(planes["alpha"]), t(planes["diffuse"])
            self.sg_colors = {i: t(planes["sg_colors"][i]) for i in range(num_lobes)}
            self.lambdas = {i: t(planes["lambdas"][i]) for i in range(num_lobes)}
        else:
            from PIL import Image
            Image.MAX_IMAGE_PIXELS = 1000000000
            rd = lambda n: torch.from_numpy(np.array(Image.open(path + n))).to(self.device).contiguous()
            self.alpha, self.diffuse = rd("alpha.png"), rd("diffuse.png")
            self.sg_colors = {i: rd("color_{}.png".format(i)) for i in range(num_lobes)}
            self.lambdas = {i: rd("lambda_axis_{}.png".format(i)) for i in range(num_lobes)}
        self.texture_size = self.alpha.shape[0]
        self._handle = None

    def native(self):
        """Interleaved texel records on the device (csrc/baked.cu); rebuilt by `repack()` after the planes change."""
        if self._handle is None:
            lib = _lib.load()
            L = self.num_lobes
            cols = (C.c_void_p * max(L, 1))(*[self.sg_colors[i].data_ptr() for i in range(L)])
            lams = (C.c_void_p * max(L, 1))(*[self.lambdas[i].data_ptr() for i in range(L)])
            h = C.c_void_p()
            _lib.check(lib.qf_texture_create(self.texture_size, L, _lib.ptr(self.alpha), _lib.ptr(self.diffuse), cols, lams,
                                             1 if self.compression_type == "sigma" else 0,     # quirk Q5
                                             float(self.lambda_thres), _lib.stream(self.device), C.byref(h)),
                       "qf_texture_create")
            self._handle = h
        return self._handle

    def repack(self):
        if self._handle is not None:
            _lib.load().qf_texture_destroy(self._handle)
            self._handle = None
        return self.native()

    @torch.no_grad()
    def get_features_from_texture_map(self, indices):
        """texture_utils.py:149-175: (M,2) long texel indices -> (M, 3+7L+1) fp32 [diffuse, L x (axis, lambda, c), sigma]."""
        lib = _lib.load()
        idx = _lib.i64(indices.to(self.device))
        M = idx.shape[0]
        out = torch.empty((M, 3 + 7 * self.num_lobes + 1), dtype=torch.float32, device=self.device)
        _lib.check(lib.qf_texture_decode(self.native(), _lib.ptr(idx), M, _lib.ptr(out), _lib.stream(self.device)),
                   "qf_texture_decode")
        return out

    def _compress_into(self, features, indices, alpha, diffuse, colors, lambdas, size):
        lib = _lib.load()
        L = self.num_lobes
        f = _lib.f32(features.reshape(-1, 3 + 7 * L + 1), self.device)
        idx = _lib.i64(indices.to(self.device)) if indices is not None else None
        cols = (C.c_void_p * max(L, 1))(*[colors[i].data_ptr() for i in range(L)])
        lams = (C.c_void_p * max(L, 1))(*[lambdas[i].data_ptr() for i in range(L)])
        _lib.check(lib.qf_texture_compress(_lib.ptr(f), f.shape[0], L, 1 if self.compression_type == "sigma" else 0,
                                           float(self.lambda_thres), _lib.ptr(idx), int(size), _lib.ptr(alpha), _lib.ptr(diffuse),
                                           cols, lams, _lib.stream(self.device)), "qf_texture_compress")

    @torch.no_grad()
    def compress(self, features):
        """texture_utils.py:67-98: features (N, 3+7L+1) -> dict(alpha (N,), diffuse (N,3), lambdas [L x (N,3)], colors [L x (N,3)]) uint8."""
        N = features.shape[0]
        z = lambda *s: torch.empty(s, dtype=torch.uint8, device=self.device)
        data = dict(alpha=z(N), diffuse=z(N, 3), lambdas=[z(N, 3) for _ in range(self.num_lobes)],
                    colors=[z(N, 3) for _ in range(self.num_lobes)])
        self._compress_into(features, None, data["alpha"], data["diffuse"], data["colors"], data["lambdas"], 0)
        return data

    @torch.no_grad()
    def assign_values_to_texture_map(self, features, indices):
        """texture_utils.py:100-106: quantise and scatter into the atlas planes at texel `indices` (M,2)."""
        self._compress_into(features, indices, self.alpha, self.diffuse, self.sg_colors, self.lambdas, self.texture_size)
        if self._handle is not None:
            self.repack()

    load_features_into_maps = assign_values_to_texture_map            # texture_utils.py:197-203

    def save_to_file(self, path):
        """texture_utils.py:119-124 (PNG set: alpha.png, diffuse.png, color_i.png, lambda_axis_i.png)."""
        from PIL import Image
        Image.fromarray(self.alpha.cpu().numpy()).save(path + "alpha.png")
        Image.fromarray(self.diffuse.cpu().numpy()).save(path + "diffuse.png")
        for i in range(self.num_lobes):
            Image.fromarray(self.sg_colors[i].cpu().numpy()).save(path + "color_{}.png".format(i))
            Image.fromarray(self.lambdas[i].cpu().numpy()).save(path + "lambda_axis_{}.png".format(i))

    @torch.no_grad()
    def features_to_rgb(self, features, dir):
        """texture_utils.py:144-147."""
        return spherical_gaussian_features_to_rgb(features, dir, self.num_lobes)

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().qf_texture_destroy(self._handle)
                self._handle = None
        except Exception:
            pass
