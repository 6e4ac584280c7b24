"""Drop-in for the reference's `examples/radiance_fields/ngp.py` hot-path classes.

`NGPRadianceField` keeps the constructor signature, buffers and call surface of ngp.py:657-809
(`normalize`, `query_density`, `_query_rgb`, `forward`) and the tinycudann parameter layout
(`mlp_base.params` = [MLP weights | grid table], `mlp_head.params`), but evaluates through libquadfield's
fused hash-grid + tensor-core MLP kernel (csrc/field.cu) — no tinycudann.  CUDA only.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Callable, List, Union

import numpy as np
import torch
from torch import nn

from .. import _lib
from . import grid as _grid


class _TruncExp(torch.autograd.Function):
    """ngp.py:146-159: forward exp(x), backward g*exp(clamp(x, max=15))."""

    @staticmethod
    def forward(ctx, x):
        x = x.float()
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(torch.clamp(x, max=15))


trunc_exp = _TruncExp.apply


class _TcnnParams(nn.Module):
    """Holds one flat fp32 `params` tensor, like a tinycudann module (state_dict key `<name>.params`)."""

    def __init__(self, n: int):
        super().__init__()
        self.params = nn.Parameter(torch.zeros(n, dtype=torch.float32))


class _DirectionEncoding(nn.Module):
    """SphericalHarmonics degree 4 (ngp.py:694-707).  A tinycudann module always registers a flat `params` Parameter, empty
    for a parameter-free encoding, so the reference's NeRF-stage state dicts carry `direction_encoding.params` of shape (0,);
    the same key is kept here so that `load_state_dict(strict=True)` works in both directions."""
    n_output_dims = 16

    def __init__(self):
        super().__init__()
        self.params = nn.Parameter(torch.zeros(0, dtype=torch.float32))


class _NGPForwardFn(torch.autograd.Function):
    """NGPRadianceField.forward under autograd: gradients for the flat tinycudann parameter tensors
    (`mlp_base.params` = [MLP weights | grid table], `mlp_head.params`) and, when `positions` requires grad, for the
    positions (tinycudann's grid input gradient; the finetune step moves the sample points with the deformation field).
    Directions get no gradient."""

    @staticmethod
    def forward(ctx, field, positions, directions, ray_indices, base_params, head_params):
        lib = _lib.load()
        h = field._native()
        pos = _lib.f32(positions.reshape(-1, 3))
        dirs = _lib.f32(directions.reshape(-1, 3))
        ri = _lib.i64(ray_indices) if ray_indices is not None else None
        M = pos.shape[0]
        rgb = torch.empty((M, 3), dtype=torch.float32, device=pos.device)
        density = torch.empty((M, 1), dtype=torch.float32, device=pos.device)
        _lib.check(lib.qf_ngp_forward(h, _lib.ptr(pos), _lib.ptr(dirs), _lib.ptr(ri), M, _lib.ptr(rgb), _lib.ptr(density),
                                      _lib.stream(pos.device)), "qf_ngp_forward")
        ctx.field = field
        ctx.pos_shape = positions.shape
        ctx.save_for_backward(pos, dirs, ri)
        return rgb, density

    @staticmethod
    def backward(ctx, g_rgb, g_density):
        lib = _lib.load()
        field = ctx.field
        pos, dirs, ri = ctx.saved_tensors
        M = pos.shape[0]
        dev = pos.device
        h = field._native()
        # `accumulate_grad_in_place`: the kernels add into `.grad` directly (they accumulate anyway: atomics into the table
        # gradient, `+=` for the matrices) instead of into fresh zero buffers that autograd then adds to `.grad` — saves a
        # 50 MB memset and a 150 MB read-modify-write per step at T=2^19; needs preallocated `.grad` tensors
        pb, ph = field.mlp_base.params, field.mlp_head.params
        inplace = (getattr(field, "accumulate_grad_in_place", False) and pb.grad is not None and ph.grad is not None
                   and pb.grad.is_contiguous() and ph.grad.is_contiguous())
        g_base = pb.grad if inplace else torch.zeros_like(pb)
        g_head = ph.grad if inplace else torch.zeros_like(ph)
        g_pos = torch.zeros((M, 3), dtype=torch.float32, device=dev) if ctx.needs_input_grad[1] else None
        if M:
            g_rgb = _lib.f32(g_rgb) if g_rgb is not None else torch.zeros((M, 3), device=dev)
            g_den = _lib.f32(g_density.reshape(-1)) if g_density is not None else None
            nb = field._n_base
            ws = _lib.workspace(dev, lib.qf_ngp_backward_workspace_bytes(M), "ngp_bwd")
            _lib.check(lib.qf_ngp_backward_inputs(h, _lib.ptr(pos), _lib.ptr(dirs), _lib.ptr(ri), M, _lib.ptr(g_rgb), _lib.ptr(g_den),
                                                  _lib.ptr(g_base[nb:]), _lib.ptr(g_base[:nb]), _lib.ptr(g_head), _lib.ptr(g_pos),
                                                  _lib.ptr(ws), ws.numel(), _lib.stream(dev)), "qf_ngp_backward_inputs")
        if g_pos is not None:
            g_pos = g_pos.view(ctx.pos_shape)
        if inplace:
            return None, g_pos, None, None, None, None
        return None, g_pos, None, None, g_base, g_head


class _NGPDensityFeatFn(torch.autograd.Function):
    """`query_density(x, return_feat=True)` under autograd: gradients of sigma and of the 15 geo features reach the flat
    `mlp_base.params` tensor ([MLP weights | grid table]) and, when requested, the positions."""

    @staticmethod
    def forward(ctx, field, positions, base_params):
        lib = _lib.load()
        h = field._native()
        pos = _lib.f32(positions.detach().reshape(-1, 3))
        M = pos.shape[0]
        density = torch.empty((M, 1), dtype=torch.float32, device=pos.device)
        feat = torch.empty((M, 15), dtype=torch.float32, device=pos.device)
        _lib.check(lib.qf_ngp_query_density(h, _lib.ptr(pos), M, _lib.ptr(density), _lib.ptr(feat), _lib.stream(pos.device)),
                   "qf_ngp_query_density")
        ctx.field, ctx.pos_shape = field, positions.shape
        ctx.save_for_backward(pos)
        return density, feat

    @staticmethod
    def backward(ctx, g_density, g_feat):
        lib = _lib.load()
        field = ctx.field
        (pos,) = ctx.saved_tensors
        M, dev = pos.shape[0], pos.device
        h = field._native()
        g_base = torch.zeros_like(field.mlp_base.params)
        g_pos = torch.zeros((M, 3), dtype=torch.float32, device=dev) if ctx.needs_input_grad[1] else None
        if M:
            gf = _lib.f32(g_feat) if g_feat is not None else torch.zeros((M, 15), device=dev)
            gd = _lib.f32(g_density.reshape(-1)) if g_density is not None else None
            nb = field._n_base
            ws = _lib.workspace(dev, lib.qf_ngp_backward_workspace_bytes(M), "ngp_bwd")
            _lib.check(lib.qf_ngp_backward_features(h, _lib.ptr(pos), M, _lib.ptr(gd), _lib.ptr(gf), _lib.ptr(g_base[nb:]),
                                                    _lib.ptr(g_base[:nb]), _lib.ptr(g_pos), _lib.ptr(ws), ws.numel(),
                                                    _lib.stream(dev)), "qf_ngp_backward_features")
        if g_pos is not None:
            g_pos = g_pos.view(ctx.pos_shape)
        return None, g_pos, g_base


_BASE_SHAPES = [(64, 32), (16, 64)]
_HEAD_SHAPES = [(64, 32), (64, 64), (16, 64)]


def _xavier_flat(shapes, gen):
    out = []
    for o, i in shapes:
        b = math.sqrt(6.0 / (i + o))
        out.append((torch.rand(o * i, generator=gen) * 2 - 1) * b)
    return torch.cat(out)


class NGPRadianceField(nn.Module):
    """Instant-NGP radiance field (ngp.py:657-809)."""

    def __init__(self, aabb: Union[torch.Tensor, List[float]], num_dim: int = 3, use_viewdirs: bool = True,
                 density_activation: Callable = lambda x: trunc_exp(x - 1), unbounded: bool = False,
                 base_resolution: int = 16, max_resolution: int = 4096, geo_feat_dim: int = 15, n_levels: int = 16,
                 log2_hashmap_size: int = 19, num_layers=2, hidden_size=64, seed: int = 1337) -> None:
        super().__init__()
        if not isinstance(aabb, torch.Tensor):
            aabb = torch.tensor(aabb, dtype=torch.float32)
        self.register_buffer("aabb", aabb)
        if num_dim != 3 or not use_viewdirs or unbounded or geo_feat_dim != 15 or n_levels != 16 or hidden_size != 64:
            raise NotImplementedError("the fused kernel covers the configuration every Quadfield script uses: 3-D, "
                                      "viewdirs, bounded, 16 levels x 2 features, 64-wide MLPs, 15 geo features")
        self.num_dim, self.use_viewdirs, self.density_activation, self.unbounded = num_dim, use_viewdirs, density_activation, unbounded
        self.base_resolution, self.max_resolution, self.geo_feat_dim = base_resolution, max_resolution, geo_feat_dim
        self.n_levels, self.log2_hashmap_size = n_levels, log2_hashmap_size
        # num_layers is ignored by the reference as well (quirk Q8, ngp.py:744)
        self._desc = _grid.make_grid_desc(aabb.tolist(), n_levels, base_resolution, max_resolution, log2_hashmap_size)
        self._n_entries = _grid.n_entries(self._desc)
        self._n_base = sum(o * i for o, i in _BASE_SHAPES)
        self.direction_encoding = _DirectionEncoding()
        self.mlp_base = _TcnnParams(self._n_base + 2 * self._n_entries)
        self.mlp_head = _TcnnParams(sum(o * i for o, i in _HEAD_SHAPES))
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            self.mlp_base.params[: self._n_base] = _xavier_flat(_BASE_SHAPES, g)
            self.mlp_base.params[self._n_base:] = (torch.rand(2 * self._n_entries, generator=g) * 2 - 1) * 1e-4
            self.mlp_head.params.copy_(_xavier_flat(_HEAD_SHAPES, g))
        self._handle = None
        self._handle_key = None

    # ---- native handle management -------------------------------------------------------------
    def _native(self):
        p_base, p_head = self.mlp_base.params, self.mlp_head.params
        if not p_base.is_cuda:
            raise RuntimeError("NGPRadianceField runs on CUDA only: call .to('cuda') first (no CPU path)")
        key = (p_base.data_ptr(), p_head.data_ptr(), p_base._version, p_head._version, self.aabb._version)
        if self._handle is not None and key == self._handle_key:
            return self._handle
        lib = _lib.load()
        aabb_key = (self.aabb.data_ptr(), self.aabb._version)
        if getattr(self, "_aabb_key", None) != aabb_key:      # a device->host copy synchronises: only when the box changed
            aabb = self.aabb.detach().cpu().tolist()
            for i in range(6):
                self._desc.aabb[i] = float(aabb[i])
            self._aabb_key = aabb_key
        base = p_base.detach()
        table, base_w, head_w = base[self._n_base:], base[: self._n_base], p_head.detach()
        st = _lib.stream(p_base.device)
        if self._handle is None or key[:2] != self._handle_key[:2]:
            self._free()
            h = C.c_void_p()
            _lib.check(lib.qf_ngp_create(C.byref(self._desc), _lib.ptr(table), self._n_entries, _lib.ptr(base_w),
                                         _lib.ptr(head_w), st, C.byref(h)), "qf_ngp_create")
            self._handle = h
        else:
            _lib.check(lib.qf_ngp_update(self._handle, _lib.ptr(table), _lib.ptr(base_w), _lib.ptr(head_w), st), "qf_ngp_update")
        self._handle_key = key
        return self._handle

    def mark_parameters_changed(self):
        """The parameters were modified behind Python's tensor version counters (a CUDA-graph replay of the optimiser step):
        the next use refreshes the fp16 working copies."""
        if self._handle_key is not None:
            self._handle_key = self._handle_key[:2] + (-1, -1) + self._handle_key[4:]

    def _free(self):
        if getattr(self, "_handle", None) is not None:
            _lib.load().qf_ngp_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass

    # ---- reference call surface ----------------------------------------------------------------
    def normalize(self, x):
        """ngp.py:748-755."""
        aabb_min, aabb_max = torch.split(self.aabb, self.num_dim, dim=-1)
        x = (x - aabb_min) / (aabb_max - aabb_min)
        selector = ((x > 0.0) & (x < 1.0)).all(dim=-1)
        return selector, x

    @torch.no_grad()
    def query_density(self, x, return_feat: bool = False):
        """ngp.py:757-779 -> density (...,1) [, feat (...,15)]."""
        lib = _lib.load()
        h = self._native()
        shape = list(x.shape[:-1])
        pos = _lib.f32(x.reshape(-1, 3))
        M = pos.shape[0]
        density = torch.empty((M, 1), dtype=torch.float32, device=pos.device)
        feat = torch.empty((M, 15), dtype=torch.float32, device=pos.device) if return_feat else None
        _lib.check(lib.qf_ngp_query_density(h, _lib.ptr(pos), M, _lib.ptr(density), _lib.ptr(feat),
                                            _lib.stream(pos.device)), "qf_ngp_query_density")
        density = density.view(shape + [1])
        if return_feat:
            return density, feat.view(shape + [15])
        return density

    def forward(self, positions: torch.Tensor, directions: torch.Tensor = None, ray_indices: torch.Tensor = None):
        """ngp.py:798-809 -> (rgb (M,3), density (M,1)).  `ray_indices` (extension): gather
        directions[ray_indices] inside the kernel, as utils.py:515-529 does with a separate indexing op.
        Differentiable w.r.t. the hash table and MLP weights (training mode)."""
        if not (self.use_viewdirs and (directions is not None)):
            raise NameError("name 'rgb' is not defined")  # what the reference does (quirk Q8, ngp.py:803-809)
        if ray_indices is None:
            assert positions.shape == directions.shape, f"{positions.shape} v.s. {directions.shape}"
        if torch.is_grad_enabled() and (self.mlp_base.params.requires_grad or self.mlp_head.params.requires_grad
                                        or positions.requires_grad):
            return _NGPForwardFn.apply(self, positions, directions, ray_indices, self.mlp_base.params, self.mlp_head.params)
        with torch.no_grad():
            return self._forward_nograd(positions, directions, ray_indices)

    def _forward_nograd(self, positions, directions, ray_indices):
        lib = _lib.load()
        h = self._native()
        pos = _lib.f32(positions.reshape(-1, 3))
        dirs = _lib.f32(directions.reshape(-1, 3))
        ri = _lib.i64(ray_indices) if ray_indices is not None else None
        M = pos.shape[0]
        rgb = torch.empty((M, 3), dtype=torch.float32, device=pos.device)
        density = torch.empty((M, 1), dtype=torch.float32, device=pos.device)
        _lib.check(lib.qf_ngp_forward(h, _lib.ptr(pos), _lib.ptr(dirs), _lib.ptr(ri), M, _lib.ptr(rgb), _lib.ptr(density),
                                      _lib.stream(pos.device)), "qf_ngp_forward")
        return rgb, density

    @torch.no_grad()
    def encode_backward(self, x01: torch.Tensor, grad_enc: torch.Tensor) -> torch.Tensor:
        """tcnn HashGrid backward alone: dL/dtable (n_entries, 2) for dL/denc (M, 32) at x in [0,1]^3."""
        lib = _lib.load()
        h = self._native()
        x01 = _lib.f32(x01.reshape(-1, 3))
        g = _lib.f32(grad_enc.reshape(x01.shape[0], 2 * self.n_levels))
        out = torch.zeros((self._n_entries, 2), dtype=torch.float32, device=x01.device)
        _lib.check(lib.qf_hashgrid_backward(h, _lib.ptr(x01), _lib.ptr(g), x01.shape[0], _lib.ptr(out), _lib.stream(x01.device)),
                   "qf_hashgrid_backward")
        return out

    @torch.no_grad()
    def encode(self, x01: torch.Tensor) -> torch.Tensor:
        """tcnn HashGrid forward alone (for tests / profiling): x in [0,1]^3 -> (M, 32)."""
        lib = _lib.load()
        h = self._native()
        x01 = _lib.f32(x01.reshape(-1, 3))
        out = torch.empty((x01.shape[0], 2 * self.n_levels), dtype=torch.float32, device=x01.device)
        _lib.check(lib.qf_hashgrid_forward(h, _lib.ptr(x01), x01.shape[0], _lib.ptr(out), _lib.stream(x01.device)),
                   "qf_hashgrid_forward")
        return out

    # ---- parameter helpers ------------------------------------------------------------------------
    @torch.no_grad()
    def load_arrays(self, table: torch.Tensor, base_w, head_w):
        """Set parameters from explicit arrays: table (n_entries,2), base_w [(64,32),(16,64)], head_w [(64,32),(64,64),(16,64)]."""
        dev = self.mlp_base.params.device
        self.mlp_base.params.copy_(torch.cat([w.reshape(-1) for w in base_w] + [table.reshape(-1)]).to(dev))
        self.mlp_head.params.copy_(torch.cat([w.reshape(-1) for w in head_w]).to(dev))


class _SGMixtureFn(torch.autograd.Function):
    """rgb = sigmoid(diffuse + sum_l c_l exp(|lambda_l| (a_l/|a_l| . d - 1))), differentiable in the feature rows."""

    @staticmethod
    def forward(ctx, features, dirs, num_lobes):
        lib = _lib.load()
        f = _lib.f32(features.detach())
        d = _lib.f32(dirs.detach())
        M = f.shape[0]
        rgb = torch.empty((M, 3), dtype=torch.float32, device=f.device)
        _lib.check(lib.qf_sg_features_to_rgb(_lib.ptr(f), f.shape[1], num_lobes, _lib.ptr(d), M, _lib.ptr(rgb),
                                             _lib.stream(f.device)), "qf_sg_features_to_rgb")
        ctx.save_for_backward(f, d)
        ctx.num_lobes = num_lobes
        return rgb

    @staticmethod
    def backward(ctx, g_rgb):
        lib = _lib.load()
        f, d = ctx.saved_tensors
        M = f.shape[0]
        g_f = torch.zeros_like(f)                       # columns beyond 3+7L (e.g. a trailing density) get no gradient
        if M:
            _lib.check(lib.qf_sg_features_to_rgb_backward(_lib.ptr(f), f.shape[1], ctx.num_lobes, _lib.ptr(d), M,
                                                          _lib.ptr(_lib.f32(g_rgb)), _lib.ptr(g_f), g_f.shape[1],
                                                          _lib.stream(f.device)), "qf_sg_features_to_rgb_backward")
        return g_f, None, None


def spherical_gaussian_features_to_rgb(features: torch.Tensor, dirs: torch.Tensor, num_lobes: int) -> torch.Tensor:
    """`NGPRadianceFieldSGNew.features_to_rgb` (ngp.py:456-461, 371-393) with discretize=False; differentiable with
    respect to `features`."""
    return _SGMixtureFn.apply(features, dirs, num_lobes)


class _DecoderHead(nn.Module):
    """Parameter layout of the reference's `BasicDecoder` (ngp.py:35-143): `layers.{i}` hidden `nn.Linear` + `lout`."""

    def __init__(self, input_dim: int, output_dim: int, num_layers: int, hidden_dim: int, bias: bool = True):
        super().__init__()
        self.layers = nn.ModuleList([nn.Linear(input_dim if i == 0 else hidden_dim, hidden_dim, bias=bias) for i in range(num_layers)])
        self.lout = nn.Linear(hidden_dim, output_dim, bias=True)

    def forward(self, x):
        for layer in self.layers:
            x = torch.relu(layer(x))
        return self.lout(x)


class NGPRadianceFieldSGNew(NGPRadianceField):
    """The spherical-Gaussian radiance field (ngp.py:284-470) as every Quadfield script builds it (`use_viewdirs=False`):
    hash grid + 64-wide base MLP -> (sigma, 15 features) on the fused kernel, a torch fp32 `BasicDecoder` head
    15 -> 64 x num_layers -> 3 + 7 L (plain library GEMMs, as in the reference), and the SG mixture
    rgb = sigmoid(diffuse + sum_l c_l exp(|lambda_l| (a_l . d - 1))) on `qf_sg_features_to_rgb`.
    State-dict keys: `aabb`, `mlp_base.params`, `mlp_head.layers.{i}.weight/bias`, `mlp_head.lout.weight/bias`.
    Trainable (train_fit_sg.py:439-452): with autograd enabled the gradient flows through the SG mixture
    (`qf_sg_features_to_rgb_backward`), the torch decoder and the geo features into the base MLP and the hash table
    (`qf_ngp_backward_features`)."""

    def __init__(self, aabb=None, num_dim: int = 3, use_viewdirs: bool = False, density_activation=None, unbounded: bool = False,
                 base_resolution: int = 16, max_resolution: int = 4096, geo_feat_dim: int = 15, n_levels: int = 16,
                 log2_hashmap_size: int = 19, num_g_lobes: int = 3, hidden_size: int = 64, num_layers: int = 2,
                 output_activation: str = "sigmoid", discretize: bool = False, seed: int = 1337) -> None:
        if use_viewdirs:
            raise NotImplementedError("every Quadfield script builds the SG field with use_viewdirs=False")
        if discretize:
            raise NotImplementedError("discretize=True (fake quantisation during SG fitting) is outside the render path")
        if aabb is None:
            aabb = [-1.5, -1.5, -1.5, 1.5, 1.5, 1.5]
        super().__init__(aabb=aabb, num_dim=num_dim, use_viewdirs=True, unbounded=unbounded, base_resolution=base_resolution,
                         max_resolution=max_resolution, geo_feat_dim=geo_feat_dim, n_levels=n_levels,
                         log2_hashmap_size=log2_hashmap_size, seed=seed)
        self.use_viewdirs = False
        self.num_g_lobes, self.discretize, self.output_activation = num_g_lobes, discretize, output_activation
        # the tcnn head of the parent is replaced by the reference's torch decoder; the fused kernel only needs the base
        # matrices, so the native handle is fed zeros for the (unused) tcnn head image
        del self.mlp_head
        del self.direction_encoding
        torch.manual_seed(seed)
        self.mlp_head = _DecoderHead(geo_feat_dim, 3 + num_g_lobes * 7, num_layers, hidden_size)
        self.register_buffer("_zero_head", torch.zeros(sum(o * i for o, i in _HEAD_SHAPES)), persistent=False)

    def _native(self):
        p_base = self.mlp_base.params
        if not p_base.is_cuda:
            raise RuntimeError("NGPRadianceFieldSGNew runs on CUDA only: call .to('cuda') first (no CPU path)")
        key = (p_base.data_ptr(), p_base._version, self.aabb._version)
        if self._handle is not None and key == self._handle_key:
            return self._handle
        lib = _lib.load()
        aabb = self.aabb.detach().cpu().tolist()
        for i in range(6):
            self._desc.aabb[i] = float(aabb[i])
        base = p_base.detach()
        table, base_w = base[self._n_base:], base[: self._n_base]
        st = _lib.stream(p_base.device)
        if self._handle is None or key[0] != self._handle_key[0]:
            self._free()
            h = C.c_void_p()
            _lib.check(lib.qf_ngp_create(C.byref(self._desc), _lib.ptr(table), self._n_entries, _lib.ptr(base_w),
                                         _lib.ptr(self._zero_head), st, C.byref(h)), "qf_ngp_create")
            self._handle = h
        else:
            _lib.check(lib.qf_ngp_update(self._handle, _lib.ptr(table), _lib.ptr(base_w), _lib.ptr(self._zero_head), st), "qf_ngp_update")
        self._handle_key = key
        return self._handle

    def query_density(self, x, return_feat: bool = False):
        """ngp.py:402-428; differentiable when autograd is on and the parameters (or x) require grad."""
        if torch.is_grad_enabled() and (self.mlp_base.params.requires_grad or x.requires_grad):
            shape = list(x.shape[:-1])
            density, feat = _NGPDensityFeatFn.apply(self, x, self.mlp_base.params)
            density, feat = density.view(shape + [1]), feat.view(shape + [15])
            return (density, feat) if return_feat else density
        return super().query_density(x, return_feat=return_feat)

    def features(self, x):
        """ngp.py:445-454 -> (M, 3 + 7 L + 1): decoder output, then the density."""
        density, embedding = self.query_density(x, return_feat=True)
        feats = self.mlp_head(embedding.reshape(-1, self.geo_feat_dim))
        return torch.cat([feats.reshape(list(embedding.shape[:-1]) + [self.num_g_lobes * 7 + 3]), density], dim=-1)

    def features_to_rgb(self, features, dir):
        """ngp.py:456-461 (discretize=False)."""
        return spherical_gaussian_features_to_rgb(features, dir, self.num_g_lobes)

    def forward(self, positions: torch.Tensor, directions: torch.Tensor = None, ray_indices: torch.Tensor = None):
        """ngp.py:463-470 -> (rgb (M,3), density (M,1)).  `ray_indices` (extension): directions are per ray."""
        density, embedding = self.query_density(positions, return_feat=True)
        feats = self.mlp_head(embedding.reshape(-1, self.geo_feat_dim))
        d = directions.reshape(-1, 3)
        if ray_indices is not None:
            d = d[ray_indices]
        rgb = spherical_gaussian_features_to_rgb(feats, d, self.num_g_lobes)
        return rgb, density
