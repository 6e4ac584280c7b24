"""Drop-in for the reference's `examples/field_rendering.py` (the nerfacc-style compositing surface).

Same function names, argument meaning, shape checks and error behaviour; the arithmetic runs in
libquadfield's warp-segmented scan kernels (csrc/composite.cu) instead of nerfacc's `pack_info` /
`exclusive_sum` / `exclusive_prod` and ATen `index_add_`.  Differentiable w.r.t. sigmas / alphas /
values like the reference (not w.r.t. t or indices, field_rendering.py:36-37).
CUDA tensors only — there is no CPU fallback.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib


# ------------------------------------------------------------------------------------------ helpers
def pack_info(ray_indices: Tensor, n_rays: Optional[int] = None) -> Tensor:
    """nerfacc `pack.pack_info` (used at field_rendering.py:201,257): (n_rays, 2) [start, count]."""
    lib = _lib.load()
    ray_indices = _lib.i64(ray_indices)
    if n_rays is None:
        n_rays = int(ray_indices.max()) + 1 if ray_indices.numel() else 0
    out = torch.empty((n_rays, 2), dtype=torch.int64, device=ray_indices.device)
    if n_rays == 0:
        return out
    nbytes = lib.qf_pack_info_workspace_bytes(n_rays)
    ws = _lib.workspace(ray_indices.device, nbytes, "pack")
    _lib.check(lib.qf_pack_info(_lib.ptr(ray_indices), ray_indices.numel(), n_rays, _lib.ptr(out), _lib.ptr(ws),
                                ws.numel(), _lib.stream(ray_indices.device)), "qf_pack_info")
    return out


def _batched_packed_info(shape, device) -> Tensor:
    n_rays, S = shape
    start = torch.arange(n_rays, device=device, dtype=torch.int64) * S
    return torch.stack([start, torch.full_like(start, S)], dim=-1).contiguous()


def _expand_ray_indices(packed_info: Tensor, n_samples: int) -> Tensor:
    cnt = packed_info[:, 1]
    return torch.repeat_interleave(torch.arange(packed_info.shape[0], device=packed_info.device), cnt,
                                   output_size=n_samples)


class _RenderWeights(torch.autograd.Function):
    """weights / trans / alphas from alphas (mode 0) or sigmas·(t_ends−t_starts) (mode 1)."""

    @staticmethod
    def forward(ctx, mode, x, t_starts, t_ends, packed_info, prefix_trans):
        lib = _lib.load()
        x = _lib.f32(x)
        n = x.numel()
        ts = _lib.f32(t_starts) if t_starts is not None else None
        te = _lib.f32(t_ends) if t_ends is not None else None
        pf = _lib.f32(prefix_trans) if prefix_trans is not None else None
        w, T = torch.empty_like(x), torch.empty_like(x)
        a = torch.empty_like(x) if mode == 1 else None
        _lib.check(lib.qf_render_weights(mode, _lib.ptr(x), _lib.ptr(ts), _lib.ptr(te), _lib.ptr(packed_info),
                                         packed_info.shape[0], n, _lib.ptr(pf), _lib.ptr(w), _lib.ptr(T), _lib.ptr(a),
                                         _lib.stream(x.device)), "qf_render_weights")
        ctx.mode = mode
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(x, ts, te, packed_info, pf)
        if mode == 1:
            return w, T, a
        return w, T

    @staticmethod
    def backward(ctx, gw, gT, ga=None):
        lib = _lib.load()
        x, ts, te, packed_info, pf = ctx.saved_tensors
        gw = _lib.f32(gw) if gw is not None else None
        gT = _lib.f32(gT) if gT is not None else None
        if gw is None and gT is None and ga is None:
            return None, None, None, None, None, None
        gin = torch.empty_like(x)
        _lib.check(lib.qf_render_weights_backward(ctx.mode, _lib.ptr(x), _lib.ptr(ts), _lib.ptr(te),
                                                  _lib.ptr(packed_info), packed_info.shape[0], x.numel(), _lib.ptr(pf),
                                                  _lib.ptr(gw), _lib.ptr(gT), _lib.ptr(gin), _lib.stream(x.device)),
                   "qf_render_weights_backward")
        if ga is not None:
            # alphas = 1 - exp(-sigma * dt) is an ordinary autograd output in the reference (field_rendering.py:257-261)
            dt = te - ts
            gin = gin + _lib.f32(ga) * dt * torch.exp(-x * dt)
        return None, gin, None, None, None, None


def _weights(mode, x, t_starts, t_ends, packed_info, ray_indices, n_rays, prefix_trans):
    shape = x.shape
    if ray_indices is not None and packed_info is None:
        packed_info = pack_info(ray_indices, n_rays)
    if packed_info is None:  # batched (n_rays, n_samples)
        assert x.dim() == 2, "batched inputs must be (n_rays, n_samples) when no packed_info/ray_indices is given"
        packed_info = _batched_packed_info(shape, x.device)
    flat = lambda t: None if t is None else t.reshape(-1)
    outs = _RenderWeights.apply(mode, flat(x), flat(t_starts), flat(t_ends), _lib.i64(packed_info), flat(prefix_trans))
    return tuple(o.reshape(shape) for o in outs)


# ------------------------------------------------------------------------------------------ public API
def render_transmittance_from_alpha(alphas: Tensor, packed_info: Optional[Tensor] = None,
                                    ray_indices: Optional[Tensor] = None, n_rays: Optional[int] = None,
                                    prefix_trans: Optional[Tensor] = None) -> Tensor:
    """field_rendering.py:161-206."""
    return _weights(0, alphas, None, None, packed_info, ray_indices, n_rays, prefix_trans)[1]


def render_transmittance_from_density(t_starts: Tensor, t_ends: Tensor, sigmas: Tensor,
                                      packed_info: Optional[Tensor] = None, ray_indices: Optional[Tensor] = None,
                                      n_rays: Optional[int] = None, prefix_trans: Optional[Tensor] = None
                                      ) -> Tuple[Tensor, Tensor]:
    """field_rendering.py:209-264 -> (trans, alphas)."""
    w, T, a = _weights(1, sigmas, t_starts, t_ends, packed_info, ray_indices, n_rays, prefix_trans)
    return T, a


def render_weight_from_alpha(alphas: Tensor, packed_info: Optional[Tensor] = None,
                             ray_indices: Optional[Tensor] = None, n_rays: Optional[int] = None,
                             prefix_trans: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """field_rendering.py:267-309 -> (weights, trans)."""
    return _weights(0, alphas, None, None, packed_info, ray_indices, n_rays, prefix_trans)


def render_weight_from_density(t_starts: Tensor, t_ends: Tensor, sigmas: Tensor, packed_info: Optional[Tensor] = None,
                               ray_indices: Optional[Tensor] = None, n_rays: Optional[int] = None,
                               prefix_trans: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor]:
    """field_rendering.py:312-362 -> (weights, trans, alphas)."""
    return _weights(1, sigmas, t_starts, t_ends, packed_info, ray_indices, n_rays, prefix_trans)


@torch.no_grad()
def render_visibility_from_alpha(alphas: Tensor, packed_info: Optional[Tensor] = None,
                                 ray_indices: Optional[Tensor] = None, n_rays: Optional[int] = None,
                                 early_stop_eps: float = 1e-4, alpha_thre: float = 0.0,
                                 prefix_trans: Optional[Tensor] = None) -> Tensor:
    """field_rendering.py:365-418."""
    trans = render_transmittance_from_alpha(alphas, packed_info, ray_indices, n_rays, prefix_trans)
    vis = trans >= early_stop_eps
    if alpha_thre > 0:
        vis = vis & (alphas >= alpha_thre)
    return vis


@torch.no_grad()
def render_visibility_from_density(t_starts: Tensor, t_ends: Tensor, sigmas: Tensor,
                                   packed_info: Optional[Tensor] = None, ray_indices: Optional[Tensor] = None,
                                   n_rays: Optional[int] = None, early_stop_eps: float = 1e-4, alpha_thre: float = 0.0,
                                   prefix_trans: Optional[Tensor] = None) -> Tensor:
    """field_rendering.py:421-480."""
    trans, alphas = render_transmittance_from_density(t_starts, t_ends, sigmas, packed_info, ray_indices, n_rays,
                                                      prefix_trans)
    vis = trans >= early_stop_eps
    if alpha_thre > 0:
        vis = vis & (alphas >= alpha_thre)
    return vis


class _Accumulate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weights, values, ray_indices, n_rays, outputs):
        lib = _lib.load()
        w = _lib.f32(weights)
        v = _lib.f32(values) if values is not None else None
        D = 1 if v is None else v.shape[-1]
        ri = _lib.i64(ray_indices)
        if outputs is None:
            out = torch.zeros((n_rays, D), dtype=torch.float32, device=w.device)
        else:
            out = outputs
            ctx.mark_dirty(out)
        _lib.check(lib.qf_accumulate_along_rays_indexed(_lib.ptr(w), _lib.ptr(v), D, _lib.ptr(ri), w.numel(),
                                                        _lib.ptr(out), _lib.stream(w.device)),
                   "qf_accumulate_along_rays_indexed")
        ctx.save_for_backward(w, v, ri)
        ctx.D = D
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        w, v, ri = ctx.saved_tensors
        gout = _lib.f32(gout)
        gw = torch.empty_like(w) if ctx.needs_input_grad[0] else None
        gv = torch.empty_like(v) if (v is not None and ctx.needs_input_grad[1]) else None
        if gw is not None or gv is not None:
            _lib.check(lib.qf_accumulate_along_rays_backward(_lib.ptr(w), _lib.ptr(v), ctx.D, _lib.ptr(ri), w.numel(),
                                                             _lib.ptr(gout), _lib.ptr(gw), _lib.ptr(gv),
                                                             _lib.stream(w.device)), "qf_accumulate_along_rays_backward")
        return gw, gv, None, None, gout if ctx.needs_input_grad[4] else None


def accumulate_along_rays(weights: Tensor, values: Optional[Tensor] = None, ray_indices: Optional[Tensor] = None,
                          n_rays: Optional[int] = None) -> Tensor:
    """field_rendering.py:483-547."""
    if values is not None:
        assert values.dim() == weights.dim() + 1
        assert weights.shape == values.shape[:-1]
    if ray_indices is not None:
        assert n_rays is not None, "n_rays must be provided"
        assert weights.dim() == 1, "weights must be flattened"
        return _Accumulate.apply(weights, values, ray_indices, n_rays, None)
    # batched: (n_rays, n_samples[, D]) -> one segment per row
    n, S = weights.shape
    ri = torch.arange(n, device=weights.device).repeat_interleave(S)
    v = None if values is None else values.reshape(n * S, -1)
    return _Accumulate.apply(weights.reshape(-1), v, ri, n, None)


def accumulate_along_rays_(weights: Tensor, values: Optional[Tensor] = None, ray_indices: Optional[Tensor] = None,
                           outputs: Optional[Tensor] = None) -> None:
    """In-place variant, field_rendering.py:550-573."""
    if values is not None:
        assert values.dim() == weights.dim() + 1
        assert weights.shape == values.shape[:-1]
    D = 1 if values is None else values.shape[-1]
    if ray_indices is not None:
        assert weights.dim() == 1, "weights must be flattened"
        assert outputs.dim() == 2 and outputs.shape[-1] == D, "outputs must be of shape (n_rays, D)"
        _Accumulate.apply(weights, values, ray_indices, outputs.shape[0], outputs)
    else:
        outputs.add_(accumulate_along_rays(weights, values))


def _query(t_starts, t_ends, ray_indices, n_rays, rgb_sigma_fn, rgb_alpha_fn):
    if ray_indices is not None:
        assert (
            t_starts.shape == t_ends.shape == ray_indices.shape
        ), "Since nerfacc 0.5.0, t_starts, t_ends and ray_indices must have the same shape (N,). "
    if rgb_sigma_fn is None and rgb_alpha_fn is None:
        raise ValueError("At least one of `rgb_sigma_fn` and `rgb_alpha_fn` should be specified.")
    if rgb_sigma_fn is not None:
        if t_starts.shape[0] != 0:
            rgbs, sigmas = rgb_sigma_fn(t_starts, t_ends, ray_indices)
        else:
            rgbs = torch.empty((0, 3), device=t_starts.device)
            sigmas = torch.empty((0,), device=t_starts.device)
        assert rgbs.shape[-1] == 3, "rgbs must have 3 channels, got {}".format(rgbs.shape)
        assert sigmas.shape == t_starts.shape, "sigmas must have shape of (N,)! Got {}".format(sigmas.shape)
        weights, trans, alphas = render_weight_from_density(t_starts, t_ends, sigmas, ray_indices=ray_indices,
                                                            n_rays=n_rays)
        extras = {"weights": weights, "alphas": alphas, "trans": trans, "sigmas": sigmas, "rgbs": rgbs}
    else:
        if t_starts.shape[0] != 0:
            rgbs, alphas = rgb_alpha_fn(t_starts, t_ends, ray_indices)
        else:
            rgbs = torch.empty((0, 3), device=t_starts.device)
            alphas = torch.empty((0,), device=t_starts.device)
        assert rgbs.shape[-1] == 3, "rgbs must have 3 channels, got {}".format(rgbs.shape)
        assert alphas.shape == t_starts.shape, "alphas must have shape of (N,)! Got {}".format(alphas.shape)
        weights, trans = render_weight_from_alpha(alphas, ray_indices=ray_indices, n_rays=n_rays)
        extras = {"weights": weights, "trans": trans, "rgbs": rgbs, "alphas": alphas}
    colors = accumulate_along_rays(weights, values=rgbs, ray_indices=ray_indices, n_rays=n_rays)
    opacities = accumulate_along_rays(weights, values=None, ray_indices=ray_indices, n_rays=n_rays)
    depths = accumulate_along_rays(weights, values=(t_starts + t_ends)[..., None] / 2.0, ray_indices=ray_indices,
                                   n_rays=n_rays)
    depths = depths / opacities.clamp_min(torch.finfo(rgbs.dtype).eps)
    return colors, opacities, depths, extras


def rendering(t_starts: Tensor, t_ends: Tensor, ray_indices: Optional[Tensor] = None, n_rays: Optional[int] = None,
              rgb_sigma_fn: Optional[Callable] = None, rgb_alpha_fn: Optional[Callable] = None,
              render_bkgd: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor, Dict]:
    """field_rendering.py:14-158."""
    colors, opacities, depths, extras = _query(t_starts, t_ends, ray_indices, n_rays, rgb_sigma_fn, rgb_alpha_fn)
    if render_bkgd is not None:
        colors = colors + render_bkgd * (1.0 - opacities)
    return colors, opacities, depths, extras


def rendering_field(t_starts: Tensor, t_ends: Tensor, ray_indices: Optional[Tensor] = None, n_rays: Optional[int] = None,
                    rgb_sigma_fn: Optional[Callable] = None, rgb_alpha_fn: Optional[Callable] = None,
                    render_bkgd: Optional[Tensor] = None):
    """field_rendering.py:575-733: `rendering` plus the far-to-near weights (quirk Q3 reproduced: the reverse pass
    builds pack_info from the flipped, descending ray indices)."""
    colors, opacities, depths, extras = _query(t_starts, t_ends, ray_indices, n_rays, rgb_sigma_fn, rgb_alpha_fn)
    if render_bkgd is not None:
        colors = colors + render_bkgd * (1.0 - opacities)
    weights = extras["weights"]
    sigmas = extras["sigmas"]  # NameError in the reference when only rgb_alpha_fn is given; KeyError here
    max_val = torch.max(t_starts) + torch.max(t_ends)
    ts = torch.flip(max_val - t_starts, dims=[0])
    te = torch.flip(max_val - t_ends, dims=[0])
    sg = torch.flip(sigmas, dims=[0])
    weights_rev, _, _ = render_weight_from_density(te, ts, sg, ray_indices=torch.flip(ray_indices, dims=[0]),
                                                   n_rays=n_rays)
    weights_rev = torch.flip(weights_rev, dims=[0])
    return colors, opacities, depths, weights, weights_rev
