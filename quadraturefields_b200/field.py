"""The quadrature `Field` net (reference `examples/field.py:130-259`) on the CUDA kernels of
`csrc/field_net.cu`.

Same constructor, buffers and state-dict keys as the reference (`center`, `xyz_min`, `xyz_max`, `half_size`,
`xyz_encoder.params`, `decoder_field.layers.{0,1}.{weight,bias}`, `decoder_field.lout.{weight,bias}`), so
`field_net.load_state_dict(ckpt["model"])` (train_finetune.py:407-409) keeps working.  Covered configuration: what both
reference call sites use — `back_prop=False`, 16 levels x 2 features, two hidden layers of 16 or 32 units, ELU or ReLU,
`output_dim` <= 3.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from torch import nn

from . import _lib
from .radiance_fields import grid as _grid


class FieldDesc(C.Structure):
    _fields_ = [("hidden", C.c_int32), ("out_dim", C.c_int32), ("activation", C.c_int32),
                ("xyz_min", C.c_float * 3), ("xyz_max", C.c_float * 3)]


class _GridParams(nn.Module):
    """Parameter container with tinycudann's `Encoding` state-dict layout: one flat fp32 `params` tensor."""

    def __init__(self, n: int, seed: int = 1337):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.params = nn.Parameter((torch.rand(n, generator=g) * 2 - 1) * 1e-4)   # tcnn grid init: U(-1e-4, 1e-4)


class BasicDecoder(nn.Module):
    """field.py:17-126 as a parameter container (two hidden `nn.Linear` + `lout`); evaluated by the fused kernel."""

    def __init__(self, input_dim, output_dim, activation, bias, layer=nn.Linear, num_layers=1, hidden_dim=128, skip=[],
                 bias_last=True):
        super().__init__()
        if skip:
            raise NotImplementedError("skip connections are not used by the reference Field (field.py:183)")
        self.input_dim, self.output_dim, self.activation, self.bias = input_dim, output_dim, activation, bias
        self.num_layers, self.hidden_dim, self.skip, self.bias_last = num_layers, hidden_dim, skip, bias_last
        self.layers = nn.ModuleList([layer(input_dim if i == 0 else hidden_dim, hidden_dim, bias=bias) for i in range(num_layers)])
        self.lout = layer(hidden_dim, output_dim, bias=bias_last)

    def name(self) -> str:
        return "BasicDecoder"


def _opt_ptr(t):
    return _lib.ptr(t) if t is not None else None


class _FieldFn(torch.autograd.Function):
    """(field, field_grad) = Field(x); backward differentiates both outputs w.r.t. the grid table and the MLP (the
    field_grad part is the reference's double backward, field.py:229-238 with create_graph=True)."""

    @staticmethod
    def forward(ctx, net, return_grad, x, table, w1, b1, w2, b2, w3, b3):
        lib = _lib.load()
        x = _lib.f32(x.detach().reshape(-1, 3))
        M, dev = x.shape[0], x.device
        h = net._native()
        fld = torch.empty((M, net.output_dim), dtype=torch.float32, device=dev)
        fgrad = torch.empty((M, 3), dtype=torch.float32, device=dev) if return_grad else None
        ws = [None if t is None else _lib.f32(t.detach()) for t in (w1, b1, w2, b2, w3, b3)]
        _lib.check(lib.qf_field_forward(h, C.byref(net._fdesc), *[_opt_ptr(t) for t in ws], _lib.ptr(x), M, _lib.ptr(fld),
                                        _opt_ptr(fgrad), _lib.stream(dev)), "qf_field_forward")
        ctx.net, ctx.x, ctx.ws = net, x, ws
        ctx.have_bias = (b1 is not None, b2 is not None, b3 is not None)
        return fld, fgrad

    @staticmethod
    def backward(ctx, g_field, g_fgrad):
        lib = _lib.load()
        net, x, ws = ctx.net, ctx.x, ctx.ws
        M, dev = x.shape[0], x.device
        h = net._native()
        gf = None if g_field is None else _lib.f32(g_field)
        gg = None if g_fgrad is None else _lib.f32(g_fgrad)
        # `accumulate_grad_in_place` (see NGPRadianceField): the grid scatter adds into `xyz_encoder.params.grad` directly
        pt = net.xyz_encoder.params
        inplace = getattr(net, "accumulate_grad_in_place", False) and pt.grad is not None and pt.grad.is_contiguous()
        g_table = pt.grad if inplace else torch.zeros_like(pt)
        grads = [None if t is None else torch.zeros_like(t) for t in ws]
        if M > 0 and (gf is not None or gg is not None):
            _lib.check(lib.qf_field_backward(h, C.byref(net._fdesc), *[_opt_ptr(t) for t in ws], _lib.ptr(x), M, _opt_ptr(gf),
                                             _opt_ptr(gg), _lib.ptr(g_table), _lib.ptr(grads[0]), _opt_ptr(grads[1]),
                                             _lib.ptr(grads[2]), _opt_ptr(grads[3]), _lib.ptr(grads[4]), _opt_ptr(grads[5]),
                                             _lib.stream(dev)), "qf_field_backward")
        return (None, None, None, None if inplace else g_table, *grads)


class Field(nn.Module):
    """field.py:130-259."""

    def __init__(self, scale, back_prop=0, precision=16, log2_T=19, L=16, max_res=512, output_dim=1, min_res=16,
                 hidden_size=32, num_features=2, nl="elu", bias=True, bias_last=True):
        super().__init__()
        if back_prop:
            raise NotImplementedError("back_prop=True needs the grid's double backward; both reference call sites use False")
        if L != 16 or num_features != 2 or hidden_size not in (16, 32) or not 1 <= output_dim <= 3 or nl not in ("elu", "relu"):
            raise NotImplementedError("covered: L=16, num_features=2, hidden_size 16/32, output_dim<=3, nl elu/relu")
        self.output_dim, self.scale, self.back_prop = output_dim, scale, back_prop
        self.dtype = torch.float16 if precision == 16 else torch.float32
        self.register_buffer("center", torch.zeros(1, 3))
        self.register_buffer("xyz_min", -torch.ones(1, 3) * scale)
        self.register_buffer("xyz_max", torch.ones(1, 3) * scale)
        self.register_buffer("half_size", (self.xyz_max - self.xyz_min) / 2)
        b = float(np.exp(np.log(max_res * scale / min_res) / (L - 1)))            # field.py:154
        self._desc = _grid.make_grid_desc([0, 0, 0, 1, 1, 1], L, min_res, None, log2_T, per_level_scale=b)
        self._n_entries = _grid.n_entries(self._desc)
        self.xyz_encoder = _GridParams(2 * self._n_entries)
        activation = nn.ELU() if nl == "elu" else nn.ReLU()
        self.decoder_field = BasicDecoder(input_dim=L * num_features + 3, output_dim=output_dim, activation=activation, bias=bias,
                                          num_layers=2, hidden_dim=hidden_size, skip=[], bias_last=bias_last)
        self._fdesc = FieldDesc(hidden=hidden_size, out_dim=output_dim, activation=0 if nl == "elu" else 1)
        self._handle = None
        self._handle_key = None

    # ---- native grid handle (fp16 working copy of the table, refreshed when the parameters change) ----
    def _native(self):
        p = self.xyz_encoder.params
        if not p.is_cuda:
            raise RuntimeError("Field runs on CUDA only: call .cuda() first (no CPU path)")
        box_key = (self.xyz_min.data_ptr(), self.xyz_min._version, self.xyz_max.data_ptr(), self.xyz_max._version)
        if getattr(self, "_box_key", None) != box_key:        # a device->host copy synchronises: only when the box changed
            lo, hi = self.xyz_min.detach().cpu().reshape(-1).tolist(), self.xyz_max.detach().cpu().reshape(-1).tolist()
            for c in range(3):
                self._fdesc.xyz_min[c], self._fdesc.xyz_max[c] = lo[c], hi[c]
            self._box_key = box_key
        key = (p.data_ptr(), p._version)
        if self._handle is not None and key == self._handle_key:
            return self._handle
        lib = _lib.load()
        st = _lib.stream(p.device)
        if self._handle is None or key[0] != self._handle_key[0]:
            self._free()
            h = C.c_void_p()
            _lib.check(lib.qf_grid_create(C.byref(self._desc), _lib.ptr(p.detach()), self._n_entries, st, C.byref(h)), "qf_grid_create")
            self._handle = h
        else:
            _lib.check(lib.qf_ngp_update(self._handle, _lib.ptr(p.detach()), None, None, st), "qf_ngp_update")
        self._handle_key = key
        return self._handle

    def _free(self):
        if getattr(self, "_handle", None) is not None:
            _lib.load().qf_ngp_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass

    def _apply_fn(self, x, return_grad):
        d = self.decoder_field
        return _FieldFn.apply(self, return_grad, x, self.xyz_encoder.params, d.layers[0].weight, d.layers[0].bias,
                              d.layers[1].weight, d.layers[1].bias, d.lout.weight, d.lout.bias)

    # ---- reference call surface ----
    def density(self, x):
        """field.py:178-198 (the raw MLP output, no activation)."""
        return self._apply_fn(x, False)[0]

    def field(self, x):
        return self.density(x)[:, 0:self.output_dim]

    def forward(self, x, return_grad=True):
        """field.py:203-221 -> (field (M,output_dim), field_grad (M,3) | None)."""
        fld, fgrad = self._apply_fn(x, bool(return_grad))
        return fld, (fgrad if return_grad else None)

    def compute_field_loss(self, weights, weights_rev, field_norm, view_dirs):
        """field.py:253-259."""
        view_dirs = view_dirs / torch.norm(view_dirs, dim=1, keepdim=True)
        field_loss = torch.abs(torch.maximum(weights.detach(), weights_rev.detach())
                               - torch.abs(torch.sum(field_norm * view_dirs.detach(), 1)))
        return field_loss.mean()

    def compute_abs_loss(self, field_norm):
        """field.py:261-264."""
        return torch.linalg.norm(field_norm, ord=1, dim=1).mean()
