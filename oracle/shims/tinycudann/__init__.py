"""Stand-in for tinycudann (un-pinned git dependency, absent offline).  TEST INFRASTRUCTURE ONLY.

`Encoding` (SphericalHarmonics deg 4), `Network` (FullyFusedMLP) and `NetworkWithInputEncoding`
(HashGrid + FullyFusedMLP) restated on CPU through `oracle.quadfield_oracle`.  Parameters live in ONE
flat fp32 tensor `params`, MLP weights first then the grid table, each matrix (out,in) row-major —
tcnn's layout as recalled (SURVEY §5 checkpoint row).  Values are cast to fp16 each forward like tcnn;
unlike tcnn the MLP *output* is returned un-rounded in fp32.  The stand-ins are differentiable
(parameters and inputs), with straight-through gradients at the fp16 roundings.
"""
import math

import torch

from oracle import quadfield_oracle as O


def _pad16(n):
    return (n + 15) // 16 * 16


class _MLP:
    def __init__(self, n_in, n_out, cfg):
        assert cfg["otype"] == "FullyFusedMLP" and cfg["activation"] == "ReLU" and cfg["output_activation"] == "None"
        self.n_in, self.n_out = n_in, n_out
        self.width = int(cfg["n_neurons"])
        self.n_hidden = int(cfg["n_hidden_layers"])
        self.in_pad, self.out_pad = _pad16(n_in), _pad16(n_out)
        self.shapes = [(self.width, self.in_pad)] + [(self.width, self.width)] * (self.n_hidden - 1) + \
                      [(self.out_pad, self.width)]

    @property
    def n_params(self):
        return sum(o * i for o, i in self.shapes)

    def init(self, gen):
        chunks = []
        for o, i in self.shapes:
            b = math.sqrt(6.0 / (i + o))
            chunks.append(((torch.rand(o * i, generator=gen) * 2 - 1) * b))
        return torch.cat(chunks)

    def split(self, flat):
        ws, off = [], 0
        for o, i in self.shapes:
            ws.append(O._round_h(flat[off:off + o * i].view(o, i)))       # fp16 working copy, straight-through gradient
            off += o * i
        return ws

    def forward(self, x, flat):
        x = O._round_h(x)
        if self.in_pad != self.n_in:
            pad = torch.full((x.shape[0], self.in_pad - self.n_in), O.HEAD_PAD_VALUE, dtype=torch.float32)
            x = torch.cat([x, pad], dim=-1)
        return O.mlp_forward(x, self.split(flat))[:, :self.n_out]


class Network(torch.nn.Module):
    def __init__(self, n_input_dims, n_output_dims, network_config, seed=1337):
        super().__init__()
        self.n_input_dims, self.n_output_dims = n_input_dims, n_output_dims
        self._mlp = _MLP(n_input_dims, n_output_dims, network_config)
        self.params = torch.nn.Parameter(self._mlp.init(torch.Generator().manual_seed(seed)))

    def forward(self, x):
        return self._mlp.forward(x.float(), self.params)


class Encoding(torch.nn.Module):
    """Two configurations are used by the reference: the SH(4) direction encoding (ngp.py:724-737, no parameters) and
    the "Grid"/"Hash" position encoding of the quadrature Field net (field.py:158-172; trainable, half output)."""

    def __init__(self, n_input_dims, encoding_config, seed=1337, dtype=None):
        super().__init__()
        self.n_input_dims = n_input_dims
        self.dtype = dtype
        e = encoding_config
        if e["otype"] == "Grid":
            assert e["type"] == "Hash" and e["interpolation"] == "Linear" and e["n_features_per_level"] == 2 and n_input_dims == 3
            self.kind = "grid"
            self.meta = O.make_grid_meta(n_levels=e["n_levels"], base_resolution=e["base_resolution"],
                                         log2_hashmap_size=e["log2_hashmap_size"], per_level_scale=float(e["per_level_scale"]))
            g = torch.Generator().manual_seed(seed)
            self.params = torch.nn.Parameter((torch.rand(self.meta.n_entries * 2, generator=g) * 2 - 1) * 1e-4)
            self.n_output_dims = e["n_levels"] * 2
            return
        nested = e["nested"]
        assert e["otype"] == "Composite" and len(nested) == 1
        assert nested[0]["otype"] == "SphericalHarmonics" and nested[0]["degree"] == 4
        self.kind = "sh"
        self.n_output_dims = 16
        # tcnn.Module registers `params` for every module, empty when the encoding has no parameters (recalled)
        self.params = torch.nn.Parameter(torch.zeros(0))

    def forward(self, x):
        x = x.float()
        if self.kind == "grid":
            table = O._round_h(self.params.view(-1, 2))           # fp16 working copy of the fp32 master parameters
            enc = O.hashgrid_encode(x, table, self.meta)
            return enc.to(self.dtype) if self.dtype is not None else enc.half()
        return O.sh4(x * 2.0 - 1.0).half().float()


class NetworkWithInputEncoding(torch.nn.Module):
    def __init__(self, n_input_dims, n_output_dims, encoding_config, network_config, seed=1337):
        super().__init__()
        e = encoding_config
        assert e["otype"] == "HashGrid" and e["n_features_per_level"] == 2 and n_input_dims == 3
        self.meta = O.make_grid_meta(n_levels=e["n_levels"], base_resolution=e["base_resolution"],
                                     log2_hashmap_size=e["log2_hashmap_size"],
                                     per_level_scale=float(e["per_level_scale"]))
        self._mlp = _MLP(e["n_levels"] * 2, n_output_dims, network_config)
        g = torch.Generator().manual_seed(seed)
        table = (torch.rand(self.meta.n_entries * 2, generator=g) * 2 - 1) * 1e-4
        self.params = torch.nn.Parameter(torch.cat([self._mlp.init(g), table]))
        self.n_input_dims, self.n_output_dims = n_input_dims, n_output_dims

    def forward(self, x):
        p = self.params
        table = O._round_h(p[self._mlp.n_params:].view(-1, 2))
        enc = O.hashgrid_encode(x.float(), table, self.meta)
        return self._mlp.forward(enc, p[:self._mlp.n_params])
