"""nerfacc 0.5.3 `scan.exclusive_sum` / `exclusive_prod` restated with explicit per-chunk loops."""
import torch


def _scan(inputs, packed_info, prod):
    if packed_info is None:
        ident = torch.ones_like(inputs[..., :1]) if prod else torch.zeros_like(inputs[..., :1])
        shifted = torch.cat([ident, inputs[..., :-1]], dim=-1)
        return torch.cumprod(shifted, dim=-1) if prod else torch.cumsum(shifted, dim=-1)
    assert inputs.dim() == 1 and packed_info.dim() == 2 and packed_info.shape[-1] == 2
    out = torch.zeros_like(inputs)
    vals = inputs.tolist()
    res = [0.0] * len(vals)
    for start, cnt in packed_info.tolist():
        acc = 1.0 if prod else 0.0
        for i in range(start, start + cnt):
            res[i] = acc
            acc = acc * vals[i] if prod else acc + vals[i]
    out[:] = torch.tensor(res, dtype=torch.float64).to(inputs.dtype)
    return out


def exclusive_sum(inputs, packed_info=None):
    return _scan(inputs, packed_info, False)


def exclusive_prod(inputs, packed_info=None):
    return _scan(inputs, packed_info, True)


def inclusive_sum(inputs, packed_info=None):
    return exclusive_sum(inputs, packed_info) + inputs


def inclusive_prod(inputs, packed_info=None):
    return exclusive_prod(inputs, packed_info) * inputs
