"""Stand-in for torch_scatter (absent offline).  TEST INFRASTRUCTURE ONLY.

Recalled semantics of torch_scatter 2.x for the three calls the reference makes (1-D index along dim 0, `out=` given):
  scatter_add(src, index, dim=0, out=out)   out[index[i]] += src[i]
  scatter_mean(src, index, dim=0, out=out)  out = (out + sum) / clamp(count, min=1)
  scatter_max(src, index, out=out)          out[index[i]] = max(out[index[i]], src[i]); returns (out, argmax)
"""
import torch


def _expand(index, src):
    index = index.reshape(-1).long()
    return index.reshape([-1] + [1] * (src.dim() - 1)).expand_as(src) if src.dim() > 1 else index


def scatter_add(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0
    if out is None:
        n = int(index.max()) + 1 if dim_size is None else dim_size
        out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype)
    out.scatter_add_(0, _expand(index, src), src)
    return out


def scatter_mean(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0
    out = scatter_add(src, index, dim, out, dim_size)
    count = torch.zeros(out.shape[0], dtype=src.dtype)
    count.scatter_add_(0, index.reshape(-1).long(), torch.ones(index.numel(), dtype=src.dtype))
    count.clamp_(min=1)
    out.div_(count.reshape([-1] + [1] * (out.dim() - 1)))
    return out


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    assert src.dim() == 1
    if out is None:
        n = int(index.max()) + 1 if dim_size is None else dim_size
        out = torch.full((n,), float("-inf"), dtype=src.dtype)
    out.scatter_reduce_(0, index.reshape(-1).long(), src, reduce="amax", include_self=True)
    return out, None
