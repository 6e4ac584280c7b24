"""kaolin 0.14.0 `kaolin.render.spc` pack primitives restated with explicit per-pack loops."""
import torch


def mark_pack_boundaries(pack_ids):
    b = torch.ones_like(pack_ids, dtype=torch.bool)
    if pack_ids.numel() > 1:
        b[1:] = pack_ids[1:] != pack_ids[:-1]
    return b


def _packs(boundaries):
    starts = torch.nonzero(boundaries).flatten().tolist()
    ends = starts[1:] + [boundaries.shape[0]]
    return list(zip(starts, ends))


def cumsum(feats, boundaries, exclusive=False, reverse=False):
    out = torch.zeros_like(feats)
    for s, e in _packs(boundaries):
        seg = feats[s:e].double()
        inc = torch.cumsum(seg, 0)
        out[s:e] = ((inc - seg) if exclusive else inc).to(feats.dtype)
    return out


def sum_reduce(feats, boundaries):
    packs = _packs(boundaries)
    out = torch.zeros((len(packs),) + tuple(feats.shape[1:]), dtype=feats.dtype)
    for i, (s, e) in enumerate(packs):
        out[i] = feats[s:e].double().sum(0).to(feats.dtype)
    return out


def exponential_integration(feats, tau, boundaries, exclusive=True):
    alpha = 1.0 - torch.exp(-tau.contiguous())
    transmittance = torch.exp(-1.0 * cumsum(tau.contiguous(), boundaries, exclusive=exclusive))
    transmittance = transmittance * alpha
    feats_out = sum_reduce(transmittance * feats, boundaries)
    return feats_out, transmittance
