"""nerfacc 0.5.3 `pack.pack_info` restated: counts via index_add_, starts via cumsum (assumes ascending ids)."""
import torch


def pack_info(ray_indices, n_rays=None):
    if n_rays is None:
        n_rays = int(ray_indices.max()) + 1 if ray_indices.numel() else 0
    chunk_cnts = torch.zeros((n_rays,), device=ray_indices.device, dtype=torch.long)
    chunk_cnts.index_add_(0, ray_indices.long(), torch.ones_like(ray_indices, dtype=torch.long))
    chunk_starts = chunk_cnts.cumsum(dim=0, dtype=torch.long) - chunk_cnts
    return torch.stack([chunk_starts, chunk_cnts], dim=-1)
