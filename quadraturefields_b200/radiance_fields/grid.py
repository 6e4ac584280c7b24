"""Level table of the multiresolution hash grid (tinycudann GridEncoding semantics, ngp.py:689-727)."""
from __future__ import annotations

import numpy as np

from .. import _lib


def make_grid_desc(aabb, n_levels: int = 16, base_resolution: int = 16, max_resolution: int = 4096,
                   log2_hashmap_size: int = 19, per_level_scale=None) -> _lib.GridDesc:
    """scale_l = exp2f(l*log2(b))*base - 1, res_l = ceil(scale_l)+1, size_l = min(align8(res_l^3), 2^log2_T);
    b = exp((ln max_res - ln base_res)/(L-1)) as at ngp.py:689-691; a level is hashed when res^3 > size."""
    if n_levels > _lib.QF_MAX_LEVELS:
        raise ValueError(f"n_levels={n_levels} > {_lib.QF_MAX_LEVELS}")
    if per_level_scale is None:
        per_level_scale = float(np.exp((np.log(max_resolution) - np.log(base_resolution)) / (n_levels - 1)))
    f32 = np.float32
    log2_pls = f32(np.log2(f32(per_level_scale)))
    d = _lib.GridDesc()
    d.n_levels = n_levels
    T = 1 << log2_hashmap_size
    off = 0
    for l in range(n_levels):
        s = f32(f32(np.exp2(f32(f32(l) * log2_pls))) * f32(base_resolution)) - f32(1.0)
        r = int(np.ceil(s)) + 1
        dense = r ** 3
        n = min((dense + 7) // 8 * 8, T)
        d.scale[l], d.resolution[l], d.offset[l], d.size[l], d.hashed[l] = float(s), r, off, n, int(dense > n)
        off += n
    if off >= 2 ** 32:
        raise ValueError("hash grid too large for 32-bit entry offsets")
    for i in range(6):
        d.aabb[i] = float(aabb[i])
    return d


def n_entries(desc: _lib.GridDesc) -> int:
    l = desc.n_levels - 1
    return int(desc.offset[l]) + int(desc.size[l])
