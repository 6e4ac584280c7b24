import numpy as np
import torch

from oracle import quadfield_oracle as O


def oracle_params(sc) -> O.NGPParams:
    """Oracle view of a quadraturefields_b200.scene.Scene's field parameters (same arrays)."""
    return O.NGPParams(torch.tensor(sc.aabb, dtype=torch.float32), O.make_grid_meta(log2_hashmap_size=sc.log2_T),
                       sc.table, sc.base_w, sc.head_w)


def oracle_texture(sc) -> O.TextureSet:
    p = sc.planes
    return O.TextureSet(p["alpha"], p["diffuse"], list(p["sg_colors"]), list(p["lambdas"]), sc.compressor.compression_type,
                        sc.compressor.lambda_thres)


def maxabs(a, b):
    a = a.detach().cpu().double() if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a)).double()
    b = b.detach().cpu().double() if isinstance(b, torch.Tensor) else torch.as_tensor(np.asarray(b)).double()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max()) if a.numel() else 0.0


def psnr_delta_db(img, ref, target):
    """|PSNR(img,target) - PSNR(ref,target)| in dB."""
    mse = lambda x, y: float(((x.double() - y.double()) ** 2).mean())
    p = lambda x: -10.0 * np.log10(max(mse(x, target), 1e-20))
    return abs(p(img) - p(ref))
