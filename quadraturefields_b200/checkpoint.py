"""The reference's `model.pth` checkpoints.

Every training stage writes one dict with `torch.save`:
  train_ngp_nerf_sg_occ.py:363   {"estimator", "model"}                        (the NeRF stage)
  train_field.py:413-418         {"estimator", "model"}  — here "model" is the quadrature Field net's state dict
  train_finetune.py:561-569      {"estimator", "field_model", "radiance_field"}
  train_fit_sg.py:491            {"estimator", "radiance_field"}               (the SG field)
and the next stage reads `ckpt["model"]` or `ckpt["radiance_field"]` plus `ckpt["estimator"]`
(train_field.py:258-260, train_finetune.py:407-409, train_fit_sg.py:376-378, bake_texture_images_shelly.py:255-259).
The tensors inside are tinycudann's flat fp32 `params` vectors (`mlp_base.params` = [MLP matrices row-major (out,in) |
grid table], `mlp_head.params`), nerfacc's estimator buffers (`resolution`, `aabbs`, `occs`, `binaries`) and the torch
decoder of the Field / SG head; the modules of this package keep those keys and layouts, so loading is
`load_state_dict` plus the checks below (size mismatches name the constructor argument that differs).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

RADIANCE_KEYS = ("radiance_field", "model")
FIELD_KEYS = ("field_model", "model")


def _state(module_or_dict) -> Dict[str, torch.Tensor]:
    sd = module_or_dict.state_dict() if hasattr(module_or_dict, "state_dict") else dict(module_or_dict)
    return {k: v.detach().cpu() for k, v in sd.items()}


def save_checkpoint(path: str, estimator=None, radiance_field=None, field_net=None, radiance_key: str = "radiance_field",
                    field_key: str = "field_model") -> None:
    """Write the dict the reference's stages exchange.  `radiance_key` is "model" for the NeRF stage
    (train_ngp_nerf_sg_occ.py:357-362) and "radiance_field" for the finetune / SG stages (train_finetune.py:563-567);
    `field_key` is "model" for the field stage (train_field.py:413-416) and "field_model" for finetuning."""
    if radiance_key not in RADIANCE_KEYS:
        raise ValueError(f"radiance_key must be one of {RADIANCE_KEYS}")
    if field_key not in FIELD_KEYS:
        raise ValueError(f"field_key must be one of {FIELD_KEYS}")
    if radiance_field is not None and field_net is not None and radiance_key == field_key:
        raise ValueError("radiance field and field net cannot share the key 'model'")
    out = {}
    if estimator is not None:
        out["estimator"] = _state(estimator)
    if field_net is not None:
        out[field_key] = _state(field_net)
    if radiance_field is not None:
        out[radiance_key] = _state(radiance_field)
    torch.save(out, path)


def _check_sizes(name: str, module, sd: Dict[str, torch.Tensor]) -> None:
    own = module.state_dict()
    for k, v in sd.items():
        if k in own and tuple(own[k].shape) != tuple(v.shape):
            hint = ""
            if k.endswith("mlp_base.params") or k.endswith("encoding.params"):
                hint = " (different log2_hashmap_size / n_levels / resolution than the checkpoint was trained with)"
            elif k in ("occs", "binaries", "resolution", "aabbs"):
                hint = " (different grid resolution / levels than the checkpoint's estimator)"
            raise ValueError(f"{name}: '{k}' has shape {tuple(v.shape)} in the checkpoint but {tuple(own[k].shape)} in the module{hint}")


def load_checkpoint(path_or_dict, estimator=None, radiance_field=None, field_net=None, map_location="cpu",
                    radiance_key: Optional[str] = None, field_key: Optional[str] = None, strict: bool = True) -> Dict:
    """Read a reference checkpoint into the given modules (any may be None) and return the raw dict.

    `radiance_key=None` takes "radiance_field" when present, else "model" — the order in which the later stages
    supersede the NeRF stage's weights; `field_key=None` takes "field_model" when present, else "model" (the field
    stage's own checkpoint)."""
    ckpt = torch.load(path_or_dict, map_location=map_location, weights_only=True) if isinstance(path_or_dict, str) else path_or_dict
    if radiance_field is not None:
        key = radiance_key or next((k for k in RADIANCE_KEYS if k in ckpt), None)
        if key is None or key not in ckpt:
            raise KeyError(f"checkpoint has none of {RADIANCE_KEYS} (keys: {sorted(ckpt)})")
        sd = {k: (v.float() if torch.is_floating_point(v) else v) for k, v in ckpt[key].items()}
        _check_sizes(key, radiance_field, sd)
        radiance_field.load_state_dict(sd, strict=strict)
    if estimator is not None:
        if "estimator" not in ckpt:
            raise KeyError(f"checkpoint has no 'estimator' (keys: {sorted(ckpt)})")
        _check_sizes("estimator", estimator, ckpt["estimator"])
        estimator.load_state_dict(ckpt["estimator"], strict=strict)
    if field_net is not None:
        key = field_key or next((k for k in FIELD_KEYS if k in ckpt), None)
        if key is None or key not in ckpt:
            raise KeyError(f"checkpoint has none of {FIELD_KEYS} (keys: {sorted(ckpt)})")
        sd = {k: (v.float() if torch.is_floating_point(v) else v) for k, v in ckpt[key].items()}
        _check_sizes(key, field_net, sd)
        field_net.load_state_dict(sd, strict=strict)
    return ckpt
