"""ctypes binding of libquadfield.so (include/quadfield.h).  There is no CPU fallback: every op in this
package goes through this library and raises if it is missing."""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libquadfield.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "quadfield.h")

QF_MAX_HITS = 32
QF_MAX_LEVELS = 16
QF_MAX_LOBES = 8
BG_MODES = {"white": 0, "black": 1, "random": 2}


class GridDesc(C.Structure):
    _fields_ = [("n_levels", C.c_int32), ("scale", C.c_float * QF_MAX_LEVELS),
                ("resolution", C.c_uint32 * QF_MAX_LEVELS), ("offset", C.c_uint32 * QF_MAX_LEVELS),
                ("size", C.c_uint32 * QF_MAX_LEVELS), ("hashed", C.c_uint32 * QF_MAX_LEVELS), ("aabb", C.c_float * 6)]


_P, _I, _L, _F, _SZ = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t
_SIGS = {
    "qf_last_error": (C.c_char_p, []),
    "qf_version": (_I, []),
    "qf_mesh_create": (_I, [_P, _L, _P, _L, _P, C.POINTER(_P)]),
    "qf_mesh_update_vertices": (_I, [_P, _P, _P]),
    "qf_mesh_destroy": (None, [_P]),
    "qf_mesh_info": (_I, [_P, C.POINTER(_L), C.POINTER(_F)]),
    "qf_mesh_set_restart_eps": (_I, [_P, _F]),
    "qf_trace_workspace_bytes": (_SZ, [_L]),
    "qf_trace_firstk": (_I, [_P, _P, _P, _L, _I, _P, _P, _P, _P, _P, _SZ, _P]),
    "qf_scan_workspace_bytes": (_SZ, [_L]),
    "qf_hits_offsets": (_I, [_P, _L, _P, _P, _SZ, _P]),
    "qf_hits_total": (_I, [_P, _L, C.POINTER(_L), _P]),
    "qf_hits_pack": (_I, [_P, _P, _P, _L, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "qf_hits_resort": (_I, [_P, _P, _L, _P, _P, _P]),
    "qf_ngp_create": (_I, [C.POINTER(GridDesc), _P, _L, _P, _P, _P, C.POINTER(_P)]),
    "qf_ngp_update": (_I, [_P, _P, _P, _P, _P]),
    "qf_ngp_destroy": (None, [_P]),
    "qf_hashgrid_forward": (_I, [_P, _P, _L, _P, _P]),
    "qf_hashgrid_backward": (_I, [_P, _P, _P, _L, _P, _P]),
    "qf_ngp_query_density": (_I, [_P, _P, _L, _P, _P, _P]),
    "qf_ngp_forward": (_I, [_P, _P, _P, _P, _L, _P, _P, _P]),
    "qf_ngp_backward_workspace_bytes": (_SZ, [_L]),
    "qf_ngp_backward": (_I, [_P, _P, _P, _P, _L, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "qf_ngp_backward_inputs": (_I, [_P, _P, _P, _P, _L, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "qf_texture_create": (_I, [_I, _I, _P, _P, C.POINTER(_P), C.POINTER(_P), _I, _F, _P, C.POINTER(_P)]),
    "qf_texture_destroy": (None, [_P]),
    "qf_texture_decode": (_I, [_P, _P, _L, _P, _P]),
    "qf_texture_compress": (_I, [_P, _L, _I, _I, _F, _P, _I, _P, _P, C.POINTER(_P), C.POINTER(_P), _P]),
    "qf_sg_features_to_rgb": (_I, [_P, _L, _I, _P, _L, _P, _P]),
    "qf_sg_features_to_rgb_backward": (_I, [_P, _L, _I, _P, _L, _P, _P, _L, _P]),
    "qf_ngp_backward_features": (_I, [_P, _P, _L, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "qf_grid_create": (_I, [_P, _P, _L, _P, C.POINTER(_P)]),
    "qf_field_forward": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _P, _P, _P]),
    "qf_field_backward": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "qf_occgrid_march": (_I, [_P, _P, _P, _P, _L, _P, _F, _F, _F, _F, _I, _P, _P, _P, _P, _P, _P]),
    "qf_occgrid_march_limited": (_I, [_P, _P, _P, _P, _L, _P, _F, _F, _F, _F, _I, _P, _P, _P, _P, _P, _I, _P, _P, _P]),
    "qf_triangle_accumulate": (_I, [_P, _P, _P, _L, _L, _P, _P, _P]),
    "qf_vertex_displace_workspace_bytes": (C.c_size_t, [_L]),
    "qf_vertex_displace": (_I, [_P, _P, _P, _L, _L, _F, _P, _P, C.c_size_t, _P]),
    "qf_triangle_weight_max": (_I, [_P, _L, _P, _L, _L, _P, _P]),
    "qf_hit_texels": (_I, [_P, _P, _P, _L, _P, _I, _P, _P]),
    "qf_derive_properties": (_I, [_P, _P, _P, _F, _P, _L, _I, _P, _P, _P, _P, _P, _P]),
    "qf_derive_properties_backward": (_I, [_P, _P, _P, _F, _P, _L, _L, _I, _P, _P, _P, _P, _P, _P, _P]),
    "qf_render_weights": (_I, [_I, _P, _P, _P, _P, _L, _L, _P, _P, _P, _P, _P]),
    "qf_render_weights_backward": (_I, [_I, _P, _P, _P, _P, _L, _L, _P, _P, _P, _P, _P]),
    "qf_accumulate_along_rays": (_I, [_P, _P, _I, _P, _L, _P, _I, _P]),
    "qf_accumulate_along_rays_indexed": (_I, [_P, _P, _I, _P, _L, _P, _P]),
    "qf_accumulate_along_rays_backward": (_I, [_P, _P, _I, _P, _L, _P, _P, _P, _P]),
    "qf_pack_info_workspace_bytes": (_SZ, [_L]),
    "qf_pack_info": (_I, [_P, _L, _L, _P, _P, _SZ, _P]),
    "qf_render_workspace_bytes": (_SZ, [_L, _I]),
    "qf_render_mesh_ngp": (_I, [_P, _P, _P, _P, _L, _I, _I, _F, _I, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "qf_render_mesh_baked": (_I, [_P, _P, _P, _P, _P, _L, _I, _I, _F, _I, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "qf_render_mesh_ngp_to_frame": (_I, [_P, _P, _P, _P, _L, _I, _F, _I, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _SZ, _P]),
    "qf_render_mesh_baked_to_frame": (_I, [_P, _P, _P, _P, _P, _L, _I, _F, _I, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _SZ, _P]),
    "qf_frame_to_u8": (_I, [_P, _P, _L, _P, _P, _P, _P]),
    "qf_peer_alloc": (_I, [_SZ, C.POINTER(_P)]),
    "qf_peer_free": (_I, [_P]),
    "qf_peer_export": (_I, [_P, C.c_char_p]),
    "qf_peer_open": (_I, [C.c_char_p, C.POINTER(_P)]),
    "qf_peer_close": (_I, [_P]),
    "qf_profile_enable": (_I, [_I]),
    "qf_profile_read": (_I, [C.POINTER(C.c_double), C.POINTER(_L)]),
    "qf_generate_rays": (_I, [C.POINTER(_F), _I, _I, _F, _F, _F, _I, _P, _P, _P]),
    "qf_band_rows": (_L, [_I, _I, _I, _I]),
    "qf_generate_rays_banded": (_I, [C.POINTER(_F), _I, _I, _F, _F, _F, _I, _I, _I, _I, _P, _P, _P]),
    "qf_generate_rays_indexed": (_I, [_P, _L, _P, _P, _P, _L, _F, _F, _F, _I, _P, _P, _P]),
}

_lib: Optional[C.CDLL] = None


def header_symbols() -> list:
    """Every function include/quadfield.h declares."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qf_[a-z0-9_]+)\s*\(", text)))


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python __graft_entry__.py` (build()) first. "
                           "quadraturefields_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    if missing:
        raise RuntimeError(f"libquadfield.so does not export {missing}")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().qf_last_error().decode(errors="replace")
        raise RuntimeError(f"libquadfield {what} failed (code {rc}): {msg}")


def ptr(t: Optional[torch.Tensor]):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("quadraturefields_b200 ops take CUDA tensors only (no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError("internal error: non-contiguous tensor passed to the C ABI")
    return C.c_void_p(t.data_ptr())


def _device_index(device) -> int:
    if isinstance(device, torch.device):
        if device.index is not None:
            return device.index
    elif isinstance(device, int):
        return device
    elif device is not None:
        d = torch.device(device)
        if d.index is not None:
            return d.index
    return torch.cuda.current_device()


def raw_stream(device=None) -> int:
    """cudaStream_t of torch's current stream on `device` as an integer (the C-level lookup: `torch.cuda.current_stream`
    costs ~10 us of Python per call, and a training step makes ~20 of them)."""
    return torch._C._cuda_getCurrentRawStream(_device_index(device))


def stream(device=None):
    return C.c_void_p(raw_stream(device))


def f32(t: torch.Tensor, device=None) -> torch.Tensor:
    if device is not None and t.device != device:
        t = t.to(device, non_blocking=True)
    if t.dtype is torch.float32 and t.is_contiguous():
        return t
    return t.to(torch.float32).contiguous()


def i64(t: torch.Tensor) -> torch.Tensor:
    if t.dtype is torch.int64 and t.is_contiguous():
        return t
    return t.to(torch.int64).contiguous()


_workspaces = {}


def workspace(device, nbytes: int, tag: str = "default") -> torch.Tensor:
    """Grow-only uint8 scratch buffer per (device, tag); the C ABI never allocates caller-visible memory."""
    if (device.type if isinstance(device, torch.device) else torch.device(device).type) != "cuda":
        raise RuntimeError("quadraturefields_b200 ops take CUDA tensors only (no CPU path)")
    key = (_device_index(device), tag, raw_stream(device))   # streams must not share scratch
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf
