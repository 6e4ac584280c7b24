"""CPU-side checks (no GPU, no compute calls): the C-ABI library loads and exports every symbol the header
declares, the ctypes signatures cover the header, and the host-side logic agrees with the oracle."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

import __graft_entry__ as entry
from oracle import quadfield_oracle as O


@pytest.fixture(scope="module")
def lib():
    entry.build()
    from quadraturefields_b200 import _lib
    return _lib


def test_library_exports_header_symbols(lib):
    L = lib.load()
    syms = lib.header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/quadfield.h but not exported"
        assert s in lib._SIGS, f"{s} has no ctypes signature"
    assert set(lib._SIGS) == set(syms)
    assert L.qf_version() == 100
    assert L.qf_last_error() is not None


def test_signature_arity_matches_header(lib):
    text = re.sub(r"/\*.*?\*/", "", open(lib.HEADER_PATH).read(), flags=re.S)
    for name, args in re.findall(r"\b(qf_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        n = 0 if args.strip() in ("", "void") else len(args.split(","))
        assert n == len(lib._SIGS[name][1]), (name, n, len(lib._SIGS[name][1]))


def test_argument_validation_without_gpu(lib):
    """Argument errors are reported through return codes + qf_last_error, before any CUDA call."""
    L = lib.load()
    h = ctypes.c_void_p()
    assert L.qf_mesh_create(None, 0, None, 0, None, ctypes.byref(h)) == 1
    assert b"empty mesh" in L.qf_last_error()
    assert L.qf_trace_firstk(None, None, None, 10, 8, None, None, None, None, None, 0, None) == 1
    assert L.qf_render_weights(7, None, None, None, None, 1, 1, None, None, None, None, None) == 1
    assert b"mode=7" in L.qf_last_error()
    # entry points added in round 1h: empty inputs are a no-op, bad arguments are reported before any CUDA call
    assert L.qf_generate_rays_indexed(None, 3, None, None, None, 0, 10.0, 5.0, 5.0, 1, None, None, None) == 0
    assert L.qf_generate_rays_indexed(None, 3, None, None, None, 4, 10.0, 5.0, 5.0, 1, None, None, None) == 1
    assert b"qf_generate_rays_indexed: NULL" in L.qf_last_error()
    assert L.qf_generate_rays_indexed(None, 0, None, None, None, 4, 10.0, 5.0, 5.0, 1, None, None, None) == 1
    assert b"views=0" in L.qf_last_error()
    assert L.qf_sg_features_to_rgb_backward(None, 24, 3, None, 0, None, None, 24, None) == 0
    assert L.qf_sg_features_to_rgb_backward(None, 10, 3, None, 5, None, None, 24, None) == 1
    assert b"stride=10" in L.qf_last_error()
    assert L.qf_sg_features_to_rgb_backward(None, 24, 3, None, 5, None, None, 24, None) == 1
    assert b"NULL argument" in L.qf_last_error()
    assert L.qf_ngp_backward_features(None, None, 0, None, None, None, None, None, None, 0, None) == 0
    assert L.qf_ngp_backward_features(None, None, 5, None, None, None, None, None, None, 0, None) == 1
    assert b"qf_ngp_backward_features: NULL" in L.qf_last_error()


def test_argument_validation_of_the_frame_and_peer_entry_points(lib):
    """r2i entry points (ray-sharded frames into whole-frame / peer buffers, uint8 eval images): bad arguments are refused
    with a message before anything touches the device."""
    import ctypes as C
    L = lib.load()
    assert L.qf_render_mesh_ngp_to_frame(None, None, None, None, 8, 8, 0.005, 0, None, 4, 2, 0, 8, None, None, None, None, None, 0, None) == 1
    assert b"qf_render_mesh_ngp_to_frame: NULL field" in L.qf_last_error()
    assert L.qf_render_mesh_baked_to_frame(None, None, None, None, None, 8, 8, 0.005, 0, None, 4, 2, 0, 8, None, None, None, None, None, 0, None) == 1
    assert b"qf_render_mesh_baked_to_frame: NULL texture" in L.qf_last_error()
    assert L.qf_frame_to_u8(None, None, 0, None, None, None, None) == 0            # empty frame: nothing to do
    assert L.qf_frame_to_u8(None, None, 16, None, None, None, None) == 1
    assert b"qf_frame_to_u8: NULL argument" in L.qf_last_error()
    buf = C.create_string_buffer(64)
    assert L.qf_peer_export(None, buf) == 1 and b"qf_peer_export: NULL" in L.qf_last_error()
    out = C.c_void_p()
    assert L.qf_peer_open(None, C.byref(out)) == 1 and b"qf_peer_open: NULL" in L.qf_last_error()
    assert L.qf_peer_alloc(0, C.byref(out)) == 1 and b"qf_peer_alloc" in L.qf_last_error()
    assert L.qf_peer_close(None) == 0 and L.qf_peer_free(None) == 0                # NULL handles are no-ops
    assert L.qf_band_rows(1080, 4, 8, 3) == 136 and L.qf_band_rows(1080, 4, 8, 6) == 132 and L.qf_band_rows(1080, 40, 8, 0) == 160


def test_ops_fail_loudly_on_cpu_tensors(lib):
    from quadraturefields_b200 import field_rendering as FR
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        FR.render_weight_from_alpha(torch.rand(4), ray_indices=torch.zeros(4, dtype=torch.long), n_rays=1)


def test_missing_library_raises(lib, monkeypatch):
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", "/nonexistent/libquadfield.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        lib.load()


def test_grid_desc_matches_oracle(lib):
    from quadraturefields_b200.radiance_fields import grid
    for log2_T in (12, 14, 19, 21):
        d = grid.make_grid_desc([-1.5] * 3 + [1.5] * 3, log2_hashmap_size=log2_T)
        m = O.make_grid_meta(log2_hashmap_size=log2_T)
        assert d.n_levels == m.n_levels == 16
        assert np.array_equal(np.array(d.scale[:16], dtype=np.float32), m.scale)
        assert list(d.resolution[:16]) == m.resolution.tolist()
        assert list(d.offset[:16]) == m.offset[:-1].tolist()
        assert list(d.size[:16]) == m.size.tolist()
        assert [bool(x) for x in d.hashed[:16]] == m.hashed.tolist()
        assert grid.n_entries(d) == m.n_entries


def test_scene_geometry_matches_oracle_generators():
    from quadraturefields_b200 import scene
    v, f = scene.icosphere(3)
    vo, fo = O.icosphere(3)
    # same surface: identical vertex set (order may differ) and face count
    assert v.shape == vo.shape and f.shape == fo.shape == (20 * 4 ** 3, 3)
    key = lambda a: np.sort(np.round(a, 9).view([("", a.dtype)] * 3).ravel())
    assert np.array_equal(key(v), key(vo))
    assert np.allclose(np.linalg.norm(v, axis=1), 1.0)
    c = scene.look_at_c2w((1.0, 2.0, 3.0))
    assert np.array_equal(c, O.look_at_c2w((1.0, 2.0, 3.0)))
    assert scene.shell_mesh([0.4, 0.6, 0.8, 1.0], 4)[1].shape[0] == 20480


def test_ray_intrinsics_match_oracle():
    from quadraturefields_b200.datasets import ray_gen
    assert ray_gen.intrinsics(800, 800, 0.6911112070083618) == tuple(
        float(x) if i < 3 else x for i, x in enumerate(O.pinhole_intrinsics(800, 800, 0.6911112070083618)))
