"""Occupancy-grid estimator: the call surface of nerfacc 0.5.3's `OccGridEstimator` that the reference
uses (`estimator.sampling`, `update_every_n_steps`, `binaries`, `aabbs`, `occs`, `state_dict`; utils.py:137-148,
422-433, train_field.py:217-219, 313-318) on the marcher kernel of `csrc/occgrid.cu`.

nerfacc is an absent third-party dependency, so its semantics are restated from the published code as recalled
(PARITY UNPINNED; see `oracle/quadfield_oracle.py::occgrid_march`).  The grid maintenance (`_update`) is nerfacc's
Python, mirrored with torch tensor ops; the density evaluation it calls is the fused field kernel.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Tuple, Union

import torch
from torch import Tensor

from . import _lib
from .field_rendering import render_visibility_from_alpha, render_visibility_from_density

QF_OCC_MAX_LEVELS = 8


class OccGridDesc(C.Structure):
    _fields_ = [("levels", C.c_int32), ("resolution", C.c_int32 * 3), ("aabbs", (C.c_float * 6) * QF_OCC_MAX_LEVELS)]


def _meshgrid3d(res: Tensor) -> Tensor:
    return torch.stack(torch.meshgrid([torch.arange(int(res[0])), torch.arange(int(res[1])), torch.arange(int(res[2]))],
                                      indexing="ij"), dim=-1).long()


class OccGridEstimator(torch.nn.Module):
    """Multi-level binary occupancy grid: level l covers the region of interest scaled by 2^l about its centre."""

    DIM: int = 3

    def __init__(self, roi_aabb: Union[List[float], Tensor], resolution: Union[int, List[int], Tensor] = 128,
                 levels: int = 1, **kwargs) -> None:
        super().__init__()
        if isinstance(resolution, int):
            resolution = [resolution] * self.DIM
        resolution = torch.as_tensor(resolution, dtype=torch.int32).reshape(-1)
        roi_aabb = torch.as_tensor(roi_aabb, dtype=torch.float32).reshape(-1).cpu()
        if resolution.numel() != 3 or roi_aabb.numel() != 6 or not 1 <= levels <= QF_OCC_MAX_LEVELS:
            raise ValueError("OccGridEstimator: 3-D grids with 1..8 levels")
        centre, half = (roi_aabb[:3] + roi_aabb[3:]) / 2, (roi_aabb[3:] - roi_aabb[:3]) / 2
        aabbs = torch.stack([torch.cat([centre - half * 2 ** l, centre + half * 2 ** l]) for l in range(levels)])
        self.cells_per_lvl = int(resolution.prod().item())
        self.levels = levels
        self.register_buffer("resolution", resolution)
        self.register_buffer("aabbs", aabbs)
        self.register_buffer("occs", torch.zeros(levels * self.cells_per_lvl))
        self.register_buffer("binaries", torch.zeros([levels] + resolution.tolist(), dtype=torch.bool))
        self.register_buffer("grid_coords", _meshgrid3d(resolution).reshape(self.cells_per_lvl, self.DIM), persistent=False)
        self.register_buffer("grid_indices", torch.arange(self.cells_per_lvl), persistent=False)
        self._desc_key = None
        self._desc = OccGridDesc()

    def _native_desc(self):
        key = (self.aabbs.data_ptr(), self.aabbs._version, self.resolution.data_ptr(), self.resolution._version)
        if key != self._desc_key:          # device -> host copy only when the boxes changed
            a, r = self.aabbs.detach().cpu().tolist(), self.resolution.detach().cpu().tolist()
            self._desc.levels = self.levels
            for c in range(3):
                self._desc.resolution[c] = int(r[c])
            for l in range(self.levels):
                for c in range(6):
                    self._desc.aabbs[l][c] = float(a[l][c])
            self._desc_key = key
        return self._desc

    # ---- marching -----------------------------------------------------------------------------------------------
    @torch.no_grad()
    def march(self, rays_o: Tensor, rays_d: Tensor, near_planes: Optional[Tensor], near_plane: float, far_plane: float,
              step_size: float, cone_angle: float, max_samples: int = 0, ray_mask: Optional[Tensor] = None,
              return_termination: bool = False):
        """`traverse_grids`: -> (ray_indices (M,), t_starts (M,), t_ends (M,), offsets (N+1,)) ray-major.
        `max_samples` > 0 caps the samples per ray, `ray_mask` (N,) bool retires rays, and with `return_termination` a fifth
        value gives the plane (N,) where every ray stopped (the chunked marching of `render_image_with_occgrid_test`)."""
        lib = _lib.load()
        dev = self.binaries.device
        o, d = _lib.f32(rays_o.reshape(-1, 3), dev), _lib.f32(rays_d.reshape(-1, 3), dev)
        N = o.shape[0]
        binaries = self.binaries.contiguous().view(torch.uint8)
        nears = _lib.f32(near_planes.reshape(-1), dev) if near_planes is not None else None
        desc, st = self._native_desc(), _lib.stream(dev)
        counts = torch.empty((N,), dtype=torch.int32, device=dev)
        offsets = torch.empty((N + 1,), dtype=torch.int64, device=dev)

        mask = ray_mask.to(dev).reshape(-1).to(torch.uint8).contiguous() if ray_mask is not None else None
        term = torch.empty((N,), dtype=torch.float32, device=dev) if return_termination else None

        def launch(p, ri, ts, te):
            _lib.check(lib.qf_occgrid_march_limited(C.byref(desc), _lib.ptr(binaries), _lib.ptr(o), _lib.ptr(d), N, _lib.ptr(nears),
                                                    float(near_plane), float(far_plane), float(step_size), float(cone_angle), p,
                                                    _lib.ptr(counts), _lib.ptr(offsets), _lib.ptr(ri), _lib.ptr(ts), _lib.ptr(te),
                                                    int(max_samples), _lib.ptr(mask), _lib.ptr(term), st),
                       "qf_occgrid_march_limited")
        launch(0, None, None, None)
        ws = _lib.workspace(dev, lib.qf_scan_workspace_bytes(N), "scan")
        _lib.check(lib.qf_hits_offsets(_lib.ptr(counts), N, _lib.ptr(offsets), _lib.ptr(ws), ws.numel(), st), "qf_hits_offsets")
        total = C.c_int64()
        _lib.check(lib.qf_hits_total(_lib.ptr(offsets), N, C.byref(total), st), "qf_hits_total")
        M = total.value
        ray_indices = torch.empty((M,), dtype=torch.int64, device=dev)
        t_starts = torch.empty((M,), dtype=torch.float32, device=dev)
        t_ends = torch.empty((M,), dtype=torch.float32, device=dev)
        if M:
            launch(1, ray_indices, t_starts, t_ends)
        if return_termination:
            return ray_indices, t_starts, t_ends, offsets, term
        return ray_indices, t_starts, t_ends, offsets

    @torch.no_grad()
    def sampling(self, rays_o: Tensor, rays_d: Tensor, sigma_fn: Optional[Callable] = None, alpha_fn: Optional[Callable] = None,
                 near_plane: float = 0.0, far_plane: float = 1e10, t_min: Optional[Tensor] = None, t_max: Optional[Tensor] = None,
                 render_step_size: float = 1e-3, early_stop_eps: float = 1e-4, alpha_thre: float = 0.0, stratified: bool = False,
                 cone_angle: float = 0.0) -> Tuple[Tensor, Tensor, Tensor]:
        """nerfacc `OccGridEstimator.sampling` -> (ray_indices, t_starts, t_ends): march through the occupied cells, then
        drop the samples that are invisible (transmittance below `early_stop_eps`) or transparent (alpha below `alpha_thre`)."""
        if t_max is not None:
            raise NotImplementedError("per-ray t_max is not used by any reference call site")
        dev = self.binaries.device
        n = rays_o.reshape(-1, 3).shape[0]
        nears = None
        if t_min is not None or stratified:
            nears = torch.full((n,), float(near_plane), dtype=torch.float32, device=dev)
            if t_min is not None:
                nears = torch.clamp(nears, min=t_min.to(dev).reshape(-1))
            if stratified:
                nears = nears + torch.rand_like(nears) * render_step_size
        ray_indices, t_starts, t_ends, offsets = self.march(rays_o, rays_d, nears, near_plane, far_plane, render_step_size, cone_angle)
        if (alpha_thre > 0.0 or early_stop_eps > 0.0) and (sigma_fn is not None or alpha_fn is not None):
            alpha_thre = min(alpha_thre, self.occs.mean().item())
            packed_info = torch.stack([offsets[:-1], offsets[1:] - offsets[:-1]], dim=-1).contiguous()
            if sigma_fn is not None:
                sigmas = sigma_fn(t_starts, t_ends, ray_indices) if t_starts.shape[0] != 0 else torch.empty((0,), device=dev)
                assert sigmas.shape == t_starts.shape, f"sigmas must have shape of (N,)! Got {sigmas.shape}"
                masks = render_visibility_from_density(t_starts=t_starts, t_ends=t_ends, sigmas=sigmas, packed_info=packed_info,
                                                       early_stop_eps=early_stop_eps, alpha_thre=alpha_thre)
            else:
                alphas = alpha_fn(t_starts, t_ends, ray_indices) if t_starts.shape[0] != 0 else torch.empty((0,), device=dev)
                assert alphas.shape == t_starts.shape, f"alphas must have shape of (N,)! Got {alphas.shape}"
                masks = render_visibility_from_alpha(alphas=alphas, packed_info=packed_info, early_stop_eps=early_stop_eps,
                                                     alpha_thre=alpha_thre)
            ray_indices, t_starts, t_ends = ray_indices[masks], t_starts[masks], t_ends[masks]
        return ray_indices, t_starts, t_ends

    # ---- grid maintenance (nerfacc's Python, mirrored) -----------------------------------------------------------------
    @torch.no_grad()
    def update_every_n_steps(self, step: int, occ_eval_fn: Callable, occ_thre: float = 1e-2, ema_decay: float = 0.95,
                             warmup_steps: int = 256, n: int = 16) -> None:
        if not self.training:
            raise RuntimeError("You should only call this function only during training. Please call "
                               "_update() directly if you want to update the field during inference.")
        if step % n == 0 and self.training:
            self._update(step=step, occ_eval_fn=occ_eval_fn, occ_thre=occ_thre, ema_decay=ema_decay, warmup_steps=warmup_steps)

    @torch.no_grad()
    def _get_all_cells(self) -> List[Tensor]:
        return [self.grid_indices] * self.levels

    @torch.no_grad()
    def _sample_uniform_and_occupied_cells(self, n: int) -> List[Tensor]:
        lvl_indices = []
        for lvl in range(self.levels):
            uniform_indices = torch.randint(self.cells_per_lvl, (n,), device=self.binaries.device)
            occupied_indices = torch.nonzero(self.binaries[lvl].flatten())[:, 0]
            if n < len(occupied_indices):
                selector = torch.randint(len(occupied_indices), (n,), device=self.binaries.device)
                occupied_indices = occupied_indices[selector]
            lvl_indices.append(torch.cat([uniform_indices, occupied_indices], dim=0))
        return lvl_indices

    @torch.no_grad()
    def _update(self, step: int, occ_eval_fn: Callable, occ_thre: float = 0.01, ema_decay: float = 0.95,
                warmup_steps: int = 256) -> None:
        lvl_indices = self._get_all_cells() if step < warmup_steps else self._sample_uniform_and_occupied_cells(self.cells_per_lvl // 4)
        for lvl, indices in enumerate(lvl_indices):
            grid_coords = self.grid_coords[indices]
            x = (grid_coords + torch.rand_like(grid_coords, dtype=torch.float32)) / self.resolution
            x = self.aabbs[lvl, :3] + x * (self.aabbs[lvl, 3:] - self.aabbs[lvl, :3])       # voxel [0,1]^3 -> world
            occ = occ_eval_fn(x).squeeze(-1)
            cell_ids = lvl * self.cells_per_lvl + indices
            self.occs[cell_ids] = torch.maximum(self.occs[cell_ids] * ema_decay, occ)
        thre = torch.clamp(self.occs[self.occs >= 0].mean(), max=occ_thre)
        self.binaries = (self.occs > thre).view(self.binaries.shape)
