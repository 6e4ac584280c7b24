"""Drop-in for the reference's NeRF-synthetic loader (`examples/datasets/nerf_synthetic.py:157-378`).

Same constructor, attributes (`images`, `camtoworlds`, `K`, `focal`, `WIDTH`, `HEIGHT`, `training`), `__getitem__` /
`fetch_data` / `preprocess` / `update_num_rays` and the same dictionary
`{"pixels", "rays", "color_bkgd"[, "data"]}`.  What differs is where the work runs:

* rays come from the library's kernels (`qf_generate_rays` for a whole evaluation image, `qf_generate_rays_indexed`
  for random training pixels) — the loader lives on the GPU, there is no CPU path;
* with `mesh_intersect=` the hit tuple `data["data"]` is produced on the GPU by `MeshIntersection.sampling_raytrace`
  (the reference runs Embree on the CPU inside DataLoader workers and converts numpy arrays, :243-262).  The tuple keeps
  the reference's six entries (xyzs, dirs, index_ray, ts, index_tri, origins); the ray offsets ride along as the
  `offsets` attribute of `data["data"]`.

Files: `transforms_{split}.json` + PNGs read with PIL (the reference uses imageio; same uint8 arrays).
`images=` / `camtoworlds=` / `focal=` build a loader from arrays already in memory (synthetic scenes, tests).
"""
from __future__ import annotations

import json
import os
from typing import Optional

import numpy as np
import torch

from .ray_gen import generate_rays, generate_rays_indexed
from .utils import Rays


def _load_renderings(root_fp: str, subject_id: str, split: str):
    """nerf_synthetic.py:68-105 -> images (n,h,w,4) uint8, camtoworlds (n,4,4) f64, focal (python float)."""
    from PIL import Image
    if not root_fp.startswith("/"):
        root_fp = os.path.join(os.getcwd(), root_fp)
    data_dir = os.path.join(root_fp, subject_id)
    with open(os.path.join(data_dir, "transforms_{}.json".format(split)), "r") as fp:
        meta = json.load(fp)
    images, camtoworlds = [], []
    for frame in meta["frames"]:
        fname = os.path.join(data_dir, frame["file_path"] + ".png")
        if not os.path.exists(fname):
            fname = os.path.join(data_dir, frame["file_path"])
        with Image.open(fname) as im:
            images.append(np.array(im))
        camtoworlds.append(frame["transform_matrix"])
    images = np.stack(images, axis=0)
    camtoworlds = np.stack(camtoworlds, axis=0)
    h, w = images.shape[1:3]
    focal = 0.5 * w / np.tan(0.5 * float(meta["camera_angle_x"]))
    return images, camtoworlds, focal


class SubjectLoader(torch.utils.data.Dataset):
    """nerf_synthetic.py:129-378."""

    SPLITS = ["train", "val", "trainval", "test"]
    WIDTH, HEIGHT = 800, 800
    NEAR, FAR = 2.0, 6.0
    OPENGL_CAMERA = True

    def __init__(self, subject_id: Optional[str] = None, root_fp: Optional[str] = None, split: str = "train",
                 color_bkgd_aug: str = "white", num_rays: Optional[int] = None, near: Optional[float] = None,
                 far: Optional[float] = None, batch_over_images: bool = True, device="cuda", mesh_intersect=None,
                 fine_tune_vertices: bool = False, add_ray_direction_noise: bool = False, upsample: int = 1,
                 images=None, camtoworlds=None, focal: Optional[float] = None):
        super().__init__()
        assert color_bkgd_aug in ["white", "black", "random"]
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("quadraturefields_b200 datasets generate rays on the GPU (no CPU path): pass a CUDA device")
        if fine_tune_vertices:
            raise NotImplementedError("fine_tune_vertices uses sampling_for_fine_tuning_mesh_ray_trace, which no shipped script enables")
        self.upsample = int(upsample)
        self.mesh_intersect = mesh_intersect
        self.fine_tune_vertices = fine_tune_vertices
        self.add_ray_direction_noise = add_ray_direction_noise
        self.split = split
        self.num_rays = num_rays
        self.near = self.NEAR if near is None else near
        self.far = self.FAR if far is None else far
        self.training = (num_rays is not None) and (split in ["train", "trainval"])
        self.color_bkgd_aug = color_bkgd_aug
        self.batch_over_images = batch_over_images
        if images is not None:
            imgs, c2w, f = np.asarray(images), np.asarray(camtoworlds), float(focal)
        elif split == "trainval":
            a, b = _load_renderings(root_fp, subject_id, "train"), _load_renderings(root_fp, subject_id, "val")
            imgs, c2w, f = np.concatenate([a[0], b[0]]), np.concatenate([a[1], b[1]]), a[2]
        elif split in ("train", "test"):
            imgs, c2w, f = _load_renderings(root_fp, subject_id, split)
        else:                                   # evaluation on the training set (nerf_synthetic.py:207-211)
            imgs, c2w, f = _load_renderings(root_fp, subject_id, "train")
        self.focal = f * self.upsample
        self.images = torch.from_numpy(np.ascontiguousarray(imgs)).to(torch.uint8)
        height, width = self.images.shape[1:3]
        self.WIDTH, self.HEIGHT = int(width * self.upsample), int(self.upsample * height)
        self.camtoworlds = torch.from_numpy(np.ascontiguousarray(c2w)).to(torch.float32)
        self.K = torch.tensor([[self.focal, 0, self.WIDTH / 2.0], [0, self.focal, self.HEIGHT / 2.0], [0, 0, 1]],
                              dtype=torch.float32)
        self.images = self.images.to(device)
        self.camtoworlds = self.camtoworlds.to(device).contiguous()
        self._k = (float(self.K[0, 0]), float(self.K[0, 2]), float(self.K[1, 2]))       # fp32 values of focal, cx, cy
        self.K = self.K.to(device)

    def __len__(self):
        return 10000000          # nerf_synthetic.py:232 (an endless stream of random batches)

    @torch.no_grad()
    def __getitem__(self, index):
        index = index % len(self.images)
        data = self.preprocess(self.fetch_data(index))
        if self.mesh_intersect is not None:
            tup = self.mesh_intersect.sampling_raytrace(data["rays"].viewdirs, data["rays"].origins)
            if tup is None:
                data["data"] = None       # the reference fails on the unpack here (quirk Q9); the render fills rgb=1, alpha=0
            else:
                xyzs, dirs, index_ray, ts, index_tri, _, origins = tup
                lst = _TupleList([xyzs, dirs, index_ray, ts, index_tri, origins])
                lst.offsets = getattr(tup, "offsets", None)
                data["data"] = lst
        return data

    def preprocess(self, data):
        """nerf_synthetic.py:264-287: alpha-composite the target colours over the batch's background colour (white at
        evaluation; white / black / one random colour per batch in training)."""
        dev = self.images.device
        if self.training and self.color_bkgd_aug == "random":
            color_bkgd = torch.rand(3, device=dev)
        else:
            level = 0.0 if (self.training and self.color_bkgd_aug == "black") else 1.0
            color_bkgd = torch.full((3,), level, dtype=torch.float32, device=dev)
        rgb, alpha = data["rgba"][..., :3], data["rgba"][..., 3:]
        out = {k: v for k, v in data.items() if k != "rgba"}          # "rays" and anything a caller attached
        out["pixels"] = rgb * alpha + color_bkgd * (1.0 - alpha)      # [n_rays, 3]
        out["color_bkgd"] = color_bkgd                                 # [3]
        return out

    def update_num_rays(self, num_rays):
        self.num_rays = num_rays

    def fetch_data(self, index, image_id=None, x=None, y=None):
        """nerf_synthetic.py:292-378.  `image_id`, `x`, `y` override the random draws of the training branch (tests replay
        the reference's draws through them)."""
        dev = self.images.device
        focal, cx, cy = self._k
        if self.training:
            n = self.num_rays
            if image_id is None:
                image_id = (torch.randint(0, len(self.images), size=(n,), device=dev) if self.batch_over_images
                            else torch.full((n,), int(index), dtype=torch.long, device=dev))
            if x is None:
                x = torch.randint(0, self.WIDTH, size=(n,), device=dev)
                y = torch.randint(0, self.HEIGHT, size=(n,), device=dev)
            image_id, x, y = image_id.to(dev), x.to(dev), y.to(dev)
            rgba = self.images[image_id, torch.floor(y / self.upsample).long(), torch.floor(x / self.upsample).long()] / 255.0
            if self.add_ray_direction_noise and not x.is_floating_point():
                x = x.float() + torch.rand_like(x.float())
                y = y.float() + torch.rand_like(y.float())
            rays = generate_rays_indexed(self.camtoworlds, image_id, x, y, focal, cx, cy, opengl=self.OPENGL_CAMERA)
            return {"rgba": rgba.view(-1, 4), "rays": rays}
        rays = generate_rays(self.camtoworlds[index], self.WIDTH, self.HEIGHT, focal, cx, cy, opengl=self.OPENGL_CAMERA, device=dev)
        # the target image stays at the stored resolution (every upsample-th pixel of the ray grid, :321-331)
        rgba = self.images[index].reshape(-1, 4) / 255.0
        return {"rgba": rgba, "rays": rays}


class _TupleList(list):
    """`data["data"]` (a list in the reference, nerf_synthetic.py:257-258) that can also carry the ray offsets."""
    offsets = None
