"""Stand-in for trimesh==3.23.5 (absent offline).  TEST INFRASTRUCTURE ONLY.
Only what the reference's hot path touches: `Trimesh(vertices, faces)` with `face_normals`,
and `triangles.points_to_barycentric`."""
import numpy as np

from . import triangles  # noqa: F401
from . import ray  # noqa: F401


class Trimesh:
    def __init__(self, vertices=None, faces=None, process=False):
        self.vertices = np.asarray(vertices, dtype=np.float64)
        self.faces = np.asarray(faces, dtype=np.int64)

    @property
    def face_normals(self):
        t = self.vertices[self.faces]
        n = np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 1])
        ln = np.sqrt((n * n).sum(axis=1))
        out = np.zeros_like(n)
        ok = ln > 0
        out[ok] = n[ok] / ln[ok, None]
        return out


def load(*a, **k):
    raise NotImplementedError("trimesh stand-in")
