"""trimesh 3.23.5 `triangles.points_to_barycentric` (method="cramer") restated."""
import numpy as np


def points_to_barycentric(triangles, points, method="cramer"):
    triangles = np.asanyarray(triangles, dtype=np.float64)
    points = np.asanyarray(points, dtype=np.float64)
    edge_vectors = triangles[:, 1:] - triangles[:, :1]
    w = points - triangles[:, 0].reshape((-1, 3))
    dot = lambda a, b: (a * b).sum(axis=1)
    dot00 = dot(edge_vectors[:, 0], edge_vectors[:, 0])
    dot01 = dot(edge_vectors[:, 0], edge_vectors[:, 1])
    dot02 = dot(edge_vectors[:, 0], w)
    dot11 = dot(edge_vectors[:, 1], edge_vectors[:, 1])
    dot12 = dot(edge_vectors[:, 1], w)
    inverse_denominator = 1.0 / (dot00 * dot11 - dot01 * dot01)
    barycentric = np.zeros((len(triangles), 3), dtype=np.float64)
    barycentric[:, 2] = (dot00 * dot12 - dot01 * dot02) * inverse_denominator
    barycentric[:, 1] = (dot11 * dot02 - dot01 * dot12) * inverse_denominator
    barycentric[:, 0] = 1 - barycentric[:, 1] - barycentric[:, 2]
    return barycentric
