"""Import stub (ray marcher is SURVEY §8 row f-1, out of scope)."""


def ray_aabb_intersect(*a, **k):
    raise NotImplementedError("nerfacc.grid stand-in")


def traverse_grids(*a, **k):
    raise NotImplementedError("nerfacc.grid stand-in")
