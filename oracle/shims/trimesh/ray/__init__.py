from . import ray_pyembree  # noqa: F401
