"""quadraturefields_b200 — B200-native (sm_100a) drop-in for the Quadfield render hot path.

Module names mirror the reference's `examples/` tree so a caller can switch imports:

    field_rendering            <- examples/field_rendering.py
    radiance_fields.ngp        <- examples/radiance_fields/ngp.py
    mesh_utils                 <- examples/mesh_utils.py        (MeshIntersection, RayIntersector)
    utils                      <- examples/utils.py             (derive_properties, mesh-path render drivers)
    texture_utils              <- examples/texture_utils.py     (FeatureCompression)
    datasets.ray_gen           <- examples/datasets/nerf_synthetic.py (pinhole ray generation)

All arithmetic runs in libquadfield.so (csrc/, C ABI in include/quadfield.h).  There is no CPU path.
"""
__version__ = "0.1.0"
