"""Import stub: the reference's `utils.py:22` imports these names from the wheel; its local copy is field_rendering.py."""
from field_rendering import accumulate_along_rays_, render_weight_from_density, rendering  # noqa: F401
