"""Stand-in for imageio.v2 (absent offline).  TEST INFRASTRUCTURE ONLY: `imread` through PIL, which decodes PNGs to the
same uint8 arrays imageio does."""
import numpy as np


def imread(path, *a, **k):
    from PIL import Image
    with Image.open(path) as im:
        return np.array(im)
