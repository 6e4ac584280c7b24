"""Ray container of the drivers (the reference keeps the same two-field tuple in `datasets/utils.py`)."""
from typing import Callable, NamedTuple

import torch


class Rays(NamedTuple):
    """A bundle of rays: `origins` and unit `viewdirs`, both (..., 3)."""
    origins: torch.Tensor
    viewdirs: torch.Tensor


def namedtuple_map(fn: Callable, tup):
    """Rebuild `tup` (any namedtuple) with `fn` applied to every non-None field."""
    fields = (None if field is None else fn(field) for field in tup)
    return tup.__class__(*fields)
