"""Import stub (absent offline).  TEST INFRASTRUCTURE ONLY."""
from . import v2  # noqa
