"""Import stub (absent offline).  TEST INFRASTRUCTURE ONLY."""
def measure(*a, **k): raise NotImplementedError
